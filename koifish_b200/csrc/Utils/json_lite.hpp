// json_lite.hpp -- a small self-contained JSON reader (objects keep insertion order) for the model / quantizer config.
// The reference uses nlohmann::json (vendored, src/Utils/json.hpp) through jKV()/jKV_arr() helpers
// (src/Utils/CLI_params.cpp); this reader covers what the hot path's config needs: objects, arrays, strings, numbers,
// booleans, null, and '#'-prefixed "comment" keys (kept as ordinary keys; callers skip them as the reference does,
// src/Tensor/GeQuant.cpp:1226).
#pragma once
#include <cctype>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace koifish {

class JSON {
   public:
    enum Kind { Null, Bool, Number, String, Array, Object };
    Kind kind = Null;
    bool b    = false;
    double num = 0;
    std::string str;
    std::vector<JSON> arr;
    std::vector<std::pair<std::string, JSON>> obj;

    bool is_object() const { return kind == Object; }
    bool is_array() const { return kind == Array; }
    bool is_string() const { return kind == String; }
    bool is_number() const { return kind == Number; }
    bool is_bool() const { return kind == Bool; }
    bool is_null() const { return kind == Null; }
    bool empty() const { return kind == Null || (kind == Object && obj.empty()) || (kind == Array && arr.empty()); }

    const JSON* find(const std::string& key) const {
        if (kind != Object) return nullptr;
        for (auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool contains(const std::string& key) const { return find(key) != nullptr; }
    const JSON& at(const std::string& key) const {
        const JSON* p = find(key);
        if (!p) throw std::runtime_error("json: missing key '" + key + "'");
        return *p;
    }
    // nested lookup: get({"model","parameter","Layer"})
    const JSON* path(std::initializer_list<const char*> keys) const {
        const JSON* cur = this;
        for (const char* k : keys) {
            if (!cur) return nullptr;
            cur = cur->find(k);
        }
        return cur;
    }
    int as_int(int dflt = 0) const { return kind == Number ? (int)num : kind == Bool ? (int)b : dflt; }
    double as_double(double dflt = 0) const { return kind == Number ? num : dflt; }
    bool as_bool(bool dflt = false) const { return kind == Bool ? b : kind == Number ? num != 0 : dflt; }
    std::string as_string(const std::string& dflt = "") const { return kind == String ? str : dflt; }

    static JSON parse(const std::string& text) {
        size_t pos = 0;
        JSON v     = parse_value(text, pos);
        skip_ws(text, pos);
        if (pos != text.size()) throw std::runtime_error("json: trailing characters at offset " + std::to_string(pos));
        return v;
    }

   private:
    static void skip_ws(const std::string& s, size_t& p) {
        while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\n' || s[p] == '\r')) p++;
    }
    static JSON parse_value(const std::string& s, size_t& p) {
        skip_ws(s, p);
        if (p >= s.size()) throw std::runtime_error("json: unexpected end");
        const char c = s[p];
        JSON v;
        if (c == '{') {
            v.kind = Object;
            p++;
            skip_ws(s, p);
            if (p < s.size() && s[p] == '}') {
                p++;
                return v;
            }
            for (;;) {
                skip_ws(s, p);
                if (p >= s.size() || s[p] != '"') throw std::runtime_error("json: expected key at offset " + std::to_string(p));
                std::string key = parse_string(s, p);
                skip_ws(s, p);
                if (p >= s.size() || s[p] != ':') throw std::runtime_error("json: expected ':' at offset " + std::to_string(p));
                p++;
                JSON val = parse_value(s, p);
                v.obj.emplace_back(std::move(key), std::move(val));
                skip_ws(s, p);
                if (p < s.size() && s[p] == ',') {
                    p++;
                    continue;
                }
                if (p < s.size() && s[p] == '}') {
                    p++;
                    return v;
                }
                throw std::runtime_error("json: expected ',' or '}' at offset " + std::to_string(p));
            }
        }
        if (c == '[') {
            v.kind = Array;
            p++;
            skip_ws(s, p);
            if (p < s.size() && s[p] == ']') {
                p++;
                return v;
            }
            for (;;) {
                v.arr.push_back(parse_value(s, p));
                skip_ws(s, p);
                if (p < s.size() && s[p] == ',') {
                    p++;
                    continue;
                }
                if (p < s.size() && s[p] == ']') {
                    p++;
                    return v;
                }
                throw std::runtime_error("json: expected ',' or ']' at offset " + std::to_string(p));
            }
        }
        if (c == '"') {
            v.kind = String;
            v.str  = parse_string(s, p);
            return v;
        }
        if (s.compare(p, 4, "true") == 0) {
            v.kind = Bool, v.b = true, p += 4;
            return v;
        }
        if (s.compare(p, 5, "false") == 0) {
            v.kind = Bool, v.b = false, p += 5;
            return v;
        }
        if (s.compare(p, 4, "null") == 0) {
            p += 4;
            return v;
        }
        if (c == '-' || c == '+' || std::isdigit((unsigned char)c)) {
            const char* b = s.c_str() + p;
            char* e       = nullptr;
            v.num         = std::strtod(b, &e);
            if (e == b) throw std::runtime_error("json: bad number at offset " + std::to_string(p));
            v.kind = Number;
            p += (size_t)(e - b);
            return v;
        }
        throw std::runtime_error(std::string("json: unexpected character '") + c + "' at offset " + std::to_string(p));
    }
    static std::string parse_string(const std::string& s, size_t& p) {
        std::string out;
        p++;  // opening quote
        while (p < s.size() && s[p] != '"') {
            char c = s[p++];
            if (c == '\\' && p < s.size()) {
                char e = s[p++];
                switch (e) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {
                        auto hex4 = [&](size_t at, unsigned* v) {
                            if (at + 4 > s.size()) return false;
                            unsigned x = 0;
                            for (int i = 0; i < 4; i++) {
                                const char h = s[at + i];
                                if (!std::isxdigit((unsigned char)h)) return false;
                                x = x * 16 + (unsigned)(std::isdigit((unsigned char)h) ? h - '0' : (std::tolower((unsigned char)h) - 'a' + 10));
                            }
                            *v = x;
                            return true;
                        };
                        unsigned cp = 0, lo = 0;
                        if (!hex4(p, &cp)) throw std::runtime_error("json: bad \\u escape at offset " + std::to_string(p));
                        p += 4;
                        // a UTF-16 surrogate pair written as two escapes is one code point
                        if (cp >= 0xD800 && cp <= 0xDBFF && p + 6 <= s.size() && s[p] == '\\' && s[p + 1] == 'u' && hex4(p + 2, &lo) && lo >= 0xDC00 && lo <= 0xDFFF) {
                            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                            p += 6;
                        }
                        if (cp < 0x80)
                            out += (char)cp;
                        else if (cp < 0x800)
                            out += (char)(0xC0 | (cp >> 6)), out += (char)(0x80 | (cp & 0x3F));
                        else if (cp < 0x10000)
                            out += (char)(0xE0 | (cp >> 12)), out += (char)(0x80 | ((cp >> 6) & 0x3F)), out += (char)(0x80 | (cp & 0x3F));
                        else
                            out += (char)(0xF0 | (cp >> 18)), out += (char)(0x80 | ((cp >> 12) & 0x3F)), out += (char)(0x80 | ((cp >> 6) & 0x3F)),
                                out += (char)(0x80 | (cp & 0x3F));
                        break;
                    }
                    default: out += e;
                }
            } else
                out += c;
        }
        if (p >= s.size()) throw std::runtime_error("json: unterminated string");
        p++;  // closing quote
        return out;
    }
};

}  // namespace koifish
