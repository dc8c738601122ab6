// KunFile.cpp -- the reference's own checkpoint container, "fish.kun" (CKP_KOIFISH): reference src/Manifold/Serialize.cpp
//   K_SafeTensors::Register / InitHeader / Save  :880-1010      GTensor::jDesc :61-100      K_SafeTensors::insertJS  src/Tensor/Safetensors.hpp:87-102
// It is a safetensors file -- u64 LE header length | JSON header | data -- whose per-tensor header entry is
//     {"dtype": <K_FLOATS name, src/g_float.hpp:127-151>, "shape": [...], "data_offsets": [b, e], "loAB": 0, "szGama": g, "szData": d}
// with e - b == szData + szGama: the payload is the tensor's device blob `data || gama` byte for byte (GTensor::SerialGamaData,
// src/Device/CUDA/huTensor.cu:413-458), i.e. packed codes followed by the bf16 [R_SCALE][C_SCALE][ZERO][STEP] array.  "__metadata__" is
// {"format": "pt", "writer": "koifish"}; one more entry, "__koifish__config__" (dtype U8, shape [n]), holds the writer's JSON config as msgpack
// (nlohmann::json::to_msgpack).  Training-state files (._koifish_state_.ckp) append the optimizer moments to every payload; they are refused.
// Host code only.
#include "KunFile.hpp"

#include <stdio.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <cmath>

namespace koifish {

static const char* kConfigKey = "__koifish__config__";

int kun_dtype_bits(const std::string& d) {
    if (d == "BF16(E8)" || d == "BF16" || d == "F16(E5)" || d == "F16" || d == "U16" || d == "I16") return 16;
    if (d == "FLOAT" || d == "F32" || d == "U32" || d == "I32") return 32;
    if (d == "F64" || d == "U64" || d == "I64") return 64;
    if (d == "F8E5M2" || d == "F8E4M3" || d == "U8" || d == "I8") return 8;
    if (d == "Q<4>") return 4;
    if (d == "Q<3>") return 3;
    if (d == "Q<2>" || d == "TERNARY") return 2;
    if (d == "BINARY" || d == "BOOL<1>") return 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------ JSON text
static void dump_string(const std::string& s, std::string* o) {
    *o += '"';
    for (unsigned char c : s) {
        switch (c) {
            case '"': *o += "\\\""; break;
            case '\\': *o += "\\\\"; break;
            case '\n': *o += "\\n"; break;
            case '\r': *o += "\\r"; break;
            case '\t': *o += "\\t"; break;
            default:
                if (c < 0x20) {
                    char b[8];
                    snprintf(b, sizeof(b), "\\u%04x", c);
                    *o += b;
                } else
                    *o += (char)c;
        }
    }
    *o += '"';
}
static void dump_into(const JSON& j, std::string* o) {
    switch (j.kind) {
        case JSON::Null: *o += "null"; break;
        case JSON::Bool: *o += j.b ? "true" : "false"; break;
        case JSON::Number: {
            char b[40];
            if (std::isfinite(j.num) && j.num == std::floor(j.num) && std::fabs(j.num) < 9007199254740992.0)
                snprintf(b, sizeof(b), "%lld", (long long)j.num);
            else
                snprintf(b, sizeof(b), "%.17g", j.num);
            *o += b;
            break;
        }
        case JSON::String: dump_string(j.str, o); break;
        case JSON::Array:
            *o += '[';
            for (size_t i = 0; i < j.arr.size(); i++) {
                if (i) *o += ',';
                dump_into(j.arr[i], o);
            }
            *o += ']';
            break;
        case JSON::Object:
            *o += '{';
            for (size_t i = 0; i < j.obj.size(); i++) {
                if (i) *o += ',';
                dump_string(j.obj[i].first, o);
                *o += ':';
                dump_into(j.obj[i].second, o);
            }
            *o += '}';
            break;
    }
}
std::string json_dump(const JSON& j) {
    std::string o;
    dump_into(j, &o);
    return o;
}

// ------------------------------------------------------------------------------------------------ msgpack (the subset JSON needs)
static void put_be(std::vector<uint8_t>* o, uint64_t v, int bytes) {
    for (int i = bytes - 1; i >= 0; i--) o->push_back((uint8_t)(v >> (8 * i)));
}
void msgpack_encode(const JSON& j, std::vector<uint8_t>* o) {
    switch (j.kind) {
        case JSON::Null: o->push_back(0xc0); break;
        case JSON::Bool: o->push_back(j.b ? 0xc3 : 0xc2); break;
        case JSON::Number: {
            const double x = j.num;
            if (std::isfinite(x) && x == std::floor(x) && std::fabs(x) < 9007199254740992.0) {
                const long long v = (long long)x;
                if (v >= 0) {
                    if (v < 128) o->push_back((uint8_t)v);
                    else if (v < 256) o->push_back(0xcc), put_be(o, (uint64_t)v, 1);
                    else if (v < 65536) o->push_back(0xcd), put_be(o, (uint64_t)v, 2);
                    else if (v < 4294967296ll) o->push_back(0xce), put_be(o, (uint64_t)v, 4);
                    else o->push_back(0xcf), put_be(o, (uint64_t)v, 8);
                } else {
                    if (v >= -32) o->push_back((uint8_t)(int8_t)v);
                    else if (v >= -128) o->push_back(0xd0), put_be(o, (uint64_t)(uint8_t)(int8_t)v, 1);
                    else if (v >= -32768) o->push_back(0xd1), put_be(o, (uint64_t)(uint16_t)(int16_t)v, 2);
                    else if (v >= -2147483648ll) o->push_back(0xd2), put_be(o, (uint64_t)(uint32_t)(int32_t)v, 4);
                    else o->push_back(0xd3), put_be(o, (uint64_t)v, 8);
                }
            } else {
                uint64_t bits;
                memcpy(&bits, &x, 8);
                o->push_back(0xcb), put_be(o, bits, 8);
            }
            break;
        }
        case JSON::String: {
            const size_t n = j.str.size();
            if (n < 32) o->push_back((uint8_t)(0xa0 | n));
            else if (n < 256) o->push_back(0xd9), put_be(o, n, 1);
            else if (n < 65536) o->push_back(0xda), put_be(o, n, 2);
            else o->push_back(0xdb), put_be(o, n, 4);
            o->insert(o->end(), j.str.begin(), j.str.end());
            break;
        }
        case JSON::Array: {
            const size_t n = j.arr.size();
            if (n < 16) o->push_back((uint8_t)(0x90 | n));
            else if (n < 65536) o->push_back(0xdc), put_be(o, n, 2);
            else o->push_back(0xdd), put_be(o, n, 4);
            for (const JSON& e : j.arr) msgpack_encode(e, o);
            break;
        }
        case JSON::Object: {
            const size_t n = j.obj.size();
            if (n < 16) o->push_back((uint8_t)(0x80 | n));
            else if (n < 65536) o->push_back(0xde), put_be(o, n, 2);
            else o->push_back(0xdf), put_be(o, n, 4);
            for (const auto& kv : j.obj) {
                JSON k;
                k.kind = JSON::String, k.str = kv.first;
                msgpack_encode(k, o);
                msgpack_encode(kv.second, o);
            }
            break;
        }
    }
}
namespace {
struct Reader {
    const uint8_t* p;
    size_t n, i = 0;
    std::string err;
    int depth = 0;
    bool need(size_t k) {
        if (i + k > n) {
            if (err.empty()) err = "msgpack: truncated";
            return false;
        }
        return true;
    }
    uint64_t be(int bytes) {
        uint64_t v = 0;
        for (int k = 0; k < bytes; k++) v = (v << 8) | p[i++];
        return v;
    }
    bool str(size_t len, std::string* s) {
        if (!need(len)) return false;
        s->assign((const char*)p + i, len);
        i += len;
        return true;
    }
    bool value(JSON* out);
    bool seq(size_t len, JSON* out) {
        out->kind = JSON::Array;
        for (size_t k = 0; k < len; k++) {
            JSON e;
            if (!value(&e)) return false;
            out->arr.push_back(std::move(e));
        }
        return true;
    }
    bool map(size_t len, JSON* out) {
        out->kind = JSON::Object;
        for (size_t k = 0; k < len; k++) {
            JSON key, v;
            if (!value(&key)) return false;
            if (key.kind != JSON::String) key.str = json_dump(key);  // JSON keys are strings
            if (!value(&v)) return false;
            out->obj.emplace_back(key.str, std::move(v));
        }
        return true;
    }
};
bool Reader::value(JSON* out) {
    if (++depth > 200) {
        err = "msgpack: nested too deep";
        return false;
    }
    struct Leave {
        int& d;
        ~Leave() { d--; }
    } leave{depth};
    if (!need(1)) return false;
    const uint8_t t = p[i++];
    auto num = [&](double v) {
        out->kind = JSON::Number, out->num = v;
        return true;
    };
    if (t < 0x80) return num(t);
    if (t >= 0xe0) return num((int8_t)t);
    if ((t & 0xf0) == 0x80) return map(t & 0x0f, out);
    if ((t & 0xf0) == 0x90) return seq(t & 0x0f, out);
    if ((t & 0xe0) == 0xa0) {
        out->kind = JSON::String;
        return str(t & 0x1f, &out->str);
    }
    switch (t) {
        case 0xc0: out->kind = JSON::Null; return true;
        case 0xc2: out->kind = JSON::Bool, out->b = false; return true;
        case 0xc3: out->kind = JSON::Bool, out->b = true; return true;
        case 0xc4: case 0xc5: case 0xc6: {  // bin 8 / 16 / 32: kept as a string of its bytes
            const int lb = t == 0xc4 ? 1 : t == 0xc5 ? 2 : 4;
            if (!need(lb)) return false;
            const size_t len = (size_t)be(lb);
            out->kind = JSON::String;
            return str(len, &out->str);
        }
        case 0xca: {
            if (!need(4)) return false;
            const uint32_t b = (uint32_t)be(4);
            float f;
            memcpy(&f, &b, 4);
            return num(f);
        }
        case 0xcb: {
            if (!need(8)) return false;
            const uint64_t b = be(8);
            double d;
            memcpy(&d, &b, 8);
            return num(d);
        }
        case 0xcc: return need(1) && num((double)be(1));
        case 0xcd: return need(2) && num((double)be(2));
        case 0xce: return need(4) && num((double)be(4));
        case 0xcf: return need(8) && num((double)be(8));
        case 0xd0: return need(1) && num((double)(int8_t)be(1));
        case 0xd1: return need(2) && num((double)(int16_t)be(2));
        case 0xd2: return need(4) && num((double)(int32_t)be(4));
        case 0xd3: return need(8) && num((double)(int64_t)be(8));
        case 0xd9: case 0xda: case 0xdb: {
            const int lb = t == 0xd9 ? 1 : t == 0xda ? 2 : 4;
            if (!need(lb)) return false;
            const size_t len = (size_t)be(lb);
            out->kind = JSON::String;
            return str(len, &out->str);
        }
        case 0xdc: return need(2) && seq((size_t)be(2), out);
        case 0xdd: return need(4) && seq((size_t)be(4), out);
        case 0xde: return need(2) && map((size_t)be(2), out);
        case 0xdf: return need(4) && map((size_t)be(4), out);
        default: err = "msgpack: type byte not used by JSON documents"; return false;  // ext / reserved
    }
}
}  // namespace
bool msgpack_decode(const uint8_t* p, size_t n, JSON* out, std::string* err) {
    Reader r{p, n};
    *out = JSON();
    if (!r.value(out) || r.i != n) {
        if (err) *err = r.err.empty() ? "msgpack: trailing bytes" : r.err;
        return false;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ the container
int kun_parse(const std::string& path, KunFile* out, std::string* err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) {
        *err = "cannot open '" + path + "'";
        return -1;
    }
    struct stat st;
    unsigned char lenb[8];
    if (fstat(fileno(f), &st) != 0 || st.st_size < 8 || fread(lenb, 1, 8, f) != 8) {
        fclose(f);
        *err = "'" + path + "': not a .kun / safetensors file (shorter than its 8-byte header length)";
        return -1;
    }
    uint64_t n = 0;
    for (int i = 7; i >= 0; i--) n = (n << 8) | lenb[i];
    if (n < 2 || n > (uint64_t)st.st_size - 8 || n > (100ull << 20)) {
        fclose(f);
        *err = "'" + path + "': header length " + std::to_string(n) + " does not fit the file";
        return -1;
    }
    std::string text(n, '\0');
    const bool got = fread(&text[0], 1, n, f) == n;
    fclose(f);
    if (!got) {
        *err = "'" + path + "': truncated header";
        return -1;
    }
    out->path = path, out->data_start = 8 + n, out->data_bytes = (uint64_t)st.st_size - 8 - n;
    out->entries.clear();
    out->has_config = false;
    try {
        const JSON j = JSON::parse(text);
        if (!j.is_object()) throw std::runtime_error("header is not a JSON object");
        for (auto& kv : j.obj) {
            if (kv.first == "__metadata__") continue;
            const JSON& v = kv.second;
            KunEntry e;
            e.name  = kv.first;
            e.dtype = v.at("dtype").as_string();
            for (auto& d : v.at("shape").arr) {
                if (d.as_double(-1) < 0) throw std::runtime_error("tensor '" + e.name + "': negative dimension");
                e.shape.push_back((int64_t)d.as_double());
            }
            const JSON& off = v.at("data_offsets");
            if (!off.is_array() || off.arr.size() != 2) throw std::runtime_error("tensor '" + e.name + "': data_offsets must be [begin, end]");
            e.begin = (uint64_t)off.arr[0].as_double(), e.end = (uint64_t)off.arr[1].as_double();
            if (e.end < e.begin || e.end > out->data_bytes) throw std::runtime_error("tensor '" + e.name + "': data_offsets exceed the file");
            const int bits = kun_dtype_bits(e.dtype);
            if (!bits) throw std::runtime_error("tensor '" + e.name + "': unknown dtype '" + e.dtype + "'");
            uint64_t numel = 1;
            for (int64_t d : e.shape) numel *= (uint64_t)d;
            if (numel * (uint64_t)bits % 8) throw std::runtime_error("tensor '" + e.name + "': shape x dtype is not a whole number of bytes");
            const uint64_t plain = numel * (uint64_t)bits / 8;
            if (v.contains("szData") && v.contains("szGama")) {
                e.has_sizes = true;
                e.szData = (uint64_t)v.at("szData").as_double(), e.szGama = (uint64_t)v.at("szGama").as_double();
            } else {
                e.szData = plain, e.szGama = 0;
            }
            if (e.name == kConfigKey) {  // its szData / szGama are those of an empty tensor: only the offsets count
                out->has_config = true, out->config = e;
                continue;
            }
            if (e.szData != plain) throw std::runtime_error("tensor '" + e.name + "': szData does not match shape x dtype");
            const uint64_t span = e.end - e.begin;
            if (span != e.szData + e.szGama) {
                if (e.szData + e.szGama > 0 && span > e.szData + e.szGama && span % (e.szData + e.szGama) == 0)
                    throw std::runtime_error("tensor '" + e.name + "': payload holds optimizer state after the weights (a training-state checkpoint): out of scope");
                throw std::runtime_error("tensor '" + e.name + "': data_offsets do not span szData + szGama bytes");
            }
            out->entries.push_back(std::move(e));
        }
    } catch (const std::exception& ex) {
        *err = "'" + path + "': " + ex.what();
        return -1;
    }
    std::sort(out->entries.begin(), out->entries.end(), [](const KunEntry& a, const KunEntry& b) { return a.begin < b.begin; });
    return 0;
}
int kun_read(const KunFile& file, const KunEntry& e, void* dst, std::string* err) {
    FILE* f = fopen(file.path.c_str(), "rb");
    if (!f) {
        *err = "cannot open '" + file.path + "'";
        return -1;
    }
    const uint64_t n = e.end - e.begin;
    int rc = 0;
    if (fseeko(f, (off_t)(file.data_start + e.begin), SEEK_SET) != 0 || fread(dst, 1, n, f) != n) {
        *err = "'" + file.path + "': cannot read the bytes of '" + e.name + "'";
        rc   = -1;
    }
    fclose(f);
    return rc;
}
int kun_config_json(const KunFile& f, std::string* text, std::string* err) {
    text->clear();
    if (!f.has_config) return 0;
    std::vector<uint8_t> raw(f.config.end - f.config.begin);
    if (kun_read(f, f.config, raw.data(), err) != 0) return -1;
    JSON j;
    if (!msgpack_decode(raw.data(), raw.size(), &j, err)) return -1;
    *text = json_dump(j);
    return 0;
}

int kun_write(const std::string& path, const std::string& config_json, const std::vector<KunTensorOut>& tensors, std::string* err) {
    try {
        JSON header;
        header.kind = JSON::Object;
        auto S = [](const std::string& s) {
            JSON j;
            j.kind = JSON::String, j.str = s;
            return j;
        };
        auto N = [](double v) {
            JSON j;
            j.kind = JSON::Number, j.num = v;
            return j;
        };
        JSON meta;  // K_SafeTensors::UpdateMetaData, Serialize.cpp:842-847
        meta.kind = JSON::Object;
        meta.obj.push_back({"format", S("pt")});
        meta.obj.push_back({"writer", S("koifish")});
        header.obj.push_back({"__metadata__", meta});
        uint64_t off = 0;
        auto entry = [&](const std::string& dtype, const std::vector<int64_t>& shape, uint64_t bytes, uint64_t szData, uint64_t szGama) {
            JSON e, sh, offs;
            e.kind = JSON::Object, sh.kind = JSON::Array, offs.kind = JSON::Array;
            for (int64_t d : shape) sh.arr.push_back(N((double)d));
            offs.arr.push_back(N((double)off)), offs.arr.push_back(N((double)(off + bytes)));
            e.obj.push_back({"dtype", S(dtype)});
            e.obj.push_back({"shape", sh});
            e.obj.push_back({"data_offsets", offs});
            e.obj.push_back({"loAB", N(0)});
            e.obj.push_back({"szGama", N((double)szGama)});
            e.obj.push_back({"szData", N((double)szData)});
            off += bytes;
            return e;
        };
        for (const KunTensorOut& t : tensors) {
            const int bits = kun_dtype_bits(t.dtype);
            if (!bits) throw std::runtime_error("tensor '" + t.name + "': unknown dtype '" + t.dtype + "'");
            std::vector<int64_t> shape = {t.shape[0]};
            if (t.shape[1] > 0) shape.push_back(t.shape[1]);
            uint64_t numel = 1;
            for (int64_t d : shape) numel *= (uint64_t)d;
            if (numel * (uint64_t)bits / 8 != t.szData || !t.blob) throw std::runtime_error("tensor '" + t.name + "': szData does not match shape x dtype");
            for (auto& kv : header.obj)
                if (kv.first == t.name) throw std::runtime_error("tensor '" + t.name + "' appears twice");
            header.obj.push_back({t.name, entry(t.dtype, shape, t.szData + t.szGama, t.szData, t.szGama)});
        }
        std::vector<uint8_t> cfg;
        if (!config_json.empty()) {  // K_SafeTensors::insertJS: the config as msgpack, registered after the tensors
            msgpack_encode(JSON::parse(config_json), &cfg);
            header.obj.push_back({kConfigKey, entry("U8", {(int64_t)cfg.size()}, cfg.size(), 0, 0)});
        }
        std::string text = json_dump(header);
        text.append((8 - text.size() % 8) % 8, ' ');
        FILE* f = fopen(path.c_str(), "wb");
        if (!f) throw std::runtime_error("cannot open '" + path + "' for writing");
        unsigned char lenb[8];
        for (int i = 0; i < 8; i++) lenb[i] = (unsigned char)((uint64_t)text.size() >> (8 * i));
        bool ok = fwrite(lenb, 1, 8, f) == 8 && fwrite(text.data(), 1, text.size(), f) == text.size();
        for (const KunTensorOut& t : tensors) ok = ok && fwrite(t.blob, 1, t.szData + t.szGama, f) == t.szData + t.szGama;
        if (!cfg.empty()) ok = ok && fwrite(cfg.data(), 1, cfg.size(), f) == cfg.size();
        ok = (fclose(f) == 0) && ok;
        if (!ok) throw std::runtime_error("short write to '" + path + "'");
        return 0;
    } catch (const std::exception& ex) {
        *err = ex.what();
        return -1;
    }
}

}  // namespace koifish
