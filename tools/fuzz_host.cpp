// fuzz_host.cpp -- AddressSanitizer / UBSan run over the host-only code that parses external input: the tokenizer (tokenizer.json pipeline, NFC,
// streaming decode), the msgpack codec and the fish.kun container reader.  Random texts must round-trip, random bytes / mutated files must be
// refused or parsed, never crash.  From the repo root:
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -Ikoifish_b200/csrc -Iinclude tools/fuzz_host.cpp \
//       koifish_b200/csrc/TokenSet/HF_Tokenizer.cpp koifish_b200/csrc/Tensor/KunFile.cpp -o /tmp/fuzz_host && /tmp/fuzz_host
// Last run (round 2): "ok tokens=360691", no sanitizer report.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <fstream>
#include <sstream>
#include "TokenSet/HF_Tokenizer.hpp"
#include "Tensor/KunFile.hpp"
using namespace koifish;
int main() {
    std::string err;
    auto tk = HF_Tokenizer::FromPath("tests/golden/tokenizer", &err);
    if (!tk) { printf("load failed %s\n", err.c_str()); return 1; }
    std::mt19937 rng(7);
    const char* frags[] = {"a", "Z", "9", " ", "  ", "\n", "\r\n", "\t", "'s", "'LL", "<|im_start|>", "<|im_end|>", "<think>", "\xe5\xa4\xa9", "\xf0\x9f\x98\x80", "e\xcc\x81",
                           "\xcc\x88", "\xe1\x84\x80", "\xe1\x85\xa1", "\xe1\x86\xa8", "\xc2\xa0", "\xe3\x80\x80", "!", "...", "\xea\xb0\x81", "\xd9\x8e", "\xe0\xa5\x98"};
    size_t total = 0;
    for (int it = 0; it < 20000; it++) {
        std::string s;
        int n = rng() % 24;
        for (int i = 0; i < n; i++) s += frags[rng() % (sizeof(frags) / sizeof(frags[0]))];
        auto ids = tk->encode(s);
        std::string back = tk->decode(ids, false);
        if (back != HF_Tokenizer::NFC(s)) { printf("round trip mismatch\n"); return 2; }
        std::string pending, acc;
        for (int id : ids) acc += tk->stream_push(&pending, id, false);
        acc += HF_Tokenizer::stream_flush(&pending);
        if (acc != back) { printf("stream mismatch\n"); return 3; }
        total += ids.size();
    }
    // raw random bytes: must throw or work, never crash
    for (int it = 0; it < 20000; it++) {
        std::string s;
        int n = rng() % 16;
        for (int i = 0; i < n; i++) s += (char)(rng() & 0xff);
        try { tk->encode(s); } catch (const std::exception&) {}
        try { HF_Tokenizer::NFC(s); } catch (const std::exception&) {}
        std::vector<int> ids;
        for (int i = 0; i < n; i++) ids.push_back((int)(rng() % 1200) - 50);
        tk->decode(ids, rng() & 1);
        JSON j; std::string e;
        msgpack_decode((const uint8_t*)s.data(), s.size(), &j, &e);
    }
    // msgpack of nested documents and mutated encodings
    for (int it = 0; it < 3000; it++) {
        JSON j = JSON::parse("{\"a\":[1,2.5,-3,\"x\",null,true,{\"b\":[[],{}]}],\"c\":\"" + std::string(rng() % 300, 'q') + "\"}");
        std::vector<uint8_t> mp; msgpack_encode(j, &mp);
        JSON k; std::string e;
        if (!msgpack_decode(mp.data(), mp.size(), &k, &e) || json_dump(k) != json_dump(j)) { printf("msgpack round trip\n"); return 4; }
        for (int m = 0; m < 8; m++) { auto c = mp; c[rng() % c.size()] = (uint8_t)rng(); if (rng() & 1) c.resize(rng() % (c.size() + 1)); msgpack_decode(c.data(), c.size(), &k, &e); }
    }
    // kun headers: random mutations of a valid file
    {
        std::vector<uint8_t> blob(64 + 2 * (4 + 16 + 2 * 1) , 7);
        KunTensorOut t; t.name = "w"; t.dtype = "Q<4>"; t.shape[0] = 4; t.shape[1] = 32; t.szData = 64; t.szGama = blob.size() - 64; t.blob = blob.data();
        if (kun_write("/tmp/a.kun", "{\"x\":1}", {t}, &err)) { printf("kun_write %s\n", err.c_str()); return 5; }
        std::ifstream f("/tmp/a.kun", std::ios::binary); std::stringstream ss; ss << f.rdbuf(); std::string raw = ss.str();
        for (int it = 0; it < 5000; it++) {
            std::string c = raw;
            int k = 1 + rng() % 3;
            for (int m = 0; m < k; m++) c[rng() % c.size()] = (char)rng();
            if (rng() % 4 == 0) c.resize(rng() % (c.size() + 1));
            std::ofstream o("/tmp/b.kun", std::ios::binary); o.write(c.data(), c.size()); o.close();
            KunFile kf; std::string e2, cfg;
            if (kun_parse("/tmp/b.kun", &kf, &e2) == 0) {
                kun_config_json(kf, &cfg, &e2);
                for (auto& en : kf.entries) { std::vector<uint8_t> buf(en.end - en.begin); kun_read(kf, en, buf.data(), &e2); }
            }
        }
    }
    printf("ok tokens=%zu\n", total);
    return 0;
}
