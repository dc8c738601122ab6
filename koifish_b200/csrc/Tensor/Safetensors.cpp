// Safetensors.cpp -- HF "model.safetensors" reader: header index + tensor bytes -> kf_model_set_tensor (quantised at load per the model's
// quantizer card).  Replaces the reference's Fish::LoadFolderOfST -> SAFETENSOR2Gensors -> GTensor::LoadParam (.weight branch)
// (src/Manifold/Serialize.cpp:1010-1100, :145-230; parser src/Tensor/Safetensors.cpp, after syoyo/safetensors-cpp).
// File format (huggingface/safetensors): u64 little-endian header length N | N bytes of JSON { name: {"dtype","shape","data_offsets":[b,e]},
// "__metadata__": {...} } | the byte buffer the offsets index.
#include "Safetensors.hpp"

#include <dirent.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>

#include "../Utils/json_lite.hpp"

using koifish::JSON;

namespace {
int dtype_bytes(const std::string& d) {
    if (d == "BF16" || d == "F16") return 2;
    if (d == "F32" || d == "I32" || d == "U32") return 4;
    if (d == "F64" || d == "I64" || d == "U64") return 8;
    if (d == "I8" || d == "U8" || d == "BOOL" || d == "F8_E5M2" || d == "F8_E4M3") return 1;
    if (d == "I16" || d == "U16") return 2;
    return 0;
}
}  // namespace

int kf_st_parse(const std::string& path, KfStFile* out, std::string* err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) {
        *err = "cannot open '" + path + "'";
        return -1;
    }
    struct stat st;
    if (fstat(fileno(f), &st) != 0 || st.st_size < 8) {
        fclose(f);
        *err = "'" + path + "': not a safetensors file (shorter than its 8-byte header length)";
        return -1;
    }
    unsigned char lenb[8];
    if (fread(lenb, 1, 8, f) != 8) {
        fclose(f);
        *err = "'" + path + "': read error";
        return -1;
    }
    uint64_t n = 0;
    for (int i = 7; i >= 0; i--) n = (n << 8) | lenb[i];
    if (n < 2 || n > (uint64_t)st.st_size - 8 || n > (100ull << 20)) {  // the format caps the header at 100 MB
        fclose(f);
        *err = "'" + path + "': header length " + std::to_string(n) + " does not fit the file";
        return -1;
    }
    std::string text(n, '\0');
    if (fread(&text[0], 1, n, f) != n) {
        fclose(f);
        *err = "'" + path + "': truncated header";
        return -1;
    }
    fclose(f);
    out->path = path, out->data_start = 8 + n, out->data_bytes = (uint64_t)st.st_size - 8 - n;
    out->entries.clear();
    try {
        JSON j = JSON::parse(text);
        if (!j.is_object()) throw std::runtime_error("header is not a JSON object");
        for (auto& kv : j.obj) {
            if (kv.first == "__metadata__") continue;
            const JSON& v = kv.second;
            KfStEntry e;
            e.name  = kv.first;
            e.dtype = v.at("dtype").as_string();
            for (auto& d : v.at("shape").arr) e.shape.push_back((int64_t)d.as_double());
            const JSON& off = v.at("data_offsets");
            if (!off.is_array() || off.arr.size() != 2) throw std::runtime_error("tensor '" + e.name + "': data_offsets must be [begin, end]");
            e.begin = (uint64_t)off.arr[0].as_double(), e.end = (uint64_t)off.arr[1].as_double();
            uint64_t numel = 1;
            for (int64_t d : e.shape) {
                if (d < 0) throw std::runtime_error("tensor '" + e.name + "': negative dimension");
                numel *= (uint64_t)d;
            }
            const int eb = dtype_bytes(e.dtype);
            if (!eb) throw std::runtime_error("tensor '" + e.name + "': unknown dtype '" + e.dtype + "'");
            if (e.end < e.begin || e.end > out->data_bytes || e.end - e.begin != numel * (uint64_t)eb)
                throw std::runtime_error("tensor '" + e.name + "': data_offsets do not match shape x dtype or exceed the file");
            out->entries.push_back(std::move(e));
        }
    } catch (const std::exception& ex) {
        *err = "'" + path + "': " + ex.what();
        return -1;
    }
    std::sort(out->entries.begin(), out->entries.end(), [](const KfStEntry& a, const KfStEntry& b) { return a.begin < b.begin; });
    return 0;
}

int kf_st_read(const KfStFile& file, const KfStEntry& e, void* dst, std::string* err) {
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

// "x.safetensors" itself, or every *.safetensors of a directory in name order (sharded checkpoints: model-00001-of-0000N.safetensors)
int kf_st_list(const std::string& path, std::vector<std::string>* files, std::string* err) {
    struct stat st;
    if (stat(path.c_str(), &st) != 0) {
        *err = "no such file or directory: '" + path + "'";
        return -1;
    }
    files->clear();
    if (!S_ISDIR(st.st_mode)) {
        files->push_back(path);
        return 0;
    }
    DIR* d = opendir(path.c_str());
    if (!d) {
        *err = "cannot list '" + path + "'";
        return -1;
    }
    while (struct dirent* de = readdir(d)) {
        const std::string n = de->d_name;
        const std::string suf = ".safetensors";
        if (n.size() > suf.size() && n.compare(n.size() - suf.size(), suf.size(), suf) == 0) files->push_back(path + "/" + n);
    }
    closedir(d);
    std::sort(files->begin(), files->end());
    if (files->empty()) {
        *err = "no *.safetensors file in '" + path + "'";
        return -1;
    }
    return 0;
}

// RN conversions to bf16 (what GTensor::LoadParam's typed copy does for F16 / F32 sources)
void kf_st_to_bf16(const std::string& dtype, const void* src, size_t n, uint16_t* dst) {
    auto f32_to_bf16 = [](float f) -> uint16_t {
        uint32_t u;
        memcpy(&u, &f, 4);
        if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u);  // NaN stays NaN
        u += 0x7fffu + ((u >> 16) & 1u);
        return (uint16_t)(u >> 16);
    };
    if (dtype == "BF16") {
        memcpy(dst, src, n * 2);
    } else if (dtype == "F32") {
        const float* s = (const float*)src;
        for (size_t i = 0; i < n; i++) dst[i] = f32_to_bf16(s[i]);
    } else {  // F16
        const uint16_t* s = (const uint16_t*)src;
        for (size_t i = 0; i < n; i++) {
            const uint16_t h = s[i];
            const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ffu;
            uint32_t bits;
            if (exp == 0) {
                if (man == 0) {
                    bits = sign;
                } else {
                    int e = -1;
                    uint32_t m = man;
                    do { e++, m <<= 1; } while (!(m & 0x400u));
                    bits = sign | (uint32_t)(127 - 15 - e) << 23 | (m & 0x3ffu) << 13;
                }
            } else if (exp == 31) {
                bits = sign | 0x7f800000u | man << 13;
            } else {
                bits = sign | (exp + 112) << 23 | man << 13;
            }
            float f;
            memcpy(&f, &bits, 4);
            dst[i] = f32_to_bf16(f);
        }
    }
}
