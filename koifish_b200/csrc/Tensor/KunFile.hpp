// KunFile.hpp -- see KunFile.cpp
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../Utils/json_lite.hpp"

namespace koifish {

struct KunEntry {
    std::string name, dtype;      // dtype: the reference's K_FLOATS name ("Q<4>", "TERNARY", "BF16(E8)", ...) or an HF name ("BF16")
    std::vector<int64_t> shape;
    uint64_t begin = 0, end = 0;  // byte range inside the data region
    uint64_t szData = 0, szGama = 0;
    bool has_sizes = false;       // the entry carried szData / szGama (CKP_KOIFISH); plain HF entries do not
};
struct KunFile {
    std::string path;
    uint64_t data_start = 0, data_bytes = 0;
    std::vector<KunEntry> entries;  // in file order, the config entry excluded
    bool has_config = false;
    KunEntry config;                // "__koifish__config__": U8 [n], msgpack of the writer's JSON config
};
int kun_dtype_bits(const std::string& dtype);  // 0: unknown
int kun_parse(const std::string& path, KunFile* out, std::string* err);
int kun_read(const KunFile& f, const KunEntry& e, void* dst, std::string* err);
int kun_config_json(const KunFile& f, std::string* json_text, std::string* err);  // "" when the file has no config entry

struct KunTensorOut {
    std::string name, dtype;
    int64_t shape[2] = {0, 0};  // shape[1] == 0: a vector
    uint64_t szData = 0, szGama = 0;
    const void* blob = nullptr;  // szData + szGama bytes (data || gama)
};
int kun_write(const std::string& path, const std::string& config_json, const std::vector<KunTensorOut>& tensors, std::string* err);

std::string json_dump(const JSON& j);
void msgpack_encode(const JSON& j, std::vector<uint8_t>* out);
bool msgpack_decode(const uint8_t* p, size_t n, JSON* out, std::string* err);

}  // namespace koifish
