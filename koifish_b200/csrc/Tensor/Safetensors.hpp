// Safetensors.hpp -- see Safetensors.cpp
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

struct KfStEntry {
    std::string name, dtype;
    std::vector<int64_t> shape;
    uint64_t begin = 0, end = 0;  // byte range inside the file's data region
};
struct KfStFile {
    std::string path;
    uint64_t data_start = 0, data_bytes = 0;
    std::vector<KfStEntry> entries;  // in file order
};
int kf_st_parse(const std::string& path, KfStFile* out, std::string* err);
int kf_st_read(const KfStFile& file, const KfStEntry& e, void* dst, std::string* err);
int kf_st_list(const std::string& path, std::vector<std::string>* files, std::string* err);
void kf_st_to_bf16(const std::string& dtype, const void* src, size_t n, uint16_t* dst);
