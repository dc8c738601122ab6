// ref_kun.cpp -- TEST INFRASTRUCTURE ONLY.  The reference's own checkpoint reader on a fish.kun file: K_SafeTensors(nullptr, {}, path) + MMAP(path), i.e.
// what SAFETENSOR_Load_jconfig does (reference src/Manifold/Serialize.cpp:428-520: mmap_from_file, validate_data_offsets, loadJS of the msgpack
// config entry), then every parsed tensor described by the reference's own GTensor::jDesc (:61-100).  Built by `make -C oracle refkun` from
// src/Manifold/Serialize.cpp, src/Tensor/Safetensors.cpp, src/Utils/GST_util.cpp (+ the objects of the refcpu target) into
// oracle/_ref/libkoifish_refkun.so; the rest of the framework is bound to 0 at link time and never reached.  Nothing of the reference is copied.
// Used to check that files written by csrc/Tensor/KunFile.cpp are files the reference reads (tests/test_kun_host.py).
#include <cstdio>
#include <cstring>
#include <string>
#include "Manifold/Fish.hpp"
#include "Tensor/Safetensors.hpp"
#include "Tensor/GTensor.hpp"
// what SAFETENSOR_Load_jconfig does (Serialize.cpp:496-520), keeping the parsed tensors: K_SafeTensors(nullptr, {}, path) + MMAP(path)
extern "C" int refcpu_kun_read(const char* path, char* out, int cap) {
    K_SafeTensors st(nullptr, {}, path);
    if (!st.MMAP(path, false, 0)) return -2;
    JSON j;
    j["config"] = st.jsConfig;
    JSON arr = JSON::array();
    for (size_t i = 0; i < st.tensors.size(); i++) {
        hGTensor t = st.tensors.at(i);
        JSON e    = t->jDesc(&st);  // the reference's own description of what it parsed (GTensor::jDesc, Serialize.cpp:61-100)
        e["name"] = st.tensors.keys()[i];
        arr.push_back(e);
    }
    j["tensors"] = arr;
    const std::string s = j.dump();
    if ((int)s.size() + 1 > cap) return -1;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}
