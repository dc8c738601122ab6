// ref_kun.cpp -- TEST INFRASTRUCTURE ONLY.  The reference's own checkpoint reader on a fish.kun file: K_SafeTensors(nullptr, {}, path) + MMAP(path), i.e.
// what SAFETENSOR_Load_jconfig does (reference src/Manifold/Serialize.cpp:428-520: mmap_from_file, validate_data_offsets, loadJS of the msgpack
// config entry), then every parsed tensor described by the reference's own GTensor::jDesc (:61-100).  Built by `make -C oracle refkun` from
// src/Manifold/Serialize.cpp, src/Tensor/Safetensors.cpp, src/Utils/GST_util.cpp (+ the objects of the refcpu target) into
// oracle/_ref/libkoifish_refkun.so; the rest of the framework is bound to 0 at link time and never reached.  Nothing of the reference is copied.
// Used to check that files written by csrc/Tensor/KunFile.cpp are files the reference reads, and -- second half of this file -- that files the
// reference's own writer produces are read by csrc/Tensor/KunFile.cpp (tests/test_kun_host.py, tests/golden/make_golden_refkun.py).
#include <cstdio>
#include <cstring>
#include <string>
#include "Manifold/Fish.hpp"
#include "Manifold/Serialize.hpp"
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

// ---- the reference's own WRITER: K_SafeTensors::Register + insertJS + Save (Serialize.cpp:286-360, 554-680, 860-880; Safetensors.hpp:87-102), as
// Fish::SAFETENSOR_Serialize drives them (:935-960), on tensors whose payload sits in host memory.  The writer asks every tensor for GetDataX() -- a
// CUDA routine of the framework (kernel/quantizer.cu) -- and, when that returns nullptr and FSerial::COPY_MMAP is set, copies host_data instead
// (:619-631): that branch is the one taken here, so GetDataX is answered with nullptr.
floatX* GTensor::GetDataX(int, const string&) { return nullptr; }
namespace {
struct HostTensor : public GTensor {
    void Setup(const char* nm, const char* dtype, long long d0, long long d1, size_t szD, size_t szG, const void* blob) {
        strcpy(name, nm);
        type = tpNumOf(dtype);
        if (d1 > 0)
            shape = {(int)d0, (int)d1};
        else
            shape = {(int)d0};
        szData = szD, szGama = szG;
        host_data = const_cast<void*>(blob);
    }
};
}  // namespace
extern "C" int refcpu_kun_write(const char* path, const char* config_json, int n, const char* const* names, const char* const* dtypes, const long long* shapes,
                                const unsigned long long* sz_data, const unsigned long long* sz_gama, const void* const* blobs) {
    CheckPoint_Params ckp;
    ckp.format = CKP_KOIFISH, ckp.state_type = CheckPoint_Params::FULL;
    // the container and its tensors are leaked on purpose: ~GTensor frees host_data, which belongs to the caller here
    K_SafeTensors& st = *new K_SafeTensors(nullptr, ckp, path);
    st.Clear();
    st.UpdateMetaData();
    size_t off = 0;
    for (int i = 0; i < n; i++) {
        std::shared_ptr<HostTensor> t(new HostTensor(), [](HostTensor*) {});
        t->Setup(names[i], dtypes[i], shapes[2 * i], shapes[2 * i + 1], (size_t)sz_data[i], (size_t)sz_gama[i], blobs[i]);
        off = st.Register(t, off, ckp.format);
    }
    if (config_json) st.insertJS(JSON::parse(config_json), off);
    std::string warn, err;
    size_t sz = off;
    if (!st.Save(path, sz, &warn, &err, FSerial::COPY_MMAP)) {
        fprintf(stderr, "refcpu_kun_write: %s\n", err.c_str());
        return -1;
    }
    return 0;
}
