/*
 * ref_shim.cpp -- thin extern "C" wrapper that exposes the REFERENCE's own code to the tests.
 * TEST INFRASTRUCTURE ONLY (same rules as koifish_oracle.h).
 *
 * It #includes /root/reference/src/PackedQ.hpp where it lies (no reference source is copied into this repo) and is
 * linked with the reference's src/Utils/GST_float.cpp compiled from its own location by oracle/Makefile.  Output goes
 * to oracle/_ref/libkoifish_ref.so (git-ignored, travels to the GPU box with the snapshot).
 *
 *   ref_pack / ref_unpack      -> PACK_{4,2,1}to128_ / UNPACK_128to{4,2,1}_UNSIGNED_   (src/PackedQ.hpp:99-239)
 *   ref_matvec_f32             -> D_matvec + dotprod_fp32                               (src/Utils/GST_float.cpp:278-304)
 *   ref_rmsnorm_f32            -> rmsnorm                                                (src/Utils/GST_float.cpp:575-586)
 */
#include <stddef.h>
#include <stdint.h>

#include "PackedQ.hpp"

typedef float (*dotprod_t)(void* w, int n, int i, float* x);
float dotprod_fp32(void* w, int n, int row, float* x);
void D_matvec(float* xout, float* x, void* w, float* b, int nIn, int nOut, dotprod_t dotprod);
float rmsnorm(float* o, float* x, float* weight, int size, float eps);

extern "C" {

int ref_pack(const int32_t* codes, size_t n, int bits, uint8_t* out) {
    const size_t per = 128 / bits;
    if (n % per)
        return -2;
    for (size_t w = 0; w < n / per; w++) {
        const int32_t* qq = codes + w * per;
        BIT_128* dst      = (BIT_128*)out + w;
        if (bits == 4) {
            PACK_4to128_(qq, dst);
        } else if (bits == 2) {
            PACK_2to128_(qq, dst);
        } else if (bits == 1) {
            PACK_1to128_(qq, dst);
        } else
            return -1;
    }
    return 0;
}

int ref_unpack(const uint8_t* data, size_t n, int bits, int32_t* out) {
    const size_t per = 128 / bits;
    if (n % per)
        return -2;
    for (size_t w = 0; w < n / per; w++) {
        int32_t* qq        = out + w * per;
        const BIT_128* src = (const BIT_128*)data + w;
        if (bits == 4) {
            UNPACK_128to4_UNSIGNED_(src, qq);
        } else if (bits == 2) {
            UNPACK_128to2_UNSIGNED_(src, qq);
        } else if (bits == 1) {
            UNPACK_128to1_UNSIGNED_(src, qq);
        } else
            return -1;
    }
    return 0;
}

void ref_matvec_f32(float* xout, float* x, float* w, int nIn, int nOut) { D_matvec(xout, x, (void*)w, nullptr, nIn, nOut, dotprod_fp32); }
float ref_rmsnorm_f32(float* o, float* x, float* weight, int size, float eps) { return rmsnorm(o, x, weight, size, eps); }
}
