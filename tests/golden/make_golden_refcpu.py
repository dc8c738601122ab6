"""Golden vectors from the reference's OWN CPU packers -- GeQuant::RTN_x (4- / 2-bit asymmetric and symmetric, 2-bit ternary), GeQuant::YinYang
(1-bit) and RT_NormalF / _row_lut (NormalFloat4), reference src/Tensor/GeQuant.cpp:428-533, 536-628, 706-752 -- compiled from the reference tree
into oracle/_ref/libkoifish_refcpu.so (oracle/ref_cpu_quant.cpp, `make -C oracle refcpu`).  Inputs come from the oracle's counter-based generator
(seeded), outputs are the packed bytes and the gama array exactly as the reference writes them.  Writes tests/golden/refcpu_quant.npz; the CPU
suite checks the oracle's restatement against it without needing /root/reference.
Run in the build container:  python tests/golden/make_golden_refcpu.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

CASES = [  # (tag, rows, cols, bits, mode, seed, sigma)
    ("q4_asym", 24, 512, 4, ol.RTN_ASYM, 11, 0.02), ("q4_sym", 24, 512, 4, ol.RTN_SYM, 12, 0.02), ("q2_asym", 16, 512, 2, ol.RTN_ASYM, 13, 0.02),
    ("q2_sym", 16, 512, 2, ol.RTN_SYM, 14, 0.02), ("ternary", 16, 512, 2, ol.YYANG, 15, 0.02), ("binary", 16, 512, 1, ol.YYANG, 16, 0.02),
    ("nf4", 12, 640, 4, ol.NF4, 17, 0.02), ("q4_asym_wide", 8, 1024, 4, ol.RTN_ASYM, 18, 1.5), ("ternary_tiny", 8, 256, 2, ol.YYANG, 19, 1e-4),
]


def main():
    assert ol.refcpu() is not None, "oracle/_ref/libkoifish_refcpu.so is not built (needs /root/reference)"
    out = {}
    for tag, rows, cols, bits, mode, seed, sigma in CASES:
        w = ol.fill_normal(rows * cols, seed, sigma)
        data, gama, qb = ol.refcpu_quantize(w, rows, cols, bits, 128, mode)
        out[tag + "_meta"] = np.array([rows, cols, bits, mode, seed, qb], dtype=np.int64)
        out[tag + "_sigma"] = np.array([sigma], dtype=np.float64)
        out[tag + "_data"], out[tag + "_gama"] = data, gama
    np.savez_compressed(os.path.join(HERE, "refcpu_quant.npz"), **out)
    print("wrote refcpu_quant.npz:", [c[0] for c in CASES])


if __name__ == "__main__":
    main()
