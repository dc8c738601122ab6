"""The seeded inputs behind tests/golden/refgpu_golden.npz (outputs of the reference's own CUDA kernels on a B200): shared by the generator
(tests/golden/make_golden_refgpu.py, needs a GPU + oracle/_ref) and the CPU test (tests/test_oracle_golden_refgpu.py)."""
import numpy as np

import oracle_lib as ol

CASES = {
    # bits, mode, rows, cols, seed, sigma
    "dequant": [(4, ol.RTN_ASYM, 64, 512, 11, 0.02), (4, ol.RTN_SYM, 64, 512, 12, 1.7), (2, ol.RTN_ASYM, 32, 512, 13, 0.02),
                (2, ol.YYANG, 64, 512, 14, 0.02), (1, ol.YYANG, 64, 512, 15, 3e-4)],
    "rmsnorm": [(2, 1024, 1), (2, 5120, 2)],                                   # rows, dim, seed
    "rope": [(16, 8, 128, 0, 1e6, 3), (16, 8, 128, 300, 1e6, 4), (16, 8, 128, 511, 1e4, 5)],   # n_head, n_kv, hd, pos, theta, seed
    "attention": [(16, 8, 128, 512, 5, 0, 6), (16, 8, 128, 512, 511, 0, 7), (16, 8, 128, 512, 100, 1, 8)],  # ..., max_seq, pos, score_bf16, seed
    "nf4": [(48, 512, 21), (16, 5120, 22)],                                     # rows, cols, seed
    "awq": [(256, 64, 31), (1024, 264, 32)],                                    # in_features, out_features, seed
}


def dequant_inputs(bits, mode, rows, cols, seed, sigma):
    w = ol.fill_normal(rows * cols, seed, sigma)
    data, gama = ol.quantize(w, rows, cols, bits, 128, mode)
    return data, gama, ol.qrange(bits, mode)[2]


def rmsnorm_inputs(rows, dim, seed):
    rng = np.random.default_rng(seed)
    return (ol.f32_to_bf16((rng.standard_normal((rows, dim)) * 2.5).astype(np.float32)),
            ol.f32_to_bf16((1.0 + 0.2 * rng.standard_normal(dim)).astype(np.float32)))


def rope_inputs(n_head, n_kv, hd, seed):
    rng = np.random.default_rng(seed)
    return (ol.f32_to_bf16(rng.standard_normal((n_head, hd)).astype(np.float32)), ol.f32_to_bf16(rng.standard_normal((n_kv, hd)).astype(np.float32)))


def attention_inputs(n_head, n_kv, hd, max_seq, seed):
    rng = np.random.default_rng(seed)
    return (ol.f32_to_bf16((rng.standard_normal((n_head, hd)) * 0.5).astype(np.float32)),
            ol.f32_to_bf16(rng.standard_normal((max_seq, n_kv * hd)).astype(np.float32)),
            ol.f32_to_bf16(rng.standard_normal((max_seq, n_kv * hd)).astype(np.float32)))


def nf4_inputs(rows, cols, seed):
    return ol.nf4_quantize(ol.fill_normal(rows * cols, seed, 0.02), rows, cols)


def awq_inputs(IC, OC, seed):
    return ol.awq_pack(ol.fill_normal(IC * OC, seed, 0.02), IC, OC)
