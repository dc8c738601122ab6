"""CPU test (no GPU, no /root/reference): the oracle against tests/golden/refgpu_golden.npz -- outputs of the reference's OWN CUDA kernels
(CU_Q128toX_ in both rounding builds, rms_norm_kernel, CU_rope2_v0, the three attention kernels, CU_Q42X_NF4, CU_Q42X_awq), produced on a
B200 by tests/golden/make_golden_refgpu.py from the seeded inputs of tests/golden_cases.py.  Bit-exact where the reference computes with
plain arithmetic (every dequant), the tolerance of tests/test_gpu_refkernels.py where it uses fast-math intrinsics or stochastic rounding."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from golden_cases import CASES, attention_inputs, awq_inputs, dequant_inputs, nf4_inputs, rmsnorm_inputs, rope_inputs

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refgpu_golden.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="tests/golden/refgpu_golden.npz not generated yet")


@pytest.fixture(scope="module")
def G():
    return np.load(PATH)


def ulp_diff(a, b):
    def key(x):
        x = x.astype(np.int32)
        return np.where(x & 0x8000, -(x & 0x7fff), x & 0x7fff)
    return np.abs(key(np.asarray(a).reshape(-1)) - key(np.asarray(b).reshape(-1)))


@pytest.mark.parametrize("variant", ["fma", "nofma"])
def test_dequant_bit_exact_vs_reference_kernel_outputs(G, variant):
    ol.set_dequant_fma(1 if variant == "fma" else 0)
    try:
        for (bits, mode, rows, cols, seed, sigma) in CASES["dequant"]:
            data, gama, qbias = dequant_inputs(bits, mode, rows, cols, seed, sigma)
            ref = G["dequant_%s_b%d_m%d_s%d" % (variant, bits, mode, seed)]
            assert np.array_equal(ol.dequant(data, gama, rows, cols, bits, 128, qbias).reshape(-1), ref), (variant, bits, mode)
    finally:
        ol.set_dequant_fma(1)


def test_nf4_and_awq_dequant_bit_exact_vs_reference_kernel_outputs(G):
    for (rows, cols, seed) in CASES["nf4"]:
        data, gama = nf4_inputs(rows, cols, seed)
        assert np.array_equal(ol.nf4_dequant(data, gama, rows, cols).reshape(-1), G["nf4_%d_%d_%d" % (rows, cols, seed)])
    for (IC, OC, seed) in CASES["awq"]:
        qw, qz, sc = awq_inputs(IC, OC, seed)
        assert np.array_equal(ol.awq_dequant(qw, qz, sc, IC, OC).reshape(-1), G["awq_%d_%d_%d" % (IC, OC, seed)])


def test_rmsnorm_vs_reference_kernel_outputs(G):
    # rsqrtf under -use_fast_math: <= 1 bf16 ulp, >= 99 % bit-equal
    for (rows, dim, seed) in CASES["rmsnorm"]:
        x, w = rmsnorm_inputs(rows, dim, seed)
        d = ulp_diff(ol.rmsnorm(x, w, rows, dim), G["rmsnorm_%d_%d" % (dim, seed)])
        assert d.max() <= 1 and (d == 0).mean() >= 0.99


def test_rope_vs_reference_kernel_outputs(G):
    # the reference rounds stochastically and evaluates powf / sincosf with fast-math: 2 bf16 ulps of the pair's magnitude
    for (n_head, n_kv, hd, pos, theta, seed) in CASES["rope"]:
        q, k = rope_inputs(n_head, n_kv, hd, seed)
        for src, nh, key in ((q, n_head, "rope_q_%d_%g_%d"), (k, n_kv, "rope_k_%d_%g_%d")):
            a = ol.bf16_to_f32(ol.rope(src, nh, hd, pos, theta)).reshape(-1, hd)
            b = ol.bf16_to_f32(G[key % (pos, theta, seed)]).reshape(-1, hd)
            s = np.abs(ol.bf16_to_f32(src).reshape(-1, hd))
            scale = np.maximum(s[:, :hd // 2], s[:, hd // 2:])
            scale = np.concatenate([scale, scale], axis=1)
            assert np.all(np.abs(a - b) <= 2.0 ** -6 * scale + 1e-6), (pos, theta)


def test_attention_vs_reference_kernel_outputs(G):
    for (n_head, n_kv, hd, max_seq, pos, score_bf16, seed) in CASES["attention"]:
        q, kc, vc = attention_inputs(n_head, n_kv, hd, max_seq, seed)
        b = ol.bf16_to_f32(G["attention_%d_%d_%d" % (pos, score_bf16, seed)])
        o = ol.bf16_to_f32(ol.attention_decode(q, kc, vc, pos, n_head, n_kv, hd, score_bf16).reshape(-1))
        tol = 2.0 ** (-5 if score_bf16 else -7)
        assert np.abs(o - b).max() <= tol * max(1e-3, np.abs(b).max()), (pos, score_bf16)
