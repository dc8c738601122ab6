#!/usr/bin/env python
"""Golden vectors from the reference's OWN CUDA kernels (oracle/_ref/libkoifish_refgpu*.so, libkoifish_refq.so: built from
/root/reference/src/Device/CUDA/T.cu and kernel/quantizer.cu by oracle/Makefile), run on a B200:

    gpurun -- 'python tests/golden/make_golden_refgpu.py gpurun_out/refgpu_golden.npz'      (then copy the file into tests/golden/)

Inputs are regenerated from seeds by tests/test_oracle_golden_refgpu.py (the oracle's own generator / numpy default_rng), so the file
holds only the reference kernels' OUTPUTS.  The CPU test then checks the oracle against them without a GPU and without /root/reference."""
import sys

import numpy as np

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
import koifish_b200 as kf  # noqa: E402
import oracle_lib as ol  # noqa: E402

from golden_cases import CASES, attention_inputs, awq_inputs, dequant_inputs, nf4_inputs, rmsnorm_inputs, rope_inputs  # noqa: E402


def main(path):
    ctx = kf.Context(0)
    out = {}
    for variant in ("fma", "nofma"):
        R = ol.refgpu(variant)
        assert R is not None, "oracle/_ref/libkoifish_refgpu*.so missing"
        for (bits, mode, rows, cols, seed, sigma) in CASES["dequant"]:
            data, gama, qbias = dequant_inputs(bits, mode, rows, cols, seed, sigma)
            nG = rows * cols // 128
            d, g, o = ctx.array(data.view(np.uint16)), ctx.array(gama), ctx.empty(rows * cols * 2)
            assert R.refk_q128tox(bits, nG, 128, qbias, d.ptr, g.ptr + 2 * (rows + cols), g.ptr + 2 * (rows + cols) + 2 * nG, o.ptr) == 0
            out["dequant_%s_b%d_m%d_s%d" % (variant, bits, mode, seed)] = o.numpy(np.uint16)
    R = ol.refgpu("fma")
    for (rows, dim, seed) in CASES["rmsnorm"]:
        x, w = rmsnorm_inputs(rows, dim, seed)
        xd, wd, o = ctx.array(x), ctx.array(w), ctx.empty(rows * dim * 2)
        assert R.refk_rmsnorm(o.ptr, xd.ptr, wd.ptr, rows, dim) == 0
        out["rmsnorm_%d_%d" % (dim, seed)] = o.numpy(np.uint16)
    for (n_head, n_kv, hd, pos, theta, seed) in CASES["rope"]:
        q, k = rope_inputs(n_head, n_kv, hd, seed)
        qd, kd = ctx.array(q), ctx.array(k)
        assert R.refk_rope2(qd.ptr, kd.ptr, pos, n_head, n_kv, hd, theta) == 0
        out["rope_q_%d_%g_%d" % (pos, theta, seed)] = qd.numpy(np.uint16)
        out["rope_k_%d_%g_%d" % (pos, theta, seed)] = kd.numpy(np.uint16)
    for (n_head, n_kv, hd, max_seq, pos, score_bf16, seed) in CASES["attention"]:
        q, kc, vc = attention_inputs(n_head, n_kv, hd, max_seq, seed)
        qd, kcd, vcd = ctx.array(q), ctx.array(kc), ctx.array(vc)
        att, o = ctx.empty(n_head * max_seq * 4), ctx.empty(n_head * hd * 2)
        assert R.refk_attention(o.ptr, att.ptr, qd.ptr, kcd.ptr, vcd.ptr, pos, max_seq, n_head, n_kv, hd, score_bf16) == 0
        out["attention_%d_%d_%d" % (pos, score_bf16, seed)] = o.numpy(np.uint16)
    Q = ol.refq()
    assert Q is not None, "oracle/_ref/libkoifish_refq.so missing"
    for (rows, cols, seed) in CASES["nf4"]:
        data, gama = nf4_inputs(rows, cols, seed)
        d, g, o = ctx.array(data.view(np.uint16)), ctx.array(gama), ctx.empty(rows * cols * 2)
        ctx.sync()
        assert Q.refq_nf4_dequant(g.ptr, d.ptr, o.ptr, rows, cols) == 0
        out["nf4_%d_%d_%d" % (rows, cols, seed)] = o.numpy(np.uint16)
    for (IC, OC, seed) in CASES["awq"]:
        qw, qz, sc = awq_inputs(IC, OC, seed)
        a, b, c, o = ctx.array(qw.view(np.uint16)), ctx.array(qz.view(np.uint16)), ctx.array(sc), ctx.empty(IC * OC * 2)
        ctx.sync()
        assert Q.refq_awq_dequant(b.ptr, c.ptr, a.ptr, o.ptr, IC, OC) == 0
        out["awq_%d_%d_%d" % (IC, OC, seed)] = o.numpy(np.uint16)
    np.savez_compressed(path, **out)
    print("wrote %s: %d arrays, %d bytes of outputs" % (path, len(out), sum(v.nbytes for v in out.values())))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/refgpu_golden.npz")
