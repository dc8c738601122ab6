"""CPU tests: the oracle against the reference's own macros (oracle/_ref), the committed golden fixtures, the
in-code invariants the reference asserts, and exact bf16 arithmetic."""
import os

import numpy as np
import pytest

import oracle_lib as ol

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "packq_ref.npz"))


# ------------------------------------------------------------------ bit layout (PackedQ.hpp:99-239)
@pytest.mark.parametrize("bits", [4, 2, 1])
def test_pack_matches_golden_from_reference_macros(bits):
    codes = GOLD[f"codes{bits}"]
    assert np.array_equal(ol.pack_codes(codes, bits), GOLD[f"bytes{bits}"])
    assert np.array_equal(ol.unpack_codes(GOLD[f"bytes{bits}"], codes.size, bits), codes)
    oh = GOLD[f"onehot_codes{bits}"]
    assert np.array_equal(ol.pack_codes(oh, bits), GOLD[f"onehot_bytes{bits}"])


def test_survey_probe_bytes():
    # SURVEY.md 8c probe of PACK_4to128_: codes (3+7i)%16 -> low half "5c 7e 90 b2 d4 f6 18 3a", same for the high half
    b = ol.pack_codes(GOLD["kat4_codes"], 4)
    want = bytes.fromhex("5c7e90b2d4f6183a") * 2
    assert bytes(b) == want
    assert bytes(GOLD["kat4_bytes"]) == want


@pytest.mark.parametrize("bits", [4, 2, 1])
def test_layout_table_A2(bits):
    # code j of a word -> byte / bit field
    per = 128 // bits
    for j in range(per):
        codes = np.zeros(per, dtype=np.int32)
        codes[j] = (1 << bits) - 1
        b = ol.pack_codes(codes, bits)
        half = per // 2
        jj = j if j < half else j - half
        per_byte = 8 // bits
        byte = (15 if j < half else 7) - jj // per_byte
        shift = 8 - bits * (jj % per_byte + 1)
        want = np.zeros(16, dtype=np.uint8)
        want[byte] = ((1 << bits) - 1) << shift
        assert np.array_equal(b, want), (bits, j)


@pytest.mark.skipif(ol.ref() is None, reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("bits", [4, 2, 1])
def test_pack_unpack_vs_reference_macros_live(bits):
    rng = np.random.default_rng(bits)
    codes = rng.integers(0, 1 << bits, size=(128 // bits) * 257, dtype=np.int32)
    assert np.array_equal(ol.pack_codes(codes, bits), ol.ref_pack(codes, bits))
    data = rng.integers(0, 256, size=16 * 300, dtype=np.uint8)
    n = data.size * 8 // bits
    assert np.array_equal(ol.unpack_codes(data, n, bits), ol.ref_unpack(data, n, bits))


def test_pack_rejects_ragged():
    L = ol.lib()
    codes = np.zeros(33, dtype=np.int32)
    out = np.zeros(64, dtype=np.uint8)
    assert L.kfo_pack_codes(codes, 33, 4, out) == -2
    assert L.kfo_pack_codes(codes, 32, 3, out) == -1


# ------------------------------------------------------------------ bf16 arithmetic
def _rn_bf16_from_f64(x):
    """exact RN-even of a python float to bf16 bits via integer arithmetic (independent of the oracle)."""
    import math
    import struct
    if x == 0:
        return 0x8000 if math.copysign(1, x) < 0 else 0
    s = 0x8000 if x < 0 else 0
    a = abs(x)
    m, e = math.frexp(a)  # a = m * 2^e, m in [0.5,1)
    exp = e - 1           # a = (2m) * 2^exp
    if exp < -126:
        q = a / 2.0 ** (-126 - 7)  # subnormal: units of 2^-133
        r = int(math.floor(q))
        frac = q - r
        if frac > 0.5 or (frac == 0.5 and (r & 1)):
            r += 1
        return s | r
    q = a / 2.0 ** (exp - 7)  # in [128, 256)
    r = int(math.floor(q))
    frac = q - r
    if frac > 0.5 or (frac == 0.5 and (r & 1)):
        r += 1
    if r == 256:
        r, exp = 128, exp + 1
    if exp > 127:
        return s | 0x7F80
    return s | ((exp + 127) << 7) | (r - 128)


def test_bf16_mul_sub_single_rounding():
    L = ol.lib()
    rng = np.random.default_rng(1)
    a = rng.integers(0, 1 << 16, size=4000).astype(np.uint16)
    b = rng.integers(0, 1 << 16, size=4000).astype(np.uint16)
    # add a family with close exponents (typical step*k vs zero) and a family with far exponents
    base = ol.f32_to_bf16(rng.normal(0, 0.05, size=4000).astype(np.float32))
    a = np.concatenate([a, base])
    b = np.concatenate([b, ol.f32_to_bf16((ol.bf16_to_f32(base) * rng.uniform(0.3, 3, size=4000)).astype(np.float32))])
    fa, fb = ol.bf16_to_f32(a).astype(np.float64), ol.bf16_to_f32(b).astype(np.float64)
    for i in range(a.size):
        if not (np.isfinite(fa[i]) and np.isfinite(fb[i])):
            continue
        prod, diff = float(fa[i]) * float(fb[i]), float(fa[i]) - float(fb[i])
        if abs(prod) < 3e38:
            assert L.kfo_bf16_mul(int(a[i]), int(b[i])) == _rn_bf16_from_f64(prod)
        if abs(diff) < 3e38:
            assert L.kfo_bf16_sub(int(a[i]), int(b[i])) == _rn_bf16_from_f64(diff)


def test_f32_to_bf16_vectorised_matches_scalar():
    L = ol.lib()
    rng = np.random.default_rng(2)
    x = rng.normal(0, 1, size=2000).astype(np.float32)
    v = ol.f32_to_bf16(x)
    for i in range(x.size):
        assert L.kfo_f32_to_bf16(float(x[i])) == v[i]
    # ties go to even
    assert L.kfo_f32_to_bf16(np.float32(1.0 + 2 ** -8)) == 0x3F80
    assert L.kfo_f32_to_bf16(np.float32(1.0 + 3 * 2 ** -8)) == 0x3F82


# ------------------------------------------------------------------ synthetic weights
def test_fill_normal_is_deterministic_and_roughly_normal():
    w = ol.fill_normal(1 << 16, seed=42, sigma=0.02)
    w2 = ol.fill_normal(1 << 16, seed=42, sigma=0.02)
    assert np.array_equal(w, w2)
    assert not np.array_equal(w, ol.fill_normal(1 << 16, seed=43, sigma=0.02))
    f = ol.bf16_to_f32(w)
    assert abs(f.mean()) < 5e-4 and abs(f.std() - 0.02) < 5e-4
    # prefix property: element i does not depend on n
    assert np.array_equal(w[:100], ol.fill_normal(100, seed=42, sigma=0.02))


# ------------------------------------------------------------------ quantiser (GeQuant.cpp:428-628)
@pytest.mark.parametrize("bits,mode", [(4, ol.RTN_ASYM), (4, ol.RTN_SYM), (2, ol.RTN_ASYM), (2, ol.YYANG), (1, ol.YYANG)])
def test_quantize_invariants(bits, mode):
    rows, cols, G = 24, 512, 128
    w = ol.fill_normal(rows * cols, seed=7 + bits, sigma=0.02)
    data, gama = ol.quantize(w, rows, cols, bits, G, mode)
    qmin, qmax, qbias = ol.qrange(bits, mode)
    nG = rows * cols // G
    assert gama.size == rows + cols + 2 * nG                 # szGama, GeQuant.cpp:518
    assert not gama[: rows + cols].any()                     # R/C scales unused
    codes = ol.unpack_codes(data, rows * cols, bits)
    assert codes.min() >= qmin + qbias and codes.max() <= qmax + qbias   # assert(qid>=qMin && qid<=qMax), :490
    assert np.array_equal(ol.pack_codes(codes, bits), data)  # pack->unpack round trip, :500-506
    zero = ol.bf16_to_f32(gama[rows + cols: rows + cols + nG])
    step = ol.bf16_to_f32(gama[rows + cols + nG:])
    wf = ol.bf16_to_f32(w).reshape(nG, G)
    if mode == ol.RTN_ASYM:
        assert np.allclose(zero, -wf.min(1), rtol=2 ** -8)
        assert np.allclose(step, (wf.max(1) - wf.min(1)) / (qmax - qmin), rtol=2 ** -8)
        # every group uses code 0 (its minimum) and code qmax (its maximum)
        c = codes.reshape(nG, G)
        assert (c.min(1) == 0).all() and (c.max(1) == qmax).all()
    else:
        assert not zero.any()
    if mode == ol.YYANG and bits == 2:
        assert np.allclose(step, np.maximum(1e-5, np.abs(wf).mean(1)), rtol=2 ** -8)
    if bits == 1:
        assert np.allclose(step, np.maximum(1e-5, np.sqrt((np.maximum(wf, 0) ** 2).mean(1))), rtol=2 ** -8)
    # dequant error: within half a step (+ bf16 rounding of step/zero/products) for RTN
    dq = ol.bf16_to_f32(ol.dequant(data, gama, rows, cols, bits, G, qbias)).reshape(nG, G)
    if mode in (ol.RTN_ASYM, ol.RTN_SYM):
        bound = 0.5 * step[:, None] + 2 ** -7 * (np.abs(wf).max(1)[:, None] + np.abs(zero)[:, None])
        assert (np.abs(dq - wf) <= bound + 1e-9).all()


@pytest.mark.parametrize("fused", [1, 0])
def test_dequant_rounding_modes(fused):
    # T.cu:274 (step * (bf16)k - zero) in bf16: fused = 1 -> ONE rounding, w = RN(step*k - zero) (the reference's kernel as nvcc builds it
    # for sm_90+, fma.rn.bf16; pinned on the GPU by tests/test_gpu_refkernels.py); fused = 0 -> TWO, w = RN(RN(step*k) - zero)
    # (-fmad=false / pre-sm_90 builds).  Checked against an independent float64 evaluation.
    rows, cols, G, bits = 8, 256, 128, 4
    w = ol.fill_normal(rows * cols, seed=11, sigma=0.02)
    data, gama = ol.quantize(w, rows, cols, bits, G, ol.RTN_ASYM)
    nG = rows * cols // G
    codes = ol.unpack_codes(data, rows * cols, bits).reshape(nG, G)
    zero = ol.bf16_to_f32(gama[rows + cols: rows + cols + nG]).astype(np.float64)
    step = ol.bf16_to_f32(gama[rows + cols + nG:]).astype(np.float64)
    ol.set_dequant_fma(fused)
    try:
        dq = ol.dequant(data, gama, rows, cols, bits, G, 0).reshape(nG, G)
    finally:
        ol.set_dequant_fma(1)
    for g in range(nG):
        for i in range(G):
            if fused:
                assert dq[g, i] == _rn_bf16_from_f64(float(step[g]) * int(codes[g, i]) - float(zero[g]))
            else:
                p = _rn_bf16_from_f64(float(step[g]) * int(codes[g, i]))
                pf = float(ol.bf16_to_f32(np.array([p], dtype=np.uint16))[0])
                assert dq[g, i] == _rn_bf16_from_f64(pf - float(zero[g]))


def test_ternary_and_binary_levels():
    rows, cols, G = 4, 256, 128
    w = ol.fill_normal(rows * cols, seed=5, sigma=0.02)
    for bits in (2, 1):
        data, gama = ol.quantize(w, rows, cols, bits, G, ol.YYANG)
        _, _, qbias = ol.qrange(bits, ol.YYANG)
        nG = rows * cols // G
        step = gama[rows + cols + nG:]
        dq = ol.dequant(data, gama, rows, cols, bits, G, qbias).reshape(nG, G)
        for g in range(nG):
            allowed = {0, int(step[g])} | ({int(step[g]) | 0x8000, 0x8000} if bits == 2 else set())
            assert set(int(v) for v in np.unique(dq[g])) <= allowed


def test_f8e5m2_truncation():
    w = ol.fill_normal(4096, seed=3, sigma=0.02)
    b = ol.f8_encode(w)
    d = ol.f8_decode(b)
    f, g = ol.bf16_to_f32(w), ol.bf16_to_f32(d)
    # truncation toward zero of the fp16 mantissa to 2 bits: |g| <= |f| and relative error < 2^-2
    assert (np.abs(g) <= np.abs(f) * (1 + 2 ** -10)).all()
    nz = np.abs(f) > 1e-4
    assert (np.abs(f - g)[nz] <= np.abs(f)[nz] * 0.25).all()
    # exact on values already representable in e5m2
    exact = np.array([0x3F80, 0xBF80, 0x3FC0, 0x3E80, 0x0000], dtype=np.uint16)  # 1, -1, 1.5, 0.25, 0
    assert np.array_equal(ol.f8_decode(ol.f8_encode(exact)), exact)
    # idempotent
    assert np.array_equal(ol.f8_encode(d), b)


# ------------------------------------------------------------------ ops
def test_linear_against_float64():
    rng = np.random.default_rng(4)
    M, N, K = 3, 40, 384
    w = ol.f32_to_bf16(rng.normal(0, 0.02, size=(N, K)).astype(np.float32))
    x = ol.f32_to_bf16(rng.normal(0, 1, size=(M, K)).astype(np.float32))
    y = ol.bf16_to_f32(ol.linear(w, x, M, N, K))
    ref = ol.bf16_to_f32(x).astype(np.float64) @ ol.bf16_to_f32(w).astype(np.float64).T
    assert np.allclose(y, ref, rtol=2 ** -7, atol=1e-4)


@pytest.mark.skipif(ol.ref() is None, reason="oracle/_ref not built")
def test_linear_and_rmsnorm_against_reference_cpu_primitives():
    # D_matvec + dotprod_fp32 and rmsnorm from src/Utils/GST_float.cpp (compiled unchanged into oracle/_ref)
    rng = np.random.default_rng(5)
    N, K = 64, 512
    w = ol.f32_to_bf16(rng.normal(0, 0.02, size=(N, K)).astype(np.float32))
    x = ol.f32_to_bf16(rng.normal(0, 1, size=(1, K)).astype(np.float32))
    yo = ol.linear_f32(w, x, 1, N, K)[0]
    wf, xf = np.ascontiguousarray(ol.bf16_to_f32(w)), np.ascontiguousarray(ol.bf16_to_f32(x)[0])
    yr = np.zeros(N, dtype=np.float32)
    ol.ref().ref_matvec_f32(yr, xf, wf, K, N)
    assert np.allclose(yo, yr, rtol=1e-5, atol=1e-6)
    g = ol.f32_to_bf16(rng.normal(1, 0.1, size=K).astype(np.float32))
    no = ol.bf16_to_f32(ol.rmsnorm(x, g, 1, K, 1e-6))[0]
    nr = np.zeros(K, dtype=np.float32)
    ol.ref().ref_rmsnorm_f32(nr, xf, np.ascontiguousarray(ol.bf16_to_f32(g)), K, 1e-6)
    assert np.allclose(no, nr, rtol=2 ** -7, atol=1e-6)


def test_rope_is_a_rotation_and_identity_at_pos0():
    rng = np.random.default_rng(6)
    H, hd = 4, 128
    v = ol.f32_to_bf16(rng.normal(0, 1, size=(H, hd)).astype(np.float32))
    assert np.array_equal(ol.rope(v, H, hd, 0, 1e6), v)
    r = ol.bf16_to_f32(ol.rope(v, H, hd, 37, 1e6))
    f = ol.bf16_to_f32(v)
    n0 = f[:, :64] ** 2 + f[:, 64:] ** 2
    n1 = r[:, :64] ** 2 + r[:, 64:] ** 2
    assert np.allclose(n0, n1, rtol=0.03, atol=1e-3)
    # last pair rotates slowest: angle = 37 / theta^(126/128)
    j = 63
    ang = 37.0 / (1e6 ** (126 / 128))
    assert np.allclose(r[:, j], f[:, j] * np.cos(ang) - f[:, j + 64] * np.sin(ang), rtol=2 ** -7, atol=1e-3)


def test_attention_decode_matches_numpy_and_variants_agree():
    rng = np.random.default_rng(8)
    H, KV, hd, pos, S = 8, 2, 64, 40, 64
    q = ol.f32_to_bf16(rng.normal(0, 1, size=(H, hd)).astype(np.float32))
    kc = ol.f32_to_bf16(rng.normal(0, 1, size=(S, KV * hd)).astype(np.float32))
    vc = ol.f32_to_bf16(rng.normal(0, 1, size=(S, KV * hd)).astype(np.float32))
    o32 = ol.bf16_to_f32(ol.attention_decode(q, kc, vc, pos, H, KV, hd, 0))
    o16 = ol.bf16_to_f32(ol.attention_decode(q, kc, vc, pos, H, KV, hd, 1))
    qf, kf, vf = ol.bf16_to_f32(q).astype(np.float64), ol.bf16_to_f32(kc).astype(np.float64), ol.bf16_to_f32(vc).astype(np.float64)
    want = np.zeros((H, hd))
    for h in range(H):
        kvh = h // (H // KV)
        s = kf[: pos + 1, kvh * hd:(kvh + 1) * hd] @ qf[h] / np.sqrt(hd)
        p = np.exp(s - s.max())
        p /= p.sum()
        want[h] = p @ vf[: pos + 1, kvh * hd:(kvh + 1) * hd]
    assert np.allclose(o32, want, rtol=2 ** -7, atol=2e-3)
    # the reference's two score precisions (bf16 neuron path / fp32 pipe path) agree within the logits tolerance
    assert np.allclose(o16, o32, rtol=3e-2, atol=3e-2)


def test_swiglu_add():
    rng = np.random.default_rng(9)
    g = ol.f32_to_bf16(rng.normal(0, 2, size=512).astype(np.float32))
    u = ol.f32_to_bf16(rng.normal(0, 2, size=512).astype(np.float32))
    gf, uf = ol.bf16_to_f32(g).astype(np.float64), ol.bf16_to_f32(u).astype(np.float64)
    assert np.allclose(ol.bf16_to_f32(ol.swiglu(g, u)), gf * uf / (1 + np.exp(-gf)), rtol=2 ** -7, atol=1e-6)
    assert np.allclose(ol.bf16_to_f32(ol.add(g, u)), gf + uf, rtol=2 ** -7, atol=1e-6)


# ------------------------------------------------------------------ whole model
def test_model_forward_is_deterministic_and_uses_cache():
    m = ol.OracleModel(n_layer=2, norm_sigma=0.1)
    toks = [(1000 + 37 * i) % 1024 for i in range(6)]
    logits = [m.forward(t, i) for i, t in enumerate(toks)]
    m.reset()
    logits2 = [m.forward(t, i) for i, t in enumerate(toks)]
    for a, b in zip(logits, logits2):
        assert np.array_equal(a, b)
    # position matters (KV cache + rope): same token at pos 0 vs later gives different logits
    assert not np.array_equal(logits[0], logits[-1])
    f = ol.bf16_to_f32(logits[-1])
    assert np.isfinite(f).all() and f.std() > 0
    # layer entry point == what forward does for layer 0 at pos 0
    m.reset()
    E = m.cfg.n_embd
    x0 = m.weight(0)[toks[0] * E:(toks[0] + 1) * E]
    x1 = m.layer(0, 0, x0)
    assert x1.shape == (E,) and not np.array_equal(x0, x1)


def test_model_weight_is_dequant_of_quantized_synthetic():
    m = ol.OracleModel(n_layer=1)
    c = m.cfg
    E, QD = c.n_embd, c.n_head * c.head_dim
    raw = ol.fill_normal(QD * E, ol.tensor_seed(c.seed, 16 + 1), c.sigma)
    data, gama = ol.quantize(raw, QD, E, 4, 128, ol.RTN_ASYM)
    assert np.array_equal(m.weight(17), ol.dequant(data, gama, QD, E, 4, 128, 0).reshape(-1))
    assert np.array_equal(m.weight(16), np.full(E, 0x3F80, dtype=np.uint16))  # norms are 1 (FIX_1)


# ---------------------------------------------------------------------------------------------- sampler (GoPT.cpp:614-630)
def _sampler_logits(seed, vocab=4096, sigma=2.0):
    rng = np.random.default_rng(seed)
    return ol.f32_to_bf16((rng.standard_normal(vocab) * sigma).astype(np.float32))


def test_sampler_port_greedy_and_distribution():
    lg = _sampler_logits(1)
    f = ol.bf16_to_f32(lg)
    st = [42]
    assert ol.sample(lg, 0.0, 50, 0.95, st)[0] == int(np.argmax(f)) and st[0] == 42  # temperature 0: argmax, generator untouched
    assert ol.sample(lg, 0.7, 1, 0.95, st)[0] == int(np.argmax(f))                    # top_k 1: argmax
    # top_p -> 0 keeps only the best candidate; the draw is then always the arg-max
    for _ in range(8):
        tok, npick = ol.sample(lg, 0.7, 50, 1e-6, st)
        assert tok == int(np.argmax(f)) and npick == 1
    # the empirical distribution over many draws follows softmax(top-50 / T) restricted to the top-p prefix
    T, k, top_p = 0.8, 50, 0.9
    order = np.lexsort((np.arange(f.size), -f))[:k]
    p = np.exp((f[order] - f[order[0]]) / T)
    p /= p.sum()
    keep = int(np.argmax(np.cumsum(p) > top_p)) + 1
    st = [12345]
    draws = [ol.sample(lg, T, k, top_p, st) for _ in range(4000)]
    assert all(n == keep for _, n in draws)
    counts = np.array([sum(1 for t, _ in draws if t == int(order[i])) for i in range(keep)], dtype=np.float64)
    assert counts.sum() == len(draws)  # never outside the kept prefix
    expect = p[:keep] / p[:keep].sum() * len(draws)
    assert np.all(np.abs(counts - expect) <= 5 * np.sqrt(expect) + 5)


def test_sampler_port_generator_is_xorshift64star():
    # random_u32 (GoPT.cpp:594-599) by hand
    s = 42
    s ^= s >> 12
    s ^= (s << 25) & 0xFFFFFFFFFFFFFFFF
    s ^= s >> 27
    st = [42]
    ol.sample(_sampler_logits(2), 0.7, 10, 1.0, st)
    assert st[0] == s


def test_sampler_port_reference_heap_selection():
    # selection 1 = TOPK_heap::Select as written (GoPT.cpp:667-700): candidates are indices 0 .. k-2 plus the first maximum of the rest
    lg = _sampler_logits(3)
    f = ol.bf16_to_f32(lg)
    k = 8
    allowed = set(range(k - 1)) | {int(k - 1 + np.argmax(f[k - 1:]))}
    st = [7]
    assert all(ol.sample(lg, 5.0, k, 1.0, st, selection=1)[0] in allowed for _ in range(200))


# ---------------------------------------------------------------------------------------------- NormalFloat4 (QUANT_MODE::RTNf)
def test_nf4_port_layout_codebook_and_nearest_code():
    rows, cols = 6, 64
    w = ol.fill_normal(rows * cols, 11, 0.02).reshape(rows, cols).copy()
    w[3, :] = 0  # an all-zero row: scale = 1, codebook = the table itself, every code = 7 (0.0)
    data, gama = ol.nf4_quantize(w, rows, cols)
    f = ol.bf16_to_f32(w).reshape(rows, cols)
    table = np.array([-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635, -0.18477343022823334,
                      -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725, 0.24611230194568634, 0.33791524171829224,
                      0.44070982933044434, 0.5626170039176941, 0.7229568362236023, 1.0], dtype=np.float32)  # NF4_LUT, g_float.hpp:543-558
    assert np.all(gama[:rows + cols] == 0)
    for r in range(rows):
        amax = np.float32(np.abs(f[r]).max())
        scale = np.float32(1.0) / amax if amax > 0 else np.float32(1.0)
        cb = (table / scale).astype(np.float32)
        assert np.array_equal(gama[rows + cols + 16 * r: rows + cols + 16 * (r + 1)], ol.f32_to_bf16(cb))  # the LUT is bf16(codebook)
        codes = np.argmin(np.abs(f[r][:, None] - cb[None, :]), axis=1)                                      # first minimum, fp32 codebook
        byts = data[r * cols // 2:(r + 1) * cols // 2]
        assert np.array_equal(byts >> 4, codes[0::2]) and np.array_equal(byts & 15, codes[1::2])            # even element = high nibble
    assert np.all(data[3 * cols // 2:4 * cols // 2] == 0x77)
    deq = ol.nf4_dequant(data, gama, rows, cols)
    lut = gama[rows + cols:].reshape(rows, 16)
    for r in range(rows):
        byts = data[r * cols // 2:(r + 1) * cols // 2]
        assert np.array_equal(deq[r, 0::2], lut[r][byts >> 4]) and np.array_equal(deq[r, 1::2], lut[r][byts & 15])
    # the extreme of every row is reproduced exactly (code 0 or 15 = -+abs_max)
    fd = ol.bf16_to_f32(deq).reshape(rows, cols)
    for r in (0, 1, 2, 4, 5):
        k = int(np.argmax(np.abs(f[r])))
        assert fd[r, k] == f[r, k]


# ---------------------------------------------------------------------------------------------- vendor AWQ layout (CU_Q42X_awq)
def test_awq_port_nibble_order_and_arithmetic():
    M, N = 256, 32  # [in_features][out_features]
    rng = np.random.default_rng(4)
    qw = rng.integers(0, 2 ** 32, size=M * N // 8, dtype=np.uint64).astype(np.uint32)
    qz = rng.integers(0, 2 ** 32, size=M // 128 * N // 8, dtype=np.uint64).astype(np.uint32)
    sc = (rng.uniform(0.001, 0.02, size=M // 128 * N)).astype(np.float16)
    got = ol.bf16_to_f32(ol.awq_dequant(qw, qz, sc.view(np.uint16), M, N)).reshape(M, N)
    order = [0, 4, 1, 5, 2, 6, 3, 7]  # AWQ_REVERSE_ORDER, packedN.cuh:110
    q = np.stack([(qw.reshape(M, N // 8) >> (4 * order[k])) & 15 for k in range(8)], axis=-1).reshape(M, N).astype(np.int32)
    z = np.stack([(qz.reshape(M // 128, N // 8) >> (4 * order[k])) & 15 for k in range(8)], axis=-1).reshape(M // 128, N).astype(np.int32)
    want = (q - np.repeat(z, 128, axis=0)).astype(np.float32) * np.repeat(sc.reshape(M // 128, N).astype(np.float32), 128, axis=0)
    assert np.array_equal(got, ol.bf16_to_f32(ol.f32_to_bf16(want)).reshape(M, N))
    # the test-side packer round-trips within half a step
    w = ol.fill_normal(M * N, 9, 0.02)
    pw, pz, ps = ol.awq_pack(w, M, N)
    back = ol.bf16_to_f32(ol.awq_dequant(pw, pz, ps, M, N)).reshape(M, N)
    step = np.repeat(ps.view(np.float16).astype(np.float32).reshape(M // 128, N), 128, axis=0)
    assert np.all(np.abs(back - ol.bf16_to_f32(w).reshape(M, N)) <= 0.51 * step + 1e-4)


def test_awq_port_against_the_reference_python_unpack():
    # tests/golden/awq_ref_py.npz: the reference's OWN unpack_awq / reverse_awq_order / Dequant_1 (src/Python/test_awq.py:33-134) run on seeded
    # arrays (tests/golden/make_awq_golden.py).  Case c has unit scales, so the weights ARE the code differences q - z: bit-exact, which pins the
    # nibble order of both qweight and qzeros.  Cases a / b have real scales: the Python path multiplies in fp16 and then rounds to bf16, the CUDA
    # kernel this port follows (CU_Q42X_awq, quantizer.cu:132-156) multiplies in fp32 -- equal except where the fp16 product is inexact, and then by
    # one bf16 ulp at most.
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "awq_ref_py.npz"))
    for tag in "abc":
        qw, qz, sc = g[tag + "_qweight"].view(np.uint32), g[tag + "_qzeros"].view(np.uint32), g[tag + "_scales"]
        IC, OC = qw.shape[0], sc.shape[1]
        got = ol.awq_dequant(qw.reshape(-1), qz.reshape(-1), sc.reshape(-1), IC, OC)
        want = g[tag + "_deq_bf16"]
        if tag == "c":
            assert np.array_equal(got, want)
            codes = g["c_codes"].astype(np.int32) - np.repeat(g["c_zeros"].astype(np.int32), 128, axis=0)
            assert np.array_equal(ol.bf16_to_f32(got), codes.astype(np.float32))
            continue
        gf, wf = ol.bf16_to_f32(got), ol.bf16_to_f32(want)
        differ = gf != wf
        assert differ.mean() < 0.10, differ.mean()
        ulp = np.abs(wf) * 2.0 ** -7
        assert np.all(np.abs(gf - wf)[differ] <= ulp[differ] * 1.001)
        # recomputing the Python arithmetic from the golden codes gives the golden values exactly: the difference above is the rounding, not the layout
        q = g[tag + "_codes"].astype(np.float16) - np.repeat(g[tag + "_zeros"].astype(np.float16), 128, axis=0)
        prod16 = (q * np.repeat(sc.view(np.float16), 128, axis=0)).astype(np.float16)
        assert np.array_equal(ol.f32_to_bf16(prod16.astype(np.float32)).reshape(IC, OC), want)


# ---------------------------------------------------------------------------------------------- the reference's own CPU packers
def _restated(w, rows, cols, bits, mode):
    if mode == ol.NF4:
        return ol.nf4_quantize(w, rows, cols)
    return ol.quantize(w, rows, cols, bits, 128, mode)


def test_packers_against_golden_bytes_of_the_reference_cpu_packers():
    # tests/golden/refcpu_quant.npz: GeQuant::RTN_x / YinYang / RT_NormalF of the reference itself (compiled from src/Tensor/GeQuant.cpp,
    # tests/golden/make_golden_refcpu.py) on seeded inputs.  Packed bytes and the written part of gama (ZERO / STEP, or the per-row codebooks) must be
    # equal bit for bit; R_SCALE / C_SCALE are never written by either (NO_NORMAL) and are not compared.
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refcpu_quant.npz"))
    tags = sorted(k[:-5] for k in g.files if k.endswith("_meta"))
    assert len(tags) == 9
    for tag in tags:
        rows, cols, bits, mode, seed, qb = (int(x) for x in g[tag + "_meta"])
        w = ol.fill_normal(rows * cols, seed, float(g[tag + "_sigma"][0]))
        data, gama = _restated(w, rows, cols, bits, mode)
        assert np.array_equal(data, g[tag + "_data"]), tag
        assert np.array_equal(gama[rows + cols:], g[tag + "_gama"][rows + cols:]), tag
        if mode != ol.NF4:
            assert ol.qrange(bits, mode)[2] == qb, tag


def test_packers_against_the_reference_cpu_packers_live():
    # the same comparison against the compiled reference code itself on more shapes, when oracle/_ref/libkoifish_refcpu.so exists (it is built where
    # /root/reference is present and travels with the repository snapshot)
    if ol.refcpu() is None:
        pytest.skip("oracle/_ref/libkoifish_refcpu.so not built (reference tree absent at build time)")
    rng = np.random.default_rng(5)
    for rows, cols in ((8, 128), (40, 384), (64, 2048), (3, 5120)):
        for bits, mode in ((4, ol.RTN_ASYM), (4, ol.RTN_SYM), (2, ol.RTN_ASYM), (2, ol.YYANG), (1, ol.YYANG), (4, ol.NF4)):
            for sigma in (0.02, 3.0):
                w = ol.fill_normal(rows * cols, int(rng.integers(1, 1 << 30)), sigma)
                if rng.random() < 0.5:  # a few exact zeros and one outlier per call
                    w = w.copy()
                    w[rng.integers(0, w.size, 16)] = 0
                    w[int(rng.integers(0, w.size))] = ol.f32_to_bf16(np.array([sigma * 40], dtype=np.float32))[0]
                want_d, want_g, qb = ol.refcpu_quantize(w, rows, cols, bits, 128, mode)
                got_d, got_g = _restated(w, rows, cols, bits, mode)
                assert np.array_equal(got_d, want_d), (rows, cols, bits, mode, sigma)
                assert np.array_equal(got_g[rows + cols:], want_g[rows + cols:]), (rows, cols, bits, mode, sigma)


def test_sampler_port_against_the_reference_sampler_compiled_from_its_tree():
    # GeneratOnPrompt::Sample = LogitsInfo::TopK / UpdateLogits / TopP / Qu_FlipCoin of src/Manifold/GoPT.cpp compiled where it lies
    # (oracle/_ref/libkoifish_refcpu.so) against kfo_sample(selection = 1), the port of that code AS WRITTEN (its heap orders indices).  Same coin,
    # same cut: the generator state and the nucleus size must be equal, and the token equal -- or, where the k candidates hold equal logits, a
    # candidate with the same logit (the reference's std::sort on `a > b` leaves the order inside a tie unspecified).  top_p >= 1 is left out: the
    # reference then never sets nPick and Qu_FlipCoin reads picks[-2].
    if ol.refcpu() is None:
        pytest.skip("oracle/_ref/libkoifish_refcpu.so not built (reference tree absent at build time)")
    rng = np.random.default_rng(3)
    total = same = 0
    for vocab in (1024, 4096, 151936):
        for trial in range(12):
            lg = ol.f32_to_bf16((rng.standard_normal(vocab) * rng.choice([0.5, 2.0, 6.0])).astype(np.float32))
            for T, k, p in ((0.6, 50, 0.95), (0.9, 20, 0.8), (1.5, 100, 0.99), (0.3, 8, 0.5), (0.6, 2, 0.95), (2.0, 300, 0.9)):
                s1, s2 = [1234 + trial], [1234 + trial]
                for _ in range(3):
                    tok_r, n_r = ol.refcpu_sample(lg, T, k, p, s1)
                    tok_o, n_o = ol.sample(lg, T, k, p, s2, selection=1)
                    assert (n_r, s1) == (n_o, s2), (vocab, T, k, p)
                    assert tok_r == tok_o or lg[tok_r] == lg[tok_o], (vocab, T, k, p, tok_r, tok_o)
                    total += 1
                    same += tok_r == tok_o
    assert same >= 0.97 * total
