"""GPU parity tests (run with -m gpu on the B200 box): every kernel behind the C ABI against the CPU oracle on the same seeded
inputs.  Integer / byte / bf16-dequant work is bit-exact; dot products are compared within a tolerance written in each test."""
import os

import numpy as np
import pytest

import koifish_b200 as kf
import oracle_lib as ol

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "packq_ref.npz"))

KF_TYPE = {(4, ol.RTN_ASYM): kf.KF_T_Q4, (4, ol.RTN_SYM): kf.KF_T_Q4, (2, ol.RTN_ASYM): kf.KF_T_Q2, (2, ol.RTN_SYM): kf.KF_T_Q2,
           (2, ol.YYANG): kf.KF_T_SIGN, (1, ol.YYANG): kf.KF_T_BINARY}
ALL_QUANT = [(4, ol.RTN_ASYM), (4, ol.RTN_SYM), (2, ol.RTN_ASYM), (2, ol.YYANG), (1, ol.YYANG)]


@pytest.fixture(scope="module")
def ctx():
    c = kf.Context(0)
    yield c
    c.close()


@pytest.fixture(autouse=True)
def _skinny_kernel_for_gemv_tests(request, ctx):
    """tests named test_gemv_* / test_linear_* pin the mma.sync skinny kernel at every token count up to 64; everything else runs with
    the product's routing (tcgen05 kernel from 9 tokens, bf16 weights always).  The tests of this module check the BIT-EXACT arithmetic
    (ctx knob gemv_exact = 1: the reference's per-weight bf16 dequant inside the matmul) unless their name says *_fast_*: those run the
    product's default decode arithmetic (fp16 codes + affine map on the group sums) against its stated tolerance."""
    name = request.node.name
    ctx.set_int("gemv_exact", 0 if "_fast_" in name else 1)
    if name.startswith("test_gemv_") or name.startswith("test_linear_"):
        ctx.set_int("tc_min_m", 0)
        yield
        ctx.set_int("tc_min_m", -1)
    else:
        yield
    ctx.set_int("gemv_exact", 0)


def oracle_qtensor(ctx, rows, cols, bits, mode, seed, group=128, sigma=0.02):
    """weights quantised by the ORACLE packer, uploaded as-is: the kernels must read the reference's byte layout"""
    w = ol.fill_normal(rows * cols, seed, sigma)
    data, gama = ol.quantize(w, rows, cols, bits, group, mode)
    _, _, qbias = ol.qrange(bits, mode)
    t = kf.QTensor.from_packed(ctx, data, gama, rows, cols, KF_TYPE[(bits, mode)], group, qbias)
    wdq = ol.dequant(data, gama, rows, cols, bits, group, qbias)
    return t, wdq


def make_weight(ctx, kind, rows, cols, seed):
    """returns (QTensor on device, dequantised bf16 weights [rows, cols] from the oracle)"""
    if kind == "bf16":
        w = ol.fill_normal(rows * cols, seed, 0.02)
        return kf.QTensor.from_packed(ctx, w.view(np.uint8), None, rows, cols, kf.KF_T_BF16), w.reshape(rows, cols)
    if kind == "f8":
        w = ol.fill_normal(rows * cols, seed, 0.02)
        b = ol.f8_encode(w)
        return kf.QTensor.from_packed(ctx, b, None, rows, cols, kf.KF_T_F8E5M2), ol.f8_decode(b).reshape(rows, cols)
    bits, mode = kind
    return oracle_qtensor(ctx, rows, cols, bits, mode, seed)


WEIGHT_KINDS = ["bf16", "f8"] + ALL_QUANT


def rand_bf16(rng, shape, scale=1.0):
    return ol.f32_to_bf16((rng.standard_normal(shape) * scale).astype(np.float32))


# ---------------------------------------------------------------------------------------------- generator / packer / dequant
def test_fill_normal_bit_exact(ctx):
    n = 1 << 18
    for seed, sigma, mean in ((42, 0.02, 0.0), (7, 0.1, 1.0), (123456789, 1.0, -0.5)):
        g = kf.fill_normal(ctx, n, seed, sigma, mean).numpy(np.uint16)
        assert np.array_equal(g, ol.fill_normal(n, seed, sigma, mean))
    # 2-D window == slice of the full tensor (tensor-parallel shards)
    full = ol.fill_normal(96 * 640, 5, 0.02).reshape(96, 640)
    win = kf.fill_normal_2d(ctx, 32, 256, 640, 16, 128, 5, 0.02).numpy(np.uint16, (32, 256))
    assert np.array_equal(win, full[16:48, 128:384])


@pytest.mark.parametrize("bits,mode", ALL_QUANT)
@pytest.mark.parametrize("rows,cols,group", [(48, 512, 128), (16, 1024, 256), (128, 384, 128)])
def test_quantize_bit_exact_vs_oracle_packer(ctx, bits, mode, rows, cols, group):
    w = ol.fill_normal(rows * cols, 100 + bits, 0.02)
    data_o, gama_o = ol.quantize(w, rows, cols, bits, group, mode)
    t = kf.quantize(ctx, ctx.array(w), rows, cols, KF_TYPE[(bits, mode)], group, mode)
    assert t.qbias == ol.qrange(bits, mode)[2]
    assert np.array_equal(t.gama_numpy(), gama_o)
    assert np.array_equal(t.data_numpy(), data_o)


def test_quantize_edge_cases(ctx):
    # a constant group (step == 0) and exact-tie values; the oracle documents its step==0 convention
    rows, cols = 16, 256
    w = ol.fill_normal(rows * cols, 9, 0.02).reshape(rows, cols)
    w[0, :128] = 0x3C00          # constant group
    w[1, :] = 0                  # all zeros
    w[2, :128] = ol.f32_to_bf16(np.linspace(-1, 1, 128).astype(np.float32))
    for bits, mode in ALL_QUANT:
        data_o, gama_o = ol.quantize(w.reshape(-1), rows, cols, bits, 128, mode)
        t = kf.quantize(ctx, ctx.array(w), rows, cols, KF_TYPE[(bits, mode)], 128, mode)
        assert np.array_equal(t.data_numpy(), data_o), (bits, mode)
        assert np.array_equal(t.gama_numpy(), gama_o), (bits, mode)
    # ragged shapes are rejected, not silently mis-packed
    with pytest.raises(kf.KoifishError):
        kf.quantize(ctx, ctx.array(w), rows, cols, kf.KF_T_Q4, 96, ol.RTN_ASYM)
    with pytest.raises(kf.KoifishError):
        kf.quantize(ctx, ctx.array(w), rows, cols, kf.KF_T_BINARY, 128, ol.RTN_ASYM)


@pytest.mark.parametrize("bits", [4, 2, 1])
def test_unpack_layout_against_reference_golden_bytes(ctx, bits):
    # bytes produced by the REFERENCE's PACK_*to128_ macros (tests/golden/make_golden.py); step = 1, zero = 0 => dequant == code
    codes, data = GOLD[f"codes{bits}"], GOLD[f"bytes{bits}"]
    n = codes.size
    rows, cols = n // 128, 128
    gama = np.zeros(rows + cols + 2 * rows, dtype=np.uint16)
    gama[rows + cols + rows:] = 0x3F80
    tp = {4: kf.KF_T_Q4, 2: kf.KF_T_Q2, 1: kf.KF_T_BINARY}[bits]
    t = kf.QTensor.from_packed(ctx, data, gama, rows, cols, tp, 128, 0)
    got = ol.bf16_to_f32(kf.dequant(ctx, t).numpy(np.uint16)).astype(np.int32)
    assert np.array_equal(got, codes)


@pytest.mark.parametrize("bits,mode", ALL_QUANT)
def test_dequant_bit_exact(ctx, bits, mode):
    rows, cols = 64, 1024
    t, wdq = oracle_qtensor(ctx, rows, cols, bits, mode, 300 + bits)
    got = kf.dequant(ctx, t).numpy(np.uint16, (rows, cols))
    assert np.array_equal(got, wdq)


def test_f8_bit_exact(ctx):
    w = ol.fill_normal(1 << 16, 77, 0.02)
    w[:8] = [0, 0x8000, 0x3F80, 0xBF80, 0x0001, 0x3380, 0x3800, 0x4700]  # zeros, +-1, tiny / fp16-subnormal range, large
    t = kf.quantize(ctx, ctx.array(w), 256, 256, kf.KF_T_F8E5M2, 0, 0)
    assert np.array_equal(t.data_numpy(), ol.f8_encode(w))
    assert np.array_equal(kf.dequant(ctx, t).numpy(np.uint16), ol.f8_decode(ol.f8_encode(w)))


# ---------------------------------------------------------------------------------------------- fused dequant GEMV
@pytest.mark.parametrize("kind", WEIGHT_KINDS, ids=str)
@pytest.mark.parametrize("M", [1, 3, 8])
def test_gemv_onehot_reproduces_dequantised_weights_bit_exact(ctx, kind, M):
    # x = e_k  =>  y[n] = w[n][k] exactly: pins the in-kernel unpack + dequant (and the k-permutation) against the oracle
    rows, cols = 160, 512
    t, wdq = make_weight(ctx, kind, rows, cols, 900)
    rng = np.random.default_rng(1)
    for trial in range(6):
        ks = rng.integers(0, cols, size=M)
        x = np.zeros((M, cols), dtype=np.uint16)
        x[np.arange(M), ks] = 0x3F80
        y = kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16, (M, rows))
        for m in range(M):
            # compare as numbers: a weight of -0.0 (E5M2 truncation of a tiny negative value) sums to +0.0 with the other zero terms
            assert np.array_equal(ol.bf16_to_f32(y[m]), ol.bf16_to_f32(wdq[:, ks[m]])), (kind, M, trial, m)


def _check_linear(y_bits, w_bits, x_bits, M, N, K, noise=2e-3):
    ref = ol.linear_f32(w_bits, x_bits, M, N, K)
    got = ol.bf16_to_f32(y_bits).reshape(M, N)
    # tolerance: one bf16 rounding of the result (2^-8 relative, half-ulp is 2^-9) + fp32 accumulation-order noise
    tol = np.abs(ref) * 2.0 ** -8 + noise * np.sqrt(np.mean(ref ** 2, axis=1, keepdims=True))
    bad = np.abs(got - ref) > tol
    assert not bad.any(), "max err %g at %s" % (np.abs(got - ref).max(), np.argwhere(bad)[:4])


@pytest.mark.parametrize("kind", WEIGHT_KINDS, ids=str)
@pytest.mark.parametrize("M,N,K", [(1, 256, 1024), (1, 1040, 4096), (2, 128, 512), (8, 384, 2048), (16, 256, 1024), (33, 128, 1024), (64, 256, 512)])
def test_gemv_matches_oracle(ctx, kind, M, N, K):
    t, wdq = make_weight(ctx, kind, N, K, 1000 + M)
    x = rand_bf16(np.random.default_rng(M * 7 + N), (M, K))
    y = kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16)
    _check_linear(y, wdq, x, M, N, K)


# ---- the product's default decode arithmetic for 4-bit weights (MODE_FAST, gemv.cu): fp16 codes into the tensor cores, the group's
#      affine map applied to the fp32 group sums, i.e. y = sum_g (step_g * sum_k c x - zero_g * sum_k x) with NO rounding of the individual
#      dequantised weights to bf16.  The reference rounds each weight (RN_bf16(step * c - zero), relative error uniform in +-2^-9, rms
#      1.5e-3 of the weight), so the two differ by a random walk of those roundings: rms 1.5e-3 of the rms output, whatever K is.  The
#      gate is 5 sigma of that noise (FAST_NOISE) on top of the bf16 rounding of the result; kf_dequant / the tcgen05 GEMM / gemv_exact = 1
#      stay bit-faithful to the reference's weights.
FAST_NOISE = 8e-3


FAST_KINDS = [(4, ol.RTN_ASYM), (4, ol.RTN_SYM), (2, ol.RTN_ASYM), (2, ol.YYANG), (1, ol.YYANG)]


def _fast_noise(kind):
    # yyang ternary / binary weights are step * k with k in {-1, 0, 1}: exact in bf16, so the fast arithmetic IS the reference's and the gate
    # is the bit-faithful mode's; the RTN kinds differ by the reference's per-weight rounding noise
    return 2e-3 if kind[1] == ol.YYANG else FAST_NOISE


@pytest.mark.parametrize("kind", FAST_KINDS, ids=str)
@pytest.mark.parametrize("M,N,K", [(1, 256, 1024), (1, 1040, 4096), (2, 128, 512), (3, 5120, 2048), (8, 384, 2048), (16, 256, 1024), (64, 256, 512)])
def test_gemv_fast_matches_oracle(ctx, kind, M, N, K):
    t, wdq = make_weight(ctx, kind, N, K, 1000 + M)
    x = rand_bf16(np.random.default_rng(M * 7 + N), (M, K))
    y = kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16)
    _check_linear(y, wdq, x, M, N, K, noise=_fast_noise(kind))
    # ... and the noise really is that small on average: rms error <= 3e-3 of the rms output (1.5e-3 expected + the bf16 result rounding)
    ref = ol.linear_f32(wdq, x, M, N, K)
    got = ol.bf16_to_f32(y).reshape(M, N)
    assert np.sqrt(np.mean((got - ref) ** 2)) <= 3e-3 * np.sqrt(np.mean(ref ** 2))


@pytest.mark.parametrize("kind", FAST_KINDS, ids=str)
def test_gemv_fast_one_signed_activations(ctx, kind):
    # all-positive activations with a large mean (mean / std = 5): the group sums of x are as large as they get.  Two things are pinned:
    # (a) the subnormal code operands lose nothing visible (the tensor cores keep 24 bits below a product's NOMINAL exponent, which for a
    #     subnormal sits up to 2^10 above a small code: the yyang kinds, which have no rounding difference to hide behind, still pass the
    #     bit-faithful gate);
    # (b) for the RTN kinds the difference to the reference is LARGER here than for zero-mean activations: the reference's rounding error
    #     of a weight is shared by every element of the group that carries the same code (16 distinct weights per group), so against
    #     same-signed activations it does not average out over k -- sigma = 4.7e-3 of the rms output in this set-up (4-bit asym: 16 codes x
    #     32 groups), gate 5 sigma.  The fast arithmetic is the one without that error.
    M, N, K = 2, 384, 4096
    t, wdq = make_weight(ctx, kind, N, K, 77)
    rng = np.random.default_rng(8)
    x = ol.f32_to_bf16((np.abs(rng.standard_normal((M, K))) + 3.0).astype(np.float32))
    y = kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16)
    _check_linear(y, wdq, x, M, N, K, noise=2e-3 if kind[1] == ol.YYANG else 2.5e-2)


@pytest.mark.parametrize("kind", [(2, ol.YYANG), (1, ol.YYANG), (2, ol.RTN_ASYM)], ids=str)
def test_gemv_fast_onehot_low_bit(ctx, kind):
    # one-hot rows through the fast arithmetic: y = step * k - zero evaluated in fp32 and rounded once -- the dequantised weight itself
    # for the yyang kinds (exact in bf16), within one bf16 ulp of it for 2-bit RTN; pins the field / activation-slot pairing of every bit position
    rows, cols = 160, 1024
    t, wdq = make_weight(ctx, kind, rows, cols, 901)
    for k0 in (0, 37, 128 + 5, 511, 1023):
        x = np.zeros((8, cols), dtype=np.uint16)
        ks = [(k0 + 3 * m) % cols for m in range(8)]
        x[np.arange(8), ks] = 0x3F80
        y = kf.linear(ctx, t, ctx.array(x), 8).numpy(np.uint16).reshape(8, rows)
        for m in range(8):
            a, b = ol.bf16_to_f32(y[m]), ol.bf16_to_f32(wdq[:, ks[m]])
            if kind[1] == ol.YYANG:
                assert np.array_equal(y[m], wdq[:, ks[m]]), (k0, m)
            else:
                assert np.all(np.abs(a - b) <= np.maximum(np.abs(a), np.abs(b)) * 2.0 ** -7 + 1e-12)


def test_gemv_fast_wide_dynamic_range_and_onehot(ctx):
    """activations spanning 2^40 inside one row (outliers next to tiny values: the per-group power-of-two scale of the fp16 staging), and
    one-hot rows: y = RN_bf16(step * k - zero), i.e. the reference's fused dequant of that weight (deq_fma = 1) up to the accumulation
    order of the two terms -- at most one bf16 ulp from the dequantised weight"""
    rows, cols = 160, 1024
    t, wdq = make_weight(ctx, (4, ol.RTN_ASYM), rows, cols, 900)
    rng = np.random.default_rng(5)
    for scale in (1.0, 3e4, 1e-6):
        x = rng.standard_normal((4, cols)).astype(np.float32) * scale
        x[:, ::37] *= 1e4
        x[:, 3::29] *= 1e-8
        xb = ol.f32_to_bf16(x)
        y = kf.linear(ctx, t, ctx.array(xb), 4).numpy(np.uint16)
        _check_linear(y, wdq, xb, 4, rows, cols, noise=FAST_NOISE)
    M = 3
    ks = rng.integers(0, cols, size=M)
    x = np.zeros((M, cols), dtype=np.uint16)
    x[np.arange(M), ks] = 0x3F80
    y = kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16, (M, rows))
    for m in range(M):
        a, b = ol.bf16_to_f32(y[m]), ol.bf16_to_f32(wdq[:, ks[m]])
        assert np.all(np.abs(a - b) <= np.maximum(np.abs(a), np.abs(b)) * 2.0 ** -7 + 1e-12)


def test_gemv_fast_epilogues_and_fused_norm(ctx):
    M, N, K = 2, 512, 2048
    rng = np.random.default_rng(17)
    wg, gq = make_weight(ctx, (4, ol.RTN_ASYM), N, K, 31)
    wu, uq = make_weight(ctx, (4, ol.RTN_ASYM), N, K, 32)
    x = rand_bf16(rng, (M, K))
    nw = ol.f32_to_bf16((1.0 + 0.1 * rng.standard_normal(K)).astype(np.float32))
    xd, nwd = ctx.array(x), ctx.array(nw)
    xn = ol.rmsnorm(x, nw, M, K)
    # fused RMSNorm + gate/up + SwiGLU vs the oracle chain
    got = ol.bf16_to_f32(kf.rmsnorm_linear(ctx, [wg, wu], xd, nwd, M, 1e-6, swiglu=True).numpy(np.uint16)).reshape(M, N)
    g = ol.linear(gq, xn, M, N, K)
    u = ol.linear(uq, xn, M, N, K)
    want = ol.bf16_to_f32(ol.swiglu(g, u)).reshape(M, N)
    assert np.abs(got - want).max() <= 2e-2 * np.abs(want).max()
    # residual epilogue == plain output + residual, with the single-GPU rounding points
    res = rand_bf16(rng, (M, N))
    plain = kf.linear(ctx, wg, xd, M).numpy(np.uint16)
    withres = kf.linear(ctx, wg, xd, M, kf.KF_EPI_RESIDUAL, ctx.array(res)).numpy(np.uint16)
    assert np.array_equal(withres.reshape(M, N), ol.add(res, plain).reshape(M, N))


@pytest.mark.parametrize("splitk", [1, 2, 3, 7])
def test_gemv_splitk_deterministic_and_correct(ctx, splitk):
    M, N, K = 4, 384, 4096
    t, wdq = make_weight(ctx, (4, ol.RTN_ASYM), N, K, 55)
    x = rand_bf16(np.random.default_rng(3), (M, K))
    xd = ctx.array(x)
    ctx.set_int("gemv_splitk", splitk)
    try:
        y1 = kf.linear(ctx, t, xd, M).numpy(np.uint16)
        y2 = kf.linear(ctx, t, xd, M).numpy(np.uint16)
    finally:
        ctx.set_int("gemv_splitk", 0)
    assert np.array_equal(y1, y2)  # fixed-order reduction: bit-reproducible run to run
    _check_linear(y1, wdq, x, M, N, K)


def test_gemv_group_256_and_large_group(ctx):
    M, N, K = 2, 128, 2048
    for group in (256, 512):
        w = ol.fill_normal(N * K, 66, 0.02)
        data, gama = ol.quantize(w, N, K, 4, group, ol.RTN_ASYM)
        t = kf.QTensor.from_packed(ctx, data, gama, N, K, kf.KF_T_Q4, group, 0)
        wdq = ol.dequant(data, gama, N, K, 4, group, 0)
        x = rand_bf16(np.random.default_rng(group), (M, K))
        _check_linear(kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16), wdq, x, M, N, K)


def test_gemv_epilogues(ctx):
    M, N, K = 3, 256, 1024
    rng = np.random.default_rng(11)
    t, wdq = make_weight(ctx, (4, ol.RTN_ASYM), N, K, 70)
    x = rand_bf16(rng, (M, K))
    xd = ctx.array(x)
    plain = kf.linear(ctx, t, xd, M).numpy(np.uint16, (M, N))
    # residual: out = RN(res + RN_bf16(acc))  (reference: bf16 GEMM output, then CU_add3)
    res = rand_bf16(rng, (M, N))
    got = kf.linear(ctx, t, xd, M, kf.KF_EPI_RESIDUAL, ctx.array(res)).numpy(np.uint16, (M, N))
    assert np.array_equal(got, ol.add(res, plain).reshape(M, N))
    # in-place residual (y aliases residual), as the runtime uses it
    buf = ctx.array(res)
    kf.linear(ctx, t, xd, M, kf.KF_EPI_RESIDUAL, buf, out=buf)
    assert np.array_equal(buf.numpy(np.uint16, (M, N)), got)
    # fp32 partial sums (tensor-parallel epilogue): rounding them gives the plain result
    f32 = kf.linear(ctx, t, xd, M, kf.KF_EPI_F32).numpy(np.float32, (M, N))
    assert np.array_equal(ol.f32_to_bf16(f32), plain)


def test_linear_multi_equals_separate_calls(ctx):
    M, K = 5, 1024
    ws = [make_weight(ctx, (4, ol.RTN_ASYM), n, K, 80 + i)[0] for i, n in enumerate((512, 128, 128))]
    xd = ctx.array(rand_bf16(np.random.default_rng(12), (M, K)))
    outs = kf.linear_multi(ctx, ws, xd, M)
    for w, o in zip(ws, outs):
        assert np.array_equal(o.numpy(np.uint16), kf.linear(ctx, w, xd, M).numpy(np.uint16))


@pytest.mark.parametrize("kind", [(4, ol.RTN_ASYM), (2, ol.YYANG), "f8"], ids=str)
def test_linear_swiglu_fused(ctx, kind):
    M, N, K = 4, 320, 1024
    wg, _ = make_weight(ctx, kind, N, K, 90)
    wu, _ = make_weight(ctx, kind, N, K, 91)
    xd = ctx.array(rand_bf16(np.random.default_rng(13), (M, K)))
    g = kf.linear(ctx, wg, xd, M).numpy(np.uint16)
    u = kf.linear(ctx, wu, xd, M).numpy(np.uint16)
    fused = kf.linear_swiglu(ctx, wg, wu, xd, M).numpy(np.uint16)
    want = ol.swiglu(g, u)
    # same op order as the reference (bf16 gate/up, fp32 SwiGLU); expf may differ by an ulp between libm and CUDA
    diff = np.abs(ol.bf16_to_f32(fused) - ol.bf16_to_f32(want))
    assert (diff <= np.abs(ol.bf16_to_f32(want)) * 2.0 ** -7 + 1e-30).all()
    assert (fused == want).mean() > 0.999


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("kind", WEIGHT_KINDS, ids=str)
def test_gemv_tile_variants(ctx, variant, kind):
    # 16 or 32 rows per warp (RT = 1 / 2), ragged row blocks (N = 400 is neither a multiple of 128 nor of 256), M = 1 and 6
    N, K = 400, 1024
    t, wdq = make_weight(ctx, kind, N, K, 321)
    ctx.set_int("gemv_variant", variant)
    try:
        for M in (1, 6):
            x = rand_bf16(np.random.default_rng(M + variant), (M, K))
            y = kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16)
            _check_linear(y, wdq, x, M, N, K)
            ks = np.arange(M) * 97 % K
            x1 = np.zeros((M, K), dtype=np.uint16)
            x1[np.arange(M), ks] = 0x3F80
            y1 = kf.linear(ctx, t, ctx.array(x1), M).numpy(np.uint16, (M, N))
            for m in range(M):
                assert np.array_equal(ol.bf16_to_f32(y1[m]), ol.bf16_to_f32(wdq[:, ks[m]]))
    finally:
        ctx.set_int("gemv_variant", 0)


@pytest.mark.parametrize("M", [1, 5, 20])
def test_rmsnorm_folded_into_linear_is_bit_identical_to_unfused(ctx, M):
    rng = np.random.default_rng(M)
    K = 2048
    x, nw = rand_bf16(rng, (M, K), 2.0), rand_bf16(rng, (K,), 0.5)
    xd, nwd = ctx.array(x), ctx.array(nw)
    xn = kf.rmsnorm(ctx, xd, nwd, M, K, 1e-6)
    ws = [make_weight(ctx, (4, ol.RTN_ASYM), n, K, 500 + i)[0] for i, n in enumerate((512, 128, 128))]
    want = kf.linear_multi(ctx, ws, xn, M)
    got = kf.rmsnorm_linear(ctx, ws, xd, nwd, M, 1e-6)
    for a, b in zip(got, want):
        assert np.array_equal(a.numpy(np.uint16), b.numpy(np.uint16))
    wg, wu = make_weight(ctx, (4, ol.RTN_ASYM), 640, K, 600)[0], make_weight(ctx, (4, ol.RTN_ASYM), 640, K, 601)[0]
    assert np.array_equal(kf.rmsnorm_linear(ctx, [wg, wu], xd, nwd, M, 1e-6, swiglu=True).numpy(np.uint16),
                          kf.linear_swiglu(ctx, wg, wu, xn, M).numpy(np.uint16))
    head = make_weight(ctx, "bf16", 1024, K, 700)[0]
    assert np.array_equal(kf.rmsnorm_linear(ctx, [head], xd, nwd, M, 1e-6)[0].numpy(np.uint16), kf.linear(ctx, head, xn, M).numpy(np.uint16))


# ---------------------------------------------------------------------------------------------- tcgen05 / TMEM dequant GEMM (M > 64)
@pytest.mark.parametrize("kind", WEIGHT_KINDS, ids=str)
@pytest.mark.parametrize("M,N,K", [(65, 128, 256), (128, 256, 512), (200, 384, 1024), (300, 128, 2048), (513, 144, 512)])
def test_gemm_tc_matches_oracle(ctx, kind, M, N, K):
    t, wdq = make_weight(ctx, kind, N, K, 2000 + M)
    x = rand_bf16(np.random.default_rng(M + N), (M, K))
    y = kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16)
    _check_linear(y, wdq, x, M, N, K)


@pytest.mark.parametrize("kind", WEIGHT_KINDS, ids=str)
def test_gemm_tc_onehot_reproduces_dequantised_weights(ctx, kind):
    # every token selects one column k: y[m][n] == w[n][k] exactly -> pins the dequant, the k permutation of A and B, the swizzled
    # shared-memory layout and the TMEM lane/column mapping of the tensor-core path
    M, N, K = 192, 256, 512
    t, wdq = make_weight(ctx, kind, N, K, 3000)
    ks = (np.arange(M) * 37 + 5) % K
    x = np.zeros((M, K), dtype=np.uint16)
    x[np.arange(M), ks] = 0x3F80
    y = kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16, (M, N))
    for m in range(M):
        assert np.array_equal(ol.bf16_to_f32(y[m]), ol.bf16_to_f32(wdq[:, ks[m]])), (kind, m)


def test_gemm_tc_equals_skinny_panels_and_epilogues(ctx):
    M, N, K = 160, 384, 1024
    rng = np.random.default_rng(77)
    t, wdq = make_weight(ctx, (4, ol.RTN_ASYM), N, K, 4000)
    x, res = rand_bf16(rng, (M, K)), rand_bf16(rng, (M, N))
    xd = ctx.array(x)
    y_tc = kf.linear(ctx, t, xd, M).numpy(np.uint16)
    ctx.set_int("tc_min_m", 0)  # same call through 64-token panels of the skinny kernel
    try:
        y_sk = kf.linear(ctx, t, xd, M).numpy(np.uint16)
    finally:
        ctx.set_int("tc_min_m", -1)
    d = np.abs(ol.bf16_to_f32(y_tc) - ol.bf16_to_f32(y_sk))
    assert (d <= np.abs(ol.bf16_to_f32(y_sk)) * 2.0 ** -7 + 1e-3).all() and (y_tc == y_sk).mean() > 0.97
    got = kf.linear(ctx, t, xd, M, kf.KF_EPI_RESIDUAL, ctx.array(res)).numpy(np.uint16)
    assert np.array_equal(got, ol.add(res, y_tc))
    f32 = kf.linear(ctx, t, xd, M, kf.KF_EPI_F32).numpy(np.float32)
    assert np.array_equal(ol.f32_to_bf16(f32), y_tc)
    # fused norm + multi / swiglu entry points take the tensor-core path for M > 64 as well
    nw = rand_bf16(rng, (K,), 0.5)
    xn = kf.rmsnorm(ctx, xd, ctx.array(nw), M, K, 1e-6)
    wg, wu = make_weight(ctx, (4, ol.RTN_ASYM), 256, K, 4001)[0], make_weight(ctx, (4, ol.RTN_ASYM), 256, K, 4002)[0]
    a = kf.rmsnorm_linear(ctx, [wg, wu], xd, ctx.array(nw), M, 1e-6, swiglu=True).numpy(np.uint16)
    g = kf.linear(ctx, wg, xn, M).numpy(np.uint16)
    u = kf.linear(ctx, wu, xn, M).numpy(np.uint16)
    want = ol.swiglu(g, u)
    assert (a == want).mean() > 0.999


def test_linear_rejects_bad_shapes(ctx):
    t, _ = make_weight(ctx, (4, ol.RTN_ASYM), 128, 512, 1)
    xd = ctx.array(np.zeros((1, 512), dtype=np.uint16))
    with pytest.raises(kf.KoifishError):
        kf.linear(ctx, t, xd, 0)
    bad = kf.QTensor.from_packed(ctx, np.zeros(24 * 512 // 2, dtype=np.uint8), np.zeros(24 + 512 + 2 * 96, dtype=np.uint16), 24, 512, kf.KF_T_Q4)
    with pytest.raises(kf.KoifishError):
        kf.linear(ctx, bad, xd, 1)  # rows not a multiple of 16


def test_gemv_full_size_properties_qwen3_32b_down_proj(ctx):
    # BASELINE full size: K = 25600, N = 5120, 4-bit.  Size-independent properties: one-hot columns reproduce the dequantised
    # weights bit-exactly; x = ones gives the row sums; linearity y(a+b) ~ y(a)+y(b).
    N, K = 5120, 25600
    w = kf.fill_normal(ctx, N * K, 4242, 0.02)
    t = kf.quantize(ctx, w, N, K, kf.KF_T_Q4, 128, ol.RTN_ASYM)
    wdq = kf.dequant(ctx, t).numpy(np.uint16, (N, K))
    rng = np.random.default_rng(5)
    ks = rng.integers(0, K, size=8)
    x = np.zeros((8, K), dtype=np.uint16)
    x[np.arange(8), ks] = 0x3F80
    y = kf.linear(ctx, t, ctx.array(x), 8).numpy(np.uint16, (8, N))
    for m in range(8):
        assert np.array_equal(ol.bf16_to_f32(y[m]), ol.bf16_to_f32(wdq[:, ks[m]]))
    ones = np.full((1, K), 0x3F80, dtype=np.uint16)
    ysum = ol.bf16_to_f32(kf.linear(ctx, t, ctx.array(ones), 1).numpy(np.uint16))
    want = ol.bf16_to_f32(wdq).astype(np.float64).sum(1)
    assert np.allclose(ysum, want, rtol=2.0 ** -7, atol=2e-2)
    a, b = rand_bf16(rng, (1, K), 0.5), rand_bf16(rng, (1, K), 0.5)
    ab = ol.f32_to_bf16(ol.bf16_to_f32(a) + ol.bf16_to_f32(b))
    ya, yb, yab = (kf.linear(ctx, t, ctx.array(v), 1, kf.KF_EPI_F32).numpy(np.float32) for v in (a, b, ab))
    # ab is a+b rounded to bf16, so compare against the exact linear image of the rounded input
    assert np.allclose(yab, ol.linear_f32(wdq, ab, 1, N, K)[0], rtol=1e-3, atol=2e-3)
    assert np.allclose(ya + yb, yab, rtol=0.05, atol=0.05)


# ---------------------------------------------------------------------------------------------- small ops
@pytest.mark.parametrize("rows,dim", [(1, 1024), (1, 5120), (7, 4096), (64, 256)])
def test_rmsnorm(ctx, rows, dim):
    rng = np.random.default_rng(dim)
    x, w = rand_bf16(rng, (rows, dim), 3.0), rand_bf16(rng, (dim,), 0.5)
    got = kf.rmsnorm(ctx, ctx.array(x), ctx.array(w), rows, dim, 1e-6).numpy(np.uint16, (rows, dim))
    want = ol.rmsnorm(x, w, rows, dim, 1e-6)
    # identical formula; only the order of the fp32 sum of squares differs -> at most one bf16 ulp on a few elements
    d = np.abs(ol.bf16_to_f32(got) - ol.bf16_to_f32(want))
    assert (d <= np.abs(ol.bf16_to_f32(want)) * 2.0 ** -7).all()
    assert (got == want).mean() > 0.98


@pytest.mark.parametrize("hd,n_head,n_kv", [(128, 16, 8), (64, 4, 2)])
@pytest.mark.parametrize("M", [3, 40])
@pytest.mark.parametrize("theta", [1e4, 1e6])
def test_qknorm_rope_kvappend(ctx, hd, n_head, n_kv, theta, M):
    rng = np.random.default_rng(hd)
    max_seq = 64
    pos = np.array([5, 17, 63], dtype=np.int32) if M == 3 else (np.arange(M, dtype=np.int32) + 10)  # M >= 16: warp-per-head kernel
    q, k, v = rand_bf16(rng, (M, n_head * hd)), rand_bf16(rng, (M, n_kv * hd)), rand_bf16(rng, (M, n_kv * hd))
    qw, kw = rand_bf16(rng, (hd,), 0.3), rand_bf16(rng, (hd,), 0.3)
    qd = ctx.array(q)
    kc, vc = ctx.zeros(max_seq * n_kv * hd * 2), ctx.zeros(max_seq * n_kv * hd * 2)
    table = kf.rope_table(ctx, max_seq, hd, theta)
    kf.qknorm_rope_kvappend(ctx, qd, ctx.array(k), ctx.array(v), ctx.array(qw), ctx.array(kw), kc, vc, table, ctx.array(pos), M, n_head, n_kv,
                            hd, max_seq)
    q_got = qd.numpy(np.uint16, (M, n_head * hd))
    k_got = kc.numpy(np.uint16, (max_seq, n_kv * hd))
    v_got = vc.numpy(np.uint16, (max_seq, n_kv * hd))
    for m in range(M):
        q_want = ol.rope(ol.rmsnorm(q[m].reshape(n_head, hd), qw, n_head, hd, 1e-6), n_head, hd, int(pos[m]), theta).reshape(-1)
        k_want = ol.rope(ol.rmsnorm(k[m].reshape(n_kv, hd), kw, n_kv, hd, 1e-6), n_kv, hd, int(pos[m]), theta).reshape(-1)
        for got, want in ((q_got[m], q_want), (k_got[pos[m]], k_want)):
            d = np.abs(ol.bf16_to_f32(got) - ol.bf16_to_f32(want))
            assert (d <= np.abs(ol.bf16_to_f32(want)) * 2.0 ** -6 + 2e-3).all()
            assert (got == want).mean() > 0.95
        assert np.array_equal(v_got[pos[m]], v[m])
    untouched = np.setdiff1d(np.arange(max_seq), pos)
    assert not k_got[untouched].any() and not v_got[untouched].any()


@pytest.mark.parametrize("hd,n_head,n_kv", [(128, 16, 8), (128, 64, 8), (64, 4, 2)])
@pytest.mark.parametrize("pos", [0, 1, 31, 200, 1500])
def test_attention_decode(ctx, hd, n_head, n_kv, pos):
    # includes pos >= 1024, which the reference's attention_qk_kernel silently truncates
    rng = np.random.default_rng(pos + hd)
    max_seq = 2048
    q = rand_bf16(rng, (1, n_head * hd))
    kc, vc = rand_bf16(rng, (max_seq, n_kv * hd)), rand_bf16(rng, (max_seq, n_kv * hd))
    kcd, vcd = ctx.array(kc), ctx.array(vc)
    want = ol.bf16_to_f32(ol.attention_decode(q, kc, vc, pos, n_head, n_kv, hd, 0)).reshape(-1)
    for split in (0, 1, 5):
        ctx.set_int("attn_split", split)
        try:
            got = kf.attn_decode(ctx, ctx.array(q), kcd, vcd, ctx.array(np.array([pos], dtype=np.int32)), 1, n_head, n_kv, hd, max_seq, pos)
        finally:
            ctx.set_int("attn_split", 0)
        g = ol.bf16_to_f32(got.numpy(np.uint16))
        assert np.allclose(g, want, rtol=2.0 ** -6, atol=4e-3), (split, np.abs(g - want).max())


def test_attention_batched_sequences_and_prefill_panel(ctx):
    rng = np.random.default_rng(21)
    hd, n_head, n_kv, max_seq, M = 128, 16, 8, 256, 4
    q = rand_bf16(rng, (M, n_head * hd))
    pos = np.array([3, 77, 150, 255], dtype=np.int32)
    # (a) M independent sequences: cache [M][max_seq][kv_dim]
    kc, vc = rand_bf16(rng, (M, max_seq, n_kv * hd)), rand_bf16(rng, (M, max_seq, n_kv * hd))
    got = kf.attn_decode(ctx, ctx.array(q), ctx.array(kc), ctx.array(vc), ctx.array(pos), M, n_head, n_kv, hd, max_seq, 255,
                         seq_stride=max_seq * n_kv * hd).numpy(np.uint16, (M, n_head * hd))
    for m in range(M):
        want = ol.bf16_to_f32(ol.attention_decode(q[m], kc[m], vc[m], int(pos[m]), n_head, n_kv, hd, 0)).reshape(-1)
        assert np.allclose(ol.bf16_to_f32(got[m]), want, rtol=2.0 ** -6, atol=4e-3)
    # (b) one sequence, causal panel: token m attends to 0..pos[m]
    got = kf.attn_decode(ctx, ctx.array(q), ctx.array(kc[0]), ctx.array(vc[0]), ctx.array(pos), M, n_head, n_kv, hd, max_seq, 255).numpy(
        np.uint16, (M, n_head * hd))
    for m in range(M):
        want = ol.bf16_to_f32(ol.attention_decode(q[m], kc[0], vc[0], int(pos[m]), n_head, n_kv, hd, 0)).reshape(-1)
        assert np.allclose(ol.bf16_to_f32(got[m]), want, rtol=2.0 ** -6, atol=4e-3)


@pytest.mark.parametrize("hd,n_head,n_kv", [(128, 16, 8), (128, 64, 8), (64, 4, 2)])
@pytest.mark.parametrize("split", [0, 1, 3, 8, 12])  # <= 8 slices: cluster / DSMEM merge; more: global workspace + last-CTA merge
def test_fused_qkv_attention_equals_unfused_path(ctx, hd, n_head, n_kv, split):
    rng = np.random.default_rng(hd + n_head)
    M, max_seq, theta = 3, 256, 1e6
    pos = np.array([0, 37, 255], dtype=np.int32)
    q, k, v = rand_bf16(rng, (M, n_head * hd)), rand_bf16(rng, (M, n_kv * hd)), rand_bf16(rng, (M, n_kv * hd))
    qw, kw = rand_bf16(rng, (hd,), 0.3), rand_bf16(rng, (hd,), 0.3)
    kc0, vc0 = rand_bf16(rng, (M, max_seq, n_kv * hd)), rand_bf16(rng, (M, max_seq, n_kv * hd))
    table = kf.rope_table(ctx, max_seq, hd, theta)
    stride = max_seq * n_kv * hd
    posd = ctx.array(pos)
    # unfused reference path
    qd, kc, vc = ctx.array(q), ctx.array(kc0), ctx.array(vc0)
    kf.qknorm_rope_kvappend(ctx, qd, ctx.array(k), ctx.array(v), ctx.array(qw), ctx.array(kw), kc, vc, table, posd, M, n_head, n_kv, hd, max_seq,
                            seq_stride=stride)
    want = kf.attn_decode(ctx, qd, kc, vc, posd, M, n_head, n_kv, hd, max_seq, 255, seq_stride=stride).numpy(np.uint16)
    # fused
    kc2, vc2 = ctx.array(kc0), ctx.array(vc0)
    ctx.set_int("attn_split", split)
    try:
        got = kf.qkv_attention(ctx, ctx.array(q), ctx.array(k), ctx.array(v), ctx.array(qw), ctx.array(kw), kc2, vc2, table, posd, M, n_head, n_kv,
                               hd, max_seq, 255, seq_stride=stride).numpy(np.uint16)
    finally:
        ctx.set_int("attn_split", 0)
    g, w = ol.bf16_to_f32(got), ol.bf16_to_f32(want)
    assert np.allclose(g, w, rtol=2.0 ** -6, atol=4e-3), np.abs(g - w).max()
    assert np.array_equal(vc2.numpy(np.uint16), vc.numpy(np.uint16))          # V rows are plain copies
    kd = np.abs(ol.bf16_to_f32(kc2.numpy(np.uint16)) - ol.bf16_to_f32(kc.numpy(np.uint16)))
    assert kd.max() <= 4e-2 and (kc2.numpy(np.uint16) == kc.numpy(np.uint16)).mean() > 0.999  # K rows: same formula, other sum order
    # the oracle agrees as well
    for m in range(M):
        kco, vco = kc.numpy(np.uint16).reshape(M, max_seq, -1)[m], vc.numpy(np.uint16).reshape(M, max_seq, -1)[m]
        qn = ol.rope(ol.rmsnorm(q[m].reshape(n_head, hd), qw, n_head, hd, 1e-6), n_head, hd, int(pos[m]), theta)
        wo = ol.bf16_to_f32(ol.attention_decode(qn, kco, vco, int(pos[m]), n_head, n_kv, hd, 0)).reshape(-1)
        assert np.allclose(g.reshape(M, -1)[m], wo, rtol=2.0 ** -6, atol=6e-3)
    with pytest.raises(kf.KoifishError):  # several tokens of ONE sequence must use the two-step path
        kf.qkv_attention(ctx, ctx.array(q), ctx.array(k), ctx.array(v), ctx.array(qw), ctx.array(kw), kc2, vc2, table, posd, M, n_head, n_kv, hd,
                         max_seq, 255, seq_stride=0)


def test_swiglu_add_embed_argmax(ctx):
    rng = np.random.default_rng(31)
    n = 5000
    g, u = rand_bf16(rng, (n,), 2.0), rand_bf16(rng, (n,), 2.0)
    sw = kf.swiglu(ctx, ctx.array(g), ctx.array(u), n).numpy(np.uint16)
    want = ol.swiglu(g, u)
    assert (sw == want).mean() > 0.999 and np.allclose(ol.bf16_to_f32(sw), ol.bf16_to_f32(want), rtol=2.0 ** -7, atol=1e-30)
    assert np.array_equal(kf.add(ctx, ctx.array(g), ctx.array(u), n).numpy(np.uint16), ol.add(g, u))
    # embedding rows, plain and quantised tables
    rows, cols = 1024, 512
    toks = np.array([0, 1023, 77, 77], dtype=np.int32)
    for kind in ("bf16", "f8", (4, ol.RTN_ASYM), (2, ol.YYANG), (1, ol.YYANG)):
        t, wdq = make_weight(ctx, kind, rows, cols, 44)
        e = kf.embed(ctx, t, ctx.array(toks), toks.size).numpy(np.uint16, (toks.size, cols))
        assert np.array_equal(e, wdq[toks]), kind
    # argmax with ties -> lowest index
    logits = rand_bf16(rng, (3, 151936))
    logits[1, 5] = logits[1, 100000] = 0x4700
    logits[2, :] = 0x3F80
    am = kf.argmax(ctx, ctx.array(logits), 3, 151936).numpy(np.int32)
    assert am[0] == int(np.argmax(ol.bf16_to_f32(logits[0]))) and am[1] == 5 and am[2] == 0
    odd = rand_bf16(rng, (2, 1001))  # not a multiple of 8: scalar path
    odd[1, 1000] = 0x4700
    am = kf.argmax(ctx, ctx.array(odd), 2, 1001).numpy(np.int32)
    assert am[0] == int(np.argmax(ol.bf16_to_f32(odd[0]))) and am[1] == 1000


@pytest.mark.parametrize("kind", WEIGHT_KINDS, ids=str)
@pytest.mark.parametrize("M", [1, 7, 16, 24, 48, 64])
def test_gemm_tc_small_token_counts(ctx, kind, M):
    # the same kernel with 16 / 32 / 64-token tiles (ctx knob tc_min_m = 1 routes every M to it); K = 1536 gives a ragged last raw
    # stage for the 1-bit / 2-bit formats, N = 400 a ragged last row tile
    N, K = 400, 1536
    t, wdq = make_weight(ctx, kind, N, K, 5000 + M)
    x = rand_bf16(np.random.default_rng(M), (M, K))
    ctx.set_int("tc_min_m", 1)
    try:
        y = kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16)
    finally:
        ctx.set_int("tc_min_m", -1)
    _check_linear(y, wdq, x, M, N, K)


@pytest.mark.parametrize("splitk", [2, 3, 5])
@pytest.mark.parametrize("M", [3, 40, 130])
def test_gemm_tc_splitk_is_deterministic_and_matches(ctx, splitk, M):
    N, K = 272, 2048
    rng = np.random.default_rng(splitk * 100 + M)
    t, wdq = make_weight(ctx, (4, ol.RTN_ASYM), N, K, 6000)
    x, res = rand_bf16(rng, (M, K)), rand_bf16(rng, (M, N))
    xd = ctx.array(x)
    ctx.set_int("tc_min_m", 1)
    try:
        f1 = kf.linear(ctx, t, xd, M, kf.KF_EPI_F32).numpy(np.float32)
        ctx.set_int("gemv_splitk", splitk)
        ys = [kf.linear(ctx, t, xd, M).numpy(np.uint16) for _ in range(3)]
        fs = kf.linear(ctx, t, xd, M, kf.KF_EPI_F32).numpy(np.float32)
        yr = kf.linear(ctx, t, xd, M, kf.KF_EPI_RESIDUAL, ctx.array(res)).numpy(np.uint16)
    finally:
        ctx.set_int("gemv_splitk", 0)
        ctx.set_int("tc_min_m", -1)
    assert all(np.array_equal(ys[0], y) for y in ys[1:])  # ordered reduction by the last CTA: run-to-run identical
    _check_linear(ys[0], wdq, x, M, N, K)
    assert np.allclose(fs, f1, rtol=1e-5, atol=1e-4)
    assert np.array_equal(yr, ol.add(res, ys[0]))


def test_gemm_tc_onehot_small_m(ctx):
    M, N, K = 16, 256, 1024
    for kind in WEIGHT_KINDS:
        t, wdq = make_weight(ctx, kind, N, K, 7000)
        ks = (np.arange(M) * 61 + 3) % K
        x = np.zeros((M, K), dtype=np.uint16)
        x[np.arange(M), ks] = 0x3F80
        ctx.set_int("tc_min_m", 1)
        try:
            y = kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16, (M, N))
        finally:
            ctx.set_int("tc_min_m", -1)
        for m in range(M):
            assert np.array_equal(ol.bf16_to_f32(y[m]), ol.bf16_to_f32(wdq[:, ks[m]])), (kind, m)


# ---------------------------------------------------------------------------------------------- prefill (flash) attention
@pytest.mark.parametrize("hd,n_head,n_kv", [(128, 8, 2), (64, 4, 4)])
@pytest.mark.parametrize("M,pos0", [(16, 0), (64, 0), (100, 37), (200, 300), (65, 511)])
def test_attn_prefill_equals_per_token_decode_attention(ctx, hd, n_head, n_kv, M, pos0):
    # the panel's queries at positions pos0 .. pos0 + M - 1 over cache rows [0, pos0 + M): the tensor-core flash kernel against the
    # per-token decode kernel (itself pinned to the oracle above) and against the oracle for a few rows
    rng = np.random.default_rng(M * 7 + pos0 + hd)
    max_seq = 768
    q = rand_bf16(rng, (M, n_head * hd))
    kc, vc = rand_bf16(rng, (max_seq, n_kv * hd)), rand_bf16(rng, (max_seq, n_kv * hd))
    pos = np.arange(pos0, pos0 + M, dtype=np.int32)
    qd, kcd, vcd, posd = ctx.array(q), ctx.array(kc), ctx.array(vc), ctx.array(pos)
    want = kf.attn_decode(ctx, qd, kcd, vcd, posd, M, n_head, n_kv, hd, max_seq, pos0 + M - 1).numpy(np.uint16)
    got = kf.attn_prefill(ctx, qd, kcd, vcd, posd, M, n_head, n_kv, hd, max_seq).numpy(np.uint16)
    g, w = ol.bf16_to_f32(got), ol.bf16_to_f32(want)
    assert np.allclose(g, w, rtol=2.0 ** -6, atol=6e-3), np.abs(g - w).max()
    for m in (0, M // 2, M - 1):
        ref = ol.attention_decode(q[m], kc, vc, int(pos[m]), n_head, n_kv, hd)
        assert np.allclose(g.reshape(M, -1)[m], ol.bf16_to_f32(ref).reshape(-1), rtol=2.0 ** -6, atol=6e-3)


@pytest.mark.parametrize("hd,n_head,n_kv", [(128, 64, 8), (128, 8, 2), (64, 16, 8), (128, 16, 1)])
@pytest.mark.parametrize("split", [0, 1, 3])
def test_attn_decode_gqa_equals_per_head_decode_attention(ctx, hd, n_head, n_kv, split):
    # batched decode: M sequences at different positions; the kv-group tensor-core kernel against the per-head kernel
    rng = np.random.default_rng(hd + n_head + split)
    M, max_seq = 5, 320
    pos = np.array([0, 1, 63, 200, 319], dtype=np.int32)
    q = rand_bf16(rng, (M, n_head * hd))
    kc, vc = rand_bf16(rng, (M, max_seq, n_kv * hd)), rand_bf16(rng, (M, max_seq, n_kv * hd))
    stride = max_seq * n_kv * hd
    qd, kcd, vcd, posd = ctx.array(q), ctx.array(kc), ctx.array(vc), ctx.array(pos)
    want = kf.attn_decode(ctx, qd, kcd, vcd, posd, M, n_head, n_kv, hd, max_seq, 319, seq_stride=stride).numpy(np.uint16)
    ctx.set_int("attn_split", split)
    try:
        got = kf.attn_decode_gqa(ctx, qd, kcd, vcd, posd, M, n_head, n_kv, hd, max_seq, 319, seq_stride=stride).numpy(np.uint16)
    finally:
        ctx.set_int("attn_split", 0)
    g, w = ol.bf16_to_f32(got), ol.bf16_to_f32(want)
    assert np.allclose(g, w, rtol=2.0 ** -6, atol=6e-3), np.abs(g - w).max()
    ref = ol.attention_decode(q[3], kc[3], vc[3], 200, n_head, n_kv, hd)
    assert np.allclose(g.reshape(M, -1)[3], ol.bf16_to_f32(ref).reshape(-1), rtol=2.0 ** -6, atol=6e-3)


@pytest.mark.parametrize("kind", [(4, ol.RTN_ASYM), (2, ol.YYANG), "bf16"], ids=str)
@pytest.mark.parametrize("M", [16, 48, 300])
def test_gemm_tc_multi_weight_launch_equals_single_launches(ctx, kind, M):
    # Q / K / V (and gate / up) share ONE tensor-core launch above the crossover: row tiles of the three weights form one item space
    K = 1024
    rows = (400, 128, 144)  # ragged last tiles
    ws = [make_weight(ctx, kind, n, K, 8000 + i)[0] for i, n in enumerate(rows)]
    x = ctx.array(rand_bf16(np.random.default_rng(M), (M, K)))
    ctx.set_int("tc_min_m", 1)
    try:
        fused = kf.linear_multi(ctx, ws, x, M)
        single = [kf.linear(ctx, w, x, M) for w in ws]
        sw_f = kf.linear_swiglu(ctx, ws[1], ws[1], x, M).numpy(np.uint16)
    finally:
        ctx.set_int("tc_min_m", -1)
    for a, b in zip(fused, single):
        assert np.array_equal(a.numpy(np.uint16), b.numpy(np.uint16))
    g = single[1].numpy(np.uint16)
    assert np.array_equal(sw_f, ol.swiglu(g, g)) or (sw_f == ol.swiglu(g, g)).mean() > 0.999


@pytest.mark.parametrize("kind", WEIGHT_KINDS, ids=str)
@pytest.mark.parametrize("M,N,K", [(20, 48, 128), (9, 16, 256), (70, 16, 128), (33, 272, 384)])
def test_gemm_tc_edge_shapes(ctx, kind, M, N, K):
    # fewer rows than one 128-row tile, the shortest K the formats allow (one raw stage or less for the 1- / 2-bit formats), odd token counts
    t, wdq = make_weight(ctx, kind, N, K, 9000 + M)
    x = rand_bf16(np.random.default_rng(M + K), (M, K))
    ctx.set_int("tc_min_m", 1)
    try:
        y = kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16)
    finally:
        ctx.set_int("tc_min_m", -1)
    _check_linear(y, wdq, x, M, N, K)


def test_attn_prefill_at_the_end_of_the_cache(ctx):
    # the last KV tile reaches past max_seq: rows are clamped for the load and masked for the math
    hd, n_head, n_kv, max_seq, pos0, M = 128, 8, 2, 100, 60, 40
    rng = np.random.default_rng(5)
    q = rand_bf16(rng, (M, n_head * hd))
    kc, vc = rand_bf16(rng, (max_seq, n_kv * hd)), rand_bf16(rng, (max_seq, n_kv * hd))
    pos = np.arange(pos0, pos0 + M, dtype=np.int32)
    qd, kcd, vcd, posd = ctx.array(q), ctx.array(kc), ctx.array(vc), ctx.array(pos)
    want = kf.attn_decode(ctx, qd, kcd, vcd, posd, M, n_head, n_kv, hd, max_seq, max_seq - 1).numpy(np.uint16)
    got = kf.attn_prefill(ctx, qd, kcd, vcd, posd, M, n_head, n_kv, hd, max_seq).numpy(np.uint16)
    g, w = ol.bf16_to_f32(got), ol.bf16_to_f32(want)
    assert np.allclose(g, w, rtol=2.0 ** -6, atol=6e-3), np.abs(g - w).max()


# ---------------------------------------------------------------------------------------------- device sampler (GoPT.cpp:614-630)
@pytest.mark.parametrize("selection", [0, 1])
@pytest.mark.parametrize("vocab,top_k,top_p,T", [(151936, 50, 0.95, 0.6), (4096, 1000, 0.9, 1.3), (32768, 7, 1.0, 0.7), (151936, 64, 0.5, 2.0)])
def test_device_sampler_matches_cpu_port(ctx, selection, vocab, top_k, top_p, T):
    # same logits, same seeds: the device sampler must draw the token the CPU port of GeneratOnPrompt::Sample draws.  The two use different
    # expf implementations (last-ulp differences of the probabilities), so a draw whose coin lands within 1e-5 of a cdf boundary may differ:
    # allowed for at most 1 draw in 100.  bf16 logits tie massively (8 mantissa bits over 152 K values): the tie rules are exercised.
    M, rounds = 4, 25
    rng = np.random.default_rng(vocab + top_k)
    lg = ol.f32_to_bf16((rng.standard_normal((M, vocab)) * 2.5).astype(np.float32))
    lgd = ctx.array(lg)
    seeds = np.array([42, 43, 0x9E3779B97F4A7C15, 7], dtype=np.uint64)
    std = ctx.array(seeds.view(np.uint16))
    states = [[int(s)] for s in seeds]
    diff = 0
    for _ in range(rounds):
        got = kf.sample(ctx, lgd, M, vocab, T, top_k, top_p, std, selection).numpy(np.int32)
        for m in range(M):
            want, _ = ol.sample(lg[m], T, top_k, top_p, states[m], selection)
            diff += int(got[m] != want)
    assert diff <= 1, diff
    assert np.array_equal(std.numpy(np.uint64), np.array([s[0] for s in states], dtype=np.uint64))  # generator states advanced identically


def test_device_sampler_greedy_paths_and_errors(ctx):
    vocab = 5000
    lg = ol.f32_to_bf16(np.random.default_rng(3).standard_normal((2, vocab)).astype(np.float32))
    lgd = ctx.array(lg)
    st = ctx.array(np.array([1, 2], dtype=np.uint64).view(np.uint16))
    want = np.argmax(ol.bf16_to_f32(lg).reshape(2, vocab), axis=1)
    assert np.array_equal(kf.sample(ctx, lgd, 2, vocab, 0.0, 50, 0.9, st).numpy(np.int32), want)
    assert np.array_equal(kf.sample(ctx, lgd, 2, vocab, 0.7, 1, 0.9, st).numpy(np.int32), want)
    with pytest.raises(kf.KoifishError):
        kf.sample(ctx, lgd, 2, vocab, 0.7, 2000, 0.9, st)  # more than 1024 candidates


@pytest.mark.parametrize("kind", [(4, ol.RTN_ASYM), "bf16", (2, ol.YYANG)], ids=str)
@pytest.mark.parametrize("M", [1, 5, 96])
def test_linear_axb_alpha_beta_bias(ctx, kind, M):
    # TASKA_AxB (GTensor.hpp:698-741): d = alpha * x . w^T + beta * d + bias, fp32 epilogue, one rounding (GEMV for small M, tcgen05 for 96)
    N, K = 256, 1024
    t, wdq = make_weight(ctx, kind, N, K, 4242)
    rng = np.random.default_rng(M)
    x, d0, bias = rand_bf16(rng, (M, K)), rand_bf16(rng, (M, N)), rand_bf16(rng, (N,))
    alpha, beta = 0.5, -1.25
    ctx.set_int("gemv_exact", 1)
    dd = ctx.array(d0)
    kf.linear_axb(ctx, t, ctx.array(x), M, dd, alpha, beta, ctx.array(bias))
    got = ol.bf16_to_f32(dd.numpy(np.uint16)).reshape(M, N)
    acc = ol.linear_f32(wdq, x, M, N, K).reshape(M, N)
    want = alpha * acc + beta * ol.bf16_to_f32(d0).reshape(M, N) + ol.bf16_to_f32(bias)[None, :]
    assert np.all(np.abs(got - want) <= np.abs(want) * 2.0 ** -8 + 2e-3 * np.sqrt(np.mean(want ** 2)))
    # alpha 1, beta 0, no bias is kf_linear
    dd = ctx.array(d0)
    kf.linear_axb(ctx, t, ctx.array(x), M, dd)
    assert np.array_equal(dd.numpy(np.uint16), kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16))


# ---------------------------------------------------------------------------------------------- NormalFloat4 ({"bits": 4} without a quant_method)
def _nf4_tensor(ctx, rows, cols, seed):
    w = ol.fill_normal(rows * cols, seed, 0.02)
    data, gama = ol.nf4_quantize(w, rows, cols)
    t = kf.QTensor.from_packed(ctx, data, gama, rows, cols, kf.KF_T_NF4, 0, 0)
    return w, data, gama, t, ol.nf4_dequant(data, gama, rows, cols)


@pytest.mark.parametrize("rows,cols", [(64, 256), (300, 1024), (16, 5120)])
def test_nf4_quantize_and_dequant_bit_exact(ctx, rows, cols):
    w, data, gama, t, wdq = _nf4_tensor(ctx, rows, cols, rows + cols)
    q = kf.quantize(ctx, ctx.array(w), rows, cols, kf.KF_T_NF4, 0, 0)
    assert np.array_equal(q.gama_numpy(), gama)  # per-row bf16 codebooks
    assert np.array_equal(q.data_numpy(), data)  # codes: nearest entry of the fp32 codebook, MSB-first nibble stream
    assert np.array_equal(kf.dequant(ctx, t).numpy(np.uint16).reshape(rows, cols), wdq)
    toks = np.array([0, rows - 1, rows // 2], dtype=np.int32)
    got = kf.embed(ctx, t, ctx.array(toks.view(np.uint16)), 3).numpy(np.uint16).reshape(3, cols)
    assert np.array_equal(got, wdq[toks])


def test_nf4_dequant_matches_reference_kernel(ctx):
    # the reference's own CU_Q42X_NF4 (src/Device/CUDA/kernel/quantizer.cu:612-654, compiled where it lies: oracle/ref_kernels_q.cu) on the
    # same packed bytes and codebooks
    ref = ol.refq()
    if ref is None:
        pytest.skip("oracle/_ref/libkoifish_refq.so not built (reference tree absent at build time)")
    rows, cols = 96, 2048
    w, data, gama, t, wdq = _nf4_tensor(ctx, rows, cols, 5)
    out = ctx.empty(rows * cols * 2)
    ctx.sync()
    assert ref.refq_nf4_dequant(t.gama_ptr, t.data_ptr, out.ptr, rows, cols) == 0
    assert np.array_equal(out.numpy(np.uint16).reshape(rows, cols), wdq)
    assert np.array_equal(kf.dequant(ctx, t).numpy(np.uint16).reshape(rows, cols), wdq)


@pytest.mark.parametrize("M", [1, 3, 8, 40])
def test_nf4_linear_matches_oracle(ctx, M):
    N, K = 272, 2048
    w, data, gama, t, wdq = _nf4_tensor(ctx, N, K, 17)
    rng = np.random.default_rng(M)
    x = rand_bf16(rng, (M, K))
    y = kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16)
    _check_linear(y, wdq, x, M, N, K)
    res = rand_bf16(rng, (M, N))
    yr = kf.linear(ctx, t, ctx.array(x), M, kf.KF_EPI_RESIDUAL, ctx.array(res)).numpy(np.uint16)
    assert np.array_equal(yr.reshape(M, N), ol.add(res, y).reshape(M, N))
    # fused norm + SwiGLU entry point through the NF4 route
    nw = ol.f32_to_bf16((1.0 + 0.1 * rng.standard_normal(K)).astype(np.float32))
    w2, d2, g2, t2, wdq2 = _nf4_tensor(ctx, N, K, 18)
    got = ol.bf16_to_f32(kf.rmsnorm_linear(ctx, [t, t2], ctx.array(x), ctx.array(nw), M, 1e-6, swiglu=True).numpy(np.uint16)).reshape(M, N)
    xn = ol.rmsnorm(x, nw, M, K)
    want = ol.bf16_to_f32(ol.swiglu(ol.linear(wdq, xn, M, N, K), ol.linear(wdq2, xn, M, N, K))).reshape(M, N)
    assert np.abs(got - want).max() <= 2e-2 * np.abs(want).max()


# ---------------------------------------------------------------------------------------------- round-2 regressions
@pytest.mark.parametrize("M", [1, 5, 40])
def test_mixed_storage_types_share_a_call(ctx, M):
    # the quantizer card selects the type per tensor-name substring (Init4Neuron), so Q / K / V or gate / up of one block may differ:
    # kf_rmsnorm_linear / kf_linear_multi group the weights that agree into one launch each; results equal the separate calls
    K = 1024
    rng = np.random.default_rng(99 + M)
    ws = [make_weight(ctx, kind, n, K, 600 + i)[0] for i, (kind, n) in enumerate((((4, ol.RTN_ASYM), 256), ("f8", 128), ((4, ol.RTN_ASYM), 128)))]
    xd = ctx.array(rand_bf16(rng, (M, K)))
    outs = kf.linear_multi(ctx, ws, xd, M)
    for w, o in zip(ws, outs):
        assert np.array_equal(o.numpy(np.uint16), kf.linear(ctx, w, xd, M).numpy(np.uint16))
    nw = ctx.array(ol.f32_to_bf16((1.0 + 0.1 * rng.standard_normal(K)).astype(np.float32)))
    xn = kf.rmsnorm(ctx, xd, nw, M, K)
    for w, o in zip(ws, kf.rmsnorm_linear(ctx, ws, xd, nw, M)):
        assert np.array_equal(o.numpy(np.uint16), kf.linear(ctx, w, xn, M).numpy(np.uint16))
    # SwiGLU of a mixed gate / up pair: two matmuls + the stand-alone CU_swiglu_v0
    wg, wu = make_weight(ctx, (4, ol.RTN_ASYM), 256, K, 610)[0], make_weight(ctx, "bf16", 256, K, 611)[0]
    fused = kf.rmsnorm_linear(ctx, [wg, wu], xd, nw, M, swiglu=True).numpy(np.uint16)
    g, u = kf.linear(ctx, wg, xn, M), kf.linear(ctx, wu, xn, M)
    want = kf.swiglu(ctx, g, u, M * 256).numpy(np.uint16)
    # a bf16 weight on its own goes to the tensor-core kernel, inside the mixed pair to the skinny one: same values, another fp32 summation
    # order -- equal to within one bf16 ulp of the gate / up products
    a, b = ol.bf16_to_f32(fused), ol.bf16_to_f32(want)
    assert np.all(np.abs(a - b) <= np.maximum(np.abs(a), np.abs(b)) * 2.0 ** -6 + 1e-3 * np.sqrt(np.mean(b ** 2)))
    assert (fused == want).mean() > 0.97


@pytest.mark.parametrize("M,N,K", [(1, 512, 2048), (4, 384, 4096), (8, 256, 1024)])
def test_gemv_tma_exact_variant_matches_oracle(ctx, M, N, K):
    # the opt-in persistent TMA-fed stream-K kernel (gemv_tma.cu, knob gemv_tma = 1), bit-faithful arithmetic
    t, wdq = make_weight(ctx, (4, ol.RTN_ASYM), N, K, 1700 + M)
    x = rand_bf16(np.random.default_rng(M), (M, K))
    ctx.set_int("gemv_tma", 1)
    try:
        y = kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16)
    finally:
        ctx.set_int("gemv_tma", 0)
    _check_linear(y, wdq, x, M, N, K)


@pytest.mark.parametrize("M,N,K", [(1, 512, 2048), (8, 256, 1024)])
def test_gemv_tma_fast_variant_matches_oracle(ctx, M, N, K):
    t, wdq = make_weight(ctx, (4, ol.RTN_ASYM), N, K, 1800 + M)
    x = rand_bf16(np.random.default_rng(M), (M, K))
    ctx.set_int("gemv_tma", 1)
    try:
        y = kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16)
    finally:
        ctx.set_int("gemv_tma", 0)
    _check_linear(y, wdq, x, M, N, K, noise=FAST_NOISE)


@pytest.mark.parametrize("kind", [(4, ol.RTN_ASYM), (2, ol.YYANG), (1, ol.YYANG)], ids=str)
@pytest.mark.parametrize("M", [3, 4, 7, 8])
def test_gemv_fast_prestaged_activations_equal_per_cta_staging(ctx, kind, M):
    # 4..8 tokens: the activations are converted to fragment order ONCE by kf_gemv_xprep_kernel and fetched from global memory in the k-loop
    # (knob gemv_xg_min_m), instead of every CTA staging its slice in shared memory.  Same staging code, same k-split -> the same bits;
    # with the fused RMSNorm, the residual epilogue and SwiGLU as well
    N, K = 384, 2048
    rng = np.random.default_rng(40 + M)
    t, wdq = make_weight(ctx, kind, N, K, 2100 + M)
    t2, _ = make_weight(ctx, kind, N, K, 2200 + M)
    xd = ctx.array(rand_bf16(rng, (M, K)))
    nw = ctx.array(ol.f32_to_bf16((1.0 + 0.1 * rng.standard_normal(K)).astype(np.float32)))
    res = ctx.array(rand_bf16(rng, (M, N)))
    outs = {}
    ctx.set_int("gemv_splitk", 2)
    try:
        for xg in (0, 3):
            ctx.set_int("gemv_xg_min_m", -xg)  # negative: always from |value| tokens (the default only where the plan needs two waves)
            outs[xg] = (kf.linear(ctx, t, xd, M).numpy(np.uint16), kf.linear(ctx, t, xd, M, kf.KF_EPI_RESIDUAL, res).numpy(np.uint16),
                        kf.rmsnorm_linear(ctx, [t, t2], xd, nw, M, swiglu=True).numpy(np.uint16),
                        kf.rmsnorm_linear(ctx, [t, t2], xd, nw, M)[1].numpy(np.uint16))
    finally:
        ctx.set_int("gemv_splitk", 0)
        ctx.set_int("gemv_xg_min_m", 3)
    for a, b in zip(outs[0], outs[3]):
        assert np.array_equal(a, b)
    _check_linear(outs[3][0], wdq, xd.numpy(np.uint16).reshape(M, K), M, N, K, noise=_fast_noise(kind))


# ---------------------------------------------------------------------------------------------- vendor AWQ layout (device level)
def _awq_tensor(ctx, IC, OC, seed):
    w_io = ol.fill_normal(IC * OC, seed, 0.02)  # [in][out]
    qw, qz, sc = ol.awq_pack(w_io, IC, OC)
    deq_io = ol.awq_dequant(qw, qz, sc, IC, OC)  # bf16 [in][out], the reference's GetDataX order
    return kf.AwqTensor(ctx, qw, qz, sc, IC, OC), np.ascontiguousarray(deq_io.T), deq_io


@pytest.mark.parametrize("IC,OC", [(256, 512), (1024, 264), (4096, 1024)])
def test_awq_dequant_bit_exact_and_matches_reference_kernel(ctx, IC, OC):
    t, wdq, deq_io = _awq_tensor(ctx, IC, OC, IC + OC)
    assert np.array_equal(kf.dequant(ctx, t).numpy(np.uint16).reshape(OC, IC), wdq)  # [out][in], like every other type
    ref = ol.refq()
    if ref is None:
        pytest.skip("oracle/_ref/libkoifish_refq.so not built (reference tree absent at build time)")
    out = ctx.empty(IC * OC * 2)
    ctx.sync()
    # the reference's own CU_Q42X_awq (quantizer.cu:132-156) on the same buffers: [in][out]
    assert ref.refq_awq_dequant(t.qz.ptr, t.sc.ptr, t.qw.ptr, out.ptr, IC, OC) == 0
    assert np.array_equal(out.numpy(np.uint16).reshape(IC, OC), deq_io)


@pytest.mark.parametrize("M", [1, 3, 8, 40])
def test_awq_linear_matches_oracle(ctx, M):
    IC, OC = 2048, 528
    t, wdq, _ = _awq_tensor(ctx, IC, OC, 77)
    rng = np.random.default_rng(M)
    x = rand_bf16(rng, (M, IC))
    y = kf.linear(ctx, t, ctx.array(x), M).numpy(np.uint16)
    _check_linear(y, wdq, x, M, OC, IC)
    res = rand_bf16(rng, (M, OC))
    yr = kf.linear(ctx, t, ctx.array(x), M, kf.KF_EPI_RESIDUAL, ctx.array(res)).numpy(np.uint16)
    assert np.array_equal(yr.reshape(M, OC), ol.add(res, y).reshape(M, OC))
