"""Parity PIN: koifish_b200's kernels against the reference's OWN CUDA kernels run on the same B200.

oracle/ref_kernels.cu compiles /root/reference/src/Device/CUDA/T.cu (and the kernel headers it includes) where they lie, for sm_100a,
twice: with the reference's nvcc flags ("fma": -use_fast_math => -fmad=true, the dequant's bf16 multiply-subtract is contracted to one
fma.rn.bf16) and with -fmad=false ("nofma": two roundings, IEEE division -- the arithmetic of the source as written / of pre-sm_90
builds).  koifish_b200 reproduces either with the context knob deq_fma (default 1 = what the reference computes on a B200).

Bit-exact: dequantised weights (kf_dequant, both roundings, every packed format), the E5M2 byte codec, and the CPU oracle's dequant
against the same reference kernel.  Tolerance (written in each test): everything the reference computes with fast-math intrinsics or
stochastic rounding (RMSNorm's rsqrtf, RoPE's powf/sincosf + SquirrelNoise rounding, softmax's expf) and its GPU packer, whose float
arithmetic differs from the CPU packer the model loader uses (documented below)."""
import ctypes as C

import numpy as np
import pytest

import koifish_b200 as kf
import oracle_lib as ol

pytestmark = pytest.mark.gpu

KF_TYPE = {(4, ol.RTN_ASYM): kf.KF_T_Q4, (4, ol.RTN_SYM): kf.KF_T_Q4, (2, ol.YYANG): kf.KF_T_SIGN, (1, ol.YYANG): kf.KF_T_BINARY}


@pytest.fixture(scope="module")
def ctx():
    c = kf.Context(0)
    yield c
    c.set_int("deq_fma", 1)
    c.close()


def need(variant):
    R = ol.refgpu(variant)
    if R is None:
        pytest.skip("oracle/_ref/libkoifish_refgpu*.so was not built (reference tree absent at build time)")
    return R


def ulp_diff(a, b):
    """distance in bf16 ulps between two bf16 bit arrays (sign-magnitude -> monotone integer)"""
    def key(x):
        x = x.astype(np.int32)
        return np.where(x & 0x8000, -(x & 0x7fff), x & 0x7fff)
    return np.abs(key(np.asarray(a)) - key(np.asarray(b)))


def packed_case(ctx, rows, cols, bits, mode, seed, sigma):
    w = ol.fill_normal(rows * cols, seed, sigma)
    data, gama = ol.quantize(w, rows, cols, bits, 128, mode)
    _, _, qbias = ol.qrange(bits, mode)
    t = kf.QTensor.from_packed(ctx, data, gama, rows, cols, KF_TYPE[(bits, mode)], 128, qbias)
    return w, data, gama, qbias, t


# ------------------------------------------------------------------------------------------------ a8: CU_Q128toX_ (T.cu:245-294)
@pytest.mark.parametrize("variant", ["fma", "nofma"])
@pytest.mark.parametrize("bits,mode", [(4, ol.RTN_ASYM), (4, ol.RTN_SYM), (2, ol.YYANG), (1, ol.YYANG)])
def test_dequant_bit_exact_vs_reference_kernel(ctx, variant, bits, mode):
    R = need(variant)
    ctx.set_int("deq_fma", 1 if variant == "fma" else 0)
    ol.set_dequant_fma(1 if variant == "fma" else 0)
    try:
        rows, cols = 384, 1024
        for seed, sigma in ((11, 0.02), (12, 1.7), (13, 3e-4)):
            w, data, gama, qbias, t = packed_case(ctx, rows, cols, bits, mode, seed, sigma)
            nG = rows * cols // 128
            zero_off = t.data_bytes + 2 * (rows + cols)
            out_ref = ctx.empty(rows * cols * 2)
            rc = R.refk_q128tox(bits, nG, 128, qbias, t.data_ptr, t.blob.ptr + zero_off, t.blob.ptr + zero_off + 2 * nG, out_ref.ptr)
            assert rc == 0, "reference kernel launch failed: %d" % rc
            ref = out_ref.numpy(np.uint16, (rows, cols))
            ours = kf.dequant(ctx, t).numpy(np.uint16, (rows, cols))
            assert np.array_equal(ours, ref), "kf_dequant differs from the reference's CU_Q128toX_ (%s build): %d of %d elements" % (
                variant, int((ours != ref).sum()), ours.size)
            # the CPU oracle restates the same kernel: pinned to it here
            assert np.array_equal(ol.dequant(data, gama, rows, cols, bits, 128, qbias), ref), "CPU oracle dequant differs from the reference kernel"
    finally:
        ctx.set_int("deq_fma", 1)
        ol.set_dequant_fma(1)


def test_dequant_roundings_differ_between_reference_builds(ctx):
    """the two builds of the reference kernel really differ (so the test above discriminates), and only by the last bf16 bit"""
    Rf, Rn = need("fma"), need("nofma")
    rows, cols = 256, 1024
    w, data, gama, qbias, t = packed_case(ctx, rows, cols, 4, ol.RTN_ASYM, 21, 0.02)
    nG, zo = rows * cols // 128, t.data_bytes + 2 * (rows + cols)
    outs = []
    for R in (Rf, Rn):
        o = ctx.empty(rows * cols * 2)
        assert R.refk_q128tox(4, nG, 128, qbias, t.data_ptr, t.blob.ptr + zo, t.blob.ptr + zo + 2 * nG, o.ptr) == 0
        outs.append(o.numpy(np.uint16))
    d = ulp_diff(outs[0], outs[1])
    # differences of opposite sign near zero can be many "ulps" apart in this metric; compare values instead there
    fa, fb = ol.bf16_to_f32(outs[0]), ol.bf16_to_f32(outs[1])
    assert (d > 0).sum() > 0, "fused and two-rounding builds agree everywhere: the pin cannot tell them apart"
    # the extra rounding is that of the product p = step * k (half a bf16 ulp of p, |p| <= 15 * step), which may exceed an ulp of the
    # (cancelled) result p - zero: bound it by the group's scale instead
    step = np.repeat(np.abs(ol.bf16_to_f32(gama[rows + cols + nG:])), 128)
    assert np.all(np.abs(fa - fb) <= 2.0 ** -8 * 15.0 * step + 2.0 ** -8 * np.maximum(np.abs(fa), np.abs(fb)) + 1e-30)


# ------------------------------------------------------------------------------------------------ a10: E5M2 byte codec
def test_f8_codec_bit_exact_vs_reference_kernels(ctx):
    R = need("fma")
    n = 1 << 16
    w = ol.fill_normal(n, 31, 0.5)
    w[:256] = np.arange(256, dtype=np.uint16) << 8  # every exponent / sign pattern
    wd = ctx.array(w)
    enc_ref = ctx.empty(n)
    assert R.refk_f8_encode(wd.ptr, enc_ref.ptr, n) == 0
    t = kf.quantize(ctx, wd, 1, n, kf.KF_T_F8E5M2)
    b_ref = enc_ref.numpy(np.uint8)
    assert np.array_equal(t.data_numpy(), b_ref), "E5M2 encode differs from CU_Float2F8<bf16>"
    assert np.array_equal(ol.f8_encode(w), b_ref), "oracle E5M2 encode differs from CU_Float2F8<bf16>"
    dec_ref = ctx.empty(n * 2)
    assert R.refk_f8_decode(enc_ref.ptr, dec_ref.ptr, n) == 0
    d_ref = dec_ref.numpy(np.uint16)
    ours = kf.dequant(ctx, t).numpy(np.uint16)
    finite = (b_ref & 0x7c) != 0x7c  # inf / nan payloads: compare the finite codes bit for bit
    assert np.array_equal(ours[finite], d_ref[finite]), "E5M2 decode differs from CU_F82Float"
    assert np.array_equal(ol.f8_decode(b_ref)[finite], d_ref[finite])


# ------------------------------------------------------------------------------------------------ N1: CU_XtoQ128_ / CU_XtoYYang_ (T.cu:105-242)
@pytest.mark.parametrize("bits,mode,yyang", [(4, ol.RTN_ASYM, 0), (4, ol.RTN_SYM, 0), (2, ol.YYANG, 3), (1, ol.YYANG, 1)])
def test_gpu_packer_cross_check(ctx, bits, mode, yyang):
    """The reference has TWO packers: the CPU one (GeQuant::RTN_x / YinYang, GeQuant.cpp:428-628), whose bytes the loader uploads
    (LowBit_worker :875-878) and which kf_quantize reproduces, and the GPU one (SetDataX -> CU_XtoQ128_), used for re-quantisation.
    They are the same algorithm with different float arithmetic: the GPU kernel forms (max - min) in bf16 before dividing by the code
    range and sums |a| in fp32 instead of double.  So: zero must agree exactly, step within one bf16 ulp, codes within one level, and a
    weight dequantised from either set of bytes must agree within one step.  Layout and dequant of the REFERENCE's bytes: bit-exact."""
    R = need("nofma")  # IEEE division, as the CPU packer
    rows, cols = 256, 1024
    nG = rows * cols // 128
    qmin, qmax, qbias = ol.qrange(bits, mode)
    w = ol.fill_normal(rows * cols, 51, 0.02)
    wd = ctx.array(w)
    tr = kf.QTensor(ctx, rows, cols, KF_TYPE[(bits, mode)], 128, qbias)
    ctx.check(ctx.lib.kf_memset(ctx.h, tr.blob.ptr, 0, tr.blob.nbytes), "memset")
    zo = tr.data_bytes + 2 * (rows + cols)
    rc = R.refk_xtoq128(bits, nG, 128, qmin, qmax, qbias, 1 if mode == ol.RTN_SYM else 0, yyang, wd.ptr, tr.data_ptr, tr.blob.ptr + zo,
                        tr.blob.ptr + zo + 2 * nG)
    assert rc == 0
    ours = kf.quantize(ctx, wd, rows, cols, KF_TYPE[(bits, mode)], 128, {ol.RTN_ASYM: kf.KF_Q_RTN_ASYM, ol.RTN_SYM: kf.KF_Q_RTN_SYM,
                                                                           ol.YYANG: kf.KF_Q_YYANG}[mode])
    g_ref, g_our = tr.gama_numpy()[rows + cols:], ours.gama_numpy()[rows + cols:]
    assert np.array_equal(g_ref[:nG], g_our[:nG]), "zero differs from the reference's GPU packer"
    assert ulp_diff(g_ref[nG:], g_our[nG:]).max() <= 1, "step differs from the reference's GPU packer by more than one bf16 ulp"
    c_ref = ol.unpack_codes(tr.data_numpy(), rows * cols, bits)
    c_our = ol.unpack_codes(ours.data_numpy(), rows * cols, bits)
    assert np.abs(c_ref - c_our).max() <= 1
    agree = float((c_ref == c_our).mean())
    assert agree >= 0.97, "only %.4f of the codes agree with the reference's GPU packer" % agree
    # the reference's bytes through OUR dequant == through ITS dequant (layout + arithmetic), bit for bit
    o_ref = ctx.empty(rows * cols * 2)
    assert R.refk_q128tox(bits, nG, 128, qbias, tr.data_ptr, tr.blob.ptr + zo, tr.blob.ptr + zo + 2 * nG, o_ref.ptr) == 0
    ctx.set_int("deq_fma", 0)
    try:
        assert np.array_equal(kf.dequant(ctx, tr).numpy(np.uint16), o_ref.numpy(np.uint16))
    finally:
        ctx.set_int("deq_fma", 1)


# ------------------------------------------------------------------------------------------------ a13 / a14: RMSNorm kernels
@pytest.mark.parametrize("dim", [1024, 4096, 5120])
def test_rmsnorm_vs_reference_kernel(ctx, dim):
    """rms_norm_kernel<256> (layernorm.cuh:801-846) uses cub::BlockReduce and rsqrtf under -use_fast_math (rsqrt.approx, 2 ulp fp32):
    the bf16 result may differ from an IEEE 1/sqrt in the last bit.  Tolerance: <= 1 bf16 ulp, and >= 99 % of the elements bit-equal."""
    R = need("fma")
    rng = np.random.default_rng(dim)
    rows = 3
    x = ol.f32_to_bf16((rng.standard_normal((rows, dim)) * 2.5).astype(np.float32))
    wn = ol.f32_to_bf16((1.0 + 0.2 * rng.standard_normal(dim)).astype(np.float32))
    xd, wd = ctx.array(x), ctx.array(wn)
    ref = ctx.empty(rows * dim * 2)
    assert R.refk_rmsnorm(ref.ptr, xd.ptr, wd.ptr, rows, dim) == 0
    ours = kf.rmsnorm(ctx, xd, wd, rows, dim).numpy(np.uint16)
    d = ulp_diff(ours, ref.numpy(np.uint16))
    assert d.max() <= 1 and (d == 0).mean() >= 0.99, "max %d ulp, %.4f equal" % (d.max(), (d == 0).mean())
    assert ulp_diff(ol.rmsnorm(x, wn, rows, dim).reshape(-1), ref.numpy(np.uint16)).max() <= 1  # the CPU oracle against the same kernel


def test_qknorm_vs_reference_kernel(ctx):
    """CU_rmsnorm_multihead (layernorm.cuh:750-798) against the QK-norm of kf_qknorm_rope_kvappend at position 0 (RoPE is the identity
    there: cos = 1, sin = 0).  Same tolerance as above."""
    R = need("fma")
    rng = np.random.default_rng(5)
    n_head, n_kv, hd, max_seq = 16, 8, 128, 32
    q = ol.f32_to_bf16(rng.standard_normal((n_head, hd)).astype(np.float32))
    k = ol.f32_to_bf16(rng.standard_normal((n_kv, hd)).astype(np.float32))
    v = ol.f32_to_bf16(rng.standard_normal((n_kv, hd)).astype(np.float32))
    qw = ol.f32_to_bf16((1.0 + 0.3 * rng.standard_normal(hd)).astype(np.float32))
    kw = ol.f32_to_bf16((1.0 + 0.3 * rng.standard_normal(hd)).astype(np.float32))
    qd, kd, vd, qwd, kwd = (ctx.array(a) for a in (q, k, v, qw, kw))
    qr, kr = ctx.array(q), ctx.array(k)
    assert R.refk_rmsnorm_multihead(qr.ptr, qwd.ptr, n_head, hd, 64) == 0
    assert R.refk_rmsnorm_multihead(kr.ptr, kwd.ptr, n_kv, hd, 64) == 0
    kc, vc = ctx.zeros(max_seq * n_kv * hd * 2), ctx.zeros(max_seq * n_kv * hd * 2)
    table = kf.rope_table(ctx, max_seq, hd, 1e6)
    pos = ctx.array(np.array([0], dtype=np.int32))
    kf.qknorm_rope_kvappend(ctx, qd, kd, vd, qwd, kwd, kc, vc, table, pos, 1, n_head, n_kv, hd, max_seq)
    assert ulp_diff(qd.numpy(np.uint16), qr.numpy(np.uint16)).max() <= 1
    assert ulp_diff(kc.numpy(np.uint16, count=n_kv * hd), kr.numpy(np.uint16)).max() <= 1
    assert np.array_equal(vc.numpy(np.uint16, count=n_kv * hd), v.reshape(-1))


# ------------------------------------------------------------------------------------------------ a15: RoPE
@pytest.mark.parametrize("theta", [1e4, 1e6])
def test_rope_vs_reference_kernel(ctx, theta):
    """CU_rope2_v0 (operator.cuh:735-772) rounds STOCHASTICALLY (CU_Float2T<bf16>, packedN.cuh:62-72: SquirrelNoise keyed on the launch
    geometry) and evaluates powf / sincosf under -use_fast_math; koifish_b200 rounds to nearest with a host-built (cos, sin) table.
    Tolerance: each component within 2 bf16 ulps of its own magnitude scale, i.e. |a - b| <= 2^-6 * max(|q_j|, |q_j+64|) -- one ulp
    for stochastic vs nearest rounding, one for the fast-math angle at positions up to 512."""
    R = need("fma")
    rng = np.random.default_rng(9)
    n_head, n_kv, hd, max_seq = 16, 8, 128, 1024
    table = kf.rope_table(ctx, max_seq, hd, theta)
    for posv in (0, 1, 17, 300, 511):
        q = ol.f32_to_bf16(rng.standard_normal((n_head, hd)).astype(np.float32))
        k = ol.f32_to_bf16(rng.standard_normal((n_kv, hd)).astype(np.float32))
        v = np.zeros((n_kv, hd), dtype=np.uint16)
        qr, kr = ctx.array(q), ctx.array(k)
        assert R.refk_rope2(qr.ptr, kr.ptr, posv, n_head, n_kv, hd, theta) == 0
        qd, kd, vd = ctx.array(q), ctx.array(k), ctx.array(v)
        kc, vc = ctx.zeros(max_seq * n_kv * hd * 2), ctx.zeros(max_seq * n_kv * hd * 2)
        pos = ctx.array(np.array([posv], dtype=np.int32))
        kf.qknorm_rope_kvappend(ctx, qd, kd, vd, None, None, kc, vc, table, pos, 1, n_head, n_kv, hd, max_seq)
        for ours, ref, src in ((qd.numpy(np.uint16), qr.numpy(np.uint16), q),
                               (kc.numpy(np.uint16, offset=posv * n_kv * hd * 2, count=n_kv * hd), kr.numpy(np.uint16), k)):
            a, b = ol.bf16_to_f32(ours).reshape(-1, hd), ol.bf16_to_f32(ref).reshape(-1, hd)
            s = np.abs(ol.bf16_to_f32(src).reshape(-1, hd))
            scale = np.maximum(s[:, :hd // 2], s[:, hd // 2:])
            scale = np.concatenate([scale, scale], axis=1)
            assert np.all(np.abs(a - b) <= 2.0 ** -6 * scale + 1e-6), "pos %d: max excess %g" % (posv, float((np.abs(a - b) - 2.0 ** -6 * scale).max()))


# ------------------------------------------------------------------------------------------------ a17: decode attention
@pytest.mark.parametrize("score_bf16", [0, 1])
def test_attention_vs_reference_kernels(ctx, score_bf16):
    """attention_qk_kernel + CU_softmax_multihead + attention_v_kernel as SelfAttention::cuInfer launches them (QKV.cu:667-672) against
    kf_attn_decode.  The reference's expf is the fast-math ex2.approx path; its neuron path keeps scores in bf16 (score_bf16 = 1), the
    pipe path in fp32 (0).  koifish_b200 keeps scores in fp32: tolerance 2^-7 of the output scale against the fp32-score reference,
    2^-5 against the bf16-score one (bf16 probabilities carry 2^-9 relative error each)."""
    R = need("fma")
    rng = np.random.default_rng(3)
    n_head, n_kv, hd, max_seq = 16, 8, 128, 1024
    kc = ol.f32_to_bf16(rng.standard_normal((max_seq, n_kv * hd)).astype(np.float32))
    vc = ol.f32_to_bf16(rng.standard_normal((max_seq, n_kv * hd)).astype(np.float32))
    kcd, vcd = ctx.array(kc), ctx.array(vc)
    att = ctx.empty(n_head * max_seq * 4)
    for posv in (0, 5, 100, 511, 1000):
        q = ol.f32_to_bf16((rng.standard_normal((n_head, hd)) * 0.5).astype(np.float32))
        qd = ctx.array(q)
        ref = ctx.empty(n_head * hd * 2)
        assert R.refk_attention(ref.ptr, att.ptr, qd.ptr, kcd.ptr, vcd.ptr, posv, max_seq, n_head, n_kv, hd, score_bf16) == 0
        pos = ctx.array(np.array([posv], dtype=np.int32))
        ours = kf.attn_decode(ctx, qd, kcd, vcd, pos, 1, n_head, n_kv, hd, max_seq, posv)
        a, b = ol.bf16_to_f32(ours.numpy(np.uint16)), ol.bf16_to_f32(ref.numpy(np.uint16))
        tol = 2.0 ** (-5 if score_bf16 else -7)
        assert np.abs(a - b).max() <= tol * max(1e-3, np.abs(b).max()), "pos %d: %g vs scale %g" % (posv, np.abs(a - b).max(), np.abs(b).max())
        # and the CPU oracle (which restates these three kernels) against them
        o = ol.bf16_to_f32(ol.attention_decode(q, kc, vc, posv, n_head, n_kv, hd, score_bf16).reshape(-1))
        assert np.abs(o - b).max() <= tol * max(1e-3, np.abs(b).max())
