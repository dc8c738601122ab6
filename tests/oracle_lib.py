"""ctypes bindings for the CPU oracle (oracle/libkoifish_oracle.so) and the reference shim (oracle/_ref).

TEST INFRASTRUCTURE ONLY -- the product (koifish_b200/) never imports this module.
bf16 tensors are numpy uint16 arrays holding the raw bit patterns.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libkoifish_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libkoifish_ref.so")
# the reference's own CUDA kernels (src/Device/CUDA/T.cu + headers) compiled for sm_100a: with the reference's nvcc flags ("fma": the bf16
# multiply-subtract of the dequant is contracted to one fma.rn.bf16), and with -fmad=false / no fast-math ("nofma": two roundings, IEEE division)
REFGPU_SO = {"fma": os.path.join(ORACLE_DIR, "_ref", "libkoifish_refgpu.so"), "nofma": os.path.join(ORACLE_DIR, "_ref", "libkoifish_refgpu_nofma.so")}
REFQ_SO = os.path.join(ORACLE_DIR, "_ref", "libkoifish_refq.so")  # the reference's quantizer.cu (NF4 dequant kernel)
REFKUN_SO = os.path.join(ORACLE_DIR, "_ref", "libkoifish_refkun.so")  # the reference's fish.kun reader (Serialize.cpp / Safetensors.cpp)
REFTOK_SO = os.path.join(ORACLE_DIR, "_ref", "libkoifish_reftok.so")  # the reference's tokenizer (HF_Tokenizer.cpp + its vendored oniguruma / utf8proc)
REFCPU_SO = os.path.join(ORACLE_DIR, "_ref", "libkoifish_refcpu.so")  # the reference's CPU packers (GeQuant.cpp: RTN_x, YinYang, RT_NormalF)

RTN_ASYM, RTN_SYM, YYANG, NF4 = 0, 1, 2, 3


def build_oracle(force=False):
    """Compile oracle/ (and oracle/_ref when /root/reference is present). Building the checker is not using it."""
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("koifish_oracle.cpp", "koifish_oracle.h")]
    if force or not os.path.exists(ORACLE_SO) or any(os.path.getmtime(f) > os.path.getmtime(ORACLE_SO) for f in srcs):
        subprocess.check_call(["make", "-B" if force else "-s", "-C", ORACLE_DIR, "libkoifish_oracle.so"], stdout=subprocess.DEVNULL)
    shim = os.path.join(ORACLE_DIR, "ref_shim.cpp")
    if os.path.exists("/root/reference/src/PackedQ.hpp") and (force or not os.path.exists(REF_SO) or os.path.getmtime(shim) > os.path.getmtime(REF_SO)):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)
    shim_q = os.path.join(ORACLE_DIR, "ref_cpu_quant.cpp")
    if os.path.exists("/root/reference/src/Tensor/GeQuant.cpp") and (force or not os.path.exists(REFCPU_SO) or
                                                                      os.path.getmtime(shim_q) > os.path.getmtime(REFCPU_SO)):
        subprocess.call(["make", "-C", ORACLE_DIR, "refcpu"], stdout=subprocess.DEVNULL)  # optional checker: tests skip when it did not build
    shim_k = os.path.join(ORACLE_DIR, "ref_kun.cpp")
    if os.path.exists("/root/reference/src/Manifold/Serialize.cpp") and (force or not os.path.exists(REFKUN_SO) or
                                                                         os.path.getmtime(shim_k) > os.path.getmtime(REFKUN_SO)):
        subprocess.call(["make", "-C", ORACLE_DIR, "refkun"], stdout=subprocess.DEVNULL)  # optional checker: tests skip when it did not build
    shim_t = os.path.join(ORACLE_DIR, "ref_tokenizer.cpp")
    if os.path.exists("/root/reference/src/TokenSet/HF_Tokenizer.cpp") and (force or not os.path.exists(REFTOK_SO) or
                                                                            os.path.getmtime(shim_t) > os.path.getmtime(REFTOK_SO)):
        subprocess.call(["make", "-C", ORACLE_DIR, "reftok"], stdout=subprocess.DEVNULL)  # optional checker: tests skip when it did not build
    if os.path.exists("/root/reference/src/Device/CUDA/T.cu"):
        src = os.path.join(ORACLE_DIR, "ref_kernels.cu")
        stale = any(not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src) for so in REFGPU_SO.values())
        srcq = os.path.join(ORACLE_DIR, "ref_kernels_q.cu")
        stale = stale or not os.path.exists(REFQ_SO) or os.path.getmtime(REFQ_SO) < os.path.getmtime(srcq)
        if force or stale:
            subprocess.check_call(["make", "-C", ORACLE_DIR, "refgpu"], stdout=subprocess.DEVNULL)


class ModelConfig(C.Structure):
    _fields_ = [
        ("n_layer", C.c_int), ("n_embd", C.c_int), ("n_ff", C.c_int), ("n_head", C.c_int), ("n_kv_head", C.c_int),
        ("head_dim", C.c_int), ("vocab", C.c_int), ("max_seq", C.c_int),
        ("rope_theta", C.c_float), ("rms_eps", C.c_float),
        ("tie_embed", C.c_int),
        ("attn_bits", C.c_int), ("attn_mode", C.c_int),
        ("mlp_bits", C.c_int), ("mlp_mode", C.c_int),
        ("embed_bits", C.c_int), ("embed_mode", C.c_int),
        ("group", C.c_int),
        ("seed", C.c_uint64),
        ("sigma", C.c_float), ("norm_sigma", C.c_float),
        ("score_bf16", C.c_int),
    ]


class QRange(C.Structure):
    _fields_ = [("qmin", C.c_int), ("qmax", C.c_int), ("qbias", C.c_int)]


_u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")

_lib = None
_ref = None
_refgpu = {}


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        L.kfo_f32_to_bf16.restype = C.c_uint16
        L.kfo_f32_to_bf16.argtypes = [C.c_float]
        L.kfo_bf16_to_f32.restype = C.c_float
        L.kfo_bf16_to_f32.argtypes = [C.c_uint16]
        L.kfo_bf16_mul.restype = C.c_uint16
        L.kfo_bf16_mul.argtypes = [C.c_uint16, C.c_uint16]
        L.kfo_bf16_sub.restype = C.c_uint16
        L.kfo_bf16_sub.argtypes = [C.c_uint16, C.c_uint16]
        L.kfo_fill_normal.argtypes = [_u16p, C.c_size_t, C.c_uint64, C.c_float, C.c_float]
        L.kfo_qrange_of.argtypes = [C.c_int, C.c_int, C.POINTER(QRange)]
        L.kfo_gama_elems.restype = C.c_size_t
        L.kfo_gama_elems.argtypes = [C.c_int, C.c_int, C.c_int]
        L.kfo_quantize.argtypes = [_u16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _u16p]
        L.kfo_pack_codes.argtypes = [_i32p, C.c_size_t, C.c_int, _u8p]
        L.kfo_unpack_codes.argtypes = [_u8p, C.c_size_t, C.c_int, _i32p]
        L.kfo_dequant.argtypes = [_u8p, _u16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u16p]
        L.kfo_f8e5m2_encode.argtypes = [_u16p, C.c_size_t, _u8p]
        L.kfo_f8e5m2_decode.argtypes = [_u8p, C.c_size_t, _u16p]
        L.kfo_linear.argtypes = [_u16p, _u16p, _u16p, C.c_int, C.c_int, C.c_int]
        L.kfo_linear_f32.argtypes = [_f32p, _u16p, _u16p, C.c_int, C.c_int, C.c_int]
        L.kfo_rmsnorm.argtypes = [_u16p, _u16p, _u16p, C.c_int, C.c_int, C.c_float]
        L.kfo_rope.argtypes = [_u16p, C.c_int, C.c_int, C.c_int, C.c_float]
        L.kfo_swiglu.argtypes = [_u16p, _u16p, _u16p, C.c_size_t]
        L.kfo_add.argtypes = [_u16p, _u16p, _u16p, C.c_size_t]
        _u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
        L.kfo_awq_dequant.argtypes = [_u32p, _u32p, _u16p, C.c_int, C.c_int, _u16p]
        L.kfo_awq_pack.argtypes = [_u16p, C.c_int, C.c_int, _u32p, _u32p, _u16p]
        L.kfo_nf4_quantize.argtypes = [_u16p, C.c_int, C.c_int, _u8p, _u16p]
        L.kfo_nf4_dequant.argtypes = [_u8p, _u16p, C.c_int, C.c_int, _u16p]
        L.kfo_sample.restype = C.c_int
        L.kfo_sample.argtypes = [_u16p, C.c_int, C.c_float, C.c_int, C.c_float, C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_int)]
        L.kfo_attention_decode.argtypes = [_u16p, _u16p, _u16p, _u16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.kfo_model_create.restype = C.c_void_p
        L.kfo_model_create.argtypes = [C.POINTER(ModelConfig)]
        L.kfo_model_destroy.argtypes = [C.c_void_p]
        L.kfo_model_reset.argtypes = [C.c_void_p]
        L.kfo_model_forward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.kfo_model_layer.argtypes = [C.c_void_p, C.c_int, C.c_int, _u16p]
        L.kfo_model_weight.restype = C.POINTER(C.c_uint16)
        L.kfo_model_weight.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]
        L.kfo_model_kcache.restype = C.POINTER(C.c_uint16)
        L.kfo_model_kcache.argtypes = [C.c_void_p, C.c_int]
        L.kfo_model_vcache.restype = C.POINTER(C.c_uint16)
        L.kfo_model_vcache.argtypes = [C.c_void_p, C.c_int]
        L.kfo_tensor_seed.restype = C.c_uint64
        L.kfo_tensor_seed.argtypes = [C.c_uint64, C.c_int]
        L.kfo_num_threads.restype = C.c_int
        L.kfo_set_dequant_fma.argtypes = [C.c_int]
        L.kfo_get_dequant_fma.restype = C.c_int
        L.kfo_bf16_fms.restype = C.c_uint16
        L.kfo_bf16_fms.argtypes = [C.c_uint16, C.c_uint16, C.c_uint16]
        _lib = L
    return _lib


def set_dequant_fma(fused):
    """1 (default): one bf16 rounding in the dequant, as the reference built for sm_90+; 0: two roundings (see koifish_oracle.h)."""
    lib().kfo_set_dequant_fma(int(fused))


_refq = None


def refq():
    """the reference's quantizer.cu compiled for sm_100a (CU_Q42X_NF4), or None when it was not built (reference tree absent)"""
    global _refq
    if _refq is None and os.path.exists(REFQ_SO):
        _refq = C.CDLL(REFQ_SO)
        _refq.refq_nf4_dequant.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _refq.refq_awq_dequant.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    return _refq


class RefTokenizer:
    """the reference's own HF_Tokenizer (oracle/ref_tokenizer.cpp) on a tokenizer.json text"""

    def __init__(self, json_text):
        self.lib = C.CDLL(REFTOK_SO)
        self.lib.reftok_load.restype = C.c_void_p
        self.lib.reftok_load.argtypes = [C.c_char_p]
        self.lib.reftok_encode.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.c_int]
        self.lib.reftok_decode.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_char_p, C.c_int]
        self.h = self.lib.reftok_load(json_text.encode())
        assert self.h, "the reference tokenizer refused the file"

    def encode(self, text):
        ids = (C.c_int * (4 * len(text.encode()) + 16))()
        n = self.lib.reftok_encode(self.h, text.encode(), ids, len(ids))
        assert n >= 0
        return list(ids[:n])

    def decode(self, ids, skip_special_tokens=False):
        a = (C.c_int * max(1, len(ids)))(*ids)
        buf = C.create_string_buffer(1 << 16)
        assert self.lib.reftok_decode(self.h, a, len(ids), int(skip_special_tokens), buf, len(buf)) >= 0
        return buf.value.decode("utf-8", errors="replace")


def reftok(json_text):
    """RefTokenizer, or None when oracle/_ref/libkoifish_reftok.so was never built"""
    build_oracle()
    return RefTokenizer(json_text) if os.path.exists(REFTOK_SO) else None


def refkun_read(path):
    """the reference's own reader on a fish.kun (K_SafeTensors::MMAP + GTensor::jDesc of what it parsed) -> {"config": ..., "tensors": [...]};
    None when oracle/_ref/libkoifish_refkun.so was never built.  The reader logs to stdout."""
    import json
    build_oracle()
    if not os.path.exists(REFKUN_SO):
        return None
    L = C.CDLL(REFKUN_SO)
    L.refcpu_kun_read.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    buf = C.create_string_buffer(1 << 22)
    n = L.refcpu_kun_read(str(path).encode(), buf, len(buf))
    assert n > 0, n
    return json.loads(buf.value.decode())


def refkun_write(path, config, tensors):
    """the reference's own WRITER (K_SafeTensors::Register + insertJS + Save) on host blobs; tensors as koifish_b200.kun_write takes them.  Returns
    False when oracle/_ref/libkoifish_refkun.so was never built."""
    import json
    build_oracle()
    if not os.path.exists(REFKUN_SO):
        return False
    L = C.CDLL(REFKUN_SO)
    L.refcpu_kun_write.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_longlong),
                                   C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), C.POINTER(C.c_void_p)]
    n = len(tensors)
    names = (C.c_char_p * n)(*[t[0].encode() for t in tensors])
    dtypes = (C.c_char_p * n)(*[t[1].encode() for t in tensors])
    shapes = (C.c_longlong * (2 * n))(*[d for t in tensors for d in (t[2][0], t[2][1] if len(t[2]) > 1 else 0)])
    szd = (C.c_ulonglong * n)(*[t[3] for t in tensors])
    szg = (C.c_ulonglong * n)(*[t[4] for t in tensors])
    keep = [np.frombuffer(bytes(t[5]), dtype=np.uint8).copy() for t in tensors]
    blobs = (C.c_void_p * n)(*[k.ctypes.data for k in keep])
    rc = L.refcpu_kun_write(str(path).encode(), json.dumps(config).encode() if config is not None else None, n, names, dtypes, shapes, szd, szg, blobs)
    assert rc == 0, rc
    return True


_refcpu = None


def refcpu():
    """the reference's own CPU packers compiled from src/Tensor/GeQuant.cpp (oracle/ref_cpu_quant.cpp), or None when the library was never built"""
    global _refcpu
    if _refcpu is None:
        build_oracle()
        if not os.path.exists(REFCPU_SO):
            return None
        _refcpu = C.CDLL(REFCPU_SO)
        _refcpu.refcpu_quantize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        _refcpu.refcpu_vendor2jsonx.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        _refcpu.refcpu_sample.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        _refcpu.refcpu_init4neuron.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_float)]
        _refcpu.refcpu_tochatml.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_char_p, C.c_int]
        _refcpu.refcpu_prefill_templates.argtypes = [C.c_int, C.c_char_p, C.c_int]
    return _refcpu


def refcpu_quantize(w, rows, cols, bits, group=128, mode=RTN_ASYM):
    """GeQuant::RTN_x / YinYang / RT_NormalF of the reference itself -> (data bytes, gama bf16 bits, qBias); mode as kfo_quantize (NF4 = 3)"""
    w = np.ascontiguousarray(w, dtype=np.uint16).reshape(-1)
    data = np.zeros(rows * cols * bits // 8, dtype=np.uint8)
    gama = np.zeros(rows + cols + (16 * rows if mode == NF4 else 2 * (rows * cols // group)), dtype=np.uint16)
    qb = C.c_int(0)
    rc = refcpu().refcpu_quantize(w.ctypes.data, rows, cols, bits, group, mode, data.ctypes.data, gama.ctypes.data, C.byref(qb))
    assert rc == 0, rc
    return data, gama, qb.value


def refcpu_vendor2jsonx(vendor_block):
    """QUANT_CARD::Vendor2JSONx of the reference itself: dict in, dict out (key order preserved)"""
    import json
    buf = C.create_string_buffer(1 << 16)
    n = refcpu().refcpu_vendor2jsonx(json.dumps(vendor_block).encode(), buf, len(buf))
    assert n >= 0
    return json.loads(buf.value.decode())


def refcpu_init4neuron(tensor_name, quantizer_block):
    """QUANT_CARD::Init4Neuron of the reference itself -> ([selected, mode, bits, group, yyang, sym, zero_point, vendor], T_errQ)"""
    import json
    out, errq = (C.c_int * 8)(), C.c_float(0)
    assert refcpu().refcpu_init4neuron(tensor_name.encode(), json.dumps(quantizer_block).encode(), out, C.byref(errq)) == 0
    return list(out), errq.value


def refcpu_sample(logits, temperature, top_k, top_p, state):
    """GeneratOnPrompt::Sample of the reference itself (LogitsInfo::TopK / UpdateLogits / TopP / Qu_FlipCoin); state as in sample()"""
    lg = np.ascontiguousarray(logits, dtype=np.uint16).reshape(-1)
    st, npick = C.c_uint64(state[0]), C.c_int(0)
    tok = refcpu().refcpu_sample(lg.ctypes.data, lg.size, temperature, top_k, top_p, C.byref(st), C.byref(npick))
    state[0] = st.value
    return int(tok), int(npick.value)


def refcpu_tochatml(lines, enable_thinking):
    """CHAT_SAMPLER::toChatML of the reference itself"""
    n = len(lines)
    roles = (C.c_char_p * max(1, n))(*[r.encode() for r, _ in lines])
    texts = (C.c_char_p * max(1, n))(*[c.encode() for _, c in lines])
    buf = C.create_string_buffer(1 << 16)
    assert refcpu().refcpu_tochatml(roles, texts, n, int(enable_thinking), buf, len(buf)) >= 0
    return buf.value.decode()


def refcpu_prefill_templates(enable_thinking):
    """CHAT_SAMPLER::InitPrefillTemplate of the reference itself -> (prompt_template, system_prompt_template), printf-style"""
    buf = C.create_string_buffer(1 << 12)
    assert refcpu().refcpu_prefill_templates(int(enable_thinking), buf, len(buf)) >= 0
    a, b = buf.value.decode().split("\x01")
    return a, b


def refgpu(variant="fma"):
    """The reference's own CUDA kernels (oracle/ref_kernels.cu); needs a GPU to call.  None when the library was never built."""
    if variant not in _refgpu:
        build_oracle()
        so = REFGPU_SO[variant]
        if not os.path.exists(so):
            return None
        R = C.CDLL(so)
        vp, i = C.c_void_p, C.c_int
        R.refk_q128tox.argtypes = [i, i, i, i, vp, vp, vp, vp]
        R.refk_xtoq128.argtypes = [i, i, i, i, i, i, i, i, vp, vp, vp, vp]
        R.refk_rmsnorm.argtypes = [vp, vp, vp, i, i]
        R.refk_rmsnorm_multihead.argtypes = [vp, vp, i, i, i]
        R.refk_rope2.argtypes = [vp, vp, i, i, i, i, C.c_float]
        R.refk_attention.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, i]
        R.refk_f8_decode.argtypes = [vp, vp, C.c_size_t]
        R.refk_f8_encode.argtypes = [vp, vp, C.c_size_t]
        _refgpu[variant] = R
    return _refgpu[variant]


def ref():
    """The reference's own code (PackedQ.hpp macros, GST_float.cpp primitives); None when it was never built."""
    global _ref
    if _ref is None:
        build_oracle()
        if not os.path.exists(REF_SO):
            return None
        R = C.CDLL(REF_SO)
        R.ref_pack.argtypes = [_i32p, C.c_size_t, C.c_int, _u8p]
        R.ref_unpack.argtypes = [_u8p, C.c_size_t, C.c_int, _i32p]
        R.ref_matvec_f32.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int]
        R.ref_rmsnorm_f32.restype = C.c_float
        R.ref_rmsnorm_f32.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_float]
        _ref = R
    return _ref


# ---------------------------------------------------------------- numpy-level helpers
def bf16_to_f32(a):
    a = np.ascontiguousarray(a, dtype=np.uint16)
    return (a.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16(a):
    """round-to-nearest-even, vectorised (matches kfo_f32_to_bf16 for finite values)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    lsb = (u >> 16) & 1
    return ((u + 0x7FFF + lsb) >> 16).astype(np.uint16)


def fill_normal(n, seed, sigma=0.02, mean=0.0):
    out = np.empty(n, dtype=np.uint16)
    lib().kfo_fill_normal(out, n, seed, sigma, mean)
    return out


def qrange(bits, mode):
    r = QRange()
    assert lib().kfo_qrange_of(bits, mode, C.byref(r)) == 0
    return r.qmin, r.qmax, r.qbias


def quantize(w, rows, cols, bits, group=128, mode=RTN_ASYM):
    w = np.ascontiguousarray(w, dtype=np.uint16).reshape(-1)
    data = np.zeros(rows * cols * bits // 8, dtype=np.uint8)
    gama = np.zeros(lib().kfo_gama_elems(rows, cols, group), dtype=np.uint16)
    rc = lib().kfo_quantize(w, rows, cols, bits, group, mode, data, gama)
    assert rc == 0, rc
    return data, gama


def awq_pack(w_in_out, M, N):
    """test helper: bf16 [M = in_features][N = out_features] -> (qweight int32 [M][N/8], qzeros int32 [M/128][N/8], scales fp16 [M/128][N])"""
    w = np.ascontiguousarray(w_in_out, dtype=np.uint16).reshape(-1)
    qw, qz, sc = np.zeros(M * N // 8, np.uint32), np.zeros(M // 128 * N // 8, np.uint32), np.zeros(M // 128 * N, np.uint16)
    assert lib().kfo_awq_pack(w, M, N, qw, qz, sc) == 0
    return qw, qz, sc


def awq_dequant(qw, qz, sc, M, N):
    out = np.zeros(M * N, dtype=np.uint16)
    assert lib().kfo_awq_dequant(np.ascontiguousarray(qw), np.ascontiguousarray(qz), np.ascontiguousarray(sc), M, N, out) == 0
    return out.reshape(M, N)


def nf4_quantize(w, rows, cols):
    w = np.ascontiguousarray(w, dtype=np.uint16).reshape(-1)
    data = np.zeros(rows * cols // 2, dtype=np.uint8)
    gama = np.zeros(rows + cols + 16 * rows, dtype=np.uint16)
    assert lib().kfo_nf4_quantize(w, rows, cols, data, gama) == 0
    return data, gama


def nf4_dequant(data, gama, rows, cols):
    out = np.zeros(rows * cols, dtype=np.uint16)
    assert lib().kfo_nf4_dequant(np.ascontiguousarray(data), np.ascontiguousarray(gama), rows, cols, out) == 0
    return out.reshape(rows, cols)


def pack_codes(codes, bits):
    codes = np.ascontiguousarray(codes, dtype=np.int32)
    out = np.zeros(codes.size * bits // 8, dtype=np.uint8)
    assert lib().kfo_pack_codes(codes, codes.size, bits, out) == 0
    return out


def unpack_codes(data, n, bits):
    out = np.zeros(n, dtype=np.int32)
    assert lib().kfo_unpack_codes(np.ascontiguousarray(data, dtype=np.uint8), n, bits, out) == 0
    return out


def dequant(data, gama, rows, cols, bits, group=128, qbias=0):
    out = np.zeros(rows * cols, dtype=np.uint16)
    rc = lib().kfo_dequant(np.ascontiguousarray(data), np.ascontiguousarray(gama), rows, cols, bits, group, qbias, out)
    assert rc == 0, rc
    return out.reshape(rows, cols)


def f8_encode(w):
    w = np.ascontiguousarray(w, dtype=np.uint16).reshape(-1)
    out = np.zeros(w.size, dtype=np.uint8)
    lib().kfo_f8e5m2_encode(w, w.size, out)
    return out


def f8_decode(b):
    b = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1)
    out = np.zeros(b.size, dtype=np.uint16)
    lib().kfo_f8e5m2_decode(b, b.size, out)
    return out


def linear(w, x, M, N, K):
    y = np.zeros((M, N), dtype=np.uint16)
    lib().kfo_linear(y, np.ascontiguousarray(w, dtype=np.uint16), np.ascontiguousarray(x, dtype=np.uint16), M, N, K)
    return y


def linear_f32(w, x, M, N, K):
    y = np.zeros((M, N), dtype=np.float32)
    lib().kfo_linear_f32(y, np.ascontiguousarray(w, dtype=np.uint16), np.ascontiguousarray(x, dtype=np.uint16), M, N, K)
    return y


def rmsnorm(x, w, rows, dim, eps=1e-6):
    out = np.zeros(rows * dim, dtype=np.uint16)
    lib().kfo_rmsnorm(out, np.ascontiguousarray(x, dtype=np.uint16).reshape(-1), np.ascontiguousarray(w, dtype=np.uint16), rows, dim, eps)
    return out.reshape(rows, dim)


def rope(v, n_heads, head_dim, pos, theta):
    v = np.array(v, dtype=np.uint16, copy=True).reshape(-1)
    lib().kfo_rope(v, n_heads, head_dim, pos, theta)
    return v.reshape(n_heads, head_dim)


def swiglu(g, u):
    g = np.ascontiguousarray(g, dtype=np.uint16).reshape(-1)
    u = np.ascontiguousarray(u, dtype=np.uint16).reshape(-1)
    out = np.zeros(g.size, dtype=np.uint16)
    lib().kfo_swiglu(out, g, u, g.size)
    return out


def sample(logits, temperature, top_k, top_p, state, selection=0):
    """CPU port of GeneratOnPrompt::Sample; state: [uint64] advanced in place (a one-element list); returns (token, n_pick)"""
    lg = np.ascontiguousarray(logits, dtype=np.uint16).reshape(-1)
    st, npick = C.c_uint64(state[0]), C.c_int(0)
    tok = lib().kfo_sample(lg, lg.size, temperature, top_k, top_p, C.byref(st), selection, C.byref(npick))
    state[0] = st.value
    return int(tok), int(npick.value)


def add(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint16).reshape(-1)
    b = np.ascontiguousarray(b, dtype=np.uint16).reshape(-1)
    out = np.zeros(a.size, dtype=np.uint16)
    lib().kfo_add(out, a, b, a.size)
    return out


def attention_decode(q, kc, vc, pos, n_head, n_kv, hd, score_bf16=0):
    out = np.zeros(n_head * hd, dtype=np.uint16)
    lib().kfo_attention_decode(out, np.ascontiguousarray(q, dtype=np.uint16).reshape(-1), np.ascontiguousarray(kc, dtype=np.uint16).reshape(-1),
                               np.ascontiguousarray(vc, dtype=np.uint16).reshape(-1), pos, n_head, n_kv, hd, score_bf16)
    return out.reshape(n_head, hd)


def ref_pack(codes, bits):
    codes = np.ascontiguousarray(codes, dtype=np.int32)
    out = np.zeros(codes.size * bits // 8, dtype=np.uint8)
    assert ref().ref_pack(codes, codes.size, bits, out) == 0
    return out


def ref_unpack(data, n, bits):
    out = np.zeros(n, dtype=np.int32)
    assert ref().ref_unpack(np.ascontiguousarray(data, dtype=np.uint8), n, bits, out) == 0
    return out


def tensor_seed(model_seed, tid):
    return lib().kfo_tensor_seed(model_seed, tid)


class OracleModel:
    """Whole-model CPU decode following the reference op order."""

    def __init__(self, **kw):
        d = dict(n_layer=2, n_embd=256, n_ff=512, n_head=4, n_kv_head=2, head_dim=64, vocab=1024, max_seq=64,
                 rope_theta=1e6, rms_eps=1e-6, tie_embed=1, attn_bits=4, attn_mode=RTN_ASYM, mlp_bits=4, mlp_mode=RTN_ASYM,
                 embed_bits=16, embed_mode=RTN_ASYM, group=128, seed=42, sigma=0.02, norm_sigma=0.0, score_bf16=0)
        d.update(kw)
        self.cfg = ModelConfig(**d)
        self.h = lib().kfo_model_create(C.byref(self.cfg))
        assert self.h

    def close(self):
        if self.h:
            lib().kfo_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        lib().kfo_model_reset(self.h)

    def forward(self, token, pos, want_logits=True):
        if want_logits:
            out = np.zeros(self.cfg.vocab, dtype=np.uint16)
            rc = lib().kfo_model_forward(self.h, token, pos, out.ctypes.data_as(C.c_void_p))
            assert rc == 0
            return out
        assert lib().kfo_model_forward(self.h, token, pos, None) == 0
        return None

    def layer(self, layer, pos, x):
        x = np.array(x, dtype=np.uint16, copy=True)
        assert lib().kfo_model_layer(self.h, layer, pos, x) == 0
        return x

    def weight(self, tid):
        n = C.c_size_t()
        p = lib().kfo_model_weight(self.h, tid, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def kcache(self, layer, npos):
        kd = self.cfg.n_kv_head * self.cfg.head_dim
        return np.ctypeslib.as_array(lib().kfo_model_kcache(self.h, layer), shape=(npos * kd,)).copy().reshape(npos, kd)

    def vcache(self, layer, npos):
        kd = self.cfg.n_kv_head * self.cfg.head_dim
        return np.ctypeslib.as_array(lib().kfo_model_vcache(self.h, layer), shape=(npos * kd,)).copy().reshape(npos, kd)
