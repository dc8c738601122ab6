"""ctypes binding of the C ABI declared in include/kf_device.h and include/kf_model.h.

The shared library is built in-tree by koifish_b200.build (nvcc, sm_100a).  There is no CPU fallback: if the library is missing
it is (re)built, and if no CUDA device is present every compute entry point returns KF_ERR_NO_DEVICE, surfaced as KoifishError.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KF_LIB_PATH") or os.path.join(HERE, "libkoifish_b200.so")  # KF_LIB_PATH: tuning builds only

KF_OK = 0
KF_ERR_NO_DEVICE, KF_ERR_CUDA, KF_ERR_BAD_ARG, KF_ERR_UNSUPPORTED, KF_ERR_OOM, KF_ERR_NCCL, KF_ERR_QUANT = -100, -101, -102, -103, -104, -105, -701
KF_T_BF16, KF_T_F8E5M2, KF_T_Q4, KF_T_Q2, KF_T_SIGN, KF_T_BINARY, KF_T_NF4, KF_T_AWQ4 = 0, 1, 2, 3, 4, 5, 6, 7
KF_Q_RTN_ASYM, KF_Q_RTN_SYM, KF_Q_YYANG = 0, 1, 2
KF_EPI_NONE, KF_EPI_RESIDUAL, KF_EPI_F32 = 0, 1, 4
TYPE_BITS = {KF_T_BF16: 16, KF_T_F8E5M2: 8, KF_T_Q4: 4, KF_T_Q2: 2, KF_T_SIGN: 2, KF_T_BINARY: 1, KF_T_NF4: 4, KF_T_AWQ4: 4}


class KoifishError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        super().__init__("%s failed: %s%s" % (where, _status_string(status), (" -- " + detail) if detail else ""))


class TensorDesc(C.Structure):
    _fields_ = [("data_dev", C.c_void_p), ("gama_dev", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("type", C.c_int),
                ("group", C.c_int), ("qbias", C.c_int), ("zero_dev", C.c_void_p), ("step_dev", C.c_void_p)]


class ModelInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("n_layers", "n_embd", "n_ff", "n_head", "n_head_kv", "head_dim", "vocab", "max_seq_len", "max_batch",
                                       "max_tokens", "tp_rank", "tp_world", "tie_word_embeddings")] + \
               [("rope_theta", C.c_float), ("norm_rms_eps", C.c_float)] + \
               [(n, C.c_uint64) for n in ("weight_bytes", "kv_bytes", "block_weight_bytes_per_layer", "head_weight_bytes")]


# every symbol include/*.h declares, with its signature: (restype, [argtypes])
_P, _I, _SZ, _U64, _F = C.c_void_p, C.c_int, C.c_size_t, C.c_uint64, C.c_float
_DESCP = C.POINTER(TensorDesc)
SIGNATURES = {
    # kf_device.h
    "kf_ctx_create": (_I, [_I, _P, C.POINTER(_P)]),
    "kf_ctx_destroy": (_I, [_P]),
    "kf_ctx_sync": (_I, [_P]),
    "kf_ctx_make_current": (_I, [_P]),
    "kf_ctx_stream": (_P, [_P]),
    "kf_ctx_sm_count": (_I, [_P]),
    "kf_status_string": (C.c_char_p, [_I]),
    "kf_last_error": (C.c_char_p, [_P]),
    "kf_launch_count": (_U64, [_P]),
    "kf_scratch_generation": (_U64, [_P]),
    "kf_ctx_get_int": (_I, [_P, C.c_char_p, _P]),
    "kf_ctx_set_int": (_I, [_P, C.c_char_p, _I]),
    "kf_malloc": (_I, [_P, _SZ, C.POINTER(_P)]),
    "kf_free": (_I, [_P, _P]),
    "kf_memset": (_I, [_P, _P, _I, _SZ]),
    "kf_h2d": (_I, [_P, _P, _P, _SZ]),
    "kf_d2h": (_I, [_P, _P, _P, _SZ]),
    "kf_d2d": (_I, [_P, _P, _P, _SZ]),
    "kf_host_alloc": (_I, [_SZ, C.POINTER(_P)]),
    "kf_host_free": (_I, [_P]),
    "kf_graph_begin": (_I, [_P]),
    "kf_graph_end": (_I, [_P, C.POINTER(_P)]),
    "kf_graph_launch": (_I, [_P, _P]),
    "kf_graph_destroy": (_I, [_P]),
    "kf_fill_normal": (_I, [_P, _P, _SZ, _U64, _F, _F]),
    "kf_fill_normal_2d": (_I, [_P, _P, _I, _I, _SZ, _SZ, _SZ, _U64, _F, _F]),
    "kf_quant_data_bytes": (_SZ, [_I, _I, _I]),
    "kf_quant_gama_bytes": (_SZ, [_I, _I, _I, _I]),
    "kf_quantize": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, C.POINTER(_I)]),
    "kf_dequant": (_I, [_P, _DESCP, _P]),
    "kf_linear": (_I, [_P, _P, _DESCP, _P, _I, _I, _P]),
    "kf_linear_multi": (_I, [_P, _I, C.POINTER(_P), _DESCP, _P, _I]),
    "kf_linear_swiglu": (_I, [_P, _P, _DESCP, _DESCP, _P, _I]),
    "kf_rmsnorm": (_I, [_P, _P, _P, _P, _I, _I, _F]),
    "kf_rmsnorm_linear": (_I, [_P, _I, C.POINTER(_P), _DESCP, _P, _P, _F, _I, _I]),
    "kf_rope_table": (_I, [_P, _P, _I, _I, _F]),
    "kf_qknorm_rope_kvappend": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _SZ]),
    "kf_attn_decode": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _SZ]),
    "kf_p2p_alloc": (_I, [_P, _SZ, _I, _P]),
    "kf_p2p_attach": (_I, [_P, _P, _I, _I]),
    "kf_p2p_ready": (_I, [_P]),
    "kf_p2p_release": (_I, [_P]),
    "kf_allreduce_residual": (_I, [_P, _P, _P, _P, _SZ]),
    "kf_relayout_wmv": (_I, [_P, _P, _P, _I, _I, _I]),
    "kf_tp_begin": (_I, [_P]),
    "kf_exchange_fused_ready": (_I, [_P, _I, _I]),
    "kf_linear_exchange": (_I, [_P, _P, _P, _P, _I, _P]),
    "kf_attn_prefill": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I]),
    "kf_attn_decode_gqa": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _SZ]),
    "kf_qkv_attention": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _SZ, _I]),
    "kf_swiglu": (_I, [_P, _P, _P, _P, _SZ]),
    "kf_add": (_I, [_P, _P, _P, _P, _SZ]),
    "kf_residual_add_f32": (_I, [_P, _P, _P, _P, _SZ]),
    "kf_advance_pos": (_I, [_P, _P, _I]),
    "kf_embed": (_I, [_P, _P, _DESCP, _P, _I]),
    "kf_linear_axb": (_I, [_P, _P, _P, _P, _I, C.c_float, C.c_float, _P]),
    "kf_argmax": (_I, [_P, _P, _P, _I, _I]),
    "kf_sample": (_I, [_P, _P, _P, _I, _I, C.c_float, _I, C.c_float, _P, _I]),
    "kf_nccl_unique_id": (_I, [_P]),
    "kf_ctx_init_nccl": (_I, [_P, _P, _I, _I]),
    "kf_allreduce_bf16": (_I, [_P, _P, _SZ]),
    "kf_allreduce_f32": (_I, [_P, _P, _SZ]),
    "kf_allgather": (_I, [_P, _P, _P, _SZ]),
    # kf_model.h
    "kf_model_create": (_I, [_P, C.c_char_p, _I, _I, C.POINTER(_P), C.POINTER(C.c_char_p)]),
    "kf_model_destroy": (_I, [_P]),
    "kf_model_error": (C.c_char_p, [_P]),
    "kf_string_free": (None, [C.c_char_p]),
    "kf_model_info_get": (_I, [_P, C.POINTER(ModelInfo)]),
    "kf_model_init_random": (_I, [_P]),
    "kf_model_set_tensor": (_I, [_P, C.c_char_p, _P, _I, _I]),
    "kf_model_set_tensor_awq": (_I, [_P, C.c_char_p, _P, _P, _P, _I, _I]),
    "kf_model_tensor_desc": (_I, [_P, C.c_char_p, _DESCP]),
    "kf_model_tensor_count": (_I, [_P]),
    "kf_model_tensor_name": (C.c_char_p, [_P, _I]),
    "kf_model_kcache": (_P, [_P, _I]),
    "kf_model_vcache": (_P, [_P, _I]),
    "kf_model_forward": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "kf_model_decode_loop": (_I, [_P, _I, _I]),
    "kf_model_read_state": (_I, [_P, _P, _P, _I]),
    "kf_model_save": (_I, [_P, C.c_char_p]),
    "kf_model_load": (_I, [_P, C.c_char_p]),
    "kf_model_generate": (_I, [_P, _P, _I, _I, _I, _I, _P, C.POINTER(_I), C.POINTER(_I)]),
    "kf_model_save_kun": (_I, [_P, C.c_char_p]),
    "kf_model_load_kun": (_I, [_P, C.c_char_p, C.POINTER(_I), C.POINTER(_I)]),
    "kf_kun_index": (_I, [C.c_char_p, C.POINTER(_P), C.POINTER(_P)]),
    "kf_kun_config": (_I, [C.c_char_p, C.POINTER(_P), C.POINTER(_P)]),
    "kf_kun_write": (_I, [C.c_char_p, C.c_char_p, _I, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(_U64), C.POINTER(_U64),
                          C.POINTER(_P), C.POINTER(_P)]),
    "kf_model_set_graphs": (_I, [_P, _I]),
    "kf_model_load_safetensors": (_I, [_P, C.c_char_p, C.POINTER(_I), C.POINTER(_I)]),
    "kf_safetensors_index": (_I, [C.c_char_p, C.POINTER(_P), C.POINTER(_P)]),
    "kf_safetensors_read_bf16": (_I, [C.c_char_p, C.c_char_p, _P, _SZ, C.POINTER(_P)]),
    "kf_model_set_sampler": (_I, [_P, C.c_float, _I, C.c_float, _U64, _I]),
    "kf_tokenizer_load": (_I, [C.c_char_p, C.POINTER(_P), C.POINTER(_P)]),
    "kf_tokenizer_from_json": (_I, [C.c_char_p, C.c_char_p, C.POINTER(_P), C.POINTER(_P)]),
    "kf_tokenizer_destroy": (_I, [_P]),
    "kf_tokenizer_encode": (_I, [_P, C.c_char_p, _SZ, _P, _SZ, C.POINTER(_SZ)]),
    "kf_tokenizer_decode": (_I, [_P, _P, _SZ, _I, C.POINTER(_P)]),
    "kf_decode_stream_create": (_I, [C.POINTER(_P)]),
    "kf_decode_stream_destroy": (_I, [_P]),
    "kf_decode_stream_push": (_I, [_P, _P, _I, _I, C.POINTER(_P)]),
    "kf_decode_stream_flush": (_I, [_P, C.POINTER(_P)]),
    "kf_tokenizer_token_to_id": (_I, [_P, C.c_char_p]),
    "kf_tokenizer_id_to_token": (_I, [_P, _I, C.POINTER(_P)]),
    "kf_tokenizer_vocab_size": (_I, [_P]),
    "kf_tokenizer_eos_id": (_I, [_P]),
    "kf_tokenizer_bos_id": (_I, [_P]),
    "kf_tokenizer_pad_id": (_I, [_P]),
    "kf_tokenizer_is_special": (_I, [_P, _I]),
    "kf_text_nfc": (_I, [C.c_char_p, _SZ, C.POINTER(_P)]),
    "kf_tokenizer_pre_tokenize": (_I, [_P, C.c_char_p, _SZ, C.POINTER(_P)]),
    "kf_chatml_prompt": (_I, [C.c_char_p, C.c_char_p, _I, C.POINTER(_P)]),
    "kf_chatml_render": (_I, [C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), _I, _I, C.POINTER(_P)]),
    "kf_config_dims": (_I, [C.c_char_p, C.POINTER(ModelInfo), C.POINTER(_P)]),
    "kf_config_quantizer_json": (_I, [C.c_char_p, C.POINTER(_P), C.POINTER(_P)]),
    "kf_config_quant_card": (_I, [C.c_char_p, C.c_char_p, C.POINTER(_I), C.POINTER(C.c_float), C.POINTER(_P)]),
    "kf_config_quant_of": (_I, [C.c_char_p, C.c_char_p, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_P)]),
    "kf_config_awq_shard": (_I, [C.c_char_p, C.c_char_p, _I, _I, _P, _P, _P, _P, _SZ, C.POINTER(_SZ), C.POINTER(_P)]),
    "kf_config_shard_of": (_I, [C.c_char_p, C.c_char_p, _I, _I, C.POINTER(_I), C.POINTER(_P)]),
}

_lib = None


def load(build_if_missing=True):
    """dlopen the in-tree library and bind every declared symbol (AttributeError if one is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise KoifishError(KF_ERR_UNSUPPORTED, "load", "libkoifish_b200.so has not been built (python -m koifish_b200.build)")
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    # string-returning functions that hand out malloc'ed memory must keep the raw pointer
    lib.kf_model_create.argtypes = [_P, C.c_char_p, _I, _I, C.POINTER(_P), C.POINTER(_P)]
    lib.kf_string_free.argtypes = [_P]
    _lib = lib
    return lib


def _status_string(status):
    try:
        return load().kf_status_string(status).decode()
    except Exception:
        return "status %d" % status


def check(status, where, ctx=None):
    if status != KF_OK:
        detail = ""
        if ctx:
            try:
                detail = load().kf_last_error(ctx).decode()
            except Exception:
                pass
        raise KoifishError(status, where, detail)
