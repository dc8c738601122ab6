"""koifish_b200 -- B200-native (sm_100a) quantized-inference hot path of Koifish: fused unpack+dequant matmuls over PackedQ
weights, RMSNorm, RoPE, GQA decode attention over the KV cache, behind a C ABI (include/kf_device.h, include/kf_model.h).
The package is a ctypes view of libkoifish_b200.so; it has no CPU fallback."""
from ._lib import (KF_EPI_F32, KF_EPI_NONE, KF_EPI_RESIDUAL, KF_ERR_BAD_ARG, KF_ERR_NO_DEVICE, KF_ERR_UNSUPPORTED, KF_OK, KF_Q_RTN_ASYM,  # noqa: F401
                   KF_Q_RTN_SYM, KF_Q_YYANG, KF_T_BF16, KF_T_BINARY, KF_T_F8E5M2, KF_T_Q2, KF_T_Q4, KF_T_SIGN, KF_T_NF4, KF_T_AWQ4, LIB_PATH, SIGNATURES,
                   TYPE_BITS, KoifishError, load)
from .api import (QWEN3_DIMS, AwqTensor, Context, DevArray, Model, ModelInfo, QTensor, TensorDesc, add, argmax, sample, linear_axb, attn_decode, attn_decode_gqa, attn_prefill, dequant, embed,  # noqa: F401
                  fill_normal, fill_normal_2d, linear, linear_multi, linear_swiglu, qknorm_rope_kvappend, qkv_attention, quantize, qwen3_config,
                  rmsnorm, rmsnorm_linear, rope_table, safetensors_index, safetensors_read_bf16, swiglu, Tokenizer, nfc, chatml_prompt, chatml_render, chat_once, from_pretrained, kun_index, kun_config, kun_write)
