"""Thin Python view of the C ABI: device context, device arrays, quantised tensors, the hot-path ops and the Qwen3 runtime.
All compute happens in libkoifish_b200.so (hand-written sm_100a CUDA); numpy is only the host container.
bf16 data is carried as uint16 bit patterns."""
import ctypes as C
import json

import numpy as np

from . import _lib as L
from ._lib import KoifishError, TensorDesc, ModelInfo  # noqa: F401


class Context:
    """kf_ctx: one device + one stream (reference: InitCUDA / main_stream, src/Device/CUDA/QKV.cu:501-571)."""

    def __init__(self, device=0, stream=None):
        self.lib = L.load()
        h = C.c_void_p()
        st = self.lib.kf_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h))
        if st != L.KF_OK:
            raise KoifishError(st, "kf_ctx_create")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.kf_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, st, where):
        L.check(st, where, self.h)

    def sync(self):
        self.check(self.lib.kf_ctx_sync(self.h), "kf_ctx_sync")

    @property
    def stream(self):
        return self.lib.kf_ctx_stream(self.h)

    @property
    def sm_count(self):
        return self.lib.kf_ctx_sm_count(self.h)

    @property
    def launches(self):
        return int(self.lib.kf_launch_count(self.h))

    def set_int(self, key, value):
        self.check(self.lib.kf_ctx_set_int(self.h, key.encode(), int(value)), "kf_ctx_set_int")

    def get_int(self, key):
        v = C.c_int(0)
        self.check(self.lib.kf_ctx_get_int(self.h, key.encode(), C.byref(v)), "kf_ctx_get_int")
        return int(v.value)

    def init_tensor_parallel(self, rank, world, max_floats=64 * 8192, p2p=True):
        """NCCL communicator + peer-memory exchange buffers for this rank; the ids / IPC handles travel through torch.distributed
        (one process per GPU, process group already initialised)."""
        import torch
        import torch.distributed as dist
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = (C.c_ubyte * 128)()
            self.check(self.lib.kf_nccl_unique_id(raw), "kf_nccl_unique_id")
            idbuf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        raw = (C.c_ubyte * 128)(*idbuf.cpu().tolist())
        self.check(self.lib.kf_ctx_init_nccl(self.h, raw, rank, world), "kf_ctx_init_nccl")
        if world <= 8 and p2p:
            # peer-memory exchange buffers; if any rank cannot map its peers (no NVLink / IPC), every rank drops them and the exchange
            # takes the NCCL path -- the decision is collective, so the ranks never disagree on the kernel they run
            h = (C.c_ubyte * 64)()
            ok = self.lib.kf_p2p_alloc(self.h, max_floats, world, h) == L.KF_OK
            mine = torch.tensor(list(h), dtype=torch.uint8, device="cuda")
            allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(allh, mine)
            flat = (C.c_ubyte * (64 * world))(*[b for t in allh for b in t.cpu().tolist()])
            ok = ok and self.lib.kf_p2p_attach(self.h, flat, rank, world) == L.KF_OK
            flag = torch.tensor([1 if ok else 0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                self.lib.kf_p2p_release(self.h)
            dist.barrier()

    # ---- memory
    def empty(self, nbytes):
        return DevArray(self, nbytes)

    def array(self, a):
        a = np.ascontiguousarray(a)
        d = DevArray(self, a.nbytes)
        d.copy_from(a)
        return d

    def zeros(self, nbytes):
        d = DevArray(self, nbytes)
        self.check(self.lib.kf_memset(self.h, d.ptr, 0, nbytes), "kf_memset")
        return d


class DevArray:
    def __init__(self, ctx, nbytes):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        ctx.check(ctx.lib.kf_malloc(ctx.h, self.nbytes, C.byref(p)), "kf_malloc(%d)" % nbytes)
        self.ptr = p.value

    def free(self):
        if getattr(self, "ptr", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.kf_free(self.ctx.h, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def copy_from(self, a):
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        self.ctx.check(self.ctx.lib.kf_h2d(self.ctx.h, self.ptr, a.ctypes.data, a.nbytes), "kf_h2d")
        self.ctx.sync()  # pageable host memory: make the copy complete before `a` can go away

    def numpy(self, dtype=np.uint16, shape=None, offset=0, count=None):
        dtype = np.dtype(dtype)
        n = (self.nbytes - offset) // dtype.itemsize if count is None else count
        out = np.empty(n, dtype=dtype)
        self.ctx.check(self.ctx.lib.kf_d2h(self.ctx.h, out.ctypes.data, self.ptr + offset, out.nbytes), "kf_d2h")
        self.ctx.sync()
        return out.reshape(shape) if shape is not None else out


class QTensor:
    """A weight in the reference's storage form: packed data || gama."""

    def __init__(self, ctx, rows, cols, type_, group=128, qbias=0):
        self.ctx, self.rows, self.cols, self.type, self.group, self.qbias = ctx, rows, cols, type_, group, qbias
        self.data_bytes = ctx.lib.kf_quant_data_bytes(rows, cols, type_)
        self.gama_bytes = ctx.lib.kf_quant_gama_bytes(rows, cols, type_, group)
        self.blob = ctx.empty(self.data_bytes + self.gama_bytes + 16)

    @property
    def data_ptr(self):
        return self.blob.ptr

    @property
    def gama_ptr(self):
        return self.blob.ptr + self.data_bytes if self.gama_bytes else None

    def desc(self):
        return TensorDesc(self.data_ptr, self.gama_ptr, self.rows, self.cols, self.type, self.group, self.qbias, None, None)

    def data_numpy(self):
        return self.blob.numpy(np.uint8, count=self.data_bytes)

    def gama_numpy(self):
        return self.blob.numpy(np.uint16, offset=self.data_bytes, count=self.gama_bytes // 2)

    @property
    def nbytes(self):
        return self.data_bytes + self.gama_bytes

    @classmethod
    def from_packed(cls, ctx, data, gama, rows, cols, type_, group=128, qbias=0):
        """wrap bytes produced elsewhere (e.g. by the test oracle's packer)"""
        t = cls(ctx, rows, cols, type_, group, qbias)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        assert data.nbytes == t.data_bytes
        ctx.check(ctx.lib.kf_h2d(ctx.h, t.data_ptr, data.ctypes.data, data.nbytes), "kf_h2d")
        if t.gama_bytes:
            gama = np.ascontiguousarray(gama, dtype=np.uint16)
            assert gama.nbytes == t.gama_bytes
            ctx.check(ctx.lib.kf_h2d(ctx.h, t.gama_ptr, gama.ctypes.data, gama.nbytes), "kf_h2d")
        ctx.sync()
        return t


class AwqTensor:
    """A weight in the vendor AWQ layout (KF_T_AWQ4): qweight int32 [in][out / 8], qzeros int32 [in / 128][out / 8], scales fp16 [in / 128][out]."""

    def __init__(self, ctx, qweight, qzeros, scales_f16, in_features, out_features):
        self.ctx, self.rows, self.cols, self.type, self.group, self.qbias = ctx, out_features, in_features, L.KF_T_AWQ4, 128, 0
        self.qw = ctx.array(np.ascontiguousarray(qweight, dtype=np.uint32).view(np.uint16))
        self.qz = ctx.array(np.ascontiguousarray(qzeros, dtype=np.uint32).view(np.uint16))
        self.sc = ctx.array(np.ascontiguousarray(scales_f16, dtype=np.uint16))

    def desc(self):
        return TensorDesc(self.qw.ptr, None, self.rows, self.cols, self.type, self.group, 0, self.qz.ptr, self.sc.ptr)


def fill_normal(ctx, n, seed, sigma=0.02, mean=0.0):
    d = ctx.empty(n * 2)
    ctx.check(ctx.lib.kf_fill_normal(ctx.h, d.ptr, n, seed, sigma, mean), "kf_fill_normal")
    return d


def fill_normal_2d(ctx, rows, cols, ld, row0, col0, seed, sigma=0.02, mean=0.0):
    d = ctx.empty(rows * cols * 2)
    ctx.check(ctx.lib.kf_fill_normal_2d(ctx.h, d.ptr, rows, cols, ld, row0, col0, seed, sigma, mean), "kf_fill_normal_2d")
    return d


def quantize(ctx, w_dev, rows, cols, type_, group=128, mode=L.KF_Q_RTN_ASYM):
    """GeQuant::LowBit_worker(flag 0x100) on the device (reference src/Tensor/GeQuant.cpp:830-905)."""
    t = QTensor(ctx, rows, cols, type_, group)
    qb = C.c_int(0)
    ctx.check(ctx.lib.kf_quantize(ctx.h, w_dev.ptr, rows, cols, type_, group, mode, t.data_ptr, t.gama_ptr, C.byref(qb)), "kf_quantize")
    t.qbias = qb.value
    return t


def dequant(ctx, t):
    """GTensor::GetDataX test hook -> DevArray of bf16 [rows, cols]"""
    out = ctx.empty(t.rows * t.cols * 2)
    d = t.desc()
    ctx.check(ctx.lib.kf_dequant(ctx.h, C.byref(d), out.ptr), "kf_dequant")
    return out


def linear(ctx, w, x, M, epilogue=L.KF_EPI_NONE, residual=None, out=None):
    """TASKA_AxB::blasLt / SLP::Forw: y[M][rows] = x[M][cols] . deq(W)^T"""
    if out is None:
        out = ctx.empty(M * w.rows * (4 if epilogue == L.KF_EPI_F32 else 2))
    d = w.desc()
    ctx.check(ctx.lib.kf_linear(ctx.h, out.ptr, C.byref(d), x.ptr, M, epilogue, residual.ptr if residual is not None else None), "kf_linear")
    return out


def linear_axb(ctx, w, x, M, d, alpha=1.0, beta=0.0, bias=None):
    """TASKA_AxB in full: d = alpha * x . deq(W)^T + beta * d + bias (in place on the device buffer d)"""
    desc = w.desc()
    ctx.check(ctx.lib.kf_linear_axb(ctx.h, d.ptr, C.byref(desc), x.ptr, M, float(alpha), float(beta), bias.ptr if bias is not None else None), "kf_linear_axb")
    return d


def linear_multi(ctx, ws, x, M):
    outs = [ctx.empty(M * w.rows * 2) for w in ws]
    descs = (TensorDesc * len(ws))(*[w.desc() for w in ws])
    ys = (C.c_void_p * len(ws))(*[o.ptr for o in outs])
    ctx.check(ctx.lib.kf_linear_multi(ctx.h, len(ws), ys, descs, x.ptr, M), "kf_linear_multi")
    return outs


def linear_swiglu(ctx, wg, wu, x, M):
    out = ctx.empty(M * wg.rows * 2)
    dg, du = wg.desc(), wu.desc()
    ctx.check(ctx.lib.kf_linear_swiglu(ctx.h, out.ptr, C.byref(dg), C.byref(du), x.ptr, M), "kf_linear_swiglu")
    return out


def rmsnorm_linear(ctx, ws, x, norm_w, M, eps=1e-6, swiglu=False):
    """RMSNorm folded into the matmul(s) consuming it (norm -> Q/K/V, norm -> gate/up+SwiGLU, final norm -> lm_head)"""
    outs = [ctx.empty(M * ws[0].rows * 2)] if swiglu else [ctx.empty(M * w.rows * 2) for w in ws]
    descs = (TensorDesc * len(ws))(*[w.desc() for w in ws])
    ys = (C.c_void_p * len(ws))(*[o.ptr for o in (outs * len(ws) if swiglu else outs)])
    ctx.check(ctx.lib.kf_rmsnorm_linear(ctx.h, len(ws), ys, descs, x.ptr, norm_w.ptr, eps, M, 2 if swiglu else 0), "kf_rmsnorm_linear")
    return outs[0] if swiglu else outs


def rmsnorm(ctx, x, w, rows, dim, eps=1e-6):
    out = ctx.empty(rows * dim * 2)
    ctx.check(ctx.lib.kf_rmsnorm(ctx.h, out.ptr, x.ptr, w.ptr, rows, dim, eps), "kf_rmsnorm")
    return out


def rope_table(ctx, max_seq, head_dim, theta):
    t = ctx.empty(max_seq * (head_dim // 2) * 8)
    ctx.check(ctx.lib.kf_rope_table(ctx.h, t.ptr, max_seq, head_dim, theta), "kf_rope_table")
    return t


def qknorm_rope_kvappend(ctx, q, k, v, qw, kw, kcache, vcache, table, pos, M, n_head, n_kv, hd, max_seq, eps=1e-6, seq_stride=0):
    ctx.check(ctx.lib.kf_qknorm_rope_kvappend(ctx.h, q.ptr, k.ptr, v.ptr, qw.ptr if qw is not None else None, kw.ptr if kw is not None else None,
                                              kcache.ptr, vcache.ptr, table.ptr, pos.ptr, M, n_head, n_kv, hd, max_seq, eps, seq_stride),
              "kf_qknorm_rope_kvappend")


def attn_decode(ctx, q, kcache, vcache, pos, M, n_head, n_kv, hd, max_seq, max_pos_hint, seq_stride=0):
    out = ctx.empty(M * n_head * hd * 2)
    ctx.check(ctx.lib.kf_attn_decode(ctx.h, out.ptr, q.ptr, kcache.ptr, vcache.ptr, pos.ptr, M, n_head, n_kv, hd, max_seq, max_pos_hint, seq_stride),
              "kf_attn_decode")
    return out


def attn_decode_gqa(ctx, q, kcache, vcache, pos, M, n_head, n_kv, hd, max_seq, max_pos_hint, seq_stride=0):
    """decode attention with the query heads of a kv group processed together on the tensor cores (batched decode)"""
    out = ctx.empty(M * n_head * hd * 2)
    ctx.check(ctx.lib.kf_attn_decode_gqa(ctx.h, out.ptr, q.ptr, kcache.ptr, vcache.ptr, pos.ptr, M, n_head, n_kv, hd, max_seq, max_pos_hint,
                                         seq_stride), "kf_attn_decode_gqa")
    return out


def attn_prefill(ctx, q, kcache, vcache, pos, M, n_head, n_kv, hd, max_seq):
    """causal attention of a panel of M consecutive tokens of one sequence (pos[0] = position of the first)"""
    out = ctx.empty(M * n_head * hd * 2)
    ctx.check(ctx.lib.kf_attn_prefill(ctx.h, out.ptr, q.ptr, kcache.ptr, vcache.ptr, pos.ptr, M, n_head, n_kv, hd, max_seq), "kf_attn_prefill")
    return out


def qkv_attention(ctx, q, k, v, qw, kw, kcache, vcache, table, pos, M, n_head, n_kv, hd, max_seq, max_pos_hint, eps=1e-6, seq_stride=0):
    """QK-norm + RoPE + KV append + split-K attention in one launch (decode: one sequence per token)"""
    out = ctx.empty(M * n_head * hd * 2)
    ctx.check(ctx.lib.kf_qkv_attention(ctx.h, out.ptr, q.ptr, k.ptr, v.ptr, qw.ptr if qw is not None else None, kw.ptr if kw is not None else None,
                                       kcache.ptr, vcache.ptr, table.ptr, pos.ptr, M, n_head, n_kv, hd, max_seq, eps, seq_stride, max_pos_hint),
              "kf_qkv_attention")
    return out


def swiglu(ctx, g, u, n):
    out = ctx.empty(n * 2)
    ctx.check(ctx.lib.kf_swiglu(ctx.h, out.ptr, g.ptr, u.ptr, n), "kf_swiglu")
    return out


def add(ctx, a, b, n):
    out = ctx.empty(n * 2)
    ctx.check(ctx.lib.kf_add(ctx.h, out.ptr, a.ptr, b.ptr, n), "kf_add")
    return out


def embed(ctx, w, tokens, M):
    out = ctx.empty(M * w.cols * 2)
    d = w.desc()
    ctx.check(ctx.lib.kf_embed(ctx.h, out.ptr, C.byref(d), tokens.ptr, M), "kf_embed")
    return out


def argmax(ctx, logits, M, vocab):
    out = ctx.empty(M * 4)
    ctx.check(ctx.lib.kf_argmax(ctx.h, out.ptr, logits.ptr, M, vocab), "kf_argmax")
    return out


def sample(ctx, logits, M, vocab, temperature, top_k, top_p, rng_state, selection=0):
    """rng_state: device buffer of M uint64 xorshift64* states (advanced in place); returns the device buffer of M int32 tokens"""
    out = ctx.empty(M * 4)
    ctx.check(ctx.lib.kf_sample(ctx.h, out.ptr, logits.ptr, M, vocab, float(temperature), int(top_k), float(top_p), rng_state.ptr, int(selection)),
              "kf_sample")
    return out


def safetensors_index(path):
    """header of a .safetensors file: list of {"name", "dtype", "shape", "nbytes"} in file order (host only)"""
    lib = L.load()
    out, err = C.c_void_p(), C.c_void_p()
    st = lib.kf_safetensors_index(str(path).encode(), C.byref(out), C.byref(err))
    msg = C.cast(err, C.c_char_p).value.decode() if err.value else ""
    if err.value:
        lib.kf_string_free(err)
    if st != L.KF_OK:
        raise L.KoifishError(st, "kf_safetensors_index", msg)
    text = C.cast(out, C.c_char_p).value.decode()
    lib.kf_string_free(out)
    return json.loads(text)


def safetensors_read_bf16(path, name, n):
    """one tensor of a .safetensors file as bf16 bit patterns (BF16 / F16 / F32 sources), converted as the loader converts it (host only)"""
    lib = L.load()
    out, err = np.zeros(n, dtype=np.uint16), C.c_void_p()
    st = lib.kf_safetensors_read_bf16(str(path).encode(), name.encode(), out.ctypes.data, n, C.byref(err))
    msg = C.cast(err, C.c_char_p).value.decode() if err.value else ""
    if err.value:
        lib.kf_string_free(err)
    if st != L.KF_OK:
        raise L.KoifishError(st, "kf_safetensors_read_bf16", msg)
    return out


class Model:
    """The Qwen3 runtime behind include/kf_model.h (reference: Fish::MakeInstance + Fish::Chat's per-token ForwardOnRLS)."""

    def __init__(self, ctx, config, tp_rank=0, tp_world=1):
        self.ctx = ctx
        self.lib = ctx.lib
        text = config if isinstance(config, str) else json.dumps(config)
        h, err = C.c_void_p(), C.c_void_p()
        st = self.lib.kf_model_create(ctx.h, text.encode(), tp_rank, tp_world, C.byref(h), C.byref(err))
        if st != L.KF_OK:
            msg = C.cast(err, C.c_char_p).value.decode() if err.value else ""
            if err.value:
                self.lib.kf_string_free(err)
            raise KoifishError(st, "kf_model_create", msg)
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.kf_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st, where):
        if st != L.KF_OK:
            raise KoifishError(st, where, self.lib.kf_model_error(self.h).decode() + " | " + self.lib.kf_last_error(self.ctx.h).decode())

    @property
    def info(self):
        i = ModelInfo()
        self._check(self.lib.kf_model_info_get(self.h, C.byref(i)), "kf_model_info_get")
        return i

    def init_random(self):
        self._check(self.lib.kf_model_init_random(self.h), "kf_model_init_random")

    def set_tensor(self, name, w_bf16):
        w = np.ascontiguousarray(w_bf16, dtype=np.uint16)
        rows, cols = (1, w.size) if w.ndim == 1 else w.shape
        self._check(self.lib.kf_model_set_tensor(self.h, name.encode(), w.ctypes.data, rows, cols), "kf_model_set_tensor(%s)" % name)

    def set_tensor_awq(self, name, qweight, qzeros, scales_f16):
        """one linear in the vendor AWQ layout, full shape: qweight int32 [in][out / 8], qzeros int32 [in / 128][out / 8], scales fp16 bits [in / 128][out]"""
        qw = np.ascontiguousarray(qweight).view(np.uint32)
        qz = np.ascontiguousarray(qzeros).view(np.uint32)
        sc = np.ascontiguousarray(scales_f16).view(np.uint16)
        self._check(self.lib.kf_model_set_tensor_awq(self.h, name.encode(), qw.ctypes.data, qz.ctypes.data, sc.ctypes.data, qw.shape[0], sc.shape[1]),
                    "kf_model_set_tensor_awq(%s)" % name)

    def tensor_names(self):
        return [self.lib.kf_model_tensor_name(self.h, i).decode() for i in range(self.lib.kf_model_tensor_count(self.h))]

    def tensor_desc(self, name):
        d = TensorDesc()
        self._check(self.lib.kf_model_tensor_desc(self.h, name.encode(), C.byref(d)), "kf_model_tensor_desc(%s)" % name)
        return d

    def dequant_tensor(self, name):
        """GetDataX of a resident tensor -> numpy bf16 bits [rows, cols]"""
        d = self.tensor_desc(name)
        out = self.ctx.empty(d.rows * d.cols * 2)
        self.ctx.check(self.lib.kf_dequant(self.ctx.h, C.byref(d), out.ptr), "kf_dequant")
        return out.numpy(np.uint16, (d.rows, d.cols))

    def forward(self, tokens, pos, seq_mode=0, want_logits=True, want_next=False):
        tokens = np.ascontiguousarray(tokens, dtype=np.int32).reshape(-1)
        pos = np.ascontiguousarray(pos, dtype=np.int32).reshape(-1)
        M = tokens.size
        vocab = self.info.vocab
        R = 1 if seq_mode == 2 else M  # seq_mode 2: prefill panel, outputs of the last token only
        logits = np.empty((R, vocab), dtype=np.uint16) if want_logits else None
        nxt = np.empty(R, dtype=np.int32) if want_next else None
        self._check(self.lib.kf_model_forward(self.h, tokens.ctypes.data, pos.ctypes.data, M, seq_mode,
                                              logits.ctypes.data if want_logits else None, nxt.ctypes.data if want_next else None), "kf_model_forward")
        return logits, nxt

    def prefill(self, tokens, pos0=0, want_logits=False):
        """run a prompt through in panels of info.max_tokens; returns (logits of the last token or None, greedy next token).  Leaves
        the model ready for decode_loop(n, 1)."""
        tokens = np.ascontiguousarray(tokens, dtype=np.int32).reshape(-1)
        P = self.info.max_tokens
        logits = nxt = None
        for s in range(0, tokens.size, P):
            part = tokens[s:s + P]
            last = s + P >= tokens.size
            logits, nxt = self.forward(part, np.arange(pos0 + s, pos0 + s + part.size), seq_mode=2, want_logits=want_logits and last,
                                       want_next=last)
        return logits, (int(nxt[0]) if nxt is not None else None)

    def decode_loop(self, n_steps, M=1):
        self._check(self.lib.kf_model_decode_loop(self.h, n_steps, M), "kf_model_decode_loop")

    def read_state(self, M=1):
        t, p = np.empty(M, dtype=np.int32), np.empty(M, dtype=np.int32)
        self._check(self.lib.kf_model_read_state(self.h, t.ctypes.data, p.ctypes.data, M), "kf_model_read_state")
        return t, p

    def generate(self, prompt_ids, max_new_tokens, eos_id=-1, pos0=0):
        """Fish::Chat's generation loop -> (generated ids, stop reason: 1 eos / 2 max_new_tokens / 3 context window full)"""
        p = np.ascontiguousarray(prompt_ids, dtype=np.int32).reshape(-1)
        out = np.zeros(max(1, max_new_tokens), dtype=np.int32)
        n, why = C.c_int(0), C.c_int(0)
        self._check(self.lib.kf_model_generate(self.h, p.ctypes.data, p.size, int(pos0), int(max_new_tokens), int(eos_id), out.ctypes.data, C.byref(n),
                                               C.byref(why)), "kf_model_generate")
        return [int(x) for x in out[:n.value]], why.value

    def save(self, path):
        self._check(self.lib.kf_model_save(self.h, str(path).encode()), "kf_model_save")

    def load(self, path):
        self._check(self.lib.kf_model_load(self.h, str(path).encode()), "kf_model_load")

    def save_kun(self, path):
        """the reference's own container (fish.kun)"""
        self._check(self.lib.kf_model_save_kun(self.h, str(path).encode()), "kf_model_save_kun")

    def load_kun(self, path):
        a, b = C.c_int(0), C.c_int(0)
        self._check(self.lib.kf_model_load_kun(self.h, str(path).encode(), C.byref(a), C.byref(b)), "kf_model_load_kun")
        return a.value, b.value

    def load_safetensors(self, path_or_dir):
        """HF model.safetensors (or a directory of shards) -> (tensors loaded, tensors skipped)"""
        a, b = C.c_int(0), C.c_int(0)
        self._check(self.lib.kf_model_load_safetensors(self.h, str(path_or_dir).encode(), C.byref(a), C.byref(b)), "kf_model_load_safetensors")
        return a.value, b.value

    def set_sampler(self, temperature, top_k=50, top_p=0.95, seed=42, selection=0):
        self._check(self.lib.kf_model_set_sampler(self.h, float(temperature), int(top_k), float(top_p), int(seed), int(selection)), "kf_model_set_sampler")

    def set_graphs(self, enable):
        self._check(self.lib.kf_model_set_graphs(self.h, int(bool(enable))), "kf_model_set_graphs")

    def kcache(self, layer, npos, kv_dim):
        p = self.lib.kf_model_kcache(self.h, layer)
        out = np.empty(npos * kv_dim, dtype=np.uint16)
        self.ctx.check(self.lib.kf_d2h(self.ctx.h, out.ctypes.data, p, out.nbytes), "kf_d2h")
        self.ctx.sync()
        return out.reshape(npos, kv_dim)

    def vcache(self, layer, npos, kv_dim):
        p = self.lib.kf_model_vcache(self.h, layer)
        out = np.empty(npos * kv_dim, dtype=np.uint16)
        self.ctx.check(self.lib.kf_d2h(self.ctx.h, out.ctypes.data, p, out.nbytes), "kf_d2h")
        self.ctx.sync()
        return out.reshape(npos, kv_dim)


def qwen3_config(n_layer, n_embd, n_ff, n_head, n_kv_head, head_dim=128, vocab=151936, quantizer=None, tie=False, max_seq_len=1024,
                 max_batch=1, seed=42, rope_theta=None, sigma=None, norm_sigma=None, max_prefill=None):
    """a Koifish-style JSON config (reference cases/qwen3/qwen3_596M_q4.json layout) as a dict"""
    cfg = {
        "version": "0.1.0",
        "model": {"arch": "QWEN3", "parameter": {
            "Layer": n_layer,
            "transformer": {"Ctx": max_seq_len, "Embed": n_embd, "Ffn": n_ff, "Head": n_head, "KVHead": n_kv_head, "head_dim": head_dim},
            "tie_word_embeddings": bool(tie), "max_pos_embeddings": 32768, "vocab_size": vocab}},
        "gpt": {"max_seq_len": max_seq_len, "max_batch": max_batch},
        "seed": seed,
    }
    if max_prefill is not None:
        cfg["gpt"]["max_prefill"] = max_prefill  # tokens per prefill panel (activation buffers); default 64
    if rope_theta is not None:
        cfg["model"]["parameter"]["rope_theta"] = rope_theta
    if quantizer:
        cfg["quantizer"] = quantizer
    init = {}
    if sigma is not None:
        init["sigma"] = sigma
    if norm_sigma is not None:
        init["norm_sigma"] = norm_sigma
    if init:
        cfg["init"] = init
    return cfg


QWEN3_DIMS = {  # SURVEY.md section 8 table (HF Qwen3 configs)
    "0.6B": dict(n_layer=28, n_embd=1024, n_ff=3072, n_head=16, n_kv_head=8, tie=True),
    "8B": dict(n_layer=36, n_embd=4096, n_ff=12288, n_head=32, n_kv_head=8, tie=False),
    "32B": dict(n_layer=64, n_embd=5120, n_ff=25600, n_head=64, n_kv_head=8, tie=False),
}


# ---------------------------------------------------------------------------------------------- text side of the chat loop (host only)
def _take_string(lib, p):
    s = C.cast(p, C.c_char_p).value
    lib.kf_string_free(p)
    return s.decode("utf-8", errors="surrogateescape")


class Tokenizer:
    """include/kf_tokenizer.h: HF tokenizer.json of the Qwen family (reference src/TokenSet/HF_Tokenizer.cpp).  Host only."""

    def __init__(self, path=None, json_text=None, config_text=None):
        self.lib = L.load()
        h, err = C.c_void_p(), C.c_void_p()
        if path is not None:
            st = self.lib.kf_tokenizer_load(str(path).encode(), C.byref(h), C.byref(err))
        else:
            st = self.lib.kf_tokenizer_from_json(json_text.encode(), config_text.encode() if config_text else None, C.byref(h), C.byref(err))
        if st != L.KF_OK:
            raise KoifishError(st, "kf_tokenizer_load", _take_string(self.lib, err) if err.value else "")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.kf_tokenizer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def encode(self, text):
        raw = text if isinstance(text, bytes) else text.encode("utf-8")
        n = C.c_size_t(0)
        st = self.lib.kf_tokenizer_encode(self.h, raw, len(raw), None, 0, C.byref(n))
        if st != L.KF_OK:
            raise KoifishError(st, "kf_tokenizer_encode", "text is not valid UTF-8")
        ids = np.zeros(max(1, n.value), dtype=np.int32)
        st = self.lib.kf_tokenizer_encode(self.h, raw, len(raw), ids.ctypes.data, ids.size, C.byref(n))
        if st != L.KF_OK:
            raise KoifishError(st, "kf_tokenizer_encode")
        return [int(x) for x in ids[:n.value]]

    def decode(self, ids, skip_special_tokens=False):
        a = np.ascontiguousarray(ids, dtype=np.int32).reshape(-1)
        out = C.c_void_p()
        st = self.lib.kf_tokenizer_decode(self.h, a.ctypes.data if a.size else None, a.size, int(skip_special_tokens), C.byref(out))
        if st != L.KF_OK:
            raise KoifishError(st, "kf_tokenizer_decode")
        return _take_string(self.lib, out)

    def stream(self, ids, skip_special_tokens=False):
        """generator of text pieces, one per id (possibly empty), then the flushed rest: whole characters only"""
        h = C.c_void_p()
        self.lib.kf_decode_stream_create(C.byref(h))
        try:
            for i in ids:
                out = C.c_void_p()
                st = self.lib.kf_decode_stream_push(self.h, h, int(i), int(skip_special_tokens), C.byref(out))
                if st != L.KF_OK:
                    raise KoifishError(st, "kf_decode_stream_push")
                yield _take_string(self.lib, out)
            out = C.c_void_p()
            self.lib.kf_decode_stream_flush(h, C.byref(out))
            yield _take_string(self.lib, out)
        finally:
            self.lib.kf_decode_stream_destroy(h)

    def pre_tokenize(self, text):
        raw = text.encode("utf-8")
        out = C.c_void_p()
        st = self.lib.kf_tokenizer_pre_tokenize(self.h, raw, len(raw), C.byref(out))
        if st != L.KF_OK:
            raise KoifishError(st, "kf_tokenizer_pre_tokenize")
        return json.loads(_take_string(self.lib, out))

    def token_to_id(self, token):
        return self.lib.kf_tokenizer_token_to_id(self.h, token.encode("utf-8"))

    def id_to_token(self, i):
        out = C.c_void_p()
        st = self.lib.kf_tokenizer_id_to_token(self.h, int(i), C.byref(out))
        if st != L.KF_OK:
            raise KoifishError(st, "kf_tokenizer_id_to_token")
        return _take_string(self.lib, out)

    vocab_size = property(lambda self: self.lib.kf_tokenizer_vocab_size(self.h))
    eos_id = property(lambda self: self.lib.kf_tokenizer_eos_id(self.h))
    bos_id = property(lambda self: self.lib.kf_tokenizer_bos_id(self.h))
    pad_id = property(lambda self: self.lib.kf_tokenizer_pad_id(self.h))

    def is_special(self, i):
        return bool(self.lib.kf_tokenizer_is_special(self.h, int(i)))


def nfc(text):
    lib = L.load()
    raw = text.encode("utf-8")
    out = C.c_void_p()
    st = lib.kf_text_nfc(raw, len(raw), C.byref(out))
    if st != L.KF_OK:
        raise KoifishError(st, "kf_text_nfc")
    return _take_string(lib, out)


def chatml_prompt(user, system=None, enable_thinking=False):
    """CHAT_SAMPLER::InitPrefillTemplate (reference src/Utils/CLI_params.cpp:1990-2008)"""
    lib = L.load()
    out = C.c_void_p()
    st = lib.kf_chatml_prompt(system.encode("utf-8") if system else None, user.encode("utf-8"), int(enable_thinking), C.byref(out))
    if st != L.KF_OK:
        raise KoifishError(st, "kf_chatml_prompt")
    return _take_string(lib, out)


def chatml_render(lines, enable_thinking=False):
    """CHAT_SAMPLER::toChatML (reference src/Utils/CLI_params.cpp:2010-2031); lines: [(role, content), ...]"""
    lib = L.load()
    n = len(lines)
    roles = (C.c_char_p * max(1, n))(*[r.encode("utf-8") for r, _ in lines])
    texts = (C.c_char_p * max(1, n))(*[c.encode("utf-8") for _, c in lines])
    out = C.c_void_p()
    st = lib.kf_chatml_render(roles, texts, n, int(enable_thinking), C.byref(out))
    if st != L.KF_OK:
        raise KoifishError(st, "kf_chatml_render")
    return _take_string(lib, out)


def chat_once(model, tokenizer, user, system=None, max_new_tokens=256, enable_thinking=False, pos0=0):
    """one turn of Fish::Chat: ChatML prompt -> token ids -> kf_model_generate (stops at the tokenizer's eos) -> text"""
    ids = tokenizer.encode(chatml_prompt(user, system, enable_thinking))
    out, why = model.generate(ids, max_new_tokens, tokenizer.eos_id, pos0)
    return tokenizer.decode(out, skip_special_tokens=True), out, why


def _host_json_call(fn_name, path):
    lib = L.load()
    out, err = C.c_void_p(), C.c_void_p()
    st = getattr(lib, fn_name)(str(path).encode(), C.byref(out), C.byref(err))
    if st != L.KF_OK:
        raise KoifishError(st, fn_name, _take_string(lib, err) if err.value else "")
    return _take_string(lib, out)


def kun_index(path):
    """header of a fish.kun file (host only): [{"name", "dtype", "shape", "szData", "szGama", "offset"}, ...]"""
    return json.loads(_host_json_call("kf_kun_index", path))


def kun_config(path):
    """the msgpack "__koifish__config__" entry of a fish.kun file as a dict (None when absent)"""
    text = _host_json_call("kf_kun_config", path)
    return json.loads(text) if text else None


def kun_write(path, config, tensors):
    """write a fish.kun from host arrays (host only).  tensors: [(name, K_FLOATS dtype name, shape tuple, szData, szGama, bytes-like blob)]"""
    lib = L.load()
    n = len(tensors)
    names = (C.c_char_p * max(1, n))(*[t[0].encode() for t in tensors])
    dtypes = (C.c_char_p * max(1, n))(*[t[1].encode() for t in tensors])
    shapes = (C.c_int64 * max(2, 2 * n))(*[d for t in tensors for d in (t[2][0], t[2][1] if len(t[2]) > 1 else 0)])
    szd = (C.c_uint64 * max(1, n))(*[t[3] for t in tensors])
    szg = (C.c_uint64 * max(1, n))(*[t[4] for t in tensors])
    keep = [np.frombuffer(bytes(t[5]), dtype=np.uint8) for t in tensors]
    blobs = (C.c_void_p * max(1, n))(*[k.ctypes.data for k in keep])
    err = C.c_void_p()
    cfg = json.dumps(config).encode() if config is not None else None
    st = lib.kf_kun_write(str(path).encode(), cfg, n, names, dtypes, shapes, szd, szg, blobs, C.byref(err))
    if st != L.KF_OK:
        raise KoifishError(st, "kf_kun_write", _take_string(lib, err) if err.value else "")


def from_pretrained(ctx, model_dir, quantizer=None, max_seq_len=1024, max_batch=1, tp_rank=0, tp_world=1):
    """the reference's `--hf <dir>` path (MODEL_CARD::InitHugFace, src/Utils/CLI_params.cpp:2224-2300 + Fish::LoadFolderOfST): an HF checkpoint
    directory (config.json, *.safetensors, tokenizer.json) -> (Model with every tensor loaded, Tokenizer).  `quantizer`: a Koifish "quantizer"
    block to quantise a bf16 checkpoint at load; a vendor-quantised (AWQ) checkpoint brings its own through config.json's quantization_config."""
    import os
    with open(os.path.join(str(model_dir), "config.json")) as f:
        hf = json.load(f)
    cfg = {"hf_config": hf, "gpt": {"max_seq_len": int(max_seq_len), "max_batch": int(max_batch)}}
    if quantizer:
        cfg["quantizer"] = quantizer
    model = Model(ctx, cfg, tp_rank, tp_world)
    loaded, _ = model.load_safetensors(model_dir)
    missing = len(model.tensor_names()) - loaded
    if missing:
        raise KoifishError(L.KF_ERR_BAD_ARG, "from_pretrained", "%d of the model's tensors are not in '%s'" % (missing, model_dir))
    return model, Tokenizer(model_dir)
