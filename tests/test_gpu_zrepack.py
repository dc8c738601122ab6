"""GPU test of gpt.awq_repack = 1 (vendor AWQ tensors re-laid-out at load into the library's own 4-bit storage so the tuned decode / tensor-core
kernels run on them).  The re-layout itself is pinned bit-exactly on the CPU (tests/test_cabi_host.py::test_awq_repack_into_packedq_storage_...);
here the model built from the repacked tensors is compared with the vendor-layout model and with the bf16 model of the dequantised weights."""
import numpy as np
import pytest

import koifish_b200 as kf
import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = kf.Context(0)
    yield c
    c.close()


def test_awq_checkpoint_repacked_runs_on_the_q4_kernels(ctx, tmp_path):
    from test_gpu_model import AWQ_VENDOR_BLOCK, _awq_models, logits_close, prompt
    a, b, d, want = _awq_models(ctx, tmp_path)  # a: vendor layout, b: bf16 model of the weights CU_Q42X_awq reads
    a.load_safetensors(d)
    hf = {"hf_config": {"hidden_size": 256, "intermediate_size": 512, "num_hidden_layers": 2, "num_attention_heads": 4, "num_key_value_heads": 2,
                        "head_dim": 64, "vocab_size": 1024, "rope_theta": 1e6, "tie_word_embeddings": False, "quantization_config": AWQ_VENDOR_BLOCK},
          "gpt": {"max_seq_len": 64, "max_batch": 1, "awq_repack": 1}}
    r = kf.Model(ctx, hf)
    loaded, skipped = r.load_safetensors(d)
    assert (loaded, skipped) == (len(b.tensor_names()), 0)
    for name, w in want.items():
        t = r.tensor_desc(name)
        assert (t.type, t.rows, t.cols, t.group, t.qbias) == (kf.KF_T_Q4, w.shape[0], w.shape[1], 128, 0), name
        got, ref = ol.bf16_to_f32(r.dequant_tensor(name)), ol.bf16_to_f32(w)
        # step = bf16(scale), zero = bf16(zero_point * scale): each off by 2^-9 relative at most, times codes <= 15; a step is about max|w| / 7.5
        assert np.abs(got - ref).max() <= 3e-2 * np.abs(ref).max(), name
        assert np.abs(got - ref).mean() <= 2e-3 * np.abs(ref).max(), name
    toks = prompt(12, 1024)
    for pos, tok in enumerate(toks[:6]):
        lr, nr = r.forward([tok], [pos], want_next=True)
        lb, nb = b.forward([tok], [pos], want_next=True)
        la, _ = a.forward([tok], [pos])
        for other in (lb, la):
            err, _, _ = logits_close(lr[0], other[0])
            assert err <= 3e-2, (pos, err)
        w = ol.bf16_to_f32(lb[0])
        top2 = np.sort(w)[-2:]
        if top2[1] - top2[0] > 2 * 3e-2 * np.abs(w).max():
            assert int(nr[0]) == int(nb[0])
    lr, _ = r.forward(toks, list(range(12)))  # a panel: the tcgen05 dequant-GEMM on the repacked words
    lb, _ = b.forward(toks, list(range(12)))
    for m in range(12):
        err, _, _ = logits_close(lr[m], lb[m])
        assert err <= 3e-2, (m, err)
