"""Golden vectors for the vendor AWQ layout from the reference's OWN Python implementation: unpack_awq / reverse_awq_order / Dequant_1 of
/root/reference/src/Python/test_awq.py:33-134.  That module imports `awq` (not installed), so the three functions and the two order tables
are taken from its source with `ast` and executed as they are (nothing is copied into this repository); pandas printing is silenced.
Writes tests/golden/awq_ref_py.npz: seeded qweight / qzeros / scales -> the int codes after the order reversal and the bf16 weights, as the
reference's Python computes them (fp16 multiply, then bf16 -- its CUDA kernel CU_Q42X_awq multiplies in fp32, so values may differ by one
bf16 ulp where the fp16 product is inexact; the codes may not differ at all).
Run in the build container:  python tests/golden/make_awq_golden.py"""
import ast
import os

import numpy as np
import torch

SRC = "/root/reference/src/Python/test_awq.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "awq_ref_py.npz")


def load_reference_functions():
    tree = ast.parse(open(SRC).read())
    keep = [n for n in tree.body if (isinstance(n, ast.FunctionDef) and n.name in ("unpack_awq", "reverse_awq_order", "Dequant_1")) or
            (isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") in ("AWQ_ORDER", "AWQ_REVERSE_ORDER"))]
    ns = {"torch": torch, "np": np, "save_dequantized_to_csv": lambda *a, **k: None, "print": lambda *a, **k: None}
    exec(compile(ast.Module(body=keep, type_ignores=[]), SRC, "exec"), ns)
    return ns


def main():
    ref = load_reference_functions()
    rng = np.random.default_rng(20261018)
    out = {}
    for tag, (IC, OC), unit in (("a", (256, 64), False), ("b", (128, 264), False), ("c", (256, 128), True)):
        qw = rng.integers(-2 ** 31, 2 ** 31, size=(IC, OC // 8), dtype=np.int64).astype(np.int32)
        qz = rng.integers(-2 ** 31, 2 ** 31, size=(IC // 128, OC // 8), dtype=np.int64).astype(np.int32)
        sc = np.ones((IC // 128, OC), dtype=np.float16) if unit else (rng.random((IC // 128, OC)) * 0.02 + 1e-3).astype(np.float16)
        tq, tz, ts = torch.from_numpy(qw), torch.from_numpy(qz), torch.from_numpy(sc)
        iw, iz = ref["unpack_awq"](tq, tz, 4)
        iw, iz = torch.bitwise_and(iw, 15), torch.bitwise_and(iz, 15)
        iw, iz = ref["reverse_awq_order"](iw, iz, 4)
        deq = ref["Dequant_1"]("x." + tag, tq, ts, tz)  # bf16 [in][out]
        out.update({tag + "_qweight": qw, tag + "_qzeros": qz, tag + "_scales": sc.view(np.uint16), tag + "_codes": iw.numpy().astype(np.uint8),
                    tag + "_zeros": iz.numpy().astype(np.uint8), tag + "_deq_bf16": deq.view(torch.int16).numpy().view(np.uint16)})
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
