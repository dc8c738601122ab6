"""Generate tests/golden/packq_ref.npz from the REFERENCE's own PackedQ.hpp macros (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the build container only:  python tests/golden/make_golden.py
The fixture pins the 128-bit word layout for the oracle and for the CUDA unpackers on the GPU box,
where /root/reference does not exist."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_lib as ol  # noqa: E402

out = {}
rng = np.random.default_rng(20261017)
for bits in (4, 2, 1):
    per = 128 // bits
    n = per * 64
    codes = rng.integers(0, 1 << bits, size=n, dtype=np.int32)
    out[f"codes{bits}"] = codes
    out[f"bytes{bits}"] = ol.ref_pack(codes, bits)
    # one-hot walk: word j has only code j set to the max value
    oh = np.zeros(per * per, dtype=np.int32)
    for j in range(per):
        oh[j * per + j] = (1 << bits) - 1
    out[f"onehot_codes{bits}"] = oh
    out[f"onehot_bytes{bits}"] = ol.ref_pack(oh, bits)
# the probe quoted in SURVEY.md 8c: codes (3+7i) mod 16
kat = np.array([(3 + 7 * i) % 16 for i in range(32)], dtype=np.int32)
out["kat4_codes"] = kat
out["kat4_bytes"] = ol.ref_pack(kat, 4)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "packq_ref.npz"), **out)
print("wrote packq_ref.npz", {k: v.shape for k, v in out.items()})
