"""Generate tests/golden/x265_ref_<depth>.npz from the REFERENCE's own C primitives.

Run in the authoring container only (needs oracle/_ref, built from /root/reference by
`make -C oracle ref`):   python tests/golden/make_golden.py
The reference ships no golden vectors (all its checks are differential), so these
files are outputs of the reference itself on the fixed-seed inputs of tests/golden_cases.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from cpulibs import Reference          # noqa: E402
from golden_cases import input_checksum, run_cases   # noqa: E402

for depth in (8, 10, 12):
    out = run_cases(Reference(depth), depth)
    out["__input_crc__"] = input_checksum(depth)
    small = {}
    for k, v in out.items():
        v = np.asarray(v)
        if v.dtype == np.int64 and v.size and abs(v).max() < 2 ** 31:
            v = v.astype(np.int32)
        small[k] = v
    path = os.path.join(HERE, "x265_ref_%d.npz" % depth)
    np.savez_compressed(path, **small)
    print(path, len(small), "arrays", os.path.getsize(path), "bytes")
