"""CUDA path, through the per-call HOST entries of the C ABI, against the committed golden vectors
(outputs of the reference's own C primitives) and against the oracle on the same inputs."""
import os

import numpy as np
import pytest

from golden_cases import input_checksum, run_cases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("depth", [8, 10, 12])
def test_cuda_host_entries_match_reference_golden(depth):
    from gpulib import context
    ctx = context(depth)
    gold = np.load(os.path.join(GOLDEN, "x265_ref_%d.npz" % depth))
    assert int(gold["__input_crc__"][0]) == int(input_checksum(depth)[0])
    before = ctx.launch_count()
    got = run_cases(ctx.host, depth)
    assert ctx.launch_count() - before > 700           # every case went through a CUDA kernel
    bad = [k for k, v in got.items() if not np.array_equal(np.asarray(v).astype(np.int64), gold[k].astype(np.int64))]
    assert not bad, bad[:20]
