"""oracle/x265_oracle.c against the committed golden vectors of the motion-search / intra family (outputs of the reference's
own MotionEstimate::motionEstimate and intrapred.cpp slots, tests/golden/make_golden_me.py).  Runs anywhere gcc exists."""
import os

import numpy as np
import pytest

from cpulibs import Oracle
from golden_cases_me import input_checksum, oracle_outputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("depth", [8, 10, 12])
def test_oracle_matches_golden_me(depth):
    gold = np.load(os.path.join(GOLDEN, "x265_ref_me_%d.npz" % depth))
    assert int(gold["__input_crc__"][0]) == int(input_checksum(depth)[0]), "input generator drifted"
    got = oracle_outputs(Oracle(depth), depth, gold)
    for k, v in got.items():
        assert np.array_equal(np.asarray(v).astype(np.int64), gold[k].astype(np.int64)), k
    assert len(set(map(tuple, gold["motion_estimate"][:, :2].tolist()))) > 20      # a spread of vectors, not a constant
