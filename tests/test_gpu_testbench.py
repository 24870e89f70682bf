"""The reference's own TestBench harness classes (PixelHarness, MBDstHarness, IPFilterHarness, IntraPredHarness --
reference source/test/*.cpp, unmodified, prebuilt into oracle/_ref/testbench_b200_<depth> by
`make -C oracle harness`) run against the B200 table filled by setupB200Primitives.
This is the reference's own opt-vs-C parity check (testbench.cpp:229-277) with `opt` = CUDA."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("depth", [8, 10, 12])
def test_reference_testbench_harness_passes(depth):
    exe = os.path.join(ROOT, "oracle", "_ref", "testbench_b200_%d" % depth)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref harness not prebuilt (needs /root/reference at build time)")
    r = subprocess.run([exe, "0x265"], capture_output=True, text=True, timeout=1500)
    tail = (r.stdout[-3000:] + r.stderr[-2000:])
    assert r.returncode == 0, tail
    assert "ALL PASSED" in r.stdout, tail
    assert "0 missing, 0 still C, 0 extra" in r.stdout, tail
