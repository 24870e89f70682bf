"""Generate tests/golden/x265_ref_me_<depth>.npz: outputs of the REFERENCE's own MotionEstimate::motionEstimate (all six
search methods), intra predictors and lookahead intra estimate on the fixed cases of tests/golden_cases_me.py, plus the
lambda-scaled cost tables BitCost::setQP builds for the QPs used.  Run in the authoring container only (needs oracle/_ref):
    python tests/golden/make_golden_me.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from cpulibs import Reference                      # noqa: E402
from golden_cases_me import reference_outputs      # noqa: E402

for depth in (8, 10, 12):
    out = reference_outputs(Reference(depth), depth)
    path = os.path.join(HERE, "x265_ref_me_%d.npz" % depth)
    np.savez_compressed(path, **out)
    print(path, len(out), "arrays", os.path.getsize(path), "bytes")
