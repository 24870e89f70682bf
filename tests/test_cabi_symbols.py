"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol the header
declares.  No compute call is made here."""
import ctypes
import os

import pytest

from gpulib import pkg


def test_library_built_and_exports_all_declared_symbols():
    if not os.path.exists(pkg.LIB_PATH):
        pkg.build_library(glue=False)
    lib = pkg.load_library()
    syms = pkg.declared_symbols()
    assert len(syms) >= 35
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_glue_libraries_export_their_header():
    """the per-depth table fillers export what include/x265b200_glue.h declares (built only where the reference headers exist)"""
    syms = pkg.declared_symbols(pkg.GLUE_HEADER)
    assert {"x265b200_setup_primitives", "x265b200_glue_context", "x265b200_glue_depth", "x265b200_glue_status"} <= set(syms)
    built = [d for d in (8, 10, 12) if os.path.exists(pkg.glue_path(d))]
    if not built:
        pytest.skip("glue libraries not built (no reference headers)")
    pkg.load_library()
    for d in built:
        lib = ctypes.CDLL(pkg.glue_path(d))
        assert not [s for s in syms if not hasattr(lib, s)], d
        assert lib.x265b200_glue_depth() == d


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        pkg.Context(10)


def test_product_does_not_reference_oracle():
    """the product tree must never import / link the oracle (parity claims depend on it)"""
    root = os.path.dirname(pkg.LIB_PATH)
    pkgdir = os.path.dirname(root)
    for dp, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert "x265_oracle" not in text and "liboracle" not in text and "cpulibs" not in text, os.path.join(dp, f)
    out = os.popen("ldd %s" % pkg.LIB_PATH).read()
    assert "oracle" not in out and "x265ref" not in out


def test_docs_quote_the_current_entry_point_count():
    """README / DESIGN / INTEGRATION state how many C entry points the header declares; keep them honest"""
    import re
    n = len(pkg.declared_symbols())
    repo = os.path.dirname(os.path.dirname(os.path.dirname(pkg.LIB_PATH)))
    for doc in ("README.md", "DESIGN.md", "INTEGRATION.md"):
        text = open(os.path.join(repo, doc)).read()
        counts = [int(m) for m in re.findall(r"(\d+) entry points", text)]
        assert counts and all(c == n for c in counts), (doc, counts, n)
