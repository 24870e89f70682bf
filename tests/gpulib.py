"""loads the product package (directory name has a hyphen, so importlib is used)"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pkg = importlib.import_module("x265-mod-by-patman_b200")

_ctx = {}


def context(depth, device=0):
    key = (depth, device)
    if key not in _ctx:
        _ctx[key] = pkg.Context(depth, device)
    return _ctx[key]
