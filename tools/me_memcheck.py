"""Tiny invocation of every motion-search entry (exhaustive, pattern walks, full-resolution chain, lowres chain) on a small
picture, windows clipped exactly at the allocation's edge: the target of
    compute-sanitizer --tool memcheck python tools/me_memcheck.py
so that any read outside the planes the caller owns shows up.  The planes are allocated with no slack beyond the padded
picture."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from frames import Geometry, make_plane                        # noqa: E402
from test_oracle_vs_ref import lowres_planes, mv_cost_table    # noqa: E402
pkg = importlib.import_module("x265-mod-by-patman_b200")


def main():
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    ctx = pkg.Context(depth, 0)
    geo = Geometry(128, 64)
    vt = np.uint8 if depth == 8 else np.int16
    cw, ch = geo.coded()
    F = torch.from_numpy(make_plane(geo, depth, 1, "natural").view(vt)).cuda()
    R = torch.from_numpy(make_plane(geo, depth, 2, "natural").view(vt)).cuda()
    P = torch.from_numpy(lowres_planes(geo, depth, 3).view(vt)).cuda()
    RAD = 1024
    dtab = torch.from_numpy(mv_cost_table(8.0, RAD).view(np.int16)).cuda()
    rng = np.random.default_rng(5)
    for (w, h) in ((8, 8), (16, 16), (64, 64), (12, 16), (32, 8)):
        n = 24
        x = rng.integers(0, cw - w + 1, n); y = rng.integers(0, ch - h + 1, n)
        x[:4] = 0; y[:4] = 0; x[4:8] = cw - w; y[4:8] = ch - h           # corners: windows end at the padding's edge
        off = torch.from_numpy((geo.origin + y * geo.stride + x).astype(np.int32)).cuda()
        m = 40
        # full-pel window such that block + filter taps stay inside the plane
        # (vertically every candidate is range-checked, so 6 rows cover the filter taps; horizontally the walks may leave
        # the window by up to 3 samples before sub-pel refinement, as in the reference: 12 columns)
        win = np.stack([-np.minimum(m, x + geo.margin_x - 12), -np.minimum(m, y + geo.margin_y - 6),
                        np.minimum(m, cw + geo.margin_x - 12 - w - x), np.minimum(m, ch + geo.margin_y - 6 - h - y)], 1).astype(np.int32)
        dwin = torch.from_numpy(win.copy()).cuda()
        qmvp = torch.from_numpy(rng.integers(-60, 61, (n, 2)).astype(np.int32)).cuda()
        mvc = torch.from_numpy(rng.integers(-200, 201, (n, 3, 2)).astype(np.int32)).cuda()
        oq = torch.zeros((n, 2), dtype=torch.int32, device="cuda"); oc = torch.zeros(n, dtype=torch.int32, device="cuda")
        for method in (pkg.ME_FULL, pkg.ME_HEX, pkg.ME_DIA, pkg.ME_STAR, pkg.ME_UMH):
            for subme in (0, 2, 7):
                ctx.motion_estimate_batch(method, w, h, m if method != pkg.ME_STAR else 64, subme, F, geo.stride, R, geo.stride, off, off, dwin, qmvp,
                                          3, mvc, dtab.data_ptr() + 2 * RAD, oq, oc)
        for hint in (0, 7, 40):
            bmv = torch.zeros((n, 2), dtype=torch.int32, device="cuda"); bc = torch.full((n,), 0x7fffffff, dtype=torch.int32, device="cuda")
            ctx.me_full_batch(w, h, hint, F, geo.stride, R, geo.stride, off, off, dwin, qmvp, dtab.data_ptr() + 2 * RAD, bmv, bc)
        if (w, h) == (8, 8):
            for method in (pkg.ME_HEX, pkg.ME_STAR, pkg.ME_FULL):
                ctx.lowres_motion_estimate_batch(method, 8, 8, 16, 1, F, geo.stride, P, geo.stride, geo.plane_elems, off, off, dwin, qmvp,
                                                 dtab.data_ptr() + 2 * RAD, oq, oc)
        torch.cuda.synchronize()
    ctx.check()
    print("me_memcheck: all entries ran")


if __name__ == "__main__":
    main()
