#!/usr/bin/env python
"""lanes-per-CU sweep for x265b200_cu_satd_batch (lab tool; X265B200_CU_LANES_LAB is read by the launcher)"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from frames import Geometry, cu_descriptors, make_plane, tile_blocks
pkg = importlib.import_module("x265-mod-by-patman_b200")
D = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ctx = pkg.Context(D, 0)
F = 32
geo = Geometry(3840, 2160); pe = geo.plane_elems
vt = np.uint8 if D == 8 else np.int16
A = torch.from_numpy(make_plane(geo, D, 1, "natural").view(vt)).cuda().repeat(F)
B = torch.from_numpy(make_plane(geo, D, 2, "natural").view(vt)).cuda().repeat(F)
for S, Gs in ((8, (1, 2, 4)), (16, (2, 4, 8, 16)), (32, (4, 8, 16, 32)), (64, (8, 16, 32))):
    oF, oR5, _ = cu_descriptors(geo, S, *[tile_blocks(geo, w, h, seed=1) for (w, h) in ((S, S), (S, S // 2), (S // 2, S))])
    a = torch.from_numpy(np.concatenate([oF.astype(np.int64) + f * pe for f in range(F)]).astype(np.int32)).cuda()
    b = torch.from_numpy(np.concatenate([oR5.astype(np.int64) + f * pe for f in range(F)]).astype(np.int32)).cuda()
    out = torch.empty(5 * a.numel(), dtype=torch.int32, device="cuda")
    base = None
    for G in Gs:
        lanes = {8: [G, 4, 8, 32], 16: [2, G, 8, 32], 32: [2, 4, G, 32], 64: [2, 4, 8, G]}[S]
        os.environ["X265B200_CU_LANES_LAB"] = ",".join(map(str, lanes))
        for _ in range(3):
            ctx.cu_satd_batch(S, A, geo.stride, B, geo.stride, a, b, out)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(6):
            ctx.cu_satd_batch(S, A, geo.stride, B, geo.stride, a, b, out)
        e1.record(); torch.cuda.synchronize()
        chk = int(out.to(torch.int64).sum())
        base = chk if base is None else base
        print("S=%d G=%d  %.4f ms  %s  sum %d" % (S, G, e0.elapsed_time(e1) / 6, "ok" if chk == base else "MISMATCH", chk))
ctx.check()
