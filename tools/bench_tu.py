#!/usr/bin/env python
"""TU reconstruction chain timing: the single tcgen05 kernel (path 0, N = 32 / 16) against the two fused mma.sync kernels (path 2) and the
stage kernels (path 1), 2160p10 x F frames per launch, 8 algorithmic bytes per sample (fenc + pred in, qCoef + recon out)."""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from frames import Geometry, make_plane, tile_blocks
pkg = importlib.import_module("x265-mod-by-patman_b200")
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 10
F = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ctx = pkg.Context(depth, 0)
geo = Geometry(3840, 2160); pe = geo.plane_elems
vt = np.int16 if depth > 8 else np.uint8
A = torch.from_numpy(make_plane(geo, depth, 1, "natural").view(vt)).cuda().repeat(F)
B = torch.from_numpy(make_plane(geo, depth, 2, "natural").view(vt)).cuda().repeat(F)
recon = torch.empty_like(A)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
b = 1 if depth == 8 else 2
rows = []
for N in (32, 16, 8, 4):
    oa, ob = tile_blocks(geo, N, N, seed=1, merange=3)
    a = torch.from_numpy(np.concatenate([oa.astype(np.int64) + f * pe for f in range(F)]).astype(np.int32)).cuda()
    bb = torch.from_numpy(np.concatenate([ob.astype(np.int64) + f * pe for f in range(F)]).astype(np.int32)).cuda()
    n = a.numel(); S = n * N * N
    tshift = 15 - depth - {4: 2, 8: 3, 16: 4, 32: 5}[N]
    qbits = 14 + 5 + tshift
    qc = torch.full((N * N,), 26214, dtype=torch.int32, device="cuda")
    q = torch.empty(S, dtype=torch.int16, device="cuda"); ns = torch.empty(n, dtype=torch.int32, device="cuda")
    z = torch.empty(n, dtype=torch.int64, device="cuda"); r = torch.empty(n, dtype=torch.int64, device="cuda")
    for path in (3, 2):
        if path == 3 and N < 16:
            continue
        ctx.set_dct_path(path)
        run = lambda: ctx.tu_chain_batch(N, A, geo.stride, B, geo.stride, a, bb, qc, qbits, 171 << (qbits - 9), 40 << 5, max(1, 6 - tshift), q, ns, recon, geo.stride, a, z, r)
        for _ in range(2):
            run()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        gbs = S * (3 * b + 2) / (ms * 1e-3) / 1e9
        rows.append((N, path, ms, gbs, gbs / peak, int(ns.sum())))
        print("tu_chain %2dx%-2d %-28s %.4f ms  %.0f GB/s  %.2f of HBM (%.0f GB/s)  numSig total %d" % (N, N, "tcgen05 single kernel" if path == 3 else "mma.sync two kernels",
              ms, gbs, gbs / peak, peak, int(ns.sum())))
ctx.set_dct_path(0)
ctx.check()
