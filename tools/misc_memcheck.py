"""Tiny invocations of the per-CU SATD (both kernels) and of the TU chain (inter / intra luma, every size) on planes allocated with no slack,
blocks and vectors clipped exactly at the padded picture's edge: the target of
    compute-sanitizer --tool memcheck python tools/misc_memcheck.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from frames import Geometry, cu_descriptors, make_plane, tile_blocks     # noqa: E402
pkg = importlib.import_module("x265-mod-by-patman_b200")


def main():
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    ctx = pkg.Context(depth, 0)
    geo = Geometry(192, 128)
    vt = np.uint8 if depth == 8 else np.int16
    F = torch.from_numpy(make_plane(geo, depth, 1, "natural").view(vt)).cuda()
    R = torch.from_numpy(make_plane(geo, depth, 2, "uniform").view(vt)).cuda()
    m = geo.margin_x - 8
    for S in (8, 16, 32, 64):
        d = [tile_blocks(geo, w, h, seed=3 + S, merange=m) for (w, h) in ((S, S), (S, S // 2), (S // 2, S))]
        offF, offR5, _ = cu_descriptors(geo, S, *d)
        out = torch.zeros(5 * len(offF), dtype=torch.int32, device="cuda")
        ctx.cu_satd_batch(S, F, geo.stride, R, geo.stride, torch.from_numpy(offF).cuda(), torch.from_numpy(offR5).cuda(), out)
    for N in (4, 8, 16, 32):
        offF, offP = tile_blocks(geo, N, N, seed=9, merange=m)
        n = len(offF)
        qc = torch.full((N * N,), 16384, dtype=torch.int32, device="cuda")
        tshift = 15 - depth - {4: 2, 8: 3, 16: 4, 32: 5}[N]
        qbits = 14 + 4 + tshift
        q = torch.zeros(n * N * N, dtype=torch.int16, device="cuda"); ns = torch.zeros(n, dtype=torch.int32, device="cuda")
        z = torch.zeros(n, dtype=torch.int64, device="cuda"); r = torch.zeros(n, dtype=torch.int64, device="cuda")
        recon = torch.zeros_like(F)
        for ttype in (pkg.TU_INTER, pkg.TU_INTRA_LUMA):
            for path in (0, 2, 1):
                ctx.set_dct_path(path)
                ctx.tu_chain_batch(N, F, geo.stride, R, geo.stride, torch.from_numpy(offF).cuda(), torch.from_numpy(offP).cuda(), qc, qbits, 85 << (qbits - 9),
                                   64 << 4, 6 - tshift, q, ns, recon, geo.stride, torch.from_numpy(offF).cuda(), z, r, ttype=ttype)
        ctx.set_dct_path(0)
    torch.cuda.synchronize()
    ctx.check()
    print("misc_memcheck: all entries ran")


if __name__ == "__main__":
    main()
