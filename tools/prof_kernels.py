#!/usr/bin/env python
"""Launch a chosen list of primitives a few times each on 2160p10-sized batches (no timing): the target of
`ncu --set full -k regex:<kernel>` captures, so one GPU call profiles several kernels.

    python tools/prof_kernels.py [--frames 8] [--reps 2] satd64 satd8 sa8d16 dct8 idct8 idct4 hvpp16 vpp64 ...
"""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from frames import Geometry, make_plane, tile_blocks  # noqa: E402

pkg = importlib.import_module("x265-mod-by-patman_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--depth", type=int, default=10)
    ap.add_argument("what", nargs="+")
    args = ap.parse_args()
    D, F = args.depth, args.frames
    ctx = pkg.Context(D, 0)
    geo = Geometry(3840, 2160)
    vt = np.uint8 if D == 8 else np.int16
    pe = geo.plane_elems
    cw, ch = geo.coded()
    S = F * cw * ch
    A = torch.from_numpy(np.concatenate([make_plane(geo, D, 1 + f, "uniform") for f in range(F)]).view(vt)).cuda()
    B = torch.from_numpy(np.concatenate([make_plane(geo, D, 101 + f, "uniform") for f in range(F)]).view(vt)).cuda()

    def desc(w, h):
        oa, ob = tile_blocks(geo, w, h, seed=1)
        a = np.concatenate([oa.astype(np.int64) + f * pe for f in range(F)]).astype(np.int32)
        b = np.concatenate([ob.astype(np.int64) + f * pe for f in range(F)]).astype(np.int32)
        return torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()

    res = torch.empty(S, dtype=torch.int16, device="cuda")
    oa32, ob32 = desc(32, 32)
    ctx.residual_batch(32, 32, A, geo.stride, B, geo.stride, oa32, ob32, res)
    coef = torch.empty(S, dtype=torch.int16, device="cuda")
    rec = torch.empty(S, dtype=torch.int16, device="cuda")
    dstP = torch.empty(F * pe, dtype=A.dtype, device="cuda")
    dstS = torch.empty(F * pe, dtype=torch.int16, device="cuda")
    ops = {"sad": 0, "satd": 1, "sa8d": 2, "sse": 3}
    for name in args.what:
        for _ in range(args.reps):
            kind = name.rstrip("0123456789")
            N = int(name[len(kind):])
            if kind in ops:
                oa, ob = desc(N, N)
                out = torch.empty(oa.numel(), dtype=torch.int64 if kind == "sse" else torch.int32, device="cuda")
                ctx.pixelcmp_batch(ops[kind], N, N, A, geo.stride, B, geo.stride, oa, ob, out)
            elif kind == "dct":
                ctx.dct_batch(pkg.TR_DCT, N, res, N, None, coef, count=S // (N * N))
            elif kind == "idct":
                ctx.idct_batch(pkg.TR_DCT, N, res, rec, N, None, count=S // (N * N))
            elif kind == "tuchain":
                oa, ob = desc(N, N)
                n = oa.numel()
                qcN = torch.full((N * N,), 16384, dtype=torch.int32, device="cuda")
                tshift = 15 - D - {4: 2, 8: 3, 16: 4, 32: 5}[N]
                qbits = 14 + 4 + tshift
                qo = torch.empty(S, dtype=torch.int16, device="cuda"); ns = torch.empty(n, dtype=torch.int32, device="cuda")
                z = torch.empty(n, dtype=torch.int64, device="cuda"); r = torch.empty(n, dtype=torch.int64, device="cuda")
                ctx.tu_chain_batch(N, A, geo.stride, B, geo.stride, oa, ob, qcN, qbits, 85 << (qbits - 9), 64 << 4, 6 - tshift,
                                   qo, ns, dstP, geo.stride, oa, z, r)
            elif kind == "residual":
                oa, ob = desc(N, N)
                ctx.residual_batch(N, N, A, geo.stride, B, geo.stride, oa, ob, res)
            elif kind in ("hpp", "vpp", "hps", "vps", "p2s", "hvpp"):
                oa, _ = desc(N, N)
                n = oa.numel()
                idx = torch.randint(1, 4, (n,), dtype=torch.int32, device="cuda")
                if kind == "hvpp":
                    idx = idx | (torch.randint(1, 4, (n,), dtype=torch.int32, device="cuda") << 4)
                dst = dstP if kind in ("hpp", "vpp", "hvpp") else dstS
                ctx.interp_batch(kind, 8, N, N, A, geo.stride, oa, dst, geo.stride, oa, idx)
            else:
                raise SystemExit("unknown primitive " + name)
        torch.cuda.synchronize()
    ctx.check()
    print("launched:", " ".join(args.what))


if __name__ == "__main__":
    main()
