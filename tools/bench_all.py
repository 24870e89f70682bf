#!/usr/bin/env python
"""Per-primitive-class throughput on 2160p10-sized batches: achieved algorithmic GB/s vs the measured HBM peak.
Complements bench.py (which times the SATD+DCT headline step).  Output: one JSON line per row + a markdown table.

    python tools/bench_all.py [--depth 10] [--frames 8] > profiles/<name>.md
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from frames import Geometry, make_plane, tile_blocks  # noqa: E402

pkg = importlib.import_module("x265-mod-by-patman_b200")


def timeit(fn, reps=8, warm=3, burst=6):
    """median over `reps` bursts of `burst` back-to-back launches (per-launch mean of a burst): the stream stays
    busy, so the host's launch latency on an idle GPU is not part of the number (every launch streams > L2 bytes)"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        fn()
        e0.record()
        for _ in range(burst):
            fn()
        e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) / burst)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=10)
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--only", default="", help="comma list of sections: metrics,transforms,quant,tu,cusatd,intra,subpel,adjacent,weight,integral,interp (default: all)")
    args = ap.parse_args()
    D, F = args.depth, args.frames
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    ctx = pkg.Context(D, 0)
    geo = Geometry(3840, 2160)
    b = 1 if D == 8 else 2
    vt = np.uint8 if D == 8 else np.int16
    pe = geo.plane_elems
    cw, ch = geo.coded()
    S = F * cw * ch
    A = torch.from_numpy(np.concatenate([make_plane(geo, D, 1 + f, "natural") for f in range(F)]).view(vt)).cuda()
    B = torch.from_numpy(np.concatenate([make_plane(geo, D, 101 + f, "natural") for f in range(F)]).view(vt)).cuda()
    rows = []
    only = [x for x in args.only.split(",") if x]

    def want(name):
        return not only or name in only

    def desc(w, h):
        oa, ob = tile_blocks(geo, w, h, seed=1)
        a = np.concatenate([oa.astype(np.int64) + f * pe for f in range(F)]).astype(np.int32)
        bb = np.concatenate([ob.astype(np.int64) + f * pe for f in range(F)]).astype(np.int32)
        return torch.from_numpy(a).cuda(), torch.from_numpy(bb).cuda()

    def add(name, ms, nbytes, units):
        gbs = nbytes / ms / 1e6
        rows.append({"primitive": name, "ms": ms, "algorithmic_GB": nbytes / 1e9, "GBps": gbs, "frac_of_measured_hbm": gbs / peak,
                     "Gunits_per_s": units / ms / 1e6})
        print(json.dumps(rows[-1]), file=sys.stderr)

    # ---- block metrics
    if want("metrics"):
        for (w, h) in ((64, 64), (32, 32), (16, 16), (8, 8), (4, 4)):
            oa, ob = desc(w, h)
            n = oa.numel()
            o32 = torch.empty(n, dtype=torch.int32, device="cuda"); o64 = torch.empty(n, dtype=torch.int64, device="cuda")
            for op, nm, out, ob_ in ((0, "sad", o32, 4), (1, "satd", o32, 4), (2, "sa8d", o32, 4), (3, "sse_pp", o64, 8)):
                ms = timeit(lambda: ctx.pixelcmp_batch(op, w, h, A, geo.stride, B, geo.stride, oa, ob, out))
                add("%s %dx%d" % (nm, w, h), ms, S * 2 * b + n * ob_, S)
            if w in (16, 64):
                K = 4
                offR = torch.stack([ob + k for k in range(K)], dim=1).reshape(-1).contiguous()
                oK = torch.empty(n * K, dtype=torch.int32, device="cuda")
                ms = timeit(lambda: ctx.sad_multi_batch(w, h, A, geo.stride, B, geo.stride, oa, offR, K, oK))
                add("sad_x4 %dx%d (per candidate sample)" % (w, h), ms, S * (1 + K) * b + n * K * 4, S * K)
    # ---- transforms on block-contiguous int16
    res = torch.empty(S, dtype=torch.int16, device="cuda")
    coef = torch.empty(S, dtype=torch.int16, device="cuda")
    rec = torch.empty(S, dtype=torch.int16, device="cuda")
    oa32, ob32 = desc(32, 32)
    ctx.residual_batch(32, 32, A, geo.stride, B, geo.stride, oa32, ob32, res)
    ctx.dct_batch(pkg.TR_DCT, 32, res, 32, None, coef, count=S // 1024)
    if want("transforms"):
        oa, ob = desc(32, 32)
        ms = timeit(lambda: ctx.residual_batch(32, 32, A, geo.stride, B, geo.stride, oa, ob, res))
        add("residual (sub_ps) 32x32", ms, S * (2 * b + 2), S)
        for N in (32, 16, 8, 4):
            n = S // (N * N)
            off = torch.arange(n, dtype=torch.int32, device="cuda") * (N * N)
            ms = timeit(lambda: ctx.dct_batch(pkg.TR_DCT, N, res, N, None, coef, count=n))
            add("dct%d (IMMA)" % N, ms, S * 4, S)
            ctx.set_dct_path(1)
            ms = timeit(lambda: ctx.dct_batch(pkg.TR_DCT, N, res, N, None, coef, count=n))
            add("dct%d (CUDA-core twin)" % N, ms, S * 4, S)
            ctx.set_dct_path(0)
            ms = timeit(lambda: ctx.idct_batch(pkg.TR_DCT, N, coef, rec, N, None, count=n))
            add("idct%d (contiguous TUs)" % N, ms, S * 4, S)
            ms = timeit(lambda: ctx.idct_batch(pkg.TR_DCT, N, coef, rec, N, off))
            add("idct%d (per-TU offsets, +%.1f B/coef descriptor traffic)" % (N, 4.0 / (N * N)), ms, S * 4, S)
        n4 = S // 16
        ms = timeit(lambda: ctx.dct_batch(pkg.TR_DST, 4, res, 4, None, coef, count=n4))
        add("dst4 (IMMA)", ms, S * 4, S)
    # ---- quant family, 32x32 blocks
    if want("quant"):
        nb = S // 1024
        qc = torch.full((1024,), 26214, dtype=torch.int32, device="cuda")
        q = torch.empty(S, dtype=torch.int16, device="cuda"); du = torch.empty(S, dtype=torch.int32, device="cuda")
        sig = torch.empty(nb, dtype=torch.int32, device="cuda")
        ms = timeit(lambda: ctx.quant_batch(coef, qc, du, q, 21, 171 << 12, 1024, nb, sig)); add("quant 32x32", ms, S * 8, S)
        ms = timeit(lambda: ctx.quant_batch(coef, qc, None, q, 21, 1 << 20, 1024, nb, sig)); add("nquant 32x32", ms, S * 4, S)
        ms = timeit(lambda: ctx.dequant_normal_batch(q, coef, S, 64 << 4, 5)); add("dequant_normal", ms, S * 4, S)
        ms = timeit(lambda: ctx.dequant_scaling_batch(q, qc, coef, 1024, nb, 4, 2)); add("dequant_scaling 32x32", ms, S * 4, S)
    # ---- fused inter-luma TU chain (sub_ps, dct, quant, dequant, idct, add_ps, sse): 3b + 2 bytes per sample
    if want("tu"):
        recon = torch.empty(F * pe, dtype=A.dtype, device="cuda")
        for N in (32, 16, 8, 4):
            oa, ob = desc(N, N)
            n = oa.numel()
            qcN = torch.full((N * N,), 16384, dtype=torch.int32, device="cuda")
            tshift = 15 - D - {4: 2, 8: 3, 16: 4, 32: 5}[N]
            qbits = 14 + 4 + tshift                                  # qp 28
            qo = torch.empty(S, dtype=torch.int16, device="cuda"); ns = torch.empty(n, dtype=torch.int32, device="cuda")
            z = torch.empty(n, dtype=torch.int64, device="cuda"); r = torch.empty(n, dtype=torch.int64, device="cuda")
            for path, label in ((0, "default: tcgen05 single kernel" if N == 32 else "default: two fused mma.sync kernels"), (2, "two fused mma.sync kernels"),
                                (3, "tcgen05 single kernel")):
                if (path == 2 and N != 32) or (path == 3 and N != 16):
                    continue
                ctx.set_dct_path(path)
                ms = timeit(lambda: ctx.tu_chain_batch(N, A, geo.stride, B, geo.stride, oa, ob, qcN, qbits, 85 << (qbits - 9), 64 << 4, 6 - tshift,
                                                       qo, ns, recon, geo.stride, oa, z, r), reps=5, warm=2)
                add("tu_chain %dx%d (%s)" % (N, N, label), ms, S * (3 * b + 2) + n * 20, S)
            ctx.set_dct_path(0)
    # ---- all rectangular PUs of a CU in one pass (fenc once + three independently displaced reference blocks = 4b bytes per CU sample)
    if want("cusatd"):
        from frames import cu_descriptors
        for Scu in (64, 32, 16, 8):
            oF, oR5, _ = cu_descriptors(geo, Scu, *[tile_blocks(geo, w, h, seed=1) for (w, h) in ((Scu, Scu), (Scu, Scu // 2), (Scu // 2, Scu))])
            a_ = torch.from_numpy(np.concatenate([oF.astype(np.int64) + f * pe for f in range(F)]).astype(np.int32)).cuda()
            b_ = torch.from_numpy(np.concatenate([oR5.astype(np.int64) + f * pe for f in range(F)]).astype(np.int32)).cuda()
            o_ = torch.empty(5 * a_.numel(), dtype=torch.int32, device="cuda")
            ms = timeit(lambda: ctx.cu_satd_batch(Scu, A, geo.stride, B, geo.stride, a_, b_, o_))
            add("cu_satd %dx%d CU (2Nx2N + 2NxN + Nx2N PUs, five vectors; per shape-pass sample)" % (Scu, Scu), ms, S * 4 * b + a_.numel() * 20, 3 * S)
    # ---- intra: all 35 predictions of N x N TUs from their neighbour arrays (output-bound: 35 b bytes per sample), and the lookahead's intra estimate
    if want("intra"):
        for N in (32, 8):
            nt = min(S // (N * N), 40000)
            nb_ = torch.randint(0, 1 << D, (nt * (4 * N + 1),), dtype=torch.int32, device="cuda").to(A.dtype)
            dstI = torch.empty(nt * 35 * N * N, dtype=A.dtype, device="cuda")
            ms = timeit(lambda: ctx.intra_pred_batch(N, nb_, nt, dstI))
            add("intra_pred_all %dx%d (35 modes per TU; 35 b bytes written per sample)" % (N, N), ms, nt * N * N * 35 * b, nt * N * N * 35)
            del dstI
        lw_, lh_ = cw // 2, ch // 2                    # a lowres frame lives in the corner of a full-size plane here: only the geometry matters
        ci = torch.empty((lw_ // 8) * (lh_ // 8), dtype=torch.int32, device="cuda"); mi = torch.empty_like(ci)
        base = geo.origin
        ms = timeit(lambda: ctx.lowres_intra_batch(A, base, geo.stride, lw_ // 8, lh_ // 8, 12, ci, mi))
        add("lowres_intra (35 modes x 8x8 SATD per CU, one 1080p-lowres frame; per CU sample)", ms, lw_ * lh_ * b, lw_ * lh_)
    # ---- sub-pel candidate cost (interpolation fused with SATD): K = 4 quarter-pel candidates per block
    if want("subpel"):
        for (w, h) in ((64, 64), (16, 16), (8, 8)):
            oa, ob = desc(w, h)
            n = oa.numel()
            K = 4
            offR = torch.stack([ob + k for k in range(K)], dim=1).reshape(-1).contiguous()
            frac = (torch.randint(0, 4, (n * K,), dtype=torch.int32, device="cuda") | (torch.randint(0, 4, (n * K,), dtype=torch.int32, device="cuda") << 4))
            cost = torch.empty(n * K, dtype=torch.int32, device="cuda")
            ms = timeit(lambda: ctx.subpel_cmp_batch(1, w, h, A, geo.stride, B, geo.stride, oa, offR, frac, K, cost), reps=4, warm=2, burst=3)
            add("subpel satd %dx%d (fused hv interp + satd, per candidate sample; window + fenc/K bytes)" % (w, h), ms,
                S * K * b * ((w + 7) * (h + 7) / (w * h) + 1.0 / K) + n * K * 4, S * K)
    # ---- adjacent slots: residual add / bi-prediction averages over a 32x32 tiling, and the lowres downscale
    if want("adjacent"):
        oa, ob = desc(32, 32)
        n = oa.numel()
        s16 = torch.randint(-2000, 2000, (F * pe,), dtype=torch.int16, device="cuda")
        outP = torch.empty(F * pe, dtype=A.dtype, device="cuda"); outS = torch.empty(F * pe, dtype=torch.int16, device="cuda")
        for op, nm, a_, b_, d_, nb_ in ((0, "sub_ps", A, B, outS, 2 * b + 2), (1, "add_ps", A, s16, outP, 2 * b + 2),
                                       (2, "pixelavg_pp", A, B, outP, 3 * b), (3, "addAvg", s16, s16, outP, 4 + b)):
            ms = timeit(lambda: ctx.blockop_batch(op, 32, 32, a_, geo.stride, oa, b_, geo.stride, ob, d_, geo.stride, oa, n))
            add("%s 32x32 (adjacent slot)" % nm, ms, S * nb_, S)
        lw, lh = cw // 2, ch // 2
        low = [torch.empty(F * lw * lh, dtype=A.dtype, device="cuda") for _ in range(4)]

        def lowres_all():
            for f in range(F):
                ctx.lowres_batch(A[f * pe + geo.origin:], geo.stride, *[t[f * lw * lh:] for t in low], lw, lw, lh)
        ms = timeit(lowres_all, reps=4, warm=1, burst=2)
        add("frameInitLowres (one launch per frame, %d launches)" % F, ms, S * 2 * b, S)
    # ---- lookahead weighted-prediction cost: K = 8 candidate weights over one frame pair per launch (8 frames per timing call)
    if want("weight"):
        K = 8
        wts = torch.tensor(sum([[64 + k, 32 << (14 - D), 6 + 14 - D, k - 3] for k in range(K)], []), dtype=torch.int32, device="cuda")
        wcost = torch.empty(K, dtype=torch.int32, device="cuda")
        intra = torch.randint(0, 1 << 20, ((cw // 8) * (ch // 8),), dtype=torch.int32, device="cuda")
        nfw = min(F, 8)

        def wc_all():
            for f in range(nfw):
                ctx.weight_cost_batch(A[f * pe + geo.origin:], B[f * pe + geo.origin:], geo.stride, cw, ch, intra, wts, K, wcost)
        ms = timeit(wc_all, reps=4, warm=1, burst=2)
        add("weight_cost (K = 8 weights fused with 8x8 SATD, per candidate sample; 2b bytes per sample for all K)", ms, nfw * cw * ch * 2 * b, nfw * cw * ch * K)
    # ---- SEA integral planes: 12 uint32 planes per picture from one read of the picture (4 frames per launch: 1.8 GB out)
    if want("integral"):
        nfi = min(F, 4)
        isum = torch.empty(nfi * 12 * pe, dtype=torch.int32, device="cuda")
        ms = timeit(lambda: ctx.me_integral_batch(A, geo.stride, geo.rows, nfi, isum, pe), reps=4, warm=1, burst=3)
        add("me_integral (12 SEA planes, %d padded frames per launch; b + 48 B per padded sample)" % nfi, ms, nfi * pe * (b + 48), nfi * pe)
        del isum
    # ---- interpolation, frame-tiled 64x64 and 16x16 luma, 8x8 chroma-size blocks
    if want("interp"):
        dstP = torch.empty(F * pe, dtype=A.dtype, device="cuda"); dstS = torch.empty(F * pe, dtype=torch.int16, device="cuda")
        srcS = torch.randint(-8192, 8192, (F * pe,), dtype=torch.int16, device="cuda")
        for taps, (w, h) in ((8, (64, 64)), (8, (16, 16)), (4, (8, 8))):
            oa, _ = desc(w, h)
            n = oa.numel()
            idx = torch.randint(1, 4, (n,), dtype=torch.int32, device="cuda")
            idxhv = idx | (torch.randint(1, 4, (n,), dtype=torch.int32, device="cuda") << 4)
            tag = "%s %dx%d" % ("luma" if taps == 8 else "chroma", w, h)
            for kind, src, dst, inb, outb in (("hpp", A, dstP, b, b), ("vpp", A, dstP, b, b), ("hps", A, dstS, b, 2), ("vps", A, dstS, b, 2),
                                              ("vsp", srcS, dstP, 2, b), ("vss", srcS, dstS, 2, 2), ("p2s", A, dstS, b, 2)):
                ms = timeit(lambda: ctx.interp_batch(kind, taps, w, h, src, geo.stride, oa, dst, geo.stride, oa, idx))
                add("%s %s" % (kind, tag), ms, S * (inb + outb), S)
            if taps == 8:
                ms = timeit(lambda: ctx.interp_batch("hvpp", 8, w, h, A, geo.stride, oa, dstP, geo.stride, oa, idxhv))
                add("hvpp %s" % tag, ms, S * 2 * b, S)
    ctx.check()
    print("| primitive (2160p%d x %d frames per launch) | ms | GB/s (algorithmic) | of %s HBM %.0f GB/s | G samples/s |" % (D, F, "measured (MEASURED_PEAKS.json)" if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else "FALLBACK (B200_PROFILING.md)", peak))
    print("|---|---|---|---|---|")
    for r in rows:
        print("| %s | %.4f | %.0f | %.2f | %.0f |" % (r["primitive"], r["ms"], r["GBps"], r["frac_of_measured_hbm"], r["Gunits_per_s"]))


if __name__ == "__main__":
    main()
