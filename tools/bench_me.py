"""Exhaustive integer search (x265b200_me_full_batch) over one 2160p frame tiled into w x h PUs, +-merange window.
The kernel is issue-bound (one VABSDIFF per sample-candidate), not HBM-bound: the table reports sample-candidates/s and
the fraction of the SMs' integer issue rate (148 SMs x 128 lanes x SM clock), next to the reference's C loop on one host core.
usage: python tools/bench_me.py [--depth 10] [--merange 57] [--cpu]"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from frames import Geometry, make_plane, tile_blocks          # noqa: E402
from test_oracle_vs_ref import mv_cost_table                  # noqa: E402
pkg = importlib.import_module("x265-mod-by-patman_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=10)
    ap.add_argument("--merange", type=int, default=57)
    ap.add_argument("--subme", type=int, default=2)
    ap.add_argument("--methods", default="5,1", help="search methods for the whole-chain timing: 0 DIA, 1 HEX, 2 UMH, 3 STAR, 4 SEA, 5 FULL")
    ap.add_argument("--shapes", default="64,32,16,8", help="square PU sizes to run")
    ap.add_argument("--bidir", action="store_true", help="also time the bi-prediction cost entry")
    ap.add_argument("--lowres", action="store_true", help="also time the lookahead's lowres search")
    ap.add_argument("--cpu", action="store_true", help="also time the oracle's loop on a sample of PUs (one core)")
    args = ap.parse_args()
    D, M = args.depth, args.merange
    ctx = pkg.Context(D, 0)
    geo = Geometry(3840, 2160)
    cw, ch = geo.coded()
    vt = np.uint8 if D == 8 else np.int16
    Fh = make_plane(geo, D, 1, "natural"); Rh = make_plane(geo, D, 2, "natural")
    A = torch.from_numpy(Fh.view(vt)).cuda(); B = torch.from_numpy(Rh.view(vt)).cuda()
    RAD = 4096
    QP = 30
    refl = None
    if args.cpu:
        import cpulibs
        if cpulibs.have_reference(D):
            refl = cpulibs.Reference(D)
    tab = refl.mvcost_table(QP, RAD) if refl else mv_cost_table(12.6992, RAD)      # the reference's own table when it is built
    dtab = torch.from_numpy(tab.view(np.int16)).cuda()
    sm_clock = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else 0
    rows = []
    sums, integral_ms = None, None
    for (w, h) in [(int(v), int(v)) for v in args.shapes.split(",")]:
        oa, _ = tile_blocks(geo, w, h, seed=1)
        n = oa.size
        px = (oa - geo.origin) % geo.stride; py = (oa - geo.origin) // geo.stride
        minx = -np.minimum(M, px + geo.margin_x - 8); maxx = np.minimum(M, cw + geo.margin_x - 8 - w - px)
        miny = -np.minimum(M, py + geo.margin_y - 8); maxy = np.minimum(M, ch + geo.margin_y - 8 - h - py)
        rng = np.stack([minx, miny, maxx, maxy], 1).astype(np.int32).copy()
        mvp = np.zeros((n, 2), np.int32)
        cands = int(((maxx - minx + 1) * (maxy - miny + 1)).sum())
        d = [torch.from_numpy(a).cuda() for a in (oa.astype(np.int32), rng, mvp)]
        bmv = torch.zeros((n, 2), dtype=torch.int32, device="cuda"); bc = torch.full((n,), 0x7fffffff, dtype=torch.int32, device="cuda")

        def run():
            bc.fill_(0x7fffffff)
            ctx.me_full_batch(w, h, M, A, geo.stride, B, geo.stride, d[0], d[0], d[1], d[2], dtab.data_ptr() + 2 * RAD, bmv, bc)
        run(); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[1]
        row = {"pu": "%dx%d" % (w, h), "n_pu": int(n), "merange": M, "candidates": cands, "ms_per_frame": ms,
               "G_sample_candidates_per_s": cands * w * h / ms / 1e6, "M_candidates_per_s": cands / ms / 1e3}
        # the whole motionEstimate chain (2 neighbour candidates, subme 2 = the medium preset's workload)
        qmvp = torch.randint(-40, 41, (n, 2), dtype=torch.int32, device="cuda")
        mvc = torch.randint(-40, 41, (n, 2, 2), dtype=torch.int32, device="cuda")
        oq = torch.zeros((n, 2), dtype=torch.int32, device="cuda"); oc = torch.zeros((n,), dtype=torch.int32, device="cuda")

        for method in [int(v) for v in args.methods.split(",")]:
            if method == 4 and (w, h) in ((32, 8), (8, 32), (8, 4), (4, 8)):
                continue
            if method == 4 and sums is None:
                # successive elimination reads the twelve integral planes of the reference picture (framefilter.cpp:737-833); built once, timed
                pitch = geo.plane_elems
                sums = torch.empty(12 * pitch, dtype=torch.int32, device="cuda")
                ctx.me_integral_batch(B, geo.stride, geo.rows, 1, sums, pitch); torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); ctx.me_integral_batch(B, geo.stride, geo.rows, 1, sums, pitch); e1.record(); e1.synchronize()
                integral_ms = e0.elapsed_time(e1)

            def run_me():
                if method == 4:
                    ctx.motion_estimate_sea_batch(w, h, M, args.subme, A, geo.stride, B, geo.stride, d[0], d[0], d[1], qmvp, 2, mvc,
                                                  dtab.data_ptr() + 2 * RAD, sums, geo.plane_elems, oq, oc)
                else:
                    ctx.motion_estimate_batch(method, w, h, M, args.subme, A, geo.stride, B, geo.stride, d[0], d[0], d[1], qmvp, 2, mvc,
                                              dtab.data_ptr() + 2 * RAD, oq, oc)
            run_me(); torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); run_me(); e1.record(); e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            name = {0: "dia", 1: "hex", 2: "umh", 3: "star", 4: "sea", 5: "full"}[method]
            row["motion_estimate_%s_ms_per_frame" % name] = sorted(ts)[1]
            if method == 4:
                row["sea_integral_planes_ms_per_reference_picture"] = integral_ms
            if refl:
                # the reference's own MotionEstimate::motionEstimate on every host core, on a sample of the same PUs
                k = min(n, 400 if method == 5 else 4000 if method == 4 else 40000)
                if method == 5: k = max(8, min(k, int(4e9 / (cands / n * w * h))))
                if method == 4:
                    continue                                 # the reference's SEA needs its own integral planes: pinned in the CPU suite, not timed here
                sel = np.linspace(0, n - 1, k).astype(np.int64)
                hq = qmvp.cpu().numpy()[sel].copy(); hm = mvc.cpu().numpy()[sel].copy()
                so = oa[sel].astype(np.int32)
                nt = os.cpu_count() or 1
                t0 = time.perf_counter()
                rmv, rc = refl.motion_estimate_batch(method, args.subme, w, h, Fh, geo.stride, so, Rh, geo.stride, so, rng[sel].copy(), hq, 2, hm, M, QP, nt)
                dt = time.perf_counter() - t0
                row["ref_%s_us_per_pu_%d_threads" % (name, nt)] = dt / k * 1e6
                row["ref_%s_ms_per_frame_extrapolated" % name] = dt / k * n * 1e3
                row["ref_%s_sample_matches" % name] = bool(np.array_equal(oq.cpu().numpy()[sel], rmv) and np.array_equal(oc.cpu().numpy()[sel], rc))
        row["motion_estimate_subme"] = args.subme
        if args.cpu:
            orc = cpulibs.Oracle(D)
            k = max(1, min(n, int(2e9 / (cands / n * w * h))))       # ~2 G sample-candidates of CPU work
            sel = np.linspace(0, n - 1, k).astype(np.int64)
            mv = np.zeros((k, 2), np.int32); c = np.full(k, 0x7fffffff, np.int32)
            t0 = time.perf_counter()
            orc.me_full_batch(w, h, Fh, geo.stride, Rh, geo.stride, oa[sel].astype(np.int32), oa[sel].astype(np.int32), rng[sel].copy(), mvp[sel].copy(),
                              tab, RAD, mv, c)
            dt = time.perf_counter() - t0
            cc = int(((maxx - minx + 1) * (maxy - miny + 1))[sel].sum())
            row["cpu_port_1core_G_sample_candidates_per_s"] = cc * w * h / dt / 1e9
            got_mv = bmv.cpu().numpy()[sel]; got_c = bc.cpu().numpy()[sel]
            row["cpu_sample_matches"] = bool(np.array_equal(got_mv, mv) and np.array_equal(got_c, c))
        rows.append(row)
        print(json.dumps(row), file=sys.stderr)
    if args.bidir:
        # bi-prediction cost of every 16x16 PU of the frame: two quarter-pel vectors per PU (search.cpp:442-448)
        oa, ob = tile_blocks(geo, 16, 16, seed=1)
        _, oc_ = tile_blocks(geo, 16, 16, seed=2)
        n = oa.size
        g = np.random.default_rng(4)
        f0 = torch.from_numpy((g.integers(0, 4, n) | (g.integers(0, 4, n) << 4)).astype(np.int32)).cuda()
        f1 = torch.from_numpy((g.integers(0, 4, n) | (g.integers(0, 4, n) << 4)).astype(np.int32)).cuda()
        dA, dB0, dB1 = (torch.from_numpy(a.astype(np.int32)).cuda() for a in (oa, ob, oc_))
        cost = torch.zeros(n, dtype=torch.int32, device="cuda")

        def run_bi():
            ctx.bidir_satd_batch(16, 16, A, geo.stride, dA, B, geo.stride, dB0, f0, B, geo.stride, dB1, f1, cost)
        run_bi(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); run_bi(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        row = {"bidir_satd": "16x16 PUs of one 2160p frame, random quarter-pel vectors", "n_pu": int(n), "ms_per_frame": sorted(ts)[2],
               "G_samples_per_s": n * 256 / sorted(ts)[2] / 1e6}
        rows.append(row)
        print(json.dumps(row), file=sys.stderr)
    if args.lowres:
        # the lookahead's search: 1920x1080 lowres frame, every 8x8 block, HEX, subme 1, merange 16 (slicetype.cpp:4484-4566)
        from test_oracle_vs_ref import lowres_planes
        lgeo = Geometry(1920, 1080)
        P = lowres_planes(lgeo, D, 7); pitch = lgeo.plane_elems
        Fl = np.roll(P[:pitch], 3 * lgeo.stride - 5).copy()
        dPl = torch.from_numpy(P.view(vt)).cuda(); dFl = torch.from_numpy(Fl.view(vt)).cuda()
        oa, _ = tile_blocks(lgeo, 8, 8, seed=1)
        n = oa.size; lcw, lch = lgeo.coded(); LM = 16
        px = (oa - lgeo.origin) % lgeo.stride; py = (oa - lgeo.origin) // lgeo.stride
        rng = np.stack([-np.minimum(LM, px + lgeo.margin_x - 8), -np.minimum(LM, py + lgeo.margin_y - 8),
                        np.minimum(LM, lcw + lgeo.margin_x - 16 - px), np.minimum(LM, lch + lgeo.margin_y - 16 - py)], 1).astype(np.int32).copy()
        hq = np.random.default_rng(3).integers(-12, 13, (n, 2)).astype(np.int32)
        d = [torch.from_numpy(a).cuda() for a in (oa.astype(np.int32), rng, hq)]
        oq = torch.zeros((n, 2), dtype=torch.int32, device="cuda"); oc = torch.zeros((n,), dtype=torch.int32, device="cuda")

        def run_lr():
            ctx.lowres_motion_estimate_batch(1, 8, 8, LM, 1, dFl, lgeo.stride, dPl, lgeo.stride, pitch, d[0], d[0], d[1], d[2], dtab.data_ptr() + 2 * RAD, oq, oc)
        run_lr(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); run_lr(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        row = {"lowres": "1920x1080, 8x8, HEX, subme 1, merange 16", "n_cu": int(n), "ms_per_frame": sorted(ts)[2]}
        if refl:
            k = 3000
            sel = np.linspace(0, n - 1, k).astype(np.int64)
            gq = oq.cpu().numpy(); gcst = oc.cpu().numpy()
            t0 = time.perf_counter(); ok = True
            for i in sel:
                a = refl.lowres_motion_estimate_ref(1, 1, 8, 8, Fl, int(oa[i]), lgeo.stride, P, int(oa[i]), lgeo.stride, pitch, rng[i], hq[i], LM, QP)
                ok = ok and a == (int(gq[i, 0]), int(gq[i, 1]), int(gcst[i]))
            dt = time.perf_counter() - t0
            row["ref_us_per_cu_1_thread_incl_ctypes"] = dt / k * 1e6
            row["ref_sample_matches"] = bool(ok)
        rows.append(row)
        print(json.dumps(row), file=sys.stderr)
    ctx.check()
    print(json.dumps({"depth": D, "rows": rows}))


if __name__ == "__main__":
    main()
