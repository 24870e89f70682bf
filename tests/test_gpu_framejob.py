"""The host-buffer layer of the C ABI (planes + frame jobs, include/x265b200.h) against the oracle / the reference:
host pictures in, host results out, nothing but numpy on this side of the boundary."""
import os
import subprocess

import numpy as np
import pytest

from cpulibs import OP_SAD, OP_SATD, OP_SA8D, Oracle, Reference, have_reference
from frames import Geometry, make_plane, tile_blocks

pytestmark = pytest.mark.gpu

FLAT = [26214, 23302, 20560, 18396, 16384, 14564]          # s_quantScales, reference common/scalinglist.cpp


def quant_params(depth, N, qp):
    """flat quant table, qBits, add of Quant::transformNxN for an inter TU (quant.cpp:465-466, rounding 171 for P/B slices)"""
    per, rem = qp // 6, qp % 6
    tshift = 15 - depth - {4: 2, 8: 3, 16: 4, 32: 5}[N]
    qbits = 14 + per + tshift
    return np.full(N * N, FLAT[rem], np.int32), qbits, 171 << (qbits - 9)


def checker(depth):
    return Reference(depth) if have_reference(depth) else Oracle(depth)


@pytest.mark.parametrize("depth", [8, 10, 12])
def test_plane_geometry_upload_and_border_extension(depth):
    from gpulib import context, pkg
    ctx = context(depth); orc = Oracle(depth)
    for (w, h, ctu) in ((352, 288, 32), (416, 240, 64), (1920, 1080, 64)):
        geo = Geometry(w, h, ctu)
        pl = pkg.Plane(ctx, w, h, ctu)
        assert (pl.stride, pl.rows, pl.origin, pl.elems) == (geo.stride, geo.rows, geo.origin, geo.plane_elems)
        # whole padded plane: bytes round-trip
        A = make_plane(geo, depth, 5)
        pl.upload_padded(A)
        assert np.array_equal(pl.download_padded(), A)
        # picture only (from a host buffer with its own stride): the margins are formed on the device like extendPicBorder
        hs = w + 13
        pic = make_plane(Geometry(w, h, ctu), depth, 6)[:hs * h].copy()
        pl2 = pkg.Plane(ctx, w, h, ctu)
        pl2.upload_picture(pic, hs)
        want = np.zeros(geo.plane_elems, orc.pix)
        want.reshape(geo.rows, geo.stride)[geo.margin_y:geo.margin_y + h, geo.margin_x:geo.margin_x + w] = pic.reshape(h, hs)[:, :w]
        orc.extend_pic_border(want, geo.origin, geo.stride, w, h, geo.margin_x, geo.margin_y)
        assert np.array_equal(pl2.download_padded(), want), (w, h, ctu)
        # the same picture sitting inside a padded host buffer with arbitrary margins: rows upload
        padded = make_plane(geo, depth, 8)
        padded.reshape(geo.rows, geo.stride)[geo.margin_y:geo.margin_y + h, geo.margin_x:geo.margin_x + w] = pic.reshape(h, hs)[:, :w]
        pl3 = pkg.Plane(ctx, w, h, ctu)
        pl3.upload_rows(padded)
        want3 = np.zeros(geo.plane_elems, orc.pix)
        want3.reshape(geo.rows, geo.stride)[geo.margin_y:geo.margin_y + h] = padded.reshape(geo.rows, geo.stride)[geo.margin_y:geo.margin_y + h]
        orc.extend_pic_border(want3, geo.origin, geo.stride, w, h, geo.margin_x, geo.margin_y)
        assert np.array_equal(pl3.download_padded(), want3), (w, h, ctu)
        pl.destroy(); pl2.destroy(); pl3.destroy()
    # chroma plane of a 4:2:0 picture (picyuv.cpp:106-110)
    pc = pkg.Plane(ctx, 1920, 1080, 64, 1, 1)
    assert pc.stride == 1920 // 2 + 2 * 96 and pc.rows == 1088 // 2 + 2 * 40 and pc.origin == 40 * pc.stride + 96
    pc.destroy()
    ctx.check()


@pytest.mark.parametrize("depth", [8, 10, 12])
def test_frame_job_all_pass_kinds(depth):
    """metric, dense-coefficient and sparse-level passes of several frames in flight vs the CPU checker"""
    from gpulib import context, pkg
    ctx = context(depth); chk = checker(depth); orc = Oracle(depth)
    geo = Geometry(640, 384)
    job = pkg.FrameJob(ctx, 640, 384, 64, slots=3)
    cmps = [(OP_SATD, 16, 16), (OP_SATD, 8, 4), (OP_SAD, 32, 16), (OP_SA8D, 8, 8), (OP_SATD, 4, 8), (OP_SATD, 64, 64)]
    desc = {}
    for op, w, h in cmps:
        desc[(w, h)] = tile_blocks(geo, w, h, seed=9, merange=24)
        job.add_cmp(op, w, h, *desc[(w, h)])
    trs = [(pkg.PASS_COEF, 16, 0), (pkg.PASS_LEVELS, 32, 30), (pkg.PASS_LEVELS, 8, 38), (pkg.PASS_LEVELS, 4, 22), (pkg.PASS_COEF, 4, 0), (pkg.PASS_LEVELS, 16, 51)]
    for kind, N, qp in trs:
        desc[("t", N)] = tile_blocks(geo, N, N, seed=4, merange=11)
        qc, qbits, add = quant_params(depth, N, qp) if kind == pkg.PASS_LEVELS else (None, 0, 0)
        job.add_transform(kind, N, *desc[("t", N)], qc=qc, qbits=qbits, add=add)
    frames = []
    for f in range(5):
        F = make_plane(geo, depth, 100 + f, "natural")
        R = make_plane(geo, depth, 200 + f, "natural") if f != 3 else F.copy()       # frame 3: zero residual where MV = 0
        frames.append((F, R))
    planes = [(pkg.Plane(ctx, 640, 384), pkg.Plane(ctx, 640, 384)) for _ in range(3)]
    results = {}
    slots = {}
    for f, (F, R) in enumerate(frames):
        k = f % 3
        if f >= 3:
            results[f - 3] = job.wait(slots[f - 3])
        planes[k][0].upload_padded(F); planes[k][1].upload_padded(R)
        slots[f] = job.submit(*planes[k])
    for f in range(len(frames) - 3, len(frames)):
        results[f] = job.wait(slots[f])
    nz_seen = zero_seen = False
    for f, (F, R) in enumerate(frames):
        res = results[f]
        for i, (op, w, h) in enumerate(cmps):
            want = chk.pixelcmp_batch(op, w, h, F, geo.stride, R, geo.stride, *desc[(w, h)])
            assert np.array_equal(res[i]["cost"], want), (f, op, w, h)
        for i, (kind, N, qp) in enumerate(trs):
            r = res[len(cmps) + i]
            oF, oR = desc[("t", N)]
            if kind == pkg.PASS_COEF:
                rr = orc.residual_batch(N, N, F, geo.stride, R, geo.stride, oF, oR)
                want = orc.dct_batch(N, rr, N, (np.arange(len(oF)) * N * N).astype(np.int32))
                assert np.array_equal(r["coef"], want), (f, N)
            else:
                qc, qbits, add = quant_params(depth, N, qp)
                if isinstance(chk, Reference):
                    lv, ns = chk.tu_forward_batch(N, F, geo.stride, R, geo.stride, oF, oR, qc, qbits, add)
                else:
                    lv, ns, _, _ = orc.tu_chain_batch(N, F, geo.stride, R, geo.stride, oF, oR, qc, qbits, add, 40, 1, np.zeros(geo.plane_elems, orc.pix), geo.stride, oF)
                assert np.array_equal(r["numSig"].astype(np.uint32), ns), (f, N)
                assert r["nlevels"] == int(ns.sum())
                assert np.array_equal(pkg.expand_levels(r), lv), (f, N)
                assert np.array_equal(r["levels"], lv[lv != 0])
                nz_seen |= r["nlevels"] > 0
                zero_seen |= bool((ns == 0).any())
    assert nz_seen and zero_seen
    # new motion vectors for a registered pass
    oF, oR = tile_blocks(geo, 16, 16, seed=77, merange=30)
    job.set_blocks(0, oF, oR)
    F, R = frames[1]
    planes[0][0].upload_padded(F); planes[0][1].upload_padded(R)
    res = job.wait(job.submit(*planes[0]))
    assert np.array_equal(res[0]["cost"], chk.pixelcmp_batch(OP_SATD, 16, 16, F, geo.stride, R, geo.stride, oF, oR))
    h2d, d2h = ctx.transfer_stats()
    assert h2d > 6 * 2 * geo.plane_elems * (1 if depth == 8 else 2) and d2h > 0
    ctx.check()
    job.destroy()
    for a, b in planes:
        a.destroy(); b.destroy()


def test_frame_job_misuse_is_reported():
    from gpulib import pkg
    ctx = pkg.Context(10)           # own context: the sticky status is part of the test
    geo = Geometry(352, 288)
    job = pkg.FrameJob(ctx, 352, 288, 64, slots=2)
    oF, oR = tile_blocks(geo, 16, 16, seed=1)
    bad = oR.copy(); bad[5] = geo.plane_elems          # descriptor outside the plane
    with pytest.raises(RuntimeError):
        job.add_cmp(OP_SATD, 16, 16, oF, bad)
    ctx2 = pkg.Context(10)
    job2 = pkg.FrameJob(ctx2, 352, 288, 64, slots=2)
    job2.add_cmp(OP_SATD, 16, 16, oF, oR)
    a, b = pkg.Plane(ctx2, 352, 288), pkg.Plane(ctx2, 352, 288)
    s0 = job2.submit(a, b); s1 = job2.submit(a, b)
    assert (s0, s1) == (0, 1)
    with pytest.raises(RuntimeError):
        job2.submit(a, b)                               # slot 0 was not waited for
    wrong = pkg.Plane(ctx2, 640, 384)
    ctx3 = pkg.Context(10)
    job3 = pkg.FrameJob(ctx3, 352, 288, 64, slots=1)
    job3.add_cmp(OP_SATD, 16, 16, oF, oR)
    with pytest.raises(RuntimeError):
        job3.submit(pkg.Plane(ctx3, 640, 384), pkg.Plane(ctx3, 640, 384))
    for c in (ctx, ctx2, ctx3):
        c.close()


def test_frame_job_c_example():
    """tools/frame_job_example.c: a C caller (no CUDA, no Python) drives frames through the host-buffer layer"""
    from gpulib import pkg
    exe = os.path.join(os.path.dirname(pkg.LIB_PATH), "frame_job_example")
    if not os.path.exists(exe):
        pytest.skip("example not built")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "frame job example ok" in out.stdout


@pytest.mark.parametrize("depth", [8, 10])
def test_tme_search_batch(depth):
    """the ThreadedME-shaped entry: host PU records of mixed shapes, references and candidate counts in, MEData-style results out; every search
    vs the oracle's motionEstimate and the bit / cost bookkeeping of search.cpp:392-394 redone in numpy"""
    from gpulib import context, pkg
    from frames import smooth_field
    from test_oracle_vs_ref import mv_cost_table
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(416, 240)
    rng = np.random.default_rng(1500 + depth)
    S = smooth_field(geo, depth, 7)
    F = S
    refs_np = [np.clip(np.roll(S, dy * geo.stride + dx).astype(np.int64) + rng.integers(-2, 3, S.size), 0, orc.pmax).astype(S.dtype)
               for (dx, dy) in ((3, -2), (-7, 5), (0, 0))]
    pF = pkg.Plane(ctx, 416, 240); pF.upload_padded(F)
    pR = []
    for r in refs_np:
        p = pkg.Plane(ctx, 416, 240); p.upload_padded(r); pR.append(p)
    RAD = 2048
    lam_tab = 11.3137
    tab = mv_cost_table(lam_tab, RAD)
    i = np.arange(RAD + 1, dtype=np.float32)
    bits_half = np.log(i + np.float32(1)).astype(np.float32) * np.float32(2.0 / np.log(np.float32(2.0))) + np.float32(1.718)
    bits_half[0] = np.float32(0.718)
    bits_tab = np.concatenate([bits_half[:0:-1], bits_half]).astype(np.float32)
    lam = int(np.floor(256.0 * lam_tab))
    shapes = [(8, 4), (4, 8), (8, 8), (16, 4), (16, 12), (4, 16), (12, 16), (16, 8), (8, 16), (16, 16), (32, 8), (32, 24), (8, 32), (24, 32), (32, 16),
              (16, 32), (32, 32), (64, 16), (64, 48), (16, 64), (48, 64), (64, 32), (32, 64), (64, 64)]          # g_puLookup, threadedme.h:67-92
    cw, ch = geo.coded()
    n = 160
    for method, subme, merange in ((1, 2, 16), (2, 3, 24), (3, 1, 16), (0, 5, 12)):
        pus = (pkg.TmePU * n)()
        want = []
        for k in range(n):
            w, h = shapes[int(rng.integers(0, len(shapes)))]
            x = int(rng.integers(0, cw - w + 1)); y = int(rng.integers(0, ch - h + 1))
            off = geo.origin + y * geo.stride + x
            m = int(rng.integers(4, 20))
            rr = [-min(m, x + geo.margin_x - 8), -min(m, y + geo.margin_y - 8), min(m, cw + geo.margin_x - 8 - w - x), min(m, ch + geo.margin_y - 8 - h - y)]
            nc = int(rng.integers(0, 4))
            ref = int(rng.integers(0, 3))
            mvp = rng.integers(-4 * m, 4 * m + 1, 2).astype(np.int32)
            mvc = rng.integers(-4 * m, 4 * m + 1, (nc, 2)).astype(np.int32)
            bits0 = int(rng.integers(1, 12))
            p = pus[k]
            p.w, p.h, p.ref, p.numCand, p.offF, p.offR, p.bits = w, h, ref, nc, off, off, bits0
            p.mvmin[0], p.mvmin[1], p.mvmax[0], p.mvmax[1] = rr[0], rr[1], rr[2], rr[3]
            p.mvp[0], p.mvp[1] = int(mvp[0]), int(mvp[1])
            for c in range(nc):
                p.mvc[c][0], p.mvc[c][1] = int(mvc[c][0]), int(mvc[c][1])
            mx, my, satd = orc.motion_estimate_full(subme, w, h, F, off, geo.stride, refs_np[ref], off, geo.stride, np.array(rr, np.int32), mvp, mvc, tab, RAD, method, merange)
            dx, dy = mx - int(mvp[0]), my - int(mvp[1])
            mvcost = (int(tab[RAD + dx]) + int(tab[RAD + dy])) & 0xffff
            bits = bits0 + int(np.float32(bits_tab[RAD + dx] + bits_tab[RAD + dy]) + np.float32(0.5))
            cost = ((satd - mvcost) + ((bits * lam + 128) >> 8)) & 0xffffffff
            want.append((mx, my, mvcost, bits, cost, satd))
        res = ctx.tme_search_batch(method, merange, subme, pF, pR, tab, bits_tab, RAD, lam, pus)
        got = [(r.mv[0], r.mv[1], r.mvCost, r.bits, r.cost, r.satdCost) for r in res]
        assert got == want, (method, [i for i in range(n) if got[i] != want[i]][:5])
    ctx.check()
    pF.destroy()
    for p in pR:
        p.destroy()
