"""Batched DEVICE entries (n > 1) of the remaining primitive classes against the oracle:
sse_ss / ssd_s, sa8d chroma shapes, ADS, quant / nquant / dequant, lowpass DCT, IDST, and every
interpolation variant for luma and chroma block sizes."""
import numpy as np
import pytest

from cpulibs import CHROMA_ONLY_420, CHROMA_ONLY_422, LUMA_PU, Oracle

pytestmark = pytest.mark.gpu
DEPTHS = [8, 10, 12]


def dev(a):
    import torch
    return torch.from_numpy(a).cuda()


def pix_view(a, depth):
    return a.view(np.int16) if depth > 8 else a


@pytest.mark.parametrize("depth", DEPTHS)
def test_sse_ss_ssd_s_sa8d_batches(depth):
    import torch
    from gpulib import context
    ctx = context(depth); orc = Oracle(depth)
    rng = np.random.default_rng(depth)
    stride, rows, n = 192, 200, 97
    for lo, hi in ((-orc.pmax, orc.pmax + 1), (-20000, 20000)):
        A = rng.integers(lo, hi, stride * rows).astype(np.int16)
        B = rng.integers(lo, hi, stride * rows).astype(np.int16)
        for w in (4, 8, 16, 32, 64):
            offA = (rng.integers(0, rows - w, n) * stride + rng.integers(0, stride - w, n)).astype(np.int32)
            offB = (rng.integers(0, rows - w, n) * stride + rng.integers(0, stride - w, n)).astype(np.int32)
            out = torch.zeros(n, dtype=torch.int64, device="cuda")
            ctx.sse_ss_batch(w, w, dev(A), stride, dev(B), stride, dev(offA), dev(offB), out)
            ref = np.array([orc.sse_ss(w, A, int(a), stride, B, int(b), stride) for a, b in zip(offA, offB)], np.uint64)
            got = out.cpu().numpy().astype(np.uint64)
            if depth == 8:
                got &= np.uint64(0xFFFFFFFF)
            assert np.array_equal(got, ref), w
            ctx.ssd_s_batch(w, dev(A), stride, dev(offA), out)
            ref = np.array([orc.ssd_s(w, A, int(a), stride) for a in offA], np.uint64)
            got = out.cpu().numpy().astype(np.uint64)
            if depth == 8:
                got &= np.uint64(0xFFFFFFFF)
            assert np.array_equal(got, ref), w
    # sa8d with the chroma CU shapes (4:2:2 w x 2w) and satd fallbacks
    P = rng.integers(0, orc.pmax + 1, stride * rows).astype(orc.pix)
    Q = rng.integers(0, orc.pmax + 1, stride * rows).astype(orc.pix)
    for (w, h) in ((4, 8), (8, 16), (16, 32), (32, 64), (8, 8), (16, 16), (32, 32), (64, 64), (4, 4), (12, 16), (24, 32)):
        offA = (rng.integers(0, rows - h, n) * stride + rng.integers(0, stride - w, n)).astype(np.int32)
        offB = (rng.integers(0, rows - h, n) * stride + rng.integers(0, stride - w, n)).astype(np.int32)
        out = torch.zeros(n, dtype=torch.int32, device="cuda")
        ctx.pixelcmp_batch(2, w, h, dev(pix_view(P, depth)), stride, dev(pix_view(Q, depth)), stride, dev(offA), dev(offB), out)
        ref = np.array([orc.sa8d(w, h, P, int(a), stride, Q, int(b), stride) for a, b in zip(offA, offB)], np.int32)
        assert np.array_equal(out.cpu().numpy(), ref), (w, h)
        # odd plane strides take the generic (non-vectorised) kernels
        ctx.pixelcmp_batch(1, w, h, dev(pix_view(P, depth)), stride - 1, dev(pix_view(Q, depth)), stride + 1, dev(offA), dev(offB), out)
        ref = np.array([orc.satd(w, h, P, int(a), stride - 1, Q, int(b), stride + 1) for a, b in zip(offA, offB)], np.int32)
        assert np.array_equal(out.cpu().numpy(), ref), (w, h, "odd stride")
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_ads_batch(depth):
    import torch
    from gpulib import context
    ctx = context(depth); orc = Oracle(depth)
    rng = np.random.default_rng(5)
    stride = 4128
    sums = rng.integers(0, 1 << 24, stride * 64, dtype=np.int64).astype(np.uint32)
    cost = rng.integers(0, 4000, 4096).astype(np.uint16)
    for (w, h) in LUMA_PU:
        terms = {1: 1, 2: 2, 4: 4}[orc.lib.orc_ads_terms(w, h)]
        n = 40
        enc = rng.integers(0, 1 << 24, (n, 4)).astype(np.int32)
        sumOff = (rng.integers(0, 20, n) * stride + rng.integers(0, 200, n)).astype(np.int32)
        delta = np.full(n, (h >> 1) * stride, np.int32)
        costOff = rng.integers(0, 3000, n).astype(np.int32)
        width = (rng.integers(1, 30, n) * 4).astype(np.int32)
        thresh = rng.integers(0, 1 << 25, n).astype(np.int32)
        pitch = 128
        mvs = torch.full((n * pitch,), -1, dtype=torch.int16, device="cuda")
        cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
        ctx.ads_batch(terms, (w >> 1) if terms == 4 else 0, dev(enc.ravel()), dev(sums), dev(sumOff), dev(delta), dev(cost),
                      dev(costOff), dev(width), dev(thresh), mvs, pitch, cnt)
        mv = mvs.cpu().numpy().reshape(n, pitch); c = cnt.cpu().numpy()
        for i in range(n):
            k, m = orc.ads(w, h, enc[i], sums, int(sumOff[i]), int(delta[i]), cost[costOff[i]:], int(width[i]), int(thresh[i]))
            assert c[i] == k and np.array_equal(mv[i, :k], m), (w, h, i)
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_quant_family_batches(depth):
    import torch
    from gpulib import context
    ctx = context(depth); orc = Oracle(depth)
    rng = np.random.default_rng(depth + 40)
    for numCoeff, qbits in ((16, 9), (64, 14), (256, 17), (1024, 21), (24, 12)):
        n = 53
        coef = rng.integers(-32768, 32768, n * numCoeff).astype(np.int16)
        for qc in (np.full(numCoeff, 26214, np.int32), rng.integers(-orc.pmax, orc.pmax + 1, numCoeff).astype(np.int32),
                   rng.integers(1, 1 << 17, numCoeff).astype(np.int32)):
            add = 171 << (qbits - 9)
            q = torch.zeros(n * numCoeff, dtype=torch.int16, device="cuda")
            du = torch.zeros(n * numCoeff, dtype=torch.int32, device="cuda")
            sig = torch.zeros(n, dtype=torch.int32, device="cuda")
            ctx.quant_batch(dev(coef), dev(qc), du, q, qbits, add, numCoeff, n, sig)
            hs, hq, hdu = sig.cpu().numpy(), q.cpu().numpy(), du.cpu().numpy()
            for i in range(n):                                  # every block of the batch
                r, rq, rdu = orc.quant(coef[i * numCoeff:(i + 1) * numCoeff].copy(), qc, qbits, add, numCoeff)
                assert int(hs[i]) == r, (numCoeff, i)
                assert np.array_equal(hq[i * numCoeff:(i + 1) * numCoeff], rq), (numCoeff, i)
                assert np.array_equal(hdu[i * numCoeff:(i + 1) * numCoeff], rdu), (numCoeff, i)
            ctx.quant_batch(dev(coef), dev(qc), None, q, qbits, 1 << (qbits - 1), numCoeff, n, sig)
            hs, hq = sig.cpu().numpy(), q.cpu().numpy()
            for i in range(n):
                r, rq = orc.nquant(coef[i * numCoeff:(i + 1) * numCoeff].copy(), qc, qbits, 1 << (qbits - 1), numCoeff)
                assert int(hs[i]) == r and np.array_equal(hq[i * numCoeff:(i + 1) * numCoeff], rq), (numCoeff, i)
        if numCoeff % 8 == 0:
            out = torch.zeros(n * numCoeff, dtype=torch.int16, device="cuda")
            for scale, shift in ((40, 1), (72 << 4, 5), (64 << 8, 10)):
                ctx.dequant_normal_batch(dev(coef), out, n * numCoeff, scale, shift)
                assert np.array_equal(out.cpu().numpy(), orc.dequant_normal(coef, n * numCoeff, scale, shift))
            dq = rng.integers(1, 1 << 12, numCoeff).astype(np.int32)
            for per, shift in ((0, 1), (3, 6), (12, 2), (8, 4)):
                ctx.dequant_scaling_batch(dev(coef), dev(dq), out, numCoeff, n, per, shift)
                ref = np.concatenate([orc.dequant_scaling(coef[i * numCoeff:(i + 1) * numCoeff].copy(), dq, numCoeff, per, shift) for i in range(n)])
                assert np.array_equal(out.cpu().numpy(), ref)
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_lowpass_idst_batches(depth):
    import torch
    from gpulib import context, pkg
    ctx = context(depth); orc = Oracle(depth)
    rng = np.random.default_rng(depth + 70)
    n = 61
    for src in (rng.integers(-orc.pmax, orc.pmax + 1, n * 1024 + 64).astype(np.int16), rng.integers(-32768, 32768, n * 1024 + 64).astype(np.int16)):
        for N in (8, 16, 32):
            off = (np.arange(n) * N * N).astype(np.int32)
            out = torch.zeros(n * N * N, dtype=torch.int16, device="cuda")
            ctx.dct_batch(pkg.TR_LOWPASS, N, dev(src), N, dev(off), out)
            ref = np.concatenate([orc.lowpass_dct(N, src, int(o), N) for o in off])
            assert np.array_equal(out.cpu().numpy(), ref), N
        stride = 12
        plane = torch.zeros(n * 4 * stride, dtype=torch.int16, device="cuda")
        offD = (np.arange(n) * 4 * stride + 3).astype(np.int32)
        ctx.idct_batch(pkg.TR_DST, 4, dev(src), plane, stride, dev(offD))
        ref = orc.idct_batch(4, src, np.zeros(n * 4 * stride, np.int16), stride, offD, dst4=1)
        assert np.array_equal(plane.cpu().numpy(), ref)
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_interp_batches(depth):
    import torch
    from gpulib import context
    ctx = context(depth); orc = Oracle(depth)
    rng = np.random.default_rng(depth + 90)
    ss, srows = 208, 160
    P = rng.integers(0, orc.pmax + 1, ss * srows).astype(orc.pix)
    S = rng.integers(-32768, 32768, ss * srows).astype(np.int16)
    ds = 80
    n = 9
    cases = [(8, s) for s in [(4, 4), (8, 8), (16, 16), (64, 64), (12, 16), (24, 32), (48, 64), (16, 4), (8, 32), (64, 16)]] + \
            [(4, s) for s in [(2, 4), (4, 2), (6, 8), (8, 6), (4, 4), (32, 32), (12, 32), (24, 64), (32, 48), (2, 16)]]
    for taps, (w, h) in cases:
        rmax = h + taps
        offS = ((rng.integers(4, srows - rmax - 4, n)) * ss + rng.integers(4, ss - w - 8, n)).astype(np.int32)
        offD = (np.arange(n) * (rmax * ds) + 5).astype(np.int32)
        for kind in ("hpp", "hps", "vpp", "vps", "vsp", "vss", "hvpp", "p2s"):
            if kind == "hvpp" and taps == 4:
                continue
            src = S if kind in ("vsp", "vss") else P
            pix_out = kind in ("hpp", "vpp", "vsp", "hvpp")
            nidx = 4 if taps == 8 else 8
            idx = rng.integers(1 if taps == 8 else 0, nidx, n)
            extra = rng.integers(1, 4, n) if kind == "hvpp" else rng.integers(0, 2, n) if kind == "hps" else np.zeros(n, np.int64)
            packed = (idx | (extra << 4 if kind == "hvpp" else extra << 8)).astype(np.int32)
            dt = orc.pix if pix_out else np.int16
            dst0 = np.full(n * rmax * ds + 64, 77, dt)
            d_dst = dev(pix_view(dst0.copy(), depth) if pix_out else dst0.copy())
            d_src = dev(pix_view(src, depth) if src is P else src)
            ctx.interp_batch(kind, taps, w, h, d_src, ss, dev(offS), d_dst, ds, dev(offD), dev(packed))
            ref = dst0.copy()
            for i in range(n):
                if kind == "p2s":
                    orc.p2s(w, h, src, int(offS[i]), ss, ref, int(offD[i]), ds)
                else:
                    orc.interp(kind, taps, w, h, src, int(offS[i]), ss, ref, int(offD[i]), ds, int(idx[i]), int(extra[i]))
            got = d_dst.cpu().numpy()
            got = got.view(dt) if pix_out and depth > 8 else got
            assert np.array_equal(got, ref), (taps, w, h, kind)
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_tu_chain_batch(depth):
    """fused inter-luma TU chain vs the oracle's composition, all TU sizes, QPs from near-lossless to 51 so that
    the cbf == 0, DC-only and full inverse paths all occur; 1080p-sized tiling so several scratch chunks run."""
    import torch
    from gpulib import context
    from frames import Geometry, make_plane, tile_blocks
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(832, 480)
    F = make_plane(geo, depth, 41, "natural")
    Pn = make_plane(geo, depth, 42, "natural")
    rng = np.random.default_rng(depth)
    # prediction = mixture: mostly close to fenc (small residual), some areas far
    mix = rng.integers(0, 4, geo.plane_elems)
    P = np.where(mix == 0, Pn, np.clip(F.astype(np.int64) + rng.integers(-3, 4, geo.plane_elems) * (1 << (depth - 8)), 0, orc.pmax)).astype(orc.pix)
    flat = [26214, 23302, 20560, 18396, 16384, 14564]; inv = [40, 45, 51, 57, 64, 72]
    seen = set()
    Pflat = np.clip(F.astype(np.int64) - (9 << (depth - 8)), 0, orc.pmax).astype(orc.pix)      # constant residual -> DC-only TUs
    for N, qp in ((32, 37), (16, 27), (8, 45), (4, 22), (32, 51), (8, 4), (16, 40), (16, 30), (32, 31), (4, 33)):
        offF, offMV = tile_blocks(geo, N, N, seed=4)
        offP = offF.copy() if qp != 27 else offMV         # zero-MV prediction, or displaced prediction (large residual)
        Pcur = Pflat if qp in (30, 31, 33) else P
        n = len(offF)
        per, rem = qp // 6, qp % 6
        tshift = 15 - depth - {4: 2, 8: 3, 16: 4, 32: 5}[N]
        qbits = 14 + per + tshift
        add = 85 << (qbits - 9)
        # position-dependent table (as with scaling lists) so that a wrong coefficient-position map cannot hide behind a flat one
        qc = (flat[rem] + rng.integers(-3000, 3001, N * N)).astype(np.int32) if qp not in (37, 45) else np.full(N * N, flat[rem], np.int32)
        scale, shift = inv[rem] << per, 20 - 14 - tshift
        recon0 = np.full(geo.plane_elems, 5, orc.pix)
        rq, rns, rz, rr = orc.tu_chain_batch(N, F, geo.stride, Pcur, geo.stride, offF, offP, qc, qbits, add, scale, shift, recon0, geo.stride, offF)
        d_recon = dev(pix_view(np.full(geo.plane_elems, 5, orc.pix), depth))
        q = torch.zeros(n * N * N, dtype=torch.int16, device="cuda"); ns = torch.zeros(n, dtype=torch.int32, device="cuda")
        z = torch.zeros(n, dtype=torch.int64, device="cuda"); r = torch.zeros(n, dtype=torch.int64, device="cuda")
        # path 0: tcgen05 single kernel for N = 32, fused mma.sync pair below; path 3: tcgen05 for 16 too; path 2: the mma.sync pair for every
        # size; path 1: the stage kernels (validation twin)
        for path in (0, 3, 2, 1):
            ctx.set_dct_path(path)
            d_recon = dev(pix_view(np.full(geo.plane_elems, 5, orc.pix), depth))
            q.zero_(); ns.zero_(); z.zero_(); r.zero_()
            ctx.tu_chain_batch(N, dev(pix_view(F, depth)), geo.stride, dev(pix_view(Pcur, depth)), geo.stride, dev(offF), dev(offP), dev(qc),
                               qbits, add, scale, shift, q, ns, d_recon, geo.stride, dev(offF), z, r)
            assert np.array_equal(ns.cpu().numpy().astype(np.uint32), rns), (N, qp, path)
            assert np.array_equal(q.cpu().numpy(), rq), (N, qp, path)
            assert np.array_equal(z.cpu().numpy().astype(np.uint64), rz), (N, qp, path)
            assert np.array_equal(r.cpu().numpy().astype(np.uint64), rr), (N, qp, path)
            got = d_recon.cpu().numpy()
            got = got.view(orc.pix) if depth > 8 else got
            assert np.array_equal(got, recon0), (N, qp, path)
        ctx.set_dct_path(0)
        qm = rq.reshape(n, N * N)
        seen |= {"zero"} if (rns == 0).any() else set()
        seen |= {"dc"} if ((rns == 1) & (qm[:, 0] != 0)).any() else set()
        seen |= {"full"} if (rns > 1).any() else set()
    assert seen == {"zero", "dc", "full"}, seen
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_tu_chain_intra_luma_and_chroma(depth):
    """x265b200_tu_chain_tt_batch: intra luma (DST-VII and no DC-only shortcut at 4x4; 8x8 identical to inter) on every path, and the chain on
    4:2:0 / 4:2:2 chroma planes (their own stride and margins, TU sizes 4 ... 32, the stacked 4:2:2 sub-TUs as separate TUs)"""
    import torch
    from gpulib import context, pkg
    from frames import ChromaGeometry, Geometry, make_plane, tile_blocks
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(832, 480)
    flat = [26214, 23302, 20560, 18396, 16384, 14564]; inv = [40, 45, 51, 57, 64, 72]
    rng = np.random.default_rng(100 + depth)

    def run(g, F, P, N, qp, ttype, paths, intra_add):
        offF, _ = tile_blocks(g, N, N, seed=6)
        n = len(offF)
        per, rem = qp // 6, qp % 6
        tshift = 15 - depth - {4: 2, 8: 3, 16: 4, 32: 5}[N]
        qbits = 14 + per + tshift
        add = (171 if intra_add else 85) << (qbits - 9)
        qc = (flat[rem] + rng.integers(-2000, 2001, N * N)).astype(np.int32)
        scale, shift = inv[rem] << per, 20 - 14 - tshift
        recon0 = np.full(g.plane_elems, 3, orc.pix)
        rq, rns, rz, rr = orc.tu_chain_batch(N, F, g.stride, P, g.stride, offF, offF, qc, qbits, add, scale, shift, recon0, g.stride, offF, ttype=ttype)
        q = torch.zeros(n * N * N, dtype=torch.int16, device="cuda"); ns = torch.zeros(n, dtype=torch.int32, device="cuda")
        z = torch.zeros(n, dtype=torch.int64, device="cuda"); r = torch.zeros(n, dtype=torch.int64, device="cuda")
        for path in paths:
            ctx.set_dct_path(path)
            d_recon = dev(pix_view(np.full(g.plane_elems, 3, orc.pix), depth))
            q.zero_(); ns.zero_(); z.zero_(); r.zero_()
            ctx.tu_chain_batch(N, dev(pix_view(F, depth)), g.stride, dev(pix_view(P, depth)), g.stride, dev(offF), dev(offF), dev(qc),
                               qbits, add, scale, shift, q, ns, d_recon, g.stride, dev(offF), z, r, ttype=ttype)
            assert np.array_equal(ns.cpu().numpy().astype(np.uint32), rns), (N, qp, ttype, path)
            assert np.array_equal(q.cpu().numpy(), rq), (N, qp, ttype, path)
            assert np.array_equal(z.cpu().numpy().astype(np.uint64), rz), (N, qp, ttype, path)
            assert np.array_equal(r.cpu().numpy().astype(np.uint64), rr), (N, qp, ttype, path)
            got = d_recon.cpu().numpy()
            assert np.array_equal(got.view(orc.pix) if depth > 8 else got, recon0), (N, qp, ttype, path)
        ctx.set_dct_path(0)
        return rq.reshape(n, N * N), rns

    F = make_plane(geo, depth, 51, "natural")
    P = np.clip(F.astype(np.int64) + rng.integers(-6, 7, geo.plane_elems) * (1 << (depth - 8)), 0, orc.pmax).astype(orc.pix)
    Pflat = np.clip(F.astype(np.int64) - (11 << (depth - 8)), 0, orc.pmax).astype(orc.pix)
    Pflat3 = np.clip(F.astype(np.int64) - (3 << (depth - 8)), 0, orc.pmax).astype(orc.pix)
    dc_seen = False
    for qp, Pc in ((22, P), (30, P), (38, Pflat), (44, Pflat), (12, P), (38, Pflat3), (44, Pflat3)):
        qm, rns = run(geo, F, Pc, 4, qp, pkg.TU_INTRA_LUMA, (0, 1), True)
        dc_seen |= bool(((rns == 1) & (qm[:, 0] != 0)).any())      # the case where the DST path must NOT take the DC-only shortcut
    assert dc_seen
    # the DST differs from the DCT (a wrong dispatch cannot pass both), and 8x8 intra luma is the inter chain
    a, _ = run(geo, F, P, 4, 22, pkg.TU_INTRA_LUMA, (0,), True)
    b, _ = run(geo, F, P, 4, 22, pkg.TU_INTER, (0,), True)
    assert not np.array_equal(a, b)
    run(geo, F, P, 8, 27, pkg.TU_INTRA_LUMA, (0, 1), True)
    # chroma planes: 4:2:0 and 4:2:2
    for hs, vs in ((1, 1), (1, 0)):
        cg = ChromaGeometry(geo, hs, vs)
        Fc = make_plane(cg, depth, 61 + vs, "natural")
        Pc = np.clip(Fc.astype(np.int64) + rng.integers(-5, 6, cg.plane_elems) * (1 << (depth - 8)), 0, orc.pmax).astype(orc.pix)
        for N, qp in ((4, 24), (8, 31), (16, 35), (32, 29)):
            run(cg, Fc, Pc, N, qp, pkg.TU_INTER, (0, 1) if N != 32 else (0, 2), False)
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_blockop_and_lowres_batches(depth):
    """adjacent slots: sub_ps / add_ps / pixelavg_pp / addAvg over descriptor batches (luma and odd chroma
    shapes, misaligned offsets, odd strides) and the whole-plane lowres downscale, element-wise vs the oracle"""
    import torch
    from gpulib import context
    ctx = context(depth); orc = Oracle(depth)
    rng = np.random.default_rng(100 + depth)
    stride, rows, n = 200, 180, 61
    N = stride * rows
    pixA = rng.integers(0, orc.pmax + 1, N).astype(orc.pix); pixB = rng.integers(0, orc.pmax + 1, N).astype(orc.pix)
    s16A = rng.integers(-32768, 32768, N).astype(np.int16); s16B = rng.integers(-32768, 32768, N).astype(np.int16)
    shapes = [(64, 64), (32, 8), (16, 16), (8, 8), (4, 4), (12, 16), (6, 8), (2, 4), (2, 2), (24, 32), (8, 2)]
    for op, A, B, dt in ((0, pixA, pixB, np.int16), (1, pixA, s16B, orc.pix), (2, pixA, pixB, orc.pix), (3, s16A, s16B, orc.pix)):
        for (w, h) in shapes:
            for sa, sb, sd in ((stride, stride, stride), (stride - 3, stride, stride - 1)):
                offA = (rng.integers(0, rows - h - 2, n) * sa + rng.integers(0, sa - w, n)).astype(np.int32)
                offB = (rng.integers(0, rows - h - 2, n) * sb + rng.integers(0, sb - w, n)).astype(np.int32)
                # destination blocks must not overlap: one block per 64 x 64 cell of a destination plane of its own
                per_row = sd // 64
                ND = sd * 64 * ((n + per_row - 1) // per_row) + 64
                cells = rng.permutation(n)
                offD = ((cells // per_row) * 64 * sd + (cells % per_row) * 64 + rng.integers(0, min(3, 64 - w + 1), n)).astype(np.int32)
                ref = orc.blockop_batch(op, w, h, A, sa, offA, B, sb, offB, np.full(ND, 7, dt), sd, offD)
                dA = dev(A.view(np.int16) if A.dtype == np.uint16 else A); dB = dev(B.view(np.int16) if B.dtype == np.uint16 else B)
                dD = dev(np.full(ND, 7, dt).view(np.int16) if np.dtype(dt) == np.uint16 else np.full(ND, 7, dt))
                ctx.blockop_batch(op, w, h, dA, sa, dev(offA), dB, sb, dev(offB), dD, sd, dev(offD), n)
                got = dD.cpu().numpy()
                got = got.view(np.uint16) if np.dtype(dt) == np.uint16 else got
                assert np.array_equal(got, ref), (op, w, h, sa)
    # contiguous blocks (NULL offset arrays)
    w = h = 8
    ref = np.zeros(n * 64, np.int16)
    orc.blockop_batch(0, w, h, pixA, w, (np.arange(n) * 64).astype(np.int32), pixB, w, (np.arange(n) * 64).astype(np.int32), ref, w,
                      (np.arange(n) * 64).astype(np.int32))
    dD = torch.zeros(n * 64, dtype=torch.int16, device="cuda")
    ctx.blockop_batch(0, w, h, dev(pix_view(pixA, depth)), w, None, dev(pix_view(pixB, depth)), w, None, dD, w, None, n)
    assert np.array_equal(dD.cpu().numpy(), ref)
    # lowres of a 352x288-like plane with odd dimensions
    lw, lh, ds = 87, 71, 96
    outs = [np.full(ds * lh, 3, orc.pix) for _ in range(4)]
    orc.lowres(pixA, 5, stride, *outs, ds, lw, lh)
    douts = [dev(pix_view(np.full(ds * lh, 3, orc.pix), depth)) for _ in range(4)]
    src = dev(pix_view(pixA, depth))
    ctx.lowres_batch(src[5:], stride, *douts, ds, lw, lh)
    for a, b in zip(outs, douts):
        g = b.cpu().numpy()
        assert np.array_equal(g.view(np.uint16) if depth > 8 else g, a)
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_subpel_cmp_batch(depth):
    """fused interpolation + SAD / SATD (subpelCompare) vs the oracle: every luma PU shape, all 16 fractions, K candidates
    per block sharing one fenc block, candidate origins at arbitrary (misaligned) integer positions"""
    import torch
    from gpulib import context
    from frames import Geometry, make_plane
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(416, 240)
    F = make_plane(geo, depth, 51, "natural"); R = make_plane(geo, depth, 52, "natural")
    Ru = make_plane(geo, depth, 53, "uniform")
    rng = np.random.default_rng(200 + depth)
    cw, ch = geo.coded()
    dF = dev(pix_view(F, depth))
    for (w, h) in LUMA_PU:
        n, K = 37, 4
        x = rng.integers(0, cw - w, n); y = rng.integers(0, ch - h, n)
        offF = (geo.origin + y * geo.stride + x).astype(np.int32)
        mvx = rng.integers(-40, 41, n * K); mvy = rng.integers(-40, 41, n * K)
        offR = (geo.origin + (np.repeat(y, K) + mvy) * geo.stride + np.repeat(x, K) + mvx).astype(np.int32)
        frac = (rng.integers(0, 4, n * K) | (rng.integers(0, 4, n * K) << 4)).astype(np.int32)
        frac[:8] = [0, 1, 2, 3, 0x10, 0x20, 0x30, 0x33]
        for op, ref_plane in ((0, R), (1, R), (1, Ru)):
            want = orc.subpel_cmp_batch(op, w, h, F, geo.stride, ref_plane, geo.stride, offF, offR, frac, K)
            cost = torch.zeros(n * K, dtype=torch.int32, device="cuda")
            ctx.subpel_cmp_batch(op, w, h, dF, geo.stride, dev(pix_view(ref_plane, depth)), geo.stride, dev(offF), dev(offR), dev(frac), K, cost)
            assert np.array_equal(cost.cpu().numpy(), want), (w, h, op)
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_me_integral_batch(depth):
    """twelve SEA integral planes of two padded pictures in one launch vs the oracle's row loop (defined region + zero row),
    and the planes feeding `ads` exactly like a reference search row would"""
    import torch
    from gpulib import context
    ctx = context(depth); orc = Oracle(depth)
    rng = np.random.default_rng(300 + depth)
    stride, rows, nf = 224, 150, 2
    pix = rng.integers(0, orc.pmax + 1, nf * stride * rows).astype(orc.pix)
    pitch = stride * rows
    sums = torch.full((nf * 12 * pitch,), 0x5a5a5a5a, dtype=torch.int32, device="cuda")
    ctx.me_integral_batch(dev(pix_view(pix, depth)), stride, rows, nf, sums, pitch)
    got = sums.cpu().numpy().view(np.uint32)
    W = [32, 32, 32, 24, 16, 16, 16, 12, 8, 8, 4, 4]; H = [32, 24, 8, 32, 16, 12, 4, 16, 32, 8, 16, 4]
    for f in range(nf):
        ref = np.zeros(12 * pitch, np.uint32)
        orc.me_integral(pix[f * pitch:(f + 1) * pitch], stride, rows, ref, pitch)
        for k in range(12):
            G = got[(f * 12 + k) * pitch:(f * 12 + k + 1) * pitch].reshape(rows, stride)
            R = ref[k * pitch:(k + 1) * pitch].reshape(rows, stride)
            assert not G[0].any()
            assert np.array_equal(G[1:rows - H[k], :stride - W[k]], R[1:rows - H[k], :stride - W[k]]), (f, k)
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_weight_and_weight_cost_batches(depth):
    """weight_pp / weight_sp planes and the fused K-candidate weighted-prediction cost vs the oracle"""
    import torch
    from gpulib import context
    ctx = context(depth); orc = Oracle(depth)
    rng = np.random.default_rng(400 + depth)
    corr = 14 - depth
    stride, rows = 480, 300
    pix = rng.integers(0, orc.pmax + 1, stride * rows).astype(orc.pix)
    s16 = rng.integers(-8192, 8192, stride * rows).astype(np.int16)
    for trial in range(4):
        shift = int(rng.integers(1, 7)); w0 = int(rng.integers(1, 127)); offset = int(rng.integers(-128, 128)) << (depth - 8)
        rnd = 1 << (shift - 1)
        want = np.full(stride * rows, 3, orc.pix)
        orc.weight_pp(pix, 0, want, 0, stride, 464, 290, w0, rnd << corr, shift + corr, offset)
        got = dev(pix_view(np.full(stride * rows, 3, orc.pix), depth))
        ctx.weight_batch(0, dev(pix_view(pix, depth)), stride, got, stride, 464, 290, w0, rnd << corr, shift + corr, offset)
        g = got.cpu().numpy()
        assert np.array_equal(g.view(np.uint16) if depth > 8 else g, want)
        want = np.full(stride * rows, 3, orc.pix)
        orc.weight_sp(s16, 0, want, 0, stride, stride, 301, 177, w0, rnd << corr, shift + corr, offset)
        got = dev(pix_view(np.full(stride * rows, 3, orc.pix), depth))
        ctx.weight_batch(1, dev(s16), stride, got, stride, 301, 177, w0, rnd << corr, shift + corr, offset)
        g = got.cpu().numpy()
        assert np.array_equal(g.view(np.uint16) if depth > 8 else g, want)
    fenc = rng.integers(0, orc.pmax + 1, stride * rows).astype(orc.pix)
    ref = np.clip(fenc.astype(np.int64) * 3 // 4 + rng.integers(-20, 21, stride * rows), 0, orc.pmax).astype(orc.pix)
    W, H = 444, 270
    intra = rng.integers(0, 3000 << (depth - 8), ((W + 7) // 8) * ((H + 7) // 8)).astype(np.int32)
    weights = np.array([0, 0, -1, 0] + [64, 32 << corr, 6 + corr, 0] + [85, 32 << corr, 6 + corr, 3 << (depth - 8)]
                       + [43, 16 << corr, 5 + corr, -(2 << (depth - 8))] + [100, 1 << corr, 1 + corr, 0], np.int32)
    K = len(weights) // 4
    for ic in (intra, None):
        want = orc.weight_cost(fenc, 0, ref, 0, stride, W, H, ic, weights)
        cost = torch.zeros(K, dtype=torch.int32, device="cuda")
        ctx.weight_cost_batch(dev(pix_view(fenc, depth)), dev(pix_view(ref, depth)), stride, W, H, dev(ic) if ic is not None else None, dev(weights), K, cost)
        assert np.array_equal(cost.cpu().numpy().view(np.uint32), want)
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_blockcopy_batch(depth):
    """copy family over descriptor batches vs the oracle (odd chroma shapes, misaligned offsets, NULL offset arrays)"""
    import torch
    from gpulib import context
    ctx = context(depth); orc = Oracle(depth)
    rng = np.random.default_rng(500 + depth)
    stride, rows, n = 200, 160, 45
    N = stride * rows
    pix = rng.integers(0, orc.pmax + 1, N).astype(orc.pix)
    s16 = rng.integers(-32768, 32768, N).astype(np.int16)
    s16pix = rng.integers(0, orc.pmax + 1, N).astype(np.int16)
    for kind, src, dt, param in ((0, pix, orc.pix, 0), (1, s16, np.int16, 0), (2, s16pix, orc.pix, 0), (3, pix, np.int16, 0), (4, None, np.int16, -77),
                                 (5, s16, np.int16, 3), (6, s16, np.int16, 5)):
        for (w, h) in ((64, 64), (32, 8), (12, 16), (6, 8), (2, 4), (4, 4)):
            offS = (rng.integers(0, rows - h, n) * stride + rng.integers(0, stride - w, n)).astype(np.int32)
            per_row = stride // 64
            ND = stride * 64 * ((n + per_row - 1) // per_row) + 64
            cells = rng.permutation(n)
            offD = ((cells // per_row) * 64 * stride + (cells % per_row) * 64 + rng.integers(0, min(3, 64 - w + 1), n)).astype(np.int32)
            want = orc.blockcopy_batch(kind, w, h, src, stride, offS, np.full(ND, 7, dt), stride, offD, param)
            dS = dev(pix_view(src, depth) if src is not None and src.dtype == np.uint16 else src) if src is not None else None
            dD = dev(np.full(ND, 7, dt).view(np.int16) if np.dtype(dt) == np.uint16 else np.full(ND, 7, dt))
            ctx.blockcopy_batch(kind, w, h, dS, stride, dev(offS), dD, stride, dev(offD), n, param)
            got = dD.cpu().numpy()
            got = got.view(np.uint16) if np.dtype(dt) == np.uint16 else got
            assert np.array_equal(got, want), (kind, w, h)
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_block_scalar_batches(depth):
    """var / psy_cost_pp / count_nonzero / copy_cnt / denoiseDct batches vs the oracle"""
    import torch
    from gpulib import context
    ctx = context(depth); orc = Oracle(depth)
    rng = np.random.default_rng(600 + depth)
    stride, rows, n = 200, 200, 53
    N = stride * rows
    A = rng.integers(0, orc.pmax + 1, N).astype(orc.pix); B = rng.integers(0, orc.pmax + 1, N).astype(orc.pix)
    A[:N // 2] = orc.pmax                                   # saturated half: the uint32 sum of squares wraps at 12 bit, 64 x 64
    dA, dB = dev(pix_view(A, depth)), dev(pix_view(B, depth))
    resi = (rng.integers(-300, 300, N) * (rng.integers(0, 3, N) == 0)).astype(np.int16)
    for size in (4, 8, 16, 32, 64):
        offA = (rng.integers(0, rows - size, n) * stride + rng.integers(0, stride - size, n)).astype(np.int32)
        offB = (rng.integers(0, rows - size, n) * stride + rng.integers(0, stride - size, n)).astype(np.int32)
        out = torch.zeros(n, dtype=torch.int64, device="cuda")
        ctx.var_batch(size, dA, stride, dev(offA), n, out)
        assert np.array_equal(out.cpu().numpy().view(np.uint64), np.array([orc.var(size, A, int(a), stride) for a in offA], np.uint64)), size
        cost = torch.zeros(n, dtype=torch.int32, device="cuda")
        ctx.psy_cost_batch(size, dA, stride, dev(offA), dB, stride, dev(offB), n, cost)
        assert np.array_equal(cost.cpu().numpy(), np.array([orc.psy_cost_pp(size, A, int(a), stride, B, int(b), stride) for a, b in zip(offA, offB)], np.int32)), size
        if size <= 32:
            coeff = torch.zeros(n * size * size, dtype=torch.int16, device="cuda"); cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
            ctx.count_nonzero_batch(size, dev(resi), stride, dev(offA), n, coeff, cnt)
            want_c = np.zeros(n * size * size, np.int16)
            want_n = np.array([orc.copy_cnt(size, want_c[i * size * size:(i + 1) * size * size], resi, int(a), stride) for i, a in enumerate(offA)], np.uint32)
            assert np.array_equal(cnt.cpu().numpy().view(np.uint32), want_n) and np.array_equal(coeff.cpu().numpy(), want_c), size
            num = size * size
            d0 = rng.integers(-32768, 32768, n * num).astype(np.int16)
            rs0 = rng.integers(0, 1 << 20, num).astype(np.uint32); offs = rng.integers(0, 65535, num).astype(np.uint16)
            wd = d0.copy(); wr = rs0.copy()
            for i in range(n):
                orc.denoise_dct(wd[i * num:(i + 1) * num], wr, offs, num)
            gd = dev(d0.copy()); gr = dev(rs0.view(np.int32).copy())
            ctx.denoise_dct_batch(gd, gr, dev(offs.view(np.int16)), num, n)
            assert np.array_equal(gd.cpu().numpy(), wd) and np.array_equal(gr.cpu().numpy().view(np.uint32), wr), size
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_me_full_batch(depth):
    """exhaustive integer search of a batch of PUs vs the oracle's raster loop: windows larger and smaller than a candidate
    tile, single-column / single-row / empty windows, lambda-scaled and random (wrapping) cost tables, and flat pictures
    where every candidate ties and the tie-break order decides the vector"""
    import torch
    from gpulib import context
    from frames import Geometry, make_plane
    from test_oracle_vs_ref import mv_cost_table
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(416, 240)
    F = make_plane(geo, depth, 61, "natural"); R = make_plane(geo, depth, 62, "natural")
    flat = np.full_like(F, orc.pmax // 3)
    rng = np.random.default_rng(400 + depth)
    cw, ch = geo.coded()
    RAD = 2048
    tabs = [mv_cost_table(12.6992, RAD), rng.integers(0, 65536, 2 * RAD + 1).astype(np.uint16), np.zeros(2 * RAD + 1, np.uint16)]
    dtabs = [dev(t.view(np.int16)) for t in tabs]
    for (w, h) in LUMA_PU:
        n = 14
        x = rng.integers(0, cw - w + 1, n); y = rng.integers(0, ch - h + 1, n)
        offF = (geo.origin + y * geo.stride + x).astype(np.int32)
        x2 = rng.integers(0, cw - w + 1, n); y2 = rng.integers(0, ch - h + 1, n)        # co-located block of the reference picture
        offR = (geo.origin + y2 * geo.stride + x2).astype(np.int32)
        # windows clipped to the padded picture like setSearchRange: up to +-70 columns, +-45 rows
        minx = -np.minimum(rng.integers(0, 71, n), x2 + geo.margin_x - 8); maxx = np.minimum(rng.integers(0, 71, n), cw + geo.margin_x - 8 - w - x2)
        miny = -np.minimum(rng.integers(0, 46, n), y2 + geo.margin_y - 8); maxy = np.minimum(rng.integers(0, 46, n), ch + geo.margin_y - 8 - h - y2)
        maxx[0] = minx[0]; maxy[1] = miny[1]                  # one column, one row
        maxx[2] = minx[2] - 1                                 # empty: PU stays untouched
        minx[3], maxx[3], miny[3], maxy[3] = 0, 0, 0, 0       # a single candidate
        rngs = np.stack([minx, miny, maxx, maxy], 1).astype(np.int32).copy()
        mvp = rng.integers(-200, 201, (n, 2)).astype(np.int32)
        bmv0 = np.stack([rng.integers(minx, np.maximum(maxx, minx) + 1), rng.integers(miny, maxy + 1)], 1).astype(np.int32)
        for ti, tab in enumerate(tabs):
            merange = (70, 16, 0)[ti]                         # staging hint: whole window, super-tiles, minimum
            for fe, rf in ((F, R), (flat, flat)):
                bc0 = np.full(n, 0x7fffffff, np.int32)
                bc0[4] = 0                                    # nothing can beat the starting point
                if fe is flat: bc0[5::3] = int(tab[RAD - int(mvp[5, 0])]) if ti == 2 else 37       # ties / low starting costs
                want_mv, want_c = bmv0.copy(), bc0.copy()
                orc.me_full_batch(w, h, fe, geo.stride, rf, geo.stride, offF, offR, rngs, mvp, tab, RAD, want_mv, want_c)
                gmv, gc = dev(bmv0.copy()), dev(bc0.copy())
                ctx.me_full_batch(w, h, merange, dev(pix_view(fe, depth)), geo.stride, dev(pix_view(rf, depth)), geo.stride, dev(offF), dev(offR),
                                  dev(rngs), dev(mvp), dtabs[ti].data_ptr() + 2 * RAD, gmv, gc)
                assert np.array_equal(gc.cpu().numpy(), want_c), (w, h, ti)
                assert np.array_equal(gmv.cpu().numpy(), want_mv), (w, h, ti)
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_motion_estimate_batch(depth):
    """whole motionEstimate (predictor candidates, full / hexagon / diamond / star / uneven multi-hexagon search, sub-pel refinement, zero-vector chance) for a batch of
    PUs vs the oracle's restatement -- itself pinned to the reference's MotionEstimate::motionEstimate by the CPU suite --
    for every SubpelWorkload level, with and without neighbour candidates, including PUs that leave early on zero residual"""
    import torch
    from gpulib import context
    from frames import Geometry, make_plane
    from test_oracle_vs_ref import mv_cost_table
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(416, 240)
    F = make_plane(geo, depth, 81, "natural"); R = make_plane(geo, depth, 82, "natural")
    R2 = np.roll(F, 2 * geo.stride + 3)                   # exact copy displaced by (+3, +2)
    rng = np.random.default_rng(500 + depth)
    from frames import smooth_field
    S = smooth_field(geo, depth, 83)                      # smooth texture: hexagon / diamond walks take many steps
    S2 = np.clip(np.roll(S, -6 * geo.stride + 9) + rng.integers(0, 3, S.size).astype(S.dtype), 0, orc.pmax).astype(S.dtype)
    S3 = np.roll(S, 11 * geo.stride - 14)                 # far match (-14, +11): star rings of 8 / 16, the raster pass
    walked = 0
    cw, ch = geo.coded()
    RAD = 2048
    tab = mv_cost_table(8.9797, RAD); dtab = dev(tab.view(np.int16))
    dF = dev(pix_view(F, depth))
    shapes = [(8, 8), (16, 16), (32, 32), (64, 64), (16, 8), (8, 16), (32, 24), (12, 16), (64, 32), (16, 4), (8, 4), (24, 32), (4, 8), (4, 4)]
    for si, (w, h) in enumerate(shapes):
        subme = si % 8
        n, nc = 40, (0, 2, 5)[si % 3]
        x = rng.integers(0, cw - w + 1, n); y = rng.integers(0, ch - h + 1, n)
        off = (geo.origin + y * geo.stride + x).astype(np.int32)
        m = int(rng.integers(3, 20))
        minx = -np.minimum(m, x + geo.margin_x - 8); maxx = np.minimum(m, cw + geo.margin_x - 8 - w - x)
        miny = -np.minimum(m, y + geo.margin_y - 8); maxy = np.minimum(m, ch + geo.margin_y - 8 - h - y)
        rngs = np.stack([minx, miny, maxx, maxy], 1).astype(np.int32).copy()
        qmvp = rng.integers(-4 * m - 6, 4 * m + 7, (n, 2)).astype(np.int32)
        qmvp[::7] = 0; qmvp[1::9] = (12, 8)               # (12, 8) q-pel is the displaced copy: zero residual on R2
        mvc = rng.integers(-4 * m - 6, 4 * m + 7, (n, max(nc, 1), 2)).astype(np.int32)
        mvc[::4, 0] = (12, 8)
        for fen, ref_plane, method in ((F, R, 5), (F, R2, 5), (F, R2, 1), (S, S2, 1), (S, S2, 0), (F, R, 0), (S, S2, 5),
                                       (S, S2, 3), (S, S3, 3), (F, R, 3), (F, R2, 3), (S, S2, 2), (S, S3, 2), (F, R, 2), (F, R2, 2)):
            merange = m if method == 5 else int(rng.integers(1, 40))
            if ref_plane is S3: merange = int(rng.integers(16, 64))
            want_mv = np.zeros((n, 2), np.int32); want_c = np.zeros(n, np.int32)
            for i in range(n):
                a = orc.motion_estimate_full(subme, w, h, fen, int(off[i]), geo.stride, ref_plane, int(off[i]), geo.stride, rngs[i], qmvp[i],
                                             mvc[i, :nc], tab, RAD, method, merange)
                want_mv[i] = a[:2]; want_c[i] = a[2]
            gmv = torch.full((n, 2), -7777, dtype=torch.int32, device="cuda"); gc = torch.full((n,), -7777, dtype=torch.int32, device="cuda")
            ctx.motion_estimate_batch(method, w, h, merange, subme, dev(pix_view(fen, depth)), geo.stride, dev(pix_view(ref_plane, depth)), geo.stride,
                                      dev(off), dev(off), dev(rngs), dev(qmvp), nc, dev(np.ascontiguousarray(mvc[:, :nc])) if nc else None,
                                      dtab.data_ptr() + 2 * RAD, gmv, gc)
            assert np.array_equal(gc.cpu().numpy(), want_c), (w, h, subme, method)
            assert np.array_equal(gmv.cpu().numpy(), want_mv), (w, h, subme, method)
            if method != 5 and fen is S:
                walked += int((np.abs(want_mv - np.clip(qmvp, 4 * rngs[:, :2], 4 * rngs[:, 2:])).max(1) >= 16).sum())
    assert walked > 100                                   # pattern searches ended four or more pels from their start
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_motion_estimate_sea_batch(depth):
    """X265_SEA through the whole chain: integral planes from x265b200_me_integral_batch, ads pre-filter and survivor SADs in
    me_sea_kernel, vs the oracle (pinned to the reference's motionEstimate with its own integral planes by the CPU suite)"""
    import torch
    from gpulib import context
    from frames import Geometry, make_plane, smooth_field
    from test_oracle_vs_ref import mv_cost_table
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(416, 240)
    rng = np.random.default_rng(900 + depth)
    S = smooth_field(geo, depth, 141, box=9)
    N = make_plane(geo, depth, 142, "natural")
    cw, ch = geo.coded()
    pitch = geo.plane_elems
    RAD = 4096
    tab = mv_cost_table(7.1, RAD); dtab = dev(tab.view(np.int16))
    shapes = [(16, 16), (8, 8), (32, 32), (64, 64), (16, 8), (8, 16), (32, 16), (16, 32), (64, 32), (32, 64), (32, 24), (24, 32),
              (64, 48), (48, 64), (64, 16), (16, 64), (16, 12), (12, 16), (16, 4), (4, 16), (4, 4)]
    moved = 0
    for si, (w, h) in enumerate(shapes):
        dx, dy = int(rng.integers(-10, 11)), int(rng.integers(-8, 9))
        if si % 3 == 2:
            F, R = N, make_plane(geo, depth, 143 + si, "natural")
        else:
            F = S
            R = np.clip(np.roll(S, dy * geo.stride + dx).astype(np.int64) + rng.integers(-3, 4, S.size), 0, orc.pmax).astype(S.dtype)
        dR = dev(pix_view(R, depth))
        dsums = torch.zeros(12 * pitch, dtype=torch.int32, device="cuda")
        ctx.me_integral_batch(dR, geo.stride, geo.rows, 1, dsums, pitch)
        sums = dsums.cpu().numpy().view(np.uint32)
        n, nc = 24, (0, 2, 1)[si % 3]
        subme = si % 4
        x = rng.integers(0, cw - w + 1, n); y = rng.integers(0, ch - h + 1, n)
        off = (geo.origin + y * geo.stride + x).astype(np.int32)
        m = int(rng.integers(4, 24))
        minx = -np.minimum(m, x + geo.margin_x - 12); maxx = np.minimum(m, cw + geo.margin_x - 12 - w - x)
        miny = -np.minimum(m, y + geo.margin_y - 12); maxy = np.minimum(m, ch + geo.margin_y - 12 - h - y)
        rngs = np.stack([minx, miny, maxx, maxy], 1).astype(np.int32).copy()
        qmvp = rng.integers(-4 * m, 4 * m + 1, (n, 2)).astype(np.int32)
        qmvp[::5] = 0
        mvc = rng.integers(-4 * m, 4 * m + 1, (n, max(nc, 1), 2)).astype(np.int32)
        merange = int(rng.integers(2, 24))
        want = np.array([orc.motion_estimate_sea(merange, subme, w, h, F, int(off[i]), geo.stride, R, int(off[i]), geo.stride, sums, pitch, rngs[i], qmvp[i],
                                                 mvc[i, :nc], tab, RAD) for i in range(n)], np.int32)
        gmv = torch.full((n, 2), -7777, dtype=torch.int32, device="cuda"); gc = torch.full((n,), -7777, dtype=torch.int32, device="cuda")
        ctx.motion_estimate_sea_batch(w, h, merange, subme, dev(pix_view(F, depth)), geo.stride, dR, geo.stride, dev(off), dev(off), dev(rngs), dev(qmvp),
                                      nc, dev(np.ascontiguousarray(mvc[:, :nc])) if nc else None, dtab.data_ptr() + 2 * RAD, dsums, pitch, gmv, gc)
        assert np.array_equal(gc.cpu().numpy(), want[:, 2]), (w, h, subme)
        assert np.array_equal(gmv.cpu().numpy(), want[:, :2]), (w, h, subme)
        moved += int((np.abs(want[:, :2] - np.clip(qmvp, 4 * rngs[:, :2], 4 * rngs[:, 2:])).max(1) >= 8).sum())
    assert moved > 60
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_lowres_motion_estimate_batch(depth):
    """the lookahead's motionEstimate on a lowres reference (four half-pel planes, quarter-pel cost on the average of two)
    for batches of 8x8 blocks vs the oracle (pinned to the reference with ref->isLowres by the CPU suite): every search
    method built, several subme levels, fenc on full- and half-pel phases, windows clipped at the picture edge"""
    import torch
    from gpulib import context
    from frames import Geometry
    from test_oracle_vs_ref import lowres_planes, mv_cost_table
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(416, 240)
    rng = np.random.default_rng(600 + depth)
    P = lowres_planes(geo, depth, 97)
    pitch = geo.plane_elems
    cw, ch = geo.coded()
    RAD = 2048
    tab = mv_cost_table(6.3496, RAD); dtab = dev(tab.view(np.int16))
    dP = dev(pix_view(P, depth))
    sub = 0
    for case in range(12):
        method = (1, 0, 3, 5)[case % 4]
        subme = (1, 2, 0, 5, 7, 1)[case % 6]
        dx, dy = int(rng.integers(-12, 13)), int(rng.integers(-10, 11))
        src = P[(case % 4) * pitch:(case % 4 + 1) * pitch]
        F = np.clip(np.roll(src, dy * geo.stride + dx).astype(np.int64) + rng.integers(-2, 3, pitch), 0, orc.pmax).astype(P.dtype)
        n = 150
        x = rng.integers(0, cw - 8 + 1, n); y = rng.integers(0, ch - 8 + 1, n)
        x[:10] = 0; y[10:20] = ch - 8                      # blocks on the picture edge
        off = (geo.origin + y * geo.stride + x).astype(np.int32)
        m = int(rng.integers(4, 26))
        minx = -np.minimum(m, x + geo.margin_x - 8); maxx = np.minimum(m, cw + geo.margin_x - 16 - x)
        miny = -np.minimum(m, y + geo.margin_y - 8); maxy = np.minimum(m, ch + geo.margin_y - 16 - y)
        rngs = np.stack([minx, miny, maxx, maxy], 1).astype(np.int32).copy()
        qmvp = rng.integers(-4 * m - 6, 4 * m + 7, (n, 2)).astype(np.int32)
        qmvp[::6] = 0
        merange = m if method == 5 else int(rng.integers(1, 33))
        want_mv = np.zeros((n, 2), np.int32); want_c = np.zeros(n, np.int32)
        for i in range(n):
            a = orc.lowres_motion_estimate(method, merange, subme, 8, 8, F, int(off[i]), geo.stride, P, int(off[i]), geo.stride, pitch, rngs[i],
                                           qmvp[i], tab, RAD)
            want_mv[i] = a[:2]; want_c[i] = a[2]
        gmv = torch.full((n, 2), -7777, dtype=torch.int32, device="cuda"); gc = torch.full((n,), -7777, dtype=torch.int32, device="cuda")
        ctx.lowres_motion_estimate_batch(method, 8, 8, merange, subme, dev(pix_view(F, depth)), geo.stride, dP, geo.stride, pitch, dev(off), dev(off),
                                         dev(rngs), dev(qmvp), dtab.data_ptr() + 2 * RAD, gmv, gc)
        assert np.array_equal(gc.cpu().numpy(), want_c), (case, method, subme)
        assert np.array_equal(gmv.cpu().numpy(), want_mv), (case, method, subme)
        sub += int(((want_mv[:, 0] | want_mv[:, 1]) & 1).sum())
    assert sub > 100                                      # quarter-pel winners: the two-plane average decided them
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_subpel_cmp_chroma_batch(depth):
    """chroma term of subpelCompare: fused 4-tap interpolation + SATD vs the oracle for every chroma block shape with a SATD
    slot, all 64 fractions, overwrite and accumulate modes, skipped candidates (negative frac)"""
    import torch
    from gpulib import context
    from frames import Geometry, make_plane
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(416, 240)
    F = make_plane(geo, depth, 91, "natural"); R = make_plane(geo, depth, 92, "natural"); Ru = make_plane(geo, depth, 93, "uniform")
    rng = np.random.default_rng(700 + depth)
    cw, ch = geo.coded()
    dF = dev(pix_view(F, depth))
    shapes = [(4, 4), (8, 8), (16, 16), (32, 32), (8, 4), (4, 8), (16, 8), (8, 16), (32, 16), (16, 32), (16, 12), (12, 16), (16, 4), (4, 16),
              (32, 24), (24, 32), (32, 8), (8, 32)]
    for (w, h) in shapes:
        n, K = 23, 5
        x = rng.integers(0, cw - w, n); y = rng.integers(0, ch - h, n)
        offF = (geo.origin + y * geo.stride + x).astype(np.int32)
        mvx = rng.integers(-30, 31, n * K); mvy = rng.integers(-30, 31, n * K)
        offR = (geo.origin + (np.repeat(y, K) + mvy) * geo.stride + np.repeat(x, K) + mvx).astype(np.int32)
        frac = (rng.integers(0, 8, n * K) | (rng.integers(0, 8, n * K) << 4)).astype(np.int32)
        frac[:10] = [0, 1, 7, 0x10, 0x70, 0x77, 0x34, 4, 0x40, 0x44]
        skip = rng.random(n * K) < 0.15
        for ref_plane in (R, Ru):
            want = np.array([orc.subpel_cmp_chroma(w, h, F, int(offF[i // K]), geo.stride, ref_plane, int(offR[i]), geo.stride,
                                                   int(frac[i] & 7), int(frac[i] >> 4)) for i in range(n * K)], np.int32)
            dR = dev(pix_view(ref_plane, depth))
            cost = torch.full((n * K,), -5, dtype=torch.int32, device="cuda")
            ctx.subpel_cmp_chroma_batch(w, h, dF, geo.stride, dR, geo.stride, dev(offF), dev(offR), dev(frac), K, cost)
            assert np.array_equal(cost.cpu().numpy(), want), (w, h)
            base = rng.integers(0, 1000, n * K).astype(np.int32)
            cost = dev(base.copy())
            ctx.subpel_cmp_chroma_batch(w, h, dF, geo.stride, dR, geo.stride, dev(offF), dev(offR), dev(np.where(skip, -1, frac).astype(np.int32)), K,
                                        cost, accumulate=1)
            assert np.array_equal(cost.cpu().numpy(), base + np.where(skip, 0, want)), (w, h, "accumulate")
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_motion_estimate_chroma_batch(depth):
    """motionEstimate with the chroma SATD term (subme > 2, 4:2:0) for batches of PUs vs the oracle (pinned to the reference's
    encoder-style setSourcePU path by the CPU suite), incl. shapes / subme levels where the term is off"""
    import torch
    from gpulib import context
    from frames import Geometry
    from test_oracle_vs_ref import mv_cost_table, yuv420_planes
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(416, 240)
    rng = np.random.default_rng(800 + depth)
    cgeo, FY, FCb, FCr = yuv420_planes(geo, depth, 111)
    refs = [yuv420_planes(geo, depth, 111, shift=(6, -4))[1:], yuv420_planes(geo, depth, 211)[1:]]
    refs = [tuple(np.clip(p.astype(np.int64) + rng.integers(-3, 4, p.size), 0, orc.pmax).astype(p.dtype) for p in r3) for r3 in refs]
    cw, ch = geo.coded()
    RAD = 2048
    tab = mv_cost_table(10.0794, RAD); dtab = dev(tab.view(np.int16))
    dFY, dFCb, dFCr = (dev(pix_view(p, depth)) for p in (FY, FCb, FCr))
    cases = [((16, 16), 3, 1), ((8, 8), 5, 3), ((32, 32), 4, 1), ((64, 64), 7, 5), ((16, 8), 3, 0), ((32, 24), 6, 3), ((16, 12), 3, 1),
             ((12, 16), 4, 1), ((8, 16), 2, 1), ((24, 32), 5, 3), ((64, 32), 3, 1), ((16, 4), 3, 1)]
    decided = 0
    for ci, ((w, h), subme, method) in enumerate(cases):
        RY, RCb, RCr = refs[ci % 3 == 2]
        n, nc = 30, ci % 3
        x = rng.integers(0, (cw - w) // 2 + 1, n) * 2; y = rng.integers(0, (ch - h) // 2 + 1, n) * 2
        offY = (geo.origin + y * geo.stride + x).astype(np.int32)
        offC = (cgeo.origin + (y // 2) * cgeo.stride + x // 2).astype(np.int32)
        m = int(rng.integers(3, 14))
        minx = -np.minimum(m, x + geo.margin_x - 16); maxx = np.minimum(m, cw + geo.margin_x - 16 - w - x)
        miny = -np.minimum(m, y + geo.margin_y - 16); maxy = np.minimum(m, ch + geo.margin_y - 16 - h - y)
        rngs = np.stack([minx, miny, maxx, maxy], 1).astype(np.int32).copy()
        qmvp = rng.integers(-4 * m - 6, 4 * m + 7, (n, 2)).astype(np.int32); qmvp[::7] = 0
        mvc = rng.integers(-4 * m - 6, 4 * m + 7, (n, max(nc, 1), 2)).astype(np.int32)
        merange = m if method == 5 else int(rng.integers(1, 33))
        want = np.array([orc.motion_estimate_chroma(method, merange, subme, w, h, FY, int(offY[i]), geo.stride, RY, int(offY[i]), geo.stride,
                                                    FCb, FCr, int(offC[i]), cgeo.stride, RCb, RCr, int(offC[i]), cgeo.stride, 1, 1,
                                                    rngs[i], qmvp[i], mvc[i, :nc], tab, RAD) for i in range(n)], np.int32)
        luma = np.array([orc.motion_estimate_full(subme, w, h, FY, int(offY[i]), geo.stride, RY, int(offY[i]), geo.stride, rngs[i], qmvp[i],
                                                  mvc[i, :nc], tab, RAD, method, merange) for i in range(n)], np.int32)
        gmv = torch.full((n, 2), -7777, dtype=torch.int32, device="cuda"); gc = torch.full((n,), -7777, dtype=torch.int32, device="cuda")
        ctx.motion_estimate_chroma_batch(method, w, h, merange, subme, dFY, geo.stride, dev(pix_view(RY, depth)), geo.stride, dev(offY), dev(offY),
                                         dFCb, dFCr, cgeo.stride, dev(pix_view(RCb, depth)), dev(pix_view(RCr, depth)), cgeo.stride,
                                         dev(offC), dev(offC), 1, 1, dev(rngs), dev(qmvp), nc,
                                         dev(np.ascontiguousarray(mvc[:, :nc])) if nc else None, dtab.data_ptr() + 2 * RAD, gmv, gc)
        assert np.array_equal(gc.cpu().numpy(), want[:, 2]), (w, h, subme, method)
        assert np.array_equal(gmv.cpu().numpy(), want[:, :2]), (w, h, subme, method)
        on = subme > 2 and (w // 2) % 4 == 0 and (h // 2) % 4 == 0
        if not on: assert np.array_equal(want, luma)
        else: decided += int((want[:, :2] != luma[:, :2]).any(1).sum())
    assert decided > 5                                    # the chroma term moved vectors, not only costs
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_bidir_satd_batch(depth):
    """bi-prediction candidate cost (two motion-compensated blocks, rounded average, SATD) vs the oracle's composition --
    pinned to predInterSearch's slot sequence on the reference table by the CPU suite -- for every luma PU shape, all
    fraction pairs incl. zero fractions, different reference planes and strides for the two lists"""
    import torch
    from gpulib import context
    from frames import Geometry, make_plane
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(416, 240); geo1 = Geometry(480, 256)
    F = make_plane(geo, depth, 121, "natural"); R0 = make_plane(geo, depth, 122, "natural"); R1 = make_plane(geo1, depth, 123, "uniform")
    rng = np.random.default_rng(900 + depth)
    cw, ch = geo.coded()
    dF, dR0, dR1 = dev(pix_view(F, depth)), dev(pix_view(R0, depth)), dev(pix_view(R1, depth))
    for (w, h) in LUMA_PU:
        n = 41
        x = rng.integers(0, cw - w, n); y = rng.integers(0, ch - h, n)
        offF = (geo.origin + y * geo.stride + x).astype(np.int32)
        off0 = (geo.origin + (y + rng.integers(-30, 31, n)) * geo.stride + x + rng.integers(-30, 31, n)).astype(np.int32)
        off1 = (geo1.origin + (y + rng.integers(-30, 31, n)) * geo1.stride + x + rng.integers(-30, 31, n)).astype(np.int32)
        f0 = (rng.integers(0, 4, n) | (rng.integers(0, 4, n) << 4)).astype(np.int32)
        f1 = (rng.integers(0, 4, n) | (rng.integers(0, 4, n) << 4)).astype(np.int32)
        f0[:6] = [0, 0, 1, 0x20, 0x33, 0]; f1[:6] = [0, 0x12, 0, 0, 0x33, 3]
        want = orc.bidir_satd_batch(w, h, F, geo.stride, offF, R0, geo.stride, off0, f0, R1, geo1.stride, off1, f1)
        cost = torch.full((n,), -3, dtype=torch.int32, device="cuda")
        ctx.bidir_satd_batch(w, h, dF, geo.stride, dev(offF), dR0, geo.stride, dev(off0), dev(f0), dR1, geo1.stride, dev(off1), dev(f1), cost)
        assert np.array_equal(cost.cpu().numpy(), want), (w, h)
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_lowres_intra_batch(depth):
    """the lookahead's intra estimate (cost + mode per 8x8 lowres CU, all CUs of a frame in one launch) vs the oracle, which the
    CPU suite pins to lowresIntraEstimate's slot sequence: natural, noise and smooth pictures, picture-edge CUs included"""
    import torch
    from gpulib import context
    from frames import Geometry, make_plane, smooth_field
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(416, 240)
    cw, ch = geo.coded()
    wcu, hcu = cw // 8, ch // 8
    modes = set()
    for seed, kind in ((1, "natural"), (2, "uniform"), (3, "smooth")):
        P = smooth_field(geo, depth, seed, box=5) if kind == "smooth" else make_plane(geo, depth, seed, kind)
        want_c, want_m = orc.lowres_intra_frame(P, geo.origin, geo.stride, wcu, hcu, 37)
        cost = torch.full((wcu * hcu,), -1, dtype=torch.int32, device="cuda"); mode = torch.full((wcu * hcu,), -1, dtype=torch.int32, device="cuda")
        ctx.lowres_intra_batch(dev(pix_view(P, depth)), geo.origin, geo.stride, wcu, hcu, 37, cost, mode)
        assert np.array_equal(cost.cpu().numpy(), want_c), kind
        assert np.array_equal(mode.cpu().numpy(), want_m), kind
        modes |= set(want_m.tolist())
    assert len(modes) > 12
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_intra_pred_batch(depth):
    """all 35 luma intra predictions of a batch of TUs (4x4 .. 32x32) vs the oracle composite (pinned to intrapred.cpp's slots
    and the analysis' filter rules by the CPU suite), then the mode costs the analysis takes from them: sa8d of every
    prediction against the TU's source block through the batched metric entry"""
    import torch
    from gpulib import context
    from cpulibs import OP_SA8D
    ctx = context(depth); orc = Oracle(depth)
    rng = np.random.default_rng(1000 + depth)
    for N in (4, 8, 16, 32):
        n, L = 14, 4 * N + 1
        nbs = np.zeros((n, L), orc.pix)
        for i in range(n):
            k = i % 4
            if k == 0: nbs[i] = rng.integers(0, orc.pmax + 1, L)
            elif k == 1: nbs[i] = np.clip(np.cumsum(rng.integers(-9, 10, L)) + orc.pmax // 2, 0, orc.pmax)
            elif k == 2: nbs[i] = (np.arange(L) % 2) * orc.pmax
            else: nbs[i] = orc.pmax if i % 8 == 3 else 0
        want = np.concatenate([orc.intra_pred_all(N, nbs[i]) for i in range(n)])
        dst = torch.zeros(n * 35 * N * N, dtype=torch.uint8 if depth == 8 else torch.int16, device="cuda")
        ctx.intra_pred_batch(N, dev(pix_view(nbs.ravel(), depth)), n, dst)
        got = dst.cpu().numpy()
        assert np.array_equal(got.view(np.uint16) if depth > 8 else got, want), N
        if N >= 8:
            fenc = rng.integers(0, orc.pmax + 1, n * N * N).astype(orc.pix)
            offA = np.repeat(np.arange(n) * N * N, 35).astype(np.int32)
            offB = (np.arange(n * 35) * N * N).astype(np.int32)
            cost = torch.zeros(n * 35, dtype=torch.int32, device="cuda")
            ctx.pixelcmp_batch(OP_SA8D, N, N, dev(pix_view(fenc, depth)), N, dst, N, dev(offA), dev(offB), cost)
            ref = np.array([orc.sa8d(N, N, fenc, int(a), N, want, int(b), N) for a, b in zip(offA, offB)], np.int32)
            assert np.array_equal(cost.cpu().numpy(), ref), (N, "sa8d")
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_intra_host_slots(depth):
    """the per-call intra slots (intra_pred[35], intra_filter, intra_pred_allangs) vs the oracle, every size and mode; the reference's own
    IntraPredHarness runs against the same entries through the table in test_gpu_testbench.py"""
    import ctypes as C
    from gpulib import context
    ctx = context(depth); orc = Oracle(depth)
    rng = np.random.default_rng(1100 + depth)
    for N in (4, 8, 16, 32):
        L = 4 * N + 1
        for trial in range(3):
            nb = rng.integers(0, orc.pmax + 1, L).astype(orc.pix) if trial < 2 else ((np.arange(L) % 3 == 0) * orc.pmax).astype(orc.pix)
            filt = orc.intra_filter(N, nb)
            got = np.zeros(L, orc.pix)
            ctx.lib.x265b200_intra_filter(ctx.h, N, C.c_void_p(nb.ctypes.data), C.c_void_p(got.ctypes.data))
            assert np.array_equal(got, filt), N
            for mode in range(35):
                for bFilter in (0, 1):
                    stride = N + 5
                    out = np.full(N * stride, 7, orc.pix)
                    ctx.lib.x265b200_intra_pred(ctx.h, N, mode, C.c_void_p(out.ctypes.data), C.c_ssize_t(stride), C.c_void_p(nb.ctypes.data), bFilter)
                    want = orc.intra_pred(N, mode, nb, bFilter).reshape(N, N)
                    assert np.array_equal(out.reshape(N, stride)[:, :N], want), (N, mode, bFilter)
                    assert (out.reshape(N, stride)[:, N:] == 7).all()
            from cpulibs import Reference, have_reference
            if have_reference(depth):
                for bLuma in (0, 1):
                    want = Reference(depth).intra_allangs(N, nb.copy(), filt.copy(), bLuma)
                    got33 = np.zeros(33 * N * N, orc.pix)
                    ctx.lib.x265b200_intra_pred_allangs(ctx.h, N, C.c_void_p(got33.ctypes.data), C.c_void_p(nb.ctypes.data), C.c_void_p(filt.ctypes.data), bLuma)
                    assert np.array_equal(got33, want), (N, bLuma)
    ctx.check()


@pytest.mark.parametrize("depth", DEPTHS)
def test_lookahead_mvp_and_bidir_batches(depth):
    """predictor selection and bi-directional candidates of the lookahead for every 8x8 CU of a lowres frame vs the oracle (pinned to the
    reference's lowresMC / bufSATD / pixelavg_pp by the CPU suite)"""
    import torch
    from gpulib import context
    from frames import Geometry
    from test_oracle_vs_ref import lowres_planes
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(416, 240)
    rng = np.random.default_rng(1300 + depth)
    P0 = lowres_planes(geo, depth, 95); P1 = lowres_planes(geo, depth, 96)
    pitch = geo.plane_elems
    cw, ch = geo.coded()
    xs, ys = np.meshgrid(np.arange(0, cw, 8), np.arange(0, ch, 8))
    off = (geo.origin + ys.ravel() * geo.stride + xs.ravel()).astype(np.int32)
    n = len(off)
    F = P1[:pitch].copy()
    numc = rng.integers(0, 6, n).astype(np.int32)
    mvc = rng.integers(-40, 41, (n, 5, 2)).astype(np.int32)
    mvc[::3, 0] = 0; mvc[1::5, 2] = 0
    for bid in (0, 1):
        mvp = torch.full((n, 2), -99, dtype=torch.int32, device="cuda"); mc = torch.zeros(n, dtype=torch.int32, device="cuda"); sk = torch.zeros(n, dtype=torch.int32, device="cuda")
        ctx.lowres_mvp_batch(dev(pix_view(F, depth)), geo.stride, dev(off), dev(pix_view(P0, depth)), geo.stride, pitch, dev(off), dev(mvc.ravel()), dev(numc), bid, mvp, mc, sk)
        want = np.array([orc.lowres_mvp(F, int(off[i]), geo.stride, P0, int(off[i]), geo.stride, pitch, mvc[i, :numc[i]], bid) for i in range(n)], np.int32)
        assert np.array_equal(mvp.cpu().numpy(), want[:, :2]), bid
        assert np.array_equal(mc.cpu().numpy(), want[:, 2]) and np.array_equal(sk.cpu().numpy(), want[:, 3]), bid
    mv0 = rng.integers(-40, 41, (n, 2)).astype(np.int32); mv1 = rng.integers(-40, 41, (n, 2)).astype(np.int32)
    mv0[::4] = 0; mv1[::6] &= ~1
    cost = torch.zeros(2 * n, dtype=torch.int32, device="cuda")
    ctx.lowres_bidir_cost_batch(dev(pix_view(F, depth)), geo.stride, dev(off), dev(pix_view(P0, depth)), geo.stride, pitch, dev(pix_view(P1, depth)), geo.stride, pitch,
                                dev(off), dev(mv0.ravel()), dev(mv1.ravel()), cost)
    want = np.array([orc.lowres_bidir(F, int(off[i]), geo.stride, P0, int(off[i]), geo.stride, pitch, P1, int(off[i]), geo.stride, pitch, mv0[i], mv1[i]) for i in range(n)], np.int32)
    assert np.array_equal(cost.cpu().numpy().reshape(n, 2), want)
    ctx.check()
