"""Pins oracle/x265_oracle.c against the reference's own C primitives (oracle/_ref).

Runs wherever oracle/_ref/libx265ref_<depth>.so exists (the authoring container, and the GPU box
because the prebuilt .so travels); skipped otherwise.  Covers the TestBench input classes
(random / all-min / all-max) plus the cases TestBench never feeds:
full-range int16 into IDCT / ss / ps filters, real quant tables, ADS.
"""
import numpy as np
import pytest

from cpulibs import (CHROMA_ONLY_420, CHROMA_ONLY_422, LUMA_PU, SATD_CHROMA_422, Oracle, Reference,
                     have_reference)

pytestmark = pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built (needs /root/reference)")

DEPTHS = [8, 10, 12]


@pytest.fixture(scope="module", params=DEPTHS)
def libs(request):
    d = request.param
    return Oracle(d), Reference(d)


def pixel_bufs(rng, depth, n):
    dt = np.uint8 if depth == 8 else np.uint16
    pmax = (1 << depth) - 1
    yield rng.integers(0, pmax + 1, n).astype(dt), rng.integers(0, pmax + 1, n).astype(dt)
    yield np.zeros(n, dt), np.full(n, pmax, dt)
    yield np.full(n, pmax, dt), np.zeros(n, dt)
    chk = ((np.arange(n) + np.arange(n) // 64) & 1).astype(dt) * pmax      # max-contrast checkerboard
    yield chk, (pmax - chk).astype(dt)


def test_tables(libs):
    o, r = libs
    for n in (4, 8, 16, 32):
        assert np.array_equal(o.dct_matrix(n), r.dct_matrix(n))
    assert np.array_equal(o.luma_taps(), r.luma_taps())
    assert np.array_equal(o.chroma_taps(), r.chroma_taps())


def test_sad_satd_all_pu(libs):
    o, r = libs
    rng = np.random.default_rng(1)
    for a, b in pixel_bufs(rng, o.depth, 64 * 80 + 200 * 80):
        for (w, h) in LUMA_PU:
            for oa, ob, sa, sb in ((0, 0, 64, 64), (5, 37, 64, 171), (64, 3, 70, 64)):
                assert o.sad(w, h, a, oa, sa, b, ob, sb) == r.sad(w, h, a, oa, sa, b, ob, sb)
                assert o.satd(w, h, a, oa, sa, b, ob, sb) == r.satd(w, h, a, oa, sa, b, ob, sb)


def test_satd_chroma_shapes(libs):
    o, r = libs
    rng = np.random.default_rng(2)
    for a, b in pixel_bufs(rng, o.depth, 100 * 100):
        for (w, h) in SATD_CHROMA_422 + [(8, 4), (4, 4), (16, 12)]:
            assert o.satd(w, h, a, 3, 100, b, 7, 97) == r.satd(w, h, a, 3, 100, b, 7, 97), (w, h)


def test_sa8d(libs):
    o, r = libs
    rng = np.random.default_rng(3)
    for a, b in pixel_bufs(rng, o.depth, 64 * 80 + 200 * 80):
        for w in (4, 8, 16, 32, 64):
            assert o.sa8d(w, w, a, 1, 64, b, 9, 150) == r.sa8d(w, w, a, 1, 64, b, 9, 150)
        for w in (4, 8, 16, 32):                       # 4:2:0 chroma cu slots (cu = 2w)
            assert o.sa8d(w, w, a, 1, 64, b, 9, 150) == r.sa8d(w, w, a, 1, 64, b, 9, 150, chroma=1)
        for w in (4, 8, 16, 32):                       # 4:2:2 chroma cu slots: w x 2w
            assert o.sa8d(w, 2 * w, a, 1, 64, b, 9, 150) == r.sa8d(w, 2 * w, a, 1, 64, b, 9, 150, chroma=2)


def test_sad_x3_x4(libs):
    o, r = libs
    rng = np.random.default_rng(4)
    for a, b in pixel_bufs(rng, o.depth, 64 * 80 + 200 * 80):
        for (w, h) in LUMA_PU:
            offs = [0, 1, 2, 3]
            assert np.array_equal(o.sad_x3(w, h, a, 0, b, offs, 59), r.sad_x3(w, h, a, 0, b, offs, 59))
            offs = [200, 7, 64 * 3 + 1, 1000]
            assert np.array_equal(o.sad_x4(w, h, a, 64, b, offs, 131), r.sad_x4(w, h, a, 64, b, offs, 131))


def test_sse(libs):
    o, r = libs
    rng = np.random.default_rng(5)
    pmax = o.pmax
    for a, b in pixel_bufs(rng, o.depth, 64 * 80 + 200 * 80):
        for w in (4, 8, 16, 32, 64):
            assert o.sse_pp(w, w, a, 2, 64, b, 5, 99) == r.sse_pp(w, w, a, 2, 64, b, 5, 99)
        for w in (4, 8, 16, 32):                       # chroma cu slots: 4:2:0 w x w, 4:2:2 w x 2w
            if True:
                assert o.sse_pp(w, w, a, 2, 64, b, 5, 99) == r.sse_pp(w, w, a, 2, 64, b, 5, 99, chroma=1)
                assert o.sse_pp(w, 2 * w, a, 2, 64, b, 5, 99) == r.sse_pp(w, 2 * w, a, 2, 64, b, 5, 99, chroma=2)
    for lo, hi in ((-pmax, pmax + 1), (-4096, 4097), (-23000, 23000)):
        sa_ = rng.integers(lo, hi, 64 * 80).astype(np.int16)
        sb_ = rng.integers(lo, hi, 64 * 80).astype(np.int16)
        for w in (4, 8, 16, 32, 64):
            assert o.sse_ss(w, sa_, 3, 64, sb_, 1, 70) == r.sse_ss(w, sa_, 3, 64, sb_, 1, 70)
            assert o.ssd_s(w, sa_, 5, 66) == r.ssd_s(w, sa_, 5, 66)
    full = rng.integers(-32768, 32768, 64 * 80).astype(np.int16)
    for w in (4, 8, 16, 32, 64):
        assert o.ssd_s(w, full, 5, 66) == r.ssd_s(w, full, 5, 66)


def test_ads(libs):
    o, r = libs
    rng = np.random.default_rng(6)
    stride = 300
    for trial in range(6):
        scale = [1 << 10, 1 << 16, 1 << 22, 0xFFFFFFFF][trial % 4]
        sums = rng.integers(0, scale, stride * 80, dtype=np.uint64).astype(np.uint32)
        cost = rng.integers(0, 4000, 256).astype(np.uint16)
        for (w, h) in LUMA_PU:
            enc = rng.integers(0, min(scale, 1 << 30), 4).astype(np.int32)
            delta = (h >> 1) * stride if trial % 2 else rng.integers(1, 4000)
            width = int(rng.integers(1, 29)) * 4
            for thresh in (0, int(scale // 2 % (1 << 31)), (1 << 31) - 1):
                n0, m0 = o.ads(w, h, enc, sums, 17, delta, cost, width, thresh)
                n1, m1 = r.ads(w, h, enc, sums, 17, delta, cost, width, thresh)
                assert n0 == n1 and np.array_equal(m0, m1), (w, h, trial, thresh)


def residual_inputs(rng, pmax, n):
    yield rng.integers(-pmax, pmax + 1, n).astype(np.int16)
    yield np.full(n, -pmax, np.int16)
    yield np.full(n, pmax, np.int16)
    yield rng.integers(-32768, 32768, n).astype(np.int16)
    yield np.full(n, 32767, np.int16)
    yield np.full(n, -32768, np.int16)
    alt = np.where(np.arange(n) & 1, 32767, -32768).astype(np.int16)
    yield alt


def test_dct_idct_dst(libs):
    o, r = libs
    rng = np.random.default_rng(7)
    for src in residual_inputs(rng, o.pmax, 64 * 40 + 64):
        for n in (4, 8, 16, 32):
            for off, stride in ((0, n), (16, 64), (3, 37 if n <= 32 else 64)):
                if stride < n:
                    continue
                assert np.array_equal(o.dct(n, src, off, stride), r.dct(n, src, off, stride)), n
            for stride in (n, 64):
                assert np.array_equal(o.idct(n, src, stride), r.idct(n, src, stride)), n
        assert np.array_equal(o.dst4(src, 5, 4), r.dst4(src, 5, 4))
        assert np.array_equal(o.dst4(src, 5, 64), r.dst4(src, 5, 64))
        assert np.array_equal(o.idst4(src, 4), r.idst4(src, 4))
        assert np.array_equal(o.idst4(src, 9), r.idst4(src, 9))


def test_lowpass_dct(libs):
    o, r = libs
    rng = np.random.default_rng(8)
    for src in residual_inputs(rng, o.pmax, 64 * 40 + 64):
        for n in (8, 16, 32):
            for off, stride in ((0, n), (7, 64)):
                assert np.array_equal(o.lowpass_dct(n, src, off, stride), r.lowpass_dct(n, src, off, stride)), n


def test_quant_family(libs):
    o, r = libs
    rng = np.random.default_rng(9)
    pmax = o.pmax
    flat = np.array([26214, 23302, 20560, 18396, 16384, 14564], np.int32)     # scalinglist.cpp:129
    for trial in range(40):
        n = [16, 64, 256, 1024][trial % 4]
        kind = trial % 5
        if kind == 0:
            coef = rng.integers(-pmax, pmax + 1, n).astype(np.int16)
            qc = rng.integers(-pmax, pmax + 1, n).astype(np.int32)
        elif kind == 1:
            coef = np.full(n, -pmax, np.int16); qc = np.full(n, -pmax, np.int32)
        elif kind == 2:
            coef = rng.integers(-32768, 32768, n).astype(np.int16)
            qc = np.full(n, flat[trial % 6], np.int32)
        elif kind == 3:
            coef = rng.integers(-32768, 32768, n).astype(np.int16)
            qc = rng.integers(1, 1 << 17, n).astype(np.int32)               # scaling-list magnitude
        else:
            coef = np.full(n, -32768, np.int16); qc = np.full(n, 26214 << 2, np.int32)
        for qbits in (8, 9, 14, 17, 21, 25):
            for add in ((171 << (qbits - 9)) if qbits >= 9 else 85, 1 << (qbits - 1)):
                a0 = o.quant(coef, qc, qbits, add, n); a1 = r.quant(coef, qc, qbits, add, n)
                assert a0[0] == a1[0] and np.array_equal(a0[1], a1[1]) and np.array_equal(a0[2], a1[2])
                b0 = o.nquant(coef, qc, qbits, add, n); b1 = r.nquant(coef, qc, qbits, add, n)
                assert b0[0] == b1[0] and np.array_equal(b0[1], b1[1])
        for shift in (1, 2, 3, 5, 8, 10):
            for scale in (40, 72 << 4, 64 << 8, 32767):
                assert np.array_equal(o.dequant_normal(coef, n, scale, shift), r.dequant_normal(coef, n, scale, shift))
        dq = rng.integers(1, 1 << 12, n).astype(np.int32)
        for per in (0, 3, 8, 12):
            for shift in (0, 1, 4, 6):
                assert np.array_equal(o.dequant_scaling(coef, dq, n, per, shift), r.dequant_scaling(coef, dq, n, per, shift))


def interp_case(o, r, kind, N, w, h, src, ss, ds, idx, extra, src_off):
    dt = o.pix if kind in ("hpp", "vpp", "vsp", "hvpp") else np.int16
    d0 = np.full(200 * 200, 123, dt); d1 = d0.copy()
    r0 = o.interp(kind, N, w, h, src, src_off, ss, d0, 5, ds, idx, extra)
    r1 = r.interp(kind, N, w, h, src, src_off, ss, d1, 5, ds, idx, extra)
    assert r1 == 0, (kind, N, w, h)
    assert np.array_equal(d0, d1), (kind, N, w, h, idx, extra)


def test_interp_luma(libs):
    o, r = libs
    rng = np.random.default_rng(10)
    pmax = o.pmax
    psrcs = [rng.integers(0, pmax + 1, 200 * 200).astype(o.pix), np.zeros(200 * 200, o.pix), np.full(200 * 200, pmax, o.pix)]
    ssrcs = [rng.integers(-4096, 4096, 200 * 200).astype(np.int16), rng.integers(-32768, 32768, 200 * 200).astype(np.int16)]
    for (w, h) in LUMA_PU:
        for idx in (1, 2, 3):
            ss = int(rng.integers(w + 8, 130)); ds = int(rng.integers(64, 160))
            off = 4 * ss + 8
            for ps in psrcs:
                interp_case(o, r, "hpp", 8, w, h, ps, ss, ds, idx, 0, off)
                interp_case(o, r, "vpp", 8, w, h, ps, ss, ds, idx, 0, off)
                interp_case(o, r, "hps", 8, w, h, ps, ss, ds, idx, 0, off)
                interp_case(o, r, "hps", 8, w, h, ps, ss, ds, idx, 1, off)
                interp_case(o, r, "vps", 8, w, h, ps, ss, ds, idx, 0, off)
                interp_case(o, r, "hvpp", 8, w, h, ps, ss, ds, idx, 1 + (idx % 3), off)
            for s16 in ssrcs:
                interp_case(o, r, "vsp", 8, w, h, s16, ss, ds, idx, 0, off)
                interp_case(o, r, "vss", 8, w, h, s16, ss, ds, idx, 0, off)
        dp0 = np.full(200 * 200, 77, np.int16); dp1 = dp0.copy()
        o.p2s(w, h, psrcs[0], 3, 99, dp0, 1, 80); assert r.p2s(w, h, psrcs[0], 3, 99, dp1, 1, 80) == 0
        assert np.array_equal(dp0, dp1)


def test_interp_chroma(libs):
    o, r = libs
    rng = np.random.default_rng(11)
    pmax = o.pmax
    ps = rng.integers(0, pmax + 1, 200 * 200).astype(o.pix)
    s16 = rng.integers(-32768, 32768, 200 * 200).astype(np.int16)
    shapes = LUMA_PU + CHROMA_ONLY_420 + CHROMA_ONLY_422
    for (w, h) in shapes:
        for idx in range(8):
            ss = int(rng.integers(w + 4, 120)); ds = int(rng.integers(64, 160))
            off = 2 * ss + 4
            interp_case(o, r, "hpp", 4, w, h, ps, ss, ds, idx, 0, off)
            interp_case(o, r, "vpp", 4, w, h, ps, ss, ds, idx, 0, off)
            interp_case(o, r, "hps", 4, w, h, ps, ss, ds, idx, idx & 1, off)
            interp_case(o, r, "vps", 4, w, h, ps, ss, ds, idx, 0, off)
            interp_case(o, r, "vsp", 4, w, h, s16, ss, ds, idx, 0, off)
            interp_case(o, r, "vss", 4, w, h, s16, ss, ds, idx, 0, off)
        dp0 = np.full(200 * 200, 77, np.int16); dp1 = dp0.copy()
        o.p2s(w, h, ps, 3, 99, dp0, 1, 80); assert r.p2s(w, h, ps, 3, 99, dp1, 1, 80) == 0
        assert np.array_equal(dp0, dp1)


def test_batched_drivers(libs):
    o, r = libs
    rng = np.random.default_rng(12)
    stride, rows = 256, 160
    A = rng.integers(0, o.pmax + 1, stride * rows).astype(o.pix)
    B = rng.integers(0, o.pmax + 1, stride * rows).astype(o.pix)
    for (w, h) in ((8, 8), (16, 16), (32, 32), (64, 64), (16, 8), (24, 32)):
        n = 50
        offA = (rng.integers(0, rows - h, n) * stride + rng.integers(0, stride - w, n)).astype(np.int32)
        offB = (rng.integers(0, rows - h, n) * stride + rng.integers(0, stride - w, n)).astype(np.int32)
        ops = [0, 1] + ([2, 3] if w == h else [])
        for op in ops:
            assert np.array_equal(o.pixelcmp_batch(op, w, h, A, stride, B, stride, offA, offB),
                                  r.pixelcmp_batch(op, w, h, A, stride, B, stride, offA, offB, nthreads=3))
        if w == h and w <= 32:
            res = o.residual_batch(w, h, A, stride, B, stride, offA, offB)
            off = (np.arange(n) * w * w).astype(np.int32)
            assert np.array_equal(o.dct_batch(w, res, w, off), r.residual_dct_batch(w, A, stride, B, stride, offA, offB, 2))


def test_tu_chain(libs):
    """inter luma TU chain: oracle composition vs the reference's own slots called in x265's order"""
    o, r = libs
    rng = np.random.default_rng(13)
    D = o.depth
    stride = 96
    flat = [26214, 23302, 20560, 18396, 16384, 14564]
    inv = [40, 45, 51, 57, 64, 72]
    for trial in range(60):
        N = [4, 8, 16, 32][trial % 4]
        qp = [4, 17, 22, 27, 32, 37, 45, 51][trial % 8]
        per, rem = qp // 6, qp % 6
        log2 = {4: 2, 8: 3, 16: 4, 32: 5}[N]
        tshift = 15 - D - log2
        qbits = 14 + per + tshift
        add = (171 if trial % 3 == 0 else 85) << (qbits - 9)
        qc = np.full(N * N, flat[rem], np.int32)
        base = rng.integers(0, o.pmax + 1, stride * 40).astype(np.int64)
        noise = rng.integers(-(3 << (D - 8)) * (1 + trial % 5), (3 << (D - 8)) * (1 + trial % 5) + 1, stride * 40)
        fenc = np.clip(base + noise, 0, o.pmax).astype(o.pix)
        pred = base.astype(o.pix) if trial % 7 else fenc.copy()
        if trial % 11 == 5:
            pred = np.clip(base + 9, 0, o.pmax).astype(o.pix); fenc = base.astype(o.pix)   # DC-only residual
        r0 = np.full(stride * 40, 3, o.pix); r1 = r0.copy()
        a = o.tu_chain(N, fenc, 5, stride, pred, 5, stride, qc, qbits, add, inv[rem] << per, 20 - 14 - tshift, r0, 7, stride)
        b = r.tu_chain(N, fenc, 5, stride, pred, 5, stride, qc, qbits, add, inv[rem] << per, 20 - 14 - tshift, r1, 7, stride)
        assert np.array_equal(a[0], b[0]) and a[1:] == b[1:], (N, qp, trial)
        assert np.array_equal(r0, r1), (N, qp, trial)
        # intra luma: DST-VII at 4x4 and no DC-only shortcut (quant.cpp:430-441, :585-603); the other sizes unchanged
        r2 = np.full(stride * 40, 3, o.pix); r3 = r2.copy()
        c = o.tu_chain(N, fenc, 5, stride, pred, 5, stride, qc, qbits, add, inv[rem] << per, 20 - 14 - tshift, r2, 7, stride, ttype=1)
        d = r.tu_chain(N, fenc, 5, stride, pred, 5, stride, qc, qbits, add, inv[rem] << per, 20 - 14 - tshift, r3, 7, stride, ttype=1)
        assert np.array_equal(c[0], d[0]) and c[1:] == d[1:], (N, qp, trial, "intra")
        assert np.array_equal(r2, r3), (N, qp, trial, "intra")
        if N != 4:
            assert np.array_equal(c[0], a[0]) and np.array_equal(r2, r0)


def test_adjacent_slots(libs):
    """sub_ps / add_ps / pixelavg_pp / addAvg through the reference's luma, 4:2:0 and 4:2:2 slots, and frameInitLowres"""
    o, r = libs
    rng = np.random.default_rng(21)
    D = o.depth
    stride, rows = 100, 140
    n = stride * rows
    pixA = rng.integers(0, o.pmax + 1, n).astype(o.pix); pixB = rng.integers(0, o.pmax + 1, n).astype(o.pix)
    # add_ps residuals up to the full int16 range so the clip is exercised; addAvg inputs like ps-filter outputs and extremes
    s16A = rng.integers(-32768, 32768, n).astype(np.int16); s16B = rng.integers(-32768, 32768, n).astype(np.int16)
    cu_shapes = [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (2, 2), (2, 4), (4, 8), (8, 16), (16, 32), (32, 64)]
    pu_shapes = list(LUMA_PU) + list(CHROMA_ONLY_420) + list(CHROMA_ONLY_422)
    for op, shapes, A, B, dt in ((0, cu_shapes, pixA, pixB, np.int16), (1, cu_shapes, pixA, s16B, o.pix),
                                 (2, LUMA_PU, pixA, pixB, o.pix), (3, pu_shapes, s16A, s16B, o.pix)):
        for (w, h) in shapes:
            for oa, ob, od, sa, sb, sd in ((0, 0, 0, stride, stride, stride), (7, 301, 5, 97, stride, 83)):
                d0 = np.full(n, 5, dt); d1 = d0.copy()
                assert r.blockop(op, w, h, A, oa, sa, B, ob, sb, d1, od, sd) == 0, (op, w, h)
                o.blockop(op, w, h, A, oa, sa, B, ob, sb, d0, od, sd)
                assert np.array_equal(d0, d1), (op, w, h)
    for (lw, lh) in ((32, 32), (45, 17), (8, 60)):
        outs = [[np.full(64 * 64, 9, o.pix) for _ in range(4)] for _ in range(2)]
        o.lowres(pixA, 3, stride, *outs[0], 64, lw, lh)
        r.lowres(pixA, 3, stride, *outs[1], 64, lw, lh)
        for a, b in zip(*outs):
            assert np.array_equal(a, b), (lw, lh)


def test_subpel_cmp(libs):
    """sub-pel candidate cost: oracle composition vs subpelCompare's slot sequence on the reference table"""
    o, r = libs
    rng = np.random.default_rng(31)
    stride, rows = 160, 160
    for fenc, ref in pixel_bufs(rng, o.depth, stride * rows):
        for (w, h) in LUMA_PU:
            for op in (0, 1):
                xf, yf = int(rng.integers(0, 4)), int(rng.integers(0, 4))
                for (xF, yF) in ((xf, yf), (0, yf), (xf, 0), (0, 0)):
                    a = o.subpel_cmp(op, w, h, fenc, 64 * 5 + 3, 64, ref, 20 * stride + 21, stride, xF, yF)
                    b = r.subpel_cmp(op, w, h, fenc, 64 * 5 + 3, 64, ref, 20 * stride + 21, stride, xF, yF)
                    assert a == b and a >= 0, (w, h, op, xF, yF)


def test_sea_integral(libs):
    """integral_init{4..32}h / v rows and the whole-picture plane loop vs the reference's own framefilter.cpp functions"""
    o, r = libs
    rng = np.random.default_rng(41)
    stride, rows = 96, 75
    pix = rng.integers(0, o.pmax + 1, stride * rows).astype(o.pix)
    for size in (4, 8, 12, 16, 24, 32):
        a = rng.integers(0, 2 ** 32, stride * 40, dtype=np.uint64).astype(np.uint32); b = a.copy()
        o.integral_inith(size, a, 5 * stride + 3, pix, 7, stride); assert r.integral_inith(size, b, 5 * stride + 3, pix, 7, stride) == 0
        assert np.array_equal(a, b), size
        o.integral_initv(size, a, 2 * stride + 1, stride); assert r.integral_initv(size, b, 2 * stride + 1, stride) == 0
        assert np.array_equal(a, b), size
    pitch = stride * rows + 13
    init = rng.integers(0, 2 ** 32, 12 * pitch, dtype=np.uint64).astype(np.uint32)      # the reference's planes start uninitialised
    sa, sb = init.copy(), init.copy()
    o.me_integral(pix, stride, rows, sa, pitch)
    assert r.me_integral(pix, stride, rows, sb, pitch) == 0
    assert np.array_equal(sa, sb)
    # and the closed form the CUDA kernel implements: box sums on the defined region, zero first row
    W = [32, 32, 32, 24, 16, 16, 16, 12, 8, 8, 4, 4]; H = [32, 24, 8, 32, 16, 12, 4, 16, 32, 8, 16, 4]
    img = pix.reshape(rows, stride).astype(np.int64)
    I = np.zeros((rows + 1, stride + 1), np.int64); I[1:, 1:] = img.cumsum(0).cumsum(1)
    for k in range(12):
        S = sb[k * pitch:k * pitch + stride * rows].reshape(rows, stride)
        assert not S[0].any()
        rr = np.arange(1, rows - H[k]); xx = np.arange(0, stride - W[k])
        box = I[rr[:, None] + H[k], xx[None, :] + W[k]] - I[rr[:, None], xx[None, :] + W[k]] - I[rr[:, None] + H[k], xx[None, :]] + I[rr[:, None], xx[None, :]]
        assert np.array_equal(S[1:rows - H[k], :stride - W[k]].astype(np.int64), box), k


def test_weighted_prediction(libs):
    """weight_pp / weight_sp slots and the lookahead's weightCostLuma composition"""
    o, r = libs
    rng = np.random.default_rng(51)
    D = o.depth
    corr = 14 - D
    stride, rows = 128, 90
    pix = rng.integers(0, o.pmax + 1, stride * rows).astype(o.pix)
    s16 = rng.integers(-8192, 8192, stride * rows).astype(np.int16)
    for trial in range(12):
        shift = int(rng.integers(1, 7)); w0 = int(rng.integers(1, 127)); offset = int(rng.integers(-128, 128)) << (D - 8)
        rnd = 1 << (shift - 1)
        a = np.full(stride * rows, 3, o.pix); b = a.copy()
        o.weight_pp(pix, 5, a, 9, stride, 64, 40, w0, rnd << corr, shift + corr, offset); r.weight_pp(pix, 5, b, 9, stride, 64, 40, w0, rnd << corr, shift + corr, offset)
        assert np.array_equal(a, b)
        o.weight_sp(s16, 7, a, 2, stride, 100, 37, 21, w0, rnd << corr, shift + corr, offset); r.weight_sp(s16, 7, b, 2, stride, 100, 37, 21, w0, rnd << corr, shift + corr, offset)
        assert np.array_equal(a, b)
    fenc = rng.integers(0, o.pmax + 1, stride * rows).astype(o.pix)
    ref = np.clip(fenc.astype(np.int64) * 3 // 4 + rng.integers(-20, 21, stride * rows), 0, o.pmax).astype(o.pix)
    W, H = 104, 72
    intra = rng.integers(0, 3000 << (D - 8), ((W + 7) // 8) * ((H + 7) // 8)).astype(np.int32)
    weights = np.array([0, 0, -1, 0] + [64, 32 << corr, 6 + corr, 0] + [85, 32 << corr, 6 + corr, 3 << (D - 8)] + [43, 16 << corr, 5 + corr, -2 << (D - 8)], np.int32)
    for ic in (intra, None):
        assert np.array_equal(o.weight_cost(fenc, 0, ref, 0, stride, W, H, ic, weights), r.weight_cost(fenc, 0, ref, 0, stride, W, H, ic, weights))


def test_copy_family(libs):
    """copy_pp/ss/sp/ps, blockfill_s, cpy2Dto1D_shl/shr, cpy1Dto2D_shl/shr vs the reference slots"""
    o, r = libs
    rng = np.random.default_rng(61)
    stride, rows = 100, 100
    n = stride * rows
    pix = rng.integers(0, o.pmax + 1, n).astype(o.pix)
    s16 = rng.integers(-32768, 32768, n).astype(np.int16)
    s16pix = rng.integers(0, o.pmax + 1, n).astype(np.int16)            # copy_sp inputs are pixel-valued (pixel.cpp:784)
    for (w, h) in LUMA_PU:
        a = np.full(n, 7, o.pix); b = a.copy()
        o.blockcopy(0, w, h, a, 3, stride, pix, 11, 97); assert r.blockcopy(0, w, h, b, 3, stride, pix, 11, 97) == 0
        assert np.array_equal(a, b), (w, h)
    for w in (4, 8, 16, 32, 64):
        for kind, src, dt in ((1, s16, np.int16), (2, s16pix, o.pix), (3, pix, np.int16), (4, None, np.int16)):
            a = np.full(n, 7, dt); b = a.copy()
            o.blockcopy(kind, w, w, a, 3, stride, src, 11, 97, -1234); assert r.blockcopy(kind, w, w, b, 3, stride, src, 11, 97, -1234) == 0
            assert np.array_equal(a, b), (kind, w)
        for kind in (5, 6):
            for shift in (1, 3, 6):
                for (ds, ss) in ((w, 97), (stride, w)):                 # 2-D -> 1-D and 1-D -> 2-D
                    a = np.full(n, 7, np.int16); b = a.copy()
                    o.blockcopy(kind, w, w, a, 0, ds, s16, 8, ss, shift); assert r.blockcopy(kind, w, w, b, 0, ds, s16, 8, ss, shift) == 0
                    assert np.array_equal(a, b), (kind, w, shift, ds)


def test_block_scalars(libs):
    """var, psy_cost_pp, count_nonzero / copy_cnt, denoiseDct vs the reference slots"""
    o, r = libs
    rng = np.random.default_rng(71)
    stride, rows = 100, 100
    n = stride * rows
    for a, b in pixel_bufs(rng, o.depth, n):
        for size in (4, 8, 16, 32, 64):
            assert o.var(size, a, 7, stride) == r.var(size, a, 7, stride), size
            assert o.psy_cost_pp(size, a, 3, stride, b, 11, 97) == r.psy_cost_pp(size, a, 3, stride, b, 11, 97), size
    resi = (rng.integers(-300, 300, n) * (rng.integers(0, 3, n) == 0)).astype(np.int16)
    for size in (4, 8, 16, 32):
        ca = np.zeros(size * size, np.int16); cb = ca.copy()
        assert o.copy_cnt(size, ca, resi, 13, stride) == r.copy_cnt(size, cb, resi, 13, stride)
        assert np.array_equal(ca, cb)
        assert o.copy_cnt(size, None, resi, 13, size) == r.copy_cnt(size, None, resi, 13, size)
        num = size * size
        da = rng.integers(-32768, 32768, num).astype(np.int16); db = da.copy()
        ra = rng.integers(0, 1 << 31, num).astype(np.uint32); rb = ra.copy()
        offs = rng.integers(0, 65535, num).astype(np.uint16)
        o.denoise_dct(da, ra, offs, num); r.denoise_dct(db, rb, offs, num)
        assert np.array_equal(da, db) and np.array_equal(ra, rb)


def mv_cost_table(lam, radius):
    """BitCost::setQP's table (bitcost.cpp:44-54, 91-101) for one lambda, as a uint16 array centred at `radius`"""
    i = np.arange(radius + 1, dtype=np.float32)
    bits = np.log(i + np.float32(1)).astype(np.float32) * np.float32(2.0 / np.log(np.float32(2.0))) + np.float32(1.718)
    bits[0] = np.float32(0.718)
    half = np.minimum(bits.astype(np.float64) * lam + 0.5, (1 << 15) - 1).astype(np.uint16)
    return np.concatenate([half[:0:-1], half])


def test_me_full_search(libs):
    """exhaustive integer search: oracle loop vs the reference's sad / sad_x4 slots driven in motion.cpp's order"""
    o, r = libs
    rng = np.random.default_rng(53)
    stride, rows = 192, 160
    R = 1024
    tabs = [mv_cost_table(4.0, R), mv_cost_table(57.0175, R), rng.integers(0, 65536, 2 * R + 1).astype(np.uint16),
            np.zeros(2 * R + 1, np.uint16)]
    for bi, (fenc, ref) in enumerate(pixel_bufs(rng, o.depth, stride * rows)):
        if bi == 0:
            ref[:] = fenc                      # a natural-looking match somewhere in the window
            ref[1:] = np.where(rng.random(ref.size - 1) < 0.3, fenc[:-1], ref[1:])
        for (w, h) in ((64, 64), (32, 24), (16, 16), (16, 4), (12, 16), (8, 8), (4, 8), (4, 4)):
            for tab in tabs:
                minx, miny = -int(rng.integers(0, 20)), -int(rng.integers(0, 14))
                maxx, maxy = int(rng.integers(0, 20)), int(rng.integers(0, 14))
                if rng.random() < 0.15: maxx = minx + int(rng.integers(0, 3))          # the x4 tail path only
                mvp = rng.integers(-60, 61, 2).astype(np.int32)
                bmv0 = (int(rng.integers(minx, maxx + 1)), int(rng.integers(miny, maxy + 1)))
                of, orf = 7 * stride + 5, 40 * stride + 60
                for bcost0 in (0x7fffffff, 0, None):
                    if bcost0 is None:          # the cost of some candidate: exercises the tie with the starting point
                        bcost0 = o.me_full_search(w, h, fenc, of, stride, ref, orf, stride, [bmv0[0], bmv0[1], bmv0[0], bmv0[1]], mvp, tab, R,
                                                  (0, 0), 0x7fffffff)[2]
                    a = o.me_full_search(w, h, fenc, of, stride, ref, orf, stride, [minx, miny, maxx, maxy], mvp, tab, R, bmv0, bcost0)
                    b = r.me_full_search(w, h, fenc, of, stride, ref, orf, stride, [minx, miny, maxx, maxy], mvp, tab, R, bmv0, bcost0)
                    assert a == b, (w, h, a, b)
                    assert minx <= a[0] <= maxx and miny <= a[1] <= maxy
    # closed form on one case: every candidate's cost in numpy, first minimum in raster order
    fenc, ref = next(pixel_bufs(rng, o.depth, stride * rows))
    w, h, of, orf = 8, 8, 3 * stride + 9, 50 * stride + 70
    tab = tabs[1]; mvp = np.array([5, -9], np.int32)
    F = fenc.reshape(rows, stride)[3:3 + h, 9:9 + w].astype(np.int64)
    P = ref.reshape(rows, stride).astype(np.int64)
    costs = np.array([[np.abs(F - P[50 + y:50 + y + h, 70 + x:70 + x + w]).sum() + ((int(tab[R + 4 * x - 5]) + int(tab[R + 4 * y + 9])) & 0xffff)
                       for x in range(-9, 12)] for y in range(-6, 8)])
    k = int(costs.argmin())
    got = o.me_full_search(w, h, fenc, of, stride, ref, orf, stride, [-9, -6, 11, 7], mvp, tab, R, (0, 0), 0x7fffffff)
    assert got == (k % 21 - 9, k // 21 - 6, int(costs.min()))


def test_motion_estimate(libs):
    """the oracle's motionEstimate restatement (full, hexagon, diamond and star searches) vs the reference's MotionEstimate::motionEstimate itself
    (compiled from encoder/motion.cpp), every subme level, with neighbour candidates and predictors inside / outside the range"""
    o, r = libs
    from frames import Geometry, make_plane
    geo = Geometry(192, 128)
    rng = np.random.default_rng(61)
    F = make_plane(geo, o.depth, 71, "natural"); R = make_plane(geo, o.depth, 72, "natural")
    R2 = np.roll(F, 3 * geo.stride + 5)                   # a real match at (+5, +3): zero-cost exits and tight refinement
    from frames import smooth_field
    S = smooth_field(geo, o.depth, 73)                    # smooth texture: the pattern searches walk many steps
    S2 = np.roll(S, -6 * geo.stride + 9) + (rng.integers(0, 3, S.size)).astype(S.dtype)
    np.clip(S2, 0, o.pmax, out=S2)
    moved = 0
    RAD = 2048
    cw, ch = geo.coded()
    nonzero_exit = 0
    for case in range(160):
        w, h = [(8, 8), (16, 16), (16, 8), (32, 32), (8, 16), (64, 64), (24, 32), (12, 16)][case % 8]
        subme = case % 8 if case < 128 else int(rng.integers(0, 8))
        qp = int(rng.integers(0, 52))
        tab = r.mvcost_table(qp, RAD)
        ref = R2 if case % 5 == 0 else R
        fen = F
        if case % 3 == 1: fen, ref = S, S2
        x = int(rng.integers(0, cw - w + 1)); y = int(rng.integers(0, ch - h + 1))
        of = geo.origin + y * geo.stride + x
        m = int(rng.integers(1, 17)) if case % 4 else int(rng.integers(17, 40))      # big windows: star rings of 16 / 32, raster pass
        # keep block + 8-tap margins inside the padded plane
        minx = -min(m, x + geo.margin_x - 8); maxx = min(m, cw + geo.margin_x - 8 - w - x)
        miny = -min(m, y + geo.margin_y - 8); maxy = min(m, ch + geo.margin_y - 8 - h - y)
        qmvp = rng.integers(-4 * m - 6, 4 * m + 7, 2)
        if case % 7 == 0: qmvp[:] = 0
        if case % 11 == 0: qmvp = (qmvp // 4) * 4
        if case % 5 == 0 and case % 2 == 0: qmvp[:] = (20, 12)          # exactly the shifted copy: bcost == 0 at the start
        nc = int(rng.integers(0, 5))
        mvc = rng.integers(-4 * m - 6, 4 * m + 7, (nc, 2))
        if nc > 1: mvc[1] = qmvp
        for method in (5, 1, 0, 3, 2):                    # X265_FULL_SEARCH, X265_HEX_SEARCH, X265_DIA_SEARCH, X265_STAR_SEARCH, X265_UMH_SEARCH
            merange = m if method == 5 else int(rng.integers(1, 40))
            a = o.motion_estimate_full(subme, w, h, fen, of, geo.stride, ref, of, geo.stride, [minx, miny, maxx, maxy], qmvp, mvc, tab, RAD,
                                       method, merange)
            b = r.motion_estimate(method, subme, w, h, fen, of, geo.stride, ref, of, geo.stride, [minx, miny, maxx, maxy], qmvp, mvc, merange, qp)
            assert a == b, (case, method, w, h, subme, a, b)
            nonzero_exit += a[2] > 0
            if method != 5:
                moved += max(abs(a[0] - int(np.clip(qmvp[0], 4 * minx, 4 * maxx))), abs(a[1] - int(np.clip(qmvp[1], 4 * miny, 4 * maxy)))) >= 16
    assert nonzero_exit > 500 and moved > 80              # dozens of pattern walks ended four or more pels from their start


def test_motion_estimate_star_far(libs):
    """star search with the match far from the start: big rings, the stride-5 raster pass (with the reference's
    `mvcost(tmv << 3)` on every fourth column) and re-centred passes all decide the result"""
    o, r = libs
    from frames import Geometry, smooth_field
    geo = Geometry(192, 128)
    rng = np.random.default_rng(67)
    S = smooth_field(geo, o.depth, 91, box=21)
    RAD = 2048
    cw, ch = geo.coded()
    far = 0
    for case in range(90):
        dx, dy = int(rng.integers(-30, 31)), int(rng.integers(-24, 25))
        S2 = np.roll(S, dy * geo.stride + dx)              # block content found at (+dx, +dy)
        w, h = [(16, 16), (8, 8), (32, 32), (16, 8), (64, 64), (8, 16)][case % 6]
        qp = int(rng.integers(0, 40))
        tab = r.mvcost_table(qp, RAD)
        x = int(rng.integers(0, cw - w + 1)); y = int(rng.integers(0, ch - h + 1))
        of = geo.origin + y * geo.stride + x
        m = int(rng.integers(20, 48))
        minx = -min(m, x + geo.margin_x - 8); maxx = min(m, cw + geo.margin_x - 8 - w - x)
        miny = -min(m, y + geo.margin_y - 8); maxy = min(m, ch + geo.margin_y - 8 - h - y)
        qmvp = rng.integers(-12, 13, 2)
        mvc = rng.integers(-40, 41, (int(rng.integers(0, 3)), 2))
        merange = int(rng.integers(8, 64))
        a = o.motion_estimate_full(2, w, h, S, of, geo.stride, S2, of, geo.stride, [minx, miny, maxx, maxy], qmvp, mvc, tab, RAD, 3, merange)
        b = r.motion_estimate(3, 2, w, h, S, of, geo.stride, S2, of, geo.stride, [minx, miny, maxx, maxy], qmvp, mvc, merange, qp)
        assert a == b, (case, w, h, a, b)
        u = o.motion_estimate_full(2, w, h, S, of, geo.stride, S2, of, geo.stride, [minx, miny, maxx, maxy], qmvp, mvc, tab, RAD, 2, merange)
        v = r.motion_estimate(2, 2, w, h, S, of, geo.stride, S2, of, geo.stride, [minx, miny, maxx, maxy], qmvp, mvc, merange, qp)
        assert u == v, ("umh", case, w, h, u, v)
        far += max(abs(a[0] - qmvp[0]), abs(a[1] - qmvp[1])) >= 40
    assert far > 30


def lowres_planes(geo, depth, seed):
    """four half-pel planes of a smooth field, `geo.plane_elems` apart, shaped like frameInitLowres's output
    (full-pel, half-pel right, half-pel down, half-pel diagonal) over the padded geometry"""
    from frames import smooth_field
    S = smooth_field(geo, depth, seed, box=7).reshape(geo.rows, geo.stride).astype(np.int64)
    right = np.roll(S, -1, 1); down = np.roll(S, -1, 0); diag = np.roll(right, -1, 0)
    H = (S + right + 1) >> 1; V = (S + down + 1) >> 1; Cc = (S + right + down + diag + 2) >> 2
    dt = np.uint8 if depth == 8 else np.uint16
    return np.concatenate([p.astype(dt).ravel() for p in (S, H, V, Cc)])


def test_lowres_motion_estimate(libs):
    """the lookahead's motionEstimate (lowres reference, 8x8 blocks, quarter-pel cost = average of two half-pel planes) vs the
    reference's MotionEstimate::motionEstimate with ref->isLowres"""
    o, r = libs
    from frames import Geometry
    geo = Geometry(192, 128)
    rng = np.random.default_rng(71)
    P = lowres_planes(geo, o.depth, 95)
    pitch = geo.plane_elems
    RAD = 2048
    cw, ch = geo.coded()
    sub = 0
    for case in range(240):
        dx, dy = int(rng.integers(-12, 13)), int(rng.integers(-10, 11))
        src = P[(case % 4) * pitch:(case % 4 + 1) * pitch] if case % 3 == 0 else P[:pitch]      # fenc sits on a half-pel phase sometimes
        F = np.clip(np.roll(src, dy * geo.stride + dx).astype(np.int64) + rng.integers(-2, 3, pitch), 0, o.pmax).astype(P.dtype)
        qp = int(rng.integers(0, 52))
        tab = r.mvcost_table(qp, RAD)
        x = int(rng.integers(0, cw - 8 + 1)); y = int(rng.integers(0, ch - 8 + 1))
        of = geo.origin + y * geo.stride + x
        m = int(rng.integers(2, 30))
        minx = -min(m, x + geo.margin_x - 8); maxx = min(m, cw + geo.margin_x - 16 - x)
        miny = -min(m, y + geo.margin_y - 8); maxy = min(m, ch + geo.margin_y - 16 - y)
        qmvp = rng.integers(-4 * m - 6, 4 * m + 7, 2)
        if case % 7 == 0: qmvp[:] = 0
        method = (1, 1, 0, 3, 5)[case % 5]
        subme = (1, 1, 2, 0, 5, 7)[case % 6]
        merange = m if method == 5 else int(rng.integers(1, 33))
        a = o.lowres_motion_estimate(method, merange, subme, 8, 8, F, of, geo.stride, P, of, geo.stride, pitch, [minx, miny, maxx, maxy], qmvp, tab, RAD)
        b = r.lowres_motion_estimate_ref(method, subme, 8, 8, F, of, geo.stride, P, of, geo.stride, pitch, [minx, miny, maxx, maxy], qmvp, merange, qp)
        assert a == b, (case, method, subme, a, b)
        sub += (a[0] | a[1]) & 1
    assert sub > 40                                       # quarter-pel winners: the two-plane average decided them


def yuv420_planes(geo, depth, seed, shift=(0, 0)):
    """luma plane over `geo` plus Cb / Cr planes over the half-size geometry (same relative margins), all displaced by
    `shift` luma samples (even) so that a displaced copy matches in all three planes"""
    from frames import Geometry, smooth_field
    cgeo = Geometry(geo.width // 2, geo.height // 2, geo.ctu // 2)
    Y = np.roll(smooth_field(geo, depth, seed, box=9), shift[1] * geo.stride + shift[0])
    Cb = np.roll(smooth_field(cgeo, depth, seed + 1, box=5), (shift[1] // 2) * cgeo.stride + shift[0] // 2)
    Cr = np.roll(smooth_field(cgeo, depth, seed + 2, box=5), (shift[1] // 2) * cgeo.stride + shift[0] // 2)
    return cgeo, Y, Cb, Cr


def test_motion_estimate_chroma(libs):
    """from subme 3 the encoder's motionEstimate charges chroma SATD in every sub-pel cost: oracle vs the reference driven
    through the encoder-style setSourcePU (Yuv source, PicYuv reference), 4:2:0, incl. PU shapes whose chroma block has
    no SATD slot (term off) and subme <= 2 (term off)"""
    o, r = libs
    from frames import Geometry
    geo = Geometry(192, 128)
    rng = np.random.default_rng(73)
    cgeo, FY, FCb, FCr = yuv420_planes(geo, o.depth, 101)
    refs = [yuv420_planes(geo, o.depth, 101, shift=(6, -4))[1:], yuv420_planes(geo, o.depth, 201)[1:]]
    for k in range(2):      # sensor noise so that costs are not all zero at the match
        refs[k] = tuple(np.clip(p.astype(np.int64) + rng.integers(-3, 4, p.size), 0, o.pmax).astype(p.dtype) for p in refs[k])
    RAD = 2048
    cw, ch = geo.coded()
    on_count = 0; chroma_decided = 0
    for case in range(200):
        w, h = [(16, 16), (8, 8), (32, 32), (16, 8), (64, 64), (8, 16), (32, 24), (16, 12), (12, 16), (16, 4), (64, 32), (24, 32)][case % 12]
        subme = (3, 4, 5, 7, 6, 2, 3, 5)[case % 8]
        method = (1, 3, 0, 5, 1)[case % 5]
        qp = int(rng.integers(10, 45))
        tab = r.mvcost_table(qp, RAD)
        RY, RCb, RCr = refs[case % 3 == 2]
        x = int(rng.integers(0, (cw - w) // 2 + 1)) * 2; y = int(rng.integers(0, (ch - h) // 2 + 1)) * 2
        ofY = geo.origin + y * geo.stride + x
        ofC = cgeo.origin + (y // 2) * cgeo.stride + x // 2
        m = int(rng.integers(2, 14))
        minx = -min(m, x + geo.margin_x - 16); maxx = min(m, cw + geo.margin_x - 16 - w - x)
        miny = -min(m, y + geo.margin_y - 16); maxy = min(m, ch + geo.margin_y - 16 - h - y)
        qmvp = rng.integers(-4 * m - 6, 4 * m + 7, 2)
        if case % 7 == 0: qmvp[:] = 0
        mvc = rng.integers(-4 * m - 6, 4 * m + 7, (int(rng.integers(0, 4)), 2))
        merange = m if method == 5 else int(rng.integers(1, 33))
        a = o.motion_estimate_chroma(method, merange, subme, w, h, FY, ofY, geo.stride, RY, ofY, geo.stride, FCb, FCr, ofC, cgeo.stride,
                                     RCb, RCr, ofC, cgeo.stride, 1, 1, [minx, miny, maxx, maxy], qmvp, mvc, tab, RAD)
        b = r.motion_estimate_chroma_ref(method, subme, 1, w, h, FY, ofY, geo.stride, RY, ofY, geo.stride, FCb, FCr, ofC, cgeo.stride,
                                         RCb, RCr, ofC, cgeo.stride, [minx, miny, maxx, maxy], qmvp, mvc, merange, qp)
        assert a == b[:3], (case, method, w, h, subme, a, b)
        assert b[3] == (subme > 2 and (w // 2) % 4 == 0 and (h // 2) % 4 == 0), (w, h, subme)
        on_count += b[3]
        if b[3]:
            luma_only = o.motion_estimate_full(subme, w, h, FY, ofY, geo.stride, RY, ofY, geo.stride, [minx, miny, maxx, maxy], qmvp, mvc, tab, RAD,
                                               method, merange)
            chroma_decided += luma_only[:2] != a[:2]
    assert on_count > 100 and chroma_decided > 5          # the chroma term changed winners, not only costs


def test_bidir_satd(libs):
    """bi-prediction candidate cost: the oracle's composition vs predInterSearch's slot sequence on the reference table
    (motion compensation of both lists by vector fraction, pixelavg_pp, satd)"""
    o, r = libs
    rng = np.random.default_rng(83)
    stride, rows = 192, 176
    for fenc, ref in pixel_bufs(rng, o.depth, stride * rows):
        ref1 = np.roll(ref, 7 * stride + 3)
        for (w, h) in LUMA_PU:
            for _ in range(3):
                f0 = int(rng.integers(0, 4)) | (int(rng.integers(0, 4)) << 4); f1 = int(rng.integers(0, 4)) | (int(rng.integers(0, 4)) << 4)
                for (a0, a1) in ((f0, f1), (0, f1), (f0, 0), (0, 0), (f0 & 3, f1 & 0x30)):
                    x = o.bidir_satd(w, h, fenc, 64 * 5 + 3, 64, ref, 20 * stride + 21, stride, a0, ref1, 31 * stride + 40, stride, a1)
                    y = r.bidir_satd(w, h, fenc, 64 * 5 + 3, 64, ref, 20 * stride + 21, stride, a0, ref1, 31 * stride + 40, stride, a1)
                    assert x == y and x >= 0, (w, h, a0, a1)


def test_intra_pred(libs):
    """reference-sample smoothing and all 35 intra predictors, 4x4 .. 32x32, with and without edge filtering, vs the
    reference's intrapred.cpp slots; the all-angles slot as a cross-check of the filter-flag table"""
    o, r = libs
    rng = np.random.default_rng(89)
    flags = [0x38, 0x00] + ([0x38] + [0x30] * 6 + [0x20, 0x00, 0x20] + [0x30] * 6) * 2 + [0x38]
    for N in (4, 8, 16, 32):
        for kind in range(4):
            if kind == 0: s = rng.integers(0, o.pmax + 1, 4 * N + 1).astype(o.pix)
            elif kind == 1: s = np.full(4 * N + 1, o.pmax, o.pix)
            elif kind == 2: s = (np.arange(4 * N + 1) % 2 * o.pmax).astype(o.pix)
            else: s = np.clip(np.cumsum(rng.integers(-6, 7, 4 * N + 1)) + o.pmax // 2, 0, o.pmax).astype(o.pix)
            f = o.intra_filter(N, s)
            assert np.array_equal(f, r.intra_filter(N, s)), N
            for mode in range(35):
                for bf in (0, 1):
                    for src in (s, f):
                        assert np.array_equal(o.intra_pred(N, mode, src, bf), r.intra_pred(N, mode, src, bf)), (N, mode, bf)
            allang = r.intra_allangs(N, s.copy(), f.copy(), 1)
            for mode in range(2, 35):
                want = o.intra_pred(N, mode, f if flags[mode] & N else s, 1).reshape(N, N)
                got = allang[(mode - 2) * N * N:(mode - 1) * N * N].reshape(N, N)
                assert np.array_equal(got if mode >= 18 else got.T, want), (N, mode)     # all_angs leaves horizontal modes transposed
            # the analysis' 35 predictions of a TU (search.cpp:1703-1727) from the reference's slots vs the oracle's composite
            every = o.intra_pred_all(N, s).reshape(35, N, N)
            assert np.array_equal(every[1].ravel(), r.intra_pred(N, 1, s, int(N <= 16)))
            assert np.array_equal(every[0].ravel(), r.intra_pred(N, 0, f if N >= 8 else s, 0))
            allang = r.intra_allangs(N, s.copy(), f.copy(), int(N <= 16))
            for mode in range(2, 35):
                got = allang[(mode - 2) * N * N:(mode - 1) * N * N].reshape(N, N)
                assert np.array_equal(got if mode >= 18 else got.T, every[mode]), (N, mode)


def test_lowres_intra_estimate(libs):
    """the lookahead's intra cost / mode per 8x8 lowres CU: oracle vs lowresIntraEstimate's slot sequence on the reference table"""
    o, r = libs
    from frames import Geometry, make_plane, smooth_field
    geo = Geometry(128, 64)
    modes = set()
    for seed, kind in ((1, "natural"), (2, "uniform"), (3, "smooth")):
        P = smooth_field(geo, o.depth, seed, box=5) if kind == "smooth" else make_plane(geo, o.depth, seed, kind)
        cw, ch = geo.coded()
        for cy in range(ch // 8):
            for cx in range(cw // 8):
                a = o.lowres_intra_cu(P, geo.origin, geo.stride, cx, cy, 24)
                b = r.lowres_intra_cu(P, geo.origin, geo.stride, cx, cy, 24)
                assert a == b, (kind, cx, cy, a, b)
                modes.add(a[1])
    assert len(modes) > 12                                 # DC, planar and a spread of angular winners


def test_motion_estimate_umh_ladder(libs):
    """UMH's early-termination ladder and adaptive cross: predictors at / near the true displacement and noise levels that
    put the start cost on either side of the 500 / 1000 / 2000 / 4000 thresholds (scaled by the PU height)"""
    o, r = libs
    from frames import Geometry, smooth_field
    geo = Geometry(192, 128)
    rng = np.random.default_rng(97)
    S = smooth_field(geo, o.depth, 131, box=11)
    RAD = 2048
    cw, ch = geo.coded()
    for case in range(320):
        dx, dy = int(rng.integers(-9, 10)), int(rng.integers(-7, 8))
        amp = (0, 1, 2, 3, 5, 8, 12, 20)[case % 8] << (o.depth - 8)
        S2 = np.clip(np.roll(S, dy * geo.stride + dx).astype(np.int64) + rng.integers(-amp, amp + 1, S.size), 0, o.pmax).astype(S.dtype)
        w, h = [(16, 16), (8, 8), (32, 32), (16, 8), (64, 64), (8, 16), (32, 16), (16, 32)][(case // 8) % 8]
        qp = int(rng.integers(0, 30))
        tab = r.mvcost_table(qp, RAD)
        x = int(rng.integers(0, cw - w + 1)); y = int(rng.integers(0, ch - h + 1))
        of = geo.origin + y * geo.stride + x
        m = int(rng.integers(12, 40))
        minx = -min(m, x + geo.margin_x - 8); maxx = min(m, cw + geo.margin_x - 8 - w - x)
        miny = -min(m, y + geo.margin_y - 8); maxy = min(m, ch + geo.margin_y - 8 - h - y)
        off = int(rng.integers(0, 4))
        qmvp = np.array([4 * dx + int(rng.integers(-off, off + 1)) * 4, 4 * dy + int(rng.integers(-off, off + 1)) * 4])
        nc = int(rng.integers(0, 4))
        mvc = qmvp + rng.integers(-30, 31, (nc, 2)) * int(rng.integers(0, 3))
        merange = int(rng.integers(8, 64))
        subme = case % 3
        a = o.motion_estimate_full(subme, w, h, S, of, geo.stride, S2, of, geo.stride, [minx, miny, maxx, maxy], qmvp, mvc, tab, RAD, 2, merange)
        b = r.motion_estimate(2, subme, w, h, S, of, geo.stride, S2, of, geo.stride, [minx, miny, maxx, maxy], qmvp, mvc, merange, qp)
        assert a == b, (case, w, h, amp, a, b)


def test_motion_estimate_sea(libs):
    """successive elimination (X265_SEA) through the whole motionEstimate: oracle vs the reference, both reading the same
    twelve integral planes; PU shapes whose DC sub-blocks lie inside the PU (the reference's 32x8, 8x32, 8x4 and 4x8 cases
    read beyond the block in its cache and are left out)"""
    o, r = libs
    from frames import Geometry, make_plane, smooth_field
    geo = Geometry(192, 128)
    rng = np.random.default_rng(101)
    S = smooth_field(geo, o.depth, 141, box=9)
    N = make_plane(geo, o.depth, 142, "natural")
    RAD = 4096
    cw, ch = geo.coded()
    pitch = geo.plane_elems
    shapes = [(16, 16), (8, 8), (32, 32), (64, 64), (16, 8), (8, 16), (32, 16), (16, 32), (64, 32), (32, 64), (32, 24), (24, 32),
              (64, 48), (48, 64), (64, 16), (16, 64), (16, 12), (12, 16), (16, 4), (4, 16)]
    moved = 0
    for case in range(120):
        dx, dy = int(rng.integers(-10, 11)), int(rng.integers(-8, 9))
        if case % 3 == 2: F, R = N, make_plane(geo, o.depth, 143, "natural")
        else:
            F = S
            R = np.clip(np.roll(S, dy * geo.stride + dx).astype(np.int64) + rng.integers(-3, 4, S.size), 0, o.pmax).astype(S.dtype)
        sums = np.zeros(12 * pitch, np.uint32)
        assert r.me_integral(R, geo.stride, geo.rows, sums, pitch) == 0
        w, h = shapes[case % len(shapes)]
        qp = int(rng.integers(0, 40))
        tab = r.mvcost_table(qp, RAD)
        x = int(rng.integers(0, cw - w + 1)); y = int(rng.integers(0, ch - h + 1))
        of = geo.origin + y * geo.stride + x
        m = int(rng.integers(4, 24))
        minx = -min(m, x + geo.margin_x - 12); maxx = min(m, cw + geo.margin_x - 12 - w - x)
        miny = -min(m, y + geo.margin_y - 12); maxy = min(m, ch + geo.margin_y - 12 - h - y)
        qmvp = rng.integers(-4 * m, 4 * m + 1, 2)
        if case % 5 == 0: qmvp[:] = 0
        mvc = rng.integers(-4 * m, 4 * m + 1, (int(rng.integers(0, 3)), 2))
        merange = int(rng.integers(2, 24))
        subme = case % 4
        a = o.motion_estimate_sea(merange, subme, w, h, F, of, geo.stride, R, of, geo.stride, sums, pitch, [minx, miny, maxx, maxy], qmvp, mvc, tab, RAD)
        b = r.motion_estimate_sea_ref(subme, w, h, F, of, geo.stride, R, of, geo.stride, sums, pitch, [minx, miny, maxx, maxy], qmvp, mvc, merange, qp)
        assert a == b, (case, w, h, a, b)
        moved += max(abs(a[0] - int(np.clip(qmvp[0], 4 * minx, 4 * maxx))), abs(a[1] - int(np.clip(qmvp[1], 4 * miny, 4 * maxy)))) >= 8
    assert moved > 25


def test_extend_pic_border(libs):
    """border extension of an uploaded picture (x265b200_plane_upload_picture) follows extendPicBorder, pixel.cpp:1044-1061"""
    o, r = libs
    rng = np.random.default_rng(77)
    for (w, h, mx, my) in ((96, 40, 24, 9), (352, 288, 96, 80), (33, 17, 5, 3)):
        stride, rows = w + 2 * mx + 8, h + 2 * my + 4
        base = rng.integers(0, o.pmax + 1, stride * rows).astype(o.pix)
        origin = my * stride + mx
        a, b = base.copy(), base.copy()
        o.extend_pic_border(a, origin, stride, w, h, mx, my)
        r.extend_pic_border(b, origin, stride, w, h, mx, my)
        assert np.array_equal(a, b), (w, h)
        assert not np.array_equal(a, base)


def test_intrinsic_dct_tier_equals_c(libs):
    """the SSE intrinsic transforms the CPU baseline may use (common/vec/dct-ssse3.cpp, dct-sse3.cpp) equal the C slots"""
    o, r = libs
    rng = np.random.default_rng(5)
    lim = 1 << (o.depth)
    for n in (8, 16, 32):
        for trial in range(4):
            src = rng.integers(-lim + 1, lim, n * n).astype(np.int16)
            assert np.array_equal(r.tier_dct(1, n, src, 0, n), r.tier_dct(0, n, src, 0, n)), n
            assert np.array_equal(r.tier_dct(1, n, src, 0, n), o.dct(n, src, 0, n)), n
            coef = o.dct(n, src, 0, n)
            assert np.array_equal(r.tier_idct(1, n, coef, n), r.tier_idct(0, n, coef, n)), n


def test_tu_forward_batch(libs):
    """sub_ps -> dct -> quant through the reference's slots (both tiers) vs the oracle's chain"""
    from frames import Geometry, make_plane, tile_blocks
    o, r = libs
    geo = Geometry(352, 288)
    F = make_plane(geo, o.depth, 11, "natural"); P = make_plane(geo, o.depth, 12, "natural")
    flat = [26214, 23302, 20560, 18396, 16384, 14564]
    for N, qp in ((32, 30), (16, 36), (8, 24), (4, 41)):
        offF, offP = tile_blocks(geo, N, N, seed=3, merange=9)
        per, rem = qp // 6, qp % 6
        tshift = 15 - o.depth - {4: 2, 8: 3, 16: 4, 32: 5}[N]
        qbits = 14 + per + tshift
        add = 171 << (qbits - 9)
        qc = np.full(N * N, flat[rem], np.int32)
        recon = np.zeros(geo.plane_elems, o.pix)
        rq, rns, _, _ = o.tu_chain_batch(N, F, geo.stride, P, geo.stride, offF, offP, qc, qbits, add, 40 << per, 20 - 14 - tshift if 20 - 14 - tshift > 0 else 1,
                                         recon, geo.stride, offF)
        for tier in (0, 1):
            r.set_tier(tier)
            lv, ns = r.tu_forward_batch(N, F, geo.stride, P, geo.stride, offF, offP, qc, qbits, add, nthreads=3)
            assert np.array_equal(lv, rq), (N, tier)
            assert np.array_equal(ns, rns), (N, tier)
        r.set_tier(0)
        assert (rns > 0).any()


def test_lookahead_mvp_and_bidir_costs(libs):
    """predictor selection (slicetype.cpp:4520-4558) and the bi-directional candidates (:4577-4596) of the lookahead: the oracle vs the
    reference's own lowresMC / bufSATD / pixelavg_pp driven with the reference's loop bodies"""
    o, r = libs
    from frames import Geometry
    geo = Geometry(192, 128)
    rng = np.random.default_rng(311)
    P0 = lowres_planes(geo, o.depth, 95); P1 = lowres_planes(geo, o.depth, 96)
    pitch = geo.plane_elems
    cw, ch = geo.coded()
    zero_pred = skip_set = 0
    for case in range(300):
        x = int(rng.integers(0, cw // 8)) * 8; y = int(rng.integers(0, ch // 8)) * 8
        of = geo.origin + y * geo.stride + x
        F = P0[(case % 4) * pitch:(case % 4 + 1) * pitch] if case % 3 == 0 else P1[:pitch]
        numc = int(rng.integers(0, 6))
        mvc = rng.integers(-40, 41, (numc, 2)).astype(np.int32)
        if numc and case % 4 == 0: mvc[int(rng.integers(0, numc))] = 0
        if numc and case % 7 == 0: mvc[0] = 0
        bid = case & 1
        a = o.lowres_mvp(F, of, geo.stride, P0, of, geo.stride, pitch, mvc, bid)
        b = r.lowres_mvp_ref(F, of, geo.stride, P0, of, geo.stride, pitch, mvc, bid)
        assert np.array_equal(a, b), (case, a, b)
        zero_pred += int(numc > 0 and a[0] == 0 and a[1] == 0); skip_set += int(a[3] != 0x7fffffff)
        mv0 = rng.integers(-40, 41, 2).astype(np.int32); mv1 = rng.integers(-40, 41, 2).astype(np.int32)
        a = o.lowres_bidir(F, of, geo.stride, P0, of, geo.stride, pitch, P1, of, geo.stride, pitch, mv0, mv1)
        b = r.lowres_bidir_ref(F, of, geo.stride, P0, geo.stride, pitch, P1, geo.stride, pitch, of, mv0, mv1)
        assert np.array_equal(a, b), (case, a, b)
    assert zero_pred > 5 and skip_set > 5
