"""Batched DEVICE entries of the C ABI against the oracle on seeded full-plane inputs.

Element-wise diffs at sizes the oracle finishes in seconds (BASELINE config #2 geometry, 1080p),
size-independent properties at BASELINE's full 2160p10 size."""
import numpy as np
import pytest

from cpulibs import OP_SAD, OP_SATD, OP_SA8D, OP_SSE_PP, LUMA_PU, Oracle
from frames import Geometry, make_plane, tile_blocks

pytestmark = pytest.mark.gpu


def dev(a):
    import torch
    return torch.from_numpy(a).cuda()


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    return torch


@pytest.mark.parametrize("depth", [8, 10, 12])
def test_pixelcmp_frame_all_shapes(depth, torch_mod):
    torch = torch_mod
    from gpulib import context
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(640, 384)
    A = make_plane(geo, depth, 0x265, "natural"); B = make_plane(geo, depth, 0x266, "uniform")
    dA, dB = dev(A.view(np.int16 if depth > 8 else np.uint8)), dev(B.view(np.int16 if depth > 8 else np.uint8))
    for (w, h) in LUMA_PU:
        offA, offB = tile_blocks(geo, w, h, seed=3)
        dOA, dOB = dev(offA), dev(offB)
        ops = [OP_SAD, OP_SATD] + ([OP_SA8D, OP_SSE_PP] if w == h else [])
        for op in ops:
            out = torch.zeros(len(offA), dtype=torch.int64 if op == OP_SSE_PP else torch.int32, device="cuda")
            ctx.pixelcmp_batch(op, w, h, dA, geo.stride, dB, geo.stride, dOA, dOB, out)
            ref = orc.pixelcmp_batch(op, w, h, A, geo.stride, B, geo.stride, offA, offB)
            assert np.array_equal(out.cpu().numpy().astype(np.int64), ref.astype(np.int64)), (w, h, op)
    ctx.check()


@pytest.mark.parametrize("depth", [8, 10, 12])
def test_sad_multi_and_ragged(depth, torch_mod):
    torch = torch_mod
    from gpulib import context
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(256, 128)
    A = make_plane(geo, depth, 1); B = make_plane(geo, depth, 2)
    dA, dB = dev(A.view(np.int16 if depth > 8 else np.uint8)), dev(B.view(np.int16 if depth > 8 else np.uint8))
    for (w, h), K in (((16, 16), 4), ((8, 8), 3), ((64, 32), 9), ((12, 16), 5), ((8, 4), 4), ((16, 8), 3)):   # last two: strip kernel
        offA, offB0 = tile_blocks(geo, w, h, seed=5)
        n = len(offA)
        offR = np.stack([offB0 + k * 3 - (k % 2) * geo.stride for k in range(K)], axis=1).astype(np.int32).ravel()
        out = torch.zeros(n * K, dtype=torch.int32, device="cuda")
        ctx.sad_multi_batch(w, h, dA, geo.stride, dB, geo.stride, dev(offA), dev(offR), K, out)
        ref = orc.pixelcmp_batch(OP_SAD, w, h, A, geo.stride, B, geo.stride, np.repeat(offA, K).astype(np.int32), offR)
        assert np.array_equal(out.cpu().numpy(), ref)
    # empty batch is a no-op
    e = torch.zeros(0, dtype=torch.int32, device="cuda")
    ctx.pixelcmp_batch(OP_SATD, 8, 8, dA, geo.stride, dB, geo.stride, e, e, e)
    # n = 1 and n not a multiple of the lane-group packing
    for (w, h) in ((8, 4), (16, 8)):                     # strip kernel: ragged counts, SAD / SATD / SSE
        offA, offB = tile_blocks(geo, w, h, seed=11)
        for n in (1, 5, 130):
            for op, dt in ((OP_SAD, torch.int32), (OP_SATD, torch.int32), (3, torch.int64)):
                out = torch.zeros(n, dtype=dt, device="cuda")
                ctx.pixelcmp_batch(op, w, h, dA, geo.stride, dB, geo.stride, dev(offA[:n].copy()), dev(offB[:n].copy()), out)
                ref = orc.pixelcmp_batch(op, w, h, A, geo.stride, B, geo.stride, offA[:n].copy(), offB[:n].copy())
                assert np.array_equal(out.cpu().numpy().astype(np.uint64), np.asarray(ref).astype(np.uint64)), (w, h, n, op)
    for n in (1, 3, 33):
        offA, offB = tile_blocks(geo, 8, 8, seed=9)
        out = torch.zeros(n, dtype=torch.int32, device="cuda")
        ctx.pixelcmp_batch(OP_SATD, 8, 8, dA, geo.stride, dB, geo.stride, dev(offA[:n].copy()), dev(offB[:n].copy()), out)
        assert np.array_equal(out.cpu().numpy(), orc.pixelcmp_batch(OP_SATD, 8, 8, A, geo.stride, B, geo.stride, offA[:n].copy(), offB[:n].copy()))
    ctx.check()


@pytest.mark.parametrize("depth", [8, 10, 12])
def test_residual_dct_idct_frame(depth, torch_mod):
    torch = torch_mod
    from gpulib import context, pkg
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(320, 192)
    A = make_plane(geo, depth, 11, "natural"); B = make_plane(geo, depth, 12, "uniform")
    vt = np.int16 if depth > 8 else np.uint8
    dA, dB = dev(A.view(vt)), dev(B.view(vt))
    for N in (4, 8, 16, 32):
        offA, offB = tile_blocks(geo, N, N, seed=7)
        n = len(offA)
        res = torch.zeros(n * N * N, dtype=torch.int16, device="cuda")
        ctx.residual_batch(N, N, dA, geo.stride, dB, geo.stride, dev(offA), dev(offB), res)
        ref_res = orc.residual_batch(N, N, A, geo.stride, B, geo.stride, offA, offB)
        assert np.array_equal(res.cpu().numpy(), ref_res)
        off = (np.arange(n) * N * N).astype(np.int32)
        coef = torch.zeros(n * N * N, dtype=torch.int16, device="cuda")
        ctx.dct_batch(pkg.TR_DCT, N, res, N, dev(off), coef)
        ref_coef = orc.dct_batch(N, ref_res, N, off)
        assert np.array_equal(coef.cpu().numpy(), ref_coef), N
        # inverse into a strided plane (dstStride > N)
        stride = 40 * N
        rows = (n + 39) // 40
        plane = torch.zeros(rows * N * stride, dtype=torch.int16, device="cuda")
        offD = ((np.arange(n) // 40) * N * stride + (np.arange(n) % 40) * N).astype(np.int32)
        ctx.idct_batch(pkg.TR_DCT, N, coef, plane, stride, dev(offD))
        ref_plane = orc.idct_batch(N, ref_coef, np.zeros(rows * N * stride, np.int16), stride, offD)
        assert np.array_equal(plane.cpu().numpy(), ref_plane), N
        if N == 4:
            ctx.dct_batch(pkg.TR_DST, 4, res, 4, dev(off), coef)
            assert np.array_equal(coef.cpu().numpy(), orc.dct_batch(4, ref_res, 4, off, dst4=1))
    ctx.check()


def test_full_size_properties_2160p10(torch_mod):
    """BASELINE config #3 size: properties that do not need the oracle at full size."""
    torch = torch_mod
    from gpulib import context
    depth = 10
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(3840, 2160)
    A = make_plane(geo, depth, 21); B = make_plane(geo, depth, 22)
    dA, dB = dev(A.view(np.int16)), dev(B.view(np.int16))
    tot = {}
    for (w, h) in ((64, 64), (32, 32), (16, 16), (8, 8), (4, 4)):
        offA, _ = tile_blocks(geo, w, h, seed=1)
        offB = offA.copy()                                  # zero motion: every tiling covers the same samples
        dO = dev(offA)
        sad = torch.zeros(len(offA), dtype=torch.int32, device="cuda")
        sse = torch.zeros(len(offA), dtype=torch.int64, device="cuda")
        ctx.pixelcmp_batch(OP_SAD, w, h, dA, geo.stride, dB, geo.stride, dO, dO, sad)
        ctx.pixelcmp_batch(OP_SSE_PP, w, h, dA, geo.stride, dB, geo.stride, dO, dO, sse)
        tot[w] = (int(sad.sum(dtype=torch.int64)), int(sse.sum()))
        # identical blocks: every metric is exactly zero
        z = torch.ones(len(offA), dtype=torch.int32, device="cuda")
        ctx.pixelcmp_batch(OP_SATD, w, h, dA, geo.stride, dA, geo.stride, dO, dO, z)
        assert int(z.abs().max()) == 0
        # sampled element-wise check against the oracle
        sel = np.linspace(0, len(offA) - 1, 64).astype(np.int64)
        satd = torch.zeros(len(offA), dtype=torch.int32, device="cuda")
        ctx.pixelcmp_batch(OP_SATD, w, h, dA, geo.stride, dB, geo.stride, dO, dO, satd)
        ref = orc.pixelcmp_batch(OP_SATD, w, h, A, geo.stride, B, geo.stride, offA[sel].copy(), offB[sel].copy())
        assert np.array_equal(satd.cpu().numpy()[sel], ref)
    # SAD and SSE are additive over any tiling: totals agree across block sizes and with numpy
    assert len(set(tot.values())) == 1
    cw, ch = geo.coded()
    a2 = A.reshape(geo.rows, geo.stride)[geo.margin_y:geo.margin_y + ch, geo.margin_x:geo.margin_x + cw].astype(np.int64)
    b2 = B.reshape(geo.rows, geo.stride)[geo.margin_y:geo.margin_y + ch, geo.margin_x:geo.margin_x + cw].astype(np.int64)
    assert tot[64] == (int(np.abs(a2 - b2).sum()), int(((a2 - b2) ** 2).sum()))
    ctx.check()


@pytest.mark.parametrize("depth", [8, 10, 12])
def test_dct_tensor_core_and_butterfly_paths_full_range(depth, torch_mod):
    """forward DCT 16/32 on the IMMA path and on the CUDA-core twin, residual-range and full-range int16
    inputs (TestBench never feeds the latter), contiguous and strided / misaligned sources."""
    torch = torch_mod
    from gpulib import context, pkg
    ctx = context(depth); orc = Oracle(depth)
    rng = np.random.default_rng(99 + depth)
    pmax = (1 << depth) - 1
    n = 150
    try:
        for N in (16, 32, 8, 4):
            inputs = [rng.integers(-pmax, pmax + 1, n * N * N + 64).astype(np.int16),
                      rng.integers(-32768, 32768, n * N * N + 64).astype(np.int16),
                      np.full(n * N * N + 64, -32768, np.int16), np.full(n * N * N + 64, 32767, np.int16)]
            for src in inputs:
                d_src = dev(src)
                layouts = [(N, (np.arange(n) * N * N).astype(np.int32)),                 # contiguous TUs
                           (2 * N, (np.arange(n // 2) * 2 * N * N + 2).astype(np.int32)),  # strided, 4-byte aligned only
                           (2 * N + 4, (np.arange(n // 3) * 2 * N * N + 1).astype(np.int32))]  # odd offsets: 2-byte aligned
                for stride, off in layouts:
                    ref = orc.dct_batch(N, src, stride, off)
                    for path in (0, 1):
                        ctx.set_dct_path(path)
                        out = torch.zeros(len(off) * N * N, dtype=torch.int16, device="cuda")
                        ctx.dct_batch(pkg.TR_DCT, N, d_src, stride, dev(off), out)
                        assert np.array_equal(out.cpu().numpy(), ref), (N, stride, path)
                        if stride == N:         # implicit contiguous descriptors (off == NULL)
                            out.zero_()
                            ctx.dct_batch(pkg.TR_DCT, N, d_src, N, None, out, count=len(off))
                            assert np.array_equal(out.cpu().numpy(), ref), (N, "implicit", path)
    finally:
        ctx.set_dct_path(0)
    ctx.check()


@pytest.mark.parametrize("depth", [8, 10, 12])
def test_idct_tensor_core_and_butterfly_paths_full_range(depth, torch_mod):
    """inverse DCT 4..32 and IDST on the IMMA path and the CUDA-core twin; full-range int16 coefficients
    exercise both saturation stages (TestBench only feeds +-PIXEL_MAX)."""
    torch = torch_mod
    from gpulib import context, pkg
    ctx = context(depth); orc = Oracle(depth)
    rng = np.random.default_rng(199 + depth)
    pmax = (1 << depth) - 1
    n = 131
    try:
        for N in (32, 16, 8, 4):
            inputs = [rng.integers(-pmax, pmax + 1, n * N * N).astype(np.int16), rng.integers(-32768, 32768, n * N * N).astype(np.int16),
                      np.full(n * N * N, -32768, np.int16), np.full(n * N * N, 32767, np.int16)]
            for src in inputs:
                for stride, base in ((N, 0), (3 * N, 2), (3 * N + 1, 1)):
                    offD = (np.arange(n) * N * stride + base).astype(np.int32)
                    size = n * N * stride + 8
                    for kind, d4 in ((pkg.TR_DCT, 0),) + (((pkg.TR_DST, 1),) if N == 4 else ()):
                        ref = orc.idct_batch(N, src, np.full(size, 7, np.int16), stride, offD, dst4=d4)
                        for path in (0, 1):
                            ctx.set_dct_path(path)
                            plane = torch.full((size,), 7, dtype=torch.int16, device="cuda")
                            ctx.idct_batch(kind, N, dev(src), plane, stride, dev(offD))
                            assert np.array_equal(plane.cpu().numpy(), ref), (N, stride, kind, path)
    finally:
        ctx.set_dct_path(0)
    ctx.check()


@pytest.mark.parametrize("width,height,depth", [(1920, 1080, 8), (7680, 4320, 12)])
def test_full_size_properties_other_configs(width, height, depth, torch_mod):
    """BASELINE configs #2 (1080p 8-bit) and #5 (4320p 12-bit) at full size: additivity over tilings vs numpy,
    identical-block zeros, sampled oracle checks for SATD / sa8d, residual -> DCT -> IDCT round trip bounded error."""
    torch = torch_mod
    from gpulib import context, pkg
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(width, height)
    A = make_plane(geo, depth, 31, "natural"); B = make_plane(geo, depth, 32, "natural")
    vt = np.uint8 if depth == 8 else np.int16
    dA, dB = dev(A.view(vt)), dev(B.view(vt))
    tot = {}
    for (w, h) in ((64, 64), (16, 16), (8, 8)):
        offA, offMV = tile_blocks(geo, w, h, seed=2)
        dO, dM = dev(offA), dev(offMV)
        sad = torch.zeros(len(offA), dtype=torch.int32, device="cuda")
        sse = torch.zeros(len(offA), dtype=torch.int64, device="cuda")
        ctx.pixelcmp_batch(OP_SAD, w, h, dA, geo.stride, dB, geo.stride, dO, dO, sad)
        ctx.pixelcmp_batch(OP_SSE_PP, w, h, dA, geo.stride, dB, geo.stride, dO, dO, sse)
        tot[w] = (int(sad.sum(dtype=torch.int64)), int(sse.sum()))
        z = torch.ones(len(offA), dtype=torch.int32, device="cuda")
        ctx.pixelcmp_batch(OP_SA8D, w, h, dB, geo.stride, dB, geo.stride, dM, dM, z)
        assert int(z.abs().max()) == 0
        sel = np.linspace(0, len(offA) - 1, 48).astype(np.int64)
        for op in (OP_SATD, OP_SA8D):
            out = torch.zeros(len(offA), dtype=torch.int32, device="cuda")
            ctx.pixelcmp_batch(op, w, h, dA, geo.stride, dB, geo.stride, dO, dM, out)
            ref = orc.pixelcmp_batch(op, w, h, A, geo.stride, B, geo.stride, offA[sel].copy(), offMV[sel].copy())
            assert np.array_equal(out.cpu().numpy()[sel], ref), (w, op)
    assert len(set(tot.values())) == 1
    cw, ch = geo.coded()
    a2 = A.reshape(geo.rows, geo.stride)[geo.margin_y:geo.margin_y + ch, geo.margin_x:geo.margin_x + cw].astype(np.int64)
    b2 = B.reshape(geo.rows, geo.stride)[geo.margin_y:geo.margin_y + ch, geo.margin_x:geo.margin_x + cw].astype(np.int64)
    want_sse = int(((a2 - b2) ** 2).sum())
    if depth == 8:
        pass                                   # per-block values are exact; only the 64-bit total is compared
    assert tot[64] == (int(np.abs(a2 - b2).sum()), want_sse)
    # residual -> DCT -> IDCT: the HEVC core transform pair reconstructs the residual up to the rounding of its
    # 16-bit intermediate (coarser at higher depth, where the forward shifts grow); loose sanity bound here, the
    # exact values are compared with the oracle on a sample below
    for N in (32, 8):
        offA, offMV = tile_blocks(geo, N, N, seed=3)
        n = len(offA)
        res = torch.zeros(n * N * N, dtype=torch.int16, device="cuda")
        ctx.residual_batch(N, N, dA, geo.stride, dB, geo.stride, dev(offA), dev(offMV), res)
        coef = torch.zeros_like(res); rec = torch.zeros_like(res)
        ctx.dct_batch(pkg.TR_DCT, N, res, N, None, coef, count=n)
        off = torch.arange(n, dtype=torch.int32, device="cuda") * (N * N)
        ctx.idct_batch(pkg.TR_DCT, N, coef, rec, N, off)
        err = (rec.to(torch.int32) - res.to(torch.int32)).abs().max()
        assert int(err) <= (4 << (depth - 8)), (N, int(err))
        sel = np.linspace(0, n - 1, 16).astype(np.int64)
        rres = orc.residual_batch(N, N, A, geo.stride, B, geo.stride, offA[sel].copy(), offMV[sel].copy())
        rcoef = orc.dct_batch(N, rres, N, (np.arange(len(sel)) * N * N).astype(np.int32))
        got = coef.cpu().numpy().reshape(n, N * N)[sel].ravel()
        assert np.array_equal(got, rcoef), N
        rrec = orc.idct_batch(N, rcoef, np.zeros(len(sel) * N * N, np.int16), N, (np.arange(len(sel)) * N * N).astype(np.int32))
        assert np.array_equal(rec.cpu().numpy().reshape(n, N * N)[sel].ravel(), rrec), N
    ctx.check()


@pytest.mark.parametrize("depth", [8, 10, 12])
def test_cu_satd_batch_equals_per_shape_costs(depth, torch_mod):
    """x265b200_cu_satd_batch (2Nx2N + 2NxN + Nx2N PUs of a CU in one pass, five motion vectors per CU) vs the oracle's per-shape SATD"""
    torch = torch_mod
    from frames import cu_descriptors
    from gpulib import context
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(640, 384)
    A = make_plane(geo, depth, 31, "natural"); B = make_plane(geo, depth, 32, "uniform")
    vt = np.int16 if depth > 8 else np.uint8
    dA, dB = dev(A.view(vt)), dev(B.view(vt))
    for S in (8, 16, 32, 64):
        shapes = [(S, S), (S, S // 2), (S // 2, S)]
        d = [tile_blocks(geo, w, h, seed=13 + S) for (w, h) in shapes]
        offF, offR5, idx = cu_descriptors(geo, S, *d)
        n = len(offF)
        out = torch.full((5 * n,), -1, dtype=torch.int32, device="cuda")
        ctx.cu_satd_batch(S, dA, geo.stride, dB, geo.stride, dev(offF), dev(offR5), out)
        got = out.cpu().numpy().reshape(n, 5)
        want = [orc.pixelcmp_batch(OP_SATD, w, h, A, geo.stride, B, geo.stride, oa, ob) for (w, h), (oa, ob) in zip(shapes, d)]
        for k in range(5):
            assert np.array_equal(got[:, k], want[0 if k == 0 else 1 if k < 3 else 2][idx[k]]), (depth, S, k)
        # ragged count
        m = 7
        out2 = torch.full((5 * m,), -1, dtype=torch.int32, device="cuda")
        ctx.cu_satd_batch(S, dA, geo.stride, dB, geo.stride, dev(offF[:m].copy()), dev(offR5[:5 * m].copy()), out2)
        assert np.array_equal(out2.cpu().numpy(), got[:m].ravel())
    # extreme pictures (TestBench cases 1 and 2: all-max against all-min, and the reverse) and a checkerboard of the two: the largest
    # differences and transform sums a legal picture can produce (the 32 / 64 wide CUs run the f16 tensor-core Hadamard at depth <= 10)
    hi = np.full(geo.plane_elems, orc.pmax, orc.pix); lo = np.zeros(geo.plane_elems, orc.pix)
    chk = np.where((np.arange(geo.plane_elems) % geo.stride + np.arange(geo.plane_elems) // geo.stride) & 1, orc.pmax, 0).astype(orc.pix)
    for P, Q in ((hi, lo), (lo, hi), (chk, lo), (hi, chk), (chk, B)):
        dP, dQ = dev(P.view(vt)), dev(Q.view(vt))
        for S in (16, 32, 64):
            shapes = [(S, S), (S, S // 2), (S // 2, S)]
            d = [tile_blocks(geo, w, h, seed=40 + S) for (w, h) in shapes]
            offF, offR5, idx = cu_descriptors(geo, S, *d)
            out = torch.full((5 * len(offF),), -1, dtype=torch.int32, device="cuda")
            ctx.cu_satd_batch(S, dP, geo.stride, dQ, geo.stride, dev(offF), dev(offR5), out)
            got = out.cpu().numpy().reshape(-1, 5)
            want = [orc.pixelcmp_batch(OP_SATD, w, h, P, geo.stride, Q, geo.stride, oa, ob) for (w, h), (oa, ob) in zip(shapes, d)]
            for k in range(5):
                assert np.array_equal(got[:, k], want[0 if k == 0 else 1 if k < 3 else 2][idx[k]]), (depth, S, k, "extreme")
    ctx.check()
