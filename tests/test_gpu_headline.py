"""The bench's own configuration, element-wise: every SATD cost of the 12 preset-slow PU shapes with seeded +-57 motion vectors
and every DCT coefficient of the four TU sizes at 3840x2160 10-bit (frames stacked as bench.py stacks them, so the
descriptors reach offsets near 3e8), plus 1920x1080 8-bit and 7680x4320 12-bit -- against the reference's own C primitives
(oracle/_ref; the oracle port where that library is absent).  Mirrors the reference's TestBench loop
(source/test/testbench.cpp:224-265, pixelharness.cpp:2649) at full-frame scale.  Also: host slots hammered from many threads."""
import ctypes as C
import threading

import numpy as np
import pytest

from cpulibs import OP_SAD, OP_SATD, Oracle, Reference, have_reference
from frames import Geometry, make_plane, tile_blocks

pytestmark = pytest.mark.gpu

SATD_SHAPES = [(64, 64), (64, 32), (32, 64), (32, 32), (32, 16), (16, 32), (16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8)]
DCT_SIZES = [32, 16, 8, 4]


def cpu_satd(depth, w, h, A, B, stride, oa, ob):
    if have_reference(depth):
        return Reference(depth).pixelcmp_batch(OP_SATD, w, h, A, stride, B, stride, oa, ob, 8)
    return Oracle(depth).pixelcmp_batch(OP_SATD, w, h, A, stride, B, stride, oa, ob)


def cpu_dct(depth, N, res, off):
    if have_reference(depth):
        return Reference(depth).dct_batch(N, res, N, off, 8)
    return Oracle(depth).dct_batch(N, res, N, off)


@pytest.mark.parametrize("depth,width,height,stack", [(10, 3840, 2160, 32), (8, 1920, 1080, 6), (12, 7680, 4320, 2)])
def test_headline_config_elementwise(depth, width, height, stack):
    import torch
    from gpulib import context, pkg
    ctx = context(depth); orc = Oracle(depth)
    geo = Geometry(width, height)
    pe = geo.plane_elems
    vt = np.int16 if depth > 8 else np.uint8
    tt = torch.int16 if depth > 8 else torch.uint8
    # `stack` frame slots as in bench.py; only the first and the last hold pictures (the others are never addressed)
    frames = sorted({0, stack - 1})
    dF = torch.zeros(stack * pe, dtype=tt, device="cuda"); dR = torch.zeros(stack * pe, dtype=tt, device="cuda")
    host = {}
    for f in frames:
        A = make_plane(geo, depth, 700 + f, "natural"); B = make_plane(geo, depth, 900 + f, "natural")
        host[f] = (A, B)
        dF[f * pe:(f + 1) * pe] = torch.from_numpy(A.view(vt)).cuda()
        dR[f * pe:(f + 1) * pe] = torch.from_numpy(B.view(vt)).cuda()
    assert (stack - 1) * pe + pe < 2 ** 31
    for (w, h) in SATD_SHAPES:
        oa, ob = tile_blocks(geo, w, h, seed=2)                # +-57 vectors, clamped to the padded plane
        assert (oa != ob).mean() > 0.9
        n = len(oa)
        a = np.concatenate([oa.astype(np.int64) + f * pe for f in frames]).astype(np.int32)
        b = np.concatenate([ob.astype(np.int64) + f * pe for f in frames]).astype(np.int32)
        out = torch.zeros(len(a), dtype=torch.int32, device="cuda")
        ctx.pixelcmp_batch(pkg.OP_SATD, w, h, dF, geo.stride, dR, geo.stride, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), out)
        got = out.cpu().numpy()
        for i, f in enumerate(frames):
            want = cpu_satd(depth, w, h, host[f][0], host[f][1], geo.stride, oa, ob)
            assert np.array_equal(got[i * n:(i + 1) * n], want), (depth, w, h, f)
    # DCT: residual of the 32x32 tiling (block-contiguous), re-read as N x N blocks for every size, exactly as bench.py does
    oa, ob = tile_blocks(geo, 32, 32, seed=2)
    n32 = len(oa)
    a = np.concatenate([oa.astype(np.int64) + f * pe for f in frames]).astype(np.int32)
    b = np.concatenate([ob.astype(np.int64) + f * pe for f in frames]).astype(np.int32)
    resid = torch.zeros(len(a) * 1024, dtype=torch.int16, device="cuda")
    ctx.residual_batch(32, 32, dF, geo.stride, dR, geo.stride, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), resid)
    res_np = {f: orc.residual_batch(32, 32, host[f][0], geo.stride, host[f][1], geo.stride, oa, ob) for f in frames}
    got_res = resid.cpu().numpy()
    per = n32 * 1024
    for i, f in enumerate(frames):
        assert np.array_equal(got_res[i * per:(i + 1) * per], res_np[f])
    coef = torch.zeros_like(resid)
    for N in DCT_SIZES:
        ctx.dct_batch(pkg.TR_DCT, N, resid, N, None, coef, count=resid.numel() // (N * N))
        got = coef.cpu().numpy()
        off = (np.arange(per // (N * N)) * N * N).astype(np.int32)
        for i, f in enumerate(frames):
            assert np.array_equal(got[i * per:(i + 1) * per], cpu_dct(depth, N, res_np[f], off)), (depth, N, f)
    ctx.check()


@pytest.mark.parametrize("depth", [8, 10])
def test_host_slots_from_many_threads(depth):
    """SURVEY 8b "Threading": the per-call slots are re-entrant from the encoder's worker threads (csrc/context.cu lane pool)"""
    from gpulib import context
    ctx = context(depth); orc = Oracle(depth)
    rng = np.random.default_rng(depth)
    NT, ITER = 16, 40
    A = rng.integers(0, orc.pmax + 1, 64 * 200).astype(orc.pix); B = rng.integers(0, orc.pmax + 1, 128 * 200).astype(orc.pix)
    R = rng.integers(-orc.pmax, orc.pmax + 1, 64 * 64).astype(np.int16)
    shapes = [(64, 64), (16, 16), (8, 4), (4, 8), (32, 24), (12, 16), (8, 8), (64, 16)]
    errors = []

    def worker(t):
        try:
            for it in range(ITER):
                w, h = shapes[(t + it) % len(shapes)]
                oa, ob = (t * 7 + it) % 50, (t * 13 + it * 3) % 60
                got = ctx.host.satd(w, h, A, oa, 64, B, ob, 128)
                want = orc.satd(w, h, A, oa, 64, B, ob, 128)
                if got != want:
                    errors.append(("satd", t, it, w, h, got, want))
                got = ctx.host.sad(w, h, A, oa, 64, B, ob, 128)
                if got != orc.sad(w, h, A, oa, 64, B, ob, 128):
                    errors.append(("sad", t, it, w, h))
                N = (4, 8, 16, 32)[(t + it) % 4]
                o = (t * 5 + it) % 30
                if not np.array_equal(ctx.host.dct(N, R, o, 64), orc.dct(N, R, o, 64)):
                    errors.append(("dct", t, it, N))
        except Exception as e:      # noqa: BLE001
            errors.append(("exception", t, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(NT)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]
    ctx.check()


def test_whole_plane_host_slots_at_1080p():
    """frameInitLowres and weight_pp are called by the encoder on whole planes (reference common/lowres.cpp:385,
    encoder/slicetype.cpp:880): the staging lane grows to fit them"""
    from gpulib import context
    depth = 8
    ctx = context(depth); orc = Oracle(depth)
    rng = np.random.default_rng(3)
    W, H = 960, 540                                  # lowres size of a 1080p source
    ss = 2 * W + 64
    src = rng.integers(0, 256, ss * (2 * H + 2)).astype(np.uint8)
    ds = W + 32
    outs = [np.zeros(ds * H, np.uint8) for _ in range(4)]
    ctx.lib.x265b200_frame_init_lowres(ctx.h, *[C.c_void_p(a.ctypes.data) for a in [src] + outs], C.c_ssize_t(ss), C.c_ssize_t(ds), W, H)
    ctx.check()
    want = [np.zeros(ds * H, np.uint8) for _ in range(4)]
    orc.lowres(src, 0, ss, want[0], want[1], want[2], want[3], ds, W, H)
    for g, w_ in zip(outs, want):
        assert np.array_equal(g, w_)
    # weight_pp over the padded lowres plane (stride x paddedLines, slicetype.cpp:880)
    stride, lines = W + 2 * 48, H + 2 * 40
    src = rng.integers(0, 256, stride * lines).astype(np.uint8)
    got = np.zeros(stride * lines, np.uint8); want = np.zeros(stride * lines, np.uint8)
    ctx.lib.x265b200_weight_pp(ctx.h, C.c_void_p(src.ctypes.data), C.c_void_p(got.ctypes.data), C.c_ssize_t(stride), stride, lines, 53, 32, 6, -3)
    ctx.check()
    orc.weight_pp(src, 0, want, 0, stride, stride, lines, 53, 32, 6, -3)
    assert np.array_equal(got, want)
