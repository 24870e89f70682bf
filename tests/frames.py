"""Deterministic synthetic planes and block-descriptor arrays with x265's plane geometry
(reference geometry: source/common/picyuv.cpp:86-118)."""
import numpy as np


def splitmix64(idx, seed):
    z = (idx.astype(np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


class Geometry:
    """luma plane geometry of a w x h picture with CTU size `ctu` (picyuv.cpp:86-92)"""

    def __init__(self, width, height, ctu=64):
        self.width, self.height, self.ctu = width, height, ctu
        self.cu_w = (width + ctu - 1) // ctu
        self.cu_h = (height + ctu - 1) // ctu
        self.margin_x = ctu + 32
        self.margin_y = ctu + 16
        self.stride = self.cu_w * ctu + 2 * self.margin_x
        self.rows = self.cu_h * ctu + 2 * self.margin_y
        self.origin = self.margin_y * self.stride + self.margin_x     # element offset of pixel (0,0)
        self.plane_elems = self.stride * self.rows

    def coded(self):
        return self.cu_w * self.ctu, self.cu_h * self.ctu


class ChromaGeometry:
    """one chroma plane of the same picture (picyuv.cpp:108-118): the horizontal margin stays the luma one ("keep 16-byte alignment for
    chroma CTUs"), the vertical one and the coded size follow the chroma shifts (4:2:0: 1, 1; 4:2:2: 1, 0; 4:4:4: 0, 0)"""

    def __init__(self, luma, hshift, vshift):
        self.width, self.height, self.ctu = luma.width >> hshift, luma.height >> vshift, luma.ctu
        self.cw, self.ch = (luma.cu_w * luma.ctu) >> hshift, (luma.cu_h * luma.ctu) >> vshift
        self.margin_x = luma.margin_x
        self.margin_y = luma.margin_y >> vshift
        self.stride = self.cw + 2 * self.margin_x
        self.rows = self.ch + 2 * self.margin_y
        self.origin = self.margin_y * self.stride + self.margin_x
        self.plane_elems = self.stride * self.rows

    def coded(self):
        return self.cw, self.ch


def make_plane(geo, depth, seed, kind="uniform"):
    dt = np.uint8 if depth == 8 else np.uint16
    pmax = (1 << depth) - 1
    idx = np.arange(geo.plane_elems, dtype=np.uint64)
    if kind == "uniform":
        return (splitmix64(idx, seed) & np.uint64(pmax)).astype(dt)
    # separable base signal: one sin per column, one cos per row (same float64 values as evaluating it per sample)
    sx = np.sin(np.arange(geo.stride, dtype=np.float64) / 97.0)
    cy = np.cos(np.arange(geo.rows, dtype=np.float64) / 61.0)
    mid, amp = 1 << (depth - 1), 1 << (depth - 3)
    noise = ((splitmix64(idx, seed) & np.uint64(31)).astype(np.int64) - 16) * (1 << (depth - 8))
    v = (mid + amp * (sx[None, :] + cy[:, None])).ravel() + noise
    return np.clip(np.rint(v), 0, pmax).astype(dt)


def tile_blocks(geo, w, h, seed, merange=57):
    """every w x h block of a full tiling of the coded area, with one MV per block drawn uniformly in
    +-merange (clamped so the reference block stays inside the padded plane).
    returns (offA, offB) int32 element offsets from the plane base."""
    cw, ch = geo.coded()
    xs = np.arange(0, cw, w, dtype=np.int64)
    ys = np.arange(0, ch, h, dtype=np.int64)
    X, Y = np.meshgrid(xs, ys)
    X = X.ravel(); Y = Y.ravel()
    n = X.size
    r = splitmix64(np.arange(n, dtype=np.uint64), seed * 7919 + w * 131 + h)
    mvx = (r % np.uint64(2 * merange + 1)).astype(np.int64) - merange
    mvy = ((r >> np.uint64(20)) % np.uint64(2 * merange + 1)).astype(np.int64) - merange
    RX = np.clip(X + mvx, -geo.margin_x + 8, cw + geo.margin_x - w - 8)
    RY = np.clip(Y + mvy, -geo.margin_y + 8, ch + geo.margin_y - h - 8)
    offA = (geo.origin + Y * geo.stride + X).astype(np.int32)
    offB = (geo.origin + RY * geo.stride + RX).astype(np.int32)
    return offA, offB


def smooth_field(geo, depth, seed, box=9):
    """a box-blurred random field over the padded plane: smooth enough that pattern searches walk many steps towards a
    displaced copy, textured enough that the match is unique"""
    rng = np.random.default_rng(seed)
    f = rng.normal(0.0, 1.0, (geo.rows + box, geo.stride + box))
    c = np.cumsum(np.cumsum(f, 0), 1)
    b = c[box:, box:] - c[:-box, box:] - c[box:, :-box] + c[:-box, :-box]
    b = b[:geo.rows, :geo.stride]
    b = (b - b.min()) / (b.max() - b.min())
    dt = np.uint8 if depth == 8 else np.uint16
    return np.rint(b * ((1 << depth) - 1)).astype(dt).ravel()


def cu_descriptors(geo, S, d2N, d2NxN, dNx2N):
    """descriptors of x265b200_cu_satd_batch for every S x S CU of the coded area from the three per-shape descriptor sets
    (each an (offA, offB) pair of tile_blocks for S x S, S x S/2 and S/2 x S).  Returns (offF[n], offR[5n], idx) where
    idx[k] maps CU i to the block index of its k-th PU in the per-shape raster order, so that
    cost5.reshape(n, 5)[:, k] == per-shape cost[idx[k]]."""
    cw, ch = geo.coded()
    cx, cy = cw // S, ch // S
    X, Y = np.meshgrid(np.arange(cx), np.arange(cy))
    X = X.ravel(); Y = Y.ravel()
    i0 = Y * cx + X
    iH = [(2 * Y + k) * cx + X for k in (0, 1)]                 # S x S/2 blocks: cx per row, 2 * cy rows
    iV = [Y * (2 * cx) + 2 * X + k for k in (0, 1)]             # S/2 x S blocks: 2 * cx per row
    offF = d2N[0][i0]
    offR = np.stack([d2N[1][i0], d2NxN[1][iH[0]], d2NxN[1][iH[1]], dNx2N[1][iV[0]], dNx2N[1][iV[1]]], axis=1).astype(np.int32)
    return offF.astype(np.int32), offR.ravel(), [i0, iH[0], iH[1], iV[0], iV[1]]
