"""Second golden family: the motion-search and intra entries.  `reference_outputs(ref, depth)` is run by
tests/golden/make_golden_me.py against the reference itself (MotionEstimate::motionEstimate and the intrapred.cpp slots
compiled into oracle/_ref); `oracle_outputs(orc, depth, gold)` recomputes the same cases with oracle/x265_oracle.c, taking
the lambda-scaled cost tables from the golden file (they are reference data, built by BitCost::setQP)."""
import zlib

import numpy as np

from frames import Geometry, make_plane, smooth_field

RAD = 4096
SHAPES = [(16, 16), (8, 8), (32, 32), (64, 64), (16, 8), (8, 16), (32, 24), (12, 16), (64, 32), (16, 64)]
QPS = [12, 22, 30, 37]


def _world(depth):
    geo = Geometry(192, 128)
    pmax = (1 << depth) - 1
    rng = np.random.default_rng(0x4d45 + depth)
    S = smooth_field(geo, depth, 11, box=9)
    R = np.clip(np.roll(S, -5 * geo.stride + 7).astype(np.int64) + rng.integers(-3, 4, S.size), 0, pmax).astype(S.dtype)
    N0 = make_plane(geo, depth, 12, "natural"); N1 = make_plane(geo, depth, 13, "natural")
    cases = []
    cw, ch = geo.coded()
    for i in range(60):
        w, h = SHAPES[i % len(SHAPES)]
        x = int(rng.integers(0, cw - w + 1)); y = int(rng.integers(0, ch - h + 1))
        m = int(rng.integers(4, 22))
        rngs = [-min(m, x + geo.margin_x - 12), -min(m, y + geo.margin_y - 12),
                min(m, cw + geo.margin_x - 12 - w - x), min(m, ch + geo.margin_y - 12 - h - y)]
        nc = int(rng.integers(0, 3))
        cases.append(dict(w=w, h=h, off=geo.origin + y * geo.stride + x, rng=rngs, qmvp=rng.integers(-4 * m, 4 * m + 1, 2),
                          mvc=rng.integers(-4 * m, 4 * m + 1, (nc, 2)), merange=int(rng.integers(4, 40)), subme=i % 8,
                          method=(1, 3, 0, 5, 2, 4)[i % 6], qp=QPS[i % 4], planes=(S, R) if i % 3 else (N0, N1)))
    nbs = {N: rng.integers(0, pmax + 1, (6, 4 * N + 1)).astype(S.dtype) for N in (4, 8, 16, 32)}
    return geo, cases, nbs, (S, R, N0, N1)


def input_checksum(depth):
    geo, cases, nbs, planes = _world(depth)
    c = 0
    for p in planes:
        c = zlib.crc32(p.tobytes(), c)
    for N in nbs:
        c = zlib.crc32(nbs[N].tobytes(), c)
    return np.array([c], np.uint32)


def reference_outputs(ref, depth):
    geo, cases, nbs, planes = _world(depth)
    out = {"__input_crc__": input_checksum(depth)}
    for qp in QPS:
        out["tab_%d" % qp] = ref.mvcost_table(qp, RAD)
    pitch = geo.plane_elems
    res = []
    for c in cases:
        F, R = c["planes"]
        rng = [c["rng"][0], c["rng"][1], c["rng"][2], c["rng"][3]]
        if c["method"] == 4:
            if c["w"] * c["h"] in (32 * 8, 8 * 4):      # never true for SHAPES; kept as a guard for the shapes SEA mis-reads
                continue
            sums = np.zeros(12 * pitch, np.uint32)
            ref.me_integral(R, geo.stride, geo.rows, sums, pitch)
            r = ref.motion_estimate_sea_ref(c["subme"], c["w"], c["h"], F, c["off"], geo.stride, R, c["off"], geo.stride, sums, pitch, rng,
                                            c["qmvp"], c["mvc"], c["merange"], c["qp"])
        else:
            r = ref.motion_estimate(c["method"], c["subme"], c["w"], c["h"], F, c["off"], geo.stride, R, c["off"], geo.stride, rng, c["qmvp"],
                                    c["mvc"], c["merange"], c["qp"])
        res.append(r)
    out["motion_estimate"] = np.array(res, np.int32)
    for N, arr in nbs.items():
        preds = []
        for s in arr:
            f = ref.intra_filter(N, s)
            one = [ref.intra_pred(N, 0, f if N >= 8 else s, 0), ref.intra_pred(N, 1, s, int(N <= 16))]
            flags = [0x38, 0x00] + ([0x38] + [0x30] * 6 + [0x20, 0x00, 0x20] + [0x30] * 6) * 2 + [0x38]
            for mode in range(2, 35):
                one.append(ref.intra_pred(N, mode, f if flags[mode] & N else s, int(N <= 16)))
            preds.append(np.concatenate(one))
        out["intra_pred_all_%d" % N] = np.concatenate(preds)
    S = planes[0]
    cw, ch = geo.coded()
    out["lowres_intra"] = np.array([ref.lowres_intra_cu(S, geo.origin, geo.stride, cx, cy, 29) for cy in range(ch // 8) for cx in range(cw // 8)],
                                   np.int32)
    return out


def oracle_outputs(orc, depth, gold):
    geo, cases, nbs, planes = _world(depth)
    out = {}
    pitch = geo.plane_elems
    res = []
    for c in cases:
        F, R = c["planes"]
        tab = gold["tab_%d" % c["qp"]]
        if c["method"] == 4:
            sums = np.zeros(12 * pitch, np.uint32)
            orc.me_integral(R, geo.stride, geo.rows, sums, pitch)
            r = orc.motion_estimate_sea(c["merange"], c["subme"], c["w"], c["h"], F, c["off"], geo.stride, R, c["off"], geo.stride, sums, pitch,
                                        c["rng"], c["qmvp"], c["mvc"], tab, RAD)
        else:
            r = orc.motion_estimate_full(c["subme"], c["w"], c["h"], F, c["off"], geo.stride, R, c["off"], geo.stride, c["rng"], c["qmvp"],
                                         c["mvc"], tab, RAD, c["method"], c["merange"])
        res.append(r)
    out["motion_estimate"] = np.array(res, np.int32)
    for N, arr in nbs.items():
        out["intra_pred_all_%d" % N] = np.concatenate([orc.intra_pred_all(N, s) for s in arr])
    S = planes[0]
    cw, ch = geo.coded()
    out["lowres_intra"] = np.array([orc.lowres_intra_cu(S, geo.origin, geo.stride, cx, cy, 29) for cy in range(ch // 8) for cx in range(cw // 8)],
                                   np.int32)
    return out
