"""Deterministic case list shared by tests/golden/make_golden.py (run against the reference's own
C primitives, oracle/_ref) and tests/test_oracle_golden.py (run against oracle/x265_oracle.c) and
tests/test_gpu_golden.py (run against the CUDA path through the C ABI).

`run_cases(lib, depth)` returns {name: ndarray}.  `lib` is any object with the cpulibs._Base
method set.  Inputs come from a fixed-seed PCG64 stream; an input checksum is stored with the
golden file so RNG drift would be detected rather than silently compared.
"""
import zlib

import numpy as np

from cpulibs import CHROMA_ONLY_420, CHROMA_ONLY_422, LUMA_PU, SATD_CHROMA_422


def _inputs(depth):
    rng = np.random.default_rng(0x265 + depth)
    pmax = (1 << depth) - 1
    dt = np.uint8 if depth == 8 else np.uint16
    n = 200 * 200
    d = {}
    d["pa"] = rng.integers(0, pmax + 1, n).astype(dt)
    d["pb"] = rng.integers(0, pmax + 1, n).astype(dt)
    d["pmin"] = np.zeros(n, dt)
    d["pmax"] = np.full(n, pmax, dt)
    d["res"] = rng.integers(-pmax, pmax + 1, n).astype(np.int16)          # residual range
    d["s4k"] = rng.integers(-4096, 4096, n).astype(np.int16)              # TestBench short_buff range
    d["full"] = rng.integers(-32768, 32768, n).astype(np.int16)           # never fed by TestBench
    d["sums"] = rng.integers(0, 1 << 22, n, dtype=np.int64).astype(np.uint32)
    d["cost"] = rng.integers(0, 3000, 256).astype(np.uint16)
    d["qc_small"] = rng.integers(-pmax, pmax + 1, 1024).astype(np.int32)
    d["qc_flat"] = np.full(1024, 26214, np.int32)
    d["dq"] = rng.integers(1, 1 << 12, 1024).astype(np.int32)
    return d


def input_checksum(depth):
    d = _inputs(depth)
    c = 0
    for k in sorted(d):
        c = zlib.crc32(d[k].tobytes(), c)
    return np.array([c], np.uint32)


def run_cases(lib, depth):
    I = _inputs(depth)
    out = {}
    pairs = [("rnd", I["pa"], I["pb"]), ("minmax", I["pmin"], I["pmax"]), ("maxmin", I["pmax"], I["pmin"])]

    # --- metrics (pixel.cpp) ---
    for tag, a, b in pairs:
        out["sad_" + tag] = np.array([lib.sad(w, h, a, 7, 64, b, 300 + 32 * i, 171)
                                      for i, (w, h) in enumerate(LUMA_PU)], np.int64)
        out["satd_" + tag] = np.array([lib.satd(w, h, a, 7, 64, b, 300 + 32 * i, 171)
                                       for i, (w, h) in enumerate(LUMA_PU)], np.int64)
        out["satd422_" + tag] = np.array([lib.satd(w, h, a, 7, 64, b, 301, 131) for (w, h) in SATD_CHROMA_422], np.int64)
        out["sa8d_" + tag] = np.array([lib.sa8d(w, w, a, 5, 64, b, 9, 150) for w in (4, 8, 16, 32, 64)], np.int64)
        out["sa8d422_" + tag] = np.array([lib.sa8d(w, 2 * w, a, 5, 64, b, 9, 150, chroma=2) for w in (4, 8, 16, 32)], np.int64)
        out["sse_pp_" + tag] = np.array([lib.sse_pp(w, w, a, 5, 64, b, 9, 150) for w in (4, 8, 16, 32, 64)], np.uint64)
        out["sad_x3_" + tag] = np.concatenate([lib.sad_x3(w, h, a, 64, b, [0, 1, 2], 59) for (w, h) in LUMA_PU])
        out["sad_x4_" + tag] = np.concatenate([lib.sad_x4(w, h, a, 64, b, [200, 7, 193, 1000], 131) for (w, h) in LUMA_PU])
    for tag, s in (("res", I["res"]), ("full", I["full"])):
        out["ssd_s_" + tag] = np.array([lib.ssd_s(w, s, 5, 66) for w in (4, 8, 16, 32, 64)], np.uint64)
    out["sse_ss_res"] = np.array([lib.sse_ss(w, I["res"], 3, 64, I["s4k"], 1, 70) for w in (4, 8, 16, 32, 64)], np.uint64)

    # --- ads (pixel.cpp:121-165; no TestBench coverage) ---
    ads_n, ads_m = [], []
    for i, (w, h) in enumerate(LUMA_PU):
        enc = (I["sums"][i * 4:i * 4 + 4] // 2).astype(np.int32)
        n, m = lib.ads(w, h, enc, I["sums"], 17 + i, (h >> 1) * 200, I["cost"], 64, 3_500_000)
        ads_n.append(n); ads_m.append(np.pad(m, (0, 64 - len(m)), constant_values=-1))
    out["ads_n"] = np.array(ads_n, np.int64)
    out["ads_mvs"] = np.concatenate(ads_m)

    # --- transforms (dct.cpp) ---
    for tag in ("res", "full"):
        s = I[tag]
        for n in (4, 8, 16, 32):
            out["dct%d_%s" % (n, tag)] = lib.dct(n, s, 11, 64)
            out["idct%d_%s" % (n, tag)] = lib.idct(n, s, 40)[:n * 40].copy()
        out["dst4_" + tag] = lib.dst4(s, 11, 64)
        out["idst4_" + tag] = lib.idst4(s, 9)
        for n in (8, 16, 32):
            out["lowpass%d_%s" % (n, tag)] = lib.lowpass_dct(n, s, 11, 64)
    for tag, coef, qc in (("small", I["res"], I["qc_small"]), ("flat", I["full"], I["qc_flat"])):
        for n, qbits in ((16, 9), (64, 14), (256, 17), (1024, 21)):
            ns, q, du = lib.quant(coef, qc, qbits, 171 << (qbits - 9), n)
            out["quant_%s_%d" % (tag, n)] = np.concatenate([[ns], q.astype(np.int64), du.astype(np.int64)])
            ns, q = lib.nquant(coef, qc, qbits, 1 << (qbits - 1), n)
            out["nquant_%s_%d" % (tag, n)] = np.concatenate([[ns], q.astype(np.int64)])
    out["dequant_normal"] = np.concatenate([lib.dequant_normal(I["full"], 1024, sc, sh)
                                            for sc, sh in ((40, 1), (72 << 4, 5), (64 << 8, 10))])
    out["dequant_scaling"] = np.concatenate([lib.dequant_scaling(I["full"], I["dq"], 1024, per, sh)
                                             for per, sh in ((0, 1), (3, 6), (12, 2), (8, 4))])

    # --- interpolation (ipfilter.cpp) ---
    def interp(kind, N, w, h, src, idx, extra=0):
        pix_out = kind in ("hpp", "vpp", "vsp", "hvpp")
        dst = np.full(80 * 100, 99, (np.uint8 if depth == 8 else np.uint16) if pix_out else np.int16)
        lib.interp(kind, N, w, h, src, 5 * 130 + 9, 130, dst, 3, 80, idx, extra)
        return dst[:80 * (h + 9)].astype(np.int64)       # rows beyond h+7 stay untouched

    luma_shapes = [(4, 4), (8, 8), (16, 16), (64, 64), (12, 16), (24, 32), (48, 64), (16, 4), (8, 32)]
    for (w, h) in luma_shapes:
        for idx in (1, 2, 3):
            key = "l%dx%d_%d" % (w, h, idx)
            out["hpp_" + key] = interp("hpp", 8, w, h, I["pa"], idx)
            out["vpp_" + key] = interp("vpp", 8, w, h, I["pa"], idx)
            out["hps_" + key] = interp("hps", 8, w, h, I["pa"], idx, idx & 1)
            out["vps_" + key] = interp("vps", 8, w, h, I["pa"], idx)
            out["vsp_" + key] = interp("vsp", 8, w, h, I["s4k"], idx)
            out["vss_" + key] = interp("vss", 8, w, h, I["full"], idx)
            out["hvpp_" + key] = interp("hvpp", 8, w, h, I["pb"], idx, 4 - idx)
    chroma_shapes = [(2, 4), (4, 2), (6, 8), (8, 6), (4, 4), (8, 8), (32, 32), (12, 32), (24, 64), (32, 48)]
    for (w, h) in chroma_shapes:
        for idx in (1, 4, 7):
            key = "c%dx%d_%d" % (w, h, idx)
            out["hpp_" + key] = interp("hpp", 4, w, h, I["pa"], idx)
            out["vpp_" + key] = interp("vpp", 4, w, h, I["pa"], idx)
            out["hps_" + key] = interp("hps", 4, w, h, I["pa"], idx, idx & 1)
            out["vps_" + key] = interp("vps", 4, w, h, I["pa"], idx)
            out["vsp_" + key] = interp("vsp", 4, w, h, I["s4k"], idx)
            out["vss_" + key] = interp("vss", 4, w, h, I["full"], idx)
    for (w, h) in luma_shapes + chroma_shapes:
        dst = np.full(80 * 100, 99, np.int16)
        lib.p2s(w, h, I["pa"], 77, 130, dst, 3, 80)
        out["p2s_%dx%d" % (w, h)] = dst[:80 * (h + 2)].astype(np.int64)
    return out
