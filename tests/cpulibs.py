"""ctypes handles for the two CPU checkers (test infrastructure only).

Oracle     -> oracle/lib/liboracle_<depth>.so   (our C restatement; built on demand with gcc)
Reference  -> oracle/_ref/libx265ref_<depth>.so (the reference's own C primitives; prebuilt in
              the authoring container, shipped to the GPU box, never rebuilt there)

Both expose the same Python methods so a test can run one against the other.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

OP_SAD, OP_SATD, OP_SA8D, OP_SSE_PP = 0, 1, 2, 3


def pixel_dtype(depth):
    return np.uint8 if depth == 8 else np.uint16


def _ptr(arr, off=0):
    """address of element `off` (in elements) of a numpy array"""
    return C.c_void_p(arr.ctypes.data + int(off) * arr.itemsize)


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)


def have_reference(depth=10):
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libx265ref_%d.so" % depth))


class _Base:
    prefix = ""

    def __init__(self, depth):
        self.depth = depth
        self.pix = pixel_dtype(depth)
        self.pmax = (1 << depth) - 1

    def _f(self, name, restype=C.c_int):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        return f

    # ---- metrics -------------------------------------------------------------------
    def sad(self, w, h, a, oa, sa, b, ob, sb):
        return self._f("sad")(w, h, _ptr(a, oa), C.c_ssize_t(sa), _ptr(b, ob), C.c_ssize_t(sb))

    def satd(self, w, h, a, oa, sa, b, ob, sb):
        return self._f("satd")(w, h, _ptr(a, oa), C.c_ssize_t(sa), _ptr(b, ob), C.c_ssize_t(sb))

    def sad_x3(self, w, h, fenc, of, ref, offs, rs):
        res = np.zeros(3, np.int32)
        self._f("sad_x3", None)(w, h, _ptr(fenc, of), _ptr(ref, offs[0]), _ptr(ref, offs[1]),
                                _ptr(ref, offs[2]), C.c_ssize_t(rs), _ptr(res))
        return res

    def sad_x4(self, w, h, fenc, of, ref, offs, rs):
        res = np.zeros(4, np.int32)
        self._f("sad_x4", None)(w, h, _ptr(fenc, of), _ptr(ref, offs[0]), _ptr(ref, offs[1]),
                                _ptr(ref, offs[2]), _ptr(ref, offs[3]), C.c_ssize_t(rs), _ptr(res))
        return res

    def ads(self, w, h, encDC, sums, osum, delta, cost, width, thresh):
        enc = np.ascontiguousarray(encDC, np.int32)
        mvs = np.full(width + 8, -1, np.int16)
        n = self._f("ads")(w, h, _ptr(enc), _ptr(sums, osum), int(delta), _ptr(cost), _ptr(mvs),
                           int(width), int(thresh))
        return n, mvs[:n].copy()

    def ssd_s(self, size, a, oa, sa):
        return int(self._f("ssd_s", C.c_uint64)(size, _ptr(a, oa), C.c_ssize_t(sa)))

    # ---- transforms ------------------------------------------------------------------
    def dct(self, n, src, osrc, stride):
        dst = np.zeros(n * n, np.int16)
        self._f("dct", None)(n, _ptr(src, osrc), _ptr(dst), C.c_ssize_t(stride))
        return dst

    def idct(self, n, src, stride, dst=None, odst=0):
        if dst is None:
            dst = np.zeros(n * stride, np.int16)
        self._f("idct", None)(n, _ptr(src), _ptr(dst, odst), C.c_ssize_t(stride))
        return dst

    def dst4(self, src, osrc, stride):
        dst = np.zeros(16, np.int16)
        self._f("dst4", None)(_ptr(src, osrc), _ptr(dst), C.c_ssize_t(stride))
        return dst

    def idst4(self, src, stride):
        dst = np.zeros(4 * stride, np.int16)
        self._f("idst4", None)(_ptr(src), _ptr(dst), C.c_ssize_t(stride))
        return dst

    def lowpass_dct(self, n, src, osrc, stride):
        dst = np.full(n * n, 0x5a5a, np.int16)
        self._f("lowpass_dct", None)(n, _ptr(src, osrc), _ptr(dst), C.c_ssize_t(stride))
        return dst

    def quant(self, coef, qc, qbits, add, n):
        deltaU = np.zeros(n, np.int32)
        q = np.zeros(n, np.int16)
        r = self._f("quant", C.c_uint32)(_ptr(coef), _ptr(qc), _ptr(deltaU), _ptr(q), qbits, add, n)
        return int(r), q, deltaU

    def nquant(self, coef, qc, qbits, add, n):
        q = np.zeros(n, np.int16)
        r = self._f("nquant", C.c_uint32)(_ptr(coef), _ptr(qc), _ptr(q), qbits, add, n)
        return int(r), q

    def dequant_normal(self, q, n, scale, shift):
        out = np.zeros(n, np.int16)
        self._f("dequant_normal", None)(_ptr(q), _ptr(out), n, scale, shift)
        return out

    def dequant_scaling(self, q, dq, n, per, shift):
        out = np.zeros(n, np.int16)
        self._f("dequant_scaling", None)(_ptr(q), _ptr(dq), _ptr(out), n, per, shift)
        return out

    # ---- interpolation: every call writes into a caller buffer `dst` at `od` ------------
    def interp(self, kind, N, w, h, src, os_, ss, dst, od, ds, idx, idy_or_ext=0):
        a = (N, w, h, _ptr(src, os_), C.c_ssize_t(ss), _ptr(dst, od), C.c_ssize_t(ds), idx)
        if kind in ("hps", "hvpp"):
            a = a + (idy_or_ext,)
        return self._f("interp_" + kind)(*a)

    def p2s(self, w, h, src, os_, ss, dst, od, ds):
        return self._f("p2s")(w, h, _ptr(src, os_), C.c_ssize_t(ss), _ptr(dst, od), C.c_ssize_t(ds))

    # ---- per-block scalars -------------------------------------------------------------------------------------
    def var(self, size, pix, op_, stride):
        return int(self._f("var", C.c_uint64)(size, _ptr(pix, op_), C.c_ssize_t(stride)))

    def psy_cost_pp(self, size, a, oa, sa, b, ob, sb):
        return self._f("psy_cost_pp")(size, _ptr(a, oa), C.c_ssize_t(sa), _ptr(b, ob), C.c_ssize_t(sb))

    def copy_cnt(self, size, coeff, resi, or_, stride):
        return int(self._f("copy_cnt", C.c_uint32)(size, _ptr(coeff) if coeff is not None else None, _ptr(resi, or_), C.c_ssize_t(stride)))

    def denoise_dct(self, dct, res_sum, offset, num):
        self._f("denoise_dct", None)(_ptr(dct), _ptr(res_sum), _ptr(offset), num)

    # ---- copy family (kind 0..6, see include/x265b200.h x265b200_blockcopy_batch) -------------------------------
    def blockcopy(self, kind, w, h, dst, od, ds, src, os_, ss, param=0):
        return self._f("blockcopy")(kind, w, h, _ptr(dst, od), C.c_ssize_t(ds), _ptr(src, os_) if src is not None else None, C.c_ssize_t(ss), param)

    # ---- weighted prediction + lookahead weight cost -----------------------------------------------------------
    def weight_pp(self, src, osrc, dst, odst, stride, width, height, w0, rnd, shift, offset):
        self._f("weight_pp", None)(_ptr(src, osrc), _ptr(dst, odst), C.c_ssize_t(stride), width, height, w0, rnd, shift, offset)

    def weight_sp(self, src, osrc, dst, odst, ss, ds, width, height, w0, rnd, shift, offset):
        self._f("weight_sp", None)(_ptr(src, osrc), _ptr(dst, odst), C.c_ssize_t(ss), C.c_ssize_t(ds), width, height, w0, rnd, shift, offset)

    def weight_cost(self, fenc, of, ref, orf, stride, width, height, intra, weights):
        K = len(weights) // 4
        cost = np.zeros(K, np.uint32)
        tmp = np.zeros(len(ref), ref.dtype)
        self._f("weight_cost", None)(_ptr(fenc, of), _ptr(ref, orf), C.c_ssize_t(stride), width, height, _ptr(intra) if intra is not None else None,
                                     _ptr(weights), K, _ptr(cost), _ptr(tmp))
        return cost

    # ---- SEA integral planes ---------------------------------------------------------------------------------
    def integral_inith(self, W, sum_, osum, pix, opix, stride):
        return self._f("integral_inith")(W, _ptr(sum_, osum), _ptr(pix, opix), C.c_ssize_t(stride))

    def integral_initv(self, H, sum_, osum, stride):
        return self._f("integral_initv")(H, _ptr(sum_, osum), C.c_ssize_t(stride))

    def me_integral(self, pix, stride, rows, sums, plane_pitch):
        return self._f("me_integral")(_ptr(pix), C.c_ssize_t(stride), rows, _ptr(sums), C.c_size_t(plane_pitch))

    # ---- exhaustive integer motion search of one PU (motion.cpp X265_FULL_SEARCH) -----------------------------
    def me_full_search(self, w, h, fenc, of, sf, ref, orf, sr, rng, mvp, cost_tab, centre, bmv, bcost):
        """rng = [minx, miny, maxx, maxy] full pel, mvp qpel (int32 arrays); cost_tab uint16 with its centre at `centre`;
        returns (bmv_x, bmv_y, bcost)"""
        mv = np.array(bmv, np.int32); bc = C.c_int32(int(bcost))
        rng = np.ascontiguousarray(rng, np.int32); mvp = np.ascontiguousarray(mvp, np.int32)
        self._f("me_full_search", None)(w, h, _ptr(fenc, of), C.c_ssize_t(sf), _ptr(ref, orf), C.c_ssize_t(sr), _ptr(rng), _ptr(mvp),
                                        _ptr(cost_tab, centre), _ptr(mv), C.byref(bc))
        return int(mv[0]), int(mv[1]), int(bc.value)

    # ---- MotionEstimate::motionEstimate, full search + sub-pel refinement ---------------------------------------
    def motion_estimate_full(self, subme, w, h, fenc, of, sf, ref, orf, sr, rng, qmvp, mvc, cost_tab, centre, method=5, merange=0):
        """oracle restatement (method: 0 DIA, 1 HEX, 5 FULL); returns (qmv_x, qmv_y, cost)"""
        out = np.zeros(2, np.int32)
        rng = np.ascontiguousarray(rng, np.int32); qmvp = np.ascontiguousarray(qmvp, np.int32)
        mvc = np.ascontiguousarray(mvc, np.int32).reshape(-1)
        c = self._f("motion_estimate")(method, merange, subme, w, h, _ptr(fenc, of), C.c_ssize_t(sf), _ptr(ref, orf), C.c_ssize_t(sr), _ptr(rng), _ptr(qmvp),
                                            len(mvc) // 2, _ptr(mvc), _ptr(cost_tab, centre), _ptr(out))
        return int(out[0]), int(out[1]), int(c)

    def motion_estimate_chroma(self, method, merange, subme, w, h, fY, ofY, sfY, rY, orY, srY, fCb, fCr, ofC, sfC, rCb, rCr, orC, srC,
                               hshift, vshift, rng, qmvp, mvc, cost_tab, centre):
        """oracle: motionEstimate with the chroma SATD term of subme > 2 (co-located luma / Cb / Cr blocks at the given offsets)"""
        out = np.zeros(2, np.int32)
        rng = np.ascontiguousarray(rng, np.int32); qmvp = np.ascontiguousarray(qmvp, np.int32)
        mvc = np.ascontiguousarray(mvc, np.int32).reshape(-1)
        c = self._f("motion_estimate_chroma")(method, merange, subme, w, h, _ptr(fY, ofY), C.c_ssize_t(sfY), _ptr(rY, orY), C.c_ssize_t(srY),
                                              _ptr(fCb, ofC), _ptr(fCr, ofC), C.c_ssize_t(sfC), _ptr(rCb, orC), _ptr(rCr, orC), C.c_ssize_t(srC),
                                              hshift, vshift, _ptr(rng), _ptr(qmvp), len(mvc) // 2, _ptr(mvc), _ptr(cost_tab, centre), _ptr(out))
        return int(out[0]), int(out[1]), int(c)

    # ---- intra prediction (intrapred.cpp) and the lookahead's intra estimate (slicetype.cpp lowresIntraEstimate) ---------
    def intra_filter(self, N, samples):
        out = np.zeros(4 * N + 1, samples.dtype)
        self._f("intra_filter", None)(N, _ptr(samples), _ptr(out))
        return out

    def intra_pred(self, N, mode, samples, bFilter):
        out = np.zeros(N * N, samples.dtype)
        self._f("intra_pred", None)(N, mode, _ptr(samples), int(bFilter), _ptr(out), C.c_ssize_t(N))
        return out

    def intra_pred_all(self, N, samples):
        out = np.zeros(35 * N * N, samples.dtype)
        self._f("intra_pred_all", None)(N, _ptr(samples), _ptr(out))
        return out

    def lowres_intra_cu(self, plane, origin, stride, cuX, cuY, penalty):
        m = C.c_int32(0)
        c = self._f("lowres_intra_cu")(_ptr(plane, origin), C.c_ssize_t(stride), cuX, cuY, penalty, C.byref(m))
        return int(c), int(m.value)

    def bidir_satd(self, w, h, fenc, of, sf, ref0, o0, sr0, frac0, ref1, o1, sr1, frac1):
        return self._f("bidir_satd")(w, h, _ptr(fenc, of), C.c_ssize_t(sf), _ptr(ref0, o0), C.c_ssize_t(sr0), frac0, _ptr(ref1, o1), C.c_ssize_t(sr1), frac1)

    def subpel_cmp_chroma(self, w, h, fenc, of, sf, ref, orf, sr, xFrac, yFrac):
        return self._f("subpel_cmp_chroma")(w, h, _ptr(fenc, of), C.c_ssize_t(sf), _ptr(ref, orf), C.c_ssize_t(sr), xFrac, yFrac)

    def motion_estimate_sea(self, merange, subme, w, h, fenc, of, sf, ref, orf, sr, sums, pitch, rng, qmvp, mvc, cost_tab, centre):
        """oracle: motionEstimate with X265_SEA; sums = 12 integral planes over the padded reference buffer, `pitch` apart"""
        out = np.zeros(2, np.int32)
        rng = np.ascontiguousarray(rng, np.int32); qmvp = np.ascontiguousarray(qmvp, np.int32)
        mvc = np.ascontiguousarray(mvc, np.int32).reshape(-1)
        c = self._f("motion_estimate_sea")(merange, subme, w, h, _ptr(fenc, of), C.c_ssize_t(sf), _ptr(ref, orf), C.c_ssize_t(sr), _ptr(sums),
                                           C.c_size_t(pitch), C.c_ssize_t(orf), _ptr(rng), _ptr(qmvp), len(mvc) // 2, _ptr(mvc),
                                           _ptr(cost_tab, centre), _ptr(out))
        return int(out[0]), int(out[1]), int(c)

    def lowres_motion_estimate(self, method, merange, subme, w, h, fenc, of, sf, planes, orf, sr, pitch, rng, qmvp, cost_tab, centre):
        """oracle: the lookahead's motionEstimate on a lowres reference (four half-pel planes `pitch` apart)"""
        out = np.zeros(2, np.int32)
        rng = np.ascontiguousarray(rng, np.int32); qmvp = np.ascontiguousarray(qmvp, np.int32)
        c = self._f("lowres_motion_estimate")(method, merange, subme, w, h, _ptr(fenc, of), C.c_ssize_t(sf), _ptr(planes, orf), C.c_ssize_t(sr),
                                              C.c_size_t(pitch), _ptr(rng), _ptr(qmvp), _ptr(cost_tab, centre), _ptr(out))
        return int(out[0]), int(out[1]), int(c)

    # ---- sub-pel candidate cost (subpelCompare): interpolation + sad (op 0) / satd (op 1) ----------------------
    def subpel_cmp(self, op, w, h, fenc, of, sf, ref, orf, sr, xFrac, yFrac):
        return self._f("subpel_cmp")(op, w, h, _ptr(fenc, of), C.c_ssize_t(sf), _ptr(ref, orf), C.c_ssize_t(sr), xFrac, yFrac)

    # ---- adjacent slots: sub_ps / add_ps / pixelavg_pp / addAvg (op 0..3) and the lowres downscale -------------
    def blockop(self, op, w, h, A, oa, sa, B, ob, sb, D, od, sd):
        return self._f("blockop")(op, w, h, _ptr(A, oa), C.c_ssize_t(sa), _ptr(B, ob), C.c_ssize_t(sb), _ptr(D, od), C.c_ssize_t(sd))

    def lowres(self, src, os_, ss, d0, dh, dv, dc, ds, width, height):
        self._f("lowres", None)(_ptr(src, os_), C.c_ssize_t(ss), _ptr(d0), _ptr(dh), _ptr(dv), _ptr(dc), C.c_ssize_t(ds), width, height)

    # ---- inter luma TU chain (sub_ps, dct, quant, dequant, DC shortcut / idct, add_ps, sse) ---------------
    def tu_chain(self, N, fenc, of, sf, pred, op_, sp, qc, qbits, add, dqscale, dqshift, recon, orr, sr, ttype=0):
        """ttype 0 = inter luma / any chroma TU, 1 = intra luma (DST-VII and no DC-only shortcut at N = 4)"""
        q = np.zeros(N * N, np.int16)
        ns = C.c_uint32(0); z = C.c_uint64(0); r = C.c_uint64(0)
        self._f("tu_chain_tt", None)(N, ttype, _ptr(fenc, of), C.c_ssize_t(sf), _ptr(pred, op_), C.c_ssize_t(sp), _ptr(qc), qbits, add,
                                     dqscale, dqshift, _ptr(q), C.byref(ns), _ptr(recon, orr), C.c_ssize_t(sr), C.byref(z), C.byref(r))
        return q, ns.value, z.value, r.value

    # ---- tables -----------------------------------------------------------------------
    def lowres_mvp(self, fenc, of, sf, planes, orf, sr, pitch, mvc, bidir):
        """oracle: predictor selection of the lookahead for one 8x8 CU -> (mvpx, mvpy, mvpCost, skipCost)"""
        mvc = np.ascontiguousarray(mvc, np.int32).reshape(-1)
        out = np.zeros(4, np.int32)
        self._f("lowres_mvp", None)(_ptr(fenc, of), C.c_ssize_t(sf), _ptr(planes, orf), C.c_ssize_t(sr), C.c_size_t(pitch), _ptr(mvc), len(mvc) // 2, int(bidir), _ptr(out))
        return out

    def lowres_bidir(self, fenc, of, sf, planes0, o0, s0, pitch0, planes1, o1, s1, pitch1, mv0, mv1):
        out = np.zeros(2, np.int32)
        mv0 = np.ascontiguousarray(mv0, np.int32); mv1 = np.ascontiguousarray(mv1, np.int32)
        self._f("lowres_bidir", None)(_ptr(fenc, of), C.c_ssize_t(sf), _ptr(planes0, o0), C.c_ssize_t(s0), C.c_size_t(pitch0), _ptr(planes1, o1), C.c_ssize_t(s1),
                                     C.c_size_t(pitch1), _ptr(mv0), _ptr(mv1), _ptr(out))
        return out

    def extend_pic_border(self, plane, origin, stride, width, height, marginX, marginY):
        """in place on `plane` (numpy, pixel dtype); origin = element offset of sample (0, 0)"""
        self._f("extend_pic_border", None)(_ptr(plane, origin), C.c_ssize_t(stride), width, height, marginX, marginY)

    def dct_matrix(self, n):
        out = np.zeros(n * n, np.int16)
        self._f("get_dct_matrix", None)(n, _ptr(out))
        return out.reshape(n, n)

    def luma_taps(self):
        out = np.zeros(32, np.int16)
        self._f("get_luma_taps", None)(_ptr(out))
        return out.reshape(4, 8)

    def chroma_taps(self):
        out = np.zeros(32, np.int16)
        self._f("get_chroma_taps", None)(_ptr(out))
        return out.reshape(8, 4)


class Oracle(_Base):
    prefix = "orc_"

    def __init__(self, depth):
        super().__init__(depth)
        path = os.path.join(ORACLE_DIR, "lib", "liboracle_%d.so" % depth)
        src = os.path.join(ORACLE_DIR, "x265_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build_oracle()
        self.lib = C.CDLL(path)
        assert self.lib.orc_depth() == depth

    def sa8d(self, w, h, a, oa, sa, b, ob, sb, chroma=0):
        return self._f("sa8d")(w, h, _ptr(a, oa), C.c_ssize_t(sa), _ptr(b, ob), C.c_ssize_t(sb))

    def sse_pp(self, w, h, a, oa, sa, b, ob, sb, chroma=0):
        rt = C.c_uint32 if self.depth == 8 else C.c_uint64
        return int(self._f("sse_pp", rt)(w, h, _ptr(a, oa), C.c_ssize_t(sa), _ptr(b, ob), C.c_ssize_t(sb)))

    def sse_ss(self, w, a, oa, sa, b, ob, sb):
        rt = C.c_uint32 if self.depth == 8 else C.c_uint64
        return int(self._f("sse_ss", rt)(w, w, _ptr(a, oa), C.c_ssize_t(sa), _ptr(b, ob), C.c_ssize_t(sb)))

    def ssd_s(self, size, a, oa, sa):
        rt = C.c_uint32 if self.depth == 8 else C.c_uint64
        return int(self._f("ssd_s", rt)(size, _ptr(a, oa), C.c_ssize_t(sa)))

    def pixelcmp_batch(self, op, w, h, A, sa, B, sb, offA, offB, nthreads=1):
        n = len(offA)
        out = np.zeros(n, np.uint64 if op == OP_SSE_PP else np.int32)
        self._f("pixelcmp_batch", None)(op, w, h, _ptr(A), C.c_ssize_t(sa), _ptr(B), C.c_ssize_t(sb),
                                        _ptr(offA), _ptr(offB), n, _ptr(out))
        return out

    def residual_batch(self, w, h, A, sa, B, sb, offA, offB):
        n = len(offA)
        out = np.zeros(n * w * h, np.int16)
        self._f("residual_batch", None)(w, h, _ptr(A), C.c_ssize_t(sa), _ptr(B), C.c_ssize_t(sb),
                                        _ptr(offA), _ptr(offB), n, _ptr(out))
        return out

    def dct_batch(self, N, src, stride, off, dst4=0):
        n = len(off)
        out = np.zeros(n * N * N, np.int16)
        self._f("dct_batch", None)(N, dst4, _ptr(src), C.c_ssize_t(stride), _ptr(off), n, _ptr(out))
        return out

    def tu_chain_batch(self, N, fenc, sf, pred, sp, offF, offP, qc, qbits, add, dqscale, dqshift, recon, sr, offR, ttype=0):
        n = len(offF)
        q = np.zeros(n * N * N, np.int16); ns = np.zeros(n, np.uint32); z = np.zeros(n, np.uint64); r = np.zeros(n, np.uint64)
        self._f("tu_chain_tt_batch", None)(N, ttype, _ptr(fenc), C.c_ssize_t(sf), _ptr(pred), C.c_ssize_t(sp), _ptr(offF), _ptr(offP), n,
                                           _ptr(qc), qbits, add, dqscale, dqshift, _ptr(q), _ptr(ns), _ptr(recon), C.c_ssize_t(sr),
                                           _ptr(offR), _ptr(z), _ptr(r))
        return q, ns, z, r

    def blockcopy_batch(self, kind, w, h, S, ss, offS, D, sd, offD, param=0):
        self._f("blockcopy_batch", None)(kind, w, h, _ptr(S) if S is not None else None, C.c_ssize_t(ss), _ptr(offS), _ptr(D), C.c_ssize_t(sd), _ptr(offD), len(offD), param)
        return D

    def subpel_cmp_batch(self, op, w, h, fenc, sf, ref, sr, offF, offR, frac, K):
        n = len(offF)
        cost = np.zeros(n * K, np.int32)
        self._f("subpel_cmp_batch", None)(op, w, h, _ptr(fenc), C.c_ssize_t(sf), _ptr(ref), C.c_ssize_t(sr), _ptr(offF), _ptr(offR), _ptr(frac), K, n, _ptr(cost))
        return cost

    def me_full_batch(self, w, h, fenc, sf, ref, sr, offF, offR, rng, mvp, cost_tab, centre, bmv, bcost):
        """bmv (n x 2 int32) and bcost (n int32) are updated in place"""
        self._f("me_full_batch", None)(w, h, _ptr(fenc), C.c_ssize_t(sf), _ptr(ref), C.c_ssize_t(sr), _ptr(offF), _ptr(offR), _ptr(rng), _ptr(mvp),
                                       _ptr(cost_tab, centre), len(offF), _ptr(bmv), _ptr(bcost))

    def bidir_satd_batch(self, w, h, fenc, sf, offF, ref0, sr0, off0, frac0, ref1, sr1, off1, frac1):
        n = len(offF)
        cost = np.zeros(n, np.int32)
        self._f("bidir_satd_batch", None)(w, h, _ptr(fenc), C.c_ssize_t(sf), _ptr(offF), _ptr(ref0), C.c_ssize_t(sr0), _ptr(off0), _ptr(frac0),
                                          _ptr(ref1), C.c_ssize_t(sr1), _ptr(off1), _ptr(frac1), n, _ptr(cost))
        return cost

    def lowres_intra_frame(self, plane, origin, stride, wcu, hcu, penalty):
        cost = np.zeros(wcu * hcu, np.int32); mode = np.zeros(wcu * hcu, np.int32)
        self._f("lowres_intra_frame", None)(_ptr(plane, origin), C.c_ssize_t(stride), wcu, hcu, penalty, _ptr(cost), _ptr(mode))
        return cost, mode

    def blockop_batch(self, op, w, h, A, sa, offA, B, sb, offB, D, sd, offD):
        self._f("blockop_batch", None)(op, w, h, _ptr(A), C.c_ssize_t(sa), _ptr(offA), _ptr(B), C.c_ssize_t(sb), _ptr(offB),
                                       _ptr(D), C.c_ssize_t(sd), _ptr(offD), len(offA))
        return D

    def idct_batch(self, N, src, dst, stride, off, dst4=0):
        n = len(off)
        self._f("idct_batch", None)(N, dst4, _ptr(src), n, _ptr(dst), C.c_ssize_t(stride), _ptr(off))
        return dst


class Reference(_Base):
    prefix = "ref_"

    def __init__(self, depth):
        super().__init__(depth)
        path = os.path.join(ORACLE_DIR, "_ref", "libx265ref_%d.so" % depth)
        self.lib = C.CDLL(path)
        assert self.lib.ref_depth() == depth

    ME_FULL = 5         # x265.h:513-518 X265_DIA_SEARCH .. X265_FULL_SEARCH

    def motion_estimate(self, method, subme, w, h, fenc, of, sf, ref, orf, sr, rng, qmvp, mvc, merange, qp):
        """the reference's own MotionEstimate::motionEstimate (oracle/ref_motion.cpp); returns (qmv_x, qmv_y, cost)"""
        out = np.zeros(2, np.int32)
        rng = np.ascontiguousarray(rng, np.int32); qmvp = np.ascontiguousarray(qmvp, np.int32)
        mvc = np.ascontiguousarray(mvc, np.int32).reshape(-1)
        c = self.lib.ref_motion_estimate(method, subme, w, h, _ptr(fenc), C.c_ssize_t(sf), C.c_ssize_t(of), _ptr(ref), C.c_ssize_t(sr),
                                         C.c_ssize_t(orf - of), _ptr(rng), _ptr(qmvp), len(mvc) // 2, _ptr(mvc), merange, qp, _ptr(out))
        return int(out[0]), int(out[1]), int(c)

    def motion_estimate_batch(self, method, subme, w, h, fenc, sf, offF, ref, sr, offR, rng, qmvp, nc, mvc, merange, qp, nthreads=1):
        n = len(offF)
        mv = np.zeros((n, 2), np.int32); cost = np.zeros(n, np.int32)
        self.lib.ref_motion_estimate_batch(method, subme, w, h, _ptr(fenc), C.c_ssize_t(sf), _ptr(offF), _ptr(ref), C.c_ssize_t(sr), _ptr(offR),
                                           _ptr(rng), _ptr(qmvp), nc, _ptr(mvc) if nc else None, merange, qp, n, _ptr(mv), _ptr(cost), nthreads)
        return mv, cost

    def lowres_motion_estimate_ref(self, method, subme, w, h, fenc, of, sf, planes, orf, sr, pitch, rng, qmvp, merange, qp):
        out = np.zeros(2, np.int32)
        rng = np.ascontiguousarray(rng, np.int32); qmvp = np.ascontiguousarray(qmvp, np.int32)
        c = self.lib.ref_lowres_motion_estimate(method, subme, w, h, _ptr(fenc), C.c_ssize_t(sf), C.c_ssize_t(of), _ptr(planes), C.c_ssize_t(sr),
                                                C.c_size_t(pitch), C.c_ssize_t(orf), _ptr(rng), _ptr(qmvp), merange, qp, _ptr(out))
        return int(out[0]), int(out[1]), int(c)

    def motion_estimate_chroma_ref(self, method, subme, csp, w, h, fY, ofY, sfY, rY, orY, srY, fCb, fCr, ofC, sfC, rCb, rCr, orC, srC,
                                   rng, qmvp, mvc, merange, qp):
        """the reference's motionEstimate through the encoder-style setSourcePU (chroma SATD from subme 3);
        returns (qmv_x, qmv_y, cost, chroma_term_was_on)"""
        out = np.zeros(2, np.int32)
        rng = np.ascontiguousarray(rng, np.int32); qmvp = np.ascontiguousarray(qmvp, np.int32)
        mvc = np.ascontiguousarray(mvc, np.int32).reshape(-1)
        c = self.lib.ref_motion_estimate_chroma(method, subme, csp, w, h, _ptr(fY, ofY), C.c_ssize_t(sfY), _ptr(fCb, ofC), _ptr(fCr, ofC),
                                                C.c_ssize_t(sfC), _ptr(rY, orY), C.c_ssize_t(srY), _ptr(rCb, orC), _ptr(rCr, orC), C.c_ssize_t(srC),
                                                _ptr(rng), _ptr(qmvp), len(mvc) // 2, _ptr(mvc), merange, qp, _ptr(out))
        on = c >= 0
        return int(out[0]), int(out[1]), int(c if on else -1 - c), bool(on)

    def intra_allangs(self, N, ref_pix, filt_pix, bLuma):
        out = np.zeros(33 * N * N, ref_pix.dtype)
        self.lib.ref_intra_allangs(N, _ptr(out), _ptr(ref_pix), _ptr(filt_pix), int(bLuma))
        return out

    def motion_estimate_sea_ref(self, subme, w, h, fenc, of, sf, ref, orf, sr, sums, pitch, rng, qmvp, mvc, merange, qp):
        out = np.zeros(2, np.int32)
        rng = np.ascontiguousarray(rng, np.int32); qmvp = np.ascontiguousarray(qmvp, np.int32)
        mvc = np.ascontiguousarray(mvc, np.int32).reshape(-1)
        c = self.lib.ref_motion_estimate_sea(subme, w, h, _ptr(fenc), C.c_ssize_t(sf), C.c_ssize_t(of), _ptr(ref), C.c_ssize_t(sr), C.c_ssize_t(orf),
                                             _ptr(sums), C.c_size_t(pitch), _ptr(rng), _ptr(qmvp), len(mvc) // 2, _ptr(mvc), merange, qp, _ptr(out))
        return int(out[0]), int(out[1]), int(c)

    def mvcost_table(self, qp, radius):
        out = np.zeros(2 * radius + 1, np.uint16)
        self.lib.ref_mvcost_table(qp, radius, _ptr(out))
        return out

    def sa8d(self, w, h, a, oa, sa, b, ob, sb, chroma=0):
        return self._f("sa8d")(chroma, w, h, _ptr(a, oa), C.c_ssize_t(sa), _ptr(b, ob), C.c_ssize_t(sb))

    def sse_pp(self, w, h, a, oa, sa, b, ob, sb, chroma=0):
        return int(self._f("sse_pp", C.c_uint64)(chroma, w, h, _ptr(a, oa), C.c_ssize_t(sa), _ptr(b, ob), C.c_ssize_t(sb)))

    def sse_ss(self, w, a, oa, sa, b, ob, sb):
        return int(self._f("sse_ss", C.c_uint64)(w, _ptr(a, oa), C.c_ssize_t(sa), _ptr(b, ob), C.c_ssize_t(sb)))

    def pixelcmp_batch(self, op, w, h, A, sa, B, sb, offA, offB, nthreads=1):
        n = len(offA)
        out = np.zeros(n, np.uint64 if op == OP_SSE_PP else np.int32)
        r = self._f("pixelcmp_batch")(op, w, h, _ptr(A), C.c_ssize_t(sa), _ptr(B), C.c_ssize_t(sb),
                                      _ptr(offA), _ptr(offB), n, _ptr(out), nthreads)
        assert r == 0
        return out

    def dct_batch(self, N, src, stride, off, nthreads=1):
        n = len(off)
        out = np.zeros(n * N * N, np.int16)
        self._f("dct_batch")(N, _ptr(src), C.c_ssize_t(stride), _ptr(off), n, _ptr(out), nthreads)
        return out

    def residual_dct_batch(self, N, A, sa, B, sb, offA, offB, nthreads=1):
        n = len(offA)
        out = np.zeros(n * N * N, np.int16)
        self._f("residual_dct_batch")(N, _ptr(A), C.c_ssize_t(sa), _ptr(B), C.c_ssize_t(sb),
                                      _ptr(offA), _ptr(offB), n, _ptr(out), nthreads)
        return out

    def lowres_mvp_ref(self, fenc, of, sf, planes, orf, sr, pitch, mvc, bidir):
        mvc = np.ascontiguousarray(mvc, np.int32).reshape(-1)
        out = np.zeros(4, np.int32)
        self.lib.ref_lowres_mvp(_ptr(fenc), C.c_ssize_t(sf), C.c_ssize_t(of), _ptr(planes), C.c_ssize_t(sr), C.c_size_t(pitch), C.c_ssize_t(orf),
                                _ptr(mvc), len(mvc) // 2, int(bidir), _ptr(out))
        return out

    def lowres_bidir_ref(self, fenc, of, sf, planes0, s0, pitch0, planes1, s1, pitch1, orf, mv0, mv1):
        out = np.zeros(2, np.int32)
        mv0 = np.ascontiguousarray(mv0, np.int32); mv1 = np.ascontiguousarray(mv1, np.int32)
        self.lib.ref_lowres_bidir(_ptr(fenc), C.c_ssize_t(sf), C.c_ssize_t(of), _ptr(planes0), C.c_ssize_t(s0), C.c_size_t(pitch0), _ptr(planes1), C.c_ssize_t(s1),
                                  C.c_size_t(pitch1), C.c_ssize_t(orf), _ptr(mv0), _ptr(mv1), _ptr(out))
        return out

    def set_tier(self, tier):
        """0 = plain C table, 1 = C + the reference's SSE intrinsic DCT tier (batch entries only)"""
        self._f("set_tier", None)(int(tier))

    def tier_dct(self, tier, n, src, osrc, stride):
        dst = np.zeros(n * n, np.int16)
        self._f("tier_dct", None)(tier, n, _ptr(src, osrc), _ptr(dst), C.c_ssize_t(stride))
        return dst

    def tier_idct(self, tier, n, src, stride):
        dst = np.zeros(n * stride, np.int16)
        self._f("tier_idct", None)(tier, n, _ptr(src), _ptr(dst), C.c_ssize_t(stride))
        return dst

    def tu_forward_batch(self, N, A, sa, B, sb, offA, offB, qc, qbits, add, nthreads=1):
        """sub_ps -> dct -> quant per TU through the reference's slots: (levels[n*N*N], numSig[n])"""
        n = len(offA)
        lv = np.zeros(n * N * N, np.int16); ns = np.zeros(n, np.uint32)
        qc = np.ascontiguousarray(qc, np.int32)
        self._f("tu_forward_batch")(N, _ptr(A), C.c_ssize_t(sa), _ptr(B), C.c_ssize_t(sb), _ptr(offA), _ptr(offB), n,
                                    _ptr(qc), int(qbits), int(add), _ptr(lv), _ptr(ns), nthreads)
        return lv, ns

    def count_nonnull_slots(self):
        return self.lib.ref_count_nonnull_slots()


LUMA_PU = [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (8, 4), (4, 8), (16, 8), (8, 16), (32, 16),
           (16, 32), (64, 32), (32, 64), (16, 12), (12, 16), (16, 4), (4, 16), (32, 24), (24, 32),
           (32, 8), (8, 32), (64, 48), (48, 64), (64, 16), (16, 64)]          # primitives.h:41-55
# chroma-only PU sizes (4:2:0 = luma/2 both ways, 4:2:2 = luma/2 horizontally)
CHROMA_ONLY_420 = [(4, 2), (2, 4), (8, 6), (6, 8), (8, 2), (2, 8)]
CHROMA_ONLY_422 = [(2, 4), (2, 8), (4, 32), (8, 64), (8, 12), (6, 16), (2, 16), (16, 24), (12, 32),
                   (32, 48), (24, 64)]
SATD_CHROMA_422 = [(4, 32), (8, 64), (8, 12), (16, 24), (12, 32), (32, 48), (24, 64)]
