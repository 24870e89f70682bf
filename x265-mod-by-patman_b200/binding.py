"""ctypes binding of include/x265b200.h (plumbing for tests and the benchmark)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "lib", "libx265b200.so")
HEADER = os.path.join(ROOT, "include", "x265b200.h")

OP_SAD, OP_SATD, OP_SA8D, OP_SSE_PP = 0, 1, 2, 3
TR_DCT, TR_DST, TR_LOWPASS = 0, 1, 2
ME_DIA, ME_HEX, ME_UMH, ME_STAR, ME_SEA, ME_FULL = 0, 1, 2, 3, 4, 5       # search methods, numbered as x265.h:511-519
IP_KINDS = {"hpp": 0, "hps": 1, "vpp": 2, "vps": 3, "vsp": 4, "vss": 5, "hvpp": 6, "p2s": 7}

_lib = None


def build_library(glue=True):
    """compile csrc/ for sm_100a (nvcc cross-compiles without a GPU)"""
    subprocess.run(["make", "-s", "-C", os.path.join(PKG_DIR, "csrc"), "all"], check=True)
    if glue and os.path.isdir("/root/reference/source"):
        subprocess.run(["make", "-s", "-C", os.path.join(PKG_DIR, "csrc"), "glue"], check=True)


GLUE_HEADER = os.path.join(ROOT, "include", "x265b200_glue.h")


def glue_path(depth):
    return os.path.join(PKG_DIR, "lib", "libx265b200_glue_%d.so" % depth)


def declared_symbols(header=HEADER):
    """every function name the header (include/x265b200.h by default) declares"""
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(x265b200_[A-Za-z0-9_]+)\s*\(", text)))


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("CUDA library %s is missing: run __graft_entry__.build() (there is no CPU fallback)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.x265b200_last_error.restype = C.c_char_p
        _lib.x265b200_launch_count.restype = C.c_uint64
        for name in ("x265b200_sse_pp", "x265b200_sse_ss", "x265b200_ssd_s", "x265b200_var"):
            getattr(_lib, name).restype = C.c_uint64
        for name in ("x265b200_quant", "x265b200_nquant", "x265b200_copy_cnt"):
            getattr(_lib, name).restype = C.c_uint32
        _lib.x265b200_host_alloc.restype = C.c_void_p
        _lib.x265b200_host_alloc.argtypes = [C.c_void_p, C.c_size_t]
        _lib.x265b200_host_free.argtypes = [C.c_void_p, C.c_void_p]
        for name in ("x265b200_host_free", "x265b200_plane_destroy", "x265b200_frame_job_destroy", "x265b200_transfer_stats"):
            getattr(_lib, name).restype = None
        for name in ("x265b200_close", "x265b200_sad_x3", "x265b200_sad_x4", "x265b200_dct", "x265b200_idct",
                     "x265b200_dequant_normal", "x265b200_dequant_scaling", "x265b200_interp", "x265b200_sub_ps", "x265b200_add_ps",
                     "x265b200_pixelavg_pp", "x265b200_addAvg", "x265b200_frame_init_lowres", "x265b200_integral_inith", "x265b200_integral_initv", "x265b200_weight_pp", "x265b200_weight_sp", "x265b200_blockcopy", "x265b200_denoise_dct",
                     "x265b200_intra_pred", "x265b200_intra_filter", "x265b200_intra_pred_allangs"):
            getattr(_lib, name).restype = None
    return _lib


def _p(arr, off=0):
    return C.c_void_p(arr.ctypes.data + int(off) * arr.itemsize)


def _dp(t, off=0):
    """device pointer of a torch CUDA tensor (or None)"""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr() + int(off) * t.element_size())


def _ss(v):
    return C.c_ssize_t(int(v))


class Context:
    """One x265b200 context = one device + one bit depth (mirrors a per-depth x265 build)."""

    def __init__(self, depth, device=0):
        self.lib = load_library()
        self.depth = depth
        self.pix = np.uint8 if depth == 8 else np.uint16
        self.pmax = (1 << depth) - 1
        h = C.c_void_p()
        r = self.lib.x265b200_open(int(device), int(depth), C.byref(h))
        if r != 0:
            raise RuntimeError("x265b200_open(device=%d, depth=%d) failed with %d: no usable sm_100 GPU "
                               "(the CUDA path has no CPU fallback)" % (device, depth, r))
        self.h = h
        self.host = HostAPI(self)

    def close(self):
        if self.h:
            self.lib.x265b200_close(self.h)
            self.h = None

    def check(self):
        st = self.lib.x265b200_status(self.h)
        if st != 0:
            raise RuntimeError("x265b200 error %d: %s" % (st, self.lib.x265b200_last_error(self.h).decode()))

    def set_dct_path(self, path):
        """0 = default (tcgen05 TU chain for N = 32, mma.sync elsewhere), 1 = CUDA-core butterfly / stage kernels (validation twin),
        2 = mma.sync kernels everywhere (the TU chain's two-kernel form for every size), 3 = tcgen05 TU chain for N = 16 as well"""
        r = self.lib.x265b200_set_dct_path(self.h, int(path))
        if r != 0:
            raise RuntimeError("x265b200_set_dct_path failed")

    def launch_count(self):
        return int(self.lib.x265b200_launch_count(self.h))

    def transfer_stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        self.lib.x265b200_transfer_stats(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def host_alloc(self, nbytes):
        """pinned host memory as a numpy uint8 array (freed with host_free)"""
        p = self.lib.x265b200_host_alloc(self.h, nbytes)
        if not p:
            raise RuntimeError("x265b200_host_alloc failed")
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (nbytes,)), p

    def host_free(self, p):
        self.lib.x265b200_host_free(self.h, p)

    def tme_search_batch(self, method, merange, subme, fenc_plane, ref_planes, cost_tab, bits_tab, radius, lam, pus):
        """pus: ctypes array of TmePU; cost_tab (uint16) / bits_tab (float32): numpy arrays of 2 * radius + 1 entries; returns a TmeResult array"""
        n = len(pus)
        res = (TmeResult * n)()
        refs = (C.c_void_p * len(ref_planes))(*[p.h for p in ref_planes])
        self._call("x265b200_tme_search_batch", int(method), int(merange), int(subme), fenc_plane.h, refs, len(ref_planes),
                   C.c_void_p(cost_tab.ctypes.data + 2 * radius), C.c_void_p(bits_tab.ctypes.data + 4 * radius), int(radius), C.c_uint64(int(lam)), pus, n, res)
        return res

    def tu_forward_batch(self, N, fenc, sf, pred, sp, offF, offP, qc, qbits, add, qCoef, numSig, sseZero=None, stream=0):
        self._call("x265b200_tu_forward_batch", N, _dp(fenc), _ss(sf), _dp(pred), _ss(sp), _dp(offF), _dp(offP), int(offF.numel()),
                   _dp(qc), qbits, add, _dp(qCoef), _dp(numSig), _dp(sseZero), C.c_void_p(stream))

    def sm_count(self):
        return int(self.lib.x265b200_sm_count(self.h))

    def _call(self, name, *args):
        r = getattr(self.lib, name)(self.h, *args)
        if r != 0:
            raise RuntimeError("%s failed (%d): %s" % (name, r, self.lib.x265b200_last_error(self.h).decode()))

    # ---------------- batched device entries: torch CUDA tensors in, results written into `out`
    def pixelcmp_batch(self, op, w, h, A, sa, B, sb, offA, offB, out, stream=0):
        self._call("x265b200_pixelcmp_batch", op, w, h, _dp(A), _ss(sa), _dp(B), _ss(sb), _dp(offA), _dp(offB),
                   int(offA.numel()), _dp(out), C.c_void_p(stream))

    def cu_satd_batch(self, S, F, sf, R, sr, offF, offR5, out, stream=0):
        self._call("x265b200_cu_satd_batch", int(S), _dp(F), _ss(sf), _dp(R), _ss(sr), _dp(offF), _dp(offR5), int(offF.numel()), _dp(out),
                   C.c_void_p(stream))

    def sad_multi_batch(self, w, h, F, sf, R, sr, offF, offR, K, out, stream=0):
        self._call("x265b200_sad_multi_batch", w, h, _dp(F), _ss(sf), _dp(R), _ss(sr), _dp(offF), _dp(offR), K,
                   int(offF.numel()), _dp(out), C.c_void_p(stream))

    def sse_ss_batch(self, w, h, A, sa, B, sb, offA, offB, out, stream=0):
        self._call("x265b200_sse_ss_batch", w, h, _dp(A), _ss(sa), _dp(B), _ss(sb), _dp(offA), _dp(offB),
                   int(offA.numel()), _dp(out), C.c_void_p(stream))

    def ssd_s_batch(self, size, A, sa, offA, out, stream=0):
        self._call("x265b200_ssd_s_batch", size, _dp(A), _ss(sa), _dp(offA), int(offA.numel()), _dp(out), C.c_void_p(stream))

    def ads_batch(self, terms, half, encDC, sums, sumOff, delta, cost, costOff, width, thresh, mvs, pitch, count, stream=0):
        self._call("x265b200_ads_batch", terms, half, _dp(encDC), _dp(sums), _dp(sumOff), _dp(delta), _dp(cost),
                   _dp(costOff), _dp(width), _dp(thresh), int(sumOff.numel()), _dp(mvs), pitch, _dp(count), C.c_void_p(stream))

    def dct_batch(self, kind, N, src, stride, off, dst, stream=0, count=None):
        """off=None: contiguous source TUs (then `count` gives the number of TUs)"""
        n = int(off.numel()) if off is not None else int(count)
        self._call("x265b200_dct_batch", kind, N, _dp(src), _ss(stride), _dp(off), n, _dp(dst), C.c_void_p(stream))

    def idct_batch(self, kind, N, src, dst, stride, off, stream=0, count=None):
        n = int(off.numel()) if off is not None else int(count)
        self._call("x265b200_idct_batch", kind, N, _dp(src), n, _dp(dst), _ss(stride), _dp(off), C.c_void_p(stream))

    def subpel_cmp_batch(self, op, w, h, fenc, sf, ref, sr, offF, offR, frac, K, cost, stream=0):
        self._call("x265b200_subpel_cmp_batch", op, w, h, _dp(fenc), _ss(sf), _dp(ref), _ss(sr), _dp(offF), _dp(offR), _dp(frac), K,
                   int(offF.numel()), _dp(cost), C.c_void_p(stream))

    def me_full_batch(self, w, h, merange, fenc, sf, ref, sr, offF, offR, rng, mvp, cost_tab_centre, bmv, bcost, stream=0):
        """cost_tab_centre: integer device address of the centre element of the uint16 mv-cost table"""
        self._call("x265b200_me_full_batch", w, h, int(merange), _dp(fenc), _ss(sf), _dp(ref), _ss(sr), _dp(offF), _dp(offR), _dp(rng), _dp(mvp),
                   C.c_void_p(int(cost_tab_centre)), int(offF.numel()), _dp(bmv), _dp(bcost), C.c_void_p(stream))

    def me_pattern_batch(self, method, w, h, merange, fenc, sf, ref, sr, offF, offR, rng, mvp, cost_tab_centre, bmv, bcost, stream=0):
        self._call("x265b200_me_pattern_batch", int(method), w, h, int(merange), _dp(fenc), _ss(sf), _dp(ref), _ss(sr), _dp(offF), _dp(offR), _dp(rng),
                   _dp(mvp), C.c_void_p(int(cost_tab_centre)), int(offF.numel()), _dp(bmv), _dp(bcost), C.c_void_p(stream))

    def motion_estimate_batch(self, method, w, h, merange, subme, fenc, sf, ref, sr, offF, offR, rng, qmvp, num_cand, mvc, cost_tab_centre,
                              out_qmv, out_cost, stream=0):
        self._call("x265b200_motion_estimate_batch", int(method), w, h, int(merange), int(subme), _dp(fenc), _ss(sf), _dp(ref), _ss(sr), _dp(offF), _dp(offR),
                   _dp(rng), _dp(qmvp), int(num_cand), _dp(mvc), C.c_void_p(int(cost_tab_centre)), int(offF.numel()), _dp(out_qmv), _dp(out_cost),
                   C.c_void_p(stream))

    def motion_estimate_sea_batch(self, w, h, merange, subme, fenc, sf, ref, sr, offF, offR, rng, qmvp, num_cand, mvc, cost_tab_centre,
                                  sums, plane_pitch, out_qmv, out_cost, stream=0):
        self._call("x265b200_motion_estimate_sea_batch", w, h, int(merange), int(subme), _dp(fenc), _ss(sf), _dp(ref), _ss(sr), _dp(offF), _dp(offR),
                   _dp(rng), _dp(qmvp), int(num_cand), _dp(mvc), C.c_void_p(int(cost_tab_centre)), _dp(sums), C.c_size_t(int(plane_pitch)),
                   int(offF.numel()), _dp(out_qmv), _dp(out_cost), C.c_void_p(stream))

    def lowres_motion_estimate_batch(self, method, w, h, merange, subme, fenc, sf, planes, sr, plane_pitch, offF, offR, rng, qmvp,
                                     cost_tab_centre, out_qmv, out_cost, stream=0):
        self._call("x265b200_lowres_motion_estimate_batch", int(method), w, h, int(merange), int(subme), _dp(fenc), _ss(sf), _dp(planes), _ss(sr),
                   C.c_size_t(int(plane_pitch)), _dp(offF), _dp(offR), _dp(rng), _dp(qmvp), C.c_void_p(int(cost_tab_centre)), int(offF.numel()),
                   _dp(out_qmv), _dp(out_cost), C.c_void_p(stream))

    def motion_estimate_chroma_batch(self, method, w, h, merange, subme, fenc, sf, ref, sr, offF, offR, fcb, fcr, sfc, rcb, rcr, src, offFC, offRC,
                                     hshift, vshift, rng, qmvp, num_cand, mvc, cost_tab_centre, out_qmv, out_cost, stream=0):
        self._call("x265b200_motion_estimate_chroma_batch", int(method), w, h, int(merange), int(subme), _dp(fenc), _ss(sf), _dp(ref), _ss(sr),
                   _dp(offF), _dp(offR), _dp(fcb), _dp(fcr), _ss(sfc), _dp(rcb), _dp(rcr), _ss(src), _dp(offFC), _dp(offRC), int(hshift), int(vshift),
                   _dp(rng), _dp(qmvp), int(num_cand), _dp(mvc), C.c_void_p(int(cost_tab_centre)), int(offF.numel()), _dp(out_qmv), _dp(out_cost),
                   C.c_void_p(stream))

    def subpel_cmp_chroma_batch(self, w, h, fenc, sf, ref, sr, offF, offR, frac, K, cost, accumulate=0, stream=0):
        self._call("x265b200_subpel_cmp_chroma_batch", w, h, _dp(fenc), _ss(sf), _dp(ref), _ss(sr), _dp(offF), _dp(offR), _dp(frac), int(K),
                   int(offF.numel()), _dp(cost), int(accumulate), C.c_void_p(stream))

    def bidir_satd_batch(self, w, h, fenc, sf, offF, ref0, sr0, off0, frac0, ref1, sr1, off1, frac1, cost, stream=0):
        self._call("x265b200_bidir_satd_batch", w, h, _dp(fenc), _ss(sf), _dp(offF), _dp(ref0), _ss(sr0), _dp(off0), _dp(frac0),
                   _dp(ref1), _ss(sr1), _dp(off1), _dp(frac1), int(offF.numel()), _dp(cost), C.c_void_p(stream))

    def lowres_intra_batch(self, plane, origin, stride, width_in_cu, height_in_cu, penalty, cost, mode, stream=0):
        self._call("x265b200_lowres_intra_batch", _dp(plane, origin), _ss(stride), int(width_in_cu), int(height_in_cu), int(penalty),
                   _dp(cost), _dp(mode), C.c_void_p(stream))

    def intra_pred_batch(self, N, neighbours, n, dst, stream=0):
        self._call("x265b200_intra_pred_batch", int(N), _dp(neighbours), int(n), _dp(dst), C.c_void_p(stream))

    def weight_batch(self, sp, src, ss, dst, ds, width, height, w0, rnd, shift, offset, stream=0):
        self._call("x265b200_weight_batch", int(sp), _dp(src), _ss(ss), _dp(dst), _ss(ds), width, height, w0, rnd, shift, offset, C.c_void_p(stream))

    def weight_cost_batch(self, fenc, ref, stride, width, height, intra, weights, K, cost, stream=0):
        self._call("x265b200_weight_cost_batch", _dp(fenc), _dp(ref), _ss(stride), width, height, _dp(intra), _dp(weights), int(K), _dp(cost), C.c_void_p(stream))

    def lowres_mvp_batch(self, fenc, sf, offF, planes, sr, pitch, offR, mvc, numc, bidir, mvp, mvp_cost, skip_cost, stream=0):
        self._call("x265b200_lowres_mvp_batch", _dp(fenc), _ss(sf), _dp(offF), _dp(planes), _ss(sr), C.c_size_t(int(pitch)), _dp(offR), _dp(mvc), _dp(numc),
                   int(bidir), int(offF.numel()), _dp(mvp), _dp(mvp_cost), _dp(skip_cost), C.c_void_p(stream))

    def lowres_bidir_cost_batch(self, fenc, sf, offF, planes0, s0, pitch0, planes1, s1, pitch1, offR, mv0, mv1, cost, stream=0):
        self._call("x265b200_lowres_bidir_cost_batch", _dp(fenc), _ss(sf), _dp(offF), _dp(planes0), _ss(s0), C.c_size_t(int(pitch0)), _dp(planes1), _ss(s1),
                   C.c_size_t(int(pitch1)), _dp(offR), _dp(mv0), _dp(mv1), int(offF.numel()), _dp(cost), C.c_void_p(stream))

    def me_integral_batch(self, pix, stride, rows, nframes, sums, plane_pitch, stream=0):
        self._call("x265b200_me_integral_batch", _dp(pix), _ss(stride), int(rows), int(nframes), _dp(sums), C.c_size_t(int(plane_pitch)), C.c_void_p(stream))

    def var_batch(self, size, pix, stride, off, n, out, stream=0):
        self._call("x265b200_var_batch", size, _dp(pix), _ss(stride), _dp(off), int(n), _dp(out), C.c_void_p(stream))

    def psy_cost_batch(self, size, src, ss, offS, rec, sr, offR, n, out, stream=0):
        self._call("x265b200_psy_cost_batch", size, _dp(src), _ss(ss), _dp(offS), _dp(rec), _ss(sr), _dp(offR), int(n), _dp(out), C.c_void_p(stream))

    def count_nonzero_batch(self, size, src, stride, off, n, coeff, count, stream=0):
        self._call("x265b200_count_nonzero_batch", size, _dp(src), _ss(stride), _dp(off), int(n), _dp(coeff), _dp(count), C.c_void_p(stream))

    def denoise_dct_batch(self, dct, res_sum, offset, num_coeff, n, stream=0):
        self._call("x265b200_denoise_dct_batch", _dp(dct), _dp(res_sum), _dp(offset), int(num_coeff), int(n), C.c_void_p(stream))

    def blockcopy_batch(self, kind, w, h, S, ss, offS, D, sd, offD, n, param=0, stream=0):
        self._call("x265b200_blockcopy_batch", kind, w, h, _dp(S), _ss(ss), _dp(offS), _dp(D), _ss(sd), _dp(offD), int(n), int(param), C.c_void_p(stream))

    def blockop_batch(self, op, w, h, A, sa, offA, B, sb, offB, D, sd, offD, n, stream=0):
        self._call("x265b200_blockop_batch", op, w, h, _dp(A), _ss(sa), _dp(offA), _dp(B), _ss(sb), _dp(offB), _dp(D), _ss(sd), _dp(offD),
                   int(n), C.c_void_p(stream))

    def lowres_batch(self, src, ss, d0, dh, dv, dc, ds, width, height, stream=0):
        self._call("x265b200_lowres_batch", _dp(src), _ss(ss), _dp(d0), _dp(dh), _dp(dv), _dp(dc), _ss(ds), width, height, C.c_void_p(stream))

    def quant_batch(self, coef, qc, deltaU, qCoef, qBits, add, numCoeff, n, numSig, stream=0):
        self._call("x265b200_quant_batch", _dp(coef), _dp(qc), _dp(deltaU), _dp(qCoef), qBits, add, numCoeff, n,
                   _dp(numSig), C.c_void_p(stream))

    def dequant_normal_batch(self, q, coef, num, scale, shift, stream=0):
        self._call("x265b200_dequant_normal_batch", _dp(q), _dp(coef), num, scale, shift, C.c_void_p(stream))

    def dequant_scaling_batch(self, q, dq, coef, num, n, per, shift, stream=0):
        self._call("x265b200_dequant_scaling_batch", _dp(q), _dp(dq), _dp(coef), num, n, per, shift, C.c_void_p(stream))

    def interp_batch(self, kind, taps, w, h, src, ss, offSrc, dst, ds, offDst, coeffIdx, stream=0):
        self._call("x265b200_interp_batch", IP_KINDS[kind], taps, w, h, _dp(src), _ss(ss), _dp(offSrc), _dp(dst), _ss(ds),
                   _dp(offDst), _dp(coeffIdx), int(offSrc.numel()), C.c_void_p(stream))

    def tu_chain_batch(self, N, fenc, sf, pred, sp, offF, offP, qc, qbits, add, dqscale, dqshift, qCoef, numSig, recon, sr, offR,
                       sseZero, sseRecon, stream=0, ttype=None):
        """ttype None: x265b200_tu_chain_batch (inter luma / chroma); TU_INTER / TU_INTRA_LUMA: x265b200_tu_chain_tt_batch"""
        if ttype is None:
            self._call("x265b200_tu_chain_batch", N, _dp(fenc), _ss(sf), _dp(pred), _ss(sp), _dp(offF), _dp(offP), int(offF.numel()),
                       _dp(qc), qbits, add, dqscale, dqshift, _dp(qCoef), _dp(numSig), _dp(recon), _ss(sr), _dp(offR),
                       _dp(sseZero), _dp(sseRecon), C.c_void_p(stream))
        else:
            self._call("x265b200_tu_chain_tt_batch", N, int(ttype), _dp(fenc), _ss(sf), _dp(pred), _ss(sp), _dp(offF), _dp(offP), int(offF.numel()),
                       _dp(qc), qbits, add, dqscale, dqshift, _dp(qCoef), _dp(numSig), _dp(recon), _ss(sr), _dp(offR),
                       _dp(sseZero), _dp(sseRecon), C.c_void_p(stream))

    def residual_batch(self, w, h, A, sa, B, sb, offA, offB, dst, stream=0):
        self._call("x265b200_residual_batch", w, h, _dp(A), _ss(sa), _dp(B), _ss(sb), _dp(offA), _dp(offB),
                   int(offA.numel()), _dp(dst), C.c_void_p(stream))


PASS_CMP, PASS_COEF, PASS_LEVELS = 0, 1, 2
TU_INTER, TU_INTRA_LUMA = 0, 1


class PassResult(C.Structure):
    """x265b200_pass_result (include/x265b200.h)"""
    _fields_ = [("kind", C.c_int), ("n", C.c_int), ("cost", C.POINTER(C.c_int32)), ("coef", C.POINTER(C.c_int16)),
                ("numSig", C.POINTER(C.c_uint16)), ("sigMap", C.POINTER(C.c_uint32)), ("levels", C.POINTER(C.c_int16)),
                ("nlevels", C.c_uint32)]


TME_MAX_CAND = 8


class TmePU(C.Structure):
    """x265b200_tme_pu"""
    _fields_ = [("w", C.c_int16), ("h", C.c_int16), ("ref", C.c_int16), ("numCand", C.c_int16), ("offF", C.c_int32), ("offR", C.c_int32),
                ("mvmin", C.c_int32 * 2), ("mvmax", C.c_int32 * 2), ("mvp", C.c_int32 * 2), ("mvc", (C.c_int32 * 2) * TME_MAX_CAND), ("bits", C.c_uint32)]


class TmeResult(C.Structure):
    """x265b200_tme_result"""
    _fields_ = [("mv", C.c_int32 * 2), ("mvCost", C.c_uint32), ("bits", C.c_uint32), ("cost", C.c_uint32), ("satdCost", C.c_uint32)]


class Plane:
    """x265b200_plane: a picture plane resident in HBM with PicYuv's geometry; uploads take host (numpy / pinned) buffers"""

    def __init__(self, ctx, width, height, ctu=64, hshift=0, vshift=0):
        self.ctx = ctx
        h = C.c_void_p()
        ctx._call("x265b200_plane_create", int(width), int(height), int(ctu), int(hshift), int(vshift), C.byref(h))
        self.h = h
        stride, rows, origin, elems, dev = C.c_ssize_t(), C.c_int(), C.c_int32(), C.c_size_t(), C.c_void_p()
        ctx.lib.x265b200_plane_info(self.h, C.byref(stride), C.byref(rows), C.byref(origin), C.byref(elems), C.byref(dev))
        self.stride, self.rows, self.origin, self.elems, self.device_ptr = stride.value, rows.value, origin.value, elems.value, dev.value

    def _r(self, name, *args):
        r = getattr(self.ctx.lib, name)(self.h, *args)
        if r != 0:
            raise RuntimeError("%s failed (%d): %s" % (name, r, self.ctx.lib.x265b200_last_error(self.ctx.h).decode()))

    def upload_padded(self, host, addr=None):
        """host: numpy array of `elems` pixels, or a raw address (pinned memory)"""
        self._r("x265b200_plane_upload_padded", C.c_void_p(addr if addr is not None else host.ctypes.data))

    def upload_picture(self, host, host_stride, addr=None):
        self._r("x265b200_plane_upload_picture", C.c_void_p(addr if addr is not None else host.ctypes.data), _ss(host_stride))

    def upload_rows(self, host, addr=None):
        self._r("x265b200_plane_upload_rows", C.c_void_p(addr if addr is not None else host.ctypes.data))

    def download_padded(self):
        out = np.empty(self.elems, self.ctx.pix)
        self._r("x265b200_plane_download_padded", _p(out))
        return out

    def destroy(self):
        if self.h:
            self.ctx.lib.x265b200_plane_destroy(self.h)
            self.h = None


class FrameJob:
    """x265b200_frame_job: analysis passes registered once, run per (fenc plane, reference plane) pair, several frames in flight"""

    def __init__(self, ctx, width, height, ctu=64, slots=3):
        self.ctx = ctx
        h = C.c_void_p()
        ctx._call("x265b200_frame_job_create", int(width), int(height), int(ctu), int(slots), C.byref(h))
        self.h = h
        self.sizes = []         # N (transform) or None per pass

    def _r(self, name, *args):
        r = getattr(self.ctx.lib, name)(self.h, *args)
        if r < 0:
            raise RuntimeError("%s failed (%d): %s" % (name, r, self.ctx.lib.x265b200_last_error(self.ctx.h).decode()))
        return r

    def add_cmp(self, op, w, h, offF, offR):
        offF = np.ascontiguousarray(offF, np.int32); offR = np.ascontiguousarray(offR, np.int32)
        self.sizes.append(None)
        return self._r("x265b200_frame_job_add_cmp", int(op), w, h, _p(offF), _p(offR), len(offF))

    def add_transform(self, kind, N, offF, offR, qc=None, qbits=0, add=0):
        offF = np.ascontiguousarray(offF, np.int32); offR = np.ascontiguousarray(offR, np.int32)
        qp = _p(np.ascontiguousarray(qc, np.int32)) if qc is not None else C.c_void_p(0)
        self.sizes.append(N)
        return self._r("x265b200_frame_job_add_transform", int(kind), N, _p(offF), _p(offR), len(offF), qp, int(qbits), int(add))

    def set_blocks(self, p, offF, offR):
        offF = np.ascontiguousarray(offF, np.int32); offR = np.ascontiguousarray(offR, np.int32)
        self._r("x265b200_frame_job_set_blocks", int(p), _p(offF), _p(offR))

    def submit(self, fenc, ref):
        return self._r("x265b200_frame_job_submit", fenc.h, ref.h)

    def wait(self, slot, copy=True):
        """list of dicts, one per pass, numpy views (copies when copy=True) of the job's pinned result memory"""
        npass = self.ctx.lib.x265b200_frame_job_pass_count(self.h)
        arr = (PassResult * npass)()
        self._r("x265b200_frame_job_wait", int(slot), arr, npass)
        out = []
        for i in range(npass):
            r = arr[i]
            fin = (lambda a: a.copy()) if copy else (lambda a: a)
            if r.kind == PASS_CMP:
                out.append({"kind": r.kind, "cost": fin(np.ctypeslib.as_array(r.cost, (r.n,)))})
            elif r.kind == PASS_COEF:
                N = self.sizes[i]
                out.append({"kind": r.kind, "coef": fin(np.ctypeslib.as_array(r.coef, (r.n * N * N,)))})
            else:
                N = self.sizes[i]
                words = (r.n * N * N + 31) // 32
                out.append({"kind": r.kind, "N": N, "n": r.n, "nlevels": int(r.nlevels),
                            "numSig": fin(np.ctypeslib.as_array(r.numSig, (r.n,))),
                            "sigMap": fin(np.ctypeslib.as_array(r.sigMap, (words,))),
                            "levels": fin(np.ctypeslib.as_array(r.levels, (max(int(r.nlevels), 1),))[:int(r.nlevels)])})
        return out

    def destroy(self):
        if self.h:
            self.ctx.lib.x265b200_frame_job_destroy(self.h)
            self.h = None


def expand_levels(res):
    """dense int16 levels of a LEVELS pass result (what the entropy coder's significance walk reconstructs)"""
    total = res["n"] * res["N"] * res["N"]
    bits = np.unpackbits(res["sigMap"].view(np.uint8), bitorder="little")[:total].astype(bool)
    dense = np.zeros(total, np.int16)
    dense[bits] = res["levels"]
    return dense


class HostAPI:
    """Per-call host entries with numpy buffers; same method names as the CPU checkers in tests/ so the
    reference-style parity cases can run against it unchanged."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.lib = ctx.lib
        self.h = ctx.h
        self.depth = ctx.depth
        self.pix = ctx.pix
        self.pmax = ctx.pmax

    def _cmp(self, name, w, h, a, oa, sa, b, ob, sb):
        r = getattr(self.lib, name)(self.h, w, h, _p(a, oa), _ss(sa), _p(b, ob), _ss(sb))
        self.ctx.check()
        return int(r)

    def sad(self, w, h, a, oa, sa, b, ob, sb): return self._cmp("x265b200_sad", w, h, a, oa, sa, b, ob, sb)
    def satd(self, w, h, a, oa, sa, b, ob, sb): return self._cmp("x265b200_satd", w, h, a, oa, sa, b, ob, sb)
    def sa8d(self, w, h, a, oa, sa, b, ob, sb, chroma=0): return self._cmp("x265b200_sa8d", w, h, a, oa, sa, b, ob, sb)

    def sse_pp(self, w, h, a, oa, sa, b, ob, sb, chroma=0):
        r = self._cmp("x265b200_sse_pp", w, h, a, oa, sa, b, ob, sb)
        return r & 0xFFFFFFFF if self.depth == 8 else r         # sse_t is uint32 at 8 bit (common.h:145-149)

    def sse_ss(self, w, a, oa, sa, b, ob, sb):
        r = self._cmp("x265b200_sse_ss", w, w, a, oa, sa, b, ob, sb)
        return r & 0xFFFFFFFF if self.depth == 8 else r

    def ssd_s(self, size, a, oa, sa):
        r = int(self.lib.x265b200_ssd_s(self.h, size, _p(a, oa), _ss(sa)))
        self.ctx.check()
        return r & 0xFFFFFFFF if self.depth == 8 else r

    def sad_x3(self, w, h, fenc, of, ref, offs, rs):
        res = np.zeros(3, np.int32)
        self.lib.x265b200_sad_x3(self.h, w, h, _p(fenc, of), _p(ref, offs[0]), _p(ref, offs[1]), _p(ref, offs[2]), _ss(rs), _p(res))
        self.ctx.check()
        return res

    def sad_x4(self, w, h, fenc, of, ref, offs, rs):
        res = np.zeros(4, np.int32)
        self.lib.x265b200_sad_x4(self.h, w, h, _p(fenc, of), _p(ref, offs[0]), _p(ref, offs[1]), _p(ref, offs[2]),
                                 _p(ref, offs[3]), _ss(rs), _p(res))
        self.ctx.check()
        return res

    def ads(self, w, h, encDC, sums, osum, delta, cost, width, thresh):
        enc = np.ascontiguousarray(encDC, np.int32)
        mvs = np.full(width + 8, -1, np.int16)
        n = self.lib.x265b200_ads(self.h, w, h, _p(enc), _p(sums, osum), int(delta), _p(cost), _p(mvs), int(width), int(thresh))
        self.ctx.check()
        return n, mvs[:n].copy()

    def dct(self, n, src, osrc, stride, kind=TR_DCT):
        dst = np.zeros(n * n, np.int16)
        self.lib.x265b200_dct(self.h, kind, n, _p(src, osrc), _p(dst), _ss(stride))
        self.ctx.check()
        return dst

    def idct(self, n, src, stride, dst=None, odst=0, kind=TR_DCT):
        if dst is None:
            dst = np.zeros(n * stride, np.int16)
        self.lib.x265b200_idct(self.h, kind, n, _p(src), _p(dst, odst), _ss(stride))
        self.ctx.check()
        return dst

    def dst4(self, src, osrc, stride): return self.dct(4, src, osrc, stride, TR_DST)
    def idst4(self, src, stride): return self.idct(4, src, stride, kind=TR_DST)

    def lowpass_dct(self, n, src, osrc, stride):
        dst = np.full(n * n, 0x5a5a, np.int16)
        self.lib.x265b200_dct(self.h, TR_LOWPASS, n, _p(src, osrc), _p(dst), _ss(stride))
        self.ctx.check()
        return dst

    def quant(self, coef, qc, qbits, add, n):
        deltaU = np.zeros(n, np.int32)
        q = np.zeros(n, np.int16)
        r = self.lib.x265b200_quant(self.h, _p(coef), _p(qc), _p(deltaU), _p(q), qbits, add, n)
        self.ctx.check()
        return int(r), q, deltaU

    def nquant(self, coef, qc, qbits, add, n):
        q = np.zeros(n, np.int16)
        r = self.lib.x265b200_nquant(self.h, _p(coef), _p(qc), _p(q), qbits, add, n)
        self.ctx.check()
        return int(r), q

    def dequant_normal(self, q, n, scale, shift):
        out = np.zeros(n, np.int16)
        self.lib.x265b200_dequant_normal(self.h, _p(q), _p(out), n, scale, shift)
        self.ctx.check()
        return out

    def dequant_scaling(self, q, dq, n, per, shift):
        out = np.zeros(n, np.int16)
        self.lib.x265b200_dequant_scaling(self.h, _p(q), _p(dq), _p(out), n, per, shift)
        self.ctx.check()
        return out

    def interp(self, kind, N, w, h, src, os_, ss, dst, od, ds, idx, idy_or_ext=0):
        self.lib.x265b200_interp(self.h, IP_KINDS[kind], N, w, h, _p(src, os_), _ss(ss), _p(dst, od), _ss(ds), idx, idy_or_ext)
        self.ctx.check()
        return 0

    def p2s(self, w, h, src, os_, ss, dst, od, ds):
        self.lib.x265b200_interp(self.h, IP_KINDS["p2s"], 8, w, h, _p(src, os_), _ss(ss), _p(dst, od), _ss(ds), 0, 0)
        self.ctx.check()
        return 0
