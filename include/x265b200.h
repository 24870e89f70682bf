/*
 * x265b200.h -- C ABI of the B200 (sm_100a) analysis-primitive path for x265.
 *
 * Drop-in boundary: x265's EncoderPrimitives function-pointer table
 * (reference source/common/primitives.h:239-432).  Two families of entry points:
 *
 *  A. per-call HOST entries (section "host") -- one per primitive typedef of
 *     primitives.h:133-182, same argument order and meaning, with (ctx, width, height)
 *     prepended because the reference bakes the block size into the slot (pu[part] / cu[size])
 *     and the pixel type into the build (X265_DEPTH).  `pixel` is uint8_t when the context was
 *     opened at 8 bit and uint16_t at 10/12 bit (common.h:127-143); it is passed as void*.
 *     setupB200Primitives() (csrc/setup_b200_primitives.cpp) binds every hot-path slot to a thunk
 *     that calls these.  They stage the block through pinned memory, run the same CUDA kernels as
 *     family B with n = 1, and copy the result back.  There is NO CPU fallback: if the device is
 *     unusable the call records a sticky error (x265b200_status) and returns 0 / leaves outputs
 *     untouched, because a slot has no way to report errors (SURVEY.md 8b "Errors").
 *
 *  B. batched DEVICE entries (section "device") -- arrays of block descriptors over planes that are
 *     already resident in HBM, one launch per primitive class and block shape.  All pointers are
 *     device pointers, offsets are element offsets from the plane base, `stream` is a cudaStream_t.
 *     These have no counterpart in the reference (it calls one block at a time); they are what a
 *     batching caller (ThreadedME, lookahead, the benchmark) uses.
 *
 * All functions return X265B200_OK (0) or a negative error code unless stated otherwise.
 */
#ifndef X265B200_H
#define X265B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct x265b200_ctx x265b200_ctx;
typedef void* x265b200_stream;          /* cudaStream_t; NULL = the legacy default stream */

enum {
    X265B200_OK = 0,
    X265B200_ERR_ARG = -1,              /* bad size / pointer / depth */
    X265B200_ERR_CUDA = -2,             /* a CUDA runtime call failed (see x265b200_last_error) */
    X265B200_ERR_NO_DEVICE = -3
};

/* block-compare operations for x265b200_pixelcmp_batch */
enum {
    X265B200_SAD = 0,                   /* pixelcmp_t  pu[].sad        pixel.cpp:40-55   */
    X265B200_SATD = 1,                  /* pixelcmp_t  pu[].satd       pixel.cpp:190-289 */
    X265B200_SA8D = 2,                  /* pixelcmp_t  cu[].sa8d       pixel.cpp:291-369 */
    X265B200_SSE_PP = 3                 /* pixel_sse_t cu[].sse_pp     pixel.cpp:167-186 */
};

/* transform kinds for x265b200_dct_batch / x265b200_idct_batch */
enum {
    X265B200_TR_DCT = 0,                /* cu[].dct / cu[].idct        dct.cpp:443-611 */
    X265B200_TR_DST = 1,                /* dst4x4 / idst4x4 (N = 4)    dct.cpp:43-81   */
    X265B200_TR_LOWPASS = 2             /* cu[].lowpass_dct (forward only, N = 8,16,32) lowpassdct.cpp:34-116 */
};

/* interpolation kinds for x265b200_interp_batch */
enum {
    X265B200_IP_HPP = 0,                /* filter_pp_t    luma_hpp / filter_hpp   ipfilter.cpp:79-118  */
    X265B200_IP_HPS = 1,                /* filter_hps_t   luma_hps / filter_hps   ipfilter.cpp:120-162 */
    X265B200_IP_VPP = 2,                /* filter_pp_t    luma_vpp / filter_vpp   ipfilter.cpp:164-203 */
    X265B200_IP_VPS = 3,                /* filter_ps_t    luma_vps / filter_vps   ipfilter.cpp:205-238 */
    X265B200_IP_VSP = 4,                /* filter_sp_t    luma_vsp / filter_vsp   ipfilter.cpp:240-283 */
    X265B200_IP_VSS = 5,                /* filter_ss_t    luma_vss / filter_vss   ipfilter.cpp:285-317 */
    X265B200_IP_HVPP = 6,               /* filter_hv_pp_t luma_hvpp               ipfilter.cpp:362-369 */
    X265B200_IP_P2S = 7                 /* filter_p2s_t   convert_p2s / p2s       ipfilter.cpp:40-57   */
};

/* two-input block operations for x265b200_blockop_batch (slots adjacent to the hot path) */
enum {
    X265B200_BOP_SUB_PS = 0,            /* pixel_sub_ps_t  cu[].sub_ps       pixel.cpp:806-818  int16 = pixel - pixel */
    X265B200_BOP_ADD_PS = 1,            /* pixel_add_ps_t  cu[].add_ps       pixel.cpp:820-832  pixel = clip(pixel + int16) */
    X265B200_BOP_PIXELAVG = 2,          /* pixelavg_pp_t   pu[].pixelavg_pp  pixel.cpp:537-549  pixel = (pixel + pixel + 1) >> 1 */
    X265B200_BOP_ADDAVG = 3             /* addAvg_t        pu[].addAvg       pixel.cpp:834-855  pixel = clip((int16 + int16 + offset) >> shift) */
};

/* ------------------------------------------------------------------ lifecycle */

/* bit_depth in {8, 10, 12}: fixes sizeof(pixel) and every depth-dependent shift, exactly as
 * X265_DEPTH does at compile time in the reference (source/CMakeLists.txt:787-798). */
int x265b200_open(int device, int bit_depth, x265b200_ctx** ctx);
void x265b200_close(x265b200_ctx* ctx);
int x265b200_bit_depth(const x265b200_ctx* ctx);
int x265b200_sm_count(const x265b200_ctx* ctx);
/* sticky status: first error recorded by any entry (host entries cannot return one) */
int x265b200_status(const x265b200_ctx* ctx);
const char* x265b200_last_error(const x265b200_ctx* ctx);
/* Implementation switch for the transforms and the TU chain: 0 (default) = tensor cores: warp-level mma.sync for the stand-alone DCT / IDCT
 * (csrc/transform_mma.cu), the single tcgen05 / tensor-memory kernel for the 32x32 TU chain (csrc/tu_umma.cuh) and the two fused mma.sync
 * kernels of csrc/tu_fused.cuh for the smaller TUs;
 * 1 = the CUDA-core partial-butterfly and stage kernels (csrc/transform.cu), kept as the validation twin;
 * 2 = mma.sync everywhere (the TU chain as the two fused kernels for every size);
 * 3 = like 0 with the tcgen05 kernel for 16x16 TUs too. */
int x265b200_set_dct_path(x265b200_ctx* ctx, int path);
/* number of kernel launches issued through this context so far (bench.py's gpu_launches) */
uint64_t x265b200_launch_count(const x265b200_ctx* ctx);

/* ------------------------------------------------------------------ device (batched) */

/* out[i] = op(planeA + offA[i], strideA, planeB + offB[i], strideB) for i < n.
 * out is int32[n] for SAD/SATD/SA8D and uint64[n] for SSE_PP.  Any w,h multiple of 4
 * (SA8D: both multiples of 16 -> per-16x16 rounding, both multiples of 8 -> per-8x8, else SATD,
 * which reproduces every luma/chroma sa8d slot binding, pixel.cpp:1180-1184,1260-1263,1339-1342). */
int x265b200_pixelcmp_batch(x265b200_ctx* ctx, int op, int w, int h,
                            const void* planeA, intptr_t strideA, const void* planeB, intptr_t strideB,
                            const int32_t* offA, const int32_t* offB, int n, void* out, x265b200_stream stream);

/* K candidates per block sharing one fenc block (sad_x3 / sad_x4 generalised, pixel.cpp:74-119):
 * out[i*K + k] = sad(fenc + offF[i], strideF, ref + offR[i*K + k], strideR). */
int x265b200_sad_multi_batch(x265b200_ctx* ctx, int w, int h, const void* fenc, intptr_t strideF,
                             const void* ref, intptr_t strideR, const int32_t* offF, const int32_t* offR,
                             int K, int n, int32_t* out, x265b200_stream stream);

/* SATD of every rectangular partition of n CUs of cuSize x cuSize (8, 16, 32, 64) in one pass, each PU against its own reference
 * block -- what the inter analysis measures per CU for SIZE_2Nx2N, SIZE_2NxN and SIZE_Nx2N (reference encoder/analysis.cpp
 * checkInter_rd0_4 -> encoder/search.cpp predInterSearch, one pu[part].satd per PU).  offF[i]: the CU in the fenc plane;
 * offR[5 * i + k] / cost[5 * i + k]: k = 0 the 2Nx2N PU, 1 / 2 the upper / lower 2NxN PU, 3 / 4 the left / right Nx2N PU, each offset
 * addressing the top-left sample of that PU's reference block (PU position + its motion vector).  Equal, cost for cost, to five
 * x265b200_pixelcmp_batch(SATD) results; the fenc CU is read once instead of three times and one launch replaces three. */
int x265b200_cu_satd_batch(x265b200_ctx* ctx, int cuSize, const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR,
                           const int32_t* offF, const int32_t* offR, int n, int32_t* cost, x265b200_stream stream);

/* pixel_sse_ss_t (pixel.cpp:167-186 with int16 inputs) and pixel_ssd_s_t (pixel.cpp:371-383);
 * out is uint64[n] (truncate to 32 bit for an 8-bit build's sse_t). */
int x265b200_sse_ss_batch(x265b200_ctx* ctx, int w, int h, const int16_t* A, intptr_t strideA,
                          const int16_t* B, intptr_t strideB, const int32_t* offA, const int32_t* offB,
                          int n, uint64_t* out, x265b200_stream stream);
int x265b200_ssd_s_batch(x265b200_ctx* ctx, int size, const int16_t* A, intptr_t strideA,
                         const int32_t* offA, int n, uint64_t* out, x265b200_stream stream);

/* SEA candidate filter (pixelcmp_ads_t, pixel.cpp:121-165).  Job i scans `width[i]` positions of the
 * integral-sum row starting at sums + sumOff[i]; `terms` in {1,2,4} DC terms (half = w>>1 as in
 * ads_x4); survivors are written in ascending order to mvs + i*mvsPitch, their count to count[i]. */
int x265b200_ads_batch(x265b200_ctx* ctx, int terms, int half, const int32_t* encDC /* n*4 */,
                       const uint32_t* sums, const int32_t* sumOff, const int32_t* delta,
                       const uint16_t* costMvX, const int32_t* costOff, const int32_t* width,
                       const int32_t* thresh, int n, int16_t* mvs, int mvsPitch, int32_t* count,
                       x265b200_stream stream);

/* forward transform of n blocks: src block i at src + off[i] with srcStride (elements);
 * dst block i contiguous at dst + i*N*N.  off == NULL means contiguous source blocks (off[i] = i*N*N,
 * use srcStride = N).  (dct_t, primitives.h:153) */
int x265b200_dct_batch(x265b200_ctx* ctx, int kind, int N, const int16_t* src, intptr_t srcStride,
                       const int32_t* off, int n, int16_t* dst, x265b200_stream stream);
/* inverse: src block i contiguous at src + i*N*N, dst block i at dst + off[i] with dstStride
 * (off == NULL: dst + i*N*N). (idct_t) */
int x265b200_idct_batch(x265b200_ctx* ctx, int kind, int N, const int16_t* src, int n,
                        int16_t* dst, intptr_t dstStride, const int32_t* off, x265b200_stream stream);

/* quant_t / nquant_t over n blocks of numCoeff coefficients each (contiguous); quantCoeff is one
 * table of numCoeff entries shared by all blocks; numSig[i] receives the return value of block i.
 * deltaU == NULL selects nquant (dct.cpp:690-715), else quant (dct.cpp:666-688). */
int x265b200_quant_batch(x265b200_ctx* ctx, const int16_t* coef, const int32_t* quantCoeff, int32_t* deltaU,
                         int16_t* qCoef, int qBits, int add, int numCoeff, int n, uint32_t* numSig,
                         x265b200_stream stream);
/* dequant_normal_t (dct.cpp:614-636) over `num` coefficients (any multiple of 8) */
int x265b200_dequant_normal_batch(x265b200_ctx* ctx, const int16_t* quantCoef, int16_t* coef, int num,
                                  int scale, int shift, x265b200_stream stream);
/* dequant_scaling_t (dct.cpp:638-664): n blocks of `num` coefficients sharing one table of `num` entries */
int x265b200_dequant_scaling_batch(x265b200_ctx* ctx, const int16_t* quantCoef, const int32_t* deQuantCoef,
                                   int16_t* coef, int num, int n, int per, int shift, x265b200_stream stream);

/* interpolation of n blocks.  taps = 8 (luma) or 4 (chroma).  src/dst element types follow the kind
 * (pixel or int16).  coeffIdx[i] packs idxX | idxY << 4 | isRowExt << 8 (idxY only for HVPP,
 * isRowExt only for HPS).  For HPS with isRowExt the block writes h + taps - 1 rows starting at
 * dst + offDst[i], exactly like the reference. */
int x265b200_interp_batch(x265b200_ctx* ctx, int kind, int taps, int w, int h,
                          const void* src, intptr_t srcStride, const int32_t* offSrc,
                          void* dst, intptr_t dstStride, const int32_t* offDst,
                          const int32_t* coeffIdx, int n, x265b200_stream stream);

/* residual = fenc - pred for n blocks (pixel_sub_ps_t, adjacent slot; used to build DCT inputs on device):
 * dst block i contiguous at dst + i*w*h. */
int x265b200_residual_batch(x265b200_ctx* ctx, int w, int h, const void* A, intptr_t strideA,
                            const void* B, intptr_t strideB, const int32_t* offA, const int32_t* offB,
                            int n, int16_t* dst, x265b200_stream stream);

/* Inter luma TU reconstruction chain for n TUs of size N in one call: the slot sequence
 * sub_ps -> dct -> quant -> dequant_normal -> (DC-only shortcut | idct) -> add_ps -> sse_pp that
 * reference encoder/search.cpp:5536-5575 drives through quant.cpp:397-480 and :543-605 (no RDOQ / psy / sign hiding /
 * transform skip, scaling lists off).  quantCoeff: N*N table; qBits/add as quant.cpp:465-466; dqScale/dqShift as
 * quant.cpp:556,567.  Outputs: qCoef[n*N*N], numSig[n], recon blocks at recon + offR[i] (the prediction when
 * numSig == 0), sseZero[n] = sse(fenc, pred) (may be NULL) and sseRecon[n] = sse(fenc, recon).  sseZero is the only
 * pointer that may be NULL: offF, offP, offR and sseRecon are required (ERR_ARG otherwise).
 * Default path (x265b200_set_dct_path 0): one tcgen05 / tensor-memory kernel for N = 32, two fused mma.sync kernels for N = 16 / 8 / 4, no
 * scratch memory.  Path 1 (validation twin) runs the stage kernels of the batched primitives over chunks whose intermediates live in an
 * L2-sized scratch taken from the stream-ordered allocator; path 2 = the mma.sync pair for every size; path 3 = tcgen05 for N = 16 too. */
int x265b200_tu_chain_batch(x265b200_ctx* ctx, int N, const void* fenc, intptr_t strideF, const void* pred, intptr_t strideP,
                            const int32_t* offF, const int32_t* offP, int n,
                            const int32_t* quantCoeff, int qBits, int add, int dqScale, int dqShift,
                            int16_t* qCoef, uint32_t* numSig, void* recon, intptr_t strideR, const int32_t* offR,
                            uint64_t* sseZero, uint64_t* sseRecon, x265b200_stream stream);

/* The same chain for the other TU classes of the residual quadtree (the planes, strides and N are the caller's):
 *   X265B200_TU_INTER       inter luma -- and every chroma TU: the chroma loop of Search::estimateResidualQT (reference
 *                           encoder/search.cpp:5638-5700) issues the identical slot sequence on the Cb / Cr planes with log2TrSizeC
 *                           (4:2:2: the two vertically stacked sub-TUs are two TUs of the batch), chroma QP in quantCoeff / qBits / dqScale;
 *   X265B200_TU_INTRA_LUMA  intra luma (Search::codeIntraLumaQT, search.cpp:327-390 residual path): a 4x4 TU takes the DST-VII pair
 *                           dst4x4 / idst4x4 instead of the DCT (reference common/quant.cpp:430-441) and never the DC-only shortcut
 *                           (quant.cpp:585-588 `useDST`); 8x8 ... 32x32 are identical to X265B200_TU_INTER.  `add` is the caller's
 *                           rounding offset (171 << (qBits - 9) for intra slices, quant.cpp:466).
 * x265b200_tu_chain_batch(..) == x265b200_tu_chain_tt_batch(.., X265B200_TU_INTER, ..). */
enum { X265B200_TU_INTER = 0, X265B200_TU_INTRA_LUMA = 1 };
int x265b200_tu_chain_tt_batch(x265b200_ctx* ctx, int N, int ttype, const void* fenc, intptr_t strideF, const void* pred, intptr_t strideP,
                               const int32_t* offF, const int32_t* offP, int n,
                               const int32_t* quantCoeff, int qBits, int add, int dqScale, int dqShift,
                               int16_t* qCoef, uint32_t* numSig, void* recon, intptr_t strideR, const int32_t* offR,
                               uint64_t* sseZero, uint64_t* sseRecon, x265b200_stream stream);

/* Sub-pel candidate cost, interpolation fused with the metric (reference encoder/motion.cpp:1780-1821,
 * MotionEstimate::subpelCompare, luma part): for candidate i of block i / K (n blocks, K candidates each),
 *   cost[i] = cmp(fenc + offF[i / K], strideF, interp(ref + offR[i], xFrac, yFrac), w)
 * where interp is a copy / luma_hpp / luma_vpp / luma_hvpp exactly as subpelCompare selects them; frac[i] = xFrac | yFrac << 4
 * (0..3 each) and offR[i] already contains the integer part (qmv >> 2) of the candidate.  op = X265B200_SAD or X265B200_SATD.
 * The interpolated block stays on chip. */
int x265b200_subpel_cmp_batch(x265b200_ctx* ctx, int op, int w, int h, const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR,
                              const int32_t* offF, const int32_t* offR, const int32_t* frac, int K, int n, int32_t* cost,
                              x265b200_stream stream);

/* Exhaustive integer motion search (reference encoder/motion.cpp:1593-1637, the X265_FULL_SEARCH case of
 * MotionEstimate::motionEstimate) for n prediction units of w x h (4..64, multiples of 4) at once.  For PU i every
 * full-pel vector (x, y) with range[4i] <= x <= range[4i+2], range[4i+1] <= y <= range[4i+3] is costed as
 *   sad(fenc + offF[i], ref + offR[i] + y * strideR + x) + (uint16_t)(costTab[(x << 2) - mvp[2i]] + costTab[(y << 2) - mvp[2i+1]])
 * (BitCost::mvcost, encoder/bitcost.h:53-56: costTab is the DEVICE copy of the lambda-scaled table, pointing at its
 * centre element; mvp is in quarter pels) and (bmv[2i..2i+1], bcost[i]) -- in/out, the search's starting point -- is
 * replaced by the cheapest candidate if that is strictly cheaper; among equal candidates the first in raster order
 * wins, exactly as the reference's COPY2_IF_LT sequence leaves it.  offR[i] addresses the co-located block (vector 0,0);
 * the caller clips range to the padded picture as MotionEstimate::setSearchRange does.  An empty range leaves PU i untouched.
 * merange (the encoder's searchRange) only sizes the on-chip window staging: windows wider than 2 * merange + 1 are still
 * searched completely, in several pieces. */
int x265b200_me_full_batch(x265b200_ctx* ctx, int w, int h, int merange, const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR,
                           const int32_t* offF, const int32_t* offR, const int32_t* range, const int32_t* mvp,
                           const uint16_t* costTab, int n, int32_t* bmv, int32_t* bcost, x265b200_stream stream);

/* search methods, numbered as x265.h:511-519 (X265_DIA_SEARCH .. X265_FULL_SEARCH) */
enum { X265B200_ME_DIA = 0, X265B200_ME_HEX = 1, X265B200_ME_UMH = 2, X265B200_ME_STAR = 3, X265B200_ME_SEA = 4, X265B200_ME_FULL = 5 };

/* The data-dependent integer searches for n PUs at once, one warp walking each PU: diamond (reference
 * encoder/motion.cpp:1016-1039), hexagon + square refinement (:1041-1138) and star (:386-630, :1327-1435, with its
 * stride-5 raster pass and the mvcost(mv << 3) it charges every fourth column), replaying the reference's decision
 * sequence step by step (same candidate order, strict-less updates, only the candidate's row range-checked, the walk
 * ending when its centre leaves the window or after merange / merange/2 - 1 steps).  Arguments as in
 * x265b200_me_full_batch; merange here is the reference's step budget, not a hint.  The window must be padded by two
 * samples horizontally, as the reference's planes are. */
int x265b200_me_pattern_batch(x265b200_ctx* ctx, int method, int w, int h, int merange, const void* fenc, intptr_t strideF,
                              const void* ref, intptr_t strideR, const int32_t* offF, const int32_t* offR, const int32_t* range,
                              const int32_t* mvp, const uint16_t* costTab, int n, int32_t* bmv, int32_t* bcost, x265b200_stream stream);

/* Uneven multi-hexagon search (reference encoder/motion.cpp:1142-1324, X265_UMH_SEARCH) followed by the hexagon search it falls into,
 * one warp per PU.  qmvp: the predictor in quarter pels (its clipped full-pel rounding centres the first diamond); mvc: numCand (0..16)
 * neighbour vectors per PU, quarter pels, whose disagreement scales the search range (motion.cpp:1211-1245).  Other arguments as
 * x265b200_me_pattern_batch. */
int x265b200_me_umh_batch(x265b200_ctx* ctx, int w, int h, int merange, const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR,
                          const int32_t* offF, const int32_t* offR, const int32_t* range, const int32_t* qmvp, int numCand, const int32_t* mvc,
                          const uint16_t* costTab, int n, int32_t* bmv, int32_t* bcost, x265b200_stream stream);

/* Successive elimination (reference encoder/motion.cpp:1438-1591, X265_SEA): every row of the window bmv +- merange (clipped to range) is
 * filtered with ads over the integral planes of the reference picture, survivors get a SAD; the reference's cost bookkeeping is kept as
 * it is (see csrc/mesearch.cu).  sums: the twelve planes x265b200_me_integral_batch wrote for the picture `ref` points into (plane k at
 * sums + k * planePitch, same stride, addressed with offR like the picture); costTab must cover indices down to -2 * |qmvp| around the
 * window, as the reference's BitCost tables do.  PU shapes 32x8, 8x32, 8x4, 4x8 are refused: the reference reads stale cache samples
 * for them. */
int x265b200_me_sea_batch(x265b200_ctx* ctx, int w, int h, int merange, const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR,
                          const int32_t* offF, const int32_t* offR, const int32_t* range, const int32_t* qmvp, const uint16_t* costTab,
                          const uint32_t* sums, size_t planePitch, int n, int32_t* bmv, int32_t* bcost, x265b200_stream stream);

/* Whole MotionEstimate::motionEstimate (reference encoder/motion.cpp:923-1773) with searchMethod DIA, HEX, UMH, STAR or FULL (SEA: pass the
 * integral planes through x265b200_motion_estimate_sea_batch) for n
 * PUs of w x h on full-resolution luma planes: SAD at the clipped predictor qmvp / its full-pel rounding / the zero vector
 * and at numCand (0..16, the same count for every PU; pad with 0,0) neighbour vectors mvc[(i * numCand + k) * 2 ..] in
 * quarter pels, the integer search (x265b200_me_pattern_batch or x265b200_me_full_batch), then the half-pel / quarter-pel refinement of
 * SubpelWorkload[subpelRefine] (0..7, motion.cpp:48-58) through the fused interpolation + SAD/SATD kernels, and the
 * zero vector's last chance.  outQMv[2i..2i+1] (quarter pel) and outCost[i] are what the reference returns in outQMv and
 * as its result, including the early exits on zero residual.  Luma only, one slice (the lookahead-style setSourcePU,
 * motion.cpp:166-189).  range, qmvp, costTab, offR as in x265b200_me_full_batch; plane strides multiples of 4.
 * The batch advances in lock step, one launch per step over all PUs; per-PU decisions live in stream-ordered scratch. */
int x265b200_motion_estimate_batch(x265b200_ctx* ctx, int searchMethod, int w, int h, int merange, int subpelRefine,
                                   const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR,
                                   const int32_t* offF, const int32_t* offR, const int32_t* range, const int32_t* qmvp,
                                   int numCand, const int32_t* mvc, const uint16_t* costTab, int n,
                                   int32_t* outQMv, int32_t* outCost, x265b200_stream stream);

/* x265b200_motion_estimate_batch with searchMethod X265_SEA (reference encoder/motion.cpp:1438-1591): sums / planePitch are the twelve
 * integral planes of the reference picture as x265b200_me_integral_batch writes them (see x265b200_me_sea_batch). */
int x265b200_motion_estimate_sea_batch(x265b200_ctx* ctx, int w, int h, int merange, int subpelRefine,
                                       const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR,
                                       const int32_t* offF, const int32_t* offR, const int32_t* range, const int32_t* qmvp,
                                       int numCand, const int32_t* mvc, const uint16_t* costTab,
                                       const uint32_t* sums, size_t planePitch, int n,
                                       int32_t* outQMv, int32_t* outCost, x265b200_stream stream);

/* The lookahead's motionEstimate (reference encoder/slicetype.cpp:4484-4566 -> motion.cpp:923-1773 with ref->isLowres): the
 * reference picture is a lowres frame = four half-pel planes (full-pel, half-pel x, half-pel y, half-pel xy, as
 * x265b200_lowres_batch / frameInitLowres writes them) planePitch samples apart starting at `planes`; a quarter-pel vector
 * is costed on the rounded average of the two nearest planes (ReferencePlanes::lowresQPelCost, common/lowres.h:95-119),
 * refinement is one SAD half-pel step, a SATD re-measure and one SATD quarter-pel step (motion.cpp:1667-1698), and there
 * are no neighbour candidates.  Other arguments and outputs as x265b200_motion_estimate_batch.  The reference runs this
 * on 8x8 blocks (its averaging buffer is 8x8); other sizes follow the same definition. */
int x265b200_lowres_motion_estimate_batch(x265b200_ctx* ctx, int searchMethod, int w, int h, int merange, int subpelRefine,
                                          const void* fenc, intptr_t strideF, const void* planes, intptr_t strideR, size_t planePitch,
                                          const int32_t* offF, const int32_t* offR, const int32_t* range, const int32_t* qmvp,
                                          const uint16_t* costTab, int n, int32_t* outQMv, int32_t* outCost, x265b200_stream stream);

/* Chroma term of MotionEstimate::subpelCompare (reference encoder/motion.cpp:1805-1865): SATD of the 4-tap interpolated chroma
 * block against the chroma fenc block, fused like x265b200_subpel_cmp_batch.  w x h is the CHROMA block; frac[i] =
 * xFrac | yFrac << 4 in eighths (0..7 each), a negative frac[i] contributes nothing; accumulate != 0 adds onto cost[i]
 * (call it for Cb and Cr after the luma entry), else cost[i] is overwritten. */
int x265b200_subpel_cmp_chroma_batch(x265b200_ctx* ctx, int w, int h, const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR,
                                     const int32_t* offF, const int32_t* offR, const int32_t* frac, int K, int n, int32_t* cost,
                                     int accumulate, x265b200_stream stream);

/* x265b200_motion_estimate_batch as the encoder proper runs it (setSourcePU of motion.cpp:222-247 with bChroma): from
 * subpelRefine 3 on, every sub-pel cost -- the predictor candidates, the refinement rounds, the zero vector's last chance
 * -- also charges the SATD of the Cb and Cr blocks at the vector scaled to chroma (bChromaSATD), provided the chroma block
 * is a multiple of 4x4 (the reference's non-NULL chroma satd slots); otherwise, and below subme 3, the result equals the
 * luma-only entry.  offFC / offRC: the PU's co-located block in the Cb / Cr planes (both planes share strides and offsets);
 * hshift, vshift = 1, 1 (4:2:0) or 0, 0 (4:4:4).  w x h is the luma PU. */
int x265b200_motion_estimate_chroma_batch(x265b200_ctx* ctx, int searchMethod, int w, int h, int merange, int subpelRefine,
                                          const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR,
                                          const int32_t* offF, const int32_t* offR,
                                          const void* fencCb, const void* fencCr, intptr_t strideFC,
                                          const void* refCb, const void* refCr, intptr_t strideRC,
                                          const int32_t* offFC, const int32_t* offRC, int hshift, int vshift,
                                          const int32_t* range, const int32_t* qmvp, int numCand, const int32_t* mvc,
                                          const uint16_t* costTab, int n, int32_t* outQMv, int32_t* outCost, x265b200_stream stream);

/* Bi-prediction candidate cost (reference encoder/search.cpp:442-448 in Search::predInterSearch): for PU i, both lists'
 * motion-compensated luma blocks (Predict::predInterLumaPixel, common/predict.cpp:279-300; off0/off1 hold the integer part
 * of each vector, frac0/frac1 = xFrac | yFrac << 4), their rounded average (pixelavg_pp) and SATD against the fenc block;
 * neither prediction nor average is written to memory. */
int x265b200_bidir_satd_batch(x265b200_ctx* ctx, int w, int h, const void* fenc, intptr_t strideF, const int32_t* offF,
                              const void* ref0, intptr_t stride0, const int32_t* off0, const int32_t* frac0,
                              const void* ref1, intptr_t stride1, const int32_t* off1, const int32_t* frac1,
                              int n, int32_t* cost, x265b200_stream stream);

/* The lookahead's intra estimate (reference encoder/slicetype.cpp:755-864, LookaheadTLD::lowresIntraEstimate) for every 8x8
 * CU of a lowres frame: neighbours taken from the plane itself (the plane must be padded, as lowres planes are), 1:2:1
 * smoothing, DC / planar / the angular modes searched coarse to fine exactly as the reference does, 8x8 SATD.
 * `plane` points at the picture's first sample.  cost[cuY * widthInCU + cuX] = the reference's fenc.intraCost (icost +
 * penalty, penalty = intraPenalty + lowresPenalty), mode[..] = fenc.intraMode.  The per-frame sums and the AQ-weighted
 * costs (slicetype.cpp:843-863) are reductions over this array and stay with the caller. */
int x265b200_lowres_intra_batch(x265b200_ctx* ctx, const void* plane, intptr_t stride, int widthInCU, int heightInCU, int penalty,
                                int32_t* cost, int32_t* mode, x265b200_stream stream);

/* Predictor selection of the lookahead (reference encoder/slicetype.cpp:4520-4558, inside CostEstimateGroup::estimateCUCost) for n 8x8 CUs: CU i has
 * numc[i] (0..5) candidate vectors mvc[(5 * i + k) * 2 ..] in quarter pels (its already-searched neighbours' vectors); each is costed by the 8x8 SATD of
 * the fenc block against ReferencePlanes::lowresMC of the vector (common/lowres.h:74-93: a half-pel plane of the lowres reference, or the rounded average
 * of the two nearest ones; planes / planePitch as in x265b200_lowres_motion_estimate_batch, offR = the CU's co-located block).  mvp[2i..] = the cheapest
 * (the first of equal costs), 0 when numc[i] == 0; mvpCost[i] = its cost (1 << 28 = MotionEstimate::COST_MAX when nothing was measured); skipCost[i] =
 * the reference's skipCost (INT_MAX unless, in a B frame, a candidate was measured while the running predictor was the zero vector). */
int x265b200_lowres_mvp_batch(x265b200_ctx* ctx, const void* fenc, intptr_t strideF, const int32_t* offF, const void* planes, intptr_t strideR,
                              size_t planePitch, const int32_t* offR, const int32_t* mvc, const int32_t* numc, int bBidir, int n,
                              int32_t* mvp, int32_t* mvpCost, int32_t* skipCost, x265b200_stream stream);

/* The bi-directional candidates of a B-frame CU (reference encoder/slicetype.cpp:4577-4596): cost[2i] = 8x8 SATD of fenc against the rounded average of
 * both lists' motion-compensated blocks (lowresMC at mv0[2i..] in reference 0 and mv1[2i..] in reference 1), cost[2i + 1] = against the average of the two
 * co-located full-pel blocks.  The caller adds lowresPenalty and takes the minimum with the list costs (COPY2_IF_LT, listused = 3). */
int x265b200_lowres_bidir_cost_batch(x265b200_ctx* ctx, const void* fenc, intptr_t strideF, const int32_t* offF,
                                     const void* planes0, intptr_t stride0, size_t planePitch0, const void* planes1, intptr_t stride1, size_t planePitch1,
                                     const int32_t* offR, const int32_t* mv0, const int32_t* mv1, int n, int32_t* cost, x265b200_stream stream);

/* All 35 luma intra predictions of n TUs of N x N (4, 8, 16, 32) as the analysis forms them before costing the modes
 * (reference encoder/search.cpp:1703-1727 on common/intrapred.cpp): neighbours[i * (4N+1) ..] = top-left, 2N above, 2N left
 * (unfiltered; the 1:2:1 smoothed copy is made on chip), DC with edge smoothing for N <= 16, planar from the smoothed
 * samples for N >= 8, angular modes per g_intraFilterFlags.  dst[(i * 35 + mode) * N * N ..] = the N x N prediction, every
 * mode in picture orientation.  Feed dst to x265b200_pixelcmp_batch (sa8d / satd) for the mode costs. */
int x265b200_intra_pred_batch(x265b200_ctx* ctx, int N, const void* neighbours, int n, void* dst, x265b200_stream stream);

/* The table's three intra slots over n TUs of N x N (4 .. 32), each TU with its own (4N + 1)-sample neighbour array (top-left, 2N above, 2N left;
 * arrays contiguous): kind 0 = intra_pred[mode] (intra_pred_t, reference common/intrapred.cpp:66-204: planar 0, DC 1, angular 2 .. 34; bFilter
 * as the slot's argument) -> dst[i * N * N ..]; kind 1 = intra_filter (intrapred.cpp:31-51) -> dst[i * (4N + 1) ..]; kind 2 =
 * intra_pred_allangs (intrapred.cpp:206-233: modes 2 .. 34, `filt` = the smoothed neighbours, horizontal modes un-flipped, bFilter = bLuma)
 * -> dst[i * 33 * N * N ..].  filt is only read by kind 2. */
int x265b200_intra_slot_batch(x265b200_ctx* ctx, int kind, int N, int mode, int bFilter, const void* src, const void* filt, int n, void* dst,
                              x265b200_stream stream);

/* D block i = op(A block i, B block i) for n blocks of w x h (1..64 each); element types follow the op (see the enum).
 * An offset array may be NULL: blocks are then contiguous (block i at i * w * h, use stride = w). */
int x265b200_blockop_batch(x265b200_ctx* ctx, int op, int w, int h, const void* A, intptr_t strideA, const int32_t* offA,
                           const void* B, intptr_t strideB, const int32_t* offB, void* D, intptr_t strideD, const int32_t* offD,
                           int n, x265b200_stream stream);

/* copy family over n blocks of w x h (offset arrays may be NULL = contiguous blocks).  kind: 0 copy_pp (pixel.cpp:751-762),
 * 1 copy_ss, 2 copy_sp ((pixel) cast), 3 copy_ps (pixel.cpp:764-804), 4 blockfill_s (param = value, src unused; pixel.cpp:385-391),
 * 5 shift left: (int16)((uint32)x << param) = cpy2Dto1D_shl / cpy1Dto2D_shl (pixel.cpp:393-408, 428-443; the 1-D side uses
 * stride = w), 6 rounding shift right: (x + (1 << (param - 1))) >> param = cpy2Dto1D_shr / cpy1Dto2D_shr (pixel.cpp:410-461). */
int x265b200_blockcopy_batch(x265b200_ctx* ctx, int kind, int w, int h, const void* src, intptr_t srcStride, const int32_t* offS,
                             void* dst, intptr_t dstStride, const int32_t* offD, int n, int param, x265b200_stream stream);

/* per-block scalars (offset arrays may be NULL = contiguous size x size blocks); size in {4, 8, 16, 32, 64}:
 * var (pixel.cpp:695-712): out[i] = sum | (uint64)sumsq << 32, both uint32 and wrapping like the reference;
 * psy_cost_pp (pixel.cpp:718-749): |AC energy(source) - AC energy(recon)| summed over the 8x8 sub-blocks (4x4: satd based);
 * count_nonzero / copy_cnt (dct.cpp:716-744): count[i] = non-zero int16 in block i; coeff != NULL additionally receives the
 *   blocks contiguously (copy_cnt); denoiseDct (dct.cpp:746-757) over n blocks of numCoeff sharing one resSum / offset table
 *   (resSum[i] accumulates |level| of every block, order-independent). */
int x265b200_var_batch(x265b200_ctx* ctx, int size, const void* pix, intptr_t stride, const int32_t* off, int n, uint64_t* out, x265b200_stream stream);
int x265b200_psy_cost_batch(x265b200_ctx* ctx, int size, const void* src, intptr_t srcStride, const int32_t* offS, const void* rec, intptr_t recStride,
                            const int32_t* offR, int n, int32_t* out, x265b200_stream stream);
int x265b200_count_nonzero_batch(x265b200_ctx* ctx, int size, const int16_t* src, intptr_t stride, const int32_t* off, int n,
                                 int16_t* coeff, uint32_t* count, x265b200_stream stream);
int x265b200_denoise_dct_batch(x265b200_ctx* ctx, int16_t* dct, uint32_t* resSum, const uint16_t* offset, int numCoeff, int n, x265b200_stream stream);

/* downscale_t frameInitLowres (pixel.cpp:595-620) over one plane: width x height are the LOWRES dimensions; reads
 * 2 * width + 1 columns of 2 * height + 1 rows of src, writes the four half-resolution planes. */
int x265b200_lowres_batch(x265b200_ctx* ctx, const void* src, intptr_t srcStride, void* dst0, void* dsth, void* dstv, void* dstc,
                          intptr_t dstStride, int width, int height, x265b200_stream stream);

/* SEA integral planes, the input of `ads` (encoder/framefilter.cpp:38-140 driven by FrameFilter::computeMEIntegral,
 * framefilter.cpp:737-835) for nframes padded pictures of `rows` x `stride` samples stored back to back.  Output: twelve
 * uint32 planes per picture in the reference's order (32x32, 32x24, 32x8, 24x32, 16x16, 16x12, 16x4, 12x16, 8x32, 8x8, 4x16,
 * 4x4), plane k of picture f at sums + (f * 12 + k) * planePitch, same stride as the picture:
 *   sum[r][x] = sum of the W x H box with top-left sample (x, r)   for 1 <= r <= rows - 1 - H, x < stride - W,
 * 0 elsewhere (the reference leaves prefix sums / uninitialised memory there; the search never reads them). */
int x265b200_me_integral_batch(x265b200_ctx* ctx, const void* pix, intptr_t stride, int rows, int nframes,
                               uint32_t* sums, size_t planePitch, x265b200_stream stream);
/* one row of integral_init{4..32}h (vertical = 0: out[x] = hsum_size(pix, x) + a[x]) or ..v (vertical = 1: out[x] = b[x] - a[x]) */
int x265b200_integral_row_batch(x265b200_ctx* ctx, int vertical, int size, const void* pix, const uint32_t* a, const uint32_t* b,
                                uint32_t* out, int count, x265b200_stream stream);

/* weightp_pp_t / weightp_sp_t (pixel.cpp:485-535) over a width x height plane region; sp = 0: pixel source, sp = 1: int16 source */
int x265b200_weight_batch(x265b200_ctx* ctx, int sp, const void* src, intptr_t srcStride, void* dst, intptr_t dstStride,
                          int width, int height, int w0, int round, int shift, int offset, x265b200_stream stream);
/* Lookahead weighted-prediction cost (encoder/slicetype.cpp:866-897 weightCostLuma; weightPrediction.cpp:171-222 luma branch)
 * for K candidate weights in one pass: weights = K x {w0, round, shift, offset} as passed to weight_pp (shift < 0: unweighted);
 * cost[k] = sum over the 8x8 blocks of the width x height picture of min(satd_8x8(weight_k(ref), fenc), intraCost[block])
 * (intraCost may be NULL: no cap).  fenc and ref share `stride` (a multiple of 4 samples). */
int x265b200_weight_cost_batch(x265b200_ctx* ctx, const void* fenc, const void* ref, intptr_t stride, int width, int height,
                               const int32_t* intraCost, const int32_t* weights, int K, uint32_t* cost, x265b200_stream stream);

/* Forward half of the chain alone -- Quant::transformNxN for an inter luma TU (reference common/quant.cpp:397-480: sub_ps ->
 * dct -> quant): qCoef[n*N*N], numSig[n], sseZero[n] = sse(fenc, pred) (may be NULL).  Arguments as x265b200_tu_chain_batch. */
int x265b200_tu_forward_batch(x265b200_ctx* ctx, int N, const void* fenc, intptr_t strideF, const void* pred, intptr_t strideP,
                              const int32_t* offF, const int32_t* offP, int n, const int32_t* quantCoeff, int qBits, int add,
                              int16_t* qCoef, uint32_t* numSig, uint64_t* sseZero, x265b200_stream stream);

/* ------------------------------------------------------------------ host-buffer layer: resident planes and frame jobs
 *
 * What a C++ caller in the encoder (ThreadedME, reference encoder/threadedme.cpp:207-261; the lookahead,
 * encoder/slicetype.cpp:4467) needs between the one-block host slots and the raw device entries: the library owns the
 * device memory, the pinned staging, the copies in both directions and the streams; the caller hands over HOST pictures and
 * reads HOST results.  No CUDA type or call appears on the caller's side.
 */

/* Pinned host memory for pictures the encoder allocates itself (PicYuv::create, reference common/picyuv.cpp:86-118, would call
 * this instead of X265_MALLOC), or page-locking of a buffer that already exists (PicYuv::m_picBuf), so that plane uploads are
 * asynchronous DMA.  Uploading from ordinary pageable memory works too, at the speed of a staged copy. */
void* x265b200_host_alloc(x265b200_ctx* ctx, size_t bytes);
void x265b200_host_free(x265b200_ctx* ctx, void* p);
int x265b200_host_register(x265b200_ctx* ctx, void* p, size_t bytes);
int x265b200_host_unregister(x265b200_ctx* ctx, void* p);

/* A picture plane resident in HBM with the reference's own geometry (PicYuv::create, common/picyuv.cpp:86-118): for a luma plane
 * (hshift = vshift = 0)  stride = ceil(width / ctu) * ctu + 2 * (ctu + 32),  rows = ceil(height / ctu) * ctu + 2 * (ctu + 16),
 * sample (0, 0) at element  origin = (ctu + 16) * stride + ctu + 32;  for a chroma plane the CTU-aligned size is shifted by
 * hshift / vshift, the horizontal margin stays ctu + 32 and the vertical margin is (ctu + 16) >> vshift (picyuv.cpp:106-110).
 * Block descriptors of the batched entries are element offsets from the plane base, e.g. origin + y * stride + x. */
typedef struct x265b200_plane x265b200_plane;
int x265b200_plane_create(x265b200_ctx* ctx, int width, int height, int ctu, int hshift, int vshift, x265b200_plane** plane);
void x265b200_plane_destroy(x265b200_plane* plane);
/* any output pointer may be NULL; *device is the device address of the plane base (for mixing with the batched entries) */
int x265b200_plane_info(const x265b200_plane* plane, intptr_t* stride, int* rows, int32_t* origin, size_t* elems, void** device);
/* Whole padded plane from a host buffer of the same geometry (PicYuv::m_picBuf).  Asynchronous when the buffer is pinned; the
 * plane's users (frame jobs) wait for it on the device, and the copy itself waits for jobs still reading the plane. */
int x265b200_plane_upload_padded(x265b200_plane* plane, const void* hostPlane);
/* width x height samples at hostPic / hostStride (elements) -> the picture area, then the margins are formed on the device
 * exactly as extendPicBorder (reference common/pixel.cpp:1044-1061) leaves them: marginX columns left and right of every
 * picture row, then marginY copies of the first and last padded row.  Cells the reference does not write (rows below
 * height + marginY when height is not a CTU multiple) keep their previous contents (zero after creation). */
int x265b200_plane_upload_picture(x265b200_plane* plane, const void* hostPic, intptr_t hostStride);
/* Same result as x265b200_plane_upload_picture for a host buffer that already has the plane's geometry (PicYuv::m_picBuf: pass its
 * base): the picture's rows travel as one linear copy of whole buffer rows (height * stride samples), the margins are formed on the
 * device.  What the encoder does for a reconstructed picture instead of extendPicBorder + a whole-plane upload. */
int x265b200_plane_upload_rows(x265b200_plane* plane, const void* hostPlane);
/* device -> host of the whole padded plane (for tests and for recon pictures the encoder wants back); synchronous */
int x265b200_plane_download_padded(x265b200_plane* plane, void* hostPlane);
/* cumulative bytes copied host -> device / device -> host by planes and frame jobs of this context */
void x265b200_transfer_stats(const x265b200_ctx* ctx, uint64_t* h2dBytes, uint64_t* d2hBytes);

/* A frame job = the analysis passes run for every frame, registered once with their block descriptors, then executed per
 * (fenc plane, reference plane) pair.  Several frames are in flight (slots); each slot has its own stream, device outputs and
 * pinned result buffers, so the upload of frame k+1 overlaps the kernels of frame k and the download of frame k-1.
 * The call shape follows ThreadedME's per-row task (encoder/threadedme.cpp:207-261: descriptors fixed by the CTU grid, one
 * result record per PU) with the whole frame as the batch.  One thread drives a job; jobs are independent. */
typedef struct x265b200_frame_job x265b200_frame_job;
enum {
    X265B200_PASS_CMP = 0,              /* op(fenc block, ref block) -> int32 cost per block (SAD / SATD / SA8D) */
    X265B200_PASS_COEF = 1,             /* residual + forward DCT -> dense int16 coefficients, N*N per TU */
    X265B200_PASS_LEVELS = 2            /* residual + forward DCT + quant (transformNxN) -> numSig per TU, one significance bit
                                           per coefficient, and the non-zero levels only (in TU order, raster order inside a TU) */
};
typedef struct x265b200_pass_result {
    int kind;                           /* X265B200_PASS_* */
    int n;                              /* blocks / TUs of the pass */
    const int32_t* cost;                /* CMP: n costs */
    const int16_t* coef;                /* COEF: n * N * N coefficients */
    const uint16_t* numSig;             /* LEVELS: n counts */
    const uint32_t* sigMap;             /* LEVELS: bit (i & 31) of word (i >> 5) is set iff coefficient i of the flat n * N * N array is non-zero */
    const int16_t* levels;              /* LEVELS: the non-zero coefficients in flat order; TU t owns numSig[t] of them */
    uint32_t nlevels;                   /* LEVELS: sum of numSig */
} x265b200_pass_result;

/* slots = frames in flight (1..8).  All planes given to x265b200_frame_job_submit must have the geometry
 * (width, height, ctu, 0, 0) -- luma. */
int x265b200_frame_job_create(x265b200_ctx* ctx, int width, int height, int ctu, int slots, x265b200_frame_job** job);
void x265b200_frame_job_destroy(x265b200_frame_job* job);
/* register a pass; offF / offR are HOST arrays of n element offsets (fenc block, reference block incl. the motion vector),
 * copied to the device here.  Returns the pass index (>= 0) or an error.  Passes run in registration order. */
int x265b200_frame_job_add_cmp(x265b200_frame_job* job, int op, int w, int h, const int32_t* offF, const int32_t* offR, int n);
/* kind = X265B200_PASS_COEF (quantCoeff / qBits / add ignored) or X265B200_PASS_LEVELS (quantCoeff: N*N HOST table,
 * qBits / add as Quant::transformNxN computes them, reference common/quant.cpp:465-466) */
int x265b200_frame_job_add_transform(x265b200_frame_job* job, int kind, int N, const int32_t* offF, const int32_t* offR, int n,
                                     const int32_t* quantCoeff, int qBits, int add);
/* new descriptors for a registered pass (same n): the motion vectors of the next frame */
int x265b200_frame_job_set_blocks(x265b200_frame_job* job, int pass, const int32_t* offF, const int32_t* offR);
/* enqueue every pass for one frame pair; returns the slot (>= 0) whose results x265b200_frame_job_wait delivers, or an
 * error (X265B200_ERR_ARG when the next slot in the ring has not been waited for yet).  Returns without waiting for the GPU. */
int x265b200_frame_job_submit(x265b200_frame_job* job, x265b200_plane* fenc, x265b200_plane* ref);
/* blocks until the slot's results are in host memory; results[p] describes pass p (maxPasses entries are filled at most).
 * The pointers address pinned memory owned by the job and stay valid until the slot is submitted again. */
int x265b200_frame_job_wait(x265b200_frame_job* job, int slot, x265b200_pass_result* results, int maxPasses);
int x265b200_frame_job_pass_count(const x265b200_frame_job* job);

/* The ThreadedME contract as one call (reference encoder/threadedme.h:112-130, threadedme.cpp:207-261 -> Analysis::deriveMVsForCTU ->
 * Search::puMotionEstimation, encoder/search.cpp:226-404): n searches, each one PU against one reference picture, given as HOST records and returned
 * as the per-reference part of MEData.  Records may mix all PU shapes, references and candidate counts; the call groups them, runs
 * x265b200_motion_estimate_batch per group on resident planes and finishes every search with search.cpp:392-394's bookkeeping:
 *   bits = pu.bits + bitcost(mv), mvCost = mvcost(mv), cost = (satdCost - mvCost) + ((bits * lambda + 128) >> 8).
 * costTab / bitsTab: HOST pointers to the CENTRE elements of BitCost's tables for the slice QP (s_costs[qp] and s_bitsizes, encoder/bitcost.h), valid for
 * indices -tabRadius .. tabRadius; lambda = RDCost::m_lambda (encoder/rdcost.h:91).  mvmin / mvmax in full pels as motionEstimate receives them after
 * setSearchRange.  Choosing the best reference per list, checkBestMVP and the bi-prediction candidate stay with the caller. */
#define X265B200_TME_MAX_CAND 8
typedef struct x265b200_tme_pu {
    int16_t w, h;                       /* PU size (g_puLookup, threadedme.h:67-92, or any multiple of 4 up to 64) */
    int16_t ref;                        /* index into refPlanes */
    int16_t numCand;                    /* neighbour vectors in mvc */
    int32_t offF, offR;                 /* the PU in the fenc plane, its co-located block in the reference plane (element offsets) */
    int32_t mvmin[2], mvmax[2];         /* search window, full pels */
    int32_t mvp[2];                     /* predictor, quarter pels */
    int32_t mvc[X265B200_TME_MAX_CAND][2];
    uint32_t bits;                      /* list / reference / mvp-index bits charged before the vector's own (search.cpp:269-271) */
} x265b200_tme_pu;
typedef struct x265b200_tme_result {
    int32_t mv[2];                      /* MEData::mv (quarter pels) */
    uint32_t mvCost;                    /* MEData::mvCost */
    uint32_t bits;                      /* MEData::bits contribution of this search */
    uint32_t cost;                      /* MEData::cost candidate */
    uint32_t satdCost;                  /* what motionEstimate returned */
} x265b200_tme_result;
int x265b200_tme_search_batch(x265b200_ctx* ctx, int searchMethod, int merange, int subpelRefine, const x265b200_plane* fencPlane,
                              const x265b200_plane* const* refPlanes, int numRefs, const uint16_t* costTab, const float* bitsTab,
                              int tabRadius, uint64_t lambda, const x265b200_tme_pu* pus, int n, x265b200_tme_result* results);

/* ------------------------------------------------------------------ host (per-call, drop-in slots) */

int x265b200_sad(x265b200_ctx*, int w, int h, const void* fenc, intptr_t fencstride, const void* fref, intptr_t frefstride);
int x265b200_satd(x265b200_ctx*, int w, int h, const void* fenc, intptr_t fencstride, const void* fref, intptr_t frefstride);
int x265b200_sa8d(x265b200_ctx*, int w, int h, const void* fenc, intptr_t fencstride, const void* fref, intptr_t frefstride);
uint64_t x265b200_sse_pp(x265b200_ctx*, int w, int h, const void* fenc, intptr_t fencstride, const void* fref, intptr_t frefstride);
uint64_t x265b200_sse_ss(x265b200_ctx*, int w, int h, const int16_t* fenc, intptr_t fencstride, const int16_t* fref, intptr_t frefstride);
uint64_t x265b200_ssd_s(x265b200_ctx*, int size, const int16_t* a, intptr_t stride);
/* fenc stride is FENC_STRIDE = 64 (common.h:71) as in the reference */
void x265b200_sad_x3(x265b200_ctx*, int w, int h, const void* fenc, const void* fref0, const void* fref1, const void* fref2, intptr_t frefstride, int32_t* res);
void x265b200_sad_x4(x265b200_ctx*, int w, int h, const void* fenc, const void* fref0, const void* fref1, const void* fref2, const void* fref3, intptr_t frefstride, int32_t* res);
int x265b200_ads(x265b200_ctx*, int w, int h, const int* encDC, const uint32_t* sums, int delta, const uint16_t* costMvX, int16_t* mvs, int width, int thresh);

void x265b200_dct(x265b200_ctx*, int kind, int N, const int16_t* src, int16_t* dst, intptr_t srcStride);
void x265b200_idct(x265b200_ctx*, int kind, int N, const int16_t* src, int16_t* dst, intptr_t dstStride);
uint32_t x265b200_quant(x265b200_ctx*, const int16_t* coef, const int32_t* quantCoeff, int32_t* deltaU, int16_t* qCoef, int qBits, int add, int numCoeff);
uint32_t x265b200_nquant(x265b200_ctx*, const int16_t* coef, const int32_t* quantCoeff, int16_t* qCoef, int qBits, int add, int numCoeff);
void x265b200_dequant_normal(x265b200_ctx*, const int16_t* quantCoef, int16_t* coef, int num, int scale, int shift);
void x265b200_dequant_scaling(x265b200_ctx*, const int16_t* src, const int32_t* dequantCoef, int16_t* dst, int num, int mcqp_miper, int shift);

/* kind = X265B200_IP_*; extra = isRowExt for HPS, idxY for HVPP (coeffIdx is then idxX), else ignored */
void x265b200_interp(x265b200_ctx*, int kind, int taps, int w, int h, const void* src, intptr_t srcStride,
                     void* dst, intptr_t dstStride, int coeffIdx, int extra);

/* adjacent slots: argument order of primitives.h:189-192 and :168 */
void x265b200_sub_ps(x265b200_ctx*, int w, int h, int16_t* dst, intptr_t dstride, const void* src0, const void* src1, intptr_t sstride0, intptr_t sstride1);
void x265b200_add_ps(x265b200_ctx*, int w, int h, void* dst, intptr_t dstride, const void* src0, const int16_t* src1, intptr_t sstride0, intptr_t sstride1);
void x265b200_pixelavg_pp(x265b200_ctx*, int w, int h, void* dst, intptr_t dstride, const void* src0, intptr_t sstride0, const void* src1, intptr_t sstride1, int weight);
void x265b200_addAvg(x265b200_ctx*, int w, int h, const int16_t* src0, const int16_t* src1, void* dst, intptr_t src0Stride, intptr_t src1Stride, intptr_t dstStride);
void x265b200_frame_init_lowres(x265b200_ctx*, const void* src0, void* dst0, void* dsth, void* dstv, void* dstc,
                                intptr_t srcStride, intptr_t dstStride, int width, int height);
/* copy_pp/ss/sp/ps, blockfill_s, cpy2Dto1D_shl/shr, cpy1Dto2D_shl/shr (primitives.h:141-150, 184-187); kind as in x265b200_blockcopy_batch */
void x265b200_blockcopy(x265b200_ctx*, int kind, int w, int h, void* dst, intptr_t dstStride, const void* src, intptr_t srcStride, int param);
/* var_t, pixelcmp_t psy_cost_pp, count_nonzero_t / copy_cnt_t (coeff == NULL: count only), denoiseDct_t with the size prepended */
uint64_t x265b200_var(x265b200_ctx*, int size, const void* pix, intptr_t stride);
int x265b200_psy_cost_pp(x265b200_ctx*, int size, const void* source, intptr_t sstride, const void* recon, intptr_t rstride);
uint32_t x265b200_copy_cnt(x265b200_ctx*, int size, int16_t* coeff, const int16_t* residual, intptr_t resiStride);
void x265b200_denoise_dct(x265b200_ctx*, int16_t* dctCoef, uint32_t* resSum, const uint16_t* offset, int numCoeff);
/* integralh_t / integralv_t (primitives.h:227-228) with the box width / height prepended */
void x265b200_weight_pp(x265b200_ctx*, const void* src, void* dst, intptr_t stride, int width, int height, int w0, int round, int shift, int offset);
void x265b200_weight_sp(x265b200_ctx*, const int16_t* src, void* dst, intptr_t srcStride, intptr_t dstStride, int width, int height, int w0, int round, int shift, int offset);
void x265b200_integral_inith(x265b200_ctx*, int W, uint32_t* sum, const void* pix, intptr_t stride);
void x265b200_integral_initv(x265b200_ctx*, int H, uint32_t* sum, intptr_t stride);

/* intra_pred_t / intra_filter_t / intra_allangs_t (primitives.h:143-145) with the TU size (and the mode, which the reference passes
 * as dirMode and also bakes into the slot index) prepended */
void x265b200_intra_pred(x265b200_ctx*, int N, int mode, void* dst, intptr_t dstStride, const void* srcPix, int bFilter);
void x265b200_intra_filter(x265b200_ctx*, int N, const void* samples, void* filtered);
void x265b200_intra_pred_allangs(x265b200_ctx*, int N, void* dst, const void* refPix, const void* filtPix, int bLuma);

#ifdef __cplusplus
}
#endif
#endif /* X265B200_H */
