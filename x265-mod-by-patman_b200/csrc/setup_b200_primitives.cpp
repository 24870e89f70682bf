// setup_b200_primitives.cpp -- binds x265's EncoderPrimitives slots to the B200 C ABI.
//
// This is the file a maintainer would add to the reference tree (e.g. source/common/b200/) next to
// the AArch64 intrinsics filler it is modelled on (reference common/aarch64/asm-primitives.cpp:697-720):
// it includes the reference's own primitives.h, so it is compiled once per bit depth / namespace
// exactly like libx265 itself, and is called between setupAssemblyPrimitives() and
// setupAliasPrimitives() (reference primitives.cpp:355-367).  See INTEGRATION.md.
//
// Every hot-path slot gets a thunk with the exact typedef'd signature
// (primitives.h:133-182) that forwards to the per-call host entry of include/x265b200.h with the
// block size baked in as a template argument -- the same way the reference bakes it into
// sad<W, H> etc.  Slots outside the hot path are left untouched.
#include "common.h"
#include "primitives.h"
#include "x265b200.h"
#include "x265b200_glue.h"
#include <mutex>

#include <stdio.h>
#include <stdlib.h>

namespace X265_NS {

static x265b200_ctx* g_b200;    // one context per process and depth, like the global `primitives` table

namespace {

// ---- metrics ----
template<int W, int H> int t_sad(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb) { return x265b200_sad(g_b200, W, H, a, sa, b, sb); }
template<int W, int H> int t_satd(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb) { return x265b200_satd(g_b200, W, H, a, sa, b, sb); }
template<int W, int H> int t_sa8d(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb) { return x265b200_sa8d(g_b200, W, H, a, sa, b, sb); }
template<int W, int H> sse_t t_sse_pp(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb) { return (sse_t)x265b200_sse_pp(g_b200, W, H, a, sa, b, sb); }
template<int W, int H> sse_t t_sse_ss(const int16_t* a, intptr_t sa, const int16_t* b, intptr_t sb) { return (sse_t)x265b200_sse_ss(g_b200, W, H, a, sa, b, sb); }
template<int W> sse_t t_ssd_s(const int16_t* a, intptr_t sa) { return (sse_t)x265b200_ssd_s(g_b200, W, a, sa); }
template<int W, int H> void t_sad_x3(const pixel* f, const pixel* r0, const pixel* r1, const pixel* r2, intptr_t rs, int32_t* res)
{ x265b200_sad_x3(g_b200, W, H, f, r0, r1, r2, rs, res); }
template<int W, int H> void t_sad_x4(const pixel* f, const pixel* r0, const pixel* r1, const pixel* r2, const pixel* r3, intptr_t rs, int32_t* res)
{ x265b200_sad_x4(g_b200, W, H, f, r0, r1, r2, r3, rs, res); }
template<int W, int H> int t_ads(int encDC[], uint32_t* sums, int delta, uint16_t* costMvX, int16_t* mvs, int width, int thresh)
{ return x265b200_ads(g_b200, W, H, encDC, sums, delta, costMvX, mvs, width, thresh); }

// ---- transforms ----
template<int N> void t_dct(const int16_t* s, int16_t* d, intptr_t st) { x265b200_dct(g_b200, X265B200_TR_DCT, N, s, d, st); }
template<int N> void t_idct(const int16_t* s, int16_t* d, intptr_t st) { x265b200_idct(g_b200, X265B200_TR_DCT, N, s, d, st); }
template<int N> void t_lowpass(const int16_t* s, int16_t* d, intptr_t st) { x265b200_dct(g_b200, X265B200_TR_LOWPASS, N, s, d, st); }
void t_dst4(const int16_t* s, int16_t* d, intptr_t st) { x265b200_dct(g_b200, X265B200_TR_DST, 4, s, d, st); }
void t_idst4(const int16_t* s, int16_t* d, intptr_t st) { x265b200_idct(g_b200, X265B200_TR_DST, 4, s, d, st); }
uint32_t t_quant(const int16_t* c, const int32_t* q, int32_t* du, int16_t* qc, int qBits, int add, int n) { return x265b200_quant(g_b200, c, q, du, qc, qBits, add, n); }
uint32_t t_nquant(const int16_t* c, const int32_t* q, int16_t* qc, int qBits, int add, int n) { return x265b200_nquant(g_b200, c, q, qc, qBits, add, n); }
void t_dequant_normal(const int16_t* q, int16_t* c, int num, int scale, int shift) { x265b200_dequant_normal(g_b200, q, c, num, scale, shift); }
void t_dequant_scaling(const int16_t* q, const int32_t* dq, int16_t* c, int num, int per, int shift) { x265b200_dequant_scaling(g_b200, q, dq, c, num, per, shift); }

// ---- intra prediction (adjacent slots: the analysis' mode scan and the lookahead's intra estimate call them) ----
template<int N> void t_intra_pred(pixel* dst, intptr_t dstStride, const pixel* srcPix, int dirMode, int bFilter) { x265b200_intra_pred(g_b200, N, dirMode, dst, dstStride, srcPix, bFilter); }
template<int N> void t_intra_planar(pixel* dst, intptr_t dstStride, const pixel* srcPix, int, int bFilter) { x265b200_intra_pred(g_b200, N, 0, dst, dstStride, srcPix, bFilter); }
template<int N> void t_intra_dc(pixel* dst, intptr_t dstStride, const pixel* srcPix, int, int bFilter) { x265b200_intra_pred(g_b200, N, 1, dst, dstStride, srcPix, bFilter); }
template<int N> void t_intra_filter(const pixel* samples, pixel* filtered) { x265b200_intra_filter(g_b200, N, samples, filtered); }
template<int N> void t_intra_allangs(pixel* dst, pixel* refPix, pixel* filtPix, int bLuma) { x265b200_intra_pred_allangs(g_b200, N, dst, refPix, filtPix, bLuma); }

// ---- interpolation ----
template<int T, int W, int H> void t_hpp(const pixel* s, intptr_t ss, pixel* d, intptr_t ds, int idx) { x265b200_interp(g_b200, X265B200_IP_HPP, T, W, H, s, ss, d, ds, idx, 0); }
template<int T, int W, int H> void t_hps(const pixel* s, intptr_t ss, int16_t* d, intptr_t ds, int idx, int ext) { x265b200_interp(g_b200, X265B200_IP_HPS, T, W, H, s, ss, d, ds, idx, ext); }
template<int T, int W, int H> void t_vpp(const pixel* s, intptr_t ss, pixel* d, intptr_t ds, int idx) { x265b200_interp(g_b200, X265B200_IP_VPP, T, W, H, s, ss, d, ds, idx, 0); }
template<int T, int W, int H> void t_vps(const pixel* s, intptr_t ss, int16_t* d, intptr_t ds, int idx) { x265b200_interp(g_b200, X265B200_IP_VPS, T, W, H, s, ss, d, ds, idx, 0); }
template<int T, int W, int H> void t_vsp(const int16_t* s, intptr_t ss, pixel* d, intptr_t ds, int idx) { x265b200_interp(g_b200, X265B200_IP_VSP, T, W, H, s, ss, d, ds, idx, 0); }
template<int T, int W, int H> void t_vss(const int16_t* s, intptr_t ss, int16_t* d, intptr_t ds, int idx) { x265b200_interp(g_b200, X265B200_IP_VSS, T, W, H, s, ss, d, ds, idx, 0); }
template<int W, int H> void t_hvpp(const pixel* s, intptr_t ss, pixel* d, intptr_t ds, int ix, int iy) { x265b200_interp(g_b200, X265B200_IP_HVPP, 8, W, H, s, ss, d, ds, ix, iy); }
template<int W, int H> void t_p2s(const pixel* s, intptr_t ss, int16_t* d, intptr_t ds) { x265b200_interp(g_b200, X265B200_IP_P2S, 8, W, H, s, ss, d, ds, 0, 0); }

// ---- adjacent slots: same chains, not named by the north star
template<int W, int H> void t_sub_ps(int16_t* d, intptr_t ds, const pixel* a, const pixel* b, intptr_t sa, intptr_t sb) { x265b200_sub_ps(g_b200, W, H, d, ds, a, b, sa, sb); }
template<int W, int H> void t_add_ps(pixel* d, intptr_t ds, const pixel* a, const int16_t* b, intptr_t sa, intptr_t sb) { x265b200_add_ps(g_b200, W, H, d, ds, a, b, sa, sb); }
template<int W, int H> void t_pixelavg(pixel* d, intptr_t ds, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb, int wt) { x265b200_pixelavg_pp(g_b200, W, H, d, ds, a, sa, b, sb, wt); }
template<int W, int H> void t_addAvg(const int16_t* a, const int16_t* b, pixel* d, intptr_t sa, intptr_t sb, intptr_t ds) { x265b200_addAvg(g_b200, W, H, a, b, d, sa, sb, ds); }
template<int W, int H> void t_copy_pp(pixel* d, intptr_t ds, const pixel* a, intptr_t sa) { x265b200_blockcopy(g_b200, 0, W, H, d, ds, a, sa, 0); }
template<int W, int H> void t_copy_ss(int16_t* d, intptr_t ds, const int16_t* a, intptr_t sa) { x265b200_blockcopy(g_b200, 1, W, H, d, ds, a, sa, 0); }
template<int W, int H> void t_copy_sp(pixel* d, intptr_t ds, const int16_t* a, intptr_t sa) { x265b200_blockcopy(g_b200, 2, W, H, d, ds, a, sa, 0); }
template<int W, int H> void t_copy_ps(int16_t* d, intptr_t ds, const pixel* a, intptr_t sa) { x265b200_blockcopy(g_b200, 3, W, H, d, ds, a, sa, 0); }
template<int W> void t_blockfill(int16_t* d, intptr_t ds, int16_t val) { x265b200_blockcopy(g_b200, 4, W, W, d, ds, NULL, 0, val); }
template<int W> void t_2Dto1D_shl(int16_t* d, const int16_t* a, intptr_t sa, int sh) { x265b200_blockcopy(g_b200, 5, W, W, d, W, a, sa, sh); }
template<int W> void t_2Dto1D_shr(int16_t* d, const int16_t* a, intptr_t sa, int sh) { x265b200_blockcopy(g_b200, 6, W, W, d, W, a, sa, sh); }
template<int W> void t_1Dto2D_shl(int16_t* d, const int16_t* a, intptr_t ds, int sh) { x265b200_blockcopy(g_b200, 5, W, W, d, ds, a, W, sh); }
template<int W> void t_1Dto2D_shr(int16_t* d, const int16_t* a, intptr_t ds, int sh) { x265b200_blockcopy(g_b200, 6, W, W, d, ds, a, W, sh); }
template<int W> void t_calcresidual(const pixel* f, const pixel* p, int16_t* r, intptr_t st) { x265b200_sub_ps(g_b200, W, W, r, st, f, p, st, st); }
template<int W> uint64_t t_var(const pixel* p, intptr_t st) { return x265b200_var(g_b200, W, p, st); }
template<int W> int t_psy_cost(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb) { return x265b200_psy_cost_pp(g_b200, W, a, sa, b, sb); }
template<int W> int t_count_nonzero(const int16_t* q) { return (int)x265b200_copy_cnt(g_b200, W, NULL, q, W); }
template<int W> uint32_t t_copy_cnt(int16_t* c, const int16_t* r, intptr_t st) { return x265b200_copy_cnt(g_b200, W, c, r, st); }
void t_denoise(int16_t* d, uint32_t* rs, const uint16_t* off, int n) { x265b200_denoise_dct(g_b200, d, rs, off, n); }
template<int W> void t_integral_h(uint32_t* sum, pixel* pix, intptr_t stride) { x265b200_integral_inith(g_b200, W, sum, pix, stride); }
template<int H> void t_integral_v(uint32_t* sum, intptr_t stride) { x265b200_integral_initv(g_b200, H, sum, stride); }
void t_weight_pp(const pixel* s, pixel* d, intptr_t st, int w, int h, int w0, int rnd, int sh, int off) { x265b200_weight_pp(g_b200, s, d, st, w, h, w0, rnd, sh, off); }
void t_weight_sp(const int16_t* s, pixel* d, intptr_t ss, intptr_t ds, int w, int h, int w0, int rnd, int sh, int off) { x265b200_weight_sp(g_b200, s, d, ss, ds, w, h, w0, rnd, sh, off); }
void t_lowres(const pixel* s, pixel* d0, pixel* dh, pixel* dv, pixel* dc, intptr_t ss, intptr_t ds, int w, int h) { x265b200_frame_init_lowres(g_b200, s, d0, dh, dv, dc, ss, ds, w, h); }

template<int W, int H> void lumaPU(EncoderPrimitives::PU& pu)
{
    pu.pixelavg_pp[NONALIGNED] = t_pixelavg<W, H>; pu.pixelavg_pp[ALIGNED] = t_pixelavg<W, H>;
    pu.copy_pp = t_copy_pp<W, H>;
    pu.addAvg[NONALIGNED] = t_addAvg<W, H>; pu.addAvg[ALIGNED] = t_addAvg<W, H>;
    pu.sad = t_sad<W, H>; pu.sad_x3 = t_sad_x3<W, H>; pu.sad_x4 = t_sad_x4<W, H>; pu.ads = t_ads<W, H>; pu.satd = t_satd<W, H>;
    pu.luma_hpp = t_hpp<8, W, H>; pu.luma_hps = t_hps<8, W, H>; pu.luma_vpp = t_vpp<8, W, H>; pu.luma_vps = t_vps<8, W, H>;
    pu.luma_vsp = t_vsp<8, W, H>; pu.luma_vss = t_vss<8, W, H>; pu.luma_hvpp = t_hvpp<W, H>;
    pu.convert_p2s[NONALIGNED] = t_p2s<W, H>; pu.convert_p2s[ALIGNED] = t_p2s<W, H>;
}

// chroma PU of size W x H (already subsampled); satd only where the reference has one (dims % 4 == 0)
template<int W, int H> void chromaPU(EncoderPrimitives::Chroma::PUChroma& pu)
{
    pu.filter_hpp = t_hpp<4, W, H>; pu.filter_hps = t_hps<4, W, H>; pu.filter_vpp = t_vpp<4, W, H>; pu.filter_vps = t_vps<4, W, H>;
    pu.filter_vsp = t_vsp<4, W, H>; pu.filter_vss = t_vss<4, W, H>;
    pu.p2s[NONALIGNED] = t_p2s<W, H>; pu.p2s[ALIGNED] = t_p2s<W, H>;
    pu.addAvg[NONALIGNED] = t_addAvg<W, H>; pu.addAvg[ALIGNED] = t_addAvg<W, H>;
    pu.copy_pp = t_copy_pp<W, H>;
    pu.satd = (W % 4 == 0 && H % 4 == 0) ? (pixelcmp_t)t_satd<(W % 4 ? 4 : W), (H % 4 ? 4 : H)> : NULL;
}

template<int W> void lumaCU(EncoderPrimitives::CU& cu)
{
    cu.sse_pp = t_sse_pp<W, W>; cu.sse_ss = t_sse_ss<W, W>; cu.sa8d = t_sa8d<W, W>;
    cu.ssd_s[NONALIGNED] = t_ssd_s<W>; cu.ssd_s[ALIGNED] = t_ssd_s<W>;
    cu.sub_ps = t_sub_ps<W, W>; cu.add_ps[NONALIGNED] = t_add_ps<W, W>; cu.add_ps[ALIGNED] = t_add_ps<W, W>;
    cu.copy_ss = t_copy_ss<W, W>; cu.copy_sp = t_copy_sp<W, W>; cu.copy_ps = t_copy_ps<W, W>;
    cu.blockfill_s[NONALIGNED] = t_blockfill<W>; cu.blockfill_s[ALIGNED] = t_blockfill<W>;
    cu.calcresidual[NONALIGNED] = t_calcresidual<W>; cu.calcresidual[ALIGNED] = t_calcresidual<W>;
    cu.var = t_var<W>; cu.psy_cost_pp = t_psy_cost<W>;
}
// the 1-D <-> 2-D coefficient copies (pixel.cpp:1081-1085, all five CU sizes)
template<int W> void lumaTU(EncoderPrimitives::CU& cu)
{
    cu.cpy2Dto1D_shl = t_2Dto1D_shl<W>; cu.cpy2Dto1D_shr = t_2Dto1D_shr<W>;
    cu.cpy1Dto2D_shl[NONALIGNED] = t_1Dto2D_shl<W>; cu.cpy1Dto2D_shl[ALIGNED] = t_1Dto2D_shl<W>; cu.cpy1Dto2D_shr = t_1Dto2D_shr<W>;
}
template<int W> void lumaCoef(EncoderPrimitives::CU& cu) { cu.count_nonzero = t_count_nonzero<W>; cu.copy_cnt = t_copy_cnt<W>; }   // dct.cpp:1100-1108
template<int W, int H> void chromaCU(EncoderPrimitives::Chroma::CUChroma& cu)
{
    cu.sa8d = t_sa8d<W, H>; cu.sse_pp = t_sse_pp<W, H>;
    cu.sub_ps = t_sub_ps<W, H>; cu.add_ps[NONALIGNED] = t_add_ps<W, H>; cu.add_ps[ALIGNED] = t_add_ps<W, H>;
    cu.copy_ss = t_copy_ss<W, H>; cu.copy_sp = t_copy_sp<W, H>; cu.copy_ps = t_copy_ps<W, H>;
}

} // anonymous namespace

#define FOR_ALL_LUMA_PU(X) \
    X(4, 4) X(8, 8) X(16, 16) X(32, 32) X(64, 64) X(8, 4) X(4, 8) X(16, 8) X(8, 16) X(32, 16) X(16, 32) X(64, 32) X(32, 64) \
    X(16, 12) X(12, 16) X(16, 4) X(4, 16) X(32, 24) X(24, 32) X(32, 8) X(8, 32) X(64, 48) X(48, 64) X(64, 16) X(16, 64)

void setupB200Primitives(EncoderPrimitives& p)
{
#define L(W, H) lumaPU<W, H>(p.pu[LUMA_ ## W ## x ## H]);
    FOR_ALL_LUMA_PU(L)
#undef L
    // 4:4:4 chroma filters use luma block sizes with the 4-tap kernel (ipfilter.cpp:391-399); the rest
    // of the 4:4:4 table is aliased from luma by setupAliasPrimitives (primitives.cpp:185-206)
#define C444(W, H) { pixelcmp_t keep = p.chroma[X265_CSP_I444].pu[LUMA_ ## W ## x ## H].satd; \
                     chromaPU<W, H>(p.chroma[X265_CSP_I444].pu[LUMA_ ## W ## x ## H]); (void)keep; }
    FOR_ALL_LUMA_PU(C444)
#undef C444
    // 4:2:0: chroma PU = (W/2) x (H/2); the reference has no entry for luma 4x4 (ipfilter.cpp:414-462)
#define C420(W, H) if (W > 4 || H > 4) chromaPU<(W) / 2, (H) / 2>(p.chroma[X265_CSP_I420].pu[LUMA_ ## W ## x ## H]);
    FOR_ALL_LUMA_PU(C420)
#undef C420
    // 4:2:2: chroma PU = (W/2) x H
#define C422(W, H) chromaPU<(W) / 2, H>(p.chroma[X265_CSP_I422].pu[LUMA_ ## W ## x ## H]);
    FOR_ALL_LUMA_PU(C422)
#undef C422

    lumaCU<4>(p.cu[BLOCK_4x4]); lumaCU<8>(p.cu[BLOCK_8x8]); lumaCU<16>(p.cu[BLOCK_16x16]);
    lumaCU<32>(p.cu[BLOCK_32x32]); lumaCU<64>(p.cu[BLOCK_64x64]);
    lumaTU<4>(p.cu[BLOCK_4x4]); lumaTU<8>(p.cu[BLOCK_8x8]); lumaTU<16>(p.cu[BLOCK_16x16]); lumaTU<32>(p.cu[BLOCK_32x32]); lumaTU<64>(p.cu[BLOCK_64x64]);

    // chroma CU slots the alias pass does not derive from luma (pixel.cpp:1260-1263,1325,1339-1342)
    chromaCU<4, 4>(p.chroma[X265_CSP_I420].cu[BLOCK_8x8]);   chromaCU<8, 8>(p.chroma[X265_CSP_I420].cu[BLOCK_16x16]);
    chromaCU<16, 16>(p.chroma[X265_CSP_I420].cu[BLOCK_32x32]); chromaCU<32, 32>(p.chroma[X265_CSP_I420].cu[BLOCK_64x64]);
    chromaCU<4, 8>(p.chroma[X265_CSP_I422].cu[BLOCK_8x8]);   chromaCU<8, 16>(p.chroma[X265_CSP_I422].cu[BLOCK_16x16]);
    chromaCU<16, 32>(p.chroma[X265_CSP_I422].cu[BLOCK_32x32]); chromaCU<32, 64>(p.chroma[X265_CSP_I422].cu[BLOCK_64x64]);

    // 4:2:0 chroma of the 4x4 luma PU: no filters (ipfilter.cpp:414-462) but a bi-prediction average (pixel.cpp:1191)
    p.chroma[X265_CSP_I420].pu[LUMA_4x4].addAvg[NONALIGNED] = t_addAvg<2, 2>; p.chroma[X265_CSP_I420].pu[LUMA_4x4].addAvg[ALIGNED] = t_addAvg<2, 2>;
    // chroma of the 4x4 luma CU: only the residual slots exist (pixel.cpp:1254, 1333)
    p.chroma[X265_CSP_I420].pu[LUMA_4x4].copy_pp = t_copy_pp<2, 2>;
    p.chroma[X265_CSP_I420].cu[BLOCK_4x4].copy_ss = t_copy_ss<2, 2>; p.chroma[X265_CSP_I420].cu[BLOCK_4x4].copy_sp = t_copy_sp<2, 2>;
    p.chroma[X265_CSP_I420].cu[BLOCK_4x4].copy_ps = t_copy_ps<2, 2>;
    p.chroma[X265_CSP_I422].cu[BLOCK_4x4].copy_ss = t_copy_ss<2, 4>; p.chroma[X265_CSP_I422].cu[BLOCK_4x4].copy_sp = t_copy_sp<2, 4>;
    p.chroma[X265_CSP_I422].cu[BLOCK_4x4].copy_ps = t_copy_ps<2, 4>;
    p.chroma[X265_CSP_I420].cu[BLOCK_4x4].sub_ps = t_sub_ps<2, 2>;
    p.chroma[X265_CSP_I420].cu[BLOCK_4x4].add_ps[NONALIGNED] = t_add_ps<2, 2>; p.chroma[X265_CSP_I420].cu[BLOCK_4x4].add_ps[ALIGNED] = t_add_ps<2, 2>;
    p.chroma[X265_CSP_I422].cu[BLOCK_4x4].sub_ps = t_sub_ps<2, 4>;
    p.chroma[X265_CSP_I422].cu[BLOCK_4x4].add_ps[NONALIGNED] = t_add_ps<2, 4>; p.chroma[X265_CSP_I422].cu[BLOCK_4x4].add_ps[ALIGNED] = t_add_ps<2, 4>;

    p.cu[BLOCK_4x4].dct = t_dct<4>;     p.cu[BLOCK_4x4].idct = t_idct<4>;
    p.cu[BLOCK_8x8].dct = t_dct<8>;     p.cu[BLOCK_8x8].idct = t_idct<8>;
    p.cu[BLOCK_16x16].dct = t_dct<16>;  p.cu[BLOCK_16x16].idct = t_idct<16>;
    p.cu[BLOCK_32x32].dct = t_dct<32>;  p.cu[BLOCK_32x32].idct = t_idct<32>;
    p.cu[BLOCK_8x8].lowpass_dct = t_lowpass<8>;
    p.cu[BLOCK_16x16].lowpass_dct = t_lowpass<16>;
    p.cu[BLOCK_32x32].lowpass_dct = t_lowpass<32>;
    p.dst4x4 = t_dst4;  p.idst4x4 = t_idst4;
    p.quant = t_quant;  p.nquant = t_nquant;
    p.dequant_normal = t_dequant_normal;  p.dequant_scaling = t_dequant_scaling;
    lumaCoef<4>(p.cu[BLOCK_4x4]); lumaCoef<8>(p.cu[BLOCK_8x8]); lumaCoef<16>(p.cu[BLOCK_16x16]); lumaCoef<32>(p.cu[BLOCK_32x32]);
    p.denoiseDct = t_denoise;
    p.frameInitLowres = t_lowres;
#define B200_INTRA(IDX, N) \
    p.cu[IDX].intra_filter = t_intra_filter<N>; p.cu[IDX].intra_pred_allangs = t_intra_allangs<N>; \
    p.cu[IDX].intra_pred[PLANAR_IDX] = t_intra_planar<N>; p.cu[IDX].intra_pred[DC_IDX] = t_intra_dc<N>; \
    for (int m = 2; m < NUM_INTRA_MODE; m++) p.cu[IDX].intra_pred[m] = t_intra_pred<N>;
    B200_INTRA(BLOCK_4x4, 4) B200_INTRA(BLOCK_8x8, 8) B200_INTRA(BLOCK_16x16, 16) B200_INTRA(BLOCK_32x32, 32)
#undef B200_INTRA
    p.weight_pp = t_weight_pp;  p.weight_sp = t_weight_sp;
    p.integral_inith[INTEGRAL_4] = t_integral_h<4>;   p.integral_initv[INTEGRAL_4] = t_integral_v<4>;
    p.integral_inith[INTEGRAL_8] = t_integral_h<8>;   p.integral_initv[INTEGRAL_8] = t_integral_v<8>;
    p.integral_inith[INTEGRAL_12] = t_integral_h<12>; p.integral_initv[INTEGRAL_12] = t_integral_v<12>;
    p.integral_inith[INTEGRAL_16] = t_integral_h<16>; p.integral_initv[INTEGRAL_16] = t_integral_v<16>;
    p.integral_inith[INTEGRAL_24] = t_integral_h<24>; p.integral_initv[INTEGRAL_24] = t_integral_v<24>;
    p.integral_inith[INTEGRAL_32] = t_integral_h<32>; p.integral_initv[INTEGRAL_32] = t_integral_v<32>;
}

} // namespace X265_NS

// C handle for drivers that cannot name the C++ symbol: opens the context (once) and fills `table`,
// which must point at an EncoderPrimitives of this build's bit depth.  Returns 0 or an X265B200_ERR_*.
extern "C" int x265b200_setup_primitives(void* table, int device)
{
    static std::mutex openLock;         // two encoder instances may initialise at the same time
    {
        std::lock_guard<std::mutex> g(openLock);
        if (!X265_NS::g_b200)
        {
            int r = x265b200_open(device, X265_DEPTH, &X265_NS::g_b200);
            if (r != X265B200_OK) return r;
        }
    }
    X265_NS::setupB200Primitives(*(X265_NS::EncoderPrimitives*)table);
    return X265B200_OK;
}
/* A slot cannot report an error (SURVEY.md 8b "Errors"): the encoder polls this at a frame boundary
 * (INTEGRATION.md shows where) and aborts loudly on a non-zero status; there is no CPU fallback to fall back to. */
extern "C" int x265b200_glue_status(const char** message)
{
    if (!X265_NS::g_b200) { if (message) *message = "x265b200_setup_primitives was not called"; return X265B200_ERR_ARG; }
    if (message) *message = x265b200_last_error(X265_NS::g_b200);
    return x265b200_status(X265_NS::g_b200);
}
extern "C" x265b200_ctx* x265b200_glue_context(void) { return X265_NS::g_b200; }
extern "C" int x265b200_glue_depth(void) { return X265_DEPTH; }
