// ipfilter.cu -- HEVC sub-pel interpolation: 8-tap luma / 4-tap chroma FIR in all seven x265
// variants plus pixel-to-short.  Bit-exact CUDA restatement of the reference's ipfilter.cpp:40-369.
//
// The variants differ only in input/output type, rounding offset, shift and clipping
// (SURVEY.md appendix C), so one templated FIR kernel covers hpp/hps/vpp/vps/vsp/vss; hvpp runs
// the horizontal pixel->short pass into a shared-memory tile and the vertical short->pixel pass
// out of it inside one CTA (the reference does the same through a stack buffer, ipfilter.cpp:362-369).
#include "internal.h"
#include "device_util.cuh"
#include "tile_kernels.cuh"

namespace b200 {

__constant__ short c_lumaTaps[4][8];
__constant__ short c_chromaTaps[8][4];
// the same taps as signed bytes for dp2a: [idx][0] = taps 0..3, [idx][1] = taps 4..7 (luma only)
__constant__ uint32_t c_lumaTapsB[4][2];
__constant__ uint32_t c_chromaTapsB[8];

int upload_filter_tables(x265b200_ctx* ctx)
{
    // HEVC fractional-sample filters (== g_lumaFilter / g_chromaFilter, constants.cpp:250-268)
    static const short luma[4][8] = { { 0, 0, 0, 64, 0, 0, 0, 0 }, { -1, 4, -10, 58, 17, -5, 1, 0 },
                                      { -1, 4, -11, 40, 40, -11, 4, -1 }, { 0, 1, -5, 17, 58, -10, 4, -1 } };
    static const short chroma[8][4] = { { 0, 64, 0, 0 }, { -2, 58, 10, -2 }, { -4, 54, 16, -2 }, { -6, 46, 28, -4 },
                                        { -4, 36, 36, -4 }, { -4, 28, 46, -6 }, { -2, 16, 54, -4 }, { -2, 10, 58, -2 } };
    uint32_t lb[4][2], cb[8];
    for (int i = 0; i < 4; i++)
        for (int hh = 0; hh < 2; hh++)
        {
            lb[i][hh] = 0;
            for (int e = 0; e < 4; e++) lb[i][hh] |= (uint32_t)(uint8_t)(int8_t)luma[i][4 * hh + e] << (8 * e);
        }
    for (int i = 0; i < 8; i++)
    {
        cb[i] = 0;
        for (int e = 0; e < 4; e++) cb[i] |= (uint32_t)(uint8_t)(int8_t)chroma[i][e] << (8 * e);
    }
    B200_CUDA(ctx, cudaMemcpyToSymbol(c_lumaTapsB, lb, sizeof(lb)));
    B200_CUDA(ctx, cudaMemcpyToSymbol(c_chromaTapsB, cb, sizeof(cb)));
    B200_CUDA(ctx, cudaMemcpyToSymbol(c_lumaTaps, luma, sizeof(luma)));
    B200_CUDA(ctx, cudaMemcpyToSymbol(c_chromaTaps, chroma, sizeof(chroma)));
    return X265B200_OK;
}

template<int TAPS> __device__ __forceinline__ int tap(int idx, int t)
{
    return TAPS == 8 ? c_lumaTaps[idx & 3][t] : c_chromaTaps[idx & 7][t];
}

struct FirParams
{
    int w, h;            // block size
    int shift, offset;   // (sum + offset) >> shift
    int maxVal;          // >= 0: cast to int16 then clip to [0, maxVal] (pixel output); < 0: store int16 (wraps)
    int rowExtKind;      // 1 for HPS: per-block isRowExt extends the block by TAPS-1 rows
};

// one thread per output sample
template<typename SRC, typename DST, int TAPS, bool VERT>
__global__ void __launch_bounds__(256)
fir_kernel(const SRC* __restrict__ src, intptr_t ss, const int32_t* __restrict__ offSrc,
           DST* __restrict__ dst, intptr_t ds, const int32_t* __restrict__ offDst,
           const int32_t* __restrict__ coeffIdx, int n, FirParams p)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int rowsMax = p.rowExtKind ? p.h + TAPS - 1 : p.h;
    int per = p.w * rowsMax;
    int blk = (int)(gid / per);
    if (blk >= n) return;
    int r = (int)(gid % per);
    int y = r / p.w, x = r % p.w;
    int ci = coeffIdx[blk];
    int idx = ci & 15;
    const SRC* s = src + offSrc[blk];
    int rows = p.h;
    if (p.rowExtKind && (ci >> 8 & 1)) { s -= (TAPS / 2 - 1) * ss; rows += TAPS - 1; }   // ipfilter.cpp:130-134
    if (y >= rows) return;
    intptr_t step = VERT ? ss : 1;
    s += (intptr_t)y * ss + x - (TAPS / 2 - 1) * step;
    int sum = 0;
#pragma unroll
    for (int t = 0; t < TAPS; t++) sum += (int)s[t * step] * tap<TAPS>(idx, t);
    int v = (int)(int16_t)((sum + p.offset) >> p.shift);          // cast to int16 BEFORE clipping (ipfilter.cpp:108-112)
    if (p.maxVal >= 0) v = min(max(v, 0), p.maxVal);
    dst[offDst[blk] + (intptr_t)y * ds + x] = (DST)v;
}

// Throughput version for blocks whose width is a multiple of 4 and strides are multiples of 4 samples:
// one thread produces a 4x4 output tile.  Horizontal filters read 4 rows x (TAPS+3) samples, vertical
// filters (TAPS+3) rows x 4 samples, through aligned chunk loads (device_util.cuh load_row_quads); every input
// sample is loaded once per tile instead of once per tap.
template<typename DST>
__device__ __forceinline__ void store4(DST* d, const int (&v)[4])
{
    if (sizeof(DST) == 2)
    {
        uint32_t lo = (uint32_t)(v[0] & 0xffff) | ((uint32_t)v[1] << 16), hi = (uint32_t)(v[2] & 0xffff) | ((uint32_t)v[3] << 16);
        uintptr_t a = (uintptr_t)d;
        if ((a & 7) == 0) *(uint2*)d = make_uint2(lo, hi);
        else if ((a & 3) == 0) { ((uint32_t*)d)[0] = lo; ((uint32_t*)d)[1] = hi; }
        else { d[0] = (DST)v[0]; d[1] = (DST)v[1]; d[2] = (DST)v[2]; d[3] = (DST)v[3]; }
    }
    else
    {
        uintptr_t a = (uintptr_t)d;
        if ((a & 3) == 0) *(uint32_t*)d = (uint32_t)(v[0] & 0xff) | ((uint32_t)(v[1] & 0xff) << 8) | ((uint32_t)(v[2] & 0xff) << 16) | ((uint32_t)v[3] << 24);
        else { d[0] = (DST)v[0]; d[1] = (DST)v[1]; d[2] = (DST)v[2]; d[3] = (DST)v[3]; }
    }
}

// Filter epilogue on sample pairs: the byte permute that packs two sums keeps their low 16 bits, which IS the
// reference's cast to int16_t before clipping (ipfilter.cpp:108-112); the clip to [0, maxVal] is one VIMNMX.S16x2.RELU
// for both samples.  mx = maxVal | maxVal << 16, or 0 for int16 outputs (stored as they are: they wrap, never saturate).
__device__ __forceinline__ uint32_t pack_clip2(int q0, int q1, uint32_t mx)
{
    uint32_t p = __byte_perm((uint32_t)q0, (uint32_t)q1, 0x5410);
    return mx ? __vimin_s16x2_relu(p, mx) : p;
}
template<typename DST>
__device__ __forceinline__ void store4p(DST* d, uint32_t p01, uint32_t p23)
{
    uintptr_t a = (uintptr_t)d;
    if (sizeof(DST) == 2)
    {
        if ((a & 7) == 0) *(uint2*)d = make_uint2(p01, p23);
        else if ((a & 3) == 0) { ((uint32_t*)d)[0] = p01; ((uint32_t*)d)[1] = p23; }
        else { d[0] = (DST)(p01 & 0xffff); d[1] = (DST)(p01 >> 16); d[2] = (DST)(p23 & 0xffff); d[3] = (DST)(p23 >> 16); }
    }
    else
    {
        uint32_t b = __byte_perm(p01, p23, 0x6420);
        if ((a & 3) == 0) *(uint32_t*)d = b;
        else { d[0] = (DST)(b & 0xff); d[1] = (DST)((b >> 8) & 0xff); d[2] = (DST)((b >> 16) & 0xff); d[3] = (DST)(b >> 24); }
    }
}

// two-way dot product of packed 16-bit samples with two signed-byte taps (IDP.2A): the sample pair is unsigned
// for pixels and signed for the int16 intermediates
template<typename SRC> __device__ __forceinline__ int dp2a_lo(uint32_t a, uint32_t b, int c)
{
    int d;
    if (sizeof(SRC) == 2 && SRC(-1) < SRC(0)) asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
template<typename SRC> __device__ __forceinline__ int dp2a_hi(uint32_t a, uint32_t b, int c)
{
    int d;
    if (sizeof(SRC) == 2 && SRC(-1) < SRC(0)) asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// sum over TAPS taps of four packed pair-words (luma) / two (chroma), starting value `acc`
template<typename SRC, int TAPS>
__device__ __forceinline__ int fir_pairs(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t t0, uint32_t t1, int acc)
{
    acc = dp2a_lo<SRC>(p0, t0, acc);
    acc = dp2a_hi<SRC>(p1, t0, acc);
    if (TAPS == 8)
    {
        acc = dp2a_lo<SRC>(p2, t1, acc);
        acc = dp2a_hi<SRC>(p3, t1, acc);
    }
    return acc;
}

// TW x TH output tile per thread: horizontal filters use TW in {4, 8} (TH = 4), vertical TH in {4, 8} (TW = 4).
// Samples stay packed two per register; taps are signed bytes, so every pair of taps is one IDP.2A
// (pixels are widened to 16-bit lanes by the loader for 8-bit builds).  FULL = every row of the tile exists
// (no per-row predicates); partial tiles (last rows of an isRowExt block) take the predicated variant.
template<typename SRC, typename DST, int TAPS, bool VERT, int TW, int TH, bool FULL>
__device__ __forceinline__ void fir_tile_body(const SRC* __restrict__ s, intptr_t ss, DST* __restrict__ d, intptr_t ds,
                                              int nr, uint32_t t0, uint32_t t1, const FirParams& p)
{
    const uint32_t mx = p.maxVal >= 0 ? (uint32_t)p.maxVal * 0x10001u : 0u;
    if (!VERT)
    {
        constexpr int NQ = (TW + TAPS - 1 + 3) / 4;              // quads covering TW + TAPS - 1 samples
        uint32_t w[TH][2 * NQ + 1];
        if (FULL)
            load_rows_quads<NQ, TH>(s, ss, w);
        else
        {
#pragma unroll
            for (int r = 0; r < TH; r++)
                if (r < nr) { load_row_quads<NQ>(s + r * ss, (uint32_t(&)[2 * NQ])w[r]); w[r][2 * NQ] = 0; }
        }
#pragma unroll
        for (int r = 0; r < TH; r++)
        {
            if (!FULL && r >= nr) break;
            // odd-phase words: (x[2i+1], x[2i+2])
            uint32_t ws[2 * NQ];
#pragma unroll
            for (int i = 0; i < 2 * NQ; i++) ws[i] = __funnelshift_r(w[r][i], w[r][i + 1], 16);
#pragma unroll
            for (int o4 = 0; o4 < TW; o4 += 4)
            {
                int v[4];
#pragma unroll
                for (int o = 0; o < 4; o++)
                {
                    int i = (o4 + o) >> 1;
                    int sum = ((o4 + o) & 1) ? fir_pairs<SRC, TAPS>(ws[i], ws[i + 1], ws[(i + 2) % (2 * NQ)], ws[(i + 3) % (2 * NQ)], t0, t1, p.offset)
                                             : fir_pairs<SRC, TAPS>(w[r][i], w[r][i + 1], w[r][(i + 2) % (2 * NQ + 1)], w[r][(i + 3) % (2 * NQ + 1)], t0, t1, p.offset);
                    v[o] = sum >> p.shift;
                }
                store4p(d + r * ds + o4, pack_clip2(v[0], v[1], mx), pack_clip2(v[2], v[3], mx));
            }
        }
    }
    else
    {
        constexpr int NR = TH + TAPS - 1;
        uint32_t w[NR][3];
        if (FULL)
            load_rows_quads<1, NR>(s, ss, w);
        else
        {
#pragma unroll
            for (int r = 0; r < NR; r++)
                if (r < nr + TAPS - 1) load_row_quads<1>(s + r * ss, (uint32_t(&)[2])w[r]);
        }
        // pair words of vertically adjacent rows: pr[r][c] = (x[r][c], x[r+1][c])
        uint32_t pr[NR - 1][4];
#pragma unroll
        for (int r = 0; r < NR - 1; r++)
        {
            pr[r][0] = __byte_perm(w[r][0], w[r + 1][0], 0x5410); pr[r][1] = __byte_perm(w[r][0], w[r + 1][0], 0x7632);
            pr[r][2] = __byte_perm(w[r][1], w[r + 1][1], 0x5410); pr[r][3] = __byte_perm(w[r][1], w[r + 1][1], 0x7632);
        }
#pragma unroll
        for (int r = 0; r < TH; r++)
        {
            if (!FULL && r >= nr) break;
            int v[4];
#pragma unroll
            for (int o = 0; o < 4; o++)
                v[o] = fir_pairs<SRC, TAPS>(pr[r][o], pr[r + 2][o], pr[(r + 4) % (NR - 1)][o], pr[(r + 6) % (NR - 1)][o], t0, t1, p.offset) >> p.shift;
            store4p(d + r * ds, pack_clip2(v[0], v[1], mx), pack_clip2(v[2], v[3], mx));
        }
    }
}

template<typename SRC, typename DST, int TAPS, bool VERT, int TW, int TH>
__global__ void __launch_bounds__(128)
fir_tile_kernel(const SRC* __restrict__ src, intptr_t ss, const int32_t* __restrict__ offSrc,
                DST* __restrict__ dst, intptr_t ds, const int32_t* __restrict__ offDst,
                const int32_t* __restrict__ coeffIdx, int n, FirParams p)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int rowsMax = p.rowExtKind ? p.h + TAPS - 1 : p.h;
    int tw = p.w / TW;
    int per = tw * ((rowsMax + TH - 1) / TH);
    int blk = (int)(gid / per);
    if (blk >= n) return;
    int t = (int)(gid % per);
    int tx = (t % tw) * TW, ty = (t / tw) * TH;
    int ci = coeffIdx[blk];
    int idx = ci & 15;
    const SRC* s = src + offSrc[blk];
    int rows = p.h;
    if (p.rowExtKind && (ci >> 8 & 1)) { s -= (TAPS / 2 - 1) * ss; rows += TAPS - 1; }
    if (ty >= rows) return;
    DST* d = dst + offDst[blk] + (intptr_t)ty * ds + tx;
    uint32_t t0 = TAPS == 8 ? c_lumaTapsB[idx & 3][0] : c_chromaTapsB[idx & 7];
    uint32_t t1 = TAPS == 8 ? c_lumaTapsB[idx & 3][1] : 0;
    s += VERT ? (intptr_t)(ty - (TAPS / 2 - 1)) * ss + tx : (intptr_t)ty * ss + tx - (TAPS / 2 - 1);
    int nr = min(TH, rows - ty);
    if (nr == TH) fir_tile_body<SRC, DST, TAPS, VERT, TW, TH, true>(s, ss, d, ds, nr, t0, t1, p);
    else fir_tile_body<SRC, DST, TAPS, VERT, TW, TH, false>(s, ss, d, ds, nr, t0, t1, p);
}

// hvpp: one CTA per block.  Pass 1 = hps(isRowExt=1) into smem (pitch w), pass 2 = vertical sp.
template<typename PIX, int TAPS>
__global__ void __launch_bounds__(256)
hv_kernel(const PIX* __restrict__ src, intptr_t ss, const int32_t* __restrict__ offSrc,
          PIX* __restrict__ dst, intptr_t ds, const int32_t* __restrict__ offDst,
          const int32_t* __restrict__ coeffIdx, int w, int h, int shift1, int offset1, int shift2, int offset2, int maxVal)
{
    extern __shared__ int16_t immed[];
    int blk = blockIdx.x;
    int ci = coeffIdx[blk];
    int idxX = ci & 15, idxY = (ci >> 4) & 15;
    const PIX* s = src + offSrc[blk] - (TAPS / 2 - 1) * ss - (TAPS / 2 - 1);
    int rows = h + TAPS - 1;
    for (int i = threadIdx.x; i < w * rows; i += blockDim.x)
    {
        int y = i / w, x = i % w;
        const PIX* q = s + (intptr_t)y * ss + x;
        int sum = 0;
#pragma unroll
        for (int t = 0; t < TAPS; t++) sum += (int)q[t] * tap<TAPS>(idxX, t);
        immed[i] = (int16_t)((sum + offset1) >> shift1);
    }
    __syncthreads();
    PIX* d = dst + offDst[blk];
    for (int i = threadIdx.x; i < w * h; i += blockDim.x)
    {
        int y = i / w, x = i % w;
        int sum = 0;
#pragma unroll
        for (int t = 0; t < TAPS; t++) sum += (int)immed[(y + t) * w + x] * tap<TAPS>(idxY, t);
        int v = (int)(int16_t)((sum + offset2) >> shift2);
        v = min(max(v, 0), maxVal);
        d[(intptr_t)y * ds + x] = (PIX)v;
    }
}

// hvpp throughput version (w % 4 == 0, h % 4 == 0, stride % 4 == 0): a group of G lanes (power of two, <= 32)
// owns one block.  Pass 1 = hps with isRowExt into the group's smem tile (pitch w), TW1 x 4 tiles through
// load_row_quads + IDP.2A; pass 2 = vertical sp out of smem (8-byte aligned LDS), 4 x TH2 output tiles.
// TW1 = 8 / TH2 = 8 (when the block allows) halve the loads per output sample of the respective pass.
template<typename PIX, int TAPS, int TW1, int TH2>
__global__ void __launch_bounds__(128)
hv_tile_kernel(const PIX* __restrict__ src, intptr_t ss, const int32_t* __restrict__ offSrc,
               PIX* __restrict__ dst, intptr_t ds, const int32_t* __restrict__ offDst,
               const int32_t* __restrict__ coeffIdx, int n, int w, int h, int G, int shift1, int offset1, int shift2, int offset2, int maxVal)
{
    extern __shared__ __align__(16) int16_t immed_all[];
    int lg = __ffs(G) - 1;
    int grp = threadIdx.x >> lg, l = threadIdx.x & (G - 1);
    int blk = blockIdx.x * (128 >> lg) + grp;
    bool live = blk < n;
    int rows = h + TAPS - 1;
    int16_t* immed = immed_all + (size_t)grp * w * rows;
    constexpr int NQ = (TW1 + TAPS - 1 + 3) / 4;
    uint32_t tx0 = 0, tx1 = 0, ty0 = 0, ty1 = 0;
    if (live)
    {
        int ci = coeffIdx[blk];
        int idxX = ci & 15, idxY = (ci >> 4) & 15;
        tx0 = TAPS == 8 ? c_lumaTapsB[idxX & 3][0] : c_chromaTapsB[idxX & 7];
        tx1 = TAPS == 8 ? c_lumaTapsB[idxX & 3][1] : 0;
        ty0 = TAPS == 8 ? c_lumaTapsB[idxY & 3][0] : c_chromaTapsB[idxY & 7];
        ty1 = TAPS == 8 ? c_lumaTapsB[idxY & 3][1] : 0;
        const PIX* s = src + offSrc[blk] - (TAPS / 2 - 1) * ss - (TAPS / 2 - 1);
        int tw = w / TW1;
        int tiles1 = tw * ((rows + 3) >> 2);
        for (int t = l; t < tiles1; t += G)
        {
            int tx = (t % tw) * TW1, ty = (t / tw) << 2;
            uint32_t wv[4][2 * NQ + 1];
            if (ty + 4 <= rows)
                load_rows_quads<NQ, 4>(s + (intptr_t)ty * ss + tx, ss, wv);
            else
            {
#pragma unroll
                for (int r = 0; r < 4; r++)
                    if (ty + r < rows) { load_row_quads<NQ>(s + (intptr_t)(ty + r) * ss + tx, (uint32_t(&)[2 * NQ])wv[r]); wv[r][2 * NQ] = 0; }
            }
#pragma unroll
            for (int r = 0; r < 4; r++)
            {
                if (ty + r >= rows) break;
                uint32_t ws[2 * NQ];
#pragma unroll
                for (int i = 0; i < 2 * NQ; i++) ws[i] = __funnelshift_r(wv[r][i], wv[r][i + 1], 16);
#pragma unroll
                for (int o4 = 0; o4 < TW1; o4 += 4)
                {
                    int v[4];
#pragma unroll
                    for (int o = 0; o < 4; o++)
                    {
                        int i = (o4 + o) >> 1;
                        int sum = ((o4 + o) & 1) ? fir_pairs<PIX, TAPS>(ws[i], ws[i + 1], ws[(i + 2) % (2 * NQ)], ws[(i + 3) % (2 * NQ)], tx0, tx1, offset1)
                                                 : fir_pairs<PIX, TAPS>(wv[r][i], wv[r][i + 1], wv[r][(i + 2) % (2 * NQ + 1)], wv[r][(i + 3) % (2 * NQ + 1)], tx0, tx1, offset1);
                        v[o] = sum >> shift1;
                    }
                    *(uint2*)(immed + (ty + r) * w + tx + o4) = make_uint2(__byte_perm((uint32_t)v[0], (uint32_t)v[1], 0x5410),
                                                                           __byte_perm((uint32_t)v[2], (uint32_t)v[3], 0x5410));
                }
            }
        }
    }
    __syncwarp();
    if (!live) return;
    PIX* d = dst + offDst[blk];
    int tw = w >> 2;
    int tiles2 = tw * (h / TH2);
    const uint32_t mx = (uint32_t)maxVal * 0x10001u;
    for (int t = l; t < tiles2; t += G)
    {
        int tx = (t % tw) << 2, ty = (t / tw) * TH2;
        constexpr int NR = TH2 + TAPS - 1;
        uint2 q[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) q[r] = *(const uint2*)(immed + (ty + r) * w + tx);
        uint32_t pr[NR - 1][4];
#pragma unroll
        for (int r = 0; r < NR - 1; r++)
        {
            pr[r][0] = __byte_perm(q[r].x, q[r + 1].x, 0x5410); pr[r][1] = __byte_perm(q[r].x, q[r + 1].x, 0x7632);
            pr[r][2] = __byte_perm(q[r].y, q[r + 1].y, 0x5410); pr[r][3] = __byte_perm(q[r].y, q[r + 1].y, 0x7632);
        }
#pragma unroll
        for (int r = 0; r < TH2; r++)
        {
            int v[4];
#pragma unroll
            for (int o = 0; o < 4; o++)
                v[o] = fir_pairs<int16_t, TAPS>(pr[r][o], pr[r + 2][o], pr[(r + 4) % (NR - 1)][o], pr[(r + 6) % (NR - 1)][o], ty0, ty1, offset2) >> shift2;
            store4p(d + (intptr_t)(ty + r) * ds + tx, pack_clip2(v[0], v[1], mx), pack_clip2(v[2], v[3], mx));
        }
    }
}

// Sub-pel candidate cost = interpolation fused with the block metric (reference encoder/motion.cpp:1780-1821,
// MotionEstimate::subpelCompare: luma_hpp / luma_vpp / luma_hvpp into a stack buffer, then sad or satd against the cached
// fenc block).  One lane group per candidate, same two passes as hv_tile_kernel, but the vertical pass leaves its 4 x TH2
// output tile in registers as packed sample pairs and feeds it straight into tile4_accumulate() with the matching fenc
// tile: the interpolated block never exists in memory.  A zero fraction runs through the identity taps {0,0,0,64,0,0,0,0}:
// for pixel inputs hps(0) == p2s and the two-pass result equals luma_vpp / luma_hpp / a plain copy bit for bit (the
// rounding offsets cancel), so one code path serves all sixteen (xFrac, yFrac) pairs.
// TAPS = 4: the chroma term of subpelCompare (motion.cpp:1805-1865), fractions in eighths (xFrac | yFrac << 4, 0..7 each);
// a negative frac marks a candidate the reference costs without its chroma term, and `accumulate` adds onto the luma cost.
template<typename PIX, int OP, int TW1, int TH2, int TAPS = 8>
__global__ void __launch_bounds__(128)
subpel_cmp_kernel(const PIX* __restrict__ fenc, intptr_t sf, const PIX* __restrict__ ref, intptr_t ss,
                  const int32_t* __restrict__ offF, const int32_t* __restrict__ offR, const int32_t* __restrict__ frac,
                  int K, int n, int w, int h, int G, int shift1, int offset1, int shift2, int offset2, int maxVal, int32_t* __restrict__ cost,
                  int accumulate = 0)
{
    extern __shared__ __align__(16) int16_t immed_all[];
    int lg = __ffs(G) - 1;
    int grp = threadIdx.x >> lg, l = threadIdx.x & (G - 1);
    int cand = blockIdx.x * (128 >> lg) + grp;
    const bool exists = cand < n;
    bool live = exists && frac[cand] >= 0;
    int rows = h + TAPS - 1;
    int16_t* immed = immed_all + (size_t)grp * w * rows;
    constexpr int NQ = (TW1 + TAPS - 1 + 3) / 4;
    uint32_t tx0 = 0, tx1 = 0, ty0 = 0, ty1 = 0;
    if (live)
    {
        int ci = frac[cand];
        int idxX = ci & 15, idxY = (ci >> 4) & 15;
        tx0 = TAPS == 8 ? c_lumaTapsB[idxX & 3][0] : c_chromaTapsB[idxX & 7]; tx1 = TAPS == 8 ? c_lumaTapsB[idxX & 3][1] : 0;
        ty0 = TAPS == 8 ? c_lumaTapsB[idxY & 3][0] : c_chromaTapsB[idxY & 7]; ty1 = TAPS == 8 ? c_lumaTapsB[idxY & 3][1] : 0;
        const PIX* s = ref + offR[cand] - (TAPS / 2 - 1) * ss - (TAPS / 2 - 1);
        int tw = w / TW1;
        int tiles1 = tw * ((rows + 3) >> 2);
        for (int t = l; t < tiles1; t += G)
        {
            int tx = (t % tw) * TW1, ty = (t / tw) << 2;
            uint32_t wv[4][2 * NQ + 1];
            if (ty + 4 <= rows)
                load_rows_quads<NQ, 4>(s + (intptr_t)ty * ss + tx, ss, wv);
            else
            {
#pragma unroll
                for (int r = 0; r < 4; r++)
                    if (ty + r < rows) { load_row_quads<NQ>(s + (intptr_t)(ty + r) * ss + tx, (uint32_t(&)[2 * NQ])wv[r]); wv[r][2 * NQ] = 0; }
            }
#pragma unroll
            for (int r = 0; r < 4; r++)
            {
                if (ty + r >= rows) break;
                uint32_t ws[2 * NQ];
#pragma unroll
                for (int i = 0; i < 2 * NQ; i++) ws[i] = __funnelshift_r(wv[r][i], wv[r][i + 1], 16);
#pragma unroll
                for (int o4 = 0; o4 < TW1; o4 += 4)
                {
                    int v[4];
#pragma unroll
                    for (int o = 0; o < 4; o++)
                    {
                        int i = (o4 + o) >> 1;
                        int sum = ((o4 + o) & 1) ? fir_pairs<PIX, TAPS>(ws[i], ws[i + 1], ws[(i + 2) % (2 * NQ)], ws[(i + 3) % (2 * NQ)], tx0, tx1, offset1)
                                                 : fir_pairs<PIX, TAPS>(wv[r][i], wv[r][i + 1], wv[r][(i + 2) % (2 * NQ + 1)], wv[r][(i + 3) % (2 * NQ + 1)], tx0, tx1, offset1);
                        v[o] = sum >> shift1;
                    }
                    *(uint2*)(immed + (ty + r) * w + tx + o4) = make_uint2(__byte_perm((uint32_t)v[0], (uint32_t)v[1], 0x5410),
                                                                           __byte_perm((uint32_t)v[2], (uint32_t)v[3], 0x5410));
                }
            }
        }
    }
    __syncwarp();
    int acc = 0;
    if (live)
    {
        const PIX* f = fenc + offF[K > 1 ? cand / K : cand];
        int tw = w >> 2;
        int tiles2 = tw * (h / TH2);
        const uint32_t mx = (uint32_t)maxVal * 0x10001u;
        for (int t = l; t < tiles2; t += G)
        {
            int tx = (t % tw) << 2, ty = (t / tw) * TH2;
            constexpr int NR = TH2 + TAPS - 1;
            uint2 q[NR];
#pragma unroll
            for (int r = 0; r < NR; r++) q[r] = *(const uint2*)(immed + (ty + r) * w + tx);
            uint32_t pr[NR - 1][4];
#pragma unroll
            for (int r = 0; r < NR - 1; r++)
            {
                pr[r][0] = __byte_perm(q[r].x, q[r + 1].x, 0x5410); pr[r][1] = __byte_perm(q[r].x, q[r + 1].x, 0x7632);
                pr[r][2] = __byte_perm(q[r].y, q[r + 1].y, 0x5410); pr[r][3] = __byte_perm(q[r].y, q[r + 1].y, 0x7632);
            }
#pragma unroll
            for (int r4 = 0; r4 < TH2; r4 += 4)
            {
                uint32_t blo[4], bhi[4], alo[4], ahi[4];
#pragma unroll
                for (int rr = 0; rr < 4; rr++)
                {
                    int r = r4 + rr;
                    int v[4];
#pragma unroll
                    for (int o = 0; o < 4; o++)
                        v[o] = fir_pairs<int16_t, TAPS>(pr[r][o], pr[r + 2][o], pr[(r + 4) % (NR - 1)][o], pr[(r + 6) % (NR - 1)][o], ty0, ty1, offset2) >> shift2;
                    blo[rr] = pack_clip2(v[0], v[1], mx);
                    bhi[rr] = pack_clip2(v[2], v[3], mx);
                }
                load_tile4x4(f + (intptr_t)(ty + r4) * sf + tx, sf, alo, ahi);
                tile4_accumulate<OP, int>(alo, ahi, blo, bhi, acc);
            }
        }
    }
    acc = group_sum(acc, G);
    if (exists && l == 0)
    {
        if (!accumulate) cost[cand] = acc;
        else if (live) cost[cand] += acc;
    }
}

// Bi-prediction candidate cost (reference encoder/search.cpp:442-448): both lists' motion-compensated luma blocks
// (predInterLumaPixel = the same copy / hpp / vpp / hvpp choice as subpelCompare, through the identity taps here), their
// rounded average (pixelavg_pp) and SATD against fenc.  Same two-pass structure as subpel_cmp_kernel with two
// intermediate tiles in shared memory; the average is taken on packed pairs between the vertical pass and the SATD.
template<typename PIX, int TW1>
__device__ __forceinline__ void bidir_stage1(const PIX* __restrict__ s, intptr_t ss, int16_t* immed, int w, int rows, int l, int G,
                                             uint32_t tx0, uint32_t tx1, int shift1, int offset1)
{
    constexpr int TAPS = 8;
    constexpr int NQ = (TW1 + TAPS - 1 + 3) / 4;
    int tw = w / TW1;
    int tiles1 = tw * ((rows + 3) >> 2);
    for (int t = l; t < tiles1; t += G)
    {
        int tx = (t % tw) * TW1, ty = (t / tw) << 2;
        uint32_t wv[4][2 * NQ + 1];
        if (ty + 4 <= rows)
            load_rows_quads<NQ, 4>(s + (intptr_t)ty * ss + tx, ss, wv);
        else
        {
#pragma unroll
            for (int r = 0; r < 4; r++)
                if (ty + r < rows) { load_row_quads<NQ>(s + (intptr_t)(ty + r) * ss + tx, (uint32_t(&)[2 * NQ])wv[r]); wv[r][2 * NQ] = 0; }
        }
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            if (ty + r >= rows) break;
            uint32_t ws[2 * NQ];
#pragma unroll
            for (int i = 0; i < 2 * NQ; i++) ws[i] = __funnelshift_r(wv[r][i], wv[r][i + 1], 16);
#pragma unroll
            for (int o4 = 0; o4 < TW1; o4 += 4)
            {
                int v[4];
#pragma unroll
                for (int o = 0; o < 4; o++)
                {
                    int i = (o4 + o) >> 1;
                    int sum = ((o4 + o) & 1) ? fir_pairs<PIX, TAPS>(ws[i], ws[i + 1], ws[(i + 2) % (2 * NQ)], ws[(i + 3) % (2 * NQ)], tx0, tx1, offset1)
                                             : fir_pairs<PIX, TAPS>(wv[r][i], wv[r][i + 1], wv[r][(i + 2) % (2 * NQ + 1)], wv[r][(i + 3) % (2 * NQ + 1)], tx0, tx1, offset1);
                    v[o] = sum >> shift1;
                }
                *(uint2*)(immed + (ty + r) * w + tx + o4) = make_uint2(__byte_perm((uint32_t)v[0], (uint32_t)v[1], 0x5410),
                                                                       __byte_perm((uint32_t)v[2], (uint32_t)v[3], 0x5410));
            }
        }
    }
}

// vertical pass of one 4 x TH2 tile out of the intermediate tile: packed, clipped pixel pairs per row
template<int TH2>
__device__ __forceinline__ void bidir_stage2(const int16_t* immed, int w, int tx, int ty, uint32_t ty0, uint32_t ty1, int shift2, int offset2,
                                             uint32_t mx, uint32_t (&lo)[TH2], uint32_t (&hi)[TH2])
{
    constexpr int TAPS = 8;
    constexpr int NR = TH2 + TAPS - 1;
    uint2 q[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) q[r] = *(const uint2*)(immed + (ty + r) * w + tx);
    uint32_t pr[NR - 1][4];
#pragma unroll
    for (int r = 0; r < NR - 1; r++)
    {
        pr[r][0] = __byte_perm(q[r].x, q[r + 1].x, 0x5410); pr[r][1] = __byte_perm(q[r].x, q[r + 1].x, 0x7632);
        pr[r][2] = __byte_perm(q[r].y, q[r + 1].y, 0x5410); pr[r][3] = __byte_perm(q[r].y, q[r + 1].y, 0x7632);
    }
#pragma unroll
    for (int r = 0; r < TH2; r++)
    {
        int v[4];
#pragma unroll
        for (int o = 0; o < 4; o++)
            v[o] = fir_pairs<int16_t, TAPS>(pr[r][o], pr[r + 2][o], pr[(r + 4) % (NR - 1)][o], pr[(r + 6) % (NR - 1)][o], ty0, ty1, offset2) >> shift2;
        lo[r] = pack_clip2(v[0], v[1], mx);
        hi[r] = pack_clip2(v[2], v[3], mx);
    }
}

template<typename PIX, int TW1, int TH2>
__global__ void __launch_bounds__(128)
bidir_satd_kernel(const PIX* __restrict__ fenc, intptr_t sf, const int32_t* __restrict__ offF,
                  const PIX* __restrict__ ref0, intptr_t ss0, const int32_t* __restrict__ off0, const int32_t* __restrict__ frac0,
                  const PIX* __restrict__ ref1, intptr_t ss1, const int32_t* __restrict__ off1, const int32_t* __restrict__ frac1,
                  int n, int w, int h, int G, int shift1, int offset1, int shift2, int offset2, int maxVal, int32_t* __restrict__ cost)
{
    constexpr int TAPS = 8;
    extern __shared__ __align__(16) int16_t immed_all[];
    int lg = __ffs(G) - 1;
    int grp = threadIdx.x >> lg, l = threadIdx.x & (G - 1);
    int cand = blockIdx.x * (128 >> lg) + grp;
    bool live = cand < n;
    int rows = h + TAPS - 1;
    int16_t* immed0 = immed_all + (size_t)grp * 2 * w * rows;
    int16_t* immed1 = immed0 + (size_t)w * rows;
    uint32_t ty00 = 0, ty01 = 0, ty10 = 0, ty11 = 0;
    if (live)
    {
        int c0 = frac0[cand], c1 = frac1[cand];
        ty00 = c_lumaTapsB[(c0 >> 4) & 3][0]; ty01 = c_lumaTapsB[(c0 >> 4) & 3][1];
        ty10 = c_lumaTapsB[(c1 >> 4) & 3][0]; ty11 = c_lumaTapsB[(c1 >> 4) & 3][1];
        bidir_stage1<PIX, TW1>(ref0 + off0[cand] - (TAPS / 2 - 1) * ss0 - (TAPS / 2 - 1), ss0, immed0, w, rows, l, G,
                               c_lumaTapsB[c0 & 3][0], c_lumaTapsB[c0 & 3][1], shift1, offset1);
        bidir_stage1<PIX, TW1>(ref1 + off1[cand] - (TAPS / 2 - 1) * ss1 - (TAPS / 2 - 1), ss1, immed1, w, rows, l, G,
                               c_lumaTapsB[c1 & 3][0], c_lumaTapsB[c1 & 3][1], shift1, offset1);
    }
    __syncwarp();
    int acc = 0;
    if (live)
    {
        const PIX* f = fenc + offF[cand];
        int tw = w >> 2;
        int tiles2 = tw * (h / TH2);
        const uint32_t mx = (uint32_t)maxVal * 0x10001u;
        for (int t = l; t < tiles2; t += G)
        {
            int tx = (t % tw) << 2, ty = (t / tw) * TH2;
            uint32_t alo[TH2], ahi[TH2], blo[TH2], bhi[TH2];
            bidir_stage2<TH2>(immed0, w, tx, ty, ty00, ty01, shift2, offset2, mx, alo, ahi);
            bidir_stage2<TH2>(immed1, w, tx, ty, ty10, ty11, shift2, offset2, mx, blo, bhi);
#pragma unroll
            for (int r = 0; r < TH2; r++)
            {   // pixelavg_pp on packed pairs (samples < 2^15: the halves cannot carry into each other)
                alo[r] = ((alo[r] + blo[r] + 0x00010001u) >> 1) & 0x7fff7fffu;
                ahi[r] = ((ahi[r] + bhi[r] + 0x00010001u) >> 1) & 0x7fff7fffu;
            }
#pragma unroll
            for (int r4 = 0; r4 < TH2; r4 += 4)
            {
                uint32_t flo[4], fhi[4], plo[4], phi[4];
#pragma unroll
                for (int rr = 0; rr < 4; rr++) { plo[rr] = alo[r4 + rr]; phi[rr] = ahi[r4 + rr]; }
                load_tile4x4(f + (intptr_t)(ty + r4) * sf + tx, sf, flo, fhi);
                tile4_accumulate<OP_SATD, int>(flo, fhi, plo, phi, acc);
            }
        }
    }
    acc = group_sum(acc, G);
    if (live && l == 0) cost[cand] = acc;
}

// p2s, 4 samples per thread
template<typename PIX>
__global__ void __launch_bounds__(256)
p2s_tile_kernel(const PIX* __restrict__ src, intptr_t ss, const int32_t* __restrict__ offSrc,
                int16_t* __restrict__ dst, intptr_t ds, const int32_t* __restrict__ offDst, int n, int w, int h, int shift)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int qw = w >> 2;
    int per = qw * h;
    int blk = (int)(gid / per);
    if (blk >= n) return;
    int r = (int)(gid % per);
    int y = r / qw, x = (r % qw) << 2;
    uint32_t wv[2];
    load_row_quads<1>(src + offSrc[blk] + (intptr_t)y * ss + x, wv);
    int a[4];
    SampleTraits<PIX>::unpack(wv[0], a[0], a[1]);
    SampleTraits<PIX>::unpack(wv[1], a[2], a[3]);
    int v[4];
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = (int)(int16_t)((int)(int16_t)(a[i] << shift) - 8192);
    store4(dst + offDst[blk] + (intptr_t)y * ds + x, v);
}

// p2s for blocks whose width is a multiple of 8 and height a multiple of 4: one thread per 8x4 strip.
// (pix << shift) - 8192 runs on packed pairs: pix << shift < 2^14, and the lane-wise subtraction is borrow-free
// (device_util.cuh psub16).
template<typename PIX>
__global__ void __launch_bounds__(256)
p2s_wide_kernel(const PIX* __restrict__ src, intptr_t ss, const int32_t* __restrict__ offSrc,
                int16_t* __restrict__ dst, intptr_t ds, const int32_t* __restrict__ offDst, int n, int w, int h, int shift)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int sw = w >> 3;
    int per = sw * (h >> 2);
    int blk = (int)(gid / per);
    if (blk >= n) return;
    int r = (int)(gid - (long long)blk * per);
    int y = (r / sw) << 2, x = (r % sw) << 3;
    uint32_t wv[4][4];
    load_rows8<4>(src + offSrc[blk] + (intptr_t)y * ss + x, ss, wv);
    int16_t* d = dst + offDst[blk] + (intptr_t)y * ds + x;
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) o[k] = psub16(wv[i][k] << shift, 0x20002000u);
        store8_s16(d + i * ds, o);
    }
}

template<typename PIX>
__global__ void __launch_bounds__(256)
p2s_kernel(const PIX* __restrict__ src, intptr_t ss, const int32_t* __restrict__ offSrc,
           int16_t* __restrict__ dst, intptr_t ds, const int32_t* __restrict__ offDst, int n, int w, int h, int shift)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int per = w * h;
    int blk = (int)(gid / per);
    if (blk >= n) return;
    int r = (int)(gid % per);
    int y = r / w, x = r % w;
    int16_t val = (int16_t)((int)src[offSrc[blk] + (intptr_t)y * ss + x] << shift);      // ipfilter.cpp:49-50
    dst[offDst[blk] + (intptr_t)y * ds + x] = (int16_t)(val - (int16_t)8192);
}

template<typename PIX, int TAPS>
static int launch_interp(x265b200_ctx* ctx, int kind, int w, int h, const void* src, intptr_t ss, const int32_t* offSrc,
                         void* dst, intptr_t ds, const int32_t* offDst, const int32_t* coeffIdx, int n, cudaStream_t st)
{
    const int D = ctx->depth;
    const int headRoom = 14 - D;                 // IF_INTERNAL_PREC - X265_DEPTH
    const int maxVal = (1 << D) - 1;
    FirParams p; p.w = w; p.h = h; p.rowExtKind = 0;
    long long threads = (long long)n * w * h;
    int grid = ceil_div(threads, 256);
    const bool tiled = !(w & 3) && !(ss & 3);
#define FIR(SRC_T, DST_T, VERT_, ROWS)                                                                          \
    do {                                                                                                        \
        if (tiled && !(VERT_) && !(w & 7))                                                                      \
            fir_tile_kernel<SRC_T, DST_T, TAPS, VERT_, 8, 4><<<ceil_div((long long)n * (w >> 3) * (((ROWS) + 3) >> 2), 128), 128, 0, st>>>( \
                (const SRC_T*)src, ss, offSrc, (DST_T*)dst, ds, offDst, coeffIdx, n, p);                         \
        else if (tiled && (VERT_) && !((ROWS) & 7))                                                             \
            fir_tile_kernel<SRC_T, DST_T, TAPS, VERT_, 4, 8><<<ceil_div((long long)n * (w >> 2) * ((ROWS) >> 3), 128), 128, 0, st>>>( \
                (const SRC_T*)src, ss, offSrc, (DST_T*)dst, ds, offDst, coeffIdx, n, p);                         \
        else if (tiled)                                                                                         \
            fir_tile_kernel<SRC_T, DST_T, TAPS, VERT_, 4, 4><<<ceil_div((long long)n * (w >> 2) * (((ROWS) + 3) >> 2), 128), 128, 0, st>>>( \
                (const SRC_T*)src, ss, offSrc, (DST_T*)dst, ds, offDst, coeffIdx, n, p);                         \
        else                                                                                                    \
            fir_kernel<SRC_T, DST_T, TAPS, VERT_><<<ceil_div((long long)n * w * (ROWS), 256), 256, 0, st>>>(     \
                (const SRC_T*)src, ss, offSrc, (DST_T*)dst, ds, offDst, coeffIdx, n, p);                         \
    } while (0)
    switch (kind)
    {
    case X265B200_IP_HPP:
        p.shift = 6; p.offset = 32; p.maxVal = maxVal;
        FIR(PIX, PIX, false, h);
        break;
    case X265B200_IP_VPP:
        p.shift = 6; p.offset = 32; p.maxVal = maxVal;
        FIR(PIX, PIX, true, h);
        break;
    case X265B200_IP_HPS:
        p.shift = 6 - headRoom; p.offset = (int)((unsigned)-8192 << p.shift); p.maxVal = -1; p.rowExtKind = 1;
        FIR(PIX, int16_t, false, h + TAPS - 1);
        break;
    case X265B200_IP_VPS:
        p.shift = 6 - headRoom; p.offset = (int)((unsigned)-8192 << p.shift); p.maxVal = -1;
        FIR(PIX, int16_t, true, h);
        break;
    case X265B200_IP_VSP:
        p.shift = 6 + headRoom; p.offset = (1 << (p.shift - 1)) + (8192 << 6); p.maxVal = maxVal;
        FIR(int16_t, PIX, true, h);
        break;
    case X265B200_IP_VSS:
        p.shift = 6; p.offset = 0; p.maxVal = -1;
        FIR(int16_t, int16_t, true, h);
        break;
    case X265B200_IP_HVPP:
    {
        int shift1 = 6 - headRoom, shift2 = 6 + headRoom;
        size_t smem = (size_t)w * (h + TAPS - 1) * sizeof(int16_t);
        if (tiled && !(h & 3))
        {
            // lanes per block: about two pass-2 tiles (4 x 8 or 4 x 4) per lane
            const bool wide = !(w & 7), tall = !(h & 7);
            int tiles2 = (w >> 2) * (tall ? h >> 3 : h >> 2);
            int G = 1;
            while (G * 2 <= tiles2 / 2 && G < 32) G <<= 1;
            int perCta = 128 / G;
#define HV(TW1_, TH2_) hv_tile_kernel<PIX, TAPS, TW1_, TH2_><<<ceil_div(n, perCta), 128, perCta * smem, st>>>((const PIX*)src, ss, offSrc, (PIX*)dst, ds, offDst, coeffIdx, n, w, h, \
                                                                          G, shift1, (int)((unsigned)-8192 << shift1), shift2, (1 << (shift2 - 1)) + (8192 << 6), maxVal)
            if (wide && tall) HV(8, 8); else if (wide) HV(8, 4); else if (tall) HV(4, 8); else HV(4, 4);
#undef HV
            break;
        }
        hv_kernel<PIX, TAPS><<<n, 256, smem, st>>>((const PIX*)src, ss, offSrc, (PIX*)dst, ds, offDst, coeffIdx, w, h,
                                                  shift1, (int)((unsigned)-8192 << shift1), shift2,
                                                  (1 << (shift2 - 1)) + (8192 << 6), maxVal);
        break;
    }
    case X265B200_IP_P2S:
        if (tiled && !(w & 7) && !(h & 3))
            p2s_wide_kernel<PIX><<<ceil_div((long long)n * (w >> 3) * (h >> 2), 256), 256, 0, st>>>((const PIX*)src, ss, offSrc, (int16_t*)dst, ds, offDst, n, w, h, headRoom);
        else if (tiled)
            p2s_tile_kernel<PIX><<<ceil_div((long long)n * (w >> 2) * h, 256), 256, 0, st>>>((const PIX*)src, ss, offSrc, (int16_t*)dst, ds, offDst, n, w, h, headRoom);
        else
            p2s_kernel<PIX><<<grid, 256, 0, st>>>((const PIX*)src, ss, offSrc, (int16_t*)dst, ds, offDst, n, w, h, headRoom);
        break;
    default:
        return fail(ctx, X265B200_ERR_ARG, "interp: unknown kind");
    }
#undef FIR
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

} // namespace b200

using namespace b200;

extern "C" int x265b200_interp_batch(x265b200_ctx* ctx, int kind, int taps, int w, int h, const void* src, intptr_t ss,
                                     const int32_t* offSrc, void* dst, intptr_t ds, const int32_t* offDst,
                                     const int32_t* coeffIdx, int n, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if ((taps != 8 && taps != 4) || w < 1 || h < 1 || w > 64 || h > 64 || n < 0)
        return fail(ctx, X265B200_ERR_ARG, "interp: bad taps / shape");
    if (n == 0) return X265B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->pixbytes == 1)
        return taps == 8 ? launch_interp<uint8_t, 8>(ctx, kind, w, h, src, ss, offSrc, dst, ds, offDst, coeffIdx, n, st)
                         : launch_interp<uint8_t, 4>(ctx, kind, w, h, src, ss, offSrc, dst, ds, offDst, coeffIdx, n, st);
    return taps == 8 ? launch_interp<uint16_t, 8>(ctx, kind, w, h, src, ss, offSrc, dst, ds, offDst, coeffIdx, n, st)
                     : launch_interp<uint16_t, 4>(ctx, kind, w, h, src, ss, offSrc, dst, ds, offDst, coeffIdx, n, st);
}

static int subpel_cmp_launch(x265b200_ctx* ctx, int taps, int accumulate, int op, int w, int h, const void* fenc, intptr_t sf, const void* ref, intptr_t sr,
                             const int32_t* offF, const int32_t* offR, const int32_t* frac, int K, int n, int32_t* cost, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (w < 4 || h < 4 || (w & 3) || (h & 3) || w > 64 || h > 64 || n < 0 || K < 1 || (op != X265B200_SAD && op != X265B200_SATD))
        return fail(ctx, X265B200_ERR_ARG, "subpel_cmp: bad shape / op");
    if ((sf | sr) & 3) return fail(ctx, X265B200_ERR_ARG, "subpel_cmp: plane strides must be multiples of 4 samples");
    if (n == 0) return X265B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int headRoom = 14 - ctx->depth, maxVal = (1 << ctx->depth) - 1;
    const int shift1 = 6 - headRoom, shift2 = 6 + headRoom;
    const int offset1 = (int)((unsigned)-8192 << shift1), offset2 = (1 << (shift2 - 1)) + (8192 << 6);
    const size_t smem = (size_t)w * (h + taps - 1) * sizeof(int16_t);
    const bool wide = !(w & 7), tall = !(h & 7);
    int tiles2 = (w >> 2) * (tall ? h >> 3 : h >> 2);
    int G = 1;
    while (G * 2 <= tiles2 / 2 && G < 32) G <<= 1;
    int perCta = 128 / G;
    long long cands = (long long)n * K;
    if (cands > 0x7fffffff) return fail(ctx, X265B200_ERR_ARG, "subpel_cmp: too many candidates");
    // the 25 PU shapes need at most 47104 bytes (16x16); other multiples of 4 (28x8: 53760) go over the 48 KB default
#define SP(PIX, OP_, TW1_, TH2_, TAPS_) do { \
        if (perCta * smem > 48 * 1024) B200_CUDA(ctx, cudaFuncSetAttribute(subpel_cmp_kernel<PIX, OP_, TW1_, TH2_, TAPS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(perCta * smem))); \
        subpel_cmp_kernel<PIX, OP_, TW1_, TH2_, TAPS_><<<ceil_div(cands, perCta), 128, perCta * smem, st>>>(  \
        (const PIX*)fenc, sf, (const PIX*)ref, sr, offF, offR, frac, K, (int)cands, w, h, G, shift1, offset1, shift2, offset2, maxVal, cost, accumulate); } while (0)
#define SP_SHAPE(PIX, OP_, TAPS_) do { if (wide && tall) SP(PIX, OP_, 8, 8, TAPS_); else if (wide) SP(PIX, OP_, 8, 4, TAPS_); \
                                       else if (tall) SP(PIX, OP_, 4, 8, TAPS_); else SP(PIX, OP_, 4, 4, TAPS_); } while (0)
    if (taps == 4)
    {   // chroma term: always SATD
        if (ctx->pixbytes == 1) SP_SHAPE(uint8_t, OP_SATD, 4); else SP_SHAPE(uint16_t, OP_SATD, 4);
    }
    else if (ctx->pixbytes == 1) { if (op == X265B200_SAD) SP_SHAPE(uint8_t, OP_SAD, 8); else SP_SHAPE(uint8_t, OP_SATD, 8); }
    else { if (op == X265B200_SAD) SP_SHAPE(uint16_t, OP_SAD, 8); else SP_SHAPE(uint16_t, OP_SATD, 8); }
#undef SP_SHAPE
#undef SP
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

extern "C" int x265b200_subpel_cmp_batch(x265b200_ctx* ctx, int op, int w, int h, const void* fenc, intptr_t sf, const void* ref, intptr_t sr,
                                         const int32_t* offF, const int32_t* offR, const int32_t* frac, int K, int n, int32_t* cost,
                                         x265b200_stream stream)
{
    return subpel_cmp_launch(ctx, 8, 0, op, w, h, fenc, sf, ref, sr, offF, offR, frac, K, n, cost, stream);
}

extern "C" int x265b200_subpel_cmp_chroma_batch(x265b200_ctx* ctx, int w, int h, const void* fenc, intptr_t sf, const void* ref, intptr_t sr,
                                                const int32_t* offF, const int32_t* offR, const int32_t* frac, int K, int n, int32_t* cost,
                                                int accumulate, x265b200_stream stream)
{
    return subpel_cmp_launch(ctx, 4, accumulate ? 1 : 0, X265B200_SATD, w, h, fenc, sf, ref, sr, offF, offR, frac, K, n, cost, stream);
}

extern "C" int x265b200_bidir_satd_batch(x265b200_ctx* ctx, int w, int h, const void* fenc, intptr_t sf, const int32_t* offF,
                                         const void* ref0, intptr_t sr0, const int32_t* off0, const int32_t* frac0,
                                         const void* ref1, intptr_t sr1, const int32_t* off1, const int32_t* frac1,
                                         int n, int32_t* cost, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (w < 4 || h < 4 || (w & 3) || (h & 3) || w > 64 || h > 64 || n < 0) return fail(ctx, X265B200_ERR_ARG, "bidir_satd: bad shape");
    if ((sf | sr0 | sr1) & 3) return fail(ctx, X265B200_ERR_ARG, "bidir_satd: plane strides must be multiples of 4 samples");
    if (n == 0) return X265B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int headRoom = 14 - ctx->depth, maxVal = (1 << ctx->depth) - 1;
    const int shift1 = 6 - headRoom, shift2 = 6 + headRoom;
    const int offset1 = (int)((unsigned)-8192 << shift1), offset2 = (1 << (shift2 - 1)) + (8192 << 6);
    const size_t smem = 2 * (size_t)w * (h + 7) * sizeof(int16_t);
    const bool wide = !(w & 7), tall = !(h & 7);
    int tiles2 = (w >> 2) * (tall ? h >> 3 : h >> 2);
    int G = 1;
    while (G * 2 <= tiles2 / 2 && G < 32) G <<= 1;
    int perCta = 128 / G;
    const size_t total = perCta * smem;
#define BD(PIX, TW1_, TH2_) do {                                                                                                        \
        if (total > 48 * 1024)                                                                                                          \
            B200_CUDA(ctx, cudaFuncSetAttribute((const void*)bidir_satd_kernel<PIX, TW1_, TH2_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total)); \
        bidir_satd_kernel<PIX, TW1_, TH2_><<<ceil_div(n, perCta), 128, total, st>>>((const PIX*)fenc, sf, offF, (const PIX*)ref0, sr0, off0, frac0, \
            (const PIX*)ref1, sr1, off1, frac1, n, w, h, G, shift1, offset1, shift2, offset2, maxVal, cost); } while (0)
#define BD_SHAPE(PIX) do { if (wide && tall) BD(PIX, 8, 8); else if (wide) BD(PIX, 8, 4); else if (tall) BD(PIX, 4, 8); else BD(PIX, 4, 4); } while (0)
    if (ctx->pixbytes == 1) BD_SHAPE(uint8_t); else BD_SHAPE(uint16_t);
#undef BD_SHAPE
#undef BD
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}
