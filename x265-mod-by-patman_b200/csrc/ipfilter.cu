// ipfilter.cu -- HEVC sub-pel interpolation: 8-tap luma / 4-tap chroma FIR in all seven x265
// variants plus pixel-to-short.  Bit-exact CUDA restatement of the reference's ipfilter.cpp:40-369.
//
// The variants differ only in input/output type, rounding offset, shift and clipping
// (SURVEY.md appendix C), so one templated FIR kernel covers hpp/hps/vpp/vps/vsp/vss; hvpp runs
// the horizontal pixel->short pass into a shared-memory tile and the vertical short->pixel pass
// out of it inside one CTA (the reference does the same through a stack buffer, ipfilter.cpp:362-369).
#include "internal.h"
#include "device_util.cuh"

namespace b200 {

__constant__ short c_lumaTaps[4][8];
__constant__ short c_chromaTaps[8][4];

int upload_filter_tables(x265b200_ctx* ctx)
{
    // HEVC fractional-sample filters (== g_lumaFilter / g_chromaFilter, constants.cpp:250-268)
    static const short luma[4][8] = { { 0, 0, 0, 64, 0, 0, 0, 0 }, { -1, 4, -10, 58, 17, -5, 1, 0 },
                                      { -1, 4, -11, 40, 40, -11, 4, -1 }, { 0, 1, -5, 17, 58, -10, 4, -1 } };
    static const short chroma[8][4] = { { 0, 64, 0, 0 }, { -2, 58, 10, -2 }, { -4, 54, 16, -2 }, { -6, 46, 28, -4 },
                                        { -4, 36, 36, -4 }, { -4, 28, 46, -6 }, { -2, 16, 54, -4 }, { -2, 10, 58, -2 } };
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
    switch (kind)
    {
    case X265B200_IP_HPP:
        p.shift = 6; p.offset = 32; p.maxVal = maxVal;
        fir_kernel<PIX, PIX, TAPS, false><<<grid, 256, 0, st>>>((const PIX*)src, ss, offSrc, (PIX*)dst, ds, offDst, coeffIdx, n, p);
        break;
    case X265B200_IP_VPP:
        p.shift = 6; p.offset = 32; p.maxVal = maxVal;
        fir_kernel<PIX, PIX, TAPS, true><<<grid, 256, 0, st>>>((const PIX*)src, ss, offSrc, (PIX*)dst, ds, offDst, coeffIdx, n, p);
        break;
    case X265B200_IP_HPS:
        p.shift = 6 - headRoom; p.offset = (int)((unsigned)-8192 << p.shift); p.maxVal = -1; p.rowExtKind = 1;
        grid = ceil_div((long long)n * w * (h + TAPS - 1), 256);
        fir_kernel<PIX, int16_t, TAPS, false><<<grid, 256, 0, st>>>((const PIX*)src, ss, offSrc, (int16_t*)dst, ds, offDst, coeffIdx, n, p);
        break;
    case X265B200_IP_VPS:
        p.shift = 6 - headRoom; p.offset = (int)((unsigned)-8192 << p.shift); p.maxVal = -1;
        fir_kernel<PIX, int16_t, TAPS, true><<<grid, 256, 0, st>>>((const PIX*)src, ss, offSrc, (int16_t*)dst, ds, offDst, coeffIdx, n, p);
        break;
    case X265B200_IP_VSP:
        p.shift = 6 + headRoom; p.offset = (1 << (p.shift - 1)) + (8192 << 6); p.maxVal = maxVal;
        fir_kernel<int16_t, PIX, TAPS, true><<<grid, 256, 0, st>>>((const int16_t*)src, ss, offSrc, (PIX*)dst, ds, offDst, coeffIdx, n, p);
        break;
    case X265B200_IP_VSS:
        p.shift = 6; p.offset = 0; p.maxVal = -1;
        fir_kernel<int16_t, int16_t, TAPS, true><<<grid, 256, 0, st>>>((const int16_t*)src, ss, offSrc, (int16_t*)dst, ds, offDst, coeffIdx, n, p);
        break;
    case X265B200_IP_HVPP:
    {
        int shift1 = 6 - headRoom, shift2 = 6 + headRoom;
        size_t smem = (size_t)w * (h + TAPS - 1) * sizeof(int16_t);
        hv_kernel<PIX, TAPS><<<n, 256, smem, st>>>((const PIX*)src, ss, offSrc, (PIX*)dst, ds, offDst, coeffIdx, w, h,
                                                  shift1, (int)((unsigned)-8192 << shift1), shift2,
                                                  (1 << (shift2 - 1)) + (8192 << 6), maxVal);
        break;
    }
    case X265B200_IP_P2S:
        p2s_kernel<PIX><<<grid, 256, 0, st>>>((const PIX*)src, ss, offSrc, (int16_t*)dst, ds, offDst, n, w, h, headRoom);
        break;
    default:
        return fail(ctx, X265B200_ERR_ARG, "interp: unknown kind");
    }
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
