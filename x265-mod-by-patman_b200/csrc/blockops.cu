// blockops.cu -- slots adjacent to the hot path that sit in the same call chains:
// the two-input block operations around the TU chain and bi-prediction, and the lookahead's lowres downscale.
//     sub_ps       cu[].sub_ps        pixel.cpp:806-818   int16 = pixel - pixel
//     add_ps       cu[].add_ps        pixel.cpp:820-832   pixel = clip(pixel + int16)
//     pixelavg_pp  pu[].pixelavg_pp   pixel.cpp:537-549   pixel = (pixel + pixel + 1) >> 1
//     addAvg       pu[].addAvg        pixel.cpp:834-855   pixel = clip((int16 + int16 + offset) >> shift)
//     lowres       frameInitLowres    pixel.cpp:595-620   four half-resolution planes (full / h / v / c half-pel phases)
#include "internal.h"
#include "device_util.cuh"

namespace b200 {

enum { BOP_SUB_PS = 0, BOP_ADD_PS = 1, BOP_PIXELAVG = 2, BOP_ADDAVG = 3 };

template<int OP> __device__ __forceinline__ int bop(int a, int b, int maxv, int shift, int offset)
{
    if (OP == BOP_SUB_PS) return (int)(int16_t)(a - b);
    if (OP == BOP_ADD_PS) return min(max(a + b, 0), maxv);
    if (OP == BOP_PIXELAVG) return (a + b + 1) >> 1;
    return min(max((a + b + offset) >> shift, 0), maxv);
}

__device__ __forceinline__ size_t blk_off(const int32_t* off, int blk, int wh) { return off ? (size_t)off[blk] : (size_t)blk * wh; }

// one thread per sample: any width, any stride
template<int OP, typename TA, typename TB, typename TD>
__global__ void __launch_bounds__(256)
blockop_kernel(const TA* __restrict__ A, intptr_t sa, const int32_t* __restrict__ offA, const TB* __restrict__ B, intptr_t sb,
               const int32_t* __restrict__ offB, TD* __restrict__ D, intptr_t sd, const int32_t* __restrict__ offD,
               int n, int w, int h, int maxv, int shift, int offset)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int per = w * h;
    int blk = (int)(gid / per);
    if (blk >= n) return;
    int r = (int)(gid - (long long)blk * per);
    int y = r / w, x = r - y * w;
    int a = A[blk_off(offA, blk, per) + (intptr_t)y * sa + x];
    int b = B[blk_off(offB, blk, per) + (intptr_t)y * sb + x];
    D[blk_off(offD, blk, per) + (intptr_t)y * sd + x] = (TD)bop<OP>(a, b, maxv, shift, offset);
}

template<typename TD> __device__ __forceinline__ void store_quad(TD* d, const int (&v)[4])
{
    uintptr_t a = (uintptr_t)d;
    if (sizeof(TD) == 2)
    {
        uint32_t lo = (uint32_t)(v[0] & 0xffff) | ((uint32_t)v[1] << 16), hi = (uint32_t)(v[2] & 0xffff) | ((uint32_t)v[3] << 16);
        if ((a & 7) == 0) *(uint2*)d = make_uint2(lo, hi);
        else if ((a & 3) == 0) { ((uint32_t*)d)[0] = lo; ((uint32_t*)d)[1] = hi; }
        else { d[0] = (TD)v[0]; d[1] = (TD)v[1]; d[2] = (TD)v[2]; d[3] = (TD)v[3]; }
    }
    else
    {
        if ((a & 3) == 0) *(uint32_t*)d = (uint32_t)(v[0] & 0xff) | ((uint32_t)(v[1] & 0xff) << 8) | ((uint32_t)(v[2] & 0xff) << 16) | ((uint32_t)v[3] << 24);
        else { d[0] = (TD)v[0]; d[1] = (TD)v[1]; d[2] = (TD)v[2]; d[3] = (TD)v[3]; }
    }
}

// one thread per 4 x 2 samples (width % 4 == 0, height % 2 == 0): widest loads the alignment allows
template<int OP, typename TA, typename TB, typename TD>
__global__ void __launch_bounds__(256)
blockop_quad_kernel(const TA* __restrict__ A, intptr_t sa, const int32_t* __restrict__ offA, const TB* __restrict__ B, intptr_t sb,
                    const int32_t* __restrict__ offB, TD* __restrict__ D, intptr_t sd, const int32_t* __restrict__ offD,
                    int n, int w, int h, int maxv, int shift, int offset)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int qw = w >> 2;
    int per = qw * (h >> 1);
    int blk = (int)(gid / per);
    if (blk >= n) return;
    int r = (int)(gid - (long long)blk * per);
    int y = (r / qw) << 1, x = (r % qw) << 2;
    const TA* a = A + blk_off(offA, blk, w * h) + (intptr_t)y * sa + x;
    const TB* b = B + blk_off(offB, blk, w * h) + (intptr_t)y * sb + x;
    TD* d = D + blk_off(offD, blk, w * h) + (intptr_t)y * sd + x;
    int va[2][4], vb[2][4];
    load4(a, va[0]); load4(a + sa, va[1]);
    load4(b, vb[0]); load4(b + sb, vb[1]);
#pragma unroll
    for (int k = 0; k < 2; k++)
    {
        int v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = bop<OP>(va[k][i], vb[k][i], maxv, shift, offset);
        store_quad(d + k * sd, v);
    }
}

// sub_ps / add_ps / pixelavg_pp on packed sample pairs, one thread per 8x4 strip (width % 8 == 0, height % 4 == 0,
// source strides % 4 == 0): grouped chunk loads for both inputs, 16-byte stores.
//   sub_ps   : lane-wise borrow-free subtraction (device_util.cuh psub16)
//   add_ps   : clip(p + r, 0, max) == clip(p + clamp(r, -max, max), 0, max) for p in [0, max] -> VIADDMNMX.S16x2.RELU
//   pixelavg : (a + b + 1) >> 1 per lane; a + b + 1 < 2^14 never carries into the neighbour lane
template<int OP, typename TA, typename TB, typename TD>
__global__ void __launch_bounds__(256)
blockop_wide_kernel(const TA* __restrict__ A, intptr_t sa, const int32_t* __restrict__ offA, const TB* __restrict__ B, intptr_t sb,
                    const int32_t* __restrict__ offB, TD* __restrict__ D, intptr_t sd, const int32_t* __restrict__ offD,
                    int n, int w, int h, int maxv)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int sw = w >> 3;
    int per = sw * (h >> 2);
    int blk = (int)(gid / per);
    if (blk >= n) return;
    int r = (int)(gid - (long long)blk * per);
    int y = (r / sw) << 2, x = (r % sw) << 3;
    uint32_t wa[4][4], wb[4][4];
    load_rows8<4>(A + blk_off(offA, blk, w * h) + (intptr_t)y * sa + x, sa, wa);
    load_rows8<4>(B + blk_off(offB, blk, w * h) + (intptr_t)y * sb + x, sb, wb);
    TD* d = D + blk_off(offD, blk, w * h) + (intptr_t)y * sd + x;
    const uint32_t mx = (uint32_t)maxv * 0x10001u, negmx = (uint32_t)(-maxv & 0xffff) * 0x10001u;
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            if (OP == BOP_SUB_PS) o[k] = psub16(wa[i][k], wb[i][k]);
            else if (OP == BOP_ADD_PS) o[k] = __viaddmin_s16x2_relu(wa[i][k], __vmins2(__vmaxs2(wb[i][k], negmx), mx), mx);
            else o[k] = ((wa[i][k] + wb[i][k] + 0x00010001u) >> 1) & 0x7fff7fffu;
        }
        TD* q = d + (intptr_t)i * sd;
        if (sizeof(TD) == 2) store8_s16((int16_t*)q, o);
        else
        {
            uint32_t b0 = __byte_perm(o[0], o[1], 0x6420), b1 = __byte_perm(o[2], o[3], 0x6420);
            if (((uintptr_t)q & 7) == 0) *(uint2*)q = make_uint2(b0, b1);
            else if (((uintptr_t)q & 3) == 0) { ((uint32_t*)q)[0] = b0; ((uint32_t*)q)[1] = b1; }
            else
            {
#pragma unroll
                for (int k = 0; k < 4; k++) { q[k] = (TD)((b0 >> (8 * k)) & 0xff); q[4 + k] = (TD)((b1 >> (8 * k)) & 0xff); }
            }
        }
    }
}

// copy family (pixel.cpp:386-461, 751-804): one thread per sample.  kind: 0 copy_pp, 1 copy_ss, 2 copy_sp ((pixel) cast),
// 3 copy_ps, 4 blockfill_s (param = value), 5 shl ((int16)((uint32)x << param): cpy2Dto1D_shl / cpy1Dto2D_shl),
// 6 shr ((x + (1 << (param - 1))) >> param: cpy2Dto1D_shr / cpy1Dto2D_shr)
template<typename TS, typename TD>
__global__ void __launch_bounds__(256)
blockcopy_kernel(int kind, const TS* __restrict__ S, intptr_t ss, const int32_t* __restrict__ offS, TD* __restrict__ D, intptr_t sd,
                 const int32_t* __restrict__ offD, int n, int w, int h, int param)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int per = w * h;
    int blk = (int)(gid / per);
    if (blk >= n) return;
    int r = (int)(gid - (long long)blk * per);
    int y = r / w, x = r - y * w;
    int v;
    if (kind == 4) v = param;
    else
    {
        v = (int)S[blk_off(offS, blk, per) + (intptr_t)y * ss + x];
        if (kind == 5) v = (int)(int16_t)((uint32_t)v << param);
        else if (kind == 6) v = (v + (int)(int16_t)(1 << (param - 1))) >> param;
    }
    D[blk_off(offD, blk, per) + (intptr_t)y * sd + x] = (TD)v;
}

// frame_init_lowres_core: one thread per lowres position (pixel.cpp:604-612; "slower than naive bilinear, but matches asm")
template<typename T>
__global__ void __launch_bounds__(256)
lowres_kernel(const T* __restrict__ src, intptr_t ss, T* __restrict__ d0, T* __restrict__ dh, T* __restrict__ dv, T* __restrict__ dc,
              intptr_t ds, int width, int height)
{
    int x = blockIdx.x * 32 + (threadIdx.x & 31);
    int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= width || y >= height) return;
    const T* s0 = src + (intptr_t)(2 * y) * ss + 2 * x;
    const T* s1 = s0 + ss;
    const T* s2 = s1 + ss;
    int a0 = s0[0], a1 = s0[1], a2 = s0[2], b0 = s1[0], b1 = s1[1], b2 = s1[2], c0 = s2[0], c1 = s2[1], c2 = s2[2];
#define LOWRES_FILTER(a, b, c, d) ((((a + b + 1) >> 1) + ((c + d + 1) >> 1) + 1) >> 1)
    intptr_t o = (intptr_t)y * ds + x;
    d0[o] = (T)LOWRES_FILTER(a0, b0, a1, b1);
    dh[o] = (T)LOWRES_FILTER(a1, b1, a2, b2);
    dv[o] = (T)LOWRES_FILTER(b0, c0, b1, c1);
    dc[o] = (T)LOWRES_FILTER(b1, c1, b2, c2);
#undef LOWRES_FILTER
}

template<int OP, typename TA, typename TB, typename TD>
static void launch_bop(const void* A, intptr_t sa, const int32_t* offA, const void* B, intptr_t sb, const int32_t* offB,
                       void* D, intptr_t sd, const int32_t* offD, int n, int w, int h, int maxv, int shift, int offset, cudaStream_t st)
{
    if (OP != BOP_ADDAVG && !(w & 7) && !(h & 3) && !((sa | sb) & 3))
        blockop_wide_kernel<OP, TA, TB, TD><<<ceil_div((long long)n * (w >> 3) * (h >> 2), 256), 256, 0, st>>>(
            (const TA*)A, sa, offA, (const TB*)B, sb, offB, (TD*)D, sd, offD, n, w, h, maxv);
    else if (!(w & 3) && !(h & 1))
        blockop_quad_kernel<OP, TA, TB, TD><<<ceil_div((long long)n * (w >> 2) * (h >> 1), 256), 256, 0, st>>>(
            (const TA*)A, sa, offA, (const TB*)B, sb, offB, (TD*)D, sd, offD, n, w, h, maxv, shift, offset);
    else
        blockop_kernel<OP, TA, TB, TD><<<ceil_div((long long)n * w * h, 256), 256, 0, st>>>(
            (const TA*)A, sa, offA, (const TB*)B, sb, offB, (TD*)D, sd, offD, n, w, h, maxv, shift, offset);
}

} // namespace b200

using namespace b200;

extern "C" int x265b200_blockop_batch(x265b200_ctx* ctx, int op, int w, int h, const void* A, intptr_t sa, const int32_t* offA,
                                      const void* B, intptr_t sb, const int32_t* offB, void* D, intptr_t sd, const int32_t* offD,
                                      int n, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (w < 1 || h < 1 || w > 64 || h > 64 || n < 0) return fail(ctx, X265B200_ERR_ARG, "blockop: bad shape");
    if (n == 0) return X265B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int maxv = (1 << ctx->depth) - 1;
    const int shift = 14 + 1 - ctx->depth, offset = (1 << (shift - 1)) + 2 * 8192;     // addAvg, pixel.cpp:839-840
    const bool p8 = ctx->pixbytes == 1;
#define BOP(OP, TA8, TB8, TD8, TA16, TB16, TD16)                                                                              \
    do { if (p8) launch_bop<OP, TA8, TB8, TD8>(A, sa, offA, B, sb, offB, D, sd, offD, n, w, h, maxv, shift, offset, st);      \
         else launch_bop<OP, TA16, TB16, TD16>(A, sa, offA, B, sb, offB, D, sd, offD, n, w, h, maxv, shift, offset, st); } while (0)
    switch (op)
    {
    case X265B200_BOP_SUB_PS:   BOP(BOP_SUB_PS, uint8_t, uint8_t, int16_t, uint16_t, uint16_t, int16_t); break;
    case X265B200_BOP_ADD_PS:   BOP(BOP_ADD_PS, uint8_t, int16_t, uint8_t, uint16_t, int16_t, uint16_t); break;
    case X265B200_BOP_PIXELAVG: BOP(BOP_PIXELAVG, uint8_t, uint8_t, uint8_t, uint16_t, uint16_t, uint16_t); break;
    case X265B200_BOP_ADDAVG:   BOP(BOP_ADDAVG, int16_t, int16_t, uint8_t, int16_t, int16_t, uint16_t); break;
    default: return fail(ctx, X265B200_ERR_ARG, "blockop: unknown op");
    }
#undef BOP
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

extern "C" int x265b200_lowres_batch(x265b200_ctx* ctx, const void* src, intptr_t srcStride, void* dst0, void* dsth, void* dstv, void* dstc,
                                     intptr_t dstStride, int width, int height, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (width < 1 || height < 1) return fail(ctx, X265B200_ERR_ARG, "lowres: bad size");
    dim3 grid((unsigned)ceil_div(width, 32), (unsigned)ceil_div(height, 8));
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->pixbytes == 1)
        lowres_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)src, srcStride, (uint8_t*)dst0, (uint8_t*)dsth, (uint8_t*)dstv, (uint8_t*)dstc, dstStride, width, height);
    else
        lowres_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)src, srcStride, (uint16_t*)dst0, (uint16_t*)dsth, (uint16_t*)dstv, (uint16_t*)dstc, dstStride, width, height);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

extern "C" int x265b200_blockcopy_batch(x265b200_ctx* ctx, int kind, int w, int h, const void* src, intptr_t ss, const int32_t* offS,
                                        void* dst, intptr_t sd, const int32_t* offD, int n, int param, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (w < 1 || h < 1 || w > 64 || h > 64 || n < 0 || kind < 0 || kind > 6 || ((kind == 5 || kind == 6) && (param < (kind == 6) || param > 15)))
        return fail(ctx, X265B200_ERR_ARG, "blockcopy: bad shape / kind");
    if (n == 0) return X265B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int grid = ceil_div((long long)n * w * h, 256);
    const bool p8 = ctx->pixbytes == 1;
#define BC(TS, TD) blockcopy_kernel<TS, TD><<<grid, 256, 0, st>>>(kind, (const TS*)src, ss, offS, (TD*)dst, sd, offD, n, w, h, param)
    switch (kind)
    {
    case 0: if (p8) BC(uint8_t, uint8_t); else BC(uint16_t, uint16_t); break;
    case 2: if (p8) BC(int16_t, uint8_t); else BC(int16_t, uint16_t); break;
    case 3: if (p8) BC(uint8_t, int16_t); else BC(uint16_t, int16_t); break;
    default: BC(int16_t, int16_t); break;
    }
#undef BC
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}
