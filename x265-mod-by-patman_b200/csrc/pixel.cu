// pixel.cu -- block-compare metrics: SAD, SATD, sa8d, SSE, ssd_s, multi-candidate SAD, ADS, residual.
//
// Restates (as CUDA, bit-exact) the reference's pixel.cpp:40-383; see include/x265b200.h for the
// slot each entry replaces.  All arithmetic is int32/int64 integer math.
//
// Work decomposition (generic kernels): a block of w x h samples is cut into 4x4 tiles (8x8 for
// sa8d); one thread owns one tile at a time and holds it entirely in registers, so the Hadamard
// transform needs no shuffles; a group of G = min(32, pow2ceil(#tiles)) lanes shares one block and
// reduces with xor-shuffles.  Adjacent lanes take horizontally adjacent tiles, so a warp-wide
// load covers whole 128-byte lines of the plane.
#include <stdio.h>
#include <stdlib.h>
#include "internal.h"
#include "device_util.cuh"
#include "tile_kernels.cuh"

namespace b200 {


__device__ __forceinline__ int hadamard4x4_abs(int (&d)[4][4])
{
#pragma unroll
    for (int y = 0; y < 4; y++)
    {
        int s0 = d[y][0] + d[y][1], s1 = d[y][0] - d[y][1];
        int s2 = d[y][2] + d[y][3], s3 = d[y][2] - d[y][3];
        d[y][0] = s0 + s2; d[y][1] = s1 + s3; d[y][2] = s0 - s2; d[y][3] = s1 - s3;
    }
    int sum = 0;
#pragma unroll
    for (int x = 0; x < 4; x++)
    {
        int s0 = d[0][x] + d[1][x], s1 = d[0][x] - d[1][x];
        int s2 = d[2][x] + d[3][x], s3 = d[2][x] - d[3][x];
        sum += abs(s0 + s2) + abs(s1 + s3) + abs(s0 - s2) + abs(s1 - s3);
    }
    return sum;
}

__device__ __forceinline__ int hadamard8x8_abs(int (&m)[8][8])
{
#pragma unroll
    for (int y = 0; y < 8; y++)
    {
#pragma unroll
        for (int step = 1; step < 8; step <<= 1)
#pragma unroll
            for (int i = 0; i < 8; i += step << 1)
#pragma unroll
                for (int j = i; j < i + step; j++)
                {
                    int u = m[y][j], v = m[y][j + step];
                    m[y][j] = u + v; m[y][j + step] = u - v;
                }
    }
    int sum = 0;
#pragma unroll
    for (int x = 0; x < 8; x++)
    {
#pragma unroll
        for (int step = 1; step < 8; step <<= 1)
#pragma unroll
            for (int i = 0; i < 8; i += step << 1)
#pragma unroll
                for (int j = i; j < i + step; j++)
                {
                    int u = m[j][x], v = m[j + step][x];
                    m[j][x] = u + v; m[j + step][x] = u - v;
                }
#pragma unroll
        for (int y = 0; y < 8; y++) sum += abs(m[y][x]);
    }
    return sum;
}

// One kernel for SAD / SATD / SSE / SSD over 4x4 tiles.  kdiv > 1: offA is indexed by block / kdiv
// (several candidates share one fenc block: sad_x3 / sad_x4).
template<typename T, int OP, typename ACC, typename OUT>
__global__ void __launch_bounds__(256)
tile4_kernel(const T* __restrict__ A, intptr_t sa, const T* __restrict__ B, intptr_t sb,
             const int32_t* __restrict__ offA, const int32_t* __restrict__ offB, int kdiv,
             int n, int w, int h, int G, OUT* __restrict__ out)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int blk = (int)(gid / G);
    int l = (int)(gid % G);
    bool live = blk < n;
    int tw = w >> 2;
    int T4 = tw * (h >> 2);
    const T* a = A;
    const T* b = B;
    if (live)
    {
        a += offA[kdiv > 1 ? blk / kdiv : blk];
        if (OP != OP_SSD) b += offB[blk];
    }
    ACC acc = 0;
    for (int t = l; t < T4; t += G)
    {
        if (!live) break;
        int tx = (t % tw) << 2, ty = (t / tw) << 2;
        const T* pa = a + (intptr_t)ty * sa + tx;
        const T* pb = b + (intptr_t)ty * sb + tx;
        int d[4][4];
#pragma unroll
        for (int y = 0; y < 4; y++)
        {
            int va[4];
            load4(pa + y * sa, va);
            if (OP == OP_SSD)
            {
#pragma unroll
                for (int x = 0; x < 4; x++) d[y][x] = va[x];
            }
            else
            {
                int vb[4];
                load4(pb + y * sb, vb);
#pragma unroll
                for (int x = 0; x < 4; x++) d[y][x] = va[x] - vb[x];
            }
        }
        if (OP == OP_SAD)
        {
            int s = 0;
#pragma unroll
            for (int y = 0; y < 4; y++)
#pragma unroll
                for (int x = 0; x < 4; x++) s += abs(d[y][x]);
            acc += s;
        }
        else if (OP == OP_SATD)
            acc += hadamard4x4_abs(d) >> 1;        // per-tile halving == the reference's 8x4 pairing (raw sums are even)
        else
        {
            // (tmp * tmp) is an int product accumulated into sse_t (pixel.cpp:171-178, 379)
#pragma unroll
            for (int y = 0; y < 4; y++)
#pragma unroll
                for (int x = 0; x < 4; x++) acc += (ACC)(long long)(int)((unsigned)d[y][x] * (unsigned)d[y][x]);
        }
    }
    acc = group_sum(acc, G);
    if (live && l == 0) out[blk] = (OUT)acc;
}

// sa8d: 8x8 Hadamard tiles.  mode16: tiles are ordered so that lanes 4q..4q+3 hold the four
// quadrants of one 16x16, whose raw sums are added before the single (x + 2) >> 2 rounding
// (pixel.cpp:342-354); otherwise each 8x8 rounds on its own (pixel.cpp:336-340).
template<typename T>
__global__ void __launch_bounds__(128)
sa8d_kernel(const T* __restrict__ A, intptr_t sa, const T* __restrict__ B, intptr_t sb,
            const int32_t* __restrict__ offA, const int32_t* __restrict__ offB,
            int n, int w, int h, int G, int mode16, int32_t* __restrict__ out)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int blk = (int)(gid / G);
    int l = (int)(gid % G);
    bool live = blk < n;
    int T8 = (w >> 3) * (h >> 3);
    const T* a = A;
    const T* b = B;
    if (live) { a += offA[blk]; b += offB[blk]; }
    int iters = (T8 + G - 1) / G;
    int acc = 0;
    for (int k = 0; k < iters; k++)
    {
        int t = l + k * G;
        bool valid = live && t < T8;
        int raw = 0;
        if (valid)
        {
            int x, y;
            if (mode16)
            {
                int b16 = t >> 2, q = t & 3, bw = w >> 4;
                x = ((b16 % bw) << 4) + ((q & 1) << 3);
                y = ((b16 / bw) << 4) + ((q >> 1) << 3);
            }
            else
            {
                int tw = w >> 3;
                x = (t % tw) << 3; y = (t / tw) << 3;
            }
            const T* pa = a + (intptr_t)y * sa + x;
            const T* pb = b + (intptr_t)y * sb + x;
            int m[8][8];
#pragma unroll
            for (int r = 0; r < 8; r++)
            {
                int va[4], vb[4];
                load4(pa + r * sa, va); load4(pb + r * sb, vb);
#pragma unroll
                for (int c = 0; c < 4; c++) m[r][c] = va[c] - vb[c];
                load4(pa + r * sa + 4, va); load4(pb + r * sb + 4, vb);
#pragma unroll
                for (int c = 0; c < 4; c++) m[r][4 + c] = va[c] - vb[c];
            }
            raw = hadamard8x8_abs(m);
        }
        if (mode16)
        {
            raw += __shfl_xor_sync(0xffffffffu, raw, 1);
            raw += __shfl_xor_sync(0xffffffffu, raw, 2);
            if (valid && (t & 3) == 0) acc += (raw + 2) >> 2;
        }
        else if (valid)
            acc += (raw + 2) >> 2;
    }
    acc = group_sum(acc, G);
    if (live && l == 0) out[blk] = acc;
}

// Throughput sa8d for planes with strides that are multiples of 4 samples: thread = one 8x8 tile, loaded as two 8x4
// strips of grouped chunk loads.  Differences and the vertical 8-point Hadamard run on packed 16-bit lanes
// (|value| <= 8 * 4095 < 2^15).  PKH (depth <= 10): the two horizontal stages that pair different words (distance 4
// and 2) stay packed as well (8 * 1023 * 4 < 2^15), and the last stage, which pairs the two lanes of a word, is folded
// into the sum of magnitudes: |a + b| + |a - b| = 2 max(|a|, |b|).  Otherwise the horizontal pass runs in int32.
template<typename T, bool PKH, int MINB = 1, bool V16 = false>
__global__ void __launch_bounds__(128, MINB)
sa8d_fast_kernel(const T* __restrict__ A, intptr_t sa, const T* __restrict__ B, intptr_t sb,
                 const int32_t* __restrict__ offA, const int32_t* __restrict__ offB,
                 int n, int w, int h, int G, int mode16, int32_t* __restrict__ out)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int lg = __ffs(G) - 1;
    int blk = (int)(gid >> lg);
    int l = (int)gid & (G - 1);
    bool live = blk < n;
    int T8 = (w >> 3) * (h >> 3);
    const T* a = A;
    const T* b = B;
    if (live) { a += offA[blk]; b += offB[blk]; }
    int iters = (T8 + G - 1) / G;
    int acc = 0;
    for (int k = 0; k < iters; k++)
    {
        int t = l + k * G;
        bool valid = live && t < T8;
        int raw = 0;
        if (valid)
        {
            int x, y;
            if (mode16)
            {
                int b16 = t >> 2, q = t & 3, bw = w >> 4;
                x = ((b16 % bw) << 4) + ((q & 1) << 3);
                y = ((b16 / bw) << 4) + ((q >> 1) << 3);
            }
            else
            {
                int tw = w >> 3;
                x = (t % tw) << 3; y = (t / tw) << 3;
            }
            const T* pa = a + (intptr_t)y * sa + x;
            const T* pb = b + (intptr_t)y * sb + x;
            uint32_t d[8][4];
#pragma unroll
            for (int half = 0; half < 2; half++)
            {
                uint32_t wa[4][4], wb[4][4];
                if (V16 && sizeof(T) == 2)
                {   // 16-byte chunk loads (strides % 8 == 0): a row of the tile is one or two requests instead of two or three
                    load_rows8_v16((const uint16_t*)pa + (intptr_t)(half * 4) * sa, sa, wa);
                    load_rows8_v16((const uint16_t*)pb + (intptr_t)(half * 4) * sb, sb, wb);
                }
                else
                {
                    load_rows8<4>(pa + (intptr_t)(half * 4) * sa, sa, wa);
                    load_rows8<4>(pb + (intptr_t)(half * 4) * sb, sb, wb);
                }
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 4; c++) d[half * 4 + r][c] = wa[r][c] - wb[r][c];
            }
            // vertical 8-point Hadamard, packed
#pragma unroll
            for (int c = 0; c < 4; c++)
            {
#pragma unroll
                for (int step = 1; step < 8; step <<= 1)
#pragma unroll
                    for (int i = 0; i < 8; i += step << 1)
#pragma unroll
                        for (int j = i; j < i + step; j++)
                        {
                            uint32_t u = d[j][c], v = d[j + step][c];
                            d[j][c] = u + v; d[j + step][c] = u - v;
                        }
            }
            if (PKH)
            {
                int half = 0;
#pragma unroll
                for (int r = 0; r < 8; r++)
                {
                    // columns (0,1) (2,3) (4,5) (6,7): distance-4 pairs words 0/2 and 1/3, distance-2 pairs 0/1 and 2/3
                    uint32_t e0 = d[r][0] + d[r][2], e1 = d[r][1] + d[r][3], e2 = d[r][0] - d[r][2], e3 = d[r][1] - d[r][3];
                    uint32_t f[4] = { e0 + e1, e0 - e1, e2 + e3, e2 - e3 };
#pragma unroll
                    for (int c = 0; c < 4; c++)
                    {
                        int p, q;
                        unpack_s16x2(f[c], p, q);
                        half += max(abs(p), abs(q));
                    }
                }
                raw = half << 1;
            }
            else
            {
#pragma unroll
                for (int r = 0; r < 8; r++)
                {
                    int m[8];
#pragma unroll
                    for (int c = 0; c < 4; c++) unpack_s16x2(d[r][c], m[2 * c], m[2 * c + 1]);
#pragma unroll
                    for (int step = 1; step < 8; step <<= 1)
#pragma unroll
                        for (int i = 0; i < 8; i += step << 1)
#pragma unroll
                            for (int j = i; j < i + step; j++)
                            {
                                int u = m[j], v = m[j + step];
                                m[j] = u + v; m[j + step] = u - v;
                            }
#pragma unroll
                    for (int c = 0; c < 8; c++) raw += abs(m[c]);
                }
            }
        }
        if (mode16)
        {
            raw += __shfl_xor_sync(0xffffffffu, raw, 1);
            raw += __shfl_xor_sync(0xffffffffu, raw, 2);
            if (valid && (t & 3) == 0) acc += (raw + 2) >> 2;
        }
        else if (valid)
            acc += (raw + 2) >> 2;
    }
    acc = group_sum(acc, G);
    if (live && l == 0) out[blk] = acc;
}

// residual = A - B for planes with strides % 4 == 0: one thread per 4 samples, chunk loads
template<typename T>
__global__ void __launch_bounds__(256)
residual_fast_kernel(const T* __restrict__ A, intptr_t sa, const T* __restrict__ B, intptr_t sb,
                     const int32_t* __restrict__ offA, const int32_t* __restrict__ offB,
                     int n, int w, int h, int16_t* __restrict__ dst)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int qw = w >> 2;
    int per = qw * h;
    int blk = (int)(gid / per);
    if (blk >= n) return;
    int r = (int)(gid % per);
    int y = r / qw, x = (r % qw) << 2;
    uint32_t wa[2], wb[2];
    load_row_quads<1>(A + offA[blk] + (intptr_t)y * sa + x, wa);
    load_row_quads<1>(B + offB[blk] + (intptr_t)y * sb + x, wb);
    int d0, d1, d2, d3;
    unpack_s16x2(wa[0] - wb[0], d0, d1);
    unpack_s16x2(wa[1] - wb[1], d2, d3);
    *(uint2*)(dst + (size_t)blk * w * h + y * w + x) = make_uint2((uint32_t)(d0 & 0xffff) | ((uint32_t)d1 << 16), (uint32_t)(d2 & 0xffff) | ((uint32_t)d3 << 16));
}

// residual for blocks whose width is a multiple of 8 and height a multiple of 4: one thread per 8x4 strip,
// 24 chunk loads issued up front, lane-wise packed subtraction, 16-byte stores into the block-contiguous output
template<typename T>
__global__ void __launch_bounds__(256)
residual_wide_kernel(const T* __restrict__ A, intptr_t sa, const T* __restrict__ B, intptr_t sb,
                     const int32_t* __restrict__ offA, const int32_t* __restrict__ offB,
                     int n, int w, int h, int16_t* __restrict__ dst)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int sw = w >> 3;
    int per = sw * (h >> 2);
    int blk = (int)(gid / per);
    if (blk >= n) return;
    int r = (int)(gid - (long long)blk * per);
    int y = (r / sw) << 2, x = (r % sw) << 3;
    uint32_t wa[4][4], wb[4][4];
    load_rows8<4>(A + offA[blk] + (intptr_t)y * sa + x, sa, wa);
    load_rows8<4>(B + offB[blk] + (intptr_t)y * sb + x, sb, wb);
    int16_t* d = dst + (size_t)blk * w * h + y * w + x;
#pragma unroll
    for (int i = 0; i < 4; i++)
        *(uint4*)(d + i * w) = make_uint4(psub16(wa[i][0], wb[i][0]), psub16(wa[i][1], wb[i][1]), psub16(wa[i][2], wb[i][2]), psub16(wa[i][3], wb[i][3]));
}

// ADS (pixel.cpp:121-165): one warp per job; ordered compaction with ballot + popc.
__global__ void __launch_bounds__(128)
ads_kernel(int terms, int half, const int32_t* __restrict__ encDC, const uint32_t* __restrict__ sums,
           const int32_t* __restrict__ sumOff, const int32_t* __restrict__ delta,
           const uint16_t* __restrict__ costMvX, const int32_t* __restrict__ costOff,
           const int32_t* __restrict__ width, const int32_t* __restrict__ thresh, int n,
           int16_t* __restrict__ mvs, int mvsPitch, int32_t* __restrict__ count)
{
    int job = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (job >= n) return;
    const uint32_t* s = sums + sumOff[job];
    const uint16_t* cost = costMvX + costOff[job];
    int dl = delta[job], wd = width[job], th = thresh[job];
    long long e0 = encDC[job * 4], e1 = encDC[job * 4 + 1], e2 = encDC[job * 4 + 2], e3 = encDC[job * 4 + 3];
    int16_t* o = mvs + (size_t)job * mvsPitch;
    int nmv = 0;
    for (int base = 0; base < wd; base += 32)
    {
        int i = base + lane;
        bool hit = false;
        if (i < wd)
        {
            long long ads;
            if (terms == 4)
                ads = llabs(e0 - (long long)s[i]) + llabs(e1 - (long long)s[i + half])
                    + llabs(e2 - (long long)s[i + dl]) + llabs(e3 - (long long)s[i + dl + half]);
            else if (terms == 2)
                ads = llabs(e0 - (long long)s[i]) + llabs(e1 - (long long)s[i + dl]);
            else
                ads = llabs(e0 - (long long)s[i]);
            hit = (int)(ads + cost[i]) < th;
        }
        unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) o[nmv + __popc(m & ((1u << lane) - 1))] = (int16_t)i;
        nmv += __popc(m);
    }
    if (lane == 0) count[job] = nmv;
}

// residual = A - B, dst contiguous per block (sub_ps semantics); one thread per 4 samples
template<typename T>
__global__ void __launch_bounds__(256)
residual_kernel(const T* __restrict__ A, intptr_t sa, const T* __restrict__ B, intptr_t sb,
                const int32_t* __restrict__ offA, const int32_t* __restrict__ offB,
                int n, int w, int h, int16_t* __restrict__ dst)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int qw = w >> 2;
    int per = qw * h;
    int blk = (int)(gid / per);
    if (blk >= n) return;
    int r = (int)(gid % per);
    int y = r / qw, x = (r % qw) << 2;
    int va[4], vb[4];
    load4(A + offA[blk] + (intptr_t)y * sa + x, va);
    load4(B + offB[blk] + (intptr_t)y * sb + x, vb);
    short4 o = make_short4((short)(va[0] - vb[0]), (short)(va[1] - vb[1]), (short)(va[2] - vb[2]), (short)(va[3] - vb[3]));
    *(short4*)(dst + (size_t)blk * w * h + y * w + x) = o;
}

// ---------------------------------------------------------------- host-side launchers

static int group_lanes(int tiles) { int g = pow2_ceil(tiles); return g > 32 ? 32 : g; }

// lanes per block for the throughput kernels, tuned with tools/satd_lab.cu on B200: about four 4x4 tiles
// per lane (two for blocks of fewer than 16 tiles) amortises the per-thread setup and the shuffle
// reduction without starving the memory pipeline; 64x64 keeps a full warp.
static int fast_group_lanes(int tiles)
{
    int per = tiles >= 16 ? 4 : 2;
    int g = 1;
    while (g * 2 * per <= tiles && g < 32) g <<= 1;
    return g;
}

template<typename T>
static int launch_pixelcmp(x265b200_ctx* ctx, int op, int w, int h, const T* A, intptr_t sa, const T* B, intptr_t sb,
                           const int32_t* offA, const int32_t* offB, int kdiv, int n, void* out, cudaStream_t st)
{
    if (n <= 0) return X265B200_OK;
    if (op == X265B200_SA8D && !((w | h) & 7))
    {
        int mode16 = !((w | h) & 15);
        int G = group_lanes((w >> 3) * (h >> 3));
        long long threads = (long long)n * G;
        if (!((sa | sb) & 3))
        {
            // lanes per block and register cap, swept on B200 (profiles/r3_all_primitives_10bit.md): four 8x8 tiles per lane for 64x64 (0.2016 ->
            // 0.1910 ms per 32 frames of 2160p10), two for 32x32 with the kernel capped at 64 registers = 8 CTAs per SM (0.2463 -> 0.2111 ms);
            // the smaller CUs keep one tile per lane and gain from the cap alone (16x16 0.2657 -> 0.2439, 8x8 0.2934 -> 0.2702 ms)
            const int tiles = (w >> 3) * (h >> 3);
            const int tpl = tiles >= 64 ? 4 : tiles >= 16 ? 2 : 1;
            G = 1;
            while (G * 2 * tpl <= tiles && G < 32) G <<= 1;
            if (mode16 && G < 4) G = 4;
            threads = (long long)n * G;
            if (ctx->depth <= 10 && tiles <= 4 && sizeof(T) == 2 && !((sa | sb) & 7))      // 16 / 8 wide CUs: 16-byte chunk loads, 0.2449 -> 0.2350 / 0.2722 -> 0.2654 ms
                sa8d_fast_kernel<T, true, 8, true><<<ceil_div(threads, 128), 128, 0, st>>>(A, sa, B, sb, offA, offB, n, w, h, G, mode16, (int32_t*)out);
            else if (ctx->depth <= 10 && tiles < 64)
                sa8d_fast_kernel<T, true, 8><<<ceil_div(threads, 128), 128, 0, st>>>(A, sa, B, sb, offA, offB, n, w, h, G, mode16, (int32_t*)out);
            else if (ctx->depth <= 10)
                sa8d_fast_kernel<T, true><<<ceil_div(threads, 128), 128, 0, st>>>(A, sa, B, sb, offA, offB, n, w, h, G, mode16, (int32_t*)out);
            else
                sa8d_fast_kernel<T, false><<<ceil_div(threads, 128), 128, 0, st>>>(A, sa, B, sb, offA, offB, n, w, h, G, mode16, (int32_t*)out);
            B200_LAUNCH_CHECK(ctx);
            return X265B200_OK;
        }
        sa8d_kernel<T><<<ceil_div(threads, 128), 128, 0, st>>>(A, sa, B, sb, offA, offB, n, w, h, G, mode16, (int32_t*)out);
        B200_LAUNCH_CHECK(ctx);
        return X265B200_OK;
    }
    int G = group_lanes((w >> 2) * (h >> 2));
    long long threads = (long long)n * G;
    int grid = ceil_div(threads, 256);
    if (sizeof(T) == 2 && !((sa | sb) & 7) && ((w == 8 && h == 4) || (w == 16 && h == 8)) && op != X265B200_SA8D)
    {
        // the narrowest 8-wide shapes: 8x4 strips through 16-byte chunk loads (tile_kernels.cuh strip8_fast_kernel):
        // 0.30 -> 0.25 ms for 8x4 and 0.209 -> 0.201 ms for 16x8 at 32 frames of 2160p10 (tools/satd_lab)
        int Gs = w == 8 ? 1 : 2;
        int sgrid = ceil_div((long long)n * Gs, 128);
        const uint16_t* A16 = (const uint16_t*)A;
        const uint16_t* B16 = (const uint16_t*)B;
        if (op == X265B200_SAD) strip8_fast_kernel<OP_SAD, int, int32_t><<<sgrid, 128, 0, st>>>(A16, sa, B16, sb, offA, offB, kdiv, n, w, h, Gs, (int32_t*)out);
        else if (op == X265B200_SATD) strip8_fast_kernel<OP_SATD, int, int32_t><<<sgrid, 128, 0, st>>>(A16, sa, B16, sb, offA, offB, kdiv, n, w, h, Gs, (int32_t*)out);
        else if (op == X265B200_SSE_PP) strip8_fast_kernel<OP_SSE, unsigned long long, unsigned long long><<<sgrid, 128, 0, st>>>(A16, sa, B16, sb, offA, offB, kdiv, n, w, h, Gs, (unsigned long long*)out);
        else return fail(ctx, X265B200_ERR_ARG, "pixelcmp: unknown op");
        B200_LAUNCH_CHECK(ctx);
        return X265B200_OK;
    }
    if (!((sa | sb) & 3))
    {
        G = fast_group_lanes((w >> 2) * (h >> 2));
        grid = ceil_div((long long)n * G, 128);
        // plane strides are multiples of 4 samples (always true for x265 planes): throughput kernels
        switch (op)
        {
        case X265B200_SAD:
            tile4_fast_kernel<T, OP_SAD, int, int32_t, FAST_UNROLL, FAST_MINBLK><<<grid, 128, 0, st>>>(A, sa, B, sb, offA, offB, kdiv, n, w, h, G, (int32_t*)out);
            break;
        case X265B200_SATD:
        case X265B200_SA8D:
            tile4_fast_kernel<T, OP_SATD, int, int32_t, FAST_UNROLL, FAST_MINBLK><<<grid, 128, 0, st>>>(A, sa, B, sb, offA, offB, kdiv, n, w, h, G, (int32_t*)out);
            break;
        case X265B200_SSE_PP:
            tile4_fast_kernel<T, OP_SSE, unsigned long long, unsigned long long, FAST_UNROLL, FAST_MINBLK><<<grid, 128, 0, st>>>(A, sa, B, sb, offA, offB, kdiv, n, w, h, G, (unsigned long long*)out);
            break;
        default:
            return fail(ctx, X265B200_ERR_ARG, "pixelcmp: unknown op");
        }
        B200_LAUNCH_CHECK(ctx);
        return X265B200_OK;
    }
    switch (op)
    {
    case X265B200_SAD:
        tile4_kernel<T, OP_SAD, int, int32_t><<<grid, 256, 0, st>>>(A, sa, B, sb, offA, offB, kdiv, n, w, h, G, (int32_t*)out);
        break;
    case X265B200_SATD:
    case X265B200_SA8D:       // sizes that are not multiples of 8 bind to satd in the reference tables
        tile4_kernel<T, OP_SATD, int, int32_t><<<grid, 256, 0, st>>>(A, sa, B, sb, offA, offB, kdiv, n, w, h, G, (int32_t*)out);
        break;
    case X265B200_SSE_PP:
        tile4_kernel<T, OP_SSE, unsigned long long, unsigned long long><<<grid, 256, 0, st>>>(A, sa, B, sb, offA, offB, kdiv, n, w, h, G, (unsigned long long*)out);
        break;
    default:
        return fail(ctx, X265B200_ERR_ARG, "pixelcmp: unknown op");
    }
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

} // namespace b200

using namespace b200;

static bool shape_ok(int w, int h) { return w >= 4 && h >= 4 && !(w & 3) && !(h & 3) && w <= 64 && h <= 64; }

extern "C" int x265b200_pixelcmp_batch(x265b200_ctx* ctx, int op, int w, int h, const void* A, intptr_t sa,
                                       const void* B, intptr_t sb, const int32_t* offA, const int32_t* offB,
                                       int n, void* out, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (!shape_ok(w, h) || n < 0) return fail(ctx, X265B200_ERR_ARG, "pixelcmp: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    return ctx->pixbytes == 1 ? launch_pixelcmp<uint8_t>(ctx, op, w, h, (const uint8_t*)A, sa, (const uint8_t*)B, sb, offA, offB, 1, n, out, st)
                              : launch_pixelcmp<uint16_t>(ctx, op, w, h, (const uint16_t*)A, sa, (const uint16_t*)B, sb, offA, offB, 1, n, out, st);
}

template<typename T>
static int launch_cu_satd(x265b200_ctx* ctx, int S, const T* A, intptr_t sa, const T* B, intptr_t sb, const int32_t* offF, const int32_t* offR,
                          int n, int32_t* out, cudaStream_t st)
{
    // lanes per CU, from tools/cu_satd_sweep.py on B200 (2160p10 x 32 frames, ms per launch): 8x8: 1 / 2 / 4 lanes 0.90 / 0.62 / 0.76;
    // 16x16: 2 / 4 / 8 / 16 lanes 0.50 / 0.41 / 0.47 / 0.61; 32x32: 8 / 16 / 32 lanes 0.39 / 0.39 / 0.43; 64x64: 16 / 32 lanes 0.42 / 0.37.
    // Few lanes with several tiles each win: a lane's three SATDs per tile amortise the five shuffle reductions of the CU.
    int G = S == 8 ? 2 : S == 16 ? 4 : S == 32 ? 8 : 32;
    if (const char* e = getenv("X265B200_CU_LANES_LAB"))       // tuning lab only (tools/cu_satd_sweep.py)
    {
        int g[4] = { 2, 4, 8, 32 };
        sscanf(e, "%d,%d,%d,%d", &g[0], &g[1], &g[2], &g[3]);
        G = g[S == 8 ? 0 : S == 16 ? 1 : S == 32 ? 2 : 3];
    }
    // 32 and 64 wide CUs of 16-bit pictures up to 10 bits: the horizontal Hadamard as f16 tensor-core MMAs (tile_kernels.cuh cu_satd_mma_kernel),
    // bit-identical: 0.3715 -> 0.3446 ms (64 wide, 16 lanes per CU) and 0.3869 -> 0.3753 ms (32 wide) per 32 frames of 2160p10; at 16 and 8 wide the
    // L1 data pipe (89 / 91 % busy), not the ALU, is the limit and the packed-integer kernel stays ahead.  X265B200_LAB="0" turns it off (lab).
    if constexpr (sizeof(T) == 2)                               // 8-bit pictures: measured slower (0.294 vs 0.285 ms at 64 wide), they keep the integer kernel
    {
        if (ctx->depth <= 10 && S >= 32 && G >= 4 && lab_knob(0, 1))
        {
            if (!getenv("X265B200_CU_LANES_LAB")) G = S == 64 ? 16 : 8;
            const int gridM = ceil_div((long long)n * G, 128);
            if (S == 64) cu_satd_mma_kernel<T, 64><<<gridM, 128, 0, st>>>(A, sa, B, sb, offF, offR, n, G, out);
            else cu_satd_mma_kernel<T, 32><<<gridM, 128, 0, st>>>(A, sa, B, sb, offF, offR, n, G, out);
            B200_LAUNCH_CHECK(ctx);
            return X265B200_OK;
        }
    }
    const int grid = ceil_div((long long)n * G, 128);
    switch (S)
    {
    // 8 and 16 wide: six resident CTAs asked for (80 registers instead of 64: more tile loads in flight), 0.6171 -> 0.5821 and 0.4050 -> 0.3971 ms
    case 8:  cu_satd_kernel<T, 8, 6><<<grid, 128, 0, st>>>(A, sa, B, sb, offF, offR, n, G, out); break;
    case 16: cu_satd_kernel<T, 16, 6><<<grid, 128, 0, st>>>(A, sa, B, sb, offF, offR, n, G, out); break;
    case 32: cu_satd_kernel<T, 32><<<grid, 128, 0, st>>>(A, sa, B, sb, offF, offR, n, G, out); break;
    default: cu_satd_kernel<T, 64><<<grid, 128, 0, st>>>(A, sa, B, sb, offF, offR, n, G, out); break;
    }
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

extern "C" int x265b200_cu_satd_batch(x265b200_ctx* ctx, int cuSize, const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR,
                                      const int32_t* offF, const int32_t* offR, int n, int32_t* cost, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if ((cuSize != 8 && cuSize != 16 && cuSize != 32 && cuSize != 64) || n < 0) return fail(ctx, X265B200_ERR_ARG, "cu_satd: CU size must be 8, 16, 32 or 64");
    if ((strideF | strideR) & 3) return fail(ctx, X265B200_ERR_ARG, "cu_satd: plane strides must be multiples of 4 samples");
    if (n == 0) return X265B200_OK;
    if (!fenc || !ref || !offF || !offR || !cost) return fail(ctx, X265B200_ERR_ARG, "cu_satd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    return ctx->pixbytes == 1 ? launch_cu_satd<uint8_t>(ctx, cuSize, (const uint8_t*)fenc, strideF, (const uint8_t*)ref, strideR, offF, offR, n, cost, st)
                              : launch_cu_satd<uint16_t>(ctx, cuSize, (const uint16_t*)fenc, strideF, (const uint16_t*)ref, strideR, offF, offR, n, cost, st);
}

extern "C" int x265b200_sad_multi_batch(x265b200_ctx* ctx, int w, int h, const void* fenc, intptr_t sf,
                                        const void* ref, intptr_t sr, const int32_t* offF, const int32_t* offR,
                                        int K, int n, int32_t* out, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (!shape_ok(w, h) || n < 0 || K < 1) return fail(ctx, X265B200_ERR_ARG, "sad_multi: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    return ctx->pixbytes == 1 ? launch_pixelcmp<uint8_t>(ctx, X265B200_SAD, w, h, (const uint8_t*)fenc, sf, (const uint8_t*)ref, sr, offF, offR, K, n * K, out, st)
                              : launch_pixelcmp<uint16_t>(ctx, X265B200_SAD, w, h, (const uint16_t*)fenc, sf, (const uint16_t*)ref, sr, offF, offR, K, n * K, out, st);
}

extern "C" int x265b200_sse_ss_batch(x265b200_ctx* ctx, int w, int h, const int16_t* A, intptr_t sa, const int16_t* B,
                                     intptr_t sb, const int32_t* offA, const int32_t* offB, int n, uint64_t* out,
                                     x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (!shape_ok(w, h) || n < 0) return fail(ctx, X265B200_ERR_ARG, "sse_ss: bad shape");
    if (n == 0) return X265B200_OK;
    int G = group_lanes((w >> 2) * (h >> 2));
    tile4_kernel<int16_t, OP_SSE, unsigned long long, unsigned long long>
        <<<ceil_div((long long)n * G, 256), 256, 0, (cudaStream_t)stream>>>(A, sa, B, sb, offA, offB, 1, n, w, h, G, (unsigned long long*)out);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

extern "C" int x265b200_ssd_s_batch(x265b200_ctx* ctx, int size, const int16_t* A, intptr_t sa, const int32_t* offA,
                                    int n, uint64_t* out, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (!shape_ok(size, size) || n < 0) return fail(ctx, X265B200_ERR_ARG, "ssd_s: bad shape");
    if (n == 0) return X265B200_OK;
    int G = group_lanes((size >> 2) * (size >> 2));
    tile4_kernel<int16_t, OP_SSD, unsigned long long, unsigned long long>
        <<<ceil_div((long long)n * G, 256), 256, 0, (cudaStream_t)stream>>>(A, sa, A, sa, offA, offA, 1, n, size, size, G, (unsigned long long*)out);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

extern "C" int x265b200_ads_batch(x265b200_ctx* ctx, int terms, int half, const int32_t* encDC, const uint32_t* sums,
                                  const int32_t* sumOff, const int32_t* delta, const uint16_t* costMvX,
                                  const int32_t* costOff, const int32_t* width, const int32_t* thresh, int n,
                                  int16_t* mvs, int mvsPitch, int32_t* count, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if ((terms != 1 && terms != 2 && terms != 4) || n < 0) return fail(ctx, X265B200_ERR_ARG, "ads: bad terms");
    if (n == 0) return X265B200_OK;
    ads_kernel<<<ceil_div((long long)n * 32, 128), 128, 0, (cudaStream_t)stream>>>(
        terms, half, encDC, sums, sumOff, delta, costMvX, costOff, width, thresh, n, mvs, mvsPitch, count);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

extern "C" int x265b200_residual_batch(x265b200_ctx* ctx, int w, int h, const void* A, intptr_t sa, const void* B,
                                       intptr_t sb, const int32_t* offA, const int32_t* offB, int n, int16_t* dst,
                                       x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (!shape_ok(w, h) || n < 0) return fail(ctx, X265B200_ERR_ARG, "residual: bad shape");
    if (n == 0) return X265B200_OK;
    long long threads = (long long)n * (w >> 2) * h;
    cudaStream_t st = (cudaStream_t)stream;
    if (!((sa | sb) & 3) && !(w & 7) && !((uintptr_t)dst & 15))
    {
        long long strips = (long long)n * (w >> 3) * (h >> 2);
        if (ctx->pixbytes == 1)
            residual_wide_kernel<uint8_t><<<ceil_div(strips, 256), 256, 0, st>>>((const uint8_t*)A, sa, (const uint8_t*)B, sb, offA, offB, n, w, h, dst);
        else
            residual_wide_kernel<uint16_t><<<ceil_div(strips, 256), 256, 0, st>>>((const uint16_t*)A, sa, (const uint16_t*)B, sb, offA, offB, n, w, h, dst);
    }
    else if (!((sa | sb) & 3))
    {
        if (ctx->pixbytes == 1)
            residual_fast_kernel<uint8_t><<<ceil_div(threads, 256), 256, 0, st>>>((const uint8_t*)A, sa, (const uint8_t*)B, sb, offA, offB, n, w, h, dst);
        else
            residual_fast_kernel<uint16_t><<<ceil_div(threads, 256), 256, 0, st>>>((const uint16_t*)A, sa, (const uint16_t*)B, sb, offA, offB, n, w, h, dst);
    }
    else if (ctx->pixbytes == 1)
        residual_kernel<uint8_t><<<ceil_div(threads, 256), 256, 0, st>>>((const uint8_t*)A, sa, (const uint8_t*)B, sb, offA, offB, n, w, h, dst);
    else
        residual_kernel<uint16_t><<<ceil_div(threads, 256), 256, 0, st>>>((const uint16_t*)A, sa, (const uint16_t*)B, sb, offA, offB, n, w, h, dst);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}
