// blockstats.cu -- the per-block scalars next to the hot path (SURVEY.md section 8a "adjacent slots"):
//     var            cu[].var            pixel.cpp:695-712   sum | sum of squares << 32 (both uint32, wrapping like the reference)
//     psy_cost_pp    cu[].psy_cost_pp    pixel.cpp:718-749   |AC energy(source) - AC energy(recon)|, energy = sa8d - (sad >> 2) per 8x8
//     count_nonzero  cu[].count_nonzero  dct.cpp:716-728
//     copy_cnt       cu[].copy_cnt       dct.cpp:730-744     strided residual -> contiguous coefficients + non-zero count
//     denoiseDct     denoiseDct          dct.cpp:746-757     |level| accumulated into resSum, offset subtracted, sign restored
// One lane group per block, one 8x8 (4x4) sub-block or a strided share of the samples per lane.
#include "internal.h"
#include "device_util.cuh"

namespace b200 {

__device__ __forceinline__ size_t bs_off(const int32_t* off, int blk, int nn) { return off ? (size_t)off[blk] : (size_t)blk * nn; }

template<typename T>
__global__ void __launch_bounds__(256)
var_kernel(const T* __restrict__ pix, intptr_t stride, const int32_t* __restrict__ off, int n, int size, unsigned long long* __restrict__ out)
{
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const T* p = pix + bs_off(off, warp, size * size);
    uint32_t sum = 0, sqr = 0;
    for (int i = lane; i < size * size; i += 32)
    {
        uint32_t v = p[(intptr_t)(i / size) * stride + (i % size)];
        sum += v; sqr += v * v;
    }
    sum = __reduce_add_sync(0xffffffffu, sum);
    sqr = __reduce_add_sync(0xffffffffu, sqr);
    if (lane == 0) out[warp] = (unsigned long long)sum + ((unsigned long long)sqr << 32);
}

// sum |H8 X H8^T| and sum X of one 8x8 block (the "difference" against the reference's zero buffer is the block itself)
template<typename T>
__device__ __forceinline__ int ac_energy8(const T* p, intptr_t stride)
{
    int m[8][8];
    int sad = 0;
#pragma unroll
    for (int y = 0; y < 8; y++)
#pragma unroll
        for (int x = 0; x < 8; x++) { m[y][x] = p[(intptr_t)y * stride + x]; sad += m[y][x]; }
#pragma unroll
    for (int y = 0; y < 8; y++)
#pragma unroll
        for (int step = 1; step < 8; step <<= 1)
#pragma unroll
            for (int i = 0; i < 8; i += step << 1)
#pragma unroll
                for (int j = i; j < i + step; j++) { int u = m[y][j], v = m[y][j + step]; m[y][j] = u + v; m[y][j + step] = u - v; }
    int raw = 0;
#pragma unroll
    for (int x = 0; x < 8; x++)
    {
#pragma unroll
        for (int step = 1; step < 8; step <<= 1)
#pragma unroll
            for (int i = 0; i < 8; i += step << 1)
#pragma unroll
                for (int j = i; j < i + step; j++) { int u = m[j][x], v = m[j + step][x]; m[j][x] = u + v; m[j + step][x] = u - v; }
#pragma unroll
        for (int y = 0; y < 8; y++) raw += abs(m[y][x]);
    }
    return ((raw + 2) >> 2) - (sad >> 2);               // sa8d_8x8 (pixel.cpp:336-340) minus DC (sad >> 2)
}
template<typename T>
__device__ __forceinline__ int ac_energy4(const T* p, intptr_t stride)
{
    int d[4][4];
    int sad = 0;
#pragma unroll
    for (int y = 0; y < 4; y++)
#pragma unroll
        for (int x = 0; x < 4; x++) { d[y][x] = p[(intptr_t)y * stride + x]; sad += d[y][x]; }
#pragma unroll
    for (int y = 0; y < 4; y++)
    {
        int s0 = d[y][0] + d[y][1], s1 = d[y][0] - d[y][1], s2 = d[y][2] + d[y][3], s3 = d[y][2] - d[y][3];
        d[y][0] = s0 + s2; d[y][1] = s1 + s3; d[y][2] = s0 - s2; d[y][3] = s1 - s3;
    }
    int raw = 0;
#pragma unroll
    for (int x = 0; x < 4; x++)
    {
        int s0 = d[0][x] + d[1][x], s1 = d[0][x] - d[1][x], s2 = d[2][x] + d[3][x], s3 = d[2][x] - d[3][x];
        raw += abs(s0 + s2) + abs(s1 + s3) + abs(s0 - s2) + abs(s1 - s3);
    }
    return (raw >> 1) - (sad >> 2);                      // satd_4x4 minus DC
}

// lanes of a group of G = min(32, (size / 8)^2) share one block; each 8x8 sub-block is one lane's work
template<typename T>
__global__ void __launch_bounds__(128)
psy_cost_kernel(const T* __restrict__ src, intptr_t ss, const int32_t* __restrict__ offS, const T* __restrict__ rec, intptr_t sr,
                const int32_t* __restrict__ offR, int n, int size, int G, int32_t* __restrict__ out)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int lg = __ffs(G) - 1;
    int blk = (int)(gid >> lg), l = (int)gid & (G - 1);
    bool live = blk < n;
    uint32_t tot = 0;
    if (live)
    {
        const T* s = src + bs_off(offS, blk, size * size);
        const T* r = rec + bs_off(offR, blk, size * size);
        if (size == 4) { if (l == 0) tot = (uint32_t)abs(ac_energy4(s, ss) - ac_energy4(r, sr)); }
        else
        {
            int bw = size >> 3;
            for (int t = l; t < bw * bw; t += G)
            {
                int i = (t / bw) << 3, j = (t % bw) << 3;
                tot += (uint32_t)abs(ac_energy8(s + (intptr_t)i * ss + j, ss) - ac_energy8(r + (intptr_t)i * sr + j, sr));
            }
        }
    }
    tot = group_sum(tot, G);
    if (live && l == 0) out[blk] = (int32_t)tot;
}

// count_nonzero (resi == nullptr... no copy) / copy_cnt: warp per block
__global__ void __launch_bounds__(256)
count_copy_kernel(const int16_t* __restrict__ src, intptr_t stride, const int32_t* __restrict__ off, int n, int size,
                  int16_t* __restrict__ coeff, uint32_t* __restrict__ count)
{
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const int16_t* p = src + bs_off(off, warp, size * size);
    uint32_t c = 0;
    for (int i = lane; i < size * size; i += 32)
    {
        int16_t v = p[(intptr_t)(i / size) * stride + (i % size)];
        if (coeff) coeff[(size_t)warp * size * size + i] = v;
        c += v != 0;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) count[warp] = c;
}

__global__ void __launch_bounds__(256)
denoise_kernel(int16_t* __restrict__ dct, uint32_t* __restrict__ resSum, const uint16_t* __restrict__ offset, int numCoeff, long long total)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    int i = (int)(gid % numCoeff);
    int level = dct[gid];
    int sign = level >> 31;
    level = (level + sign) ^ sign;
    if (level) atomicAdd(resSum + i, (uint32_t)level);
    level -= offset[i];
    dct[gid] = (int16_t)(level < 0 ? 0 : (level ^ sign) - sign);
}

} // namespace b200

using namespace b200;

static bool size_ok(int s) { return s == 4 || s == 8 || s == 16 || s == 32 || s == 64; }

extern "C" int x265b200_var_batch(x265b200_ctx* ctx, int size, const void* pix, intptr_t stride, const int32_t* off, int n, uint64_t* out,
                                  x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (!size_ok(size) || n < 0) return fail(ctx, X265B200_ERR_ARG, "var: bad size");
    if (n == 0) return X265B200_OK;
    int grid = ceil_div((long long)n * 32, 256);
    if (ctx->pixbytes == 1) var_kernel<uint8_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)pix, stride, off, n, size, (unsigned long long*)out);
    else var_kernel<uint16_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)pix, stride, off, n, size, (unsigned long long*)out);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

extern "C" int x265b200_psy_cost_batch(x265b200_ctx* ctx, int size, const void* src, intptr_t ss, const int32_t* offS, const void* rec, intptr_t sr,
                                       const int32_t* offR, int n, int32_t* out, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (!size_ok(size) || n < 0) return fail(ctx, X265B200_ERR_ARG, "psy_cost: bad size");
    if (n == 0) return X265B200_OK;
    int sub = size == 4 ? 1 : (size >> 3) * (size >> 3);
    int G = sub < 32 ? sub : 32;
    int grid = ceil_div((long long)n * G, 128);
    if (ctx->pixbytes == 1) psy_cost_kernel<uint8_t><<<grid, 128, 0, (cudaStream_t)stream>>>((const uint8_t*)src, ss, offS, (const uint8_t*)rec, sr, offR, n, size, G, out);
    else psy_cost_kernel<uint16_t><<<grid, 128, 0, (cudaStream_t)stream>>>((const uint16_t*)src, ss, offS, (const uint16_t*)rec, sr, offR, n, size, G, out);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

extern "C" int x265b200_count_nonzero_batch(x265b200_ctx* ctx, int size, const int16_t* src, intptr_t stride, const int32_t* off, int n,
                                            int16_t* coeff, uint32_t* count, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (!size_ok(size) || n < 0) return fail(ctx, X265B200_ERR_ARG, "count_nonzero: bad size");
    if (n == 0) return X265B200_OK;
    count_copy_kernel<<<ceil_div((long long)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(src, stride, off, n, size, coeff, count);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

extern "C" int x265b200_denoise_dct_batch(x265b200_ctx* ctx, int16_t* dct, uint32_t* resSum, const uint16_t* offset, int numCoeff, int n,
                                          x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (numCoeff < 1 || n < 0) return fail(ctx, X265B200_ERR_ARG, "denoise_dct: bad size");
    if (n == 0) return X265B200_OK;
    long long total = (long long)numCoeff * n;
    denoise_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(dct, resSum, offset, numCoeff, total);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}
