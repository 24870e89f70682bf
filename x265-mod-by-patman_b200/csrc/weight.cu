// weight.cu -- weighted prediction of a plane (weight_pp / weight_sp, reference common/pixel.cpp:485-535) and the
// lookahead's weighted-prediction cost (reference encoder/slicetype.cpp:866-897 LookaheadTLD::weightCostLuma and
// encoder/weightPrediction.cpp:171-222 weightCost, luma branch) fused into one pass:
//     cost[k] = sum over 8x8 blocks of min(satd_8x8(weight_k(ref), fenc), intraCost[block])        (uint32, wraps like the reference)
// for K candidate weights at once.  The reference materialises weight_k(ref) as a plane per candidate and then walks
// it with the 8x8 SATD slot; here a thread keeps the fenc and ref 8x8 block in registers and applies every candidate
// weight to the register copy, so both pictures are read once for all K candidates.
#include "internal.h"
#include "device_util.cuh"
#include "tile_kernels.cuh"

namespace b200 {

struct WeightP { int w0, round, shift, offset; };       // shift < 0: unweighted (the reference then uses ref as it is)

// one pixel: x265_clip((w0 * (int16)(pix << correction) + round >> shift) + offset), pixel.cpp:527-528
__device__ __forceinline__ int weight_pix(int pix, const WeightP& p, int correction, int maxv)
{
    int val = (int)(int16_t)(pix << correction);
    return min(max(((p.w0 * val + p.round) >> p.shift) + p.offset, 0), maxv);
}

template<typename SRC, typename PIX>
__global__ void __launch_bounds__(256)
weight_kernel(const SRC* __restrict__ src, intptr_t ss, PIX* __restrict__ dst, intptr_t ds, int width, int height,
              WeightP p, int correction, int maxv, int sp)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int x = (int)(gid % width), y = (int)(gid / width);
    if (y >= height) return;
    int s = src[(intptr_t)y * ss + x];
    int v;
    if (sp) v = min(max(((p.w0 * (s + 8192) + p.round) >> p.shift) + p.offset, 0), maxv);     // pixel.cpp:501
    else v = weight_pix(s, p, correction, maxv);
    dst[(intptr_t)y * ds + x] = (PIX)v;
}

template<typename PIX>
__global__ void __launch_bounds__(128)
weight_cost_kernel(const PIX* __restrict__ fenc, const PIX* __restrict__ ref, intptr_t stride, int bw, int bh,
                   const int32_t* __restrict__ intraCost, const WeightP* __restrict__ wp, int K, int correction, int maxv,
                   uint32_t* __restrict__ cost)
{
    int blk = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = blk < bw * bh;
    uint32_t flo[4][4], fhi[4][4], rlo[4][4], rhi[4][4];
    int cap = 0x7fffffff;
    if (live)
    {
        int by = blk / bw, bx = blk - by * bw;
        const PIX* f = fenc + (intptr_t)(by * 8) * stride + bx * 8;
        const PIX* r = ref + (intptr_t)(by * 8) * stride + bx * 8;
#pragma unroll
        for (int t = 0; t < 4; t++)
        {
            intptr_t o = (intptr_t)((t >> 1) * 4) * stride + (t & 1) * 4;
            load_tile4x4(f + o, stride, flo[t], fhi[t]);
            load_tile4x4(r + o, stride, rlo[t], rhi[t]);
        }
        if (intraCost) cap = intraCost[blk];
    }
    for (int k = 0; k < K; k++)
    {
        WeightP p = wp[k];
        int satd = 0;
        if (live)
        {
#pragma unroll
            for (int t = 0; t < 4; t++)
            {
                uint32_t wlo[4], whi[4];
#pragma unroll
                for (int rr = 0; rr < 4; rr++)
                {
                    if (p.shift < 0) { wlo[rr] = rlo[t][rr]; whi[rr] = rhi[t][rr]; }
                    else
                    {
                        int a = weight_pix((int)(rlo[t][rr] & 0xffff), p, correction, maxv), b = weight_pix((int)(rlo[t][rr] >> 16), p, correction, maxv);
                        int c = weight_pix((int)(rhi[t][rr] & 0xffff), p, correction, maxv), d = weight_pix((int)(rhi[t][rr] >> 16), p, correction, maxv);
                        wlo[rr] = (uint32_t)a | ((uint32_t)b << 16);
                        whi[rr] = (uint32_t)c | ((uint32_t)d << 16);
                    }
                }
                // the reference calls satd(weighted ref, fenc): the metric is symmetric, the argument order is kept anyway
                tile4_accumulate<OP_SATD, int>(wlo, whi, flo[t], fhi[t], satd);
            }
            satd = min(satd, cap);
        }
        uint32_t s = __reduce_add_sync(0xffffffffu, (uint32_t)satd);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(cost + k, s);
    }
}

} // namespace b200

using namespace b200;

extern "C" int x265b200_weight_batch(x265b200_ctx* ctx, int sp, const void* src, intptr_t srcStride, void* dst, intptr_t dstStride,
                                     int width, int height, int w0, int round, int shift, int offset, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (width < 1 || height < 1 || shift < 0 || shift > 31) return fail(ctx, X265B200_ERR_ARG, "weight: bad size / shift");
    WeightP p = { w0, round, shift, offset };
    const int correction = 14 - ctx->depth, maxv = (1 << ctx->depth) - 1;
    long long total = (long long)width * height;
    cudaStream_t st = (cudaStream_t)stream;
    if (sp)
    {
        if (ctx->pixbytes == 1) weight_kernel<int16_t, uint8_t><<<ceil_div(total, 256), 256, 0, st>>>((const int16_t*)src, srcStride, (uint8_t*)dst, dstStride, width, height, p, correction, maxv, 1);
        else weight_kernel<int16_t, uint16_t><<<ceil_div(total, 256), 256, 0, st>>>((const int16_t*)src, srcStride, (uint16_t*)dst, dstStride, width, height, p, correction, maxv, 1);
    }
    else
    {
        if (ctx->pixbytes == 1) weight_kernel<uint8_t, uint8_t><<<ceil_div(total, 256), 256, 0, st>>>((const uint8_t*)src, srcStride, (uint8_t*)dst, dstStride, width, height, p, correction, maxv, 0);
        else weight_kernel<uint16_t, uint16_t><<<ceil_div(total, 256), 256, 0, st>>>((const uint16_t*)src, srcStride, (uint16_t*)dst, dstStride, width, height, p, correction, maxv, 0);
    }
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

extern "C" int x265b200_weight_cost_batch(x265b200_ctx* ctx, const void* fenc, const void* ref, intptr_t stride, int width, int height,
                                          const int32_t* intraCost, const int32_t* weights, int K, uint32_t* cost, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (width < 1 || height < 1 || K < 0 || (stride & 3)) return fail(ctx, X265B200_ERR_ARG, "weight_cost: bad geometry");
    if (K == 0) return X265B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    B200_CUDA(ctx, cudaMemsetAsync(cost, 0, (size_t)K * sizeof(uint32_t), st));
    int bw = (width + 7) >> 3, bh = (height + 7) >> 3;
    const int correction = 14 - ctx->depth, maxv = (1 << ctx->depth) - 1;
    if (ctx->pixbytes == 1)
        weight_cost_kernel<uint8_t><<<ceil_div((long long)bw * bh, 128), 128, 0, st>>>((const uint8_t*)fenc, (const uint8_t*)ref, stride, bw, bh, intraCost,
                                                                                     (const WeightP*)weights, K, correction, maxv, cost);
    else
        weight_cost_kernel<uint16_t><<<ceil_div((long long)bw * bh, 128), 128, 0, st>>>((const uint16_t*)fenc, (const uint16_t*)ref, stride, bw, bh, intraCost,
                                                                                      (const WeightP*)weights, K, correction, maxv, cost);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}
