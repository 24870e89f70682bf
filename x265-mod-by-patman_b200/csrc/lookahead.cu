// lookahead.cu -- the two cost steps of CostEstimateGroup::estimateCUCost that sit around the lowres motion search
// (reference encoder/slicetype.cpp:4467-4650), batched over the 8x8 CUs of a lowres frame:
//   * predictor selection (:4520-4558): every neighbour vector is costed by the 8x8 SATD of its motion-compensated block
//     (ReferencePlanes::lowresMC, common/lowres.h:74-93: a half-pel plane, or the rounded average of the two nearest planes for a
//     quarter-pel vector) and the cheapest becomes the predictor handed to motionEstimate;
//   * the bi-directional candidates of a B frame (:4577-4596): SATD against the average of both lists' motion-compensated blocks and
//     against the average of the two co-located full-pel blocks.
// Together with x265b200_lowres_motion_estimate_batch, x265b200_lowres_intra_batch and x265b200_weight_cost_batch these are all the
// pixel-touching steps of estimateCUCost; what remains on the host is the list bookkeeping (COPY2_IF_LT over three costs, the AQ-weighted
// sums) and the order in which CUs become available (each CU's candidates are its already-searched neighbours' vectors).
// Four lanes share a CU, one 4x4 tile each; the SATD tile code is the one of the metric kernels (tile_kernels.cuh).
#include "internal.h"
#include "tile_kernels.cuh"

namespace b200 {

// the 4x4 tile (tx, ty) of the block a quarter-pel vector addresses in a lowres reference (four half-pel planes `pitch` apart)
template<typename T>
__device__ __forceinline__ void lowres_mc_tile(const T* planes, intptr_t stride, size_t pitch, int qx, int qy, int tx, int ty, uint32_t (&lo)[4], uint32_t (&hi)[4])
{
    const T* a = planes + (size_t)((qy & 2) | ((qx & 2) >> 1)) * pitch + (qx >> 2) + (intptr_t)(qy >> 2) * stride + (intptr_t)(ty << 2) * stride + (tx << 2);
    load_tile4x4(a, stride, lo, hi);
    if ((qx | qy) & 1)
    {
        const int bx = qx + (qx & 1), by = qy + (qy & 1);
        const T* b = planes + (size_t)((by & 2) | ((bx & 2) >> 1)) * pitch + (bx >> 2) + (intptr_t)(by >> 2) * stride + (intptr_t)(ty << 2) * stride + (tx << 2);
        uint32_t blo[4], bhi[4];
        load_tile4x4(b, stride, blo, bhi);
#pragma unroll
        for (int r = 0; r < 4; r++)
        {   // pixelavg_pp (pixel.cpp:586-594) on packed pairs: samples < 2^15, the halves cannot carry into each other
            lo[r] = ((lo[r] + blo[r] + 0x00010001u) >> 1) & 0x7fff7fffu;
            hi[r] = ((hi[r] + bhi[r] + 0x00010001u) >> 1) & 0x7fff7fffu;
        }
    }
}

__device__ __forceinline__ int quad_sum(int v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

template<typename T>
__global__ void __launch_bounds__(128)
lowres_mvp_kernel(const T* __restrict__ fenc, intptr_t sf, const int32_t* __restrict__ offF, const T* __restrict__ planes, intptr_t sr, size_t pitch,
                  const int32_t* __restrict__ offR, const int32_t* __restrict__ mvc, const int32_t* __restrict__ numc, int bBidir, int n,
                  int32_t* __restrict__ mvp, int32_t* __restrict__ mvpCost, int32_t* __restrict__ skipCost)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int cu = gid >> 2, t = gid & 3, tx = t & 1, ty = t >> 1;
    const bool live = cu < n;
    uint32_t flo[4], fhi[4];
    int nc = 0;
    const T* ref = planes;
    if (live)
    {
        load_tile4x4(fenc + offF[cu] + (intptr_t)(ty << 2) * sf + (tx << 2), sf, flo, fhi);
        nc = numc[cu];
        ref += offR[cu];
    }
    else
    {
#pragma unroll
        for (int r = 0; r < 4; r++) { flo[r] = 0; fhi[r] = 0; }
    }
    int px = 0, py = 0, pcost = 1 << 28, skip = 0x7fffffff;          // MotionEstimate::COST_MAX, INT_MAX
    for (int k = 0; k < 5; k++)
    {   // all four lanes of a CU share nc; lanes of other CUs in the warp simply contribute zeros to their own sums
        const bool on = k < nc;
        int c = 0, qx = 0, qy = 0;
        if (on)
        {
            qx = mvc[((size_t)cu * 5 + k) * 2]; qy = mvc[((size_t)cu * 5 + k) * 2 + 1];
            uint32_t lo[4], hi[4];
            lowres_mc_tile(ref, sr, pitch, qx, qy, tx, ty, lo, hi);
            tile4_accumulate<OP_SATD, int>(flo, fhi, lo, hi, c);
        }
        c = quad_sum(c);
        if (on)
        {
            if (c < pcost) { pcost = c; px = qx; py = qy; }          // COPY2_IF_LT: the first of equal costs stays
            if (!(px | py) && bBidir) skip = c;
        }
    }
    if (live && t == 0) { mvp[2 * cu] = px; mvp[2 * cu + 1] = py; mvpCost[cu] = pcost; skipCost[cu] = skip; }
}

template<typename T>
__global__ void __launch_bounds__(128)
lowres_bidir_kernel(const T* __restrict__ fenc, intptr_t sf, const int32_t* __restrict__ offF, const T* __restrict__ planes0, intptr_t s0, size_t pitch0,
                    const T* __restrict__ planes1, intptr_t s1, size_t pitch1, const int32_t* __restrict__ offR, const int32_t* __restrict__ mv0,
                    const int32_t* __restrict__ mv1, int n, int32_t* __restrict__ cost)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int cu = gid >> 2, t = gid & 3, tx = t & 1, ty = t >> 1;
    const bool live = cu < n;
    int cb = 0, cc = 0;
    if (live)
    {
        uint32_t flo[4], fhi[4], alo[4], ahi[4], blo[4], bhi[4];
        load_tile4x4(fenc + offF[cu] + (intptr_t)(ty << 2) * sf + (tx << 2), sf, flo, fhi);
        const T* r0 = planes0 + offR[cu];
        const T* r1 = planes1 + offR[cu];
        lowres_mc_tile(r0, s0, pitch0, mv0[2 * cu], mv0[2 * cu + 1], tx, ty, alo, ahi);
        lowres_mc_tile(r1, s1, pitch1, mv1[2 * cu], mv1[2 * cu + 1], tx, ty, blo, bhi);
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            alo[r] = ((alo[r] + blo[r] + 0x00010001u) >> 1) & 0x7fff7fffu;
            ahi[r] = ((ahi[r] + bhi[r] + 0x00010001u) >> 1) & 0x7fff7fffu;
        }
        tile4_accumulate<OP_SATD, int>(flo, fhi, alo, ahi, cb);
        // co-located candidate: the two full-pel planes at the CU itself
        load_tile4x4(r0 + (intptr_t)(ty << 2) * s0 + (tx << 2), s0, alo, ahi);
        load_tile4x4(r1 + (intptr_t)(ty << 2) * s1 + (tx << 2), s1, blo, bhi);
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            alo[r] = ((alo[r] + blo[r] + 0x00010001u) >> 1) & 0x7fff7fffu;
            ahi[r] = ((ahi[r] + bhi[r] + 0x00010001u) >> 1) & 0x7fff7fffu;
        }
        tile4_accumulate<OP_SATD, int>(flo, fhi, alo, ahi, cc);
    }
    cb = quad_sum(cb); cc = quad_sum(cc);
    if (live && t == 0) { cost[2 * cu] = cb; cost[2 * cu + 1] = cc; }
}

} // namespace b200

using namespace b200;

extern "C" int x265b200_lowres_mvp_batch(x265b200_ctx* ctx, const void* fenc, intptr_t strideF, const int32_t* offF, const void* planes, intptr_t strideR,
                                         size_t planePitch, const int32_t* offR, const int32_t* mvc, const int32_t* numc, int bBidir, int n,
                                         int32_t* mvp, int32_t* mvpCost, int32_t* skipCost, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (n < 0 || ((strideF | strideR) & 3)) return fail(ctx, X265B200_ERR_ARG, "lowres_mvp: strides must be multiples of 4 samples");
    if (n == 0) return X265B200_OK;
    if (!fenc || !planes || !offF || !offR || !mvc || !numc || !mvp || !mvpCost || !skipCost) return fail(ctx, X265B200_ERR_ARG, "lowres_mvp: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ceil_div((long long)n * 4, 128);
    if (ctx->pixbytes == 1)
        lowres_mvp_kernel<uint8_t><<<grid, 128, 0, st>>>((const uint8_t*)fenc, strideF, offF, (const uint8_t*)planes, strideR, planePitch, offR, mvc, numc, bBidir, n, mvp, mvpCost, skipCost);
    else
        lowres_mvp_kernel<uint16_t><<<grid, 128, 0, st>>>((const uint16_t*)fenc, strideF, offF, (const uint16_t*)planes, strideR, planePitch, offR, mvc, numc, bBidir, n, mvp, mvpCost, skipCost);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

extern "C" int x265b200_lowres_bidir_cost_batch(x265b200_ctx* ctx, const void* fenc, intptr_t strideF, const int32_t* offF,
                                                const void* planes0, intptr_t stride0, size_t planePitch0, const void* planes1, intptr_t stride1, size_t planePitch1,
                                                const int32_t* offR, const int32_t* mv0, const int32_t* mv1, int n, int32_t* cost, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (n < 0 || ((strideF | stride0 | stride1) & 3)) return fail(ctx, X265B200_ERR_ARG, "lowres_bidir: strides must be multiples of 4 samples");
    if (n == 0) return X265B200_OK;
    if (!fenc || !planes0 || !planes1 || !offF || !offR || !mv0 || !mv1 || !cost) return fail(ctx, X265B200_ERR_ARG, "lowres_bidir: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ceil_div((long long)n * 4, 128);
    if (ctx->pixbytes == 1)
        lowres_bidir_kernel<uint8_t><<<grid, 128, 0, st>>>((const uint8_t*)fenc, strideF, offF, (const uint8_t*)planes0, stride0, planePitch0, (const uint8_t*)planes1, stride1,
                                                         planePitch1, offR, mv0, mv1, n, cost);
    else
        lowres_bidir_kernel<uint16_t><<<grid, 128, 0, st>>>((const uint16_t*)fenc, strideF, offF, (const uint16_t*)planes0, stride0, planePitch0, (const uint16_t*)planes1, stride1,
                                                          planePitch1, offR, mv0, mv1, n, cost);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}
