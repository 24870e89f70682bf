// tme.cu -- the ThreadedME contract as one call: every (PU, reference picture) search of a CTU row -- or of a whole frame -- handed over as an
// array of records, searched on the GPU, and returned as the per-reference part of MEData (reference encoder/threadedme.h:112-130; the
// searches are the inner loop of Search::puMotionEstimation, encoder/search.cpp:264-404, which ThreadedME::findJob drives per CTU through
// Analysis::deriveMVsForCTU, encoder/threadedme.cpp:207-261).
//
// A caller has HOST records (shapes, offsets, predictors and neighbour vectors differ per PU) and wants HOST results; the planes are resident
// (x265b200_plane).  The call groups the records by (shape, reference, number of candidates) -- the batched search entries take one shape and
// one reference per launch -- runs x265b200_motion_estimate_batch per group, and finishes each search with the cost bookkeeping of
// search.cpp:392-394: bits += bitcost(mv), mvCost = mvcost(mv), cost = (satdCost - mvCost) + rdCost.getCost(bits).  Choosing the best
// reference per list, AMVP index refinement (checkBestMVP) and the bi-prediction candidate stay with the caller: they need the CU's
// neighbour context; the bi-prediction cost itself is x265b200_bidir_satd_batch.
#include "internal.h"

#include <algorithm>
#include <string.h>
#include <vector>

struct x265b200_plane;
extern "C" int x265b200_plane_info(const x265b200_plane* p, intptr_t* stride, int* rows, int32_t* origin, size_t* elems, void** device);

namespace b200 {

// per search: vector bits and costs (encoder/bitcost.h:53-63, encoder/rdcost.h:164-169)
__global__ void tme_finish_kernel(int n, const int32_t* __restrict__ qmv, const int32_t* __restrict__ satd, const int32_t* __restrict__ qmvp,
                                  const uint32_t* __restrict__ bits0, const uint16_t* __restrict__ costTab, const float* __restrict__ bitsTab,
                                  unsigned long long lambda, x265b200_tme_result* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int mx = qmv[2 * i], my = qmv[2 * i + 1], dx = mx - qmvp[2 * i], dy = my - qmvp[2 * i + 1];
    const uint32_t mvCost = (uint16_t)(costTab[dx] + costTab[dy]);
    const uint32_t bits = bits0[i] + (uint32_t)(bitsTab[dx] + bitsTab[dy] + 0.5f);
    x265b200_tme_result r;
    r.mv[0] = mx; r.mv[1] = my;
    r.mvCost = mvCost;
    r.bits = bits;
    r.satdCost = (uint32_t)satd[i];
    r.cost = ((uint32_t)satd[i] - mvCost) + (uint32_t)(((unsigned long long)bits * lambda + 128) >> 8);
    out[i] = r;
}

} // namespace b200

using namespace b200;

extern "C" int x265b200_tme_search_batch(x265b200_ctx* ctx, int searchMethod, int merange, int subpelRefine, const x265b200_plane* fencPlane,
                                         const x265b200_plane* const* refPlanes, int numRefs, const uint16_t* costTab, const float* bitsTab,
                                         int tabRadius, uint64_t lambda, const x265b200_tme_pu* pus, int n, x265b200_tme_result* results)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (n < 0 || !fencPlane || !refPlanes || numRefs < 1 || !costTab || !bitsTab || tabRadius < 1 || (n && (!pus || !results)))
        return fail(ctx, X265B200_ERR_ARG, "tme_search: bad arguments");
    if (searchMethod == X265B200_ME_SEA) return fail(ctx, X265B200_ERR_ARG, "tme_search: X265_SEA needs integral planes; use x265b200_motion_estimate_sea_batch");
    if (n == 0) return X265B200_OK;
    B200_CUDA(ctx, cudaSetDevice(ctx->device));
    intptr_t strideF = 0; void* dF = nullptr; size_t elems = 0;
    x265b200_plane_info(fencPlane, &strideF, nullptr, nullptr, &elems, &dF);
    for (int i = 0; i < n; i++)
    {
        const x265b200_tme_pu& p = pus[i];
        if (p.w < 4 || p.w > 64 || p.h < 4 || p.h > 64 || (p.w & 3) || (p.h & 3) || p.ref < 0 || p.ref >= numRefs || p.numCand < 0 || p.numCand > X265B200_TME_MAX_CAND ||
            p.offF < 0 || p.offR < 0 || (size_t)p.offF + (size_t)(p.h - 1) * strideF + p.w > elems)
            return fail(ctx, X265B200_ERR_ARG, "tme_search: bad PU record");
    }
    // group by (shape, reference, candidate count); the order inside a group is the caller's
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    auto key = [&](int i) { const x265b200_tme_pu& p = pus[i]; return ((((long long)p.w << 8 | p.h) << 8 | p.ref) << 8) | p.numCand; };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key(a) < key(b); });

    // staging: SoA arrays in the sorted order, one pinned block up and one down
    const size_t K = X265B200_TME_MAX_CAND;
    const size_t inInts = (size_t)n * (1 + 1 + 4 + 2 + 1 + 2 * K);
    int32_t *hIn = nullptr, *dIn = nullptr, *dOut = nullptr;
    x265b200_tme_result *hRes = nullptr, *dRes = nullptr;
    uint16_t* dCost = nullptr; float* dBits = nullptr;
    cudaStream_t st = nullptr;
    int rc = X265B200_OK;
    auto cleanup = [&]()
    {
        if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
        cudaFreeHost(hIn); cudaFreeHost(hRes); cudaFree(dIn); cudaFree(dOut); cudaFree(dRes); cudaFree(dCost); cudaFree(dBits);
    };
#define TME_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = fail(ctx, X265B200_ERR_CUDA, #call, e__); cleanup(); return rc; } } while (0)
    TME_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    TME_CUDA(cudaMallocHost((void**)&hIn, inInts * 4));
    TME_CUDA(cudaMallocHost((void**)&hRes, (size_t)n * sizeof(x265b200_tme_result)));
    TME_CUDA(cudaMalloc((void**)&dIn, inInts * 4));
    TME_CUDA(cudaMalloc((void**)&dOut, (size_t)n * 3 * 4));
    TME_CUDA(cudaMalloc((void**)&dRes, (size_t)n * sizeof(x265b200_tme_result)));
    const size_t tabN = 2 * (size_t)tabRadius + 1;
    TME_CUDA(cudaMalloc((void**)&dCost, tabN * 2));
    TME_CUDA(cudaMalloc((void**)&dBits, tabN * 4));
    TME_CUDA(cudaMemcpyAsync(dCost, costTab - tabRadius, tabN * 2, cudaMemcpyHostToDevice, st));
    TME_CUDA(cudaMemcpyAsync(dBits, bitsTab - tabRadius, tabN * 4, cudaMemcpyHostToDevice, st));
    int32_t* hOffF = hIn; int32_t* hOffR = hOffF + n; int32_t* hRange = hOffR + n; int32_t* hMvp = hRange + 4 * (size_t)n;
    int32_t* hBits = hMvp + 2 * (size_t)n; int32_t* hMvc = hBits + n;
    size_t mvcPos = 0;
    std::vector<size_t> mvcStart(n);
    for (int s = 0; s < n; s++)
    {
        const x265b200_tme_pu& p = pus[order[s]];
        hOffF[s] = p.offF; hOffR[s] = p.offR;
        hRange[4 * s] = p.mvmin[0]; hRange[4 * s + 1] = p.mvmin[1]; hRange[4 * s + 2] = p.mvmax[0]; hRange[4 * s + 3] = p.mvmax[1];
        hMvp[2 * s] = p.mvp[0]; hMvp[2 * s + 1] = p.mvp[1];
        hBits[s] = (int32_t)p.bits;
        mvcStart[s] = mvcPos;
        for (int k = 0; k < p.numCand; k++) { hMvc[mvcPos++] = p.mvc[k][0]; hMvc[mvcPos++] = p.mvc[k][1]; }
    }
    TME_CUDA(cudaMemcpyAsync(dIn, hIn, inInts * 4, cudaMemcpyHostToDevice, st));
    ctx->h2d_bytes.fetch_add(inInts * 4 + tabN * 6, std::memory_order_relaxed);
    int32_t* dOffF = dIn; int32_t* dOffR = dOffF + n; int32_t* dRange = dOffR + n; int32_t* dMvp = dRange + 4 * (size_t)n;
    int32_t* dBits0 = dMvp + 2 * (size_t)n; int32_t* dMvc = dBits0 + n;
    int32_t* dQmv = dOut; int32_t* dSatd = dOut + 2 * (size_t)n;
    for (int s0 = 0; s0 < n;)
    {
        int s1 = s0 + 1;
        while (s1 < n && key(order[s1]) == key(order[s0])) s1++;
        const x265b200_tme_pu& p = pus[order[s0]];
        intptr_t strideR = 0; void* dR = nullptr; size_t relems = 0;
        x265b200_plane_info(refPlanes[p.ref], &strideR, nullptr, nullptr, &relems, &dR);
        rc = x265b200_motion_estimate_batch(ctx, searchMethod, p.w, p.h, merange, subpelRefine, dF, strideF, dR, strideR, dOffF + s0, dOffR + s0,
                                            dRange + 4 * (size_t)s0, dMvp + 2 * (size_t)s0, p.numCand, p.numCand ? dMvc + mvcStart[s0] : nullptr,
                                            dCost + tabRadius, s1 - s0, dQmv + 2 * (size_t)s0, dSatd + s0, (x265b200_stream)st);
        if (rc != X265B200_OK) { cleanup(); return rc; }
        s0 = s1;
    }
    tme_finish_kernel<<<ceil_div(n, 128), 128, 0, st>>>(n, dQmv, dSatd, dMvp, (const uint32_t*)dBits0, dCost + tabRadius, dBits + tabRadius,
                                                        (unsigned long long)lambda, dRes);
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    TME_CUDA(cudaGetLastError());
    TME_CUDA(cudaMemcpyAsync(hRes, dRes, (size_t)n * sizeof(x265b200_tme_result), cudaMemcpyDeviceToHost, st));
    TME_CUDA(cudaStreamSynchronize(st));
    ctx->d2h_bytes.fetch_add((size_t)n * sizeof(x265b200_tme_result), std::memory_order_relaxed);
    for (int s = 0; s < n; s++) results[order[s]] = hRes[s];
#undef TME_CUDA
    cleanup();
    return X265B200_OK;
}
