// tu_chain.cu -- inter luma TU reconstruction chain as ONE C-ABI call:
//     resi = fenc - pred -> dct -> quant -> [numSig == 0: recon = pred]
//                              -> dequant_normal -> (DC-only shortcut | idct) -> recon = clip(pred + resi') -> sse
// i.e. the sequence reference encoder/search.cpp:5536-5575 drives through quant.cpp:397-480 (transformNxN) and
// quant.cpp:543-605 (invtransformNxN), without RDOQ / psy / sign hiding / transform skip and with scaling lists off.
//
// B200 mapping: the stage kernels are the batched primitives of this library (tensor-core DCT/IDCT); the chain is
// walked in chunks whose int16 intermediates (residual, coefficients, dequantised coefficients, reconstructed
// residual) total a few MB, so they never leave the 126 MB L2: HBM sees fenc + pred in, qCoef + recon + costs out.
#include <stdlib.h>
#include "internal.h"
#include "device_util.cuh"

namespace b200 {

int launch_quant(x265b200_ctx* ctx, int mode, const int16_t* coef, const int32_t* quantCoeff, int32_t* deltaU, int16_t* qCoef,
                 int qBits, int add, int numCoeff, int n, uint32_t* numSig, cudaStream_t st);       // transform.cu
bool launch_tu_fused(x265b200_ctx* ctx, int N, const void* fenc, intptr_t sf, const void* pred, intptr_t sp,
                     const int32_t* offF, const int32_t* offP, int n, const int32_t* quantCoeff, int qBits, int qAdd,
                     int dqScale, int dqShift, int16_t* qCoef, uint32_t* numSig, void* recon, intptr_t sr,
                     const int32_t* offR, uint64_t* sseZero, uint64_t* sseRecon, cudaStream_t st, int dst4);   // tu_fused.cuh (transform_mma.cu)

bool launch_tu_umma(x265b200_ctx* ctx, int N, const void* fenc, intptr_t sf, const void* pred, intptr_t sp,
                    const int32_t* offF, const int32_t* offP, int n, const int32_t* quantCoeff, int qBits, int qAdd,
                    int dqScale, int dqShift, int16_t* qCoef, uint32_t* numSig, void* recon, intptr_t sr,
                    const int32_t* offR, uint64_t* sseZero, uint64_t* sseRecon, cudaStream_t st);    // tu_umma.cuh (transform_mma.cu)
bool launch_tu_forward(x265b200_ctx* ctx, int N, const void* fenc, intptr_t sf, const void* pred, intptr_t sp,
                       const int32_t* offF, const int32_t* offP, int n, const int32_t* quantCoeff, int qBits, int qAdd,
                       int16_t* qCoef, uint32_t* numSig, uint64_t* sseZero, cudaStream_t st, int dst4);        // tu_fused.cuh

// recon = clip(pred + resi') with the cbf == 0 and DC-only cases, plus both distortions; 4 samples per thread
template<typename T>
__global__ void __launch_bounds__(256)
recon_kernel(const T* __restrict__ fenc, intptr_t sf, const T* __restrict__ pred, intptr_t sp,
             const int32_t* __restrict__ offF, const int32_t* __restrict__ offP, int n, int N,
             const int16_t* __restrict__ resi, const int16_t* __restrict__ dq, const int16_t* __restrict__ qCoef,
             const uint32_t* __restrict__ numSig, T* __restrict__ recon, intptr_t sr, const int32_t* __restrict__ offR,
             unsigned long long* __restrict__ sseZero, unsigned long long* __restrict__ sseRecon, int depth, int dcShortcut)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int qw = N >> 2;
    int per = qw * N;                                    // threads per TU (4, 16, 64, 256)
    int tu = (int)(gid / per);
    bool live = tu < n;
    unsigned long long z = 0, d = 0;
    if (live)
    {
        int r = (int)(gid % per);
        int y = r / qw, x = (r % qw) << 2;
        uint32_t wf[2], wp[2];
        load_row_quads<1>(fenc + offF[tu] + (intptr_t)y * sf + x, wf);
        load_row_quads<1>(pred + offP[tu] + (intptr_t)y * sp + x, wp);
        int f[4], p[4], v[4];
        SampleTraits<T>::unpack(wf[0], f[0], f[1]); SampleTraits<T>::unpack(wf[1], f[2], f[3]);
        SampleTraits<T>::unpack(wp[0], p[0], p[1]); SampleTraits<T>::unpack(wp[1], p[2], p[3]);
        uint32_t ns = numSig[tu];
        size_t base = (size_t)tu * N * N;
        int rr[4] = { 0, 0, 0, 0 };
        if (ns == 1 && qCoef[base] != 0 && dcShortcut)
        {
            // DC only, quant.cpp:588-598
            const int shift_2nd = 12 - (depth - 8) - 3, add_2nd = 1 << (shift_2nd - 1);
            int dc = (int)(int16_t)(((((int)dq[base] + 1) >> 1) * 8 + add_2nd) >> shift_2nd);
            rr[0] = rr[1] = rr[2] = rr[3] = dc;
        }
        else if (ns)
        {
            uint2 q = *(const uint2*)(resi + base + y * N + x);
            rr[0] = (int16_t)(q.x & 0xffff); rr[1] = (int)q.x >> 16; rr[2] = (int16_t)(q.y & 0xffff); rr[3] = (int)q.y >> 16;
        }
        int maxv = (1 << depth) - 1;
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            v[i] = ns ? min(max(p[i] + rr[i], 0), maxv) : p[i];                  // add_ps (pixel.cpp:821-831) or the prediction
            int e0 = f[i] - p[i], e1 = f[i] - v[i];
            z += (unsigned)(e0 * e0);
            d += (unsigned)(e1 * e1);
        }
        T* o = recon + offR[tu] + (intptr_t)y * sr + x;
        uintptr_t a = (uintptr_t)o;
        if (sizeof(T) == 2 && (a & 7) == 0)
            *(uint2*)o = make_uint2((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16));
        else if (sizeof(T) == 1 && (a & 3) == 0)
            *(uint32_t*)o = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
        else { o[0] = (T)v[0]; o[1] = (T)v[1]; o[2] = (T)v[2]; o[3] = (T)v[3]; }
    }
    int G = per < 32 ? per : 32;
    z = group_sum(z, G);
    d = group_sum(d, G);
    if (live && ((threadIdx.x & 31) & (G - 1)) == 0)
    {
        if (per <= 32) { if (sseZero) sseZero[tu] = z; sseRecon[tu] = d; }
        else { if (sseZero) atomicAdd(sseZero + tu, z); atomicAdd(sseRecon + tu, d); }
    }
}

} // namespace b200

using namespace b200;

// ttype: X265B200_TU_INTER (inter luma and every chroma TU: the chroma loop of estimateResidualQT, search.cpp:5638-5700, is this chain on the
// chroma planes with log2TrSizeC) or X265B200_TU_INTRA_LUMA (a 4x4 TU takes DST-VII and no DC-only shortcut, quant.cpp:430-441, :585-603)
extern "C" int x265b200_tu_chain_tt_batch(x265b200_ctx* ctx, int N, int ttype, const void* fenc, intptr_t strideF, const void* pred,
                                          intptr_t strideP, const int32_t* offF, const int32_t* offP, int n,
                                          const int32_t* quantCoeff, int qBits, int add, int dqScale, int dqShift,
                                          int16_t* qCoef, uint32_t* numSig, void* recon, intptr_t strideR, const int32_t* offR,
                                          uint64_t* sseZero, uint64_t* sseRecon, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if ((N != 4 && N != 8 && N != 16 && N != 32) || n < 0 || qBits < 8 || dqShift < 1)
        return fail(ctx, X265B200_ERR_ARG, "tu_chain: bad size / parameters");
    if (ttype != X265B200_TU_INTER && ttype != X265B200_TU_INTRA_LUMA) return fail(ctx, X265B200_ERR_ARG, "tu_chain: unknown TU type");
    if ((strideF | strideP) & 3) return fail(ctx, X265B200_ERR_ARG, "tu_chain: plane strides must be multiples of 4 samples");
    if (n == 0) return X265B200_OK;
    const int dst4 = ttype == X265B200_TU_INTRA_LUMA && N == 4;
    if (!fenc || !pred || !recon || !offF || !offP || !offR || !quantCoeff || !qCoef || !numSig || !sseRecon)
        return fail(ctx, X265B200_ERR_ARG, "tu_chain: only sseZero may be NULL (offF, offP, offR and sseRecon are required)");
    cudaStream_t st = (cudaStream_t)stream;
    const int NN = N * N;
    // ONE kernel on the 5th-generation tensor cores, accumulators in tensor memory (tu_umma.cuh): the default for N = 32, where it beats the
    // two mma.sync kernels (0.50 vs 0.55 ms per 16 frames of 2160p10, and 1.0 x instead of 1.75 x the algorithmic DRAM traffic); at N = 16 it is
    // 10 % slower than they are and only runs on request (path 3).
    if (((ctx->dct_path == 0 && N == 32) || (ctx->dct_path == 3 && (N == 32 || N == 16))) &&
        launch_tu_umma(ctx, N, fenc, strideF, pred, strideP, offF, offP, n, quantCoeff, qBits, add, dqScale, dqShift,
                       qCoef, numSig, recon, strideR, offR, sseZero, sseRecon, st))
        return X265B200_OK;
    if (cudaGetLastError() != cudaSuccess) return fail(ctx, X265B200_ERR_CUDA, "tu_chain tcgen05 launch");
    // N = 16 / 8 / 4 (and path 2 for every size): two fused mma.sync kernels over all TUs, no scratch (tu_fused.cuh).
    // dct_path == 1 (validation twin): the six stage kernels of the batched primitives, walked in chunks whose two
    // int16 scratch planes stay L2-resident.
    if (ctx->dct_path != 1 &&
        launch_tu_fused(ctx, N, fenc, strideF, pred, strideP, offF, offP, n, quantCoeff, qBits, add, dqScale, dqShift,
                        qCoef, numSig, recon, strideR, offR, sseZero, sseRecon, st, dst4))
        return X265B200_OK;
    if (cudaGetLastError() != cudaSuccess) return fail(ctx, X265B200_ERR_CUDA, "tu_chain fused launch");
    const int chunkTUs = (8 << 20) / NN;
    int16_t* scratch = nullptr;
    size_t chunkElems = (size_t)(n < chunkTUs ? n : chunkTUs) * NN;
    B200_CUDA(ctx, cudaMallocAsync((void**)&scratch, 2 * chunkElems * sizeof(int16_t), st));
    int16_t* s0 = scratch;
    int16_t* s1 = scratch + chunkElems;
    if (NN > 32)
    {
        if (sseZero) B200_CUDA(ctx, cudaMemsetAsync(sseZero, 0, (size_t)n * 8, st));
        B200_CUDA(ctx, cudaMemsetAsync(sseRecon, 0, (size_t)n * 8, st));
    }
    int rc = X265B200_OK;
    for (int c0 = 0; c0 < n && rc == X265B200_OK; c0 += chunkTUs)
    {
        int m = n - c0 < chunkTUs ? n - c0 : chunkTUs;
        int16_t* q = qCoef + (size_t)c0 * NN;
        rc = x265b200_residual_batch(ctx, N, N, fenc, strideF, pred, strideP, offF + c0, offP + c0, m, s0, stream);
        if (rc == X265B200_OK) rc = x265b200_dct_batch(ctx, dst4 ? X265B200_TR_DST : X265B200_TR_DCT, N, s0, N, nullptr, m, s1, stream);
        if (rc == X265B200_OK) rc = launch_quant(ctx, 2, s1, quantCoeff, nullptr, q, qBits, add, NN, m, numSig + c0, st);
        if (rc == X265B200_OK) rc = x265b200_dequant_normal_batch(ctx, q, s0, m * NN, dqScale, dqShift, stream);
        if (rc == X265B200_OK) rc = x265b200_idct_batch(ctx, dst4 ? X265B200_TR_DST : X265B200_TR_DCT, N, s0, m, s1, N, nullptr, stream);
        if (rc != X265B200_OK) break;
        long long threads = (long long)m * (NN >> 2);
        if (ctx->pixbytes == 1)
            recon_kernel<uint8_t><<<ceil_div(threads, 256), 256, 0, st>>>((const uint8_t*)fenc, strideF, (const uint8_t*)pred, strideP, offF + c0, offP + c0, m, N,
                                                                        s1, s0, q, numSig + c0, (uint8_t*)recon, strideR, offR + c0,
                                                                        (unsigned long long*)(sseZero ? sseZero + c0 : nullptr), (unsigned long long*)sseRecon + c0, ctx->depth, !dst4);
        else
            recon_kernel<uint16_t><<<ceil_div(threads, 256), 256, 0, st>>>((const uint16_t*)fenc, strideF, (const uint16_t*)pred, strideP, offF + c0, offP + c0, m, N,
                                                                         s1, s0, q, numSig + c0, (uint16_t*)recon, strideR, offR + c0,
                                                                         (unsigned long long*)(sseZero ? sseZero + c0 : nullptr), (unsigned long long*)sseRecon + c0, ctx->depth, !dst4);
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
        if (cudaGetLastError() != cudaSuccess) rc = fail(ctx, X265B200_ERR_CUDA, "recon_kernel launch");
    }
    cudaFreeAsync(scratch, st);
    return rc;
}

extern "C" int x265b200_tu_chain_batch(x265b200_ctx* ctx, int N, const void* fenc, intptr_t strideF, const void* pred,
                                       intptr_t strideP, const int32_t* offF, const int32_t* offP, int n,
                                       const int32_t* quantCoeff, int qBits, int add, int dqScale, int dqShift,
                                       int16_t* qCoef, uint32_t* numSig, void* recon, intptr_t strideR, const int32_t* offR,
                                       uint64_t* sseZero, uint64_t* sseRecon, x265b200_stream stream)
{
    return x265b200_tu_chain_tt_batch(ctx, N, X265B200_TU_INTER, fenc, strideF, pred, strideP, offF, offP, n, quantCoeff, qBits, add, dqScale, dqShift,
                                      qCoef, numSig, recon, strideR, offR, sseZero, sseRecon, stream);
}

// forward half of the chain: what Quant::transformNxN (reference common/quant.cpp:397-480) does for an inter luma TU
extern "C" int x265b200_tu_forward_batch(x265b200_ctx* ctx, int N, const void* fenc, intptr_t strideF, const void* pred, intptr_t strideP,
                                         const int32_t* offF, const int32_t* offP, int n, const int32_t* quantCoeff, int qBits, int add,
                                         int16_t* qCoef, uint32_t* numSig, uint64_t* sseZero, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if ((N != 4 && N != 8 && N != 16 && N != 32) || n < 0 || qBits < 8)
        return fail(ctx, X265B200_ERR_ARG, "tu_forward: bad size / parameters");
    if ((strideF | strideP) & 3) return fail(ctx, X265B200_ERR_ARG, "tu_forward: plane strides must be multiples of 4 samples");
    if (n == 0) return X265B200_OK;
    if (!fenc || !pred || !offF || !offP || !quantCoeff || !qCoef || !numSig)
        return fail(ctx, X265B200_ERR_ARG, "tu_forward: only sseZero may be NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->dct_path != 1 && launch_tu_forward(ctx, N, fenc, strideF, pred, strideP, offF, offP, n, quantCoeff, qBits, add, qCoef, numSig, sseZero, st, 0))
        return X265B200_OK;
    if (cudaGetLastError() != cudaSuccess) return fail(ctx, X265B200_ERR_CUDA, "tu_forward fused launch");
    // validation twin / unaligned operands: the stage kernels over chunks whose scratch stays L2-resident
    const int NN = N * N;
    const int chunkTUs = (8 << 20) / NN;
    int16_t* scratch = nullptr;
    size_t chunkElems = (size_t)(n < chunkTUs ? n : chunkTUs) * NN;
    B200_CUDA(ctx, cudaMallocAsync((void**)&scratch, 2 * chunkElems * sizeof(int16_t), st));
    int rc = X265B200_OK;
    for (int c0 = 0; c0 < n && rc == X265B200_OK; c0 += chunkTUs)
    {
        int m = n - c0 < chunkTUs ? n - c0 : chunkTUs;
        rc = x265b200_residual_batch(ctx, N, N, fenc, strideF, pred, strideP, offF + c0, offP + c0, m, scratch, stream);
        if (rc == X265B200_OK) rc = x265b200_dct_batch(ctx, X265B200_TR_DCT, N, scratch, N, nullptr, m, scratch + chunkElems, stream);
        if (rc == X265B200_OK) rc = launch_quant(ctx, 2, scratch + chunkElems, quantCoeff, nullptr, qCoef + (size_t)c0 * NN, qBits, add, NN, m, numSig + c0, st);
        if (rc == X265B200_OK && sseZero)
            rc = x265b200_pixelcmp_batch(ctx, X265B200_SSE_PP, N, N, fenc, strideF, pred, strideP, offF + c0, offP + c0, m, sseZero + c0, stream);
    }
    cudaFreeAsync(scratch, st);
    return rc;
}
