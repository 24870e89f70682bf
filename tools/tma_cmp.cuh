// tma_cmp.cuh (lab tool, not part of the library) -- SAD / SATD / SSE over pixel planes with the blocks staged in shared memory by TMA
// (cp.async.bulk.tensor.2d -> SASS UTMALDG), used by the tuning lab tools/satd_lab.cu and tools/tma_probe.cu only (TMA box origins must be 16-byte aligned, reference blocks are not).
//
// Why: a reference block sits at an arbitrary motion-vector offset, so a register-path kernel has to
// fetch the aligned superset of every row and realign it with funnel shifts, and the bytes it keeps in
// flight are bounded by its registers.  A 2-D tensor map over the plane (dim0 = stride, dim1 = rows)
// lets the TMA unit fetch the exact w x h box at element coordinates (x, y): the box lands dense and
// aligned in shared memory, and the bytes in flight are bounded by shared memory instead.
//
// Structure: every warp runs its own producer/consumer pipeline (no block-level synchronisation):
// a ring of `stages` buffers, one mbarrier each.  An "item" is `jobBytes` per plane: either one strip
// (w x hs) of a large block, or nb whole small blocks (one TMA box per block and plane, issued by
// lanes 0..nb-1).  After consuming item q the warp re-arms the same stage with item q + stages.
// A lane owns one 4x4 tile at a time, exactly like tile4_fast_kernel, and feeds the same
// tile4_accumulate(); rows of a tile are read in the order r ^ f(tile row), which makes the 8-byte
// shared loads of a half-warp hit distinct banks for every block width.  The XOR order is harmless:
// SAD/SSE sum over rows, and for SATD a dyadic shift of the inputs of the 4-point Hadamard only flips
// signs of whole output rows, which the horizontal pass and the abs() absorb.
//
// Blocks the tensor map cannot express (negative offset, row wrap) are flagged per item and read
// straight from global memory by the same warp.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "tile_kernels.cuh"

namespace b200 {

struct TmaCmpParams
{
    int n, w, h, kdiv;
    int sa, sb;            // plane strides in samples
    int hs;                // rows per item (== h when an item holds whole blocks)
    int nb;                // blocks per item (1 when an item is a strip)
    int lgStrips;          // log2(h / hs)
    int stages;
    int jobBytes;          // bytes per plane per stage
    int lgTw;              // log2(w / 4)
    int lgTpb;             // log2(tiles per block)
    int bankShift;         // see file header: f = (tile row >> bankShift) & 3
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do
    {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"((unsigned long long)map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// packed 16-bit pairs of the four samples of row `off` (bytes) of a dense shared-memory tile
__device__ __forceinline__ void lds_quad(const uint8_t* base, uint32_t off, uint16_t, uint32_t& lo, uint32_t& hi)
{
    uint2 q = *(const uint2*)(base + off);
    lo = q.x; hi = q.y;
}
__device__ __forceinline__ void lds_quad(const uint8_t* base, uint32_t off, uint8_t, uint32_t& lo, uint32_t& hi)
{
    uint32_t v = *(const uint32_t*)(base + off);
    lo = __byte_perm(v, 0, 0x4140); hi = __byte_perm(v, 0, 0x4342);
}

template<typename T, int OP, typename ACC, typename OUT>
__global__ void __launch_bounds__(512)
cmp_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const T* __restrict__ A, const T* __restrict__ B,
               const int32_t* __restrict__ offA, const int32_t* __restrict__ offB, TmaCmpParams p, OUT* __restrict__ out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int S = p.stages;
    const uint32_t stageBytes = 2u * (uint32_t)p.jobBytes;
    uint8_t* tiles = smem + (size_t)warp * S * stageBytes;
    uint64_t* bars = (uint64_t*)(smem + (size_t)W * S * stageBytes) + warp * S;
    uint32_t* flags = (uint32_t*)((uint64_t*)(smem + (size_t)W * S * stageBytes) + W * S) + warp * S;
    if (lane == 0)
        for (int s = 0; s < S; s++) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int gw = blockIdx.x * W + warp, GW = gridDim.x * W;
    const int groups = (p.n + p.nb - 1) / p.nb;
    const int mine = groups > gw ? (groups - gw + GW - 1) / GW : 0;
    const int Q = mine << p.lgStrips;
    const int strips = 1 << p.lgStrips;
    const int tw = 1 << p.lgTw;
    const uint32_t pitch = (uint32_t)p.w * sizeof(T);
    const uint32_t boxBytes = pitch * (uint32_t)p.hs;
    const int tilesPerItem = p.nb == 1 ? (p.hs >> 2) << p.lgTw : p.nb << p.lgTpb;

    auto issue = [&](int q, int stage)
    {
        int gi = gw + (q >> p.lgStrips) * GW;
        int s = q & (strips - 1);
        int blk = gi * p.nb + lane;
        bool have = lane < p.nb && blk < p.n;
        int oa = 0, ob = 0;
        if (have) { oa = offA[p.kdiv > 1 ? blk / p.kdiv : blk]; ob = offB[blk]; }
        int ya = (int)((unsigned)oa / (unsigned)p.sa), xa = oa - ya * p.sa;
        int yb = (int)((unsigned)ob / (unsigned)p.sb), xb = ob - yb * p.sb;
        bool bad = have && (oa < 0 || ob < 0 || xa + p.w > p.sa || xb + p.w > p.sb);
        unsigned anybad = __ballot_sync(0xffffffffu, bad);
        unsigned nvalid = __popc(__ballot_sync(0xffffffffu, have));
        uint32_t bar = smem_u32(&bars[stage]);
        if (lane == 0)
        {
            flags[stage] = anybad;
            if (anybad) mbar_arrive(bar);
            else mbar_arrive_expect_tx(bar, nvalid * 2u * boxBytes);
        }
        __syncwarp();
        if (!anybad && have)
        {
            uint32_t dst = smem_u32(tiles + (size_t)stage * stageBytes) + (uint32_t)lane * boxBytes;
            tma_load_2d(dst, &mapA, bar, xa, ya + s * p.hs);
            tma_load_2d(dst + (uint32_t)p.jobBytes, &mapB, bar, xb, yb + s * p.hs);
        }
    };

    int pre = Q < S ? Q : S;
    for (int q = 0; q < pre; q++) issue(q, q);

    ACC acc = 0;
    int stage = 0;
    uint32_t phase = 0;
    for (int q = 0; q < Q; q++)
    {
        mbar_wait(smem_u32(&bars[stage]), phase);
        const int gi = gw + (q >> p.lgStrips) * GW;
        const int s = q & (strips - 1);
        const uint8_t* sA = tiles + (size_t)stage * stageBytes;
        const uint8_t* sB = sA + p.jobBytes;
        if (flags[stage] == 0)
        {
            for (int t = lane; t < tilesPerItem; t += 32)
            {
                int u = t >> p.lgTw, tx = t & (tw - 1);
                uint32_t f = (uint32_t)(u >> p.bankShift) & 3u;
                uint32_t base = (uint32_t)u * 4u * pitch + (uint32_t)tx * 4u * (uint32_t)sizeof(T);
                uint32_t alo[4], ahi[4], blo[4], bhi[4];
#pragma unroll
                for (int r = 0; r < 4; r++)
                {
                    uint32_t off = base + ((uint32_t)r ^ f) * pitch;
                    lds_quad(sA, off, T(), alo[r], ahi[r]);
                    lds_quad(sB, off, T(), blo[r], bhi[r]);
                }
                tile4_accumulate<OP, ACC>(alo, ahi, blo, bhi, acc);
                if (p.lgTpb < 5)
                {   // several whole blocks per warp pass: reduce inside each group of tiles-per-block lanes
                    ACC v = group_sum(acc, 1 << p.lgTpb);
                    int blk = gi * p.nb + (t >> p.lgTpb);
                    if ((lane & ((1 << p.lgTpb) - 1)) == 0 && blk < p.n) out[blk] = (OUT)v;
                    acc = 0;
                }
            }
        }
        else
        {   // blocks outside the tensor map's reach: same arithmetic straight from global memory
            for (int bj = 0; bj < p.nb; bj++)
            {
                int blk = gi * p.nb + bj;
                if (blk >= p.n) break;
                const T* a = A + offA[p.kdiv > 1 ? blk / p.kdiv : blk] + (intptr_t)(s * p.hs) * p.sa;
                const T* b = B + offB[blk] + (intptr_t)(s * p.hs) * p.sb;
                ACC part = 0;
                int tl = (p.hs >> 2) << p.lgTw;
                for (int t = lane; t < tl; t += 32)
                {
                    int ty = t >> p.lgTw, tx = t & (tw - 1);
                    uint32_t alo[4], ahi[4], blo[4], bhi[4];
                    load_tile4x4(a + (intptr_t)(ty << 2) * p.sa + (tx << 2), p.sa, alo, ahi);
                    load_tile4x4(b + (intptr_t)(ty << 2) * p.sb + (tx << 2), p.sb, blo, bhi);
                    tile4_accumulate<OP, ACC>(alo, ahi, blo, bhi, part);
                }
                if (p.lgTpb < 5)
                {
                    part = group_sum(part, 32);
                    if (lane == 0) out[blk] = (OUT)part;
                }
                else
                    acc += part;
            }
        }
        if (p.lgTpb >= 5 && s == strips - 1)
        {
            ACC v = group_sum(acc, 32);
            if (lane == 0) out[gi] = (OUT)v;          // nb == 1: the group index is the block index
            acc = 0;
        }
        __syncwarp();
        if (q + S < Q) issue(q + S, stage);
        if (++stage == S) { stage = 0; phase ^= 1u; }
    }
}

// ---------------------------------------------------------------------------------------- host side

typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline tmap_encode_fn tmap_encoder()
{
    static tmap_encode_fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (tmap_encode_fn)p;
    }();
    return fn;
}

// 2-D map over a plane: dim0 = one row of `stride` samples, dim1 = as many rows as an int32 element offset can reach.
// The extent is only a bound for the TMA unit's out-of-range test; nothing outside the requested boxes is touched.
inline bool make_plane_map(CUtensorMap* m, const void* base, int elemBytes, intptr_t stride, int boxW, int boxH, int l2promo = 0)
{
    tmap_encode_fn enc = tmap_encoder();
    if (!enc) return false;
    cuuint64_t gdim[2] = { (cuuint64_t)stride, (cuuint64_t)(0x80000000ull / (cuuint64_t)stride + 64) };
    cuuint64_t gstr[1] = { (cuuint64_t)stride * elemBytes };
    cuuint32_t box[2] = { (cuuint32_t)boxW, (cuuint32_t)boxH };
    cuuint32_t est[2] = { 1, 1 };
    CUtensorMapL2promotion promo = l2promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : l2promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                 : l2promo == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    return enc(m, elemBytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), gdim, gstr,
               box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static inline int ilog2_exact(int v) { int l = 0; while ((1 << l) < v) l++; return (1 << l) == v ? l : -1; }

// fills p for the shape, or returns false when the TMA path does not apply (caller uses the register path)
inline bool tma_cmp_plan(TmaCmpParams& p, int elemBytes, int w, int h, intptr_t sa, intptr_t sb, const void* A, const void* B,
                         int jobBytes, int stages)
{
    int pitch = w * elemBytes, blockBytes = pitch * h;
    int tpb = (w >> 2) * (h >> 2);
    if (pitch < 16 || (pitch & 15) || ((blockBytes & 127) && blockBytes < jobBytes)) return false;
    if (ilog2_exact(w >> 2) < 0 || ilog2_exact(tpb) < 0) return false;
    if (((uintptr_t)A | (uintptr_t)B) & 15) return false;
    if (((sa * elemBytes) | (sb * elemBytes)) & 15) return false;
    if (sa < w || sb < w || sa > 0x7fffffff || sb > 0x7fffffff) return false;
    p.w = w; p.h = h; p.sa = (int)sa; p.sb = (int)sb;
    p.lgTw = ilog2_exact(w >> 2); p.lgTpb = ilog2_exact(tpb);
    p.stages = stages;
    if (blockBytes >= jobBytes)
    {
        int hs = jobBytes / pitch;
        if (hs < 4) hs = 4;
        while (h % hs) hs >>= 1;
        if (hs < 4 || (hs & 3) || ((pitch * hs) & 127)) return false;
        p.hs = hs; p.nb = 1; p.lgStrips = ilog2_exact(h / hs);
        if (p.lgStrips < 0) return false;
        p.jobBytes = pitch * hs;
        if (p.lgTpb < 5) return false;                      // a strip item needs whole-warp blocks
    }
    else
    {
        int nb = jobBytes / blockBytes;
        if (nb > 32) nb = 32;
        while ((nb * tpb) & 31) nb <<= 1;
        if (nb > 32) return false;
        p.hs = h; p.nb = nb; p.lgStrips = 0;
        p.jobBytes = nb * blockBytes;
        if (p.lgTpb >= 5 && nb != 1) { p.nb = 1; p.jobBytes = blockBytes; }
    }
    p.bankShift = 4 * pitch < 128 ? ilog2_exact(128 / (4 * pitch)) : 0;
    return true;
}

inline size_t tma_cmp_smem(const TmaCmpParams& p, int warps) { return (size_t)warps * p.stages * (2 * p.jobBytes + 16) + 16; }

// Launch on `st`.  warps per CTA and CTAs per SM are tuning knobs; returns false when the shape/planes do not fit the
// TMA path (nothing launched) so the caller can take the register path, or on a launch error (err set).
template<typename T, int OP, typename ACC, typename OUT>
inline bool launch_cmp_tma(const T* A, intptr_t sa, const T* B, intptr_t sb, const int32_t* offA, const int32_t* offB, int kdiv,
                           int n, int w, int h, OUT* out, cudaStream_t st, int smCount, int jobBytes, int stages, int warps,
                           int l2promo, cudaError_t* err)
{
    TmaCmpParams p;
    *err = cudaSuccess;
    if (!tma_cmp_plan(p, (int)sizeof(T), w, h, sa, sb, A, B, jobBytes, stages)) return false;
    p.n = n; p.kdiv = kdiv;
    CUtensorMap ma, mb;
    if (!make_plane_map(&ma, A, (int)sizeof(T), sa, w, p.hs, l2promo) || !make_plane_map(&mb, B, (int)sizeof(T), sb, w, p.hs, l2promo)) return false;
    size_t smem = tma_cmp_smem(p, warps);
    if (smem > 227 * 1024) return false;
    auto kern = cmp_tma_kernel<T, OP, ACC, OUT>;
    static size_t configured = 0;                        // per instantiation
    if (smem > configured)
    {
        *err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (*err != cudaSuccess) return false;
        configured = smem;
    }
    int perSM = (int)((227 * 1024) / (smem + 1024));
    if (perSM < 1) perSM = 1;
    if (perSM * warps > 64) perSM = 64 / warps;
    int groups = (n + p.nb - 1) / p.nb;
    int grid = smCount * perSM;
    int need = (groups + warps - 1) / warps;
    if (grid > need) grid = need;
    kern<<<grid, warps * 32, smem, st>>>(ma, mb, A, B, offA, offB, p, out);
    *err = cudaGetLastError();
    return *err == cudaSuccess;
}

} // namespace b200
