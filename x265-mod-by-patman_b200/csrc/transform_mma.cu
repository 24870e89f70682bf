// transform_mma.cu -- forward integer DCT 4/8/16/32 (and the 4x4 DST) on the tensor cores (IMMA s8 x s8/u8 -> s32).
//
// The two transform stages are dense int16 x int8 contractions (reference dct.cpp:83-240 computes them
// with partial butterflies; SURVEY.md appendix D shows the full-matrix form is bit-identical):
//     stage 1:  C1[k][j]   = (sum_i T[k][i]  * X[j][i]   + add1) >> shift1      (truncated to int16)
//     stage 2:  out[k2][k] = (sum_j T[k2][j] * C1[k][j]  + add2) >> shift2      (truncated to int16)
// Both are issued as  D = A * B  with A = T (constant, s8, |T| <= 90) and the int16 operand split
// into a signed high byte and an unsigned low byte, B = 256 * Bhi + Blo, i.e. two IMMAs
// (mma.sync m16n8k32 / m16n8k16, SASS IMMA.16832 / IMMA.16816) whose s32 results are recombined
// exactly: |sum| <= 32 * 90 * 32768 < 2^31.
//
// No shared memory and no shuffles: X (row-major) already is the "col" B operand of stage 1, and the
// accumulator fragment of stage 1 is, element for element, a B-operand fragment of stage 2 once the
// contraction index j is enumerated in the order the accumulator layout delivers it; that
// permutation is folded into the constant A operand of stage 2 (table built on the host).
#include <stdlib.h>
#include <string.h>
#include "internal.h"
#include "device_util.cuh"

namespace b200 {

// per-lane A fragments, [stage][mtile][reg][lane]; N = 32: 2 x 2 x 4 x 32, N = 16: 2 x 1 x 2 x 32.
// Global (not __constant__) memory: the index is the lane, and per-lane constant-bank reads serialise.
__device__ uint32_t c_A32[2][2][4][32];
__device__ uint32_t c_A16[2][2][32];
// N = 8 (two TUs per MMA) and N = 4 (eight TUs per MMA; [1] = DST-VII): block-diagonal A, [kind][stage][reg][lane]
__device__ uint32_t c_A8[2][2][32];
__device__ uint32_t c_A4[2][2][2][32];
// inverse transforms: stage 1 uses a constant A (T^T, K order chosen for coalesced gathers), stage 2 uses the
// stage-1 accumulators as A and a constant B.  [mt][reg][lane] / [ntile][reg][lane]
__device__ uint32_t c_IA32[2][4][32];        // stage-1 A fragments, N = 32
__device__ uint32_t c_IB32[4][2][32];        // stage-2 B fragments, N = 32
__device__ uint32_t c_IA16[2][32];
__device__ uint32_t c_IB16[2][32];
__device__ uint32_t c_IA8[2][32];            // block-diagonal: two 8x8 TUs
__device__ uint32_t c_IB8[32];
__device__ uint32_t c_IA4[2][2][32];         // [kind: 0 DCT, 1 DST][reg][lane]: eight 4x4 TUs
__device__ uint32_t c_IB4[2][32];
__device__ uint8_t c_ummaB[2][2][1024];      // tcgen05 B tiles (tu_umma.cuh)
__device__ uint8_t c_ummaAD[2][7168];        // tcgen05: the non-zero block of diag(T, T, ..) between two runs of zero row groups (tu_umma.cuh)

static const short h_cosmag[32] = { 64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67,
                                    64, 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4 };
static int h_cos128(int m)
{
    m &= 127;
    if (m > 64) m = 128 - m;
    if (m == 32) return 0;
    if (m > 32) return -h_cosmag[64 - m];
    return h_cosmag[m];
}
static int tcoef(int N, int k, int i) { return h_cos128(k * (32 / N) * (2 * i + 1)); }

int upload_mma_tables(x265b200_ctx* ctx)
{
    // N = 32, m16n8k32: a0 (row g, k 4t+e)  a1 (row g+8, k 4t+e)  a2 (row g, k 16+4t+e)  a3 (row g+8, k 16+4t+e)
    // stage 2 column order: slot s = half*16 + 4t' + e  <->  j = (2*half + (e>>1))*8 + 2t' + (e&1)
    static uint32_t a32[2][2][4][32], a16[2][2][32];
    for (int stage = 0; stage < 2; stage++)
        for (int mt = 0; mt < 2; mt++)
            for (int r = 0; r < 4; r++)
                for (int lane = 0; lane < 32; lane++)
                {
                    int g = lane >> 2, t = lane & 3;
                    int row = mt * 16 + g + (r & 1) * 8;
                    uint32_t v = 0;
                    for (int e = 0; e < 4; e++)
                    {
                        int half = r >> 1;
                        int col = stage == 0 ? half * 16 + 4 * t + e : (2 * half + (e >> 1)) * 8 + 2 * t + (e & 1);
                        v |= (uint32_t)(uint8_t)(int8_t)tcoef(32, row, col) << (8 * e);
                    }
                    a32[stage][mt][r][lane] = v;
                }
    // N = 16, m16n8k16: a0 (row g, k 4t+e)  a1 (row g+8, k 4t+e);  stage 2: slot 4t'+e <-> j = (e>>1)*8 + 2t' + (e&1)
    for (int stage = 0; stage < 2; stage++)
        for (int r = 0; r < 2; r++)
            for (int lane = 0; lane < 32; lane++)
            {
                int g = lane >> 2, t = lane & 3;
                int row = g + r * 8;
                uint32_t v = 0;
                for (int e = 0; e < 4; e++)
                {
                    int col = stage == 0 ? 4 * t + e : (e >> 1) * 8 + 2 * t + (e & 1);
                    v |= (uint32_t)(uint8_t)(int8_t)tcoef(16, row, col) << (8 * e);
                }
                a16[stage][r][lane] = v;
            }
    // N = 8: rows 0-7 = TU a, rows 8-15 = TU b.  stage 1: slot s = (tu = s >> 3, i = s & 7);
    // stage 2: slot 4t'+e = (tu = e >> 1, j = 2t' + (e & 1))
    static uint32_t a8[2][2][32], a4[2][2][2][32];
    static const int dst4[4][4] = { { 29, 55, 74, 84 }, { 74, 74, 0, -74 }, { 84, -29, -74, 55 }, { 55, -84, 74, -29 } };
    for (int stage = 0; stage < 2; stage++)
        for (int r = 0; r < 2; r++)
            for (int lane = 0; lane < 32; lane++)
            {
                int g = lane >> 2, t = lane & 3;
                int row = g + r * 8;
                uint32_t v8 = 0, v4[2] = { 0, 0 };
                for (int e = 0; e < 4; e++)
                {
                    // --- N = 8
                    int tu_s = stage == 0 ? t >> 1 : e >> 1;
                    int idx = stage == 0 ? 4 * (t & 1) + e : 2 * t + (e & 1);
                    int c8 = (row >> 3) == tu_s ? tcoef(8, row & 7, idx) : 0;
                    v8 |= (uint32_t)(uint8_t)(int8_t)c8 << (8 * e);
                    // --- N = 4.  stage 1: row = (tu = row >> 2, k = row & 3), slot = (tu = t, i = e)
                    //             stage 2: row = (set = row >> 3, h = (row >> 2) & 1, k2 = row & 3),
                    //                      slot = (set = t >> 1, h = e >> 1, j = 2 (t & 1) + (e & 1))
                    for (int kind = 0; kind < 2; kind++)
                    {
                        int c4;
                        if (stage == 0)
                            c4 = (row >> 2) == t ? (kind ? dst4[row & 3][e] : tcoef(4, row & 3, e)) : 0;
                        else
                        {
                            int j = 2 * (t & 1) + (e & 1);
                            bool on = (row >> 3) == (t >> 1) && ((row >> 2) & 1) == (e >> 1);
                            c4 = on ? (kind ? dst4[row & 3][j] : tcoef(4, row & 3, j)) : 0;
                        }
                        v4[kind] |= (uint32_t)(uint8_t)(int8_t)c4 << (8 * e);
                    }
                }
                a8[stage][r][lane] = v8;
                a4[0][stage][r][lane] = v4[0];
                a4[1][stage][r][lane] = v4[1];
            }
    // ---- inverse tables (see the kernels below for the index maps)
    static uint32_t ia32[2][4][32], ib32[4][2][32], ia16[2][32], ib16[2][32], ia8[2][32], ib8[32], ia4[2][2][32], ib4[2][32];
    for (int lane = 0; lane < 32; lane++)
    {
        int g = lane >> 2, t = lane & 3;
        for (int mt = 0; mt < 2; mt++)
            for (int r = 0; r < 4; r++)
            {
                int jp = mt * 16 + g + (r & 1) * 8, half = r >> 1;
                uint32_t v = 0;
                for (int e = 0; e < 4; e++) v |= (uint32_t)(uint8_t)(int8_t)tcoef(32, (half * 4 + e) * 4 + t, jp) << (8 * e);
                ia32[mt][r][lane] = v;
            }
        for (int it = 0; it < 4; it++)
            for (int half = 0; half < 2; half++)
            {
                uint32_t v = 0;
                for (int e = 0; e < 4; e++) v |= (uint32_t)(uint8_t)(int8_t)tcoef(32, half * 16 + 4 * t + 2 * (e & 1) + (e >> 1), it * 8 + g) << (8 * e);
                ib32[it][half][lane] = v;
            }
        for (int r = 0; r < 2; r++)
        {
            uint32_t v = 0;
            for (int e = 0; e < 4; e++) v |= (uint32_t)(uint8_t)(int8_t)tcoef(16, e * 4 + t, g + r * 8) << (8 * e);
            ia16[r][lane] = v;
        }
        for (int it = 0; it < 2; it++)
        {
            uint32_t v = 0;
            for (int e = 0; e < 4; e++) v |= (uint32_t)(uint8_t)(int8_t)tcoef(16, 4 * t + 2 * (e & 1) + (e >> 1), it * 8 + g) << (8 * e);
            ib16[it][lane] = v;
        }
        // N = 8: stage-1 A rows (tu = r, j' = g), slot 4t+e = (tu = t >> 1, k = 2e + (t & 1)); stage-2 B slot 4t+e = k' = 2t + e (e < 2)
        for (int r = 0; r < 2; r++)
        {
            uint32_t v = 0;
            for (int e = 0; e < 4; e++) v |= (uint32_t)(uint8_t)(int8_t)((t >> 1) == r ? tcoef(8, 2 * e + (t & 1), g) : 0) << (8 * e);
            ia8[r][lane] = v;
        }
        {
            uint32_t v = 0;
            for (int e = 0; e < 2; e++) v |= (uint32_t)(uint8_t)(int8_t)tcoef(8, 2 * t + e, g) << (8 * e);
            ib8[lane] = v;
        }
        // N = 4: stage-1 A rows (tu_r = (g >> 2) + 2r, j' = g & 3), slot 4t+e = (tu_s = t, k = e);
        //        stage-2 B columns n = g = (set' = g >> 2, i' = g & 3), slot 4t+e = (set = t >> 1, k' = 2 (t & 1) + e) for e < 2
        for (int kind = 0; kind < 2; kind++)
        {
            for (int r = 0; r < 2; r++)
            {
                uint32_t v = 0;
                for (int e = 0; e < 4; e++)
                {
                    int c = ((g >> 2) + 2 * r) == t ? (kind ? dst4[e][g & 3] : tcoef(4, e, g & 3)) : 0;
                    v |= (uint32_t)(uint8_t)(int8_t)c << (8 * e);
                }
                ia4[kind][r][lane] = v;
            }
            uint32_t v = 0;
            for (int e = 0; e < 2; e++)
            {
                int kp = 2 * (t & 1) + e;
                int c = (t >> 1) == (g >> 2) ? (kind ? dst4[kp][g & 3] : tcoef(4, kp, g & 3)) : 0;
                v |= (uint32_t)(uint8_t)(int8_t)c << (8 * e);
            }
            ib4[kind][lane] = v;
        }
    }
    {   // tcgen05 B tiles (tu_umma.cuh): canonical K-major, no swizzle: byte (n, k) at (n / 8) * 256 + (k / 16) * 128 + (n % 8) * 16 + k % 16
        static uint8_t ub[2][2][1024];
        memset(ub, 0, sizeof(ub));
        for (int s = 0; s < 2; s++)
        {
            const int N = s ? 16 : 32;
            for (int nn = 0; nn < N; nn++)
                for (int k = 0; k < N; k++)
                {
                    const int o = (nn >> 3) * 256 + (k >> 4) * 128 + (nn & 7) * 16 + (k & 15);
                    ub[s][0][o] = (uint8_t)(int8_t)tcoef(N, nn, k);        // forward: B[n][k] = T[n][k]
                    ub[s][1][o] = (uint8_t)(int8_t)tcoef(N, k, nn);        // inverse: B[n][k] = T[k][n]
                }
        }
        B200_CUDA(ctx, cudaMemcpyToSymbol(c_ummaB, ub, sizeof(ub)));
        // forward stage 2: A = diag(T, T, ..) over the CTA's 128 rows, K = 128 in four steps of 32.  Tile ks (128 x 32, K-major canonical: byte
        // (m, kk) at (m / 8) * 256 + (kk / 16) * 128 + (m % 8) * 16 + kk % 16) is non-zero only in rows 32 ks .. 32 ks + 31, and that 32 x 32 block
        // is the same for every ks: [T] for N = 32, [[T16, 0], [0, T16]] for N = 16.  Stored once, 12 zero row groups before and after it.
        static uint8_t ad[2][7168];
        memset(ad, 0, sizeof(ad));
        for (int s = 0; s < 2; s++)
        {
            const int N = s ? 16 : 32;
            for (int m = 0; m < 32; m++)
                for (int kk = 0; kk < 32; kk++)
                    if (m / N == kk / N)
                        ad[s][(12 + (m >> 3)) * 256 + (kk >> 4) * 128 + (m & 7) * 16 + (kk & 15)] = (uint8_t)(int8_t)tcoef(N, m % N, kk % N);
        }
        B200_CUDA(ctx, cudaMemcpyToSymbol(c_ummaAD, ad, sizeof(ad)));
    }
    B200_CUDA(ctx, cudaMemcpyToSymbol(c_IA32, ia32, sizeof(ia32)));
    B200_CUDA(ctx, cudaMemcpyToSymbol(c_IB32, ib32, sizeof(ib32)));
    B200_CUDA(ctx, cudaMemcpyToSymbol(c_IA16, ia16, sizeof(ia16)));
    B200_CUDA(ctx, cudaMemcpyToSymbol(c_IB16, ib16, sizeof(ib16)));
    B200_CUDA(ctx, cudaMemcpyToSymbol(c_IA8, ia8, sizeof(ia8)));
    B200_CUDA(ctx, cudaMemcpyToSymbol(c_IB8, ib8, sizeof(ib8)));
    B200_CUDA(ctx, cudaMemcpyToSymbol(c_IA4, ia4, sizeof(ia4)));
    B200_CUDA(ctx, cudaMemcpyToSymbol(c_IB4, ib4, sizeof(ib4)));
    B200_CUDA(ctx, cudaMemcpyToSymbol(c_A8, a8, sizeof(a8)));
    B200_CUDA(ctx, cudaMemcpyToSymbol(c_A4, a4, sizeof(a4)));
    B200_CUDA(ctx, cudaMemcpyToSymbol(c_A32, a32, sizeof(a32)));
    B200_CUDA(ctx, cudaMemcpyToSymbol(c_A16, a16, sizeof(a16)));
    return X265B200_OK;
}

__device__ __forceinline__ void imma32_ss(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void imma32_su(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void imma16_ss(int (&c)[4], const uint32_t (&a)[2], uint32_t b0)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(b0));
}
__device__ __forceinline__ void imma16_su(int (&c)[4], const uint32_t (&a)[2], uint32_t b0)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(b0));
}

// element offset of TU `tu`: explicit descriptor, or contiguous TUs when no descriptor array is given
__device__ __forceinline__ size_t tu_offset(const int32_t* off, int tu, int nn) { return off ? (size_t)off[tu] : (size_t)tu * nn; }

__device__ __forceinline__ void imma32_us(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void imma16_us(int (&c)[4], const uint32_t (&a)[2], uint32_t b0)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(b0));
}

// four consecutive int16 as two packed words, with the widest load the address allows
// (warp-uniform: all lanes of a TU share the alignment class when srcStride % 4 == 0)
__device__ __forceinline__ uint2 ldg_s16x4(const int16_t* p)
{
    uintptr_t a = (uintptr_t)p;
    if ((a & 7) == 0) return __ldg((const uint2*)p);
    if ((a & 3) == 0) return make_uint2(__ldg((const uint32_t*)p), __ldg((const uint32_t*)p + 1));
    uint32_t m = __ldg((const uint32_t*)(p + 1));
    uint32_t x = (uint32_t)(uint16_t)__ldg(p) | (m << 16);
    uint32_t y = (m >> 16) | ((uint32_t)(uint16_t)__ldg(p + 3) << 16);
    return make_uint2(x, y);
}

// four int16 (two packed words) -> their low bytes / high bytes
__device__ __forceinline__ void split4(uint2 x, uint32_t& lo, uint32_t& hi)
{
    lo = __byte_perm(x.x, x.y, 0x6420);
    hi = __byte_perm(x.x, x.y, 0x7531);
}
// bytes 0 / bytes 1 of four int32 values (the low/high byte of their int16 truncation)
__device__ __forceinline__ void pack4(int v0, int v1, int v2, int v3, uint32_t& lo, uint32_t& hi)
{
    uint32_t u01 = __byte_perm(v0, v1, 0x5140), u23 = __byte_perm(v2, v3, 0x5140);
    lo = __byte_perm(u01, u23, 0x5410);
    hi = __byte_perm(u01, u23, 0x7632);
}
__device__ __forceinline__ int recombine(int hi, int lo, int shift) { return ((hi << 8) + lo) >> shift; }

// persistent warps: each warp keeps the 16 A-fragment registers of both stages resident and walks
// 32x32 TUs with a grid stride; the next TU's rows are requested before the current one is transformed.
template<int MINB>
__global__ void __launch_bounds__(128, MINB)
dct32_imma_kernel(const int16_t* __restrict__ src, intptr_t srcStride, const int32_t* __restrict__ off, int n,
                  int16_t* __restrict__ dst, int shift1, int shift2)
{
    int lane = threadIdx.x & 31;
    int warp = blockIdx.x * 4 + (threadIdx.x >> 5);
    int nwarps = gridDim.x * 4;
    if (warp >= n) return;
    int g = lane >> 2, t = lane & 3;
    uint32_t a1[2][4], a2[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int r = 0; r < 4; r++) { a1[mt][r] = c_A32[0][mt][r][lane]; a2[mt][r] = c_A32[1][mt][r][lane]; }
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);
    intptr_t lane_off = (intptr_t)g * srcStride + 4 * t;

    uint2 x[4][2];
    {
        const int16_t* p = src + tu_offset(off, warp, 1024) + lane_off;
#pragma unroll
        for (int jt = 0; jt < 4; jt++)
        {
            const int16_t* q = p + (intptr_t)(jt * 8) * srcStride;
            x[jt][0] = ldg_s16x4(q); x[jt][1] = ldg_s16x4(q + 16);
        }
    }
    for (int tu = warp; tu < n; tu += nwarps)
    {
        // stage-1 B operand straight from the row-major block: n-tile jt = rows jt*8 + g
        uint32_t blo[4][2], bhi[4][2];
#pragma unroll
        for (int jt = 0; jt < 4; jt++)
        {
            split4(x[jt][0], blo[jt][0], bhi[jt][0]);
            split4(x[jt][1], blo[jt][1], bhi[jt][1]);
        }
        int nxt = tu + nwarps;
        if (nxt < n)
        {
            const int16_t* p = src + tu_offset(off, nxt, 1024) + lane_off;
#pragma unroll
            for (int jt = 0; jt < 4; jt++)
            {
                const int16_t* q = p + (intptr_t)(jt * 8) * srcStride;
                x[jt][0] = ldg_s16x4(q); x[jt][1] = ldg_s16x4(q + 16);
            }
        }
        uint32_t b2lo[4][2], b2hi[4][2];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
        {
            int v[4][4];
#pragma unroll
            for (int jt = 0; jt < 4; jt++)
            {
                int chi[4] = { 0, 0, 0, 0 }, clo[4] = { add1, add1, add1, add1 };
                imma32_ss(chi, a1[mt], bhi[jt][0], bhi[jt][1]);
                imma32_su(clo, a1[mt], blo[jt][0], blo[jt][1]);
#pragma unroll
                for (int r = 0; r < 4; r++) v[jt][r] = recombine(chi[r], clo[r], shift1);
            }
            // accumulator fragment -> stage-2 B fragments: rows g (c0,c1) feed n-tile 2mt, rows g+8 (c2,c3) feed 2mt+1
            pack4(v[0][0], v[0][1], v[1][0], v[1][1], b2lo[2 * mt][0], b2hi[2 * mt][0]);
            pack4(v[2][0], v[2][1], v[3][0], v[3][1], b2lo[2 * mt][1], b2hi[2 * mt][1]);
            pack4(v[0][2], v[0][3], v[1][2], v[1][3], b2lo[2 * mt + 1][0], b2hi[2 * mt + 1][0]);
            pack4(v[2][2], v[2][3], v[3][2], v[3][3], b2lo[2 * mt + 1][1], b2hi[2 * mt + 1][1]);
        }
        int16_t* o = dst + (size_t)tu * 1024 + g * 32 + 2 * t;
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
        {
#pragma unroll
            for (int nt = 0; nt < 4; nt++)
            {
                int chi[4] = { 0, 0, 0, 0 }, clo[4] = { add2, add2, add2, add2 };
                imma32_ss(chi, a2[mt], b2hi[nt][0], b2hi[nt][1]);
                imma32_su(clo, a2[mt], b2lo[nt][0], b2lo[nt][1]);
                int v0 = recombine(chi[0], clo[0], shift2), v1 = recombine(chi[1], clo[1], shift2);
                int v2 = recombine(chi[2], clo[2], shift2), v3 = recombine(chi[3], clo[3], shift2);
                *(uint32_t*)(o + (mt * 16) * 32 + nt * 8) = __byte_perm(v0, v1, 0x5410);
                *(uint32_t*)(o + (mt * 16 + 8) * 32 + nt * 8) = __byte_perm(v2, v3, 0x5410);
            }
        }
    }
}

// persistent warps over groups of TPW 16x16 TUs (TPW TUs are loaded together so their rows are in
// flight at the same time; the next group is requested before the current one is transformed)
template<int TPW>
__global__ void __launch_bounds__(128)
dct16_imma_kernel(const int16_t* __restrict__ src, intptr_t srcStride, const int32_t* __restrict__ off, int n,
                  int16_t* __restrict__ dst, int shift1, int shift2)
{
    int lane = threadIdx.x & 31;
    int warp = blockIdx.x * 4 + (threadIdx.x >> 5);
    int nwarps = gridDim.x * 4;
    int ngroups = (n + TPW - 1) / TPW;
    if (warp >= ngroups) return;
    int g = lane >> 2, t = lane & 3;
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);
    uint32_t a1[2] = { c_A16[0][0][lane], c_A16[0][1][lane] };
    uint32_t a2[2] = { c_A16[1][0][lane], c_A16[1][1][lane] };
    intptr_t lane_off = (intptr_t)g * srcStride + 4 * t;
    uint2 x[TPW][2];
#pragma unroll
    for (int u = 0; u < TPW; u++)
    {
        const int16_t* p = src + tu_offset(off, min(warp * TPW + u, n - 1), 256) + lane_off;
        x[u][0] = ldg_s16x4(p); x[u][1] = ldg_s16x4(p + 8 * srcStride);
    }
    for (int grp = warp; grp < ngroups; grp += nwarps)
    {
        uint32_t blo[TPW][2], bhi[TPW][2];
#pragma unroll
        for (int u = 0; u < TPW; u++)
        {
            split4(x[u][0], blo[u][0], bhi[u][0]);
            split4(x[u][1], blo[u][1], bhi[u][1]);
        }
        int nxt = grp + nwarps;
        if (nxt < ngroups)
        {
#pragma unroll
            for (int u = 0; u < TPW; u++)
            {
                const int16_t* p = src + tu_offset(off, min(nxt * TPW + u, n - 1), 256) + lane_off;
                x[u][0] = ldg_s16x4(p); x[u][1] = ldg_s16x4(p + 8 * srcStride);
            }
        }
#pragma unroll
        for (int u = 0; u < TPW; u++)
        {
            int tu = grp * TPW + u;
            if (tu >= n) break;
            int v[2][4];
#pragma unroll
            for (int jt = 0; jt < 2; jt++)
            {
                int chi[4] = { 0, 0, 0, 0 }, clo[4] = { add1, add1, add1, add1 };
                imma16_ss(chi, a1, bhi[u][jt]);
                imma16_su(clo, a1, blo[u][jt]);
#pragma unroll
                for (int r = 0; r < 4; r++) v[jt][r] = recombine(chi[r], clo[r], shift1);
            }
            uint32_t b2lo[2], b2hi[2];
            pack4(v[0][0], v[0][1], v[1][0], v[1][1], b2lo[0], b2hi[0]);
            pack4(v[0][2], v[0][3], v[1][2], v[1][3], b2lo[1], b2hi[1]);
            int16_t* o = dst + (size_t)tu * 256 + g * 16 + 2 * t;
#pragma unroll
            for (int nt = 0; nt < 2; nt++)
            {
                int chi[4] = { 0, 0, 0, 0 }, clo[4] = { add2, add2, add2, add2 };
                imma16_ss(chi, a2, b2hi[nt]);
                imma16_su(clo, a2, b2lo[nt]);
                int v0 = recombine(chi[0], clo[0], shift2), v1 = recombine(chi[1], clo[1], shift2);
                int v2 = recombine(chi[2], clo[2], shift2), v3 = recombine(chi[3], clo[3], shift2);
                *(uint32_t*)(o + nt * 8) = __byte_perm(v0, v1, 0x5410);
                *(uint32_t*)(o + 8 * 16 + nt * 8) = __byte_perm(v2, v3, 0x5410);
            }
        }
    }
}

// Small transforms: block-diagonal A packs two 8x8 TUs (SMALL = 8) or eight 4x4 TUs (SMALL = 4) into one
// m16n8k16 IMMA per stage and byte plane.  One MMA group = 128 coefficients = 4 per lane: an 8-byte load
// and two 4-byte stores per lane, all fully coalesced (a warp reads/writes 256 contiguous bytes when the
// TUs are contiguous).  UN groups are processed together to keep UN * 256 B per warp in flight.
template<int SMALL, int UN>
__global__ void __launch_bounds__(128)
dct_small_imma_kernel(const int16_t* __restrict__ src, intptr_t srcStride, const int32_t* __restrict__ off, int n,
                      int16_t* __restrict__ dst, int shift1, int shift2, int kind)
{
    constexpr int TPG = SMALL == 8 ? 2 : 8;                 // TUs per MMA group
    constexpr int NN = SMALL * SMALL;
    int lane = threadIdx.x & 31;
    int warp = blockIdx.x * 4 + (threadIdx.x >> 5);
    int nwarps = gridDim.x * 4;
    int ngroups = (n + TPG * UN - 1) / (TPG * UN);          // groups of UN MMA groups
    if (warp >= ngroups) return;
    int g = lane >> 2, t = lane & 3;
    const uint32_t (*A)[2][32] = SMALL == 8 ? c_A8 : c_A4[kind];
    uint32_t a1[2] = { A[0][0][lane], A[0][1][lane] };
    uint32_t a2[2] = { A[1][0][lane], A[1][1][lane] };
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);
    // which TU of the group this lane loads from, and where inside it
    const int ld_tu = SMALL == 8 ? (t >> 1) : ((g >> 2) * 4 + t);
    const intptr_t ld_off = SMALL == 8 ? (intptr_t)g * srcStride + 4 * (t & 1) : (intptr_t)(g & 3) * srcStride;
    // which TUs the two accumulator halves (rows g / rows g+8) belong to, and where the pair lands
    const int st_tu0 = SMALL == 8 ? 0 : (t >> 1) + 2 * (g >> 2);
    const int st_tu1 = SMALL == 8 ? 1 : 4 + (t >> 1) + 2 * (g >> 2);
    const int st_off = SMALL == 8 ? g * 8 + 2 * t : (g & 3) * 4 + 2 * (t & 1);

    uint2 x[UN];
#pragma unroll
    for (int u = 0; u < UN; u++)
    {
        int tu = min((warp * UN + u) * TPG + ld_tu, n - 1);
        x[u] = ldg_s16x4(src + tu_offset(off, tu, NN) + ld_off);
    }
    for (int grp = warp; grp < ngroups; grp += nwarps)
    {
        uint32_t blo[UN], bhi[UN];
#pragma unroll
        for (int u = 0; u < UN; u++) split4(x[u], blo[u], bhi[u]);
        int nxt = grp + nwarps;
        if (nxt < ngroups)
        {
#pragma unroll
            for (int u = 0; u < UN; u++)
            {
                int tu = min((nxt * UN + u) * TPG + ld_tu, n - 1);
                x[u] = ldg_s16x4(src + tu_offset(off, tu, NN) + ld_off);
            }
        }
#pragma unroll
        for (int u = 0; u < UN; u++)
        {
            int base = (grp * UN + u) * TPG;
            if (base >= n) break;
            int chi[4] = { 0, 0, 0, 0 }, clo[4] = { add1, add1, add1, add1 };
            imma16_ss(chi, a1, bhi[u]);
            imma16_su(clo, a1, blo[u]);
            uint32_t b2lo, b2hi;
            pack4(recombine(chi[0], clo[0], shift1), recombine(chi[1], clo[1], shift1),
                  recombine(chi[2], clo[2], shift1), recombine(chi[3], clo[3], shift1), b2lo, b2hi);
            int dhi[4] = { 0, 0, 0, 0 }, dlo[4] = { add2, add2, add2, add2 };
            imma16_ss(dhi, a2, b2hi);
            imma16_su(dlo, a2, b2lo);
            int v0 = recombine(dhi[0], dlo[0], shift2), v1 = recombine(dhi[1], dlo[1], shift2);
            int v2 = recombine(dhi[2], dlo[2], shift2), v3 = recombine(dhi[3], dlo[3], shift2);
            if (base + st_tu0 < n) *(uint32_t*)(dst + (size_t)(base + st_tu0) * NN + st_off) = __byte_perm(v0, v1, 0x5410);
            if (base + st_tu1 < n) *(uint32_t*)(dst + (size_t)(base + st_tu1) * NN + st_off) = __byte_perm(v2, v3, 0x5410);
        }
    }
}

// ================================================================== inverse transforms
// Reference: out1[j][i] = clip16((sum_k T[k][i] * in[k][j] + 64) >> 7), out2 likewise on out1 with shift 12 - (depth - 8)
// (dct.cpp:242-416,528-611).  With Y1 = stage-1 output:
//     stage 1:  D1[j'][k'] = sum_k  T[k][j'] * in[k][k']       A = T^T (constant), B = in           (D1 = Y1^T)
//     stage 2:  D2[j'][i'] = sum_k' D1[j'][k'] * T[k'][i']     A = D1 (accumulators), B = T (constant)
// so the final accumulator is the row-major output block.  The stage-1 B operand walks DOWN the columns of
// `in`; the contraction order (free) is chosen so that the four lanes of a quad read four consecutive rows, and
// the column order (also free, it only renames the stage-2 contraction slots) so that one 32-bit load feeds
// two n-tiles.
__device__ __forceinline__ int recombine_clip(int hi, int lo, int shift)
{
    return min(32767, max(-32768, ((hi << 8) + lo) >> shift));
}
__device__ __forceinline__ void store_pair(int16_t* p, int v0, int v1)
{
    if (((uintptr_t)p & 3) == 0) *(uint32_t*)p = __byte_perm(v0, v1, 0x5410);
    else { p[0] = (int16_t)v0; p[1] = (int16_t)v1; }
}
// two int32 -> packed int16 pair with signed saturation in one instruction (I2IP.S16.S32.SAT): the inverse transform's
// clip3(-32768, 32767, .) of both stages (dct.cpp:257, 272)
__device__ __forceinline__ uint32_t pack_sat_s16(int lo, int hi)
{
    uint32_t d;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(d) : "r"(hi), "r"(lo));
    return d;
}
__device__ __forceinline__ uint32_t recombine_sat2(int h0, int l0, int h1, int l1, int shift)
{
    return pack_sat_s16(recombine(h0, l0, shift), recombine(h1, l1, shift));
}
// low bytes / high bytes of four saturated values (the stage-2 operand of the inverse transform)
__device__ __forceinline__ void pack4_sat(int v0, int v1, int v2, int v3, uint32_t& lo, uint32_t& hi)
{
    split4(make_uint2(pack_sat_s16(v0, v1), pack_sat_s16(v2, v3)), lo, hi);
}
__device__ __forceinline__ void store_pair_packed(int16_t* p, uint32_t w)
{
    if (((uintptr_t)p & 3) == 0) *(uint32_t*)p = w;
    else { p[0] = (int16_t)(w & 0xffff); p[1] = (int16_t)(w >> 16); }
}

__global__ void __launch_bounds__(128)
idct32_imma_kernel(const int16_t* __restrict__ src, int n, int16_t* __restrict__ dst, intptr_t dstStride,
                   const int32_t* __restrict__ off, int shift1, int shift2)
{
    int lane = threadIdx.x & 31;
    int warp = blockIdx.x * 4 + (threadIdx.x >> 5);
    int nwarps = gridDim.x * 4;
    if (warp >= n) return;
    int g = lane >> 2, t = lane & 3;
    uint32_t a1[2][4], b2[4][2];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int r = 0; r < 4; r++) a1[mt][r] = c_IA32[mt][r][lane];
#pragma unroll
    for (int it = 0; it < 4; it++) { b2[it][0] = c_IB32[it][0][lane]; b2[it][1] = c_IB32[it][1][lane]; }
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);

    // B1 gather: for (pair p, half, e): in[(half*4+e)*4 + t][p*16 + 2g .. +1] (one 32-bit load = n-tiles 2p and 2p+1)
    uint32_t x[2][2][4];
    {
        const int16_t* q = src + (size_t)warp * 1024 + t * 32 + 2 * g;
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int half = 0; half < 2; half++)
#pragma unroll
                for (int e = 0; e < 4; e++) x[p][half][e] = __ldg((const uint32_t*)(q + ((half * 4 + e) * 4) * 32 + p * 16));
    }
    for (int tu = warp; tu < n; tu += nwarps)
    {
        uint32_t blo[4][2], bhi[4][2];
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int half = 0; half < 2; half++)
            {
                // even columns (low halves) -> n-tile 2p, odd columns (high halves) -> n-tile 2p+1
                uint32_t e01 = __byte_perm(x[p][half][0], x[p][half][1], 0x5140), e23 = __byte_perm(x[p][half][2], x[p][half][3], 0x5140);
                uint32_t o01 = __byte_perm(x[p][half][0], x[p][half][1], 0x7362), o23 = __byte_perm(x[p][half][2], x[p][half][3], 0x7362);
                blo[2 * p][half] = __byte_perm(e01, e23, 0x5410); bhi[2 * p][half] = __byte_perm(e01, e23, 0x7632);
                blo[2 * p + 1][half] = __byte_perm(o01, o23, 0x5410); bhi[2 * p + 1][half] = __byte_perm(o01, o23, 0x7632);
            }
        int nxt = tu + nwarps;
        if (nxt < n)
        {
            const int16_t* q = src + (size_t)nxt * 1024 + t * 32 + 2 * g;
#pragma unroll
            for (int p = 0; p < 2; p++)
#pragma unroll
                for (int half = 0; half < 2; half++)
#pragma unroll
                    for (int e = 0; e < 4; e++) x[p][half][e] = __ldg((const uint32_t*)(q + ((half * 4 + e) * 4) * 32 + p * 16));
        }
        int16_t* o = dst + tu_offset(off, tu, 1024) + (intptr_t)g * dstStride + 2 * t;
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
        {
            int v[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; nt++)
            {
                int chi[4] = { 0, 0, 0, 0 }, clo[4] = { add1, add1, add1, add1 };
                imma32_ss(chi, a1[mt], bhi[nt][0], bhi[nt][1]);
                imma32_su(clo, a1[mt], blo[nt][0], blo[nt][1]);
#pragma unroll
                for (int r = 0; r < 4; r++) v[nt][r] = recombine(chi[r], clo[r], shift1);
            }
            // accumulators -> stage-2 A fragments (rows g: c0,c1; rows g+8: c2,c3; n-tiles 0,1 -> k 0..15, 2,3 -> 16..31)
            uint32_t alo[4], ahi[4];
            pack4_sat(v[0][0], v[0][1], v[1][0], v[1][1], alo[0], ahi[0]);
            pack4_sat(v[0][2], v[0][3], v[1][2], v[1][3], alo[1], ahi[1]);
            pack4_sat(v[2][0], v[2][1], v[3][0], v[3][1], alo[2], ahi[2]);
            pack4_sat(v[2][2], v[2][3], v[3][2], v[3][3], alo[3], ahi[3]);
#pragma unroll
            for (int it = 0; it < 4; it++)
            {
                int dhi[4] = { 0, 0, 0, 0 }, dlo[4] = { add2, add2, add2, add2 };
                imma32_ss(dhi, ahi, b2[it][0], b2[it][1]);
                imma32_us(dlo, alo, b2[it][0], b2[it][1]);
                store_pair_packed(o + (intptr_t)(mt * 16) * dstStride + it * 8, recombine_sat2(dhi[0], dlo[0], dhi[1], dlo[1], shift2));
                store_pair_packed(o + (intptr_t)(mt * 16 + 8) * dstStride + it * 8, recombine_sat2(dhi[2], dlo[2], dhi[3], dlo[3], shift2));
            }
        }
    }
}

template<int TPW>
__global__ void __launch_bounds__(128)
idct16_imma_kernel(const int16_t* __restrict__ src, int n, int16_t* __restrict__ dst, intptr_t dstStride,
                   const int32_t* __restrict__ off, int shift1, int shift2)
{
    int lane = threadIdx.x & 31;
    int warp = blockIdx.x * 4 + (threadIdx.x >> 5);
    int nwarps = gridDim.x * 4;
    int ngroups = (n + TPW - 1) / TPW;
    if (warp >= ngroups) return;
    int g = lane >> 2, t = lane & 3;
    uint32_t a1[2] = { c_IA16[0][lane], c_IA16[1][lane] };
    uint32_t b2[2] = { c_IB16[0][lane], c_IB16[1][lane] };
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);
    uint32_t x[TPW][4];
#pragma unroll
    for (int u = 0; u < TPW; u++)
    {
        const int16_t* q = src + (size_t)min(warp * TPW + u, n - 1) * 256 + t * 16 + 2 * g;
#pragma unroll
        for (int e = 0; e < 4; e++) x[u][e] = __ldg((const uint32_t*)(q + e * 4 * 16));     // in[e*4 + t][2g .. 2g+1]
    }
    for (int grp = warp; grp < ngroups; grp += nwarps)
    {
        uint32_t blo[TPW][2], bhi[TPW][2];
#pragma unroll
        for (int u = 0; u < TPW; u++)
        {
            uint32_t e01 = __byte_perm(x[u][0], x[u][1], 0x5140), e23 = __byte_perm(x[u][2], x[u][3], 0x5140);
            uint32_t o01 = __byte_perm(x[u][0], x[u][1], 0x7362), o23 = __byte_perm(x[u][2], x[u][3], 0x7362);
            blo[u][0] = __byte_perm(e01, e23, 0x5410); bhi[u][0] = __byte_perm(e01, e23, 0x7632);
            blo[u][1] = __byte_perm(o01, o23, 0x5410); bhi[u][1] = __byte_perm(o01, o23, 0x7632);
        }
        int nxt = grp + nwarps;
        if (nxt < ngroups)
        {
#pragma unroll
            for (int u = 0; u < TPW; u++)
            {
                const int16_t* q = src + (size_t)min(nxt * TPW + u, n - 1) * 256 + t * 16 + 2 * g;
#pragma unroll
                for (int e = 0; e < 4; e++) x[u][e] = __ldg((const uint32_t*)(q + e * 4 * 16));
            }
        }
#pragma unroll
        for (int u = 0; u < TPW; u++)
        {
            int tu = grp * TPW + u;
            if (tu >= n) break;
            int v[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; nt++)
            {
                int chi[4] = { 0, 0, 0, 0 }, clo[4] = { add1, add1, add1, add1 };
                imma16_ss(chi, a1, bhi[u][nt]);
                imma16_su(clo, a1, blo[u][nt]);
#pragma unroll
                for (int r = 0; r < 4; r++) v[nt][r] = recombine(chi[r], clo[r], shift1);
            }
            uint32_t alo[2], ahi[2];
            pack4_sat(v[0][0], v[0][1], v[1][0], v[1][1], alo[0], ahi[0]);
            pack4_sat(v[0][2], v[0][3], v[1][2], v[1][3], alo[1], ahi[1]);
            int16_t* o = dst + tu_offset(off, tu, 256) + (intptr_t)g * dstStride + 2 * t;
#pragma unroll
            for (int it = 0; it < 2; it++)
            {
                int dhi[4] = { 0, 0, 0, 0 }, dlo[4] = { add2, add2, add2, add2 };
                imma16_ss(dhi, ahi, b2[it]);
                imma16_us(dlo, alo, b2[it]);
                store_pair_packed(o + it * 8, recombine_sat2(dhi[0], dlo[0], dhi[1], dlo[1], shift2));
                store_pair_packed(o + (intptr_t)8 * dstStride + it * 8, recombine_sat2(dhi[2], dlo[2], dhi[3], dlo[3], shift2));
            }
        }
    }
}

// 4x4 transpose of int16 across the four lanes {l, l^4, l^8, l^12} (row index r = (lane >> 2) & 3): lane r passes
// its row (a_r0..a_r3 as two packed words) and returns column r (a_0r..a_3r).  Two shuffles, four byte permutes.
__device__ __forceinline__ uint2 transpose4x4_s16(uint2 v, int r)
{
    // exchange with r ^ 2: afterwards the lane holds the 2x2 block rows {r & 1, (r & 1) + 2} x cols {2 (r >> 1), +1}
    uint32_t recv = __shfl_xor_sync(0xffffffffu, (r & 2) ? v.x : v.y, 8);
    uint32_t p0 = (r & 2) ? recv : v.x, p1 = (r & 2) ? v.y : recv;
    // exchange with r ^ 1: even lanes keep the low halves (their column) and pass the high ones, odd lanes the reverse
    uint32_t lows = __byte_perm(p0, p1, 0x5410), highs = __byte_perm(p0, p1, 0x7632);
    uint32_t got = __shfl_xor_sync(0xffffffffu, (r & 1) ? lows : highs, 4);
    uint32_t ev = (r & 1) ? got : lows, od = (r & 1) ? highs : got;      // (row 0, row 2) and (row 1, row 3) of the column
    return make_uint2(__byte_perm(ev, od, 0x5410), __byte_perm(ev, od, 0x7632));
}

// two 8x8 TUs (SMALL = 8) or eight 4x4 TUs (SMALL = 4) per MMA group, block-diagonal constant operands
template<int SMALL, int UN>
__global__ void __launch_bounds__(128)
idct_small_imma_kernel(const int16_t* __restrict__ src, int n, int16_t* __restrict__ dst, intptr_t dstStride,
                       const int32_t* __restrict__ off, int shift1, int shift2, int kind)
{
    constexpr int TPG = SMALL == 8 ? 2 : 8;
    constexpr int NN = SMALL * SMALL;
    int lane = threadIdx.x & 31;
    int warp = blockIdx.x * 4 + (threadIdx.x >> 5);
    int nwarps = gridDim.x * 4;
    int ngroups = (n + TPG * UN - 1) / (TPG * UN);
    if (warp >= ngroups) return;
    int g = lane >> 2, t = lane & 3;
    uint32_t a1[2], b2;
    if (SMALL == 8) { a1[0] = c_IA8[0][lane]; a1[1] = c_IA8[1][lane]; b2 = c_IB8[lane]; }
    else { a1[0] = c_IA4[kind][0][lane]; a1[1] = c_IA4[kind][1][lane]; b2 = c_IB4[kind][lane]; }
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);
    // stage-1 B operand, slot 4t+e.  N = 8: TU t>>1, in[2e + (t&1)][g].  N = 4: TU (g>>2)*4 + t, in[e][g&3].
    // Each lane fetches ONE 8-byte row piece -- N = 8: row 2r + (t&1), columns 4(g>>2)..+3 of TU t>>1; N = 4: row r of
    // TU (g>>2)*4 + t, with r = g & 3 -- so a warp reads the group's 256 contiguous bytes with one instruction, and
    // the lanes {g&3 = 0..3} transpose their 4x4 among themselves (transpose4x4_s16) to get the column they need.
    const int ld_tu = SMALL == 8 ? (t >> 1) : ((g >> 2) * 4 + t);
    const int ld_off = SMALL == 8 ? (2 * (g & 3) + (t & 1)) * 8 + 4 * (g >> 2) : (g & 3) * 4;
    // output: accumulator rows g / g+8.  N = 8: TU 0 / 1, out[g][2t..].  N = 4: TU (t>>1)*4 + (g>>2) (+2), out[g&3][2(t&1)..]
    const int st_tu0 = SMALL == 8 ? 0 : (t >> 1) * 4 + (g >> 2);
    const int st_tu1 = SMALL == 8 ? 1 : (t >> 1) * 4 + (g >> 2) + 2;
    const int st_row = SMALL == 8 ? g : (g & 3);
    const int st_col = SMALL == 8 ? 2 * t : 2 * (t & 1);

    uint2 x[UN];
#pragma unroll
    for (int u = 0; u < UN; u++)
        x[u] = __ldg((const uint2*)(src + (size_t)min((warp * UN + u) * TPG + ld_tu, n - 1) * NN + ld_off));
    for (int grp = warp; grp < ngroups; grp += nwarps)
    {
        uint32_t blo[UN], bhi[UN];
#pragma unroll
        for (int u = 0; u < UN; u++) split4(transpose4x4_s16(x[u], g & 3), blo[u], bhi[u]);
        int nxt = grp + nwarps;
        if (nxt < ngroups)
        {
#pragma unroll
            for (int u = 0; u < UN; u++)
                x[u] = __ldg((const uint2*)(src + (size_t)min((nxt * UN + u) * TPG + ld_tu, n - 1) * NN + ld_off));
        }
#pragma unroll
        for (int u = 0; u < UN; u++)
        {
            int base = (grp * UN + u) * TPG;
            if (base >= n) break;
            int chi[4] = { 0, 0, 0, 0 }, clo[4] = { add1, add1, add1, add1 };
            imma16_ss(chi, a1, bhi[u]);
            imma16_su(clo, a1, blo[u]);
            uint32_t alo[2], ahi[2];
            split4(make_uint2(recombine_sat2(chi[0], clo[0], chi[1], clo[1], shift1), 0u), alo[0], ahi[0]);   // slots e = 2,3 are padding (their B rows are zero)
            split4(make_uint2(recombine_sat2(chi[2], clo[2], chi[3], clo[3], shift1), 0u), alo[1], ahi[1]);
            int dhi[4] = { 0, 0, 0, 0 }, dlo[4] = { add2, add2, add2, add2 };
            imma16_ss(dhi, ahi, b2);
            imma16_us(dlo, alo, b2);
            if (base + st_tu0 < n)
                store_pair_packed(dst + tu_offset(off, base + st_tu0, NN) + (intptr_t)st_row * dstStride + st_col,
                                  recombine_sat2(dhi[0], dlo[0], dhi[1], dlo[1], shift2));
            if (base + st_tu1 < n)
                store_pair_packed(dst + tu_offset(off, base + st_tu1, NN) + (intptr_t)st_row * dstStride + st_col,
                                  recombine_sat2(dhi[2], dlo[2], dhi[3], dlo[3], shift2));
        }
    }
}

// Persistent kernels: exactly one resident wave.  With sms * 8 CTAs a kernel that fits 5 CTAs per SM ran 1.6 waves (the
// second one 60 % full): dct32 0.198 -> 0.174 ms once the grid matched the residency.
template<typename K>
static int persistent_grid(K kernel, int sms, int threads = 128)
{
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, 0) != cudaSuccess || occ < 1) occ = 4;
    return sms * occ;
}
#define PGRID(kernel) ([&]() { static int g = persistent_grid(kernel, sms); return g; }())
// whole waves per launch, measured (profiles/r2_persistent_grid.md): 1 for the 32-point kernels, 4 for the 16-point and small ones
constexpr int WAVES_16 = 4, WAVES_SMALL = 4;

bool launch_idct_imma(x265b200_ctx* ctx, int N, const int16_t* src, int n, int16_t* dst, intptr_t dstStride,
                      const int32_t* off, int shift1, int shift2, cudaStream_t st, int dst4)
{
    if ((uintptr_t)src & 7) return false;
    int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
    int grid;
    if (N == 32)
    {
        grid = PGRID(idct32_imma_kernel);
        if (grid > ceil_div(n, 4)) grid = ceil_div(n, 4);
        idct32_imma_kernel<<<grid, 128, 0, st>>>(src, n, dst, dstStride, off, shift1, shift2);
    }
    else if (N == 16)
    {
        int need = ceil_div(ceil_div(n, 4), 4);
        grid = PGRID(idct16_imma_kernel<4>) * WAVES_16;
        if (grid > need) grid = need;
        idct16_imma_kernel<4><<<grid, 128, 0, st>>>(src, n, dst, dstStride, off, shift1, shift2);
    }
    else if (N == 8)
    {
        int need = ceil_div(ceil_div(n, 2 * 4), 4);
        grid = PGRID((idct_small_imma_kernel<8, 4>)) * WAVES_SMALL;
        if (grid > need) grid = need;
        idct_small_imma_kernel<8, 4><<<grid, 128, 0, st>>>(src, n, dst, dstStride, off, shift1, shift2, 0);
    }
    else if (N == 4)
    {
        int need = ceil_div(ceil_div(n, 8 * 4), 4);
        grid = PGRID((idct_small_imma_kernel<4, 4>)) * WAVES_SMALL;
        if (grid > need) grid = need;
        idct_small_imma_kernel<4, 4><<<grid, 128, 0, st>>>(src, n, dst, dstStride, off, shift1, shift2, dst4 ? 1 : 0);
    }
    else
        return false;
    return true;
}

// returns true if it handled the request
bool launch_dct_imma(x265b200_ctx* ctx, int N, const int16_t* src, intptr_t srcStride, const int32_t* off, int n,
                     int16_t* dst, int shift1, int shift2, cudaStream_t st, int dst4)
{
    if ((uintptr_t)dst & 3) return false;
    // persistent grid: a multiple of the SM count, capped by the work available
    int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
    if (N == 32)
    {
        int grid = PGRID(dct32_imma_kernel<5>);
        if (grid > ceil_div(n, 4)) grid = ceil_div(n, 4);
        // 5 resident CTAs per SM (96 registers, no spills): 6 / 7 / 8 spill and measured 10 / 23 / 27 % slower
        dct32_imma_kernel<5><<<grid, 128, 0, st>>>(src, srcStride, off, n, dst, shift1, shift2);
    }
    else if (N == 16)
    {
        int grid = PGRID(dct16_imma_kernel<4>) * WAVES_16;
        if (grid > ceil_div(ceil_div(n, 4), 4)) grid = ceil_div(ceil_div(n, 4), 4);
        dct16_imma_kernel<4><<<grid, 128, 0, st>>>(src, srcStride, off, n, dst, shift1, shift2);
    }
    else if (N == 8)
    {
        int grid = PGRID((dct_small_imma_kernel<8, 4>)) * WAVES_SMALL;
        int need = ceil_div(ceil_div(n, 2 * 4), 4);
        if (grid > need) grid = need;
        dct_small_imma_kernel<8, 4><<<grid, 128, 0, st>>>(src, srcStride, off, n, dst, shift1, shift2, 0);
    }
    else if (N == 4)
    {
        int grid = PGRID((dct_small_imma_kernel<4, 4>)) * WAVES_SMALL;
        int need = ceil_div(ceil_div(n, 8 * 4), 4);
        if (grid > need) grid = need;
        dct_small_imma_kernel<4, 4><<<grid, 128, 0, st>>>(src, srcStride, off, n, dst, shift1, shift2, dst4 ? 1 : 0);
    }
    else
        return false;
    return true;
}

#include "tu_fused.cuh"
#include "tu_umma.cuh"

} // namespace b200
