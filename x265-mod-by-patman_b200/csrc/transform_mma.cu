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
#include "internal.h"

namespace b200 {

// per-lane A fragments, [stage][mtile][reg][lane]; N = 32: 2 x 2 x 4 x 32, N = 16: 2 x 1 x 2 x 32.
// Global (not __constant__) memory: the index is the lane, and per-lane constant-bank reads serialise.
__device__ uint32_t c_A32[2][2][4][32];
__device__ uint32_t c_A16[2][2][32];
// N = 8 (two TUs per MMA) and N = 4 (eight TUs per MMA; [1] = DST-VII): block-diagonal A, [kind][stage][reg][lane]
__device__ uint32_t c_A8[2][2][32];
__device__ uint32_t c_A4[2][2][2][32];

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
__global__ void __launch_bounds__(128)
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

// returns true if it handled the request
bool launch_dct_imma(x265b200_ctx* ctx, int N, const int16_t* src, intptr_t srcStride, const int32_t* off, int n,
                     int16_t* dst, int shift1, int shift2, cudaStream_t st, int dst4)
{
    if ((uintptr_t)dst & 3) return false;
    // persistent grid: a multiple of the SM count, capped by the work available
    int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
    if (N == 32)
    {
        int grid = sms * 8;
        if (grid > ceil_div(n, 4)) grid = ceil_div(n, 4);
        dct32_imma_kernel<<<grid, 128, 0, st>>>(src, srcStride, off, n, dst, shift1, shift2);
    }
    else if (N == 16)
    {
        int grid = sms * 8;
        if (grid > ceil_div(ceil_div(n, 4), 4)) grid = ceil_div(ceil_div(n, 4), 4);
        dct16_imma_kernel<4><<<grid, 128, 0, st>>>(src, srcStride, off, n, dst, shift1, shift2);
    }
    else if (N == 8)
    {
        int grid = sms * 8;
        int need = ceil_div(ceil_div(n, 2 * 4), 4);
        if (grid > need) grid = need;
        dct_small_imma_kernel<8, 4><<<grid, 128, 0, st>>>(src, srcStride, off, n, dst, shift1, shift2, 0);
    }
    else if (N == 4)
    {
        int grid = sms * 8;
        int need = ceil_div(ceil_div(n, 8 * 4), 4);
        if (grid > need) grid = need;
        dct_small_imma_kernel<4, 4><<<grid, 128, 0, st>>>(src, srcStride, off, n, dst, shift1, shift2, dst4 ? 1 : 0);
    }
    else
        return false;
    return true;
}

} // namespace b200
