// tile_kernels.cuh -- throughput kernels for SAD / SATD / SSE over pixel planes (shared by pixel.cu
// and the tuning lab tools/satd_lab.cu, which instantiates the variants side by side).
#pragma once
#include "device_util.cuh"

namespace b200 {

enum { OP_SAD = 0, OP_SATD = 1, OP_SA8D = 2, OP_SSE = 3, OP_SSD = 4 };

// configuration used by the library (chosen with tools/satd_lab.cu on B200, see profiles/)
constexpr int FAST_UNROLL = 1;
constexpr int FAST_MINBLK = 1;

// One 4x4 tile: a = fenc rows, b = reference rows, each row as packed 16-bit pairs (lo = samples 0,1; hi = samples 2,3).
//   SAD : |a - b| per lane as max - min (VIMNMX.U16x2 twice; the lane-wise difference of max and min never borrows),
//         lane sums stay packed for the tile (<= 4 * 4095) and are widened once with a dot product against (1, 1).
//   SATD: differences by plain 32-bit subtraction (a borrow from the low lane is undone when the word is unpacked, and
//         word-wise adds keep that representation exact), the vertical 4-point Hadamard and the horizontal stage that
//         pairs the two words of a row on packed lanes (|value| <= 8 * 4095 < 2^15), and the last stage -- it pairs the
//         two lanes of a word -- folded into the magnitude sum: |p + q| + |p - q| = 2 max(|p|, |q|), so the per-tile
//         (sum >> 1) of the reference (pixel.cpp:236-241, raw sums are even) is the plain sum of the maxima.
template<int OP, typename ACC>
__device__ __forceinline__ void tile4_accumulate(const uint32_t (&alo)[4], const uint32_t (&ahi)[4],
                                                 const uint32_t (&blo)[4], const uint32_t (&bhi)[4], ACC& acc)
{
    if constexpr (OP == OP_SAD)
    {
        uint32_t s = 0;
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            s += __vmaxu2(alo[r], blo[r]) - __vminu2(alo[r], blo[r]);
            s += __vmaxu2(ahi[r], bhi[r]) - __vminu2(ahi[r], bhi[r]);
        }
        acc += (ACC)((s & 0xffffu) + (s >> 16));
    }
    else if constexpr (OP == OP_SATD)
    {
        uint32_t dl[4], dh[4];
#pragma unroll
        for (int r = 0; r < 4; r++) { dl[r] = alo[r] - blo[r]; dh[r] = ahi[r] - bhi[r]; }
        // vertical Hadamard on packed columns (0,1) and (2,3)
        uint32_t s0 = dl[0] + dl[1], s1 = dl[0] - dl[1], s2 = dl[2] + dl[3], s3 = dl[2] - dl[3];
        dl[0] = s0 + s2; dl[1] = s1 + s3; dl[2] = s0 - s2; dl[3] = s1 - s3;
        s0 = dh[0] + dh[1]; s1 = dh[0] - dh[1]; s2 = dh[2] + dh[3]; s3 = dh[2] - dh[3];
        dh[0] = s0 + s2; dh[1] = s1 + s3; dh[2] = s0 - s2; dh[3] = s1 - s3;
        int sum = 0;
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            int p0, q0, p1, q1;
            unpack_s16x2(dl[r] + dh[r], p0, q0);          // columns (0 + 2, 1 + 3)
            unpack_s16x2(dl[r] - dh[r], p1, q1);          // columns (0 - 2, 1 - 3)
            sum += max(abs(p0), abs(q0)) + max(abs(p1), abs(q1));
        }
        acc += sum;
    }
    else
    {
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            int x0, x1, x2, x3;
            unpack_s16x2(alo[r] - blo[r], x0, x1);
            unpack_s16x2(ahi[r] - bhi[r], x2, x3);
            acc += (ACC)(x0 * x0 + x1 * x1) + (ACC)(x2 * x2 + x3 * x3);   // |x| <= 4095: no int overflow
        }
    }
}

// Differences are formed on packed 16-bit pairs (one IADD per two samples, no unpacking of the
// inputs); for SATD the vertical 4-point Hadamard also runs packed (|value| <= 4 * 4095 fits a
// 16-bit lane at every depth), then the eight words are unpacked once for the horizontal pass.
// Plane strides must be multiples of 4 samples.  UNROLL tiles are loaded before any is consumed.
template<typename T, int OP, typename ACC, typename OUT, int UNROLL, int MINBLK>
__global__ void __launch_bounds__(256, MINBLK)
tile4_fast_kernel(const T* __restrict__ A, intptr_t sa, const T* __restrict__ B, intptr_t sb,
                  const int32_t* __restrict__ offA, const int32_t* __restrict__ offB, int kdiv,
                  int n, int w, int h, int G, OUT* __restrict__ out)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int lg = __ffs(G) - 1;                                  // G is a power of two
    int blk = (int)(gid >> lg);
    int l = (int)gid & (G - 1);
    bool live = blk < n;
    int tw = w >> 2;
    int T4 = tw * (h >> 2);
    const T* a = A;
    const T* b = B;
    if (live)
    {
        a += offA[kdiv > 1 ? blk / kdiv : blk];
        b += offB[blk];
    }
    // tile coordinates advance incrementally (no division in the loop)
    int tx = l % tw, ty = l / tw;
    int dx = G % tw, dy = G / tw;
    ACC acc = 0;
    if (live)
    {
        int t = l;
        if (UNROLL == 2)
        {
            for (; t + G < T4; t += 2 * G)
            {
                uint32_t alo[2][4], ahi[2][4], blo[2][4], bhi[2][4];
#pragma unroll
                for (int u = 0; u < 2; u++)
                {
                    load_tile4x4(a + (intptr_t)(ty << 2) * sa + (tx << 2), sa, alo[u], ahi[u]);
                    load_tile4x4(b + (intptr_t)(ty << 2) * sb + (tx << 2), sb, blo[u], bhi[u]);
                    tx += dx; ty += dy;
                    if (tx >= tw) { tx -= tw; ty++; }
                }
#pragma unroll
                for (int u = 0; u < 2; u++) tile4_accumulate<OP, ACC>(alo[u], ahi[u], blo[u], bhi[u], acc);
            }
        }
        for (; t < T4; t += G)
        {
            uint32_t alo[4], ahi[4], blo[4], bhi[4];
            load_tile4x4(a + (intptr_t)(ty << 2) * sa + (tx << 2), sa, alo, ahi);
            load_tile4x4(b + (intptr_t)(ty << 2) * sb + (tx << 2), sb, blo, bhi);
            tx += dx; ty += dy;
            if (tx >= tw) { tx -= tw; ty++; }
            tile4_accumulate<OP, ACC>(alo, ahi, blo, bhi, acc);
        }
    }
    acc = group_sum(acc, G);
    if (live && l == 0) out[blk] = (OUT)acc;
}

// Narrow blocks (w = 8 or 16, 16-bit samples, strides % 8 == 0): thread = 8x4 strip loaded with 16-byte chunk loads
// (device_util.cuh load_rows8_v16), two 4x4 tiles per strip; G lanes (power of two, the strips of one block) reduce together.
template<int OP, typename ACC, typename OUT>
__global__ void __launch_bounds__(128)
strip8_fast_kernel(const uint16_t* __restrict__ A, intptr_t sa, const uint16_t* __restrict__ B, intptr_t sb,
                   const int32_t* __restrict__ offA, const int32_t* __restrict__ offB, int kdiv,
                   int n, int w, int h, int G, OUT* __restrict__ out)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int lg = __ffs(G) - 1;
    int blk = (int)(gid >> lg);
    int l = (int)gid & (G - 1);
    bool live = blk < n;
    int sw = w >> 3;
    int S = sw * (h >> 2);
    ACC acc = 0;
    if (live)
    {
        const uint16_t* a = A + offA[kdiv > 1 ? blk / kdiv : blk];
        const uint16_t* b = B + offB[blk];
        for (int t = l; t < S; t += G)
        {
            int sx = (t % sw) << 3, sy = (t / sw) << 2;
            uint32_t wa[4][4], wb[4][4];
            load_rows8_v16(a + (intptr_t)sy * sa + sx, sa, wa);
            load_rows8_v16(b + (intptr_t)sy * sb + sx, sb, wb);
            uint32_t alo[4], ahi[4], blo[4], bhi[4];
#pragma unroll
            for (int half = 0; half < 2; half++)
            {
#pragma unroll
                for (int r = 0; r < 4; r++) { alo[r] = wa[r][2 * half]; ahi[r] = wa[r][2 * half + 1]; blo[r] = wb[r][2 * half]; bhi[r] = wb[r][2 * half + 1]; }
                tile4_accumulate<OP, ACC>(alo, ahi, blo, bhi, acc);
            }
        }
    }
    acc = group_sum(acc, G);
    if (live && l == 0) out[blk] = (OUT)acc;
}

// All rectangular partitions of a CU in one pass (SATD): the 2Nx2N PU, the two 2NxN PUs and the two Nx2N PUs of an S x S
// CU, each against its OWN reference block (five motion vectors per CU), as the inter analysis costs them
// (reference encoder/search.cpp predInterSearch -> one satd per PU; encoder/analysis.cpp checkInter_rd0_4 for SIZE_2Nx2N /
// SIZE_2NxN / SIZE_Nx2N).  Every PU's SATD is the sum of its 4x4-tile SATDs (tile4_accumulate), so a lane that owns a fenc
// tile loads it ONCE and meets it with the matching tile of three reference blocks; the three per-shape launches read the
// fenc plane three times.  offR[5 * cu + k], out[5 * cu + k]: k = 0 2Nx2N, 1 / 2 upper / lower 2NxN, 3 / 4 left / right Nx2N,
// each offset addressing the top-left sample of that PU's reference block.  G lanes (power of two) share a CU.
template<typename T, int S, int MINB = 0>
__global__ void __launch_bounds__(128, MINB)
cu_satd_kernel(const T* __restrict__ A, intptr_t sa, const T* __restrict__ B, intptr_t sb,
               const int32_t* __restrict__ offF, const int32_t* __restrict__ offR, int n, int G, int32_t* __restrict__ out)
{
    constexpr int TW = S >> 2, T4 = TW * TW, HALF = S >> 1;
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int lg = __ffs(G) - 1;
    int cu = (int)(gid >> lg);
    int l = (int)gid & (G - 1);
    bool live = cu < n;
    int acc0 = 0, accH0 = 0, accH1 = 0, accV0 = 0, accV1 = 0;
    if (live)
    {
        const T* a = A + offF[cu];
        const int32_t* r = offR + 5 * (size_t)cu;
        const T* b0 = B + r[0];
        const T* bH[2] = { B + r[1], B + r[2] };
        const T* bV[2] = { B + r[3], B + r[4] };
        for (int t = l; t < T4; t += G)
        {
            const int x = (t % TW) << 2, y = (t / TW) << 2;
            const int kH = y >= HALF, kV = x >= HALF;
            uint32_t alo[4], ahi[4], b0lo[4], b0hi[4], b1lo[4], b1hi[4], b2lo[4], b2hi[4];
            load_tile4x4(a + (intptr_t)y * sa + x, sa, alo, ahi);
            load_tile4x4(b0 + (intptr_t)y * sb + x, sb, b0lo, b0hi);
            load_tile4x4((kH ? bH[1] : bH[0]) + (intptr_t)(y - kH * HALF) * sb + x, sb, b1lo, b1hi);
            load_tile4x4((kV ? bV[1] : bV[0]) + (intptr_t)y * sb + (x - kV * HALF), sb, b2lo, b2hi);
            int c0 = 0, c1 = 0, c2 = 0;
            tile4_accumulate<OP_SATD, int>(alo, ahi, b0lo, b0hi, c0);
            tile4_accumulate<OP_SATD, int>(alo, ahi, b1lo, b1hi, c1);
            tile4_accumulate<OP_SATD, int>(alo, ahi, b2lo, b2hi, c2);
            acc0 += c0;
            accH0 += kH ? 0 : c1; accH1 += kH ? c1 : 0;
            accV0 += kV ? 0 : c2; accV1 += kV ? c2 : 0;
        }
    }
    acc0 = group_sum(acc0, G);
    accH0 = group_sum(accH0, G); accH1 = group_sum(accH1, G);
    accV0 = group_sum(accV0, G); accV1 = group_sum(accV1, G);
    if (live && l < 5)
    {
        int v = l == 0 ? acc0 : l == 1 ? accH0 : l == 2 ? accH1 : l == 3 ? accV0 : accV1;
        if (G >= 8 || l == 0)
        {
            if (G >= 8) out[5 * (size_t)cu + l] = v;
            else
            {
                int32_t* o = out + 5 * (size_t)cu;
                o[0] = acc0; o[1] = accH0; o[2] = accH1; o[3] = accV0; o[4] = accV1;
            }
        }
    }
}

// ---- the same entry with the 4-point Hadamard on the tensor cores (16-bit planes, depth <= 10) -----------------------------------------
// ncu on cu_satd_kernel (profiles/r3_cu_satd_ncu_summary.txt): the ALU pipe is 68-74 % busy for 16 ... 64 wide CUs, DRAM 35 %: the packed-integer
// Hadamard, not memory, sets the time (B200 issues 16 ALU-pipe lanes per clock and scheduler: at the HBM roofline a 4-byte sample may cost about
// 11 ALU-pipe operations, this kernel spends 25 per CU sample).  Here the horizontal transform of every tile is one mma.sync.m16n8k16 (f16
// operands, f32 accumulators):
//   * a pixel p <= 1023 becomes the half-precision number 1024 + p by adding 0x6400 to its 16 bits (exponent 2^10, unit in the last place 1), two
//     samples per instruction, no conversion; every product with +-1 and every partial sum (< 2^17) is exact in f32;
//   * A row g of the MMA is the concatenation of one tile row of the four lanes 4g .. 4g + 3 (lane t supplies k = 2t, 2t + 1, 2t + 8, 2t + 9: its
//     packed words lo / hi as loaded), B is block-diagonal with the Hadamard matrix on the diagonal, so the four lanes' tiles stay apart: a lane
//     may hold ANY tile and the lane <-> tile assignment, the loads and the per-PU bookkeeping of cu_satd_kernel carry over unchanged;
//   * fenc goes through +B once per tile, each reference block through -B accumulating on fenc's result: the difference is never formed on the
//     ALU, and the 1024 offsets cancel exactly (they only reach the u = 0 outputs: +4096 - 4096);
//   * MMA rows g and g + 8 carry tile rows (0, 1) in one instruction and (2, 3) in the next, so a lane receives all four rows of two output
//     columns and finishes the vertical 4-point transform in registers: four FADD and, with the last stage folded as in tile4_accumulate
//     (|a + b| + |a - b| = 2 max(|a|, |b|), which is also the reference's per-tile >> 1), two FMNMX with free |.| operand modifiers per column.
// The two column pairs a lane receives belong to the tiles of lanes (lane & ~3) | (t >> 1) and ... | 2 | (t >> 1): their PU flags come by shuffle.
__device__ __forceinline__ void hmma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, const float (&c)[4])
{
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %11, %12, %13};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}
// sum over the two columns this lane received (c = 0, 1) of max(|s0|, |s2|) + max(|s1|, |s3|): x[j] = rows (2j, 2j + 1) x columns (c0, c1)
__device__ __forceinline__ float hadamard_cols_fold(const float (&r01)[4], const float (&r23)[4])
{
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 2; c++)
    {
        const float x0 = r01[c], x1 = r01[2 + c], x2 = r23[c], x3 = r23[2 + c];
        const float s0 = x0 + x1, s1 = x0 - x1, s2 = x2 + x3, s3 = x2 - x3;
        sum += fmaxf(fabsf(s0), fabsf(s2)) + fmaxf(fabsf(s1), fabsf(s3));
    }
    return sum;
}

template<typename T, int S>
__global__ void __launch_bounds__(128, 1)                   // (128, 1): 118 registers and all tile loads hoisted; without the 1 ptxas picks 110 and the kernel is 20 % slower
cu_satd_mma_kernel(const T* __restrict__ A, intptr_t sa, const T* __restrict__ B, intptr_t sb,
                   const int32_t* __restrict__ offF, const int32_t* __restrict__ offR, int n, int G, int32_t* __restrict__ out)
{
    constexpr int TW = S >> 2, T4 = TW * TW, HALF = S >> 1;
    constexpr uint32_t F16_1024 = 0x64006400u;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lg = __ffs(G) - 1;                            // G: power of two, 4 <= G <= 32 (a quad of lanes never spans two CUs)
    const int cu = (int)(gid >> lg);
    const int l = (int)gid & (G - 1);
    const bool live = cu < n;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    // B fragments of the two column halves (h = 0: the tiles of lanes t = 0, 1 of every quad; h = 1: t = 2, 3): column n = g of the half is
    // frequency u = g & 3 of tile tau = 2 h + (g >> 2); this lane supplies k = its own four samples, i.e. non-zero only when t == tau
    uint32_t bp[2][2], bn[2][2];
#pragma unroll
    for (int h = 0; h < 2; h++)
    {
        const int tau = 2 * h + (g >> 2), u = g & 3;
        const uint32_t one = 0x3C00u, neg = 0xBC00u;
        // H[u][c] = (-1)^popc(u & c)
        const uint32_t h0 = one, h1 = (u & 1) ? neg : one, h2 = (u & 2) ? neg : one, h3 = (__popc(u & 3) & 1) ? neg : one;
        bp[h][0] = t == tau ? (h0 | (h1 << 16)) : 0u;
        bp[h][1] = t == tau ? (h2 | (h3 << 16)) : 0u;
        bn[h][0] = bp[h][0] ^ 0x80008000u;
        bn[h][1] = bp[h][1] ^ 0x80008000u;
    }
    const T* a = A;                                         // 8-bit pictures: load_tile4x4 widens the bytes to the same packed 16-bit pairs
    const T* b0 = B;
    const T* bH[2] = { B, B };
    const T* bV[2] = { B, B };
    if (live)
    {
        const int32_t* r = offR + 5 * (size_t)cu;
        a = A + offF[cu];
        b0 = B + r[0];
        bH[0] = B + r[1]; bH[1] = B + r[2];
        bV[0] = B + r[3]; bV[1] = B + r[4];
    }
    const int src0 = (lane & ~3) | (t >> 1), src1 = src0 | 2;   // whose tiles this lane's two column pairs belong to
    float acc0 = 0.f, accH0 = 0.f, accH1 = 0.f, accV0 = 0.f, accV1 = 0.f;
    const float zero[4] = { 0.f, 0.f, 0.f, 0.f };
    for (int tt = l; tt < T4; tt += G)                      // the trip count is the same for every lane of the warp (T4 % G == 0)
    {
        const int x = (tt % TW) << 2, y = (tt / TW) << 2;
        const int kH = y >= HALF, kV = x >= HALF;
        uint32_t alo[4], ahi[4], rlo[3][4], rhi[3][4];
        load_tile4x4(a + (intptr_t)y * sa + x, sa, alo, ahi);
        load_tile4x4(b0 + (intptr_t)y * sb + x, sb, rlo[0], rhi[0]);
        load_tile4x4((kH ? bH[1] : bH[0]) + (intptr_t)(y - kH * HALF) * sb + x, sb, rlo[1], rhi[1]);
        load_tile4x4((kV ? bV[1] : bV[0]) + (intptr_t)y * sb + (x - kV * HALF), sb, rlo[2], rhi[2]);
        const int flags = kH | (kV << 1);
        const int f0 = __shfl_sync(0xffffffffu, flags, src0), f1 = __shfl_sync(0xffffffffu, flags, src1);
        float cf[2][2][4];
#pragma unroll
        for (int j = 0; j < 2; j++)
        {
            const uint32_t fa[4] = { alo[2 * j] + F16_1024, alo[2 * j + 1] + F16_1024, ahi[2 * j] + F16_1024, ahi[2 * j + 1] + F16_1024 };
            hmma_16816(cf[0][j], fa, bp[0][0], bp[0][1], zero);
            hmma_16816(cf[1][j], fa, bp[1][0], bp[1][1], zero);
        }
#pragma unroll
        for (int p = 0; p < 3; p++)
        {
            float d[2][2][4];
#pragma unroll
            for (int j = 0; j < 2; j++)
            {
                const uint32_t ra[4] = { rlo[p][2 * j] + F16_1024, rlo[p][2 * j + 1] + F16_1024, rhi[p][2 * j] + F16_1024, rhi[p][2 * j + 1] + F16_1024 };
                hmma_16816(d[0][j], ra, bn[0][0], bn[0][1], cf[0][j]);
                hmma_16816(d[1][j], ra, bn[1][0], bn[1][1], cf[1][j]);
            }
            const float s0 = hadamard_cols_fold(d[0][0], d[0][1]), s1 = hadamard_cols_fold(d[1][0], d[1][1]);
            if (p == 0) acc0 += s0 + s1;
            else if (p == 1)
            {
                accH0 += ((f0 & 1) ? 0.f : s0) + ((f1 & 1) ? 0.f : s1);
                accH1 += ((f0 & 1) ? s0 : 0.f) + ((f1 & 1) ? s1 : 0.f);
            }
            else
            {
                accV0 += ((f0 & 2) ? 0.f : s0) + ((f1 & 2) ? 0.f : s1);
                accV1 += ((f0 & 2) ? s0 : 0.f) + ((f1 & 2) ? s1 : 0.f);
            }
        }
    }
    // per-lane sums are integers below 2^24 (a tile's SATD is at most 16 * 16 * 1023 / 2): exact in f32
    int i0 = group_sum(__float2int_rn(acc0), G);
    int iH0 = group_sum(__float2int_rn(accH0), G), iH1 = group_sum(__float2int_rn(accH1), G);
    int iV0 = group_sum(__float2int_rn(accV0), G), iV1 = group_sum(__float2int_rn(accV1), G);
    if (live && l < 5)
    {
        const int v = l == 0 ? i0 : l == 1 ? iH0 : l == 2 ? iH1 : l == 3 ? iV0 : iV1;
        if (G >= 8) out[5 * (size_t)cu + l] = v;
        else if (l == 0)
        {
            int32_t* o = out + 5 * (size_t)cu;
            o[0] = i0; o[1] = iH0; o[2] = iH1; o[3] = iV0; o[4] = iV1;
        }
    }
}

} // namespace b200
