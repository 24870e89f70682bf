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
template<typename T, int S>
__global__ void __launch_bounds__(128)
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

} // namespace b200
