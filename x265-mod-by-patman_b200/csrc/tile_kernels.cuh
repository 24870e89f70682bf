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

} // namespace b200
