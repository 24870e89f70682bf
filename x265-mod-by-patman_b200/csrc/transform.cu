// transform.cu -- forward/inverse integer DCT 4/8/16/32, 4x4 DST, lowpass DCT, quant / dequant.
//
// Bit-exact CUDA restatement of the reference's dct.cpp:43-715 and lowpassdct.cpp:34-116.
//
// Transform kernels: one thread owns one row (forward) / column (inverse) of a TU in registers and
// runs the even/odd partial-butterfly factorisation as a compile-time recursion; the coefficient
// matrix lives in constant memory, and after full unrolling every coefficient is a constant-bank
// operand of an IMAD.  The N threads of a TU sit in one warp, so the inter-stage transpose goes
// through a padded shared-memory tile with only __syncwarp().  Global traffic is coalesced 32-bit
// rows in and out.
#include "internal.h"
#include "device_util.cuh"

namespace b200 {

// T32[k][n]; TN[k][n] = T32[k * 32 / N][n]  (HEVC core transform, == g_t4..g_t32 constants.cpp:270-344)
__constant__ int c_T32[32][32];
// DST-VII 4x4 (== fastForwardDst / inversedst dct.cpp:43-81)
__constant__ int c_DST4[4][4];

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

int upload_transform_tables(x265b200_ctx* ctx)
{
    int t[32][32];
    for (int k = 0; k < 32; k++)
        for (int n = 0; n < 32; n++) t[k][n] = h_cos128(k * (2 * n + 1));
    static const int dst4[4][4] = { { 29, 55, 74, 84 }, { 74, 74, 0, -74 }, { 84, -29, -74, 55 }, { 55, -84, 74, -29 } };
    B200_CUDA(ctx, cudaMemcpyToSymbol(c_T32, t, sizeof(t)));
    B200_CUDA(ctx, cudaMemcpyToSymbol(c_DST4, dst4, sizeof(dst4)));
    return X265B200_OK;
}

// y[k] = sum_i TN[k][i] * x[i]
template<int N> struct Fwd
{
    __device__ __forceinline__ static void run(const int (&x)[N], int (&y)[N])
    {
        int E[N / 2], O[N / 2], ye[N / 2];
#pragma unroll
        for (int i = 0; i < N / 2; i++) { E[i] = x[i] + x[N - 1 - i]; O[i] = x[i] - x[N - 1 - i]; }
        Fwd<N / 2>::run(E, ye);
#pragma unroll
        for (int k = 0; k < N / 2; k++)
        {
            int acc = 0;
#pragma unroll
            for (int i = 0; i < N / 2; i++) acc += c_T32[(2 * k + 1) * (32 / N)][i] * O[i];
            y[2 * k] = ye[k];
            y[2 * k + 1] = acc;
        }
    }
};
template<> struct Fwd<2>
{
    __device__ __forceinline__ static void run(const int (&x)[2], int (&y)[2])
    {
        y[0] = 64 * (x[0] + x[1]);
        y[1] = 64 * (x[0] - x[1]);
    }
};

// x[i] = sum_k TN[k][i] * c[k]
template<int N> struct Inv
{
    __device__ __forceinline__ static void run(const int (&c)[N], int (&x)[N])
    {
        int ce[N / 2], E[N / 2];
#pragma unroll
        for (int k = 0; k < N / 2; k++) ce[k] = c[2 * k];
        Inv<N / 2>::run(ce, E);
#pragma unroll
        for (int i = 0; i < N / 2; i++)
        {
            int o = 0;
#pragma unroll
            for (int k = 0; k < N / 2; k++) o += c_T32[(2 * k + 1) * (32 / N)][i] * c[2 * k + 1];
            x[i] = E[i] + o;
            x[N - 1 - i] = E[i] - o;
        }
    }
};
template<> struct Inv<2>
{
    __device__ __forceinline__ static void run(const int (&c)[2], int (&x)[2])
    {
        x[0] = 64 * (c[0] + c[1]);
        x[1] = 64 * (c[0] - c[1]);
    }
};

__device__ __forceinline__ void dst4_fwd(const int (&x)[4], int (&y)[4])
{
#pragma unroll
    for (int k = 0; k < 4; k++)
        y[k] = c_DST4[k][0] * x[0] + c_DST4[k][1] * x[1] + c_DST4[k][2] * x[2] + c_DST4[k][3] * x[3];
}
__device__ __forceinline__ void dst4_inv(const int (&c)[4], int (&x)[4])
{
#pragma unroll
    for (int i = 0; i < 4; i++)
        x[i] = c_DST4[0][i] * c[0] + c_DST4[1][i] * c[1] + c_DST4[2][i] * c[2] + c_DST4[3][i] * c[3];
}

__device__ __forceinline__ int clip16(int v) { return min(32767, max(-32768, v)); }

enum { MODE_DCT = 0, MODE_DST = 1, MODE_LOWPASS = 2 };
constexpr int TR_THREADS = 128;

template<int N> struct Tile
{
    static constexpr int LD = N + 2;                  // int16 row pitch; (N+2)/2 words is odd -> conflict-free
    static constexpr int SIZE = N * LD;
    static constexpr int PER_CTA = TR_THREADS / N;
};

// copy a strided N x N int16 block (global) <-> padded smem tile with the TU's N threads
template<int N>
__device__ __forceinline__ void tile_load(int16_t* tile, const int16_t* g, intptr_t gs, int j)
{
    constexpr int LD = Tile<N>::LD;
    if ((((uintptr_t)g | (uintptr_t)(gs * 2)) & 3) == 0)
    {
#pragma unroll
        for (int idx = j; idx < N * N / 2; idx += N)
        {
            int r = idx / (N / 2), cw = idx % (N / 2);
            *(uint32_t*)(tile + r * LD + 2 * cw) = __ldg((const uint32_t*)(g + r * gs) + cw);
        }
    }
    else
    {
#pragma unroll
        for (int idx = j; idx < N * N; idx += N)
        {
            int r = idx / N, c = idx % N;
            tile[r * LD + c] = __ldg(g + r * gs + c);
        }
    }
}
template<int N>
__device__ __forceinline__ void tile_store(const int16_t* tile, int16_t* g, intptr_t gs, int j)
{
    constexpr int LD = Tile<N>::LD;
    if ((((uintptr_t)g | (uintptr_t)(gs * 2)) & 3) == 0)
    {
#pragma unroll
        for (int idx = j; idx < N * N / 2; idx += N)
        {
            int r = idx / (N / 2), cw = idx % (N / 2);
            *((uint32_t*)(g + r * gs) + cw) = *(const uint32_t*)(tile + r * LD + 2 * cw);
        }
    }
    else
    {
#pragma unroll
        for (int idx = j; idx < N * N; idx += N)
        {
            int r = idx / N, c = idx % N;
            g[r * gs + c] = tile[r * LD + c];
        }
    }
}

// Forward transform.  MODE_LOWPASS: N is the size of the inner DCT, the TU is 2N x 2N.
template<int N, int MODE>
__global__ void __launch_bounds__(TR_THREADS)
fwd_kernel(const int16_t* __restrict__ src, intptr_t srcStride, const int32_t* __restrict__ off, int n,
           int16_t* __restrict__ dst, int shift1, int shift2, int depth)
{
    constexpr int LD = Tile<N>::LD;
    __shared__ __align__(16) int16_t s_a[Tile<N>::PER_CTA * Tile<N>::SIZE];
    __shared__ __align__(16) int16_t s_b[Tile<N>::PER_CTA * Tile<N>::SIZE];
    int g = threadIdx.x / N, j = threadIdx.x % N;
    int tu = blockIdx.x * Tile<N>::PER_CTA + g;
    bool live = tu < n;
    int16_t* ta = s_a + g * Tile<N>::SIZE;
    int16_t* tb = s_b + g * Tile<N>::SIZE;
    int total = 0;

    if (live)
    {
        const int16_t* p = src + (off ? (size_t)off[tu] : (size_t)tu * (MODE == MODE_LOWPASS ? 4 * N * N : N * N));
        if (MODE == MODE_LOWPASS)
        {
            // lowpassdct.cpp:40-49: 2x2 sums truncated to int16, average = sum >> 2
            const int16_t* r0 = p + (intptr_t)(2 * j) * srcStride;
            const int16_t* r1 = r0 + srcStride;
#pragma unroll
            for (int c = 0; c < N; c++)
            {
                int16_t s = (int16_t)(r0[2 * c] + r0[2 * c + 1] + r1[2 * c] + r1[2 * c + 1]);
                ta[j * LD + c] = (int16_t)(s >> 2);
                total += s;
            }
        }
        else
            tile_load<N>(ta, p, srcStride, j);
    }
    __syncwarp();

    int x[N], y[N];
    if (live)
    {
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = ta[j * LD + i];
        if (MODE == MODE_DST) dst4_fwd((const int(&)[4])x, (int(&)[4])y); else Fwd<N>::run(x, y);
        int add = 1 << (shift1 - 1);
#pragma unroll
        for (int k = 0; k < N; k++) tb[k * LD + j] = (int16_t)((y[k] + add) >> shift1);   // truncation, dct.cpp:113
    }
    __syncwarp();
    if (live)
    {
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = tb[j * LD + i];
        if (MODE == MODE_DST) dst4_fwd((const int(&)[4])x, (int(&)[4])y); else Fwd<N>::run(x, y);
        int add = 1 << (shift2 - 1);
#pragma unroll
        for (int k = 0; k < N; k++) ta[k * LD + j] = (int16_t)((y[k] + add) >> shift2);
    }
    __syncwarp();

    if (MODE == MODE_LOWPASS)
    {
        // block sum over the TU's N threads (N <= 16 lanes of one warp)
        for (int m = N >> 1; m > 0; m >>= 1) total += __shfl_xor_sync(0xffffffffu, total, m);
        if (live)
        {
            int16_t* o = dst + (size_t)tu * (4 * N * N);
            // rows j and j + N of the 2N x 2N output: top-left quadrant = coefficients, rest zero
#pragma unroll
            for (int c = 0; c < 2 * N; c++)
            {
                o[j * 2 * N + c] = c < N ? ta[j * LD + c] : (int16_t)0;
                o[(j + N) * 2 * N + c] = 0;
            }
            if (j == 0)     // same thread wrote row 0 above, so program order suffices
            {
                int16_t dc;
                if (N == 4)         // lowPassDct8_c: int16 running total, <<1 at 8 bit, >>(depth-9) otherwise
                    dc = depth == 8 ? (int16_t)((int)(int16_t)total << 1) : (int16_t)((int16_t)total >> (depth - 9));
                else if (N == 8)
                    dc = (int16_t)(total >> (1 + (depth - 8)));
                else
                    dc = (int16_t)(total >> (3 + (depth - 8)));
                o[0] = dc;
            }
        }
    }
    else if (live)
        tile_store<N>(ta, dst + (size_t)tu * (N * N), N, j);
}

template<int N, int MODE>
__global__ void __launch_bounds__(TR_THREADS)
inv_kernel(const int16_t* __restrict__ src, int n, int16_t* __restrict__ dst, intptr_t dstStride,
           const int32_t* __restrict__ off, int shift1, int shift2)
{
    constexpr int LD = Tile<N>::LD;
    __shared__ __align__(16) int16_t s_a[Tile<N>::PER_CTA * Tile<N>::SIZE];
    __shared__ __align__(16) int16_t s_b[Tile<N>::PER_CTA * Tile<N>::SIZE];
    int g = threadIdx.x / N, j = threadIdx.x % N;
    int tu = blockIdx.x * Tile<N>::PER_CTA + g;
    bool live = tu < n;
    int16_t* ta = s_a + g * Tile<N>::SIZE;
    int16_t* tb = s_b + g * Tile<N>::SIZE;

    if (live) tile_load<N>(ta, src + (size_t)tu * (N * N), N, j);
    __syncwarp();
    int c[N], x[N];
    if (live)
    {
#pragma unroll
        for (int k = 0; k < N; k++) c[k] = ta[k * LD + j];
        if (MODE == MODE_DST) dst4_inv((const int(&)[4])c, (int(&)[4])x); else Inv<N>::run(c, x);
        int add = 1 << (shift1 - 1);
#pragma unroll
        for (int i = 0; i < N; i++) tb[j * LD + i] = (int16_t)clip16((x[i] + add) >> shift1);   // saturation, dct.cpp:257
    }
    __syncwarp();
    if (live)
    {
#pragma unroll
        for (int k = 0; k < N; k++) c[k] = tb[k * LD + j];
        if (MODE == MODE_DST) dst4_inv((const int(&)[4])c, (int(&)[4])x); else Inv<N>::run(c, x);
        int add = 1 << (shift2 - 1);
#pragma unroll
        for (int i = 0; i < N; i++) ta[j * LD + i] = (int16_t)clip16((x[i] + add) >> shift2);
    }
    __syncwarp();
    if (live) tile_store<N>(ta, dst + (off ? (size_t)off[tu] : (size_t)tu * (N * N)), dstStride, j);
}

// ---------------------------------------------------------------- quant family
// One thread per 8 coefficients.  numSig is accumulated with a warp-segmented sum and one atomic
// per (warp, block) into a zeroed counter.
// Throughput version for numCoeff = 16 / 64 / 256 / 1024: one thread owns the same eight coefficient positions of NB
// consecutive blocks, so the quantCoeff entries are fetched once per NB blocks and NB 16-byte coefficient loads are in
// flight per thread.  Blocks of <= 256 coefficients are reduced inside one lane group and numSig is stored directly
// (no zeroing pass, no atomics); 32x32 blocks span four warps and use one atomic per warp.
template<bool NQUANT, bool STORE_DU, int NB>
__global__ void __launch_bounds__(256)
quant_multi_kernel(const int16_t* __restrict__ coef, const int32_t* __restrict__ quantCoeff, int32_t* __restrict__ deltaU,
                   int16_t* __restrict__ qCoef, int qBits, int add, int numCoeff, int n, uint32_t* __restrict__ numSig)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int per = numCoeff >> 3;                       // power of two
    int lgper = __ffs(per) - 1;
    int grp = (int)(gid >> lgper);
    int pos = ((int)gid & (per - 1)) << 3;
    int blk0 = grp * NB;
    bool live = blk0 < n;
    int sig[NB];
#pragma unroll
    for (int j = 0; j < NB; j++) sig[j] = 0;
    if (live)
    {
        int4 cv[NB];
#pragma unroll
        for (int j = 0; j < NB; j++)
            cv[j] = blk0 + j < n ? __ldg((const int4*)(coef + (size_t)(blk0 + j) * numCoeff + pos)) : make_int4(0, 0, 0, 0);
        int4 q0 = __ldg((const int4*)(quantCoeff + pos));
        int4 q1 = __ldg((const int4*)(quantCoeff + pos + 4));
        int q[8] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w };
        int qBits8 = qBits - 8;
#pragma unroll
        for (int j = 0; j < NB; j++)
        {
            if (blk0 + j >= n) break;
            int c[8] = { (int16_t)(cv[j].x & 0xffff), cv[j].x >> 16, (int16_t)(cv[j].y & 0xffff), cv[j].y >> 16,
                         (int16_t)(cv[j].z & 0xffff), cv[j].z >> 16, (int16_t)(cv[j].w & 0xffff), cv[j].w >> 16 };
            if constexpr (NQUANT)
            {   // nquant is ALU-pipe bound (about 11 ALU operations per coefficient at 4 bytes per coefficient): the tail runs on packed pairs --
                // clip3(-32768, 32767, .) of two levels is one cvt.pack.sat, (int16)abs(.) is max(w, -w) per 16-bit lane (-32768 stays 0x8000 as
                // the cast leaves it), and the non-zero count is taken from the packed result (a level is zero exactly when its output is)
                uint32_t o[4], nz = 0;
#pragma unroll
                for (int i = 0; i < 8; i += 2)
                {
                    int x[2];
#pragma unroll
                    for (int e = 0; e < 2; e++)
                    {
                        const int m = c[i + e] >> 31;
                        const int l = (int)((unsigned)abs(c[i + e]) * (unsigned)q[i + e] + (unsigned)add) >> qBits;   // int32 wrap, dct.cpp:702-703
                        x[e] = (l ^ m) - m;                                                                        // level * sign
                    }
                    uint32_t w;
                    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(w) : "r"(x[1]), "r"(x[0]));
                    w = __viaddmax_s16x2(~w, 0x00010001u, w);
                    o[i >> 1] = w;
                    nz += __vminu2(w, 0x00010001u);
                }
                sig[j] += (int)((nz & 0xffffu) + (nz >> 16));
                *(int4*)(qCoef + (size_t)(blk0 + j) * numCoeff + pos) = make_int4((int)o[0], (int)o[1], (int)o[2], (int)o[3]);
            }
            else
            {
            int lv[8], du[8];
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                int level = c[i];
                int sign = level < 0 ? -1 : 1;
                int tmplevel = (int)((unsigned)abs(level) * (unsigned)q[i]);          // int32 wrap, dct.cpp:678
                level = (int)((unsigned)tmplevel + (unsigned)add) >> qBits;
                du[i] = (int)((unsigned)tmplevel - ((unsigned)level << qBits)) >> qBits8;
                sig[j] += level != 0;
                level = (int)((unsigned)level * (unsigned)sign);
                level = clip16(level);
                lv[i] = level;
            }
            size_t base = (size_t)(blk0 + j) * numCoeff + pos;
            int4 o;
            o.x = (lv[0] & 0xffff) | (lv[1] << 16); o.y = (lv[2] & 0xffff) | (lv[3] << 16);
            o.z = (lv[4] & 0xffff) | (lv[5] << 16); o.w = (lv[6] & 0xffff) | (lv[7] << 16);
            *(int4*)(qCoef + base) = o;
            if (STORE_DU)
            {
                *(int4*)(deltaU + base) = make_int4(du[0], du[1], du[2], du[3]);
                *(int4*)(deltaU + base + 4) = make_int4(du[4], du[5], du[6], du[7]);
            }
            }
        }
    }
    int lane = threadIdx.x & 31;
    int G = per < 32 ? per : 32;
#pragma unroll
    for (int j = 0; j < NB; j++)
    {
        int sum = group_sum(sig[j], G);
        if (live && blk0 + j < n && (lane & (G - 1)) == 0)
        {
            if (per <= 32) numSig[blk0 + j] = (uint32_t)sum;
            else if (sum) atomicAdd(numSig + blk0 + j, (uint32_t)sum);
        }
    }
}

template<bool NQUANT, bool STORE_DU>
__global__ void __launch_bounds__(256)
quant_kernel(const int16_t* __restrict__ coef, const int32_t* __restrict__ quantCoeff, int32_t* __restrict__ deltaU,
             int16_t* __restrict__ qCoef, int qBits, int add, int numCoeff, int n, uint32_t* __restrict__ numSig)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int per = numCoeff >> 3;                       // threads per block of coefficients
    int blk = (int)(gid / per);
    bool live = blk < n;
    int sig = 0;
    if (live)
    {
        int pos = (int)(gid % per) << 3;
        size_t base = (size_t)blk * numCoeff + pos;
        int4 cv = __ldg((const int4*)(coef + base));
        int4 q0 = __ldg((const int4*)(quantCoeff + pos));
        int4 q1 = __ldg((const int4*)(quantCoeff + pos + 4));
        int c[8] = { (int16_t)(cv.x & 0xffff), cv.x >> 16, (int16_t)(cv.y & 0xffff), cv.y >> 16,
                     (int16_t)(cv.z & 0xffff), cv.z >> 16, (int16_t)(cv.w & 0xffff), cv.w >> 16 };
        int q[8] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w };
        int lv[8], du[8];
        int qBits8 = qBits - 8;
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            int level = c[i];
            int sign = level < 0 ? -1 : 1;
            int tmplevel = (int)((unsigned)abs(level) * (unsigned)q[i]);          // int32 wrap, dct.cpp:678
            level = (int)((unsigned)tmplevel + (unsigned)add) >> qBits;
            du[i] = (int)((unsigned)tmplevel - ((unsigned)level << qBits)) >> qBits8;
            sig += level != 0;
            level = (int)((unsigned)level * (unsigned)sign);
            level = clip16(level);
            lv[i] = NQUANT ? abs(level) : level;                                  // nquant: (int16)abs(clip), dct.cpp:711
        }
        int4 o;
        o.x = (lv[0] & 0xffff) | (lv[1] << 16); o.y = (lv[2] & 0xffff) | (lv[3] << 16);
        o.z = (lv[4] & 0xffff) | (lv[5] << 16); o.w = (lv[6] & 0xffff) | (lv[7] << 16);
        *(int4*)(qCoef + base) = o;
        if (STORE_DU)
        {
            *(int4*)(deltaU + base) = make_int4(du[0], du[1], du[2], du[3]);
            *(int4*)(deltaU + base + 4) = make_int4(du[4], du[5], du[6], du[7]);
        }
    }
    // segmented reduction over the lanes of one warp that share `blk`
    int lane = threadIdx.x & 31;
    if ((per & (per - 1)) == 0)
    {
        // usual case (numCoeff = 16/64/256/1024): a block is an aligned power-of-two lane group
        int G = per < 32 ? per : 32;
        int sum = group_sum(sig, G);
        if (live && (lane & (G - 1)) == 0 && sum) atomicAdd(numSig + blk, (uint32_t)sum);
    }
    else
    {
        unsigned peers = __match_any_sync(0xffffffffu, live ? blk : -1);
        int leader = __ffs(peers) - 1;
        int sum = 0;
        for (unsigned m = peers; m; m &= m - 1)
            sum += __shfl_sync(peers, sig, __ffs(m) - 1);
        if (live && lane == leader && sum) atomicAdd(numSig + blk, (uint32_t)sum);
    }
}

__global__ void __launch_bounds__(256)
dequant_normal_kernel(const int16_t* __restrict__ q, int16_t* __restrict__ coef, int num, int scale, int shift)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long pos = gid << 3;
    if (pos >= num) return;
    int4 v = __ldg((const int4*)(q + pos));
    int c[8] = { (int16_t)(v.x & 0xffff), v.x >> 16, (int16_t)(v.y & 0xffff), v.y >> 16,
                 (int16_t)(v.z & 0xffff), v.z >> 16, (int16_t)(v.w & 0xffff), v.w >> 16 };
    int add = 1 << (shift - 1);
#pragma unroll
    for (int i = 0; i < 8; i++)
        c[i] = clip16((int)((unsigned)c[i] * (unsigned)scale + (unsigned)add) >> shift);
    int4 o;
    o.x = (c[0] & 0xffff) | (c[1] << 16); o.y = (c[2] & 0xffff) | (c[3] << 16);
    o.z = (c[4] & 0xffff) | (c[5] << 16); o.w = (c[6] & 0xffff) | (c[7] << 16);
    *(int4*)(coef + pos) = o;
}

// 8 coefficients per thread (num % 8 == 0): 128-bit coefficient loads, two 128-bit table loads
__global__ void __launch_bounds__(256)
dequant_scaling_vec_kernel(const int16_t* __restrict__ q, const int32_t* __restrict__ dq, int16_t* __restrict__ coef,
                           int num, long long total, int per, int shift)
{
    long long pos = ((long long)blockIdx.x * blockDim.x + threadIdx.x) << 3;
    if (pos >= total) return;
    int tp = (int)(pos % num);
    int4 v = __ldg((const int4*)(q + pos));
    int4 t0 = __ldg((const int4*)(dq + tp)), t1 = __ldg((const int4*)(dq + tp + 4));
    int c[8] = { (int16_t)(v.x & 0xffff), v.x >> 16, (int16_t)(v.y & 0xffff), v.y >> 16,
                 (int16_t)(v.z & 0xffff), v.z >> 16, (int16_t)(v.w & 0xffff), v.w >> 16 };
    int d[8] = { t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w };
    shift += 4;
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        int prod = (int)((unsigned)c[i] * (unsigned)d[i]);
        c[i] = shift > per ? clip16((int)((unsigned)prod + (1u << (shift - per - 1))) >> (shift - per))
                           : clip16((int)((unsigned)clip16(prod) * (1u << (per - shift))));
    }
    int4 o;
    o.x = (c[0] & 0xffff) | (c[1] << 16); o.y = (c[2] & 0xffff) | (c[3] << 16);
    o.z = (c[4] & 0xffff) | (c[5] << 16); o.w = (c[6] & 0xffff) | (c[7] << 16);
    *(int4*)(coef + pos) = o;
}

__global__ void __launch_bounds__(256)
dequant_scaling_kernel(const int16_t* __restrict__ q, const int32_t* __restrict__ dq, int16_t* __restrict__ coef,
                       int num, long long total, int per, int shift)
{
    long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= total) return;
    int prod = (int)((unsigned)(int)q[pos] * (unsigned)dq[pos % num]);
    int r;
    shift += 4;                                                    // dct.cpp:644
    if (shift > per)
        r = clip16((int)((unsigned)prod + (1u << (shift - per - 1))) >> (shift - per));
    else
        r = clip16((int)((unsigned)clip16(prod) * (1u << (per - shift))));
    coef[pos] = (int16_t)r;
}

static int ilog2(int n) { int l = 0; while ((1 << l) < n) l++; return l; }

bool launch_dct_imma(x265b200_ctx* ctx, int N, const int16_t* src, intptr_t srcStride, const int32_t* off, int n,
                     int16_t* dst, int shift1, int shift2, cudaStream_t st, int dst4);    // transform_mma.cu
bool launch_idct_imma(x265b200_ctx* ctx, int N, const int16_t* src, int n, int16_t* dst, intptr_t dstStride,
                      const int32_t* off, int shift1, int shift2, cudaStream_t st, int dst4);

} // namespace b200

using namespace b200;

extern "C" int x265b200_dct_batch(x265b200_ctx* ctx, int kind, int N, const int16_t* src, intptr_t srcStride,
                                  const int32_t* off, int n, int16_t* dst, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (n < 0) return fail(ctx, X265B200_ERR_ARG, "dct: n < 0");
    if (n == 0) return X265B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int d8 = ctx->depth - 8;
#define FWD(NN, MODE, TN) fwd_kernel<NN, MODE><<<ceil_div(n, Tile<NN>::PER_CTA), TR_THREADS, 0, st>>>( \
        src, srcStride, off, n, dst, ilog2(TN) - 1 + d8, ilog2(TN) + 6, ctx->depth)
    if ((kind == X265B200_TR_DCT || (kind == X265B200_TR_DST && N == 4)) && ctx->dct_path != 1 &&
        launch_dct_imma(ctx, N, src, srcStride, off, n, dst, ilog2(N) - 1 + d8, ilog2(N) + 6, st, kind == X265B200_TR_DST))
    {
        // tensor-core path (transform_mma.cu)
    }
    else if (kind == X265B200_TR_DCT)
    {
        if (N == 4) FWD(4, MODE_DCT, 4); else if (N == 8) FWD(8, MODE_DCT, 8);
        else if (N == 16) FWD(16, MODE_DCT, 16); else if (N == 32) FWD(32, MODE_DCT, 32);
        else return fail(ctx, X265B200_ERR_ARG, "dct: N must be 4, 8, 16 or 32");
    }
    else if (kind == X265B200_TR_DST)
    {
        if (N != 4) return fail(ctx, X265B200_ERR_ARG, "dst: N must be 4");
        FWD(4, MODE_DST, 4);
    }
    else if (kind == X265B200_TR_LOWPASS)
    {
        if (N == 8) FWD(4, MODE_LOWPASS, 4); else if (N == 16) FWD(8, MODE_LOWPASS, 8);
        else if (N == 32) FWD(16, MODE_LOWPASS, 16);
        else return fail(ctx, X265B200_ERR_ARG, "lowpass_dct: N must be 8, 16 or 32");
    }
    else
        return fail(ctx, X265B200_ERR_ARG, "dct: unknown kind");
#undef FWD
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

extern "C" int x265b200_idct_batch(x265b200_ctx* ctx, int kind, int N, const int16_t* src, int n, int16_t* dst,
                                   intptr_t dstStride, const int32_t* off, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (n < 0) return fail(ctx, X265B200_ERR_ARG, "idct: n < 0");
    if (n == 0) return X265B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int s2 = 12 - (ctx->depth - 8);
#define INV(NN, MODE) inv_kernel<NN, MODE><<<ceil_div(n, Tile<NN>::PER_CTA), TR_THREADS, 0, st>>>(src, n, dst, dstStride, off, 7, s2)
    if ((kind == X265B200_TR_DCT || (kind == X265B200_TR_DST && N == 4)) && ctx->dct_path != 1 &&
        launch_idct_imma(ctx, N, src, n, dst, dstStride, off, 7, s2, st, kind == X265B200_TR_DST))
    {
        // tensor-core path (transform_mma.cu)
    }
    else if (kind == X265B200_TR_DCT)
    {
        if (N == 4) INV(4, MODE_DCT); else if (N == 8) INV(8, MODE_DCT);
        else if (N == 16) INV(16, MODE_DCT); else if (N == 32) INV(32, MODE_DCT);
        else return fail(ctx, X265B200_ERR_ARG, "idct: N must be 4, 8, 16 or 32");
    }
    else if (kind == X265B200_TR_DST)
    {
        if (N != 4) return fail(ctx, X265B200_ERR_ARG, "idst: N must be 4");
        INV(4, MODE_DST);
    }
    else
        return fail(ctx, X265B200_ERR_ARG, "idct: unknown kind");
#undef INV
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

namespace b200 {
// mode 0 = quant (stores deltaU), 1 = nquant, 2 = quant without storing deltaU (fused TU chain)
int launch_quant(x265b200_ctx* ctx, int mode, const int16_t* coef, const int32_t* quantCoeff, int32_t* deltaU, int16_t* qCoef,
                 int qBits, int add, int numCoeff, int n, uint32_t* numSig, cudaStream_t st)
{
    int per = numCoeff >> 3;
    const bool pow2 = (per & (per - 1)) == 0 && !(((uintptr_t)coef | (uintptr_t)qCoef | (uintptr_t)quantCoeff | (uintptr_t)deltaU) & 15);
    if (!pow2 || per > 32) B200_CUDA(ctx, cudaMemsetAsync(numSig, 0, (size_t)n * sizeof(uint32_t), st));
    if (pow2)
    {
        constexpr int NB = 4;
        long long threads = (long long)ceil_div(n, NB) * per;
        int grid = ceil_div(threads, 256);
        if (mode == 0)
            quant_multi_kernel<false, true, NB><<<grid, 256, 0, st>>>(coef, quantCoeff, deltaU, qCoef, qBits, add, numCoeff, n, numSig);
        else if (mode == 1)
            quant_multi_kernel<true, false, NB><<<grid, 256, 0, st>>>(coef, quantCoeff, nullptr, qCoef, qBits, add, numCoeff, n, numSig);
        else
            quant_multi_kernel<false, false, NB><<<grid, 256, 0, st>>>(coef, quantCoeff, nullptr, qCoef, qBits, add, numCoeff, n, numSig);
        B200_LAUNCH_CHECK(ctx);
        return X265B200_OK;
    }
    long long threads = (long long)n * (numCoeff >> 3);
    int grid = ceil_div(threads, 256);
    if (mode == 0)
        quant_kernel<false, true><<<grid, 256, 0, st>>>(coef, quantCoeff, deltaU, qCoef, qBits, add, numCoeff, n, numSig);
    else if (mode == 1)
        quant_kernel<true, false><<<grid, 256, 0, st>>>(coef, quantCoeff, nullptr, qCoef, qBits, add, numCoeff, n, numSig);
    else
        quant_kernel<false, false><<<grid, 256, 0, st>>>(coef, quantCoeff, nullptr, qCoef, qBits, add, numCoeff, n, numSig);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}
}

extern "C" int x265b200_quant_batch(x265b200_ctx* ctx, const int16_t* coef, const int32_t* quantCoeff, int32_t* deltaU,
                                    int16_t* qCoef, int qBits, int add, int numCoeff, int n, uint32_t* numSig,
                                    x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (numCoeff < 8 || (numCoeff & 7) || n < 0) return fail(ctx, X265B200_ERR_ARG, "quant: numCoeff must be a multiple of 8");
    if (n == 0) return X265B200_OK;
    return launch_quant(ctx, deltaU ? 0 : 1, coef, quantCoeff, deltaU, qCoef, qBits, add, numCoeff, n, numSig, (cudaStream_t)stream);
}

extern "C" int x265b200_dequant_normal_batch(x265b200_ctx* ctx, const int16_t* q, int16_t* coef, int num, int scale,
                                             int shift, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (num < 0 || (num & 7) || shift < 1) return fail(ctx, X265B200_ERR_ARG, "dequant_normal: num must be a multiple of 8");
    if (num == 0) return X265B200_OK;
    dequant_normal_kernel<<<ceil_div(num >> 3, 256), 256, 0, (cudaStream_t)stream>>>(q, coef, num, scale, shift);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

extern "C" int x265b200_dequant_scaling_batch(x265b200_ctx* ctx, const int16_t* q, const int32_t* dq, int16_t* coef,
                                              int num, int n, int per, int shift, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (num <= 0 || n < 0) return fail(ctx, X265B200_ERR_ARG, "dequant_scaling: bad size");
    if (n == 0) return X265B200_OK;
    long long total = (long long)num * n;
    if (!(num & 7) && !(((uintptr_t)q | (uintptr_t)coef | (uintptr_t)dq) & 15))
    {
        dequant_scaling_vec_kernel<<<ceil_div(total >> 3, 256), 256, 0, (cudaStream_t)stream>>>(q, dq, coef, num, total, per, shift);
        B200_LAUNCH_CHECK(ctx);
        return X265B200_OK;
    }
    dequant_scaling_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(q, dq, coef, num, total, per, shift);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}
