// intra.cu -- the lookahead's intra cost estimate per 8x8 lowres CU (reference encoder/slicetype.cpp:755-864,
// LookaheadTLD::lowresIntraEstimate) on top of the intra predictors of common/intrapred.cpp:31-204.
//
// The reference evaluates DC, planar and then walks the angular modes coarse to fine (5, 10 .. 30; best +-2; best +-1),
// one 8x8 prediction + SATD at a time.  A mode's cost does not depend on the walk, so here one warp owns a CU, every
// lane costs one mode (all 35 of them: prediction samples are computed straight from the neighbour arrays in shared
// memory, per 4x4 tile, and go into a scalar 4x4 Hadamard without ever forming the predicted block in memory), and
// lane 0 then replays the reference's decision sequence on the 35 costs -- same order, same strict-less updates.
#include "internal.h"

namespace b200 {

constexpr int IN_N = 8;                                  // X265_LOWRES_CU_SIZE, common.h:227
constexpr int IN_WARPS = 4;
constexpr int IN_COST_MAX = 1 << 28;                      // MotionEstimate::COST_MAX, motion.h:68

__constant__ uint8_t c_intraFilterFlags[35] = {          // constants.cpp:561-567
    0x38, 0x00,
    0x38, 0x30, 0x30, 0x30, 0x30, 0x30, 0x30, 0x20, 0x00, 0x20, 0x30, 0x30, 0x30, 0x30, 0x30, 0x30,
    0x38, 0x30, 0x30, 0x30, 0x30, 0x30, 0x30, 0x20, 0x00, 0x20, 0x30, 0x30, 0x30, 0x30, 0x30, 0x30,
    0x38 };
__constant__ int c_intraAngle[17] = { -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32 };     // intrapred.cpp:122
__constant__ int c_intraInvAngle[8] = { 4096, 1638, 910, 630, 482, 390, 315, 256 };                           // intrapred.cpp:123

// neighbour j as a horizontal mode sees it: above and left exchanged (intrapred.cpp:111-119)
__device__ __forceinline__ int intra_nb(const uint16_t* s, int hor, int j)
{
    if (!hor || j == 0) return s[j];
    return j <= 2 * IN_N ? s[2 * IN_N + j] : s[j - 2 * IN_N];
}
// reference sample i of the angular rule (intrapred.cpp:147-171): the above row, extended to the left by the left column
// projected along the inverse angle when the angle is negative
__device__ __forceinline__ int intra_ref(const uint16_t* s, int hor, int angleOffset, int i)
{
    if (i >= -1) return intra_nb(s, hor, i + 1);
    const int k = -2 - i;
    return intra_nb(s, hor, 2 * IN_N + ((128 + (k + 1) * c_intraInvAngle[-angleOffset - 1]) >> 8));
}

// predicted sample (r, c) of `mode` for an 8x8 block with edge filtering on (bFilter = cuSize <= 16)
struct IntraMode
{
    int mode, hor, angleOffset, angle, dc;
    const uint16_t* s;                                   // unfiltered or smoothed neighbours, as g_intraFilterFlags says
    __device__ int px(int r, int c, int pmax) const
    {
        if (mode == 0)      // planar, intrapred.cpp:87-100
            return ((IN_N - 1 - c) * s[2 * IN_N + 1 + r] + (IN_N - 1 - r) * s[1 + c] + (c + 1) * s[1 + IN_N] + (r + 1) * s[2 * IN_N + 1 + IN_N] + IN_N) >> 4;
        if (mode == 1)
        {                   // DC + edge smoothing, intrapred.cpp:53-85
            if (!r && !c) return (s[1] + s[2 * IN_N + 1] + 2 * dc + 2) >> 2;
            if (!r) return (s[1 + c] + 3 * dc + 2) >> 2;
            if (!c) return (s[2 * IN_N + 1 + r] + 3 * dc + 2) >> 2;
            return dc;
        }
        const int y = hor ? c : r, x = hor ? r : c;       // position in the un-flipped (vertical) frame
        if (!angle)
        {
            if (x) return intra_nb(s, hor, 1 + x);
            int v = (int)(int16_t)(intra_nb(s, hor, 1) + ((intra_nb(s, hor, 2 * IN_N + 1 + y) - intra_nb(s, hor, 0)) >> 1));
            return min(max(v, 0), pmax);
        }
        const int sum = (y + 1) * angle, off = sum >> 5, frac = sum & 31;
        int v = intra_ref(s, hor, angleOffset, off + x);
        if (frac) v = ((32 - frac) * v + frac * intra_ref(s, hor, angleOffset, off + x + 1) + 16) >> 5;
        return v;
    }
};

// sum of |H4 d H4^T| over a 4x4 tile of differences (twice the reference's satd_4x4, which halves an always even sum)
__device__ __forceinline__ int hadamard4x4_abs(int (&d)[16])
{
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        int a0 = d[4 * r] + d[4 * r + 1], a1 = d[4 * r] - d[4 * r + 1], a2 = d[4 * r + 2] + d[4 * r + 3], a3 = d[4 * r + 2] - d[4 * r + 3];
        d[4 * r] = a0 + a2; d[4 * r + 1] = a1 + a3; d[4 * r + 2] = a0 - a2; d[4 * r + 3] = a1 - a3;
    }
    int sum = 0;
#pragma unroll
    for (int c = 0; c < 4; c++)
    {
        int a0 = d[c] + d[4 + c], a1 = d[c] - d[4 + c], a2 = d[8 + c] + d[12 + c], a3 = d[8 + c] - d[12 + c];
        sum += abs(a0 + a2) + abs(a1 + a3) + abs(a0 - a2) + abs(a1 - a3);
    }
    return sum;
}

template<typename PIX>
__global__ void __launch_bounds__(IN_WARPS * 32)
lowres_intra_kernel(const PIX* __restrict__ plane, intptr_t stride, int widthInCU, int ncu, int penalty, int pmax,
                    int32_t* __restrict__ costOut, int32_t* __restrict__ modeOut)
{
    __shared__ uint16_t nb[IN_WARPS][2][4 * IN_N + 1 + 3];
    __shared__ uint16_t fencS[IN_WARPS][IN_N * IN_N];
    __shared__ int costs[IN_WARPS][36];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cu = blockIdx.x * IN_WARPS + warp;
    if (cu >= ncu) return;                               // whole warps leave; only __syncwarp below
    const int cuY = cu / widthInCU, cuX = cu - cuY * widthInCU;
    const PIX* cur = plane + (intptr_t)IN_N * cuY * stride + IN_N * cuX;
    const PIX* p = cur - stride - 1;
    uint16_t* s = nb[warp][0];
    uint16_t* f = nb[warp][1];
    // reference samples: top-left + 16 above, 16 left (slicetype.cpp:789-792)
    if (lane <= 2 * IN_N) s[lane] = p[lane];
    if (lane >= 1 && lane <= 2 * IN_N) s[2 * IN_N + lane] = p[(intptr_t)lane * stride];
    fencS[warp][lane] = cur[(intptr_t)(lane >> 3) * stride + (lane & 7)];
    fencS[warp][32 + lane] = cur[(intptr_t)(4 + (lane >> 3)) * stride + (lane & 7)];
    __syncwarp();
    // 1:2:1 smoothing, intrapred.cpp:31-51
    for (int i = lane; i <= 4 * IN_N; i += 32)
    {
        int v;
        if (i == 0) v = (2 * s[0] + s[1] + s[2 * IN_N + 1] + 2) >> 2;
        else if (i == 2 * IN_N || i == 4 * IN_N) v = s[i];
        else if (i == 2 * IN_N + 1) v = (2 * s[i] + s[0] + s[i + 1] + 2) >> 2;
        else v = (2 * s[i] + s[i - 1] + s[i + 1] + 2) >> 2;
        f[i] = (uint16_t)v;
    }
    int dc = IN_N;
    for (int i = 0; i < IN_N; i++) dc += s[1 + i] + s[2 * IN_N + 1 + i];
    dc /= 2 * IN_N;
    __syncwarp();

    // work item = (mode, 4x4 tile): 140 items over 32 lanes (five passes, the last 12 lanes wide) instead of 35 modes over 32 lanes (two passes
    // of four tiles each, the second 3 lanes wide); the four tiles of a mode sit in adjacent lanes and meet by two xor shuffles
    for (int i0 = 0; i0 < 140; i0 += 32)
    {
        const int item = i0 + lane;
        int sum = 0;
        if (item < 140)
        {
            const int mode = item >> 2, t = item & 3;
            IntraMode m;
            m.mode = mode; m.dc = dc;
            m.hor = mode >= 2 && mode < 18;
            m.angleOffset = mode < 2 ? 0 : (m.hor ? 10 - mode : mode - 26);
            m.angle = c_intraAngle[8 + m.angleOffset];
            // DC reads the unfiltered samples, planar the smoothed ones (cuSize >= 8), angular modes follow the flag table
            m.s = mode == 1 ? s : mode == 0 ? f : ((c_intraFilterFlags[mode] & IN_N) ? f : s);
            const int r0 = (t >> 1) * 4, c0 = (t & 1) * 4;
            int d[16];
#pragma unroll
            for (int i = 0; i < 16; i++)
                d[i] = (int)fencS[warp][(r0 + (i >> 2)) * IN_N + c0 + (i & 3)] - m.px(r0 + (i >> 2), c0 + (i & 3), pmax);
            sum = hadamard4x4_abs(d);
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        if (item < 140 && (lane & 3) == 0) costs[warp][item >> 2] = sum >> 1;
    }
    __syncwarp();
    if (lane == 0)
    {   // slicetype.cpp:796-841
        const int* c = costs[warp];
        int icost = IN_COST_MAX, imode = 0;
        if (c[1] < icost) { icost = c[1]; imode = 1; }
        if (c[0] < icost) { icost = c[0]; imode = 0; }
        int acost = IN_COST_MAX, amode = 4;
        for (int mode = 5; mode < 35; mode += 5)
            if (c[mode] < acost) { acost = c[mode]; amode = mode; }
        for (int dist = 2; dist >= 1; dist--)
        {
            const int minus = amode - dist, plus = amode + dist;    // both around the best before this round
            if (c[minus] < acost) { acost = c[minus]; amode = minus; }
            if (c[plus] < acost) { acost = c[plus]; amode = plus; }
        }
        if (acost < icost) { icost = acost; imode = amode; }
        costOut[cu] = icost + penalty;
        modeOut[cu] = imode;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// All 35 luma predictions of a TU, N = 4 .. 32, the way the analysis forms them before costing the modes with sa8d
// (reference encoder/search.cpp:1703-1727): same per-sample rules as above with N as a run-time value.  One CTA per TU:
// neighbours and their smoothed copy in shared memory, the threads sweep the 35 * N * N output samples.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int intra_nb_n(const uint16_t* s, int N, int hor, int j)
{
    if (!hor || j == 0) return s[j];
    return j <= 2 * N ? s[2 * N + j] : s[j - 2 * N];
}
__device__ __forceinline__ int intra_ref_n(const uint16_t* s, int N, int hor, int angleOffset, int i)
{
    if (i >= -1) return intra_nb_n(s, N, hor, i + 1);
    const int k = -2 - i;
    return intra_nb_n(s, N, hor, 2 * N + ((128 + (k + 1) * c_intraInvAngle[-angleOffset - 1]) >> 8));
}
__device__ __forceinline__ int intra_px_n(const uint16_t* s, int N, int lgN, int mode, int bFilter, int dc, int r, int c, int pmax)
{
    if (mode == 0)
        return ((N - 1 - c) * s[2 * N + 1 + r] + (N - 1 - r) * s[1 + c] + (c + 1) * s[1 + N] + (r + 1) * s[3 * N + 1] + N) >> (lgN + 1);
    if (mode == 1)
    {
        if (!bFilter) return dc;
        if (!r && !c) return (s[1] + s[2 * N + 1] + 2 * dc + 2) >> 2;
        if (!r) return (s[1 + c] + 3 * dc + 2) >> 2;
        if (!c) return (s[2 * N + 1 + r] + 3 * dc + 2) >> 2;
        return dc;
    }
    const int hor = mode < 18;
    const int angleOffset = hor ? 10 - mode : mode - 26;
    const int angle = c_intraAngle[8 + angleOffset];
    const int y = hor ? c : r, x = hor ? r : c;
    if (!angle)
    {
        if (x || !bFilter) return intra_nb_n(s, N, hor, 1 + x);
        int v = (int)(int16_t)(intra_nb_n(s, N, hor, 1) + ((intra_nb_n(s, N, hor, 2 * N + 1 + y) - intra_nb_n(s, N, hor, 0)) >> 1));
        return min(max(v, 0), pmax);
    }
    const int sum = (y + 1) * angle, off = sum >> 5, frac = sum & 31;
    int v = intra_ref_n(s, N, hor, angleOffset, off + x);
    if (frac) v = ((32 - frac) * v + frac * intra_ref_n(s, N, hor, angleOffset, off + x + 1) + 16) >> 5;
    return v;
}

template<typename PIX>
__global__ void __launch_bounds__(256)
intra_pred_all_kernel(const PIX* __restrict__ neighbours, int N, int lgN, int pmax, PIX* __restrict__ dst)
{
    __shared__ uint16_t s[4 * 32 + 1 + 3], f[4 * 32 + 1 + 3];
    const int tu = blockIdx.x, L = 4 * N + 1;
    const PIX* nb = neighbours + (size_t)tu * L;
    for (int i = threadIdx.x; i < L; i += blockDim.x) s[i] = nb[i];
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += blockDim.x)
    {   // intrapred.cpp:31-51
        int v;
        if (i == 0) v = (2 * s[0] + s[1] + s[2 * N + 1] + 2) >> 2;
        else if (i == 2 * N || i == 4 * N) v = s[i];
        else if (i == 2 * N + 1) v = (2 * s[i] + s[0] + s[i + 1] + 2) >> 2;
        else v = (2 * s[i] + s[i - 1] + s[i + 1] + 2) >> 2;
        f[i] = (uint16_t)v;
    }
    int dc = N;
    for (int i = 0; i < N; i++) dc += s[1 + i] + s[2 * N + 1 + i];
    dc /= 2 * N;
    __syncthreads();
    const int bFilter = N <= 16, NN = N * N;
    PIX* out = dst + (size_t)tu * 35 * NN;
    // Main reference array of every angular mode with a non-zero angle, rm[mode - 2][i + N] = ref(i) for i = -N .. 2N: the left / top swap of the
    // horizontal modes, the smoothed-or-plain choice of the filter flag table and the projection of the side array for negative angles
    // (intrapred.cpp:150-180) are resolved ONCE per TU and mode here instead of once per predicted sample.
    __shared__ uint16_t rm[33][3 * 32 + 2];
    const int RL = 3 * N + 2;
    for (int e = threadIdx.x; e < 33 * RL; e += blockDim.x)
    {
        const int mode = 2 + e / RL, i = e % RL - N;
        const int hor = mode < 18, angleOffset = hor ? 10 - mode : mode - 26;
        const uint16_t* src = (c_intraFilterFlags[mode] & N) ? f : s;
        int v = 0;
        bool need = c_intraAngle[8 + angleOffset] != 0 && i < 2 * N && (i >= -1 || angleOffset < 0);
        if (need && i < -1) need = ((128 + (-1 - i) * c_intraInvAngle[-angleOffset - 1]) >> 8) <= 2 * N;     // projections past the side array are never read
        if (need) v = intra_ref_n(src, N, hor, angleOffset, i);
        rm[mode - 2][e % RL] = (uint16_t)v;
    }
    __syncthreads();
    // a thread forms four horizontally adjacent samples of one (mode, row) and stores them as one word pair: a warp writes 256 (128) contiguous
    // bytes instead of 64 (32)
    for (int o4 = threadIdx.x; o4 < 35 * NN / 4; o4 += blockDim.x)
    {
        const int o = o4 << 2, mode = o >> (2 * lgN), rc = o & (NN - 1), r = rc >> lgN, c = rc & (N - 1);
        int v[4];
        const int hor = mode >= 2 && mode < 18;
        const int angle = mode < 2 ? 0 : c_intraAngle[8 + (hor ? 10 - mode : mode - 26)];
        if (angle == 0)
        {   // planar, DC, pure horizontal / vertical (with their edge filters): the per-sample rule
            const uint16_t* src = mode == 1 ? s : mode == 0 ? (N >= 8 ? f : s) : ((c_intraFilterFlags[mode] & N) ? f : s);
#pragma unroll
            for (int i = 0; i < 4; i++) v[i] = intra_px_n(src, N, lgN, mode, bFilter, dc, r, c + i, pmax);
        }
        else
        {
            const uint16_t* ref = rm[mode - 2] + N;
            if (!hor)
            {   // the row shares offset and fraction: five neighbouring reference samples, four interpolations
                const int sum = (r + 1) * angle, off = sum >> 5, frac = sum & 31;
                int a = ref[off + c];
#pragma unroll
                for (int i = 0; i < 4; i++) { const int b = ref[off + c + i + 1]; v[i] = ((32 - frac) * a + frac * b + 16) >> 5; a = b; }
            }
            else
            {
#pragma unroll
                for (int i = 0; i < 4; i++)
                {
                    const int sum = (c + i + 1) * angle, off = sum >> 5, frac = sum & 31;
                    v[i] = ((32 - frac) * ref[off + r] + frac * ref[off + r + 1] + 16) >> 5;
                }
            }
        }
        if (sizeof(PIX) == 2) *(uint2*)(out + o) = make_uint2((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16));
        else *(uint32_t*)(out + o) = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
    }
}

// The three intra slots of the table as they are called one TU at a time (reference common/intrapred.cpp:31-233), batched over n TUs:
//   kind 0  intra_pred[mode](dst, dstStride, srcPix, mode, bFilter): one N x N prediction from the (4N + 1) neighbours the caller chose
//   kind 1  intra_filter(samples, filtered): the 1:2:1 smoothing of the neighbour array
//   kind 2  intra_pred_allangs(dst, refPix, filtPix, bLuma): modes 2 .. 34, each from the smoothed or the plain neighbours as
//           g_intraFilterFlags says, horizontal modes left un-flipped (the reference transposes them back, intrapred.cpp:217-231)
template<typename PIX>
__global__ void __launch_bounds__(256)
intra_slot_kernel(int kind, int N, int lgN, int mode, int bFilter, const PIX* __restrict__ src, const PIX* __restrict__ filt, int pmax, PIX* __restrict__ dst)
{
    __shared__ uint16_t s[4 * 32 + 1 + 3], f[4 * 32 + 1 + 3];
    const int tu = blockIdx.x, L = 4 * N + 1, NN = N * N;
    for (int i = threadIdx.x; i < L; i += blockDim.x) { s[i] = src[(size_t)tu * L + i]; if (filt) f[i] = filt[(size_t)tu * L + i]; }
    __syncthreads();
    if (kind == 1)
    {
        for (int i = threadIdx.x; i < L; i += blockDim.x)
        {   // intrapred.cpp:31-51
            int v;
            if (i == 0) v = (2 * s[0] + s[1] + s[2 * N + 1] + 2) >> 2;
            else if (i == 2 * N || i == 4 * N) v = s[i];
            else if (i == 2 * N + 1) v = (2 * s[i] + s[0] + s[i + 1] + 2) >> 2;
            else v = (2 * s[i] + s[i - 1] + s[i + 1] + 2) >> 2;
            dst[(size_t)tu * L + i] = (PIX)v;
        }
        return;
    }
    if (kind == 0)
    {
        int dc = N;
        if (mode == 1)
        {
            for (int i = 0; i < N; i++) dc += s[1 + i] + s[2 * N + 1 + i];
            dc /= 2 * N;
        }
        for (int o = threadIdx.x; o < NN; o += blockDim.x)
            dst[(size_t)tu * NN + o] = (PIX)intra_px_n(s, N, lgN, mode, bFilter, dc, o >> lgN, o & (N - 1), pmax);
        return;
    }
    for (int o = threadIdx.x; o < 33 * NN; o += blockDim.x)
    {
        const int m = 2 + (o >> (2 * lgN)), rc = o & (NN - 1), r = rc >> lgN, c = rc & (N - 1);
        const uint16_t* nb = (c_intraFilterFlags[m] & N) ? f : s;
        const bool hor = m < 18;
        dst[(size_t)tu * 33 * NN + o] = (PIX)intra_px_n(nb, N, lgN, m, bFilter, 0, hor ? c : r, hor ? r : c, pmax);
    }
}

} // namespace b200

using namespace b200;

extern "C" int x265b200_intra_slot_batch(x265b200_ctx* ctx, int kind, int N, int mode, int bFilter, const void* src, const void* filt, int n, void* dst,
                                         x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if ((N != 4 && N != 8 && N != 16 && N != 32) || n < 0 || kind < 0 || kind > 2 || (kind == 0 && (mode < 0 || mode > 34)) || (kind == 2 && !filt))
        return fail(ctx, X265B200_ERR_ARG, "intra_slot: bad size / kind / mode");
    if (n == 0) return X265B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int lgN = N == 4 ? 2 : N == 8 ? 3 : N == 16 ? 4 : 5, pmax = (1 << ctx->depth) - 1;
    if (ctx->pixbytes == 1) intra_slot_kernel<uint8_t><<<n, 256, 0, st>>>(kind, N, lgN, mode, bFilter, (const uint8_t*)src, (const uint8_t*)filt, pmax, (uint8_t*)dst);
    else intra_slot_kernel<uint16_t><<<n, 256, 0, st>>>(kind, N, lgN, mode, bFilter, (const uint16_t*)src, (const uint16_t*)filt, pmax, (uint16_t*)dst);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

extern "C" int x265b200_lowres_intra_batch(x265b200_ctx* ctx, const void* plane, intptr_t stride, int widthInCU, int heightInCU, int penalty,
                                           int32_t* cost, int32_t* mode, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (widthInCU < 0 || heightInCU < 0 || (long long)widthInCU * heightInCU > 0x7fffffff) return fail(ctx, X265B200_ERR_ARG, "lowres_intra: bad geometry");
    const int ncu = widthInCU * heightInCU;
    if (ncu == 0) return X265B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int pmax = (1 << ctx->depth) - 1;
    if (ctx->pixbytes == 1)
        lowres_intra_kernel<uint8_t><<<ceil_div(ncu, IN_WARPS), IN_WARPS * 32, 0, st>>>((const uint8_t*)plane, stride, widthInCU, ncu, penalty, pmax, cost, mode);
    else
        lowres_intra_kernel<uint16_t><<<ceil_div(ncu, IN_WARPS), IN_WARPS * 32, 0, st>>>((const uint16_t*)plane, stride, widthInCU, ncu, penalty, pmax, cost, mode);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

extern "C" int x265b200_intra_pred_batch(x265b200_ctx* ctx, int N, const void* neighbours, int n, void* dst, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if ((N != 4 && N != 8 && N != 16 && N != 32) || n < 0) return fail(ctx, X265B200_ERR_ARG, "intra_pred: N must be 4, 8, 16 or 32");
    if (n == 0) return X265B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int lgN = N == 4 ? 2 : N == 8 ? 3 : N == 16 ? 4 : 5, pmax = (1 << ctx->depth) - 1;
    if (((uintptr_t)dst & 7)) return fail(ctx, X265B200_ERR_ARG, "intra_pred: dst must be 8-byte aligned");
    const int threads = N == 4 ? 160 : 256;                 // 35 * N * N / 4 four-sample groups per TU: 140 at N = 4
    if (ctx->pixbytes == 1) intra_pred_all_kernel<uint8_t><<<n, threads, 0, st>>>((const uint8_t*)neighbours, N, lgN, pmax, (uint8_t*)dst);
    else intra_pred_all_kernel<uint16_t><<<n, threads, 0, st>>>((const uint16_t*)neighbours, N, lgN, pmax, (uint16_t*)dst);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}
