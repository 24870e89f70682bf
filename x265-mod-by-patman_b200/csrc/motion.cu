// motion.cu -- MotionEstimate::motionEstimate for a batch of prediction units, searchMethod DIA / HEX / STAR / FULL, luma only
// (reference encoder/motion.cpp:923-1013 start point, :1016-1138 / :1593-1637 integer search, :1643-1773 sub-pel refinement and the
// zero-vector last chance; the setSourcePU variant of motion.cpp:166-189: one slice, no vertical restriction).
//
// The reference runs this per PU as a chain of data-dependent steps.  Across thousands of PUs the steps line up: every
// PU measures its predictor candidates, then searches its window, then takes the same number of half-pel and
// quarter-pel rounds (SubpelWorkload, motion.cpp:48-58).  So the batch advances in lock step, each step one launch over
// all PUs, with the per-PU decisions (COPY2_IF_LT chains, early `break`s, the bcost == 0 exits) kept in small state
// arrays and taken by one thread per PU between the heavy launches:
//     start_gen -> subpel_cmp_batch(SAD) -> start_select -> me_pattern_batch | me_full_batch -> [round: select+gen -> subpel_cmp_batch] x R -> finish
// The heavy launches are the library's own batched entries (fused interpolation + SAD/SATD, exhaustive search).
// A PU that left the chain early (zero residual) or whose refinement loop broke keeps producing harmless candidates at
// its current vector so the launches stay dense; its state no longer changes.
#include "internal.h"
#include "device_util.cuh"
#include "tile_kernels.cuh"

namespace b200 {

enum { ME_FIN = 1, ME_SKIP = 2, ME_STOP = 4 };
constexpr int ME_MAX_CAND = 16;

// one refinement round: `dirs` neighbours at distance `step` (2 = half pel, 1 = quarter pel) around the running best
struct MeRound { int step, dirs, remeasure, newPhase, zeroSlot, K; };

__constant__ int c_square1[9][2] = { {0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {-1, 1}, {1, -1}, {1, 1} };   // motion.cpp:67

struct MeState
{
    int32_t* pmv;       // 2n  clipped predictor, quarter pel
    int32_t* bestpre;   // 2n
    int32_t* bprecost;  // n
    int32_t* bmv;       // 2n  full pel until the search is over, quarter pel afterwards
    int32_t* bcost;     // n
    int32_t* flags;     // n
    int32_t* eff;       // 4n  search window handed to the exhaustive search (empty for finished PUs)
    int32_t* candOff;   // n * KMAX
    int32_t* candFrac;  // n * KMAX
    int32_t* candCost;  // n * KMAX
    int32_t* candOffC;  // n * KMAX  chroma block of the candidate (chroma term on)
    int32_t* candFracC; // n * KMAX  xFrac | yFrac << 4 in eighths, -1 = costed without the chroma term
};

// chroma residual term of subpelCompare (bChromaSATD, motion.cpp:1805-1865); on == 0: luma only
struct MeChroma { int on, hshift, vshift; intptr_t strideRC; const int32_t* offRC; };

__device__ __forceinline__ int clip3(int lo, int hi, int v) { return v < lo ? lo : v > hi ? hi : v; }
__device__ __forceinline__ int mvcost(const uint16_t* tab, int mvpx, int mvpy, int qx, int qy)
{
    return (uint16_t)(tab[qx - mvpx] + tab[qy - mvpy]);       // bitcost.h:56
}
// pitch == 0: full-resolution reference, candidate = (integer offset, xFrac | yFrac << 4) for the interpolating kernel.
// pitch > 0: lowres reference of four half-pel planes `pitch` samples apart; candidate = the two plane blocks whose
// rounded average is the prediction (ReferencePlanes::lowresQPelCost, lowres.h:95-119), equal for half / full-pel vectors.
__device__ __forceinline__ void put_cand(const MeState& s, size_t slot, int base, intptr_t strideR, int pitch, int qx, int qy,
                                         const MeChroma& c = MeChroma{ 0, 0, 0, 0, nullptr }, int baseC = 0, bool chromaTerm = true)
{
    if (c.on)
    {
        const int mvx = (int)((unsigned)qx << (1 - c.hshift)), mvy = (int)((unsigned)qy << (1 - c.vshift));
        s.candOffC[slot] = chromaTerm ? baseC + (mvx >> 3) + (mvy >> 3) * (int)c.strideRC : baseC;
        s.candFracC[slot] = chromaTerm ? (mvx & 7) | ((mvy & 7) << 4) : -1;
    }
    if (!pitch)
    {
        s.candOff[slot] = base + (qx >> 2) + (qy >> 2) * (int)strideR;      // subpelCompare, motion.cpp:1777-1781
        s.candFrac[slot] = (qx & 3) | ((qy & 3) << 4);
        return;
    }
    const int a = ((qy & 2) | ((qx & 2) >> 1)) * pitch + base + (qx >> 2) + (qy >> 2) * (int)strideR;
    int b = a;
    if ((qx | qy) & 1)
    {
        const int bx = qx + (qx & 1), by = qy + (qy & 1);
        b = ((by & 2) | ((bx & 2) >> 1)) * pitch + base + (bx >> 2) + (by >> 2) * (int)strideR;
    }
    s.candOff[slot] = a;
    s.candFrac[slot] = b;
}

// cost[i] = SAD / SATD(fenc block i / K, (A_i + B_i + 1) >> 1): G lanes share a candidate, a lane takes 4x4 tiles
template<typename T, int OP>
__global__ void __launch_bounds__(256)
lowres_cmp_kernel(const T* __restrict__ fenc, intptr_t sf, const T* __restrict__ planes, intptr_t sr, const int32_t* __restrict__ offF,
                  const int32_t* __restrict__ offA, const int32_t* __restrict__ offB, int K, int ncand, int w, int h, int G,
                  int32_t* __restrict__ cost)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lg = __ffs(G) - 1;
    const int cand = (int)(gid >> lg), l = (int)gid & (G - 1);
    const bool live = cand < ncand;
    const int tw = w >> 2, T4 = tw * (h >> 2);
    int acc = 0;
    if (live)
    {
        const T* f = fenc + offF[cand / K];
        const T* a = planes + offA[cand];
        const T* b = planes + offB[cand];
        for (int t = l; t < T4; t += G)
        {
            const int ty = t / tw, tx = t - ty * tw;
            uint32_t flo[4], fhi[4], alo[4], ahi[4], blo[4], bhi[4];
            load_tile4x4(f + (intptr_t)(ty << 2) * sf + (tx << 2), sf, flo, fhi);
            load_tile4x4(a + (intptr_t)(ty << 2) * sr + (tx << 2), sr, alo, ahi);
            load_tile4x4(b + (intptr_t)(ty << 2) * sr + (tx << 2), sr, blo, bhi);
#pragma unroll
            for (int r = 0; r < 4; r++)
            {   // pixelavg_pp (pixel.cpp:586-594) on packed pairs: samples < 2^15, so the halves cannot carry into each other
                alo[r] = ((alo[r] + blo[r] + 0x00010001u) >> 1) & 0x7fff7fffu;
                ahi[r] = ((ahi[r] + bhi[r] + 0x00010001u) >> 1) & 0x7fff7fffu;
            }
            tile4_accumulate<OP, int>(flo, fhi, alo, ahi, acc);
        }
    }
    acc = group_sum(acc, G);
    if (live && l == 0) cost[cand] = acc;
}

static int launch_lowres_cmp(x265b200_ctx* ctx, int op, int w, int h, const void* fenc, intptr_t sf, const void* planes, intptr_t sr,
                             const int32_t* offF, const int32_t* offA, const int32_t* offB, int K, int n, int32_t* cost, cudaStream_t st)
{
    const int T4 = (w >> 2) * (h >> 2), per = T4 >= 16 ? 4 : 2;
    int G = 1;
    while (G * 2 * per <= T4 && G < 32) G <<= 1;
    const long long cands = (long long)n * K;
    if (cands * G > 0x7fffffffLL * 256) return fail(ctx, X265B200_ERR_ARG, "lowres cost: too many candidates");
    if (cands > 0x7fffffff) return fail(ctx, X265B200_ERR_ARG, "lowres cost: too many candidates");
    const int grid = ceil_div(cands * G, 256);
#define LC(T, OP_) lowres_cmp_kernel<T, OP_><<<grid, 256, 0, st>>>((const T*)fenc, sf, (const T*)planes, sr, offF, offA, offB, K, (int)cands, w, h, G, cost)
    if (ctx->pixbytes == 1) { if (op == X265B200_SAD) LC(uint8_t, OP_SAD); else LC(uint8_t, OP_SATD); }
    else { if (op == X265B200_SAD) LC(uint16_t, OP_SAD); else LC(uint16_t, OP_SATD); }
#undef LC
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

__global__ void me_start_gen(int n, int nc, int K, const int32_t* __restrict__ qmvp, const int32_t* __restrict__ mvc,
                             const int32_t* __restrict__ range, const int32_t* __restrict__ offR, intptr_t strideR, int pitch, MeState s, MeChroma c)
{
    int pu = blockIdx.x * blockDim.x + threadIdx.x;
    if (pu >= n) return;
    const int baseC = c.on ? c.offRC[pu] : 0;
    const int qminx = range[4 * pu] * 4, qminy = range[4 * pu + 1] * 4, qmaxx = range[4 * pu + 2] * 4, qmaxy = range[4 * pu + 3] * 4;
    const int px = clip3(qminx, qmaxx, qmvp[2 * pu]), py = clip3(qminy, qmaxy, qmvp[2 * pu + 1]);
    s.pmv[2 * pu] = px; s.pmv[2 * pu + 1] = py;
    const int base = offR[pu];
    const size_t o = (size_t)pu * K;
    put_cand(s, o, base, strideR, pitch, px, py, c, baseC, true);
    put_cand(s, o + 1, base, strideR, pitch, ((px + 2) >> 2) * 4, ((py + 2) >> 2) * 4, c, baseC, false);    // plain sad()
    put_cand(s, o + 2, base, strideR, pitch, 0, 0, c, baseC, false);                                          // plain sad()
    for (int i = 0; i < nc; i++)
        put_cand(s, o + 3 + i, base, strideR, pitch, clip3(qminx, qmaxx, mvc[((size_t)pu * nc + i) * 2]), clip3(qminy, qmaxy, mvc[((size_t)pu * nc + i) * 2 + 1]),
                 c, baseC, true);
}

__global__ void me_start_select(int n, int nc, int K, const int32_t* __restrict__ qmvp, const int32_t* __restrict__ mvc,
                                const int32_t* __restrict__ range, const uint16_t* __restrict__ tab, MeState s,
                                int32_t* __restrict__ outQMv, int32_t* __restrict__ outCost)
{
    int pu = blockIdx.x * blockDim.x + threadIdx.x;
    if (pu >= n) return;
    const int mvpx = qmvp[2 * pu], mvpy = qmvp[2 * pu + 1];
    const int minx = range[4 * pu], miny = range[4 * pu + 1], maxx = range[4 * pu + 2], maxy = range[4 * pu + 3];
    const int pmvx = s.pmv[2 * pu], pmvy = s.pmv[2 * pu + 1];
    const int32_t* cost = s.candCost + (size_t)pu * K;
    // motion.cpp:954-975
    int bestprex = pmvx, bestprey = pmvy, bprecost = cost[0];
    int bmvx = (pmvx + 2) >> 2, bmvy = (pmvy + 2) >> 2, bcost = bprecost;
    if ((pmvx | pmvy) & 3) bcost = cost[1] + mvcost(tab, mvpx, mvpy, bmvx * 4, bmvy * 4);
    if (pmvx | pmvy)
    {   // :978-988
        int c = cost[2] + mvcost(tab, mvpx, mvpy, 0, 0);
        if (c < bcost) { bcost = c; bmvx = 0; bmvy = max(min(0, maxy), miny); }
    }
    for (int i = 0; i < nc; i++)
    {   // :992-1004
        int mx = clip3(minx * 4, maxx * 4, mvc[((size_t)pu * nc + i) * 2]), my = clip3(miny * 4, maxy * 4, mvc[((size_t)pu * nc + i) * 2 + 1]);
        if ((mx | my) && (mx != pmvx || my != pmvy) && (mx != bestprex || my != bestprey))
        {
            int c = cost[3 + i] + mvcost(tab, mvpx, mvpy, mx, my);
            if (c < bprecost) { bprecost = c; bestprex = mx; bestprey = my; }
        }
    }
    s.bestpre[2 * pu] = bestprex; s.bestpre[2 * pu + 1] = bestprey; s.bprecost[pu] = bprecost;
    s.bmv[2 * pu] = bmvx; s.bmv[2 * pu + 1] = bmvy; s.bcost[pu] = bcost;
    const bool fin = bcost == 0;                               // :1008-1012
    s.flags[pu] = fin ? ME_FIN : 0;
    if (fin)
    {
        outQMv[2 * pu] = bmvx * 4; outQMv[2 * pu + 1] = bmvy * 4;
        outCost[pu] = mvcost(tab, mvpx, mvpy, bmvx * 4, bmvy * 4);
    }
    s.eff[4 * pu] = fin ? 1 : minx; s.eff[4 * pu + 1] = miny; s.eff[4 * pu + 2] = fin ? 0 : maxx; s.eff[4 * pu + 3] = maxy;
}

// mode bit 0: the step after the integer search (:1643-1666); bit 1: take the decisions of round `prev` (:1700-1757);
// bit 2: emit the candidates of round `next`; bit 3: zero-vector last chance and outputs (:1762-1772)
__global__ void me_round_kernel(int n, int mode, MeRound prev, MeRound next, const int32_t* __restrict__ qmvp,
                                const int32_t* __restrict__ range, const int32_t* __restrict__ offR, intptr_t strideR, int pitch,
                                const uint16_t* __restrict__ tab, MeState s, int32_t* __restrict__ outQMv, int32_t* __restrict__ outCost, MeChroma c)
{
    int pu = blockIdx.x * blockDim.x + threadIdx.x;
    if (pu >= n) return;
    const int baseC = c.on ? c.offRC[pu] : 0;
    int flags = s.flags[pu];
    const int mvpx = qmvp[2 * pu], mvpy = qmvp[2 * pu + 1];
    const int qminy = range[4 * pu + 1] * 4, qmaxy = range[4 * pu + 3] * 4;
    int bmvx = s.bmv[2 * pu], bmvy = s.bmv[2 * pu + 1], bcost = s.bcost[pu];
    int zcost = 0;

    if ((mode & 1) && !(flags & ME_FIN))
    {
        const int bprecost = s.bprecost[pu];
        if (bprecost < bcost) { bmvx = s.bestpre[2 * pu]; bmvy = s.bestpre[2 * pu + 1]; bcost = bprecost; }
        else { bmvx *= 4; bmvy *= 4; }
        if (!bcost) { bcost = mvcost(tab, mvpx, mvpy, bmvx, bmvy); flags |= ME_SKIP; }
    }
    if (mode & 2)
    {
        const int32_t* cost = s.candCost + (size_t)pu * prev.K;
        if (prev.zeroSlot) zcost = cost[prev.K - 1];
        if (prev.newPhase) flags &= ~ME_STOP;
        if (!(flags & (ME_FIN | ME_SKIP | ME_STOP)))
        {
            if (prev.remeasure) bcost = cost[0] + mvcost(tab, mvpx, mvpy, bmvx, bmvy);
            int bdir = 0;
            for (int i = 1; i <= prev.dirs; i++)
            {
                int qx = bmvx + c_square1[i][0] * prev.step, qy = bmvy + c_square1[i][1] * prev.step;
                if (qy < qminy || qy > qmaxy) continue;
                int c = cost[i] + mvcost(tab, mvpx, mvpy, qx, qy);
                if (c < bcost) { bcost = c; bdir = i; }
            }
            if (bdir) { bmvx += c_square1[bdir][0] * prev.step; bmvy += c_square1[bdir][1] * prev.step; }
            else if (prev.dirs) flags |= ME_STOP;
        }
    }
    if (mode & 4)
    {
        const size_t o = (size_t)pu * next.K;
        const int base = offR[pu];
        const bool live = !(flags & ME_FIN);                   // a finished PU's bmv is still in full pels: park it on the zero vector
        const int cx = live ? bmvx : 0, cy = live ? bmvy : 0;
        put_cand(s, o, base, strideR, pitch, cx, cy, c, baseC);
        for (int i = 1; i <= next.dirs; i++)
        {
            int qx = cx + c_square1[i][0] * next.step, qy = cy + c_square1[i][1] * next.step;
            if (!live || qy < qminy || qy > qmaxy) { qx = cx; qy = cy; }      // never measured by the reference: stay on a valid block
            put_cand(s, o + i, base, strideR, pitch, qx, qy, c, baseC);
        }
        if (next.zeroSlot) put_cand(s, o + next.K - 1, base, strideR, pitch, 0, 0, c, baseC);
    }
    if ((mode & 8) && !(flags & ME_FIN))
    {
        if (bmvx | bmvy)
        {
            int c = zcost + mvcost(tab, mvpx, mvpy, 0, 0);
            if (c <= bcost) { bmvx = 0; bmvy = 0; }             // the returned cost stays the winner's
        }
        outQMv[2 * pu] = bmvx; outQMv[2 * pu + 1] = bmvy; outCost[pu] = bcost;
    }
    s.bmv[2 * pu] = bmvx; s.bmv[2 * pu + 1] = bmvy; s.bcost[pu] = bcost; s.flags[pu] = flags;
}

} // namespace b200

using namespace b200;

// pitch == 0: full-resolution reference (interpolating sub-pel costs); pitch > 0: lowres reference, four half-pel planes
struct MeChromaArgs { const void* fencCb; const void* fencCr; intptr_t strideFC; const void* refCb; const void* refCr; const int32_t* offFC; };

static int motion_chain(x265b200_ctx* ctx, int pitch, const MeChroma& chroma, const MeChromaArgs& cargs,
                        int searchMethod, int w, int h, int merange, int subpelRefine,
                        const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR,
                        const int32_t* offF, const int32_t* offR, const int32_t* range, const int32_t* qmvp,
                        int numCand, const int32_t* mvc, const uint16_t* costTab, int n,
                        int32_t* outQMv, int32_t* outCost, x265b200_stream stream, const uint32_t* sums = nullptr, size_t sumsPitch = 0)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (subpelRefine < 0 || subpelRefine > 7 || numCand < 0 || numCand > ME_MAX_CAND || n < 0 || (numCand && !mvc))
        return fail(ctx, X265B200_ERR_ARG, "motion_estimate: bad arguments");
    if (searchMethod < X265B200_ME_DIA || searchMethod > X265B200_ME_FULL)
        return fail(ctx, X265B200_ERR_ARG, "motion_estimate: unknown search method");
    if (searchMethod == X265B200_ME_SEA && !sums)
        return fail(ctx, X265B200_ERR_ARG, "motion_estimate: X265_SEA needs the integral planes (x265b200_motion_estimate_sea_batch)");
    if (n == 0) return X265B200_OK;
    cudaStream_t st = (cudaStream_t)stream;

    // SubpelWorkload, motion.cpp:48-58
    static const int workload[8][5] = { { 1, 4, 0, 4, 0 }, { 1, 4, 1, 4, 0 }, { 1, 4, 1, 4, 1 }, { 2, 4, 1, 4, 1 },
                                        { 2, 4, 2, 4, 1 }, { 1, 8, 1, 8, 1 }, { 2, 8, 1, 8, 1 }, { 2, 8, 2, 8, 1 } };
    const int* wl = workload[subpelRefine];
    MeRound rounds[8]; int ops[8]; int R = 0;
    if (pitch)
    {   // lowres reference (:1667-1698): one SAD half-pel step, SATD re-measure, one SATD quarter-pel step
        rounds[R] = MeRound{ 2, wl[1], 0, 1, 0, 0 }; ops[R++] = X265B200_SAD;
        rounds[R] = MeRound{ 1, wl[3], 1, 1, 0, 0 }; ops[R++] = X265B200_SATD;
    }
    else
    {
        for (int it = 0; it < wl[0]; it++)
        {
            rounds[R] = MeRound{ 2, wl[1], it == 0 && wl[4], it == 0, 0, 0 };
            ops[R++] = wl[4] ? X265B200_SATD : X265B200_SAD;
        }
        for (int it = 0; it < (wl[2] ? wl[2] : 1); it++)
        {   // with no quarter-pel iterations a SAD half-pel search is still re-measured with SATD (:1729-1731)
            if (!wl[2] && wl[4]) break;
            rounds[R] = MeRound{ 1, wl[2] ? wl[3] : 0, it == 0 && !wl[4], it == 0, 0, 0 };
            ops[R++] = X265B200_SATD;
        }
    }
    rounds[R - 1].zeroSlot = 1;                                // the last round is always a SATD round
    int KMAX = 3 + numCand;
    for (int r = 0; r < R; r++)
    {
        rounds[r].K = 1 + rounds[r].dirs + rounds[r].zeroSlot;
        if (rounds[r].K > KMAX) KMAX = rounds[r].K;
    }

    int32_t* scratch = nullptr;
    const size_t per = 2 + 2 + 1 + 2 + 1 + 1 + 4 + 5 * (size_t)KMAX;
    B200_CUDA(ctx, cudaMallocAsync((void**)&scratch, per * n * sizeof(int32_t), st));
    MeState s;
    int32_t* p = scratch;
    s.pmv = p; p += 2 * (size_t)n; s.bestpre = p; p += 2 * (size_t)n; s.bprecost = p; p += n; s.bmv = p; p += 2 * (size_t)n;
    s.bcost = p; p += n; s.flags = p; p += n; s.eff = p; p += 4 * (size_t)n;
    s.candOff = p; p += (size_t)KMAX * n; s.candFrac = p; p += (size_t)KMAX * n; s.candCost = p; p += (size_t)KMAX * n;
    s.candOffC = p; p += (size_t)KMAX * n; s.candFracC = p;

    const int T = 128, G = ceil_div(n, T);
    int rc = X265B200_OK;
    auto bail = [&](int code) { cudaFreeAsync(scratch, st); return code; };

    const int K0 = 3 + numCand;
    auto cand_costs = [&](int op, int K)
    {
        if (pitch) return launch_lowres_cmp(ctx, op, w, h, fenc, strideF, ref, strideR, offF, s.candOff, s.candFrac, K, n, s.candCost, st);
        int r = x265b200_subpel_cmp_batch(ctx, op, w, h, fenc, strideF, ref, strideR, offF, s.candOff, s.candFrac, K, n, s.candCost, stream);
        if (r || !chroma.on) return r;
        const int cw = w >> chroma.hshift, ch = h >> chroma.vshift;      // + SATD of both chroma blocks, whatever `op` is
        r = x265b200_subpel_cmp_chroma_batch(ctx, cw, ch, cargs.fencCb, cargs.strideFC, cargs.refCb, chroma.strideRC, cargs.offFC, s.candOffC,
                                             s.candFracC, K, n, s.candCost, 1, stream);
        if (r) return r;
        return x265b200_subpel_cmp_chroma_batch(ctx, cw, ch, cargs.fencCr, cargs.strideFC, cargs.refCr, chroma.strideRC, cargs.offFC, s.candOffC,
                                                s.candFracC, K, n, s.candCost, 1, stream);
    };
    me_start_gen<<<G, T, 0, st>>>(n, numCand, K0, qmvp, mvc, range, offR, strideR, pitch, s, chroma);
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    if ((rc = cand_costs(X265B200_SAD, K0))) return bail(rc);
    me_start_select<<<G, T, 0, st>>>(n, numCand, K0, qmvp, mvc, range, costTab, s, outQMv, outCost);
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    if (searchMethod == X265B200_ME_FULL)
        rc = x265b200_me_full_batch(ctx, w, h, merange, fenc, strideF, ref, strideR, offF, offR, s.eff, qmvp, costTab, n, s.bmv, s.bcost, stream);
    else if (searchMethod == X265B200_ME_UMH)
        rc = x265b200_me_umh_batch(ctx, w, h, merange, fenc, strideF, ref, strideR, offF, offR, s.eff, qmvp, numCand, mvc, costTab, n, s.bmv, s.bcost, stream);
    else if (searchMethod == X265B200_ME_SEA)
        rc = x265b200_me_sea_batch(ctx, w, h, merange, fenc, strideF, ref, strideR, offF, offR, s.eff, qmvp, costTab, sums, sumsPitch, n, s.bmv, s.bcost, stream);
    else
        rc = x265b200_me_pattern_batch(ctx, searchMethod, w, h, merange, fenc, strideF, ref, strideR, offF, offR, s.eff, qmvp, costTab, n, s.bmv, s.bcost, stream);
    if (rc) return bail(rc);
    for (int r = 0; r <= R; r++)
    {
        const int mode = (r == 0 ? 1 : 2) | (r < R ? 4 : 8);
        me_round_kernel<<<G, T, 0, st>>>(n, mode, rounds[r ? r - 1 : 0], rounds[r < R ? r : R - 1], qmvp, range, offR, strideR, pitch, costTab, s, outQMv, outCost, chroma);
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
        if (r < R && (rc = cand_costs(ops[r], rounds[r].K))) return bail(rc);
    }
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    if (e != cudaSuccess) return fail(ctx, X265B200_ERR_CUDA, "motion_estimate launch", e);
    return X265B200_OK;
}

extern "C" int x265b200_motion_estimate_batch(x265b200_ctx* ctx, int searchMethod, int w, int h, int merange, int subpelRefine,
                                              const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR,
                                              const int32_t* offF, const int32_t* offR, const int32_t* range, const int32_t* qmvp,
                                              int numCand, const int32_t* mvc, const uint16_t* costTab, int n,
                                              int32_t* outQMv, int32_t* outCost, x265b200_stream stream)
{
    return motion_chain(ctx, 0, MeChroma{ 0, 0, 0, 0, nullptr }, MeChromaArgs{}, searchMethod, w, h, merange, subpelRefine, fenc, strideF, ref, strideR,
                        offF, offR, range, qmvp, numCand, mvc, costTab, n, outQMv, outCost, stream);
}

extern "C" int x265b200_motion_estimate_sea_batch(x265b200_ctx* ctx, int w, int h, int merange, int subpelRefine,
                                                  const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR,
                                                  const int32_t* offF, const int32_t* offR, const int32_t* range, const int32_t* qmvp,
                                                  int numCand, const int32_t* mvc, const uint16_t* costTab,
                                                  const uint32_t* sums, size_t planePitch, int n,
                                                  int32_t* outQMv, int32_t* outCost, x265b200_stream stream)
{
    return motion_chain(ctx, 0, MeChroma{ 0, 0, 0, 0, nullptr }, MeChromaArgs{}, X265B200_ME_SEA, w, h, merange, subpelRefine, fenc, strideF, ref, strideR,
                        offF, offR, range, qmvp, numCand, mvc, costTab, n, outQMv, outCost, stream, sums, planePitch);
}

extern "C" int x265b200_lowres_motion_estimate_batch(x265b200_ctx* ctx, int searchMethod, int w, int h, int merange, int subpelRefine,
                                                     const void* fenc, intptr_t strideF, const void* planes, intptr_t strideR, size_t planePitch,
                                                     const int32_t* offF, const int32_t* offR, const int32_t* range, const int32_t* qmvp,
                                                     const uint16_t* costTab, int n, int32_t* outQMv, int32_t* outCost, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (planePitch == 0 || planePitch > 0x1fffffff || ((strideF | strideR) & 3) || w < 4 || w > 64 || h < 4 || h > 64 || (w & 3) || (h & 3))
        return fail(ctx, X265B200_ERR_ARG, "lowres_motion_estimate: bad geometry");
    return motion_chain(ctx, (int)planePitch, MeChroma{ 0, 0, 0, 0, nullptr }, MeChromaArgs{}, searchMethod, w, h, merange, subpelRefine, fenc, strideF,
                        planes, strideR, offF, offR, range, qmvp, 0, nullptr, costTab, n, outQMv, outCost, stream);
}

extern "C" int x265b200_motion_estimate_chroma_batch(x265b200_ctx* ctx, int searchMethod, int w, int h, int merange, int subpelRefine,
                                                     const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR,
                                                     const int32_t* offF, const int32_t* offR,
                                                     const void* fencCb, const void* fencCr, intptr_t strideFC,
                                                     const void* refCb, const void* refCr, intptr_t strideRC,
                                                     const int32_t* offFC, const int32_t* offRC, int hshift, int vshift,
                                                     const int32_t* range, const int32_t* qmvp, int numCand, const int32_t* mvc,
                                                     const uint16_t* costTab, int n, int32_t* outQMv, int32_t* outCost, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (!((hshift == 1 && vshift == 1) || (hshift == 0 && vshift == 0)) || ((strideFC | strideRC) & 3))
        return fail(ctx, X265B200_ERR_ARG, "motion_estimate_chroma: 4:2:0 or 4:4:4 planes with strides that are multiples of 4");
    // bChromaSATD (motion.cpp:240): subme > 2 and a chroma block that has a SATD slot (a multiple of 4x4, pixel.cpp:1217-1243)
    const int on = subpelRefine > 2 && !((w >> hshift) & 3) && !((h >> vshift) & 3);
    if (on && n > 0 && (!fencCb || !fencCr || !refCb || !refCr || !offFC || !offRC))
        return fail(ctx, X265B200_ERR_ARG, "motion_estimate_chroma: chroma planes and offsets are required from subme 3 on");
    MeChroma c{ on, hshift, vshift, strideRC, offRC };
    MeChromaArgs a{ fencCb, fencCr, strideFC, refCb, refCr, offFC };
    return motion_chain(ctx, 0, c, a, searchMethod, w, h, merange, subpelRefine, fenc, strideF, ref, strideR, offF, offR, range, qmvp, numCand, mvc,
                        costTab, n, outQMv, outCost, stream);
}
