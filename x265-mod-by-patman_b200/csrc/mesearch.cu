// mesearch.cu -- exhaustive integer motion search for a batch of prediction units (reference encoder/motion.cpp:1593-1637,
// the X265_FULL_SEARCH case of MotionEstimate::motionEstimate; vector cost from encoder/bitcost.h:53-56).
//
// The reference walks the search window in raster order calling sad_x4 on four neighbouring candidates at a time, so
// every candidate re-reads the whole block from cache.  Here one CTA owns one PU: the part of the reference picture its
// search window touches is staged in shared memory once (in one piece when it fits, else in super-tiles), a warp takes
// work items of 32 candidate columns (one per lane) x J candidate rows, and register-tiles the J rows: a window sample
// fetched from shared memory is compared against J different fenc rows held in a register ring, so the inner loop is
// ~one VABSDIFF per sample-candidate (8-bit pictures: one VABSDIFF4 per four).  The kernel is integer-issue bound, not
// HBM bound: a PU reads its window once (tens of KB) and spends ~10^5..10^7 ALU operations on it.
// Raster-order tie-breaking (COPY2_IF_LT keeps the first strictly-smaller cost) is carried by the reduction key
// (cost << 32 | raster index + 1); index 0 is the caller's initial vector, which therefore wins every tie.
#include "internal.h"
#include "device_util.cuh"

namespace b200 {

constexpr int ME_WARPS = 8;
constexpr int ME_SMEM_CAP = 74 * 1024;          // per CTA: three CTAs (24 warps) per SM

__device__ __forceinline__ unsigned sad4_acc(unsigned a, unsigned b, unsigned c)
{
    unsigned d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// One (window row, fenc row) meeting for a lane: 16-bit samples one at a time, 8-bit samples four at a time.
template<typename PIX> struct MeStep;
template<> struct MeStep<uint16_t>
{
    static constexpr int XSTEP = 1;
    typedef uint16_t cell;
    __device__ static __forceinline__ const cell* win_ptr(const uint16_t* win, int lane, int x, unsigned&) { return win + lane + x; }
    __device__ static __forceinline__ unsigned load_win(const cell* p, unsigned) { return *p; }
    __device__ static __forceinline__ unsigned acc(unsigned r, unsigned f, unsigned a) { return __usad(r, f, a); }
};
template<> struct MeStep<uint8_t>
{
    static constexpr int XSTEP = 4;
    typedef uint32_t cell;
    // a lane's four samples at byte offset lane + x straddle two aligned words; the realigning shift is a per-lane constant
    __device__ static __forceinline__ const cell* win_ptr(const uint8_t* win, int lane, int x, unsigned& sh)
    {
        sh = (lane & 3) * 8;
        return (const cell*)(win + (lane & ~3) + x);
    }
    __device__ static __forceinline__ unsigned load_win(const cell* p, unsigned sh) { return __funnelshift_r(p[0], p[1], sh); }
    __device__ static __forceinline__ unsigned acc(unsigned r, unsigned f, unsigned a) { return sad4_acc(r, f, a); }
};

// SX x SY: candidates staged per super-tile; pitch: window row pitch in samples (a multiple of 4); fenc is kept at pitch w.
template<typename PIX, int J>
__global__ void __launch_bounds__(ME_WARPS * 32)
me_full_kernel(const PIX* __restrict__ fenc, intptr_t strideF, const PIX* __restrict__ ref, intptr_t strideR,
               const int32_t* __restrict__ offF, const int32_t* __restrict__ offR, const int32_t* __restrict__ range,
               const int32_t* __restrict__ mvp, const uint16_t* __restrict__ costTab, int w, int h, int SX, int SY, int pitch,
               int32_t* __restrict__ bmv, int32_t* __restrict__ bcost)
{
    typedef MeStep<PIX> S;
    typedef typename S::cell cell;
    extern __shared__ __align__(16) uint8_t me_smem[];
    __shared__ unsigned long long red[ME_WARPS];
    PIX* fencS = (PIX*)me_smem;
    PIX* win = fencS + w * h;                                   // w * h * sizeof(PIX) is a multiple of 16

    const int pu = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int minx = range[4 * pu], miny = range[4 * pu + 1], maxx = range[4 * pu + 2], maxy = range[4 * pu + 3];
    if (maxx < minx || maxy < miny) return;                     // empty window: (bmv, bcost) stay as they are
    const int mvpx = mvp[2 * pu], mvpy = mvp[2 * pu + 1];
    const PIX* f = fenc + offF[pu];
    const PIX* r0 = ref + offR[pu];
    const unsigned RW = (unsigned)(maxx - minx + 1);

    for (int y = warp; y < h; y += ME_WARPS)
        for (int x = lane; x < w; x += 32) fencS[y * w + x] = f[(intptr_t)y * strideF + x];

    unsigned long long best = ~0ull;
    const int fpitch = w / S::XSTEP, wpitch = pitch / S::XSTEP; // in cells
    for (int sy0 = miny; sy0 <= maxy; sy0 += SY)
        for (int sx0 = minx; sx0 <= maxx; sx0 += SX)
        {
            const int ncx = min(SX, maxx - sx0 + 1), ncy = min(SY, maxy - sy0 + 1);
            const int rows = ncy + h - 1, cols = ncx + w - 1;
            __syncthreads();                                    // the previous super-tile has been consumed (and fencS is written)
            const PIX* src = r0 + (intptr_t)sy0 * strideR + sx0;
            for (int y = warp; y < rows; y += ME_WARPS)
            {
#pragma unroll 4
                for (int x = lane; x < cols; x += 32) win[y * pitch + x] = src[(intptr_t)y * strideR + x];
            }
            __syncthreads();

            const int ngx = (ncx + 31) >> 5, items = ngx * ((ncy + J - 1) / J);
            for (int item = warp; item < items; item += ME_WARPS)
            {
                const int gy = item / ngx, cx0 = (item - gy * ngx) * 32, dy0 = gy * J;
                unsigned acc[J];
#pragma unroll
                for (int j = 0; j < J; j++) acc[j] = 0;
                for (int x = 0; x < w; x += S::XSTEP)
                {
                    unsigned sh = 0;
                    const cell* wp = S::win_ptr(win + dy0 * pitch + cx0, lane, x, sh);
                    const cell* fp = (const cell*)(fencS + x);
                    unsigned ring[J];
                    // window row ry (relative to dy0) meets fenc row ry - j for candidate row j
                    if (h >= J)
                    {
#pragma unroll
                        for (int k = 0; k < J; k++)
                        {   // first rows: candidate rows j > k have not started yet
                            unsigned R = S::load_win(wp + k * wpitch, sh);
                            ring[k] = fp[k * fpitch];
#pragma unroll
                            for (int j = 0; j <= k; j++) acc[j] = S::acc(R, ring[k - j], acc[j]);
                        }
                        for (int ry0 = J; ry0 < h; ry0 += J)
                        {
#pragma unroll
                            for (int k = 0; k < J; k++)
                            {
                                unsigned R = S::load_win(wp + (ry0 + k) * wpitch, sh);
                                ring[k] = fp[(ry0 + k) * fpitch];
#pragma unroll
                                for (int j = 0; j < J; j++) acc[j] = S::acc(R, ring[(k - j + J) % J], acc[j]);
                            }
                        }
#pragma unroll
                        for (int k = 0; k < J - 1; k++)
                        {   // rows below the block of candidate rows 0 .. k
                            unsigned R = S::load_win(wp + (h + k) * wpitch, sh);
#pragma unroll
                            for (int j = k + 1; j < J; j++) acc[j] = S::acc(R, ring[(k - j + J) % J], acc[j]);
                        }
                    }
                }

                const int dx = sx0 + cx0 + lane;
                if (cx0 + lane < ncx)
                {
                    const unsigned cx = costTab[(dx << 2) - mvpx];
#pragma unroll
                    for (int j = 0; j < J; j++)
                        if (dy0 + j < ncy)
                        {
                            const int dy = sy0 + dy0 + j;
                            unsigned cost = acc[j] + (uint16_t)(cx + costTab[(dy << 2) - mvpy]);   // bitcost.h:56 returns uint16_t
                            unsigned idx = (unsigned)(dy - miny) * RW + (unsigned)(dx - minx) + 1u;
                            unsigned long long key = ((unsigned long long)cost << 32) | idx;
                            best = key < best ? key : best;
                        }
                }
            }
        }

#pragma unroll
    for (int o = 16; o; o >>= 1)
    {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if (lane == 0) red[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int k = 1; k < ME_WARPS; k++) best = red[k] < best ? red[k] : best;
        const long long cost = (long long)(best >> 32);
        if (cost < (long long)bcost[pu])                        // strictly cheaper than the caller's starting point
        {
            unsigned idx = (unsigned)best - 1u;
            bcost[pu] = (int32_t)cost;
            bmv[2 * pu] = minx + (int)(idx % RW);
            bmv[2 * pu + 1] = miny + (int)(idx / RW);
        }
    }
}

template<typename PIX, int J>
static cudaError_t launch_me_full(const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR, const int32_t* offF, const int32_t* offR,
                                  const int32_t* range, const int32_t* mvp, const uint16_t* costTab, int w, int h, int span, int n,
                                  int32_t* bmv, int32_t* bcost, cudaStream_t st)
{
    // stage the whole span x span window when it fits; otherwise halve the super-tile rows, then the columns
    const int pb = (int)sizeof(PIX);
    int SX = span, SY = span, pitch, rows;
    size_t bytes;
    for (;;)
    {
        pitch = (SX + w + 3) & ~3;                              // covers SX + w - 1 samples (+ the 8-bit path's second word)
        rows = (SY + J - 1) / J * J + h - 1;
        bytes = (size_t)w * h * pb + ((size_t)rows * pitch + 64) * pb;      // + 64: lanes past the last candidate column read on
        if (bytes <= (size_t)ME_SMEM_CAP) break;
        if (SY > 32) SY = ((SY + 1) / 2 + J - 1) / J * J;
        else if (SX > 32) SX = ((SX + 1) / 2 + 31) & ~31;
        else break;
    }
    cudaError_t e = cudaFuncSetAttribute((const void*)me_full_kernel<PIX, J>, cudaFuncAttributeMaxDynamicSharedMemorySize, ME_SMEM_CAP);
    if (e != cudaSuccess) return e;
    me_full_kernel<PIX, J><<<n, ME_WARPS * 32, bytes, st>>>((const PIX*)fenc, strideF, (const PIX*)ref, strideR, offF, offR, range, mvp, costTab,
                                                           w, h, SX, SY, pitch, bmv, bcost);
    return cudaSuccess;
}


// ---------------------------------------------------------------------------------------------------------------
// Pattern searches: diamond (motion.cpp:1016-1039), hexagon + square refinement (motion.cpp:1041-1138) and star
// (motion.cpp:386-630, 1327-1435).
// Each step depends on the previous best, so a PU is a sequential walk; one warp owns one PU (fenc block in shared
// memory, the lanes split the block's samples, up to four candidates of a step are measured in one pass) and the
// batch supplies the parallelism.  Only a candidate's row is range-checked, as in the reference; a step may leave the
// window horizontally, which ends the walk.  Equal costs keep the earlier candidate (the reference's tag-in-low-bits
// comparison reduces to a strict `<` in evaluation order).
// ---------------------------------------------------------------------------------------------------------------
constexpr int MP_WARPS = 4;
__constant__ int c_hex2[8][2] = { {-1, -2}, {-2, 0}, {-1, 2}, {1, 2}, {2, 0}, {1, -2}, {-1, -2}, {-2, 0} };      // motion.cpp:65
__constant__ int c_sq1[9][2] = { {0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {-1, 1}, {1, -1}, {1, 1} }; // motion.cpp:67

template<typename PIX>
struct PatternPU
{
    const PIX* fs;              // fenc block, pitch w (shared)
    const PIX* r0;              // co-located block of the reference picture
    intptr_t strideR;
    const uint16_t* cx;         // cost table shifted by the predictor (BitCost::setMVP)
    const uint16_t* cy;
    int w, h, lane, px0, py0, dq, dr;     // a lane's sample walks 32 positions per step: dq rows and dr columns
    int minx, maxx, miny, maxy;

    __device__ bool row_ok(int y) const { return y >= miny && y <= maxy; }
    __device__ bool inside(int x, int y) const { return x >= minx && x <= maxx && y >= miny && y <= maxy; }

    // SAD + mvcost of up to four full-pel candidates in one pass over the block.  ok[k] says whether the reference
    // measures candidate k at all (it is not read otherwise); q3 marks a candidate charged mvcost(mv << 3), the
    // reference's raster quirk (motion.cpp:1392), instead of mv << 2; q3 == -2: plain SADs, no vector cost (SEA's triples).
    // wrap16 = false: the vector cost is the plain sum of the two table entries (the reference's COST_MV_X4_DIR-free ring of
    // motion.cpp:1282-1310 adds them as ints), not BitCost::mvcost's uint16_t sum.
    __device__ void eval(int nc, const int (&cand)[4][2], const bool (&ok)[4], int (&cost)[4], int q3 = -1, bool wrap16 = true) const
    {
        const PIX* rp[4];
        unsigned acc[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const bool use = k < nc && ok[k];
            rp[k] = r0 + (use ? cand[k][0] + (intptr_t)cand[k][1] * strideR : 0);
            acc[k] = 0;
        }
        int px = px0, py = py0;
        for (int idx = lane; idx < w * h; idx += 32)
        {
            const unsigned f = fs[idx];
            const intptr_t o = (intptr_t)py * strideR + px;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (k < nc) acc[k] = __usad(f, (unsigned)rp[k][o], acc[k]);
            px += dr; py += dq;
            if (px >= w) { px -= w; py++; }
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const int sh = k == q3 ? 3 : 2;
            const bool use = k < nc && ok[k];
            const int mc = (use && q3 != -2) ? (int)cx[cand[k][0] << sh] + (int)cy[cand[k][1] << sh] : 0;
            cost[k] = (int)__reduce_add_sync(0xffffffffu, acc[k]) + (wrap16 ? (int)(uint16_t)mc : mc);
        }
    }
};

// running best of the star search: vector, cost, and the point number / distance of the ring point that set it
struct StarBest { int x, y, cost, point, dist; };

// ring point idx of the pattern at distance dist (motion.cpp:405-628): offset, point number, and the distance it reports
__device__ __forceinline__ void star_point(int dist, int idx, int& dx, int& dy, int& point, int& d)
{
    if (dist == 1)
    {   // 2 4 5 7
        const int px[4] = { 0, -1, 1, 0 }, py[4] = { -1, 0, 0, 1 }, pn[4] = { 2, 4, 5, 7 };
        dx = px[idx]; dy = py[idx]; point = pn[idx]; d = 1;
    }
    else if (dist <= 8)
    {   // 2 1 3 4 5 6 8 7: axis points at dist, diagonal points at dist / 2
        const int sx[8] = { 0, -1, 1, -2, 2, -1, 1, 0 }, sy[8] = { -2, -1, -1, 0, 0, 1, 1, 2 }, pn[8] = { 2, 1, 3, 4, 5, 6, 8, 7 };
        const int h = dist >> 1;
        dx = sx[idx] * h; dy = sy[idx] * h; point = pn[idx]; d = (idx == 0 || idx == 3 || idx == 4 || idx == 7) ? dist : h;
    }
    else
    {   // top, left, right, bottom, then three points on each edge of the diamond, quarter by quarter
        const int q = dist >> 2;
        point = 0; d = dist;
        if (idx < 4) { dx = idx == 1 ? -dist : idx == 2 ? dist : 0; dy = idx == 0 ? -dist : idx == 3 ? dist : 0; }
        else
        {
            const int i = ((idx - 4) >> 2) + 1, c = (idx - 4) & 3;
            dx = (c & 1) ? q * i : -q * i;
            dy = (c & 2) ? dist - q * i : -dist + q * i;
        }
    }
}

template<typename PIX>
__device__ void star_pattern(const PatternPU<PIX>& P, StarBest& b, int earlyExit, int merange)
{
    const int ox = b.x, oy = b.y;
    int rounds = 0, saved = b.cost;
    int cand[4][2], cost[4], pt[4], dd[4]; bool ok[4];
    for (int dist = 1; dist <= 8 || dist <= (int)(int16_t)merange; dist <<= 1)
    {
        if (dist > 1) saved = b.cost;
        const int npts = dist == 1 ? 4 : dist <= 8 ? 8 : 16;
        for (int i0 = 0; i0 < npts; i0 += 4)
        {
            for (int k = 0; k < 4; k++)
            {
                int dx, dy;
                star_point(dist, i0 + k, dx, dy, pt[k], dd[k]);
                cand[k][0] = ox + dx; cand[k][1] = oy + dy;
                ok[k] = P.inside(cand[k][0], cand[k][1]);
            }
            if (!(ok[0] | ok[1] | ok[2] | ok[3])) continue;
            P.eval(4, cand, ok, cost);
            for (int k = 0; k < 4; k++)
                if (ok[k] && cost[k] < b.cost) { b.cost = cost[k]; b.x = cand[k][0]; b.y = cand[k][1]; b.point = pt[k]; b.dist = dd[k]; }
        }
        if (b.cost < saved) rounds = 0;
        else if (++rounds >= earlyExit) return;
    }
}

// the two outer neighbours of a distance-1 winner (motion.cpp:76-86 `offsets`), both taken around the winner as it was
__constant__ int c_two_point[16][2] = { {-1, 0}, {0, -1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {-1, -1},
                                        {1, -1}, {1, 1}, {-1, 0}, {0, 1}, {-1, 1}, {1, 1}, {1, 0}, {0, 1} };
template<typename PIX>
__device__ void star_two_point(const PatternPU<PIX>& P, StarBest& b)
{
    int cand[4][2], cost[4]; bool ok[4] = { false, false, false, false };
    const int p = (b.point - 1) * 2;
    for (int k = 0; k < 2; k++)
    {
        cand[k][0] = b.x + c_two_point[p + k][0]; cand[k][1] = b.y + c_two_point[p + k][1];
        ok[k] = P.inside(cand[k][0], cand[k][1]);
    }
    cand[2][0] = cand[3][0] = b.x; cand[2][1] = cand[3][1] = b.y;
    if (!(ok[0] | ok[1])) return;
    P.eval(2, cand, ok, cost);
    for (int k = 0; k < 2; k++)
        if (ok[k] && cost[k] < b.cost) { b.cost = cost[k]; b.x = cand[k][0]; b.y = cand[k][1]; }
}

// X265_STAR_SEARCH, motion.cpp:1327-1435.  phase 0: the whole search in this warp.  The stride-5 raster pass some PUs take (first pattern
// ended more than 5 away) is ~530 candidates: in one warp it lasts longer than every other walk of the launch together, so the library
// splits the search around it -- phase 1 stops at the raster decision (returns true, x / y / best = the state before it), the raster runs
// as its own CTA-per-PU kernel (me_star_raster_kernel), phase 2 resumes with the re-centred passes.
template<typename PIX>
__device__ bool star_search(const PatternPU<PIX>& P, int merange, int& x, int& y, int& best, int phase)
{
    StarBest b = { x, y, best, 0, 0 };
    bool done = false;
    if (phase != 2)
    {
        star_pattern(P, b, 3, merange);
        if (b.dist == 1)
        {
            if (!b.point) done = true;
            else
            {
                const int saved = b.cost;
                star_two_point(P, b);
                done = b.cost == saved;
            }
        }
    }
    else b.dist = 6;                                        // only PUs that were due for the raster come back; the loop below resets it
    if (!done)
    {
        if (phase != 2 && b.dist > 5)
        {
            if (phase == 1) { x = b.x; y = b.y; best = b.cost; return true; }
            // raster over the window in steps of 5, four columns per pass where the reference uses sad_x4
            int cand[4][2], cost[4]; bool ok[4];
            for (int ty = P.miny; ty <= P.maxy; ty += 5)
                for (int tx = P.minx; tx <= P.maxx; tx += 5)
                {
                    const bool quad = tx + 15 <= P.maxx;
                    for (int k = 0; k < 4; k++) { cand[k][0] = tx + (quad ? 5 * k : 0); cand[k][1] = ty; ok[k] = quad || k == 0; }
                    P.eval(quad ? 4 : 1, cand, ok, cost, quad ? 3 : -1);
                    for (int k = 0; k < (quad ? 4 : 1); k++)
                        if (cost[k] < b.cost) { b.cost = cost[k]; b.x = cand[k][0]; b.y = cand[k][1]; }
                    if (quad) tx += 15;
                }
        }
        while (b.dist > 0)
        {   // re-centred passes until one brings nothing
            b.dist = 0; b.point = 0;
            star_pattern(P, b, 32, merange);
            if (b.dist == 1)
            {
                if (b.point) star_two_point(P, b);
                break;
            }
        }
    }
    x = b.x; y = b.y; best = b.cost;
    return false;
}

// The raster pass of the star search (motion.cpp:1365-1399) for the PUs phase 1 flagged: one CTA per PU, a lane owns one candidate of the
// stride-5 grid and walks the whole block (fenc from shared memory, the lanes of a warp read reference samples 10 bytes apart: a few cache
// lines per request).  The reference measures the grid four columns at a time with sad_x4 and charges the fourth candidate of each group
// mvcost(mv << 3) (:1392); groups run while tx + 15 <= maxx, the rest of a row is single candidates.  Visiting order = raster order of the
// grid with strict-less updates: carried by the key (cost << 32 | grid index + 1), index 0 being the state before the pass.
constexpr int RS_WARPS = 8;
template<typename PIX>
__global__ void __launch_bounds__(RS_WARPS * 32)
me_star_raster_kernel(const PIX* __restrict__ fenc, intptr_t strideF, const PIX* __restrict__ ref, intptr_t strideR,
                      const int32_t* __restrict__ offF, const int32_t* __restrict__ offR, const int32_t* __restrict__ range,
                      const int32_t* __restrict__ mvp, const uint16_t* __restrict__ costTab, int w, int h,
                      const int32_t* __restrict__ flags, int32_t* __restrict__ bmv, int32_t* __restrict__ bcost)
{
    extern __shared__ __align__(16) uint8_t rs_smem[];
    __shared__ unsigned long long red[RS_WARPS];
    const int pu = blockIdx.x;
    if (!flags[pu]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int minx = range[4 * pu], miny = range[4 * pu + 1], maxx = range[4 * pu + 2], maxy = range[4 * pu + 3];
    PIX* fs = (PIX*)rs_smem;
    {
        const PIX* f = fenc + offF[pu];
        for (int idx = threadIdx.x; idx < w * h; idx += RS_WARPS * 32) fs[idx] = f[(intptr_t)(idx / w) * strideF + idx % w];
    }
    __syncthreads();
    const int ncx = (maxx - minx) / 5 + 1, ncy = (maxy - miny) / 5 + 1;
    const int grouped = maxx - minx >= 15 ? 4 * ((maxx - minx - 15) / 20 + 1) : 0;      // columns measured as groups of four
    const uint16_t* cx = costTab - mvp[2 * pu];
    const uint16_t* cy = costTab - mvp[2 * pu + 1];
    const PIX* r0 = ref + offR[pu];
    const int chunks = (ncx + 31) >> 5;
    unsigned long long best = ~0ull;
    for (int item = warp; item < ncy * chunks; item += RS_WARPS)
    {
        const int j = item / chunks, i = (item - j * chunks) * 32 + lane;
        const bool live = i < ncx;
        const int tx = minx + 5 * (live ? i : ncx - 1), ty = miny + 5 * j;
        const PIX* rp = r0 + (intptr_t)ty * strideR + tx;
        unsigned acc = 0;
        for (int yy = 0; yy < h; yy++)
        {
            const PIX* rr = rp + (intptr_t)yy * strideR;
            const PIX* ff = fs + yy * w;
#pragma unroll 4
            for (int xx = 0; xx < w; xx++) acc = __usad((unsigned)ff[xx], (unsigned)rr[xx], acc);
        }
        if (live)
        {
            const int sh = (i < grouped && (i & 3) == 3) ? 3 : 2;
            const unsigned cost = acc + (unsigned)(uint16_t)((int)cx[tx << sh] + (int)cy[ty << sh]);
            const unsigned long long key = ((unsigned long long)cost << 32) | (unsigned)(j * ncx + i + 1);
            best = key < best ? key : best;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1)
    {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if (lane == 0) red[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int k = 1; k < RS_WARPS; k++) best = red[k] < best ? red[k] : best;
        const long long cost = (long long)(best >> 32);
        if (best != ~0ull && cost < (long long)bcost[pu])
        {
            const int idx = (int)((unsigned)best - 1u);
            bcost[pu] = (int32_t)cost;
            bmv[2 * pu] = minx + 5 * (idx % ncx);
            bmv[2 * pu + 1] = miny + 5 * (idx / ncx);
        }
    }
}

// X265_HEX_SEARCH, motion.cpp:1041-1138: the six corners around the start (hex2[1..6] order), the walk, the square refinement
template<typename PIX>
__device__ void hex_search(const PatternPU<PIX>& P, int merange, int& x, int& y, int& best)
{
    const int minx = P.minx, maxx = P.maxx, miny = P.miny, maxy = P.maxy;
    int cand[4][2] = { {0, 0}, {0, 0}, {0, 0}, {0, 0} }, cost[4]; bool ok[4] = { false, false, false, false };
    auto in_range = [&](int cx, int cy) { return cx >= minx && cx <= maxx && cy >= miny && cy <= maxy; };
        int win = 0;
        for (int half = 0; half < 2; half++)
        {
            for (int k = 0; k < 3; k++) { cand[k][0] = x + c_hex2[half * 3 + k + 1][0]; cand[k][1] = y + c_hex2[half * 3 + k + 1][1]; ok[k] = P.row_ok(cand[k][1]); }
            P.eval(3, cand, ok, cost);
            for (int k = 0; k < 3; k++)
                if (ok[k] && cost[k] < best) { best = cost[k]; win = half * 3 + k + 2; }
        }
        if (win)
        {
            int dir = win - 2;
            x += c_hex2[dir + 1][0]; y += c_hex2[dir + 1][1];
            for (int i = (merange >> 1) - 1; i > 0 && in_range(x, y); i--)
            {   // the three corners the previous hexagon did not cover
                for (int k = 0; k < 3; k++) { cand[k][0] = x + c_hex2[dir + k][0]; cand[k][1] = y + c_hex2[dir + k][1]; ok[k] = P.row_ok(cand[k][1]); }
                P.eval(3, cand, ok, cost);
                win = 0;
                for (int k = 0; k < 3; k++)
                    if (ok[k] && cost[k] < best) { best = cost[k]; win = k + 1; }
                if (!win) break;
                dir = (dir + win - 2 + 6) % 6;                  // mod6m1[dir + 1]
                x += c_hex2[dir + 1][0]; y += c_hex2[dir + 1][1];
            }
        }
        // square refinement around the final centre: the cross, then the corners
        win = 0;
        for (int half = 0; half < 2; half++)
        {
            for (int k = 0; k < 4; k++) { cand[k][0] = x + c_sq1[half * 4 + k + 1][0]; cand[k][1] = y + c_sq1[half * 4 + k + 1][1]; ok[k] = P.row_ok(cand[k][1]); }
            P.eval(4, cand, ok, cost);
            for (int k = 0; k < 4; k++)
                if (ok[k] && cost[k] < best) { best = cost[k]; win = half * 4 + k + 1; }
        }
        x += c_sq1[win][0]; y += c_sq1[win][1];
}

// X265_UMH_SEARCH, motion.cpp:1142-1324 (uneven multi-hexagon, from x264): small diamonds around the predictor, the zero vector and
// the running best; an early-termination ladder on SAD thresholds scaled by the PU height; a cross whose reach adapts to how much
// the neighbour vectors disagree; the 5x5 corners; rings of the 16-point hexagon at radius 1 .. merange / 4.  Returns true when
// the search goes on into the hexagon stage with the (possibly widened) merange.  Candidate order, the row-only range check of
// the x4 groups and the strict-less updates follow the reference; four candidates are measured per pass over the block.
struct UmhBest { int x, y, cost; };
template<typename PIX>
__device__ __forceinline__ void umh_try(const PatternPU<PIX>& P, UmhBest& u, int x, int y)
{
    int cand[4][2] = { {x, y}, {x, y}, {x, y}, {x, y} }, cost[4]; bool ok[4] = { true, false, false, false };
    P.eval(1, cand, ok, cost);
    if (cost[0] < u.cost) { u.cost = cost[0]; u.x = x; u.y = y; }
}
template<typename PIX>
__device__ __forceinline__ void umh_x4(const PatternPU<PIX>& P, UmhBest& u, int ox, int oy, int d0x, int d0y, int d1x, int d1y, int d2x, int d2y, int d3x, int d3y)
{
    int cand[4][2] = { {ox + d0x, oy + d0y}, {ox + d1x, oy + d1y}, {ox + d2x, oy + d2y}, {ox + d3x, oy + d3y} }, cost[4]; bool ok[4];
    for (int k = 0; k < 4; k++) ok[k] = P.row_ok(cand[k][1]);
    if (!(ok[0] | ok[1] | ok[2] | ok[3])) return;
    P.eval(4, cand, ok, cost);
    for (int k = 0; k < 4; k++)
        if (ok[k] && cost[k] < u.cost) { u.cost = cost[k]; u.x = cand[k][0]; u.y = cand[k][1]; }
}
template<typename PIX>
__device__ void umh_cross(const PatternPU<PIX>& P, UmhBest& u, int ox, int oy, int start, int x_max, int y_max)
{   // CROSS, motion.cpp:359-385
    int i = start;
    if (x_max <= min(P.maxx - ox, ox - P.minx))
        for (; i < x_max - 2; i += 4) umh_x4(P, u, ox, oy, i, 0, -i, 0, i + 2, 0, -i - 2, 0);
    for (; i < x_max; i += 2)
    {
        if (ox + i <= P.maxx) umh_try(P, u, ox + i, oy);
        if (ox - i >= P.minx) umh_try(P, u, ox - i, oy);
    }
    i = start;
    if (y_max <= min(P.maxy - oy, oy - P.miny))
        for (; i < y_max - 2; i += 4) umh_x4(P, u, ox, oy, 0, i, 0, -i, 0, i + 2, 0, -i - 2);
    for (; i < y_max; i += 2)
    {
        if (oy + i <= P.maxy) umh_try(P, u, ox, oy + i);
        if (oy - i >= P.miny) umh_try(P, u, ox, oy - i);
    }
}
__constant__ int c_hex4[16][2] = { {0, -4}, {0, 4}, {-2, -3}, {2, -3}, {-4, -2}, {4, -2}, {-4, -1}, {4, -1},
                                   {-4, 0}, {4, 0}, {-4, 1}, {4, 1}, {-4, 2}, {4, 2}, {-2, 3}, {2, 3} };              // motion.cpp:68-74
__constant__ unsigned char c_range_mul[4][4] = { { 3, 3, 4, 4 }, { 3, 4, 4, 4 }, { 4, 4, 4, 5 }, { 4, 4, 5, 6 } };     // motion.cpp:1232
template<typename PIX>
__device__ bool umh_search(const PatternPU<PIX>& P, int& merange, int pmvx, int pmvy, const int32_t* qmvp, int numCand, const int32_t* mvc,
                           int& x, int& y, int& best)
{
    UmhBest u = { x, y, best };
    const int scale = (P.h * P.h) >> 4;                         // sizeScale, motion.cpp:124-152
#define UMH_THRESH(v) (u.cost < (((v) >> 4) * scale))
    int cross_start = 1;
    const int ucost1 = u.cost;
    umh_x4(P, u, pmvx, pmvy, 0, -1, 0, 1, -1, 0, 1, 0);
    if (pmvx | pmvy) umh_x4(P, u, 0, 0, 0, -1, 0, 1, -1, 0, 1, 0);
    const int ucost2 = u.cost;
    if ((u.x | u.y) && (u.x != pmvx || u.y != pmvy)) umh_x4(P, u, u.x, u.y, 0, -1, 0, 1, -1, 0, 1, 0);
    if (u.cost == ucost2) cross_start = 3;
    int ox = u.x, oy = u.y;
    if (u.cost == ucost2 && UMH_THRESH(2000))
    {
        umh_x4(P, u, ox, oy, 0, -2, -1, -1, 1, -1, -2, 0);
        umh_x4(P, u, ox, oy, 2, 0, -1, 1, 1, 1, 0, 2);
        if (u.cost == ucost1 && UMH_THRESH(500)) { x = u.x; y = u.y; best = u.cost; return false; }
        if (u.cost == ucost2)
        {
            const int reach = (int)(int16_t)(merange >> 1) | 1;
            umh_cross(P, u, ox, oy, 3, reach, reach);
            umh_x4(P, u, ox, oy, -1, -2, 1, -2, -2, -1, 2, -1);
            umh_x4(P, u, ox, oy, -2, 1, 2, 1, -1, 2, 1, 2);
            if (u.cost == ucost2) { x = u.x; y = u.y; best = u.cost; return false; }
            cross_start = reach + 2;
        }
    }
    if (numCand)
    {   // search range scaled by the disagreement of the predictors and by how good the match already is
        const bool is64 = P.w == 64 && P.h == 64;
        int mvd, denom = 1;
        if (numCand == 1)
            mvd = is64 ? 25 : abs(qmvp[0] - mvc[0]) + abs(qmvp[1] - mvc[1]);
        else
        {
            denom = numCand - 1;
            mvd = 0;
            if (!is64) { mvd = abs(qmvp[0] - mvc[0]) + abs(qmvp[1] - mvc[1]); denom++; }
            for (int i = 0; i < numCand - 1; i++)
                mvd += abs(mvc[2 * i] - mvc[2 * i + 2]) + abs(mvc[2 * i + 1] - mvc[2 * i + 3]);
        }
        const int sad_ctx = UMH_THRESH(1000) ? 0 : UMH_THRESH(2000) ? 1 : UMH_THRESH(4000) ? 2 : 3;
        const int mvd_ctx = mvd < 10 * denom ? 0 : mvd < 20 * denom ? 1 : mvd < 40 * denom ? 2 : 3;
        merange = (merange * c_range_mul[mvd_ctx][sad_ctx]) >> 2;
    }
    // the cross and the corners stay centred where the diamonds ended (the reference's FIXME)
    umh_cross(P, u, ox, oy, cross_start, merange, merange >> 1);
    umh_x4(P, u, ox, oy, -2, -2, -2, 2, 2, -2, 2, 2);
    // hexagon grid around the new best
    ox = u.x; oy = u.y;
    unsigned short i = 1;
    do
    {
        const int room = min(min(P.maxx - ox, ox - P.minx), min(P.maxy - oy, oy - P.miny));
        const bool whole = 4 * i <= room;                       // whole ring inside the window: plain int vector costs, no range checks
        for (int j0 = 0; j0 < 16; j0 += 4)
        {
            int cand[4][2], cost[4]; bool ok[4];
            for (int k = 0; k < 4; k++)
            {
                cand[k][0] = ox + c_hex4[j0 + k][0] * i; cand[k][1] = oy + c_hex4[j0 + k][1] * i;
                ok[k] = whole || P.inside(cand[k][0], cand[k][1]);
            }
            if (!(ok[0] | ok[1] | ok[2] | ok[3])) continue;
            P.eval(4, cand, ok, cost, -1, !whole);
            for (int k = 0; k < 4; k++)
                if (ok[k] && cost[k] < u.cost) { u.cost = cost[k]; u.x = cand[k][0]; u.y = cand[k][1]; }
        }
    }
    while (++i <= merange >> 2);
#undef UMH_THRESH
    x = u.x; y = u.y; best = u.cost;
    return P.inside(u.x, u.y);
}

// WPC warps (= PUs) per CTA.  A CTA lives as long as its longest walk: one warp per CTA lets every finished walk give its slot back at once,
// a few per cent for the uneven multi-hexagon search on small PUs (8x8: 5.1 -> 4.9 ms per 2160p frame with the reference's cost table); the
// other methods are a few per cent faster with four (profiles/r3_motion_search_10bit.json).
template<typename PIX, int WPC>
__global__ void __launch_bounds__(WPC * 32)
me_pattern_kernel(int method, int merange, const PIX* __restrict__ fenc, intptr_t strideF, const PIX* __restrict__ ref, intptr_t strideR,
                  const int32_t* __restrict__ offF, const int32_t* __restrict__ offR, const int32_t* __restrict__ range,
                  const int32_t* __restrict__ mvp, const uint16_t* __restrict__ costTab, int n, int w, int h,
                  int32_t* __restrict__ bmv, int32_t* __restrict__ bcost, int numCand, const int32_t* __restrict__ mvc,
                  int phase, int32_t* __restrict__ flags)
{
    extern __shared__ __align__(16) uint8_t mp_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pu = blockIdx.x * WPC + warp;
    if (pu >= n) return;                                        // whole warps leave; only __syncwarp below
    if (phase == 2 && !flags[pu]) return;                       // star search, second half: only the PUs that took the raster pass
    const int minx = range[4 * pu], miny = range[4 * pu + 1], maxx = range[4 * pu + 2], maxy = range[4 * pu + 3];
    if (maxx < minx || maxy < miny) return;
    PIX* fs = (PIX*)mp_smem + (size_t)warp * w * h;
    {
        const PIX* f = fenc + offF[pu];
        int px = lane % w, py = lane / w;
        for (int idx = lane; idx < w * h; idx += 32)
        {
            fs[idx] = f[(intptr_t)py * strideF + px];
            px += 32 % w; py += 32 / w;
            if (px >= w) { px -= w; py++; }
        }
    }
    __syncwarp();
    PatternPU<PIX> P;
    P.fs = fs; P.r0 = ref + offR[pu]; P.strideR = strideR;
    P.cx = costTab - mvp[2 * pu]; P.cy = costTab - mvp[2 * pu + 1];
    P.w = w; P.h = h; P.lane = lane; P.px0 = lane % w; P.py0 = lane / w; P.dq = 32 / w; P.dr = 32 % w; P.minx = minx; P.maxx = maxx; P.miny = miny; P.maxy = maxy;

    int x = bmv[2 * pu], y = bmv[2 * pu + 1], best = bcost[pu];
    int cand[4][2] = { {0, 0}, {0, 0}, {0, 0}, {0, 0} }, cost[4]; bool ok[4] = { false, false, false, false };
    auto in_range = [&](int cx, int cy) { return cx >= minx && cx <= maxx && cy >= miny && cy <= maxy; };

    if (method == X265B200_ME_STAR)
    {
        const bool raster = star_search(P, merange, x, y, best, phase);
        if (phase == 1 && lane == 0) flags[pu] = raster;
    }
    else if (method == 0)
    {   // diamond, radius 1: up, down, left, right
        int i = merange;
        do
        {
            for (int k = 0; k < 4; k++) { cand[k][0] = x + c_sq1[k + 1][0]; cand[k][1] = y + c_sq1[k + 1][1]; ok[k] = P.row_ok(cand[k][1]); }
            P.eval(4, cand, ok, cost);
            int win = -1;
            for (int k = 0; k < 4; k++)
                if (ok[k] && cost[k] < best) { best = cost[k]; win = k; }
            if (win < 0) break;
            x = cand[win][0]; y = cand[win][1];
        }
        while (--i && in_range(x, y));
    }
    else
    {
        if (method == X265B200_ME_UMH)
        {
            int mr = merange;
            const int qminx = minx * 4, qmaxx = maxx * 4, qminy = miny * 4, qmaxy = maxy * 4;
            const int pmvx = min(max(mvp[2 * pu], qminx), qmaxx), pmvy = min(max(mvp[2 * pu + 1], qminy), qmaxy);
            if (umh_search(P, mr, (pmvx + 2) >> 2, (pmvy + 2) >> 2, mvp + 2 * pu, numCand, mvc ? mvc + 2 * (size_t)numCand * pu : nullptr, x, y, best))
                hex_search(P, mr, x, y, best);
        }
        else hex_search(P, merange, x, y, best);
    }
    if (lane == 0) { bmv[2 * pu] = x; bmv[2 * pu + 1] = y; bcost[pu] = best; }
}


// ---------------------------------------------------------------------------------------------------------------
// X265_SEA, motion.cpp:1438-1591 (successive elimination): every row of the window around the running best is pre-filtered
// with `ads` over the twelve integral planes of the reference picture (csrc/integral.cu) and only the survivors get a SAD.
// One warp owns a PU and walks its rows in order, because each row's threshold is the best cost so far; inside a row the
// lanes filter 32 columns at a time (ballot + popc keeps the ascending order the reference's list has), then the survivors
// are measured four at a time.  The reference's cost bookkeeping is reproduced as it is: the row cost comes from the table
// shifted by the predictor a second time, indexed by the FULL-pel y and scaled by 4; survivors taken three at a time are
// charged only the doubly shifted x-cost while the row cost is off the running best; the window width is rounded up to a
// multiple of 4, so up to three columns right of it are examined too.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ads_terms(int w, int h)
{   // pixel.cpp:1122-1146: which ads_x{4,2,1} a PU shape is bound to
    if ((w == 8 && h == 4) || (w == 4 && h == 8) || (w == 16 && h == 8) || (w == 8 && h == 16) || (w == 32 && h == 16) || (w == 16 && h == 32) ||
        (w == 64 && h == 32) || (w == 32 && h == 64)) return 2;
    if ((w == 4 && h == 4) || (w == 8 && h == 8) || (w == 16 && h == 12) || (w == 12 && h == 16) || (w == 16 && h == 4) || (w == 4 && h == 16)) return 1;
    return 4;
}

constexpr int SEA_MAX_WIDTH = 1024;             // survivors of one row (int16 each) per warp

template<typename PIX>
__global__ void __launch_bounds__(MP_WARPS * 32)
me_sea_kernel(int merange, const PIX* __restrict__ fenc, intptr_t strideF, const PIX* __restrict__ ref, intptr_t strideR,
              const int32_t* __restrict__ offF, const int32_t* __restrict__ offR, const int32_t* __restrict__ range,
              const int32_t* __restrict__ mvp, const uint16_t* __restrict__ costTab, const uint32_t* __restrict__ sums, size_t planePitch,
              int n, int w, int h, int listCap, int32_t* __restrict__ bmv, int32_t* __restrict__ bcost)
{
    extern __shared__ __align__(16) uint8_t mp_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pu = blockIdx.x * MP_WARPS + warp;
    if (pu >= n) return;
    const int rminx = range[4 * pu], rminy = range[4 * pu + 1], rmaxx = range[4 * pu + 2], rmaxy = range[4 * pu + 3];
    if (rmaxx < rminx || rmaxy < rminy) return;
    PIX* fs = (PIX*)mp_smem + (size_t)warp * w * h;
    int16_t* list = (int16_t*)((PIX*)mp_smem + (size_t)MP_WARPS * w * h) + (size_t)warp * listCap;
    const PIX* f = fenc + offF[pu];
    {
        int px = lane % w, py = lane / w;
        for (int idx = lane; idx < w * h; idx += 32)
        {
            fs[idx] = f[(intptr_t)py * strideF + px];
            px += 32 % w; py += 32 / w;
            if (px >= w) { px -= w; py++; }
        }
    }
    __syncwarp();
    PatternPU<PIX> P;
    P.fs = fs; P.r0 = ref + offR[pu]; P.strideR = strideR;
    const int qmx = mvp[2 * pu], qmy = mvp[2 * pu + 1];
    P.cx = costTab - qmx; P.cy = costTab - qmy;
    P.w = w; P.h = h; P.lane = lane; P.px0 = lane % w; P.py0 = lane / w; P.dq = 32 / w; P.dr = 32 % w;
    P.minx = rminx; P.maxx = rmaxx; P.miny = rminy; P.maxy = rmaxy;
    const uint16_t* pcx = P.cx - qmx;
    const uint16_t* pcy = P.cy - qmy;

    int x = bmv[2 * pu], y = bmv[2 * pu + 1], best = bcost[pu];
    const int minX = max(x - merange, rminx), minY = max(y - merange, rminy), maxX = min(x + merange, rmaxx), maxY = min(y + merange, rmaxy);
    const int width = min((maxX - minX + 3) & ~3, listCap);

    // which sub-block DCs the PU's ads form compares, and on which integral plane (motion.cpp:1445-1548)
    int deltaX = w <= 8 ? w : w >> 1, deltaY = h <= 8 ? h : h >> 1;
    const bool vertical = (w == 32 && h == 64) || (w == 16 && h == 32) || (w == 8 && h == 16) || (w == 4 && h == 8);
    const bool horizontal = (w == 64 && h == 32) || (w == 32 && h == 16) || (w == 16 && h == 8) || (w == 8 && h == 4);
    const bool smallRect = (w == 4 && h == 4) || (w == 16 && h == 12) || (w == 12 && h == 16) || (w == 16 && h == 4) || (w == 4 && h == 16);
    const bool asym = (w == 12 && h == 16) || (w == 4 && h == 16) || (w == 24 && h == 32) || (w == 8 && h == 32) || (w == 48 && h == 64) ||
                      (w == 16 && h == 64) || (w == 16 && h == 12) || (w == 16 && h == 4) || (w == 32 && h == 24) || (w == 32 && h == 8) ||
                      (w == 64 && h == 48) || (w == 64 && h == 16);
    int tw, th;
    if (vertical) { tw = w; th = h >> 1; }
    else if (horizontal) { tw = w >> 1; th = h; }
    else if (asym) { tw = smallRect ? w : w >> 1; th = smallRect ? h : h >> 1; }
    else { tw = w <= 8 ? w : w >> 1; th = w <= 8 ? h : h >> 1; }
    long long encDC[4];
    {
        unsigned s4[4] = { 0, 0, 0, 0 };
        for (int idx = lane; idx < w * h; idx += 32)
        {
            const int yy = idx / w, xx = idx - yy * w;
            const unsigned v = fs[idx];
            // sub-block k starts at (k & 1 ? deltaX : 0, k & 2 ? deltaY : 0) and is tw x th; samples outside the PU count as 0
            if (xx < tw && yy < th) s4[0] += v;
            if (xx >= deltaX && xx < deltaX + tw && yy < th) s4[1] += v;
            if (xx < tw && yy >= deltaY && yy < deltaY + th) s4[2] += v;
            if (xx >= deltaX && xx < deltaX + tw && yy >= deltaY && yy < deltaY + th) s4[3] += v;
        }
        for (int k = 0; k < 4; k++) encDC[k] = (long long)__reduce_add_sync(0xffffffffu, s4[k]);
    }
    int plane;
    switch (deltaX)
    {
    case 32: plane = deltaY % 24 == 0 ? 1 : deltaY == 8 ? 2 : 0; break;
    case 24: plane = 3; break;
    case 16: plane = deltaY % 12 == 0 ? 5 : deltaY == 4 ? 6 : 4; break;
    case 12: plane = 7; break;
    case 8: plane = deltaY == 32 ? 8 : 9; break;
    case 4: plane = deltaY == 16 ? 10 : 11; break;
    default: plane = 11; break;
    }
    const uint32_t* sumsBase = sums + (size_t)plane * planePitch + offR[pu];
    intptr_t delta = deltaY;
    if ((w == h && w >= 16) || vertical || (w == 12 && h == 16) || (w == 4 && h == 16) || (w == 24 && h == 32) || (w == 8 && h == 32) ||
        (w == 48 && h == 64) || (w == 16 && h == 64))
        delta *= strideR;
    if (vertical) encDC[1] = encDC[2];
    if (horizontal) delta = deltaX;
    const int terms = ads_terms(w, h), half = w >> 1;

    for (int ty = minY; ty <= maxY; ty++)
    {
        const int ycost = (int)pcy[ty] << 2;
        if (best <= ycost) continue;
        best -= ycost;
        const uint32_t* srow = sumsBase + minX + (intptr_t)ty * strideR;
        int xn = 0;
        for (int base = 0; base < width; base += 32)
        {
            const int i = base + lane;
            bool hit = false;
            if (i < width)
            {
                long long ads;
                if (terms == 4)
                    ads = llabs(encDC[0] - (long long)srow[i]) + llabs(encDC[1] - (long long)srow[i + half])
                        + llabs(encDC[2] - (long long)srow[i + delta]) + llabs(encDC[3] - (long long)srow[i + delta + half]);
                else if (terms == 2)
                    ads = llabs(encDC[0] - (long long)srow[i]) + llabs(encDC[1] - (long long)srow[i + delta]);
                else
                    ads = llabs(encDC[0] - (long long)srow[i]);
                hit = (int)(ads + P.cx[(minX + i) * 4]) < best;         // m_fpelMvCosts: the x-cost at full-pel columns
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (hit) list[xn + __popc(m & ((1u << lane) - 1))] = (int16_t)i;
            xn += __popc(m);
        }
        __syncwarp();
        int i = 0;
        int cand[4][2], cost[4]; bool ok[4];
        for (; i < xn - 2; i += 3)
        {
            for (int k = 0; k < 4; k++) { cand[k][0] = minX + list[i + (k < 3 ? k : 2)]; cand[k][1] = ty; ok[k] = k < 3; }
            P.eval(3, cand, ok, cost, -2);
            for (int k = 0; k < 3; k++)
            {
                const int c = cost[k] + (int)pcx[cand[k][0] * 4];
                if (c < best) { best = c; x = cand[k][0]; y = ty; }
            }
        }
        best += ycost;
        if (i < xn)
        {
            const int rem = xn - i;                                     // 1 or 2
            for (int k = 0; k < 4; k++) { cand[k][0] = minX + list[i + (k < rem ? k : rem - 1)]; cand[k][1] = ty; ok[k] = k < rem; }
            P.eval(rem, cand, ok, cost);
            for (int k = 0; k < rem; k++)
                if (cost[k] < best) { best = cost[k]; x = cand[k][0]; y = ty; }
        }
        __syncwarp();
    }
    if (lane == 0) { bmv[2 * pu] = x; bmv[2 * pu + 1] = y; bcost[pu] = best; }
}

int launch_me_pattern(x265b200_ctx* ctx, int method, int w, int h, int merange, const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR,
                      const int32_t* offF, const int32_t* offR, const int32_t* range, const int32_t* mvp, const uint16_t* costTab, int n,
                      int32_t* bmv, int32_t* bcost, int numCand, const int32_t* mvc, x265b200_stream stream)
{
    if (n == 0) return X265B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int wpc = (method == X265B200_ME_UMH && w * h < 256) ? 1 : MP_WARPS;
    const size_t smem = (size_t)wpc * w * h * ctx->pixbytes;
    int phase = 0;
    int32_t* flags = nullptr;
#define MPK(PIX, WPC) me_pattern_kernel<PIX, WPC><<<ceil_div(n, WPC), WPC * 32, smem, st>>>(method, merange, (const PIX*)fenc, strideF, (const PIX*)ref, strideR, \
                                                                                 offF, offR, range, mvp, costTab, n, w, h, bmv, bcost, numCand, mvc, phase, flags)
#define MPL() do { if (ctx->pixbytes == 1) { if (wpc == 1) MPK(uint8_t, 1); else MPK(uint8_t, MP_WARPS); } \
                   else { if (wpc == 1) MPK(uint16_t, 1); else MPK(uint16_t, MP_WARPS); } } while (0)
    if (method == X265B200_ME_STAR && !lab_knob(1, 0))
    {   // star search split around its raster pass: pattern kernel up to the decision, CTA-per-PU raster for the PUs that take it, pattern kernel again
        B200_CUDA(ctx, cudaMallocAsync((void**)&flags, (size_t)n * sizeof(int32_t), st));
        B200_CUDA(ctx, cudaMemsetAsync(flags, 0, (size_t)n * sizeof(int32_t), st));
        phase = 1;
        MPL();
        if (ctx->pixbytes == 1)
            me_star_raster_kernel<uint8_t><<<n, RS_WARPS * 32, (size_t)w * h, st>>>((const uint8_t*)fenc, strideF, (const uint8_t*)ref, strideR, offF, offR, range, mvp, costTab,
                                                                                 w, h, flags, bmv, bcost);
        else
            me_star_raster_kernel<uint16_t><<<n, RS_WARPS * 32, (size_t)w * h * 2, st>>>((const uint16_t*)fenc, strideF, (const uint16_t*)ref, strideR, offF, offR, range, mvp, costTab,
                                                                                      w, h, flags, bmv, bcost);
        phase = 2;
        MPL();
        ctx->launches.fetch_add(2, std::memory_order_relaxed);
        cudaFreeAsync(flags, st);
    }
    else
        MPL();
#undef MPL
#undef MPK
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

} // namespace b200

using namespace b200;

extern "C" int x265b200_me_full_batch(x265b200_ctx* ctx, int w, int h, int merange, const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR,
                                      const int32_t* offF, const int32_t* offR, const int32_t* range, const int32_t* mvp,
                                      const uint16_t* costTab, int n, int32_t* bmv, int32_t* bcost, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (w < 4 || w > 64 || h < 4 || h > 64 || (w & 3) || (h & 3) || n < 0 || merange < 0) return fail(ctx, X265B200_ERR_ARG, "me_full: bad geometry");
    if (n == 0) return X265B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int span = 2 * (merange > 512 ? 512 : merange) + 1;   // sizing hint only: larger windows are walked in super-tiles
    cudaError_t e;
    // eight candidate rows per work item wherever the PU height allows: half the shared-memory loads per VABSDIFF
    if (ctx->pixbytes == 1 && (h & 7))
        e = launch_me_full<uint8_t, 4>(fenc, strideF, ref, strideR, offF, offR, range, mvp, costTab, w, h, span, n, bmv, bcost, st);
    else if (ctx->pixbytes == 1)
        e = launch_me_full<uint8_t, 8>(fenc, strideF, ref, strideR, offF, offR, range, mvp, costTab, w, h, span, n, bmv, bcost, st);
    else if (h & 7)
        e = launch_me_full<uint16_t, 4>(fenc, strideF, ref, strideR, offF, offR, range, mvp, costTab, w, h, span, n, bmv, bcost, st);
    else
        e = launch_me_full<uint16_t, 8>(fenc, strideF, ref, strideR, offF, offR, range, mvp, costTab, w, h, span, n, bmv, bcost, st);
    if (e != cudaSuccess) return fail(ctx, X265B200_ERR_CUDA, "me_full: shared memory attribute", e);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

extern "C" int x265b200_me_pattern_batch(x265b200_ctx* ctx, int method, int w, int h, int merange, const void* fenc, intptr_t strideF,
                                         const void* ref, intptr_t strideR, const int32_t* offF, const int32_t* offR, const int32_t* range,
                                         const int32_t* mvp, const uint16_t* costTab, int n, int32_t* bmv, int32_t* bcost, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (w < 4 || w > 64 || h < 4 || h > 64 || (w & 3) || (h & 3) || n < 0 || merange < 0) return fail(ctx, X265B200_ERR_ARG, "me_pattern: bad geometry");
    if (method != X265B200_ME_DIA && method != X265B200_ME_HEX && method != X265B200_ME_STAR)
        return fail(ctx, X265B200_ERR_ARG, "me_pattern: method must be DIA, HEX or STAR (UMH: x265b200_me_umh_batch, SEA: x265b200_me_sea_batch)");
    return b200::launch_me_pattern(ctx, method, w, h, merange, fenc, strideF, ref, strideR, offF, offR, range, mvp, costTab, n, bmv, bcost, 0, nullptr, stream);
}

extern "C" int x265b200_me_umh_batch(x265b200_ctx* ctx, int w, int h, int merange, const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR,
                                     const int32_t* offF, const int32_t* offR, const int32_t* range, const int32_t* qmvp, int numCand, const int32_t* mvc,
                                     const uint16_t* costTab, int n, int32_t* bmv, int32_t* bcost, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (w < 4 || w > 64 || h < 4 || h > 64 || (w & 3) || (h & 3) || n < 0 || merange < 0 || numCand < 0 || numCand > 16 || (numCand && !mvc))
        return fail(ctx, X265B200_ERR_ARG, "me_umh: bad geometry / candidates");
    return b200::launch_me_pattern(ctx, X265B200_ME_UMH, w, h, merange, fenc, strideF, ref, strideR, offF, offR, range, qmvp, costTab, n, bmv, bcost, numCand, mvc, stream);
}

extern "C" int x265b200_me_sea_batch(x265b200_ctx* ctx, int w, int h, int merange, const void* fenc, intptr_t strideF, const void* ref, intptr_t strideR,
                                     const int32_t* offF, const int32_t* offR, const int32_t* range, const int32_t* qmvp, const uint16_t* costTab,
                                     const uint32_t* sums, size_t planePitch, int n, int32_t* bmv, int32_t* bcost, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (w < 4 || w > 64 || h < 4 || h > 64 || (w & 3) || (h & 3) || n < 0 || merange < 0 || !sums)
        return fail(ctx, X265B200_ERR_ARG, "me_sea: bad geometry / integral planes");
    if ((w == 32 && h == 8) || (w == 8 && h == 32) || (w == 8 && h == 4) || (w == 4 && h == 8))
        return fail(ctx, X265B200_ERR_ARG, "me_sea: the reference reads stale samples of its 64-stride fenc cache for 32x8, 8x32, 8x4 and 4x8 PUs; "
                                            "these shapes have no defined result");
    if (n == 0) return X265B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int listCap = 2 * merange + 8;
    if (listCap > SEA_MAX_WIDTH) listCap = SEA_MAX_WIDTH;
    const size_t smem = (size_t)MP_WARPS * w * h * ctx->pixbytes + (size_t)MP_WARPS * listCap * sizeof(int16_t);
    if (2 * merange + 4 > SEA_MAX_WIDTH) return fail(ctx, X265B200_ERR_ARG, "me_sea: merange too large");
    if (ctx->pixbytes == 1)
        me_sea_kernel<uint8_t><<<ceil_div(n, MP_WARPS), MP_WARPS * 32, smem, st>>>(merange, (const uint8_t*)fenc, strideF, (const uint8_t*)ref, strideR, offF, offR, range,
                                                                                 qmvp, costTab, sums, planePitch, n, w, h, listCap, bmv, bcost);
    else
        me_sea_kernel<uint16_t><<<ceil_div(n, MP_WARPS), MP_WARPS * 32, smem, st>>>(merange, (const uint16_t*)fenc, strideF, (const uint16_t*)ref, strideR, offF, offR, range,
                                                                                  qmvp, costTab, sums, planePitch, n, w, h, listCap, bmv, bcost);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}
