/* ref_motion.cpp -- compiles the reference's MotionEstimate (encoder/motion.cpp), BitCost (encoder/bitcost.cpp) and the
 * Yuv block cache (common/yuv.cpp) into oracle/_ref/libx265ref_<depth>.so and drives
 * MotionEstimate::motionEstimate through the lookahead-style setSourcePU (luma only, no PicYuv), so the oracle's
 * restatement of the search is pinned against the reference's own search code, not just its SAD slots.
 * The sources are included from where they lie; nothing is copied.  Built like ref_framefilter.cpp
 * (-ffunction-sections, --gc-sections): whatever of those files needs the rest of the encoder is never linked.
 * x265_malloc / x265_free / general_log come from ref_shim.cpp's stand-ins (common/common.cpp would drag in the
 * parameter parser).  Test infrastructure, never shipped. */
#include "motion.cpp"
#include "bitcost.cpp"
#include "yuv.cpp"
#include <stdlib.h>
#include <thread>
#include <vector>

using namespace X265_NS;

extern "C" void ref_ensure(void);        /* ref_shim.cpp: C primitives into the global table */

extern "C" __attribute__((visibility("default")))
int ref_motion_estimate(int method, int subme, int w, int h, pixel* fencPlane, intptr_t strideF, intptr_t offF,
                        pixel* refPlane, intptr_t strideR, intptr_t offR, const int32_t* range /* mvmin.x, mvmin.y, mvmax.x, mvmax.y */,
                        const int32_t* qmvp, int numCand, const int32_t* mvc, int merange, int qp, int32_t* outQMv)
{
    static bool scales = false;
    if (!scales) { ref_ensure(); MotionEstimate::initScales(); scales = true; }
    MotionEstimate me;
    me.init(X265_CSP_I400);
    me.setQP(qp);
    me.setSourcePU(fencPlane, strideF, offF, w, h, method, subme);
    ReferencePlanes ref;
    ref.fpelPlane[0] = refPlane + offR;          /* setSourcePU's offset also shifts the reference block; offR is the rest */
    ref.lumaStride = strideR;
    MV mvmin(range[0], range[1]), mvmax(range[2], range[3]), mvp(qmvp[0], qmvp[1]), out;
    MV cands[16];
    for (int i = 0; i < numCand && i < 16; i++) cands[i] = MV(mvc[2 * i], mvc[2 * i + 1]);
    int cost = me.motionEstimate(&ref, mvmin, mvmax, mvp, numCand, cands, merange, out, 1, false);
    outQMv[0] = out.x; outQMv[1] = out.y;
    return cost;
}

/* the lookahead's use (slicetype.cpp:4484-4566): a lowres reference = four half-pel planes `pitch` samples apart,
 * ref->isLowres set, no neighbour candidates */
extern "C" __attribute__((visibility("default")))
int ref_lowres_motion_estimate(int method, int subme, int w, int h, pixel* fencPlane, intptr_t strideF, intptr_t offF,
                               pixel* planes, intptr_t strideR, size_t pitch, intptr_t offR, const int32_t* range,
                               const int32_t* qmvp, int merange, int qp, int32_t* outQMv)
{
    static bool scales = false;
    if (!scales) { ref_ensure(); MotionEstimate::initScales(); scales = true; }
    MotionEstimate me;
    me.init(X265_CSP_I400);
    me.setQP(qp);
    me.setSourcePU(fencPlane, strideF, offF, w, h, method, subme);
    ReferencePlanes ref;
    for (int k = 0; k < 4; k++) ref.lowresPlane[k] = planes + k * pitch + (offR - offF);
    ref.fpelPlane[0] = ref.lowresPlane[0];
    ref.lumaStride = strideR;
    ref.isLowres = true;
    MV mvmin(range[0], range[1]), mvmax(range[2], range[3]), mvp(qmvp[0], qmvp[1]), out;
    int cost = me.motionEstimate(&ref, mvmin, mvmax, mvp, 0, NULL, merange, out, 1, false);
    outQMv[0] = out.x; outQMv[1] = out.y;
    return cost;
}

/* the encoder's call: setSourcePU(const Yuv&, ctuAddr, cuPartIdx, puPartIdx, ..., bChroma) (motion.cpp:222-247), which
 * turns on the chroma SATD term of subpelCompare from subme 3.  The PU is handed over as the top-left block of a 64x64
 * CU-sized Yuv; the reference picture is a PicYuv whose CTU / partition offset tables are all zero, so every
 * get*Addr(ctuAddr = 0, absPartIdx = 0) lands on the co-located blocks given here.  csp: X265_CSP_I420 or I444. */
extern "C" __attribute__((visibility("default")))
int ref_motion_estimate_chroma(int method, int subme, int csp, int w, int h, const pixel* fencY, intptr_t sfY, const pixel* fencCb, const pixel* fencCr,
                               intptr_t sfC, pixel* refY, intptr_t srY, pixel* refCb, pixel* refCr, intptr_t srC, const int32_t* range,
                               const int32_t* qmvp, int numCand, const int32_t* mvc, int merange, int qp, int32_t* outQMv)
{
    static bool scales = false;
    if (!scales) { ref_ensure(); MotionEstimate::initScales(); scales = true; }
    MotionEstimate me;
    me.init(csp);
    me.setQP(qp);
    Yuv src;
    src.create(64, csp);
    const int hs = CHROMA_H_SHIFT(csp), vs = CHROMA_V_SHIFT(csp);
    for (int y = 0; y < h; y++) memcpy(src.m_buf[0] + y * src.m_size, fencY + y * sfY, w * sizeof(pixel));
    for (int y = 0; y < (h >> vs); y++)
    {
        memcpy(src.m_buf[1] + y * src.m_csize, fencCb + y * sfC, (w >> hs) * sizeof(pixel));
        memcpy(src.m_buf[2] + y * src.m_csize, fencCr + y * sfC, (w >> hs) * sizeof(pixel));
    }
    me.setSourcePU(src, 0, 0, 0, w, h, method, subme, true);
    static intptr_t zeros[4] = { 0, 0, 0, 0 };
    PicYuv* pic = (PicYuv*)calloc(1, sizeof(PicYuv));          /* plain data: only strides, origins and offset tables are read */
    pic->m_stride = srY; pic->m_strideC = srC;
    pic->m_picOrg[0] = refY; pic->m_picOrg[1] = refCb; pic->m_picOrg[2] = refCr;
    pic->m_cuOffsetY = pic->m_cuOffsetC = pic->m_buOffsetY = pic->m_buOffsetC = zeros;
    ReferencePlanes ref;
    ref.reconPic = pic;
    ref.fpelPlane[0] = refY; ref.fpelPlane[1] = refCb; ref.fpelPlane[2] = refCr;
    ref.lumaStride = srY; ref.chromaStride = srC;
    MV mvmin(range[0], range[1]), mvmax(range[2], range[3]), mvp(qmvp[0], qmvp[1]), out;
    MV cands[16];
    for (int i = 0; i < numCand && i < 16; i++) cands[i] = MV(mvc[2 * i], mvc[2 * i + 1]);
    int cost = me.motionEstimate(&ref, mvmin, mvmax, mvp, numCand, cands, merange, out, 1, false);
    outQMv[0] = out.x; outQMv[1] = out.y;
    int on = me.bChromaSATD ? 1 : 0;
    src.destroy();
    free(pic);
    return on ? cost : -1 - cost;                              /* negative: the chroma term was off for this PU (caller checks) */
}

/* X265_SEA: as ref_motion_estimate, plus the twelve SEA integral planes (`sums`, planePitch apart, computed over the padded
 * reference picture by ref_me_integral) handed to MotionEstimate::integral[] at the PU's co-located block, as
 * Search::predInterSearch does (search.cpp:2700) */
extern "C" __attribute__((visibility("default")))
int ref_motion_estimate_sea(int subme, int w, int h, pixel* fencPlane, intptr_t strideF, intptr_t offF,
                            pixel* refPlane, intptr_t strideR, intptr_t offR, uint32_t* sums, size_t planePitch,
                            const int32_t* range, const int32_t* qmvp, int numCand, const int32_t* mvc, int merange, int qp, int32_t* outQMv)
{
    static bool scales = false;
    if (!scales) { ref_ensure(); MotionEstimate::initScales(); scales = true; }
    MotionEstimate me;
    me.init(X265_CSP_I400);
    memset(me.fencPUYuv.m_buf[0], 0, 64 * 64 * sizeof(pixel));   /* SEA's DC sums read the whole 64-stride cache */
    me.setQP(qp);
    me.setSourcePU(fencPlane, strideF, offF, w, h, X265_SEA, subme);
    ReferencePlanes ref;
    ref.fpelPlane[0] = refPlane + (offR - offF);
    ref.lumaStride = strideR;
    for (int k = 0; k < INTEGRAL_PLANE_NUM; k++) me.integral[k] = sums + k * planePitch + offR;
    MV mvmin(range[0], range[1]), mvmax(range[2], range[3]), mvp(qmvp[0], qmvp[1]), out;
    MV cands[16];
    for (int i = 0; i < numCand && i < 16; i++) cands[i] = MV(mvc[2 * i], mvc[2 * i + 1]);
    int cost = me.motionEstimate(&ref, mvmin, mvmax, mvp, numCand, cands, merange, out, 1, false);
    outQMv[0] = out.x; outQMv[1] = out.y;
    return cost;
}

/* the lambda-scaled mv cost table BitCost::setQP builds, copied out for the other implementations: [-radius, radius] */
extern "C" __attribute__((visibility("default")))
void ref_mvcost_table(int qp, int radius, uint16_t* out)
{
    struct Peek : BitCost { const uint16_t* tab() const { return m_cost; } } bc;
    bc.setQP(qp);
    for (int i = -radius; i <= radius; i++) out[i + radius] = bc.tab()[i];
}

/* the same call for n PUs on nthreads host threads, one MotionEstimate object per thread (as the encoder's worker threads
 * hold one each): the CPU arm of tools/bench_me.py */
extern "C" __attribute__((visibility("default")))
void ref_motion_estimate_batch(int method, int subme, int w, int h, pixel* fencPlane, intptr_t strideF, const int32_t* offF,
                               pixel* refPlane, intptr_t strideR, const int32_t* offR, const int32_t* range, const int32_t* qmvp,
                               int numCand, const int32_t* mvc, int merange, int qp, int n, int32_t* outQMv, int32_t* outCost, int nthreads)
{
    static bool scales = false;
    if (!scales) { ref_ensure(); MotionEstimate::initScales(); scales = true; }
    { BitCost warm; warm.setQP(qp); }                       /* build the shared cost table before the threads race for it */
    auto work = [&](int lo, int hi)
    {
        MotionEstimate me;
        me.init(X265_CSP_I400);
        me.setQP(qp);
        for (int i = lo; i < hi; i++)
        {
            me.setSourcePU(fencPlane, strideF, offF[i], w, h, method, subme);
            ReferencePlanes ref;
            ref.fpelPlane[0] = refPlane + (offR[i] - offF[i]);
            ref.lumaStride = strideR;
            MV mvmin(range[4 * i], range[4 * i + 1]), mvmax(range[4 * i + 2], range[4 * i + 3]), mvp(qmvp[2 * i], qmvp[2 * i + 1]), out;
            MV cands[16];
            for (int k = 0; k < numCand && k < 16; k++) cands[k] = MV(mvc[((size_t)i * numCand + k) * 2], mvc[((size_t)i * numCand + k) * 2 + 1]);
            outCost[i] = me.motionEstimate(&ref, mvmin, mvmax, mvp, numCand, cands, merange, out, 1, false);
            outQMv[2 * i] = out.x; outQMv[2 * i + 1] = out.y;
        }
    };
    if (nthreads <= 1) { work(0, n); return; }
    std::vector<std::thread> th;
    int chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++)
    {
        int lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (lo >= hi) break;
        th.emplace_back(work, lo, hi);
    }
    for (auto& t : th) t.join();
}

/* the lookahead's predictor selection and bi-directional candidates (encoder/slicetype.cpp:4520-4558, 4577-4596) driven through the
 * reference's own ReferencePlanes::lowresMC (common/lowres.h:74-93), MotionEstimate::bufSATD and the pixelavg_pp slot: the loop bodies are
 * the reference's statements, only the surrounding Lowres / Lookahead objects are replaced by plain pointers */
static void ref_planes(ReferencePlanes& r, pixel* planes, intptr_t stride, size_t pitch)
{
    for (int k = 0; k < 4; k++) r.lowresPlane[k] = planes + k * pitch;
    r.fpelPlane[0] = r.lowresPlane[0];
    r.lumaStride = stride;
    r.isLowres = true;
}
extern "C" __attribute__((visibility("default")))
void ref_lowres_mvp(pixel* fencPlane, intptr_t strideF, intptr_t offF, pixel* planes, intptr_t strideR, size_t pitch, intptr_t offR,
                    const int32_t* mvcIn, int numc, int bBidir, int32_t* out)
{
    ref_ensure();
    MotionEstimate me;
    me.init(X265_CSP_I400);
    me.setSourcePU(fencPlane, strideF, offF, 8, 8, X265_HEX_SEARCH, 1);
    ReferencePlanes ref;
    ref_planes(ref, planes, strideR, pitch);
    ReferencePlanes* fref = &ref;
    const intptr_t pelOffset = offR;
    MV mvc[5], mvp;
    for (int i = 0; i < numc; i++) mvc[i] = MV(mvcIn[2 * i], mvcIn[2 * i + 1]);
    int skipCost = INT_MAX, mvpcost = MotionEstimate::COST_MAX;
    if (!numc)
        mvp = 0;
    else
    {
        ALIGN_VAR_32(pixel, subpelbuf[X265_LOWRES_CU_SIZE * X265_LOWRES_CU_SIZE]);
        for (int idx = 0; idx < numc; idx++)
        {
            intptr_t stride = X265_LOWRES_CU_SIZE;
            pixel *src = fref->lowresMC(pelOffset, mvc[idx], subpelbuf, stride, 0);
            int cost = me.bufSATD(src, stride);
            COPY2_IF_LT(mvpcost, cost, mvp, mvc[idx]);
            if (!mvp.notZero() && bBidir)
                skipCost = cost;
        }
    }
    out[0] = mvp.x; out[1] = mvp.y; out[2] = mvpcost; out[3] = skipCost;
}
extern "C" __attribute__((visibility("default")))
void ref_lowres_bidir(pixel* fencPlane, intptr_t strideF, intptr_t offF, pixel* planes0, intptr_t s0, size_t pitch0, pixel* planes1, intptr_t s1,
                      size_t pitch1, intptr_t offR, const int32_t* mv0, const int32_t* mv1, int32_t* out)
{
    ref_ensure();
    MotionEstimate me;
    me.init(X265_CSP_I400);
    me.setSourcePU(fencPlane, strideF, offF, 8, 8, X265_HEX_SEARCH, 1);
    ReferencePlanes r0, r1;
    ref_planes(r0, planes0, s0, pitch0);
    ref_planes(r1, planes1, s1, pitch1);
    ReferencePlanes *fref0 = &r0, *fref1 = &r1;
    const intptr_t pelOffset = offR;
    ALIGN_VAR_32(pixel, subpelbuf0[X265_LOWRES_CU_SIZE * X265_LOWRES_CU_SIZE]);
    ALIGN_VAR_32(pixel, subpelbuf1[X265_LOWRES_CU_SIZE * X265_LOWRES_CU_SIZE]);
    intptr_t stride0 = X265_LOWRES_CU_SIZE, stride1 = X265_LOWRES_CU_SIZE;
    pixel *src0 = fref0->lowresMC(pelOffset, MV(mv0[0], mv0[1]), subpelbuf0, stride0, 0);
    pixel *src1 = fref1->lowresMC(pelOffset, MV(mv1[0], mv1[1]), subpelbuf1, stride1, 0);
    ALIGN_VAR_32(pixel, ref[X265_LOWRES_CU_SIZE * X265_LOWRES_CU_SIZE]);
    primitives.pu[LUMA_8x8].pixelavg_pp[NONALIGNED](ref, X265_LOWRES_CU_SIZE, src0, stride0, src1, stride1, 32);
    out[0] = me.bufSATD(ref, X265_LOWRES_CU_SIZE);
    src0 = fref0->lowresPlane[0] + pelOffset;
    src1 = fref1->lowresPlane[0] + pelOffset;
    primitives.pu[LUMA_8x8].pixelavg_pp[NONALIGNED](ref, X265_LOWRES_CU_SIZE, src0, fref0->lumaStride, src1, fref1->lumaStride, 32);
    out[1] = me.bufSATD(ref, X265_LOWRES_CU_SIZE);
}
