/*
 * ref_shim.cpp -- C handle onto the REFERENCE's own C primitives (test infrastructure).
 *
 * Compiled ONLY where /root/reference exists (this container), together with the
 * reference's pixel.cpp / dct.cpp / lowpassdct.cpp / ipfilter.cpp / intrapred.cpp /
 * loopfilter.cpp / constants.cpp / primitives.cpp, straight from where they lie
 * (oracle/Makefile, target `ref`), into oracle/_ref/libx265ref_<depth>.so.
 * Nothing from the reference is copied into this repository; this file only
 * includes the reference's headers at build time and calls through the table that
 * setupCPrimitives() + setupAliasPrimitives() fill (primitives.h:471-474), exactly
 * as source/test/testbench.cpp:224-227 does.
 *
 * Uses: (1) pin oracle/x265_oracle.c (tests/test_oracle_vs_ref.py),
 *       (2) generate tests/golden/*.npz (tests/golden/make_golden.py),
 *       (3) CPU baseline "kind": "reference" in bench.py.
 * The product never loads this library.
 */
#include "common.h"
#include "primitives.h"
#include "constants.h"
#include "contexts.h"
#include "temporalfilter.h"

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <cstring>

/* ---- link stubs for subsystems that are off the hot path and not compiled in ----
 * (encoder/sao.cpp, encoder/framefilter.cpp, encoder/entropy.cpp, common/cpu.cpp,
 * common/common.cpp, common/temporalfilter.cpp pull in the whole encoder).  None of
 * these is reachable from a hot-path slot. */
namespace X265_NS {
void setupSaoPrimitives_c(EncoderPrimitives&) {}
void setupMCSTFPrimitives_scalar(MCSTFPrimitives&) {}
MCSTFPrimitives mcstfPrim;
const uint32_t g_entropyBits[128] = { 0 };
const uint8_t g_nextState[128][2] = { { 0 } };
const cpu_name_t cpu_names[] = { { "", 0 } };
uint32_t cpu_detect(bool) { return 0; }
void general_log(const x265_param*, const char*, int, const char*, ...) {}
void* x265_malloc(size_t size) { void* p = NULL; return posix_memalign(&p, 64, size) ? NULL : p; }
void x265_free(void* p) { free(p); }
}
extern "C" const uint32_t PFX(entropyStateBits)[128] = { 0 };

using namespace X265_NS;

namespace X265_NS {
extern const uint8_t lumaPartitionMapTable[];
void setupIntrinsicDCT_sse3(EncoderPrimitives&);      /* common/vec/dct-sse3.cpp  : idct 8/16/32 */
void setupIntrinsicDCT_ssse3(EncoderPrimitives&);     /* common/vec/dct-ssse3.cpp : dct 16/32    */
void setupIntrinsicDCT_sse41(EncoderPrimitives&);     /* common/vec/dct-sse41.cpp : dequant_scaling */
}

namespace {
/* persistent worker pool: bench.py's CPU arm calls a batch entry per primitive pass, and a fresh std::thread set per
 * pass (round 1) cost more than the small passes themselves.  Work is handed out in chunks from an atomic cursor, so a
 * slow core does not hold the pass back. */
class Pool
{
public:
    static Pool& get() { static Pool p; return p; }
    void run(int n, int nthreads, const std::function<void(int, int)>& f)
    {
        if (nthreads <= 1 || n < 2) { f(0, n); return; }
        std::unique_lock<std::mutex> callers(m_callers);        /* one parallel region at a time */
        grow(nthreads - 1);
        int chunk = n / (nthreads * 8);
        if (chunk < 1) chunk = 1;
        {
            std::lock_guard<std::mutex> g(m_mu);
            m_fn = &f; m_n = n; m_chunk = chunk; m_cursor = 0; m_active = nthreads - 1; m_pending = nthreads - 1; m_gen++;
        }
        m_cv.notify_all();
        work();
        std::unique_lock<std::mutex> g(m_mu);
        m_done.wait(g, [this] { return m_pending == 0; });
        m_fn = nullptr;
    }
private:
    Pool() {}
    ~Pool()
    {
        { std::lock_guard<std::mutex> g(m_mu); m_quit = true; m_gen++; }
        m_cv.notify_all();
        for (auto& t : m_threads) t.join();
    }
    void grow(int want)
    {
        while ((int)m_threads.size() < want)
        {
            int id = (int)m_threads.size();
            m_threads.emplace_back([this, id] { loop(id); });
        }
    }
    void work()
    {
        for (;;)
        {
            int lo = m_cursor.fetch_add(m_chunk);
            if (lo >= m_n) break;
            int hi = lo + m_chunk < m_n ? lo + m_chunk : m_n;
            (*m_fn)(lo, hi);
        }
    }
    void loop(int id)
    {
        uint64_t seen = 0;
        for (;;)
        {
            {
                std::unique_lock<std::mutex> g(m_mu);
                m_cv.wait(g, [&] { return m_gen != seen; });
                seen = m_gen;
                if (m_quit) return;
                if (id >= m_active) continue;
            }
            work();
            std::lock_guard<std::mutex> g(m_mu);
            if (--m_pending == 0) m_done.notify_all();
        }
    }
    std::vector<std::thread> m_threads;
    std::mutex m_mu, m_callers;
    std::condition_variable m_cv, m_done;
    const std::function<void(int, int)>* m_fn = nullptr;
    std::atomic<int> m_cursor{0};
    int m_n = 0, m_chunk = 1, m_active = 0, m_pending = 0;
    uint64_t m_gen = 0;
    bool m_quit = false;
};

template<typename F>
void parfor(int n, int nthreads, F f)
{
    Pool::get().run(n, nthreads, std::function<void(int, int)>(f));
}

EncoderPrimitives g_c;     // plain C table + aliases
EncoderPrimitives g_simd;  // g_c with the reference's SSE3 / SSSE3 / SSE4.1 intrinsic DCT tier layered on top (vec-primitives.cpp:58-81)
EncoderPrimitives g_lp;    // copy with enableLowpassDCTPrimitives applied (standard_dct bound)
bool g_ready = false;

void ensure()
{
    if (g_ready) return;
    memset(&g_c, 0, sizeof(g_c));
    setupCPrimitives(g_c);
    setupAliasPrimitives(g_c);
    /* HBD alias trampolines dispatch through the global table (primitives.cpp:98-168) */
    memcpy(&primitives, &g_c, sizeof(g_c));
    /* the best table this image can build: no nasm/yasm, so the .asm tier is out, but the intrinsic tier compiles with g++ */
    memcpy(&g_simd, &g_c, sizeof(g_c));
    setupIntrinsicDCT_sse3(g_simd);
    setupIntrinsicDCT_ssse3(g_simd);
    setupIntrinsicDCT_sse41(g_simd);
    g_ready = true;
}

int lumaPart(int w, int h)
{
    if ((w & 3) || (h & 3) || w < 4 || h < 4 || w > 64 || h > 64) return -1;
    int p = lumaPartitionMapTable[(((w >> 2) - 1) << 4) + ((h >> 2) - 1)];
    return p == 255 ? -1 : p;
}

/* find a (csp, lumaPart) whose chroma PU is w x h: 4:4:4 (w,h), 4:2:0 (2w,2h), 4:2:2 (2w,h) */
bool chromaSlot(int w, int h, int& csp, int& part, bool needSatd)
{
    int p;
    if ((p = lumaPart(w, h)) >= 0) { csp = X265_CSP_I444; part = p; return true; }
    if ((p = lumaPart(2 * w, 2 * h)) >= 0 && (!needSatd || g_c.chroma[X265_CSP_I420].pu[p].satd))
    { csp = X265_CSP_I420; part = p; return true; }
    if ((p = lumaPart(2 * w, h)) >= 0 && (!needSatd || g_c.chroma[X265_CSP_I422].pu[p].satd))
    { csp = X265_CSP_I422; part = p; return true; }
    return false;
}
}

extern "C" {

int ref_depth() { return X265_DEPTH; }
int ref_table_bytes() { return (int)sizeof(EncoderPrimitives); }

void ref_get_dct_matrix(int N, int16_t* out)
{
    const int16_t* t = N == 4 ? &g_t4[0][0] : N == 8 ? &g_t8[0][0] : N == 16 ? &g_t16[0][0] : &g_t32[0][0];
    memcpy(out, t, (size_t)N * N * sizeof(int16_t));
}
void ref_get_luma_taps(int16_t* out) { memcpy(out, g_lumaFilter, sizeof(int16_t) * 4 * 8); }
void ref_get_chroma_taps(int16_t* out) { memcpy(out, g_chromaFilter, sizeof(int16_t) * 8 * 4); }

/* fills the process-global table too (ref_motion.cpp's MotionEstimate reads it) */
void ref_ensure(void) { ensure(); }

/* ---- pixel metrics ---- */
int ref_sad(int w, int h, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{ ensure(); int p = lumaPart(w, h); return p < 0 ? -1 : g_c.pu[p].sad(a, sa, b, sb); }

void ref_sad_x3(int w, int h, const pixel* f, const pixel* r0, const pixel* r1, const pixel* r2, intptr_t rs, int32_t* res)
{ ensure(); g_c.pu[lumaPart(w, h)].sad_x3(f, r0, r1, r2, rs, res); }

void ref_sad_x4(int w, int h, const pixel* f, const pixel* r0, const pixel* r1, const pixel* r2, const pixel* r3, intptr_t rs, int32_t* res)
{ ensure(); g_c.pu[lumaPart(w, h)].sad_x4(f, r0, r1, r2, r3, rs, res); }

int ref_ads(int w, int h, int* encDC, uint32_t* sums, int delta, uint16_t* costMvX, int16_t* mvs, int width, int thresh)
{ ensure(); return g_c.pu[lumaPart(w, h)].ads(encDC, sums, delta, costMvX, mvs, width, thresh); }

int ref_satd(int w, int h, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    ensure();
    int p = lumaPart(w, h);
    if (p >= 0) return g_c.pu[p].satd(a, sa, b, sb);
    int csp, part;
    if (!chromaSlot(w, h, csp, part, true)) return -1;
    return g_c.chroma[csp].pu[part].satd(a, sa, b, sb);
}

/* sa8d slots: luma cu[] (square w == h), 4:2:0 chroma cu (w == h, cu = 2w), 4:2:2 chroma cu (h == 2w).
 * `chroma` = 0 luma table, 1 = 4:2:0 table, 2 = 4:2:2 table. */
int ref_sa8d(int chroma, int w, int h, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    ensure();
    if (chroma == 0) return g_c.cu[w == 4 ? 0 : w == 8 ? 1 : w == 16 ? 2 : w == 32 ? 3 : 4].sa8d(a, sa, b, sb);
    int cuw = 2 * w;
    int idx = cuw == 8 ? 1 : cuw == 16 ? 2 : cuw == 32 ? 3 : 4;
    return g_c.chroma[chroma == 1 ? X265_CSP_I420 : X265_CSP_I422].cu[idx].sa8d(a, sa, b, sb);
}

uint64_t ref_sse_pp(int chroma, int w, int h, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    ensure();
    (void)h;
    if (chroma == 0) return g_c.cu[w == 4 ? 0 : w == 8 ? 1 : w == 16 ? 2 : w == 32 ? 3 : 4].sse_pp(a, sa, b, sb);
    int cuw = 2 * w;
    int idx = cuw == 4 ? 0 : cuw == 8 ? 1 : cuw == 16 ? 2 : cuw == 32 ? 3 : 4;
    return g_c.chroma[chroma == 1 ? X265_CSP_I420 : X265_CSP_I422].cu[idx].sse_pp(a, sa, b, sb);
}
uint64_t ref_sse_ss(int w, const int16_t* a, intptr_t sa, const int16_t* b, intptr_t sb)
{ ensure(); return g_c.cu[w == 4 ? 0 : w == 8 ? 1 : w == 16 ? 2 : w == 32 ? 3 : 4].sse_ss(a, sa, b, sb); }
uint64_t ref_ssd_s(int w, const int16_t* a, intptr_t sa)
{ ensure(); return g_c.cu[w == 4 ? 0 : w == 8 ? 1 : w == 16 ? 2 : w == 32 ? 3 : 4].ssd_s[NONALIGNED](a, sa); }

/* ---- transforms ---- */
static int trIdx(int n) { return n == 4 ? 0 : n == 8 ? 1 : n == 16 ? 2 : 3; }
void ref_dct(int n, const int16_t* src, int16_t* dst, intptr_t stride) { ensure(); g_c.cu[trIdx(n)].dct(src, dst, stride); }
void ref_idct(int n, const int16_t* src, int16_t* dst, intptr_t stride) { ensure(); g_c.cu[trIdx(n)].idct(src, dst, stride); }
void ref_dst4(const int16_t* src, int16_t* dst, intptr_t stride) { ensure(); g_c.dst4x4(src, dst, stride); }
void ref_idst4(const int16_t* src, int16_t* dst, intptr_t stride) { ensure(); g_c.idst4x4(src, dst, stride); }
void ref_lowpass_dct(int n, const int16_t* src, int16_t* dst, intptr_t stride)
{
    ensure();
    /* lowpassdct.cpp keeps file-static pointers to the standard_dct slots of the table
     * setupLowPassPrimitives_c last saw (g_c); fill them the way
     * enableLowpassDCTPrimitives does (primitives.cpp:77-88), on g_c itself. */
    for (int i = 0; i < 4; i++)
        if (!g_c.cu[i].standard_dct) g_c.cu[i].standard_dct = g_c.cu[i].dct;
    g_c.cu[trIdx(n)].lowpass_dct(src, dst, stride);
}
uint32_t ref_quant(const int16_t* coef, const int32_t* qc, int32_t* deltaU, int16_t* qCoef, int qBits, int add, int n)
{ ensure(); return g_c.quant(coef, qc, deltaU, qCoef, qBits, add, n); }
uint32_t ref_nquant(const int16_t* coef, const int32_t* qc, int16_t* qCoef, int qBits, int add, int n)
{ ensure(); return g_c.nquant(coef, qc, qCoef, qBits, add, n); }
void ref_dequant_normal(const int16_t* q, int16_t* coef, int num, int scale, int shift)
{ ensure(); g_c.dequant_normal(q, coef, num, scale, shift); }
void ref_dequant_scaling(const int16_t* q, const int32_t* dq, int16_t* coef, int num, int per, int shift)
{ ensure(); g_c.dequant_scaling(q, dq, coef, num, per, shift); }

/* ---- interpolation: N = 8 luma pu[] slots, N = 4 chroma[csp].pu[] slots ---- */
#define PICK(lumaSlot, chromaSlotName)                                                        \
    ensure();                                                                                 \
    int part = lumaPart(w, h), csp = 0;                                                       \
    if (N == 8) { if (part < 0) return -1; }                                                  \
    else if (!chromaSlot(w, h, csp, part, false)) return -1;

int ref_interp_hpp(int N, int w, int h, const pixel* s, intptr_t ss, pixel* d, intptr_t ds, int idx)
{ PICK(0, 0); (N == 8 ? g_c.pu[part].luma_hpp : g_c.chroma[csp].pu[part].filter_hpp)(s, ss, d, ds, idx); return 0; }
int ref_interp_vpp(int N, int w, int h, const pixel* s, intptr_t ss, pixel* d, intptr_t ds, int idx)
{ PICK(0, 0); (N == 8 ? g_c.pu[part].luma_vpp : g_c.chroma[csp].pu[part].filter_vpp)(s, ss, d, ds, idx); return 0; }
int ref_interp_hps(int N, int w, int h, const pixel* s, intptr_t ss, int16_t* d, intptr_t ds, int idx, int ext)
{ PICK(0, 0); (N == 8 ? g_c.pu[part].luma_hps : g_c.chroma[csp].pu[part].filter_hps)(s, ss, d, ds, idx, ext); return 0; }
int ref_interp_vps(int N, int w, int h, const pixel* s, intptr_t ss, int16_t* d, intptr_t ds, int idx)
{ PICK(0, 0); (N == 8 ? g_c.pu[part].luma_vps : g_c.chroma[csp].pu[part].filter_vps)(s, ss, d, ds, idx); return 0; }
int ref_interp_vsp(int N, int w, int h, const int16_t* s, intptr_t ss, pixel* d, intptr_t ds, int idx)
{ PICK(0, 0); (N == 8 ? g_c.pu[part].luma_vsp : g_c.chroma[csp].pu[part].filter_vsp)(s, ss, d, ds, idx); return 0; }
int ref_interp_vss(int N, int w, int h, const int16_t* s, intptr_t ss, int16_t* d, intptr_t ds, int idx)
{ PICK(0, 0); (N == 8 ? g_c.pu[part].luma_vss : g_c.chroma[csp].pu[part].filter_vss)(s, ss, d, ds, idx); return 0; }
int ref_interp_hvpp(int N, int w, int h, const pixel* s, intptr_t ss, pixel* d, intptr_t ds, int ix, int iy)
{
    ensure();
    int part = lumaPart(w, h);
    if (N != 8 || part < 0) return -1;       /* only the luma table has an hv slot */
    g_c.pu[part].luma_hvpp(s, ss, d, ds, ix, iy);
    return 0;
}
int ref_p2s(int w, int h, const pixel* s, intptr_t ss, int16_t* d, intptr_t ds)
{
    ensure();
    int part = lumaPart(w, h), csp = 0;
    if (part >= 0) { g_c.pu[part].convert_p2s[NONALIGNED](s, ss, d, ds); return 0; }
    if (!chromaSlot(w, h, csp, part, false)) return -1;
    g_c.chroma[csp].pu[part].p2s[NONALIGNED](s, ss, d, ds);
    return 0;
}

/* ---- inter luma TU chain composed from the reference's own slots, in the order search.cpp:5536-5575 / quant.cpp:397-605 call them ---- */
/* ttype 1 = intra luma: the 4x4 TU goes through the dst4x4 / idst4x4 slots and skips the DC-only shortcut, as Quant::transformNxN /
 * invtransformNxN do (quant.cpp:430-441, :585-603) */
void ref_tu_chain_tt(int N, int ttype, const pixel* fenc, intptr_t sf, const pixel* pred, intptr_t sp, const int32_t* quantCoeff,
                     int qBits, int add, int dqScale, int dqShift, int16_t* qCoef, uint32_t* numSig, pixel* recon, intptr_t sr,
                     uint64_t* sseZero, uint64_t* sseRecon)
{
    ensure();
    const bool useDST = ttype == 1 && N == 4;
    int cu = trIdx(N);
    ALIGN_VAR_32(int16_t, resi[32 * 32]);
    ALIGN_VAR_32(int16_t, coef[32 * 32]);
    ALIGN_VAR_32(int16_t, dq[32 * 32]);
    ALIGN_VAR_32(int16_t, rec[32 * 32]);
    ALIGN_VAR_32(int32_t, deltaU[32 * 32]);
    ALIGN_VAR_32(int16_t, q[32 * 32]);
    g_c.cu[cu].sub_ps(resi, N, fenc, pred, sf, sp);
    if (useDST) g_c.dst4x4(resi, coef, N);
    else g_c.cu[cu].dct(resi, coef, N);
    uint32_t ns = g_c.quant(coef, quantCoeff, deltaU, q, qBits, add, N * N);
    memcpy(qCoef, q, sizeof(int16_t) * N * N);
    *numSig = ns;
    *sseZero = g_c.cu[cu].sse_pp(fenc, sf, pred, sp);
    if (!ns)
    {
        g_c.pu[cu].copy_pp(recon, sr, pred, sp);
        *sseRecon = *sseZero;
        return;
    }
    g_c.dequant_normal(q, dq, N * N, dqScale, dqShift);
    if (ns == 1 && q[0] != 0 && !useDST)
    {
        const int shift_1st = 7 - 6, add_1st = 1 << (shift_1st - 1);
        const int shift_2nd = 12 - (X265_DEPTH - 8) - 3, add_2nd = 1 << (shift_2nd - 1);
        int dc_val = (((dq[0] * (64 >> 6) + add_1st) >> shift_1st) * (64 >> 3) + add_2nd) >> shift_2nd;
        g_c.cu[cu].blockfill_s[NONALIGNED](rec, N, (int16_t)dc_val);
    }
    else if (useDST)
        g_c.idst4x4(dq, rec, N);
    else
        g_c.cu[cu].idct(dq, rec, N);
    g_c.cu[cu].add_ps[NONALIGNED](recon, sr, pred, rec, sp, N);
    *sseRecon = g_c.cu[cu].sse_pp(fenc, sf, recon, sr);
}
void ref_tu_chain(int N, const pixel* fenc, intptr_t sf, const pixel* pred, intptr_t sp, const int32_t* quantCoeff,
                  int qBits, int add, int dqScale, int dqShift, int16_t* qCoef, uint32_t* numSig, pixel* recon, intptr_t sr,
                  uint64_t* sseZero, uint64_t* sseRecon)
{
    ref_tu_chain_tt(N, 0, fenc, sf, pred, sp, quantCoeff, qBits, add, dqScale, dqShift, qCoef, numSig, recon, sr, sseZero, sseRecon);
}

/* ---- slot census for the coverage contract (SURVEY.md section 8a checklist) ---- */
int ref_count_nonnull_slots()
{
    ensure();
    const void* const* slots = (const void* const*)&g_c;
    int n = 0;
    for (size_t i = 0; i < sizeof(g_c) / sizeof(void*); i++) n += slots[i] != NULL;
    return n;
}

/* ---- batched, multi-threaded drivers over descriptor arrays (CPU baseline) ---- */
enum { OP_SAD = 0, OP_SATD = 1, OP_SA8D = 2, OP_SSE_PP = 3 };

int ref_pixelcmp_batch(int op, int w, int h, const pixel* A, intptr_t sa, const pixel* B, intptr_t sb,
                       const int32_t* offA, const int32_t* offB, int n, void* out, int nthreads)
{
    ensure();
    int part = lumaPart(w, h);
    if (part < 0) return -1;
    int cu = w == 4 ? 0 : w == 8 ? 1 : w == 16 ? 2 : w == 32 ? 3 : 4;
    pixelcmp_t f = op == OP_SAD ? g_c.pu[part].sad : op == OP_SATD ? g_c.pu[part].satd : g_c.cu[cu].sa8d;
    pixel_sse_t fs = g_c.cu[cu].sse_pp;
    if ((op == OP_SA8D || op == OP_SSE_PP) && w != h) return -1;
    parfor(n, nthreads, [=](int lo, int hi) {
        for (int i = lo; i < hi; i++)
        {
            if (op == OP_SSE_PP) ((uint64_t*)out)[i] = fs(A + offA[i], sa, B + offB[i], sb);
            else ((int32_t*)out)[i] = f(A + offA[i], sa, B + offB[i], sb);
        }
    });
    return 0;
}

/* tier = 0: plain C table; 1: C + the SSE intrinsic DCT tier (what x265_setup_primitives gives a CPU with SSE4.1 when the
 * .asm tier is not built) */
static int g_tier = 0;
void ref_set_tier(int tier) { ensure(); g_tier = tier ? 1 : 0; }
static const EncoderPrimitives& tab() { return g_tier ? g_simd : g_c; }

/* one transform slot of the chosen tier, for pinning the intrinsic functions against the C ones */
void ref_tier_dct(int tier, int n, const int16_t* src, int16_t* dst, intptr_t stride)
{ ensure(); (tier ? g_simd : g_c).cu[trIdx(n)].dct(src, dst, stride); }
void ref_tier_idct(int tier, int n, const int16_t* src, int16_t* dst, intptr_t stride)
{ ensure(); (tier ? g_simd : g_c).cu[trIdx(n)].idct(src, dst, stride); }

/* extendPicBorder (common/pixel.cpp:1044-1061) on a padded plane; pic points at sample (0, 0) */
void ref_extend_pic_border(pixel* pic, intptr_t stride, int width, int height, int marginX, int marginY)
{ ensure(); extendPicBorder(pic, stride, width, height, marginX, marginY); }

/* Quant::transformNxN's slot sequence for an inter luma TU (common/quant.cpp:397-480 without RDOQ / sign hiding):
 * sub_ps -> dct -> quant; levels[n*N*N], numSig[n] */
int ref_tu_forward_batch(int N, const pixel* A, intptr_t sa, const pixel* B, intptr_t sb, const int32_t* offA, const int32_t* offB, int n,
                         const int32_t* quantCoeff, int qBits, int add, int16_t* levels, uint32_t* numSig, int nthreads)
{
    ensure();
    int cu = trIdx(N);
    const EncoderPrimitives& t = tab();
    parfor(n, nthreads, [=, &t](int lo, int hi) {
        ALIGN_VAR_32(int16_t, resi[32 * 32]);
        ALIGN_VAR_32(int16_t, coef[32 * 32]);
        ALIGN_VAR_32(int32_t, deltaU[32 * 32]);
        for (int i = lo; i < hi; i++)
        {
            t.cu[cu].sub_ps(resi, N, A + offA[i], B + offB[i], sa, sb);
            t.cu[cu].dct(resi, coef, N);
            numSig[i] = t.quant(coef, quantCoeff, deltaU, levels + (size_t)i * N * N, qBits, add, N * N);
        }
    });
    return 0;
}

/* residual via the reference's own sub_ps slot (pixel.cpp), then dct slot, per block */
int ref_residual_dct_batch(int N, const pixel* A, intptr_t sa, const pixel* B, intptr_t sb,
                           const int32_t* offA, const int32_t* offB, int n, int16_t* out, int nthreads)
{
    ensure();
    int cu = trIdx(N);
    const EncoderPrimitives& t = tab();
    parfor(n, nthreads, [=, &t](int lo, int hi) {
        ALIGN_VAR_32(int16_t, resi[32 * 32]);
        for (int i = lo; i < hi; i++)
        {
            t.cu[cu].sub_ps(resi, N, A + offA[i], B + offB[i], sa, sb);
            t.cu[cu].dct(resi, out + (size_t)i * N * N, N);
        }
    });
    return 0;
}

int ref_dct_batch(int N, const int16_t* src, intptr_t srcStride, const int32_t* off, int n, int16_t* out, int nthreads)
{
    ensure();
    int cu = trIdx(N);
    const EncoderPrimitives& t = tab();
    parfor(n, nthreads, [=, &t](int lo, int hi) {
        for (int i = lo; i < hi; i++)
            t.cu[cu].dct(src + off[i], out + (size_t)i * N * N, srcStride);
    });
    return 0;
}


/* ---- adjacent slots through the reference's own table ---- */
int ref_blockop(int op, int w, int h, const void* A, intptr_t sa, const void* B, intptr_t sb, void* D, intptr_t sd)
{
    ensure();
    if (op == 2 || op == 3)
    {
        int p = lumaPart(w, h), csp = 0;
        if (p < 0 && op == 3)
        {
            if (!chromaSlot(w, h, csp, p, false)) return -1;
            g_c.chroma[csp].pu[p].addAvg[NONALIGNED]((const int16_t*)A, (const int16_t*)B, (pixel*)D, sa, sb, sd);
            return 0;
        }
        if (p < 0) return -1;
        if (op == 2) g_c.pu[p].pixelavg_pp[NONALIGNED]((pixel*)D, sd, (const pixel*)A, sa, (const pixel*)B, sb, 32);
        else g_c.pu[p].addAvg[NONALIGNED]((const int16_t*)A, (const int16_t*)B, (pixel*)D, sa, sb, sd);
        return 0;
    }
    /* cu slots: square luma sizes, 4:2:0 (w == h, half size) shares them; 4:2:2 chroma CUs are w x 2w */
    int cu = -1;
    for (int i = 0; i < NUM_CU_SIZES; i++) if ((4 << i) == w) cu = i;
    if (w == h && cu >= 0)
    {
        if (op == 0) g_c.cu[cu].sub_ps((int16_t*)D, sd, (const pixel*)A, (const pixel*)B, sa, sb);
        else g_c.cu[cu].add_ps[NONALIGNED]((pixel*)D, sd, (const pixel*)A, (const int16_t*)B, sa, sb);
        return 0;
    }
    for (int i = 0; i < NUM_CU_SIZES; i++)
    {
        if (w == h && (2 << i) == w)
        {
            if (op == 0) g_c.chroma[X265_CSP_I420].cu[i].sub_ps((int16_t*)D, sd, (const pixel*)A, (const pixel*)B, sa, sb);
            else g_c.chroma[X265_CSP_I420].cu[i].add_ps[NONALIGNED]((pixel*)D, sd, (const pixel*)A, (const int16_t*)B, sa, sb);
            return 0;
        }
        if (h == 2 * w && (2 << i) == w)
        {
            if (op == 0) g_c.chroma[X265_CSP_I422].cu[i].sub_ps((int16_t*)D, sd, (const pixel*)A, (const pixel*)B, sa, sb);
            else g_c.chroma[X265_CSP_I422].cu[i].add_ps[NONALIGNED]((pixel*)D, sd, (const pixel*)A, (const int16_t*)B, sa, sb);
            return 0;
        }
    }
    return -1;
}

void ref_lowres(const pixel* src, intptr_t ss, pixel* d0, pixel* dh, pixel* dv, pixel* dc, intptr_t ds, int width, int height)
{ ensure(); g_c.frameInitLowres(src, d0, dh, dv, dc, ss, ds, width, height); }


/* subpelCompare's slot sequence (encoder/motion.cpp:1795-1811) through the reference table */
int ref_subpel_cmp(int op, int w, int h, const pixel* fenc, intptr_t sf, const pixel* fref, intptr_t sr, int xFrac, int yFrac)
{
    ensure();
    int part = lumaPart(w, h);
    if (part < 0) return -1;
    pixelcmp_t cmp = op ? g_c.pu[part].satd : g_c.pu[part].sad;
    if (!(yFrac | xFrac)) return cmp(fenc, sf, fref, sr);
    ALIGN_VAR_32(pixel, subpelbuf[64 * 64]);
    if (!yFrac) g_c.pu[part].luma_hpp(fref, sr, subpelbuf, w, xFrac);
    else if (!xFrac) g_c.pu[part].luma_vpp(fref, sr, subpelbuf, w, yFrac);
    else g_c.pu[part].luma_hvpp(fref, sr, subpelbuf, w, xFrac, yFrac);
    return cmp(fenc, sf, subpelbuf, w);
}


/* ---- SEA integral primitives: the reference's own functions (framefilter.cpp via ref_framefilter.cpp) ---- */
static int integralIdx(int size) { return size == 4 ? INTEGRAL_4 : size == 8 ? INTEGRAL_8 : size == 12 ? INTEGRAL_12 : size == 16 ? INTEGRAL_16 : size == 24 ? INTEGRAL_24 : size == 32 ? INTEGRAL_32 : -1; }
int ref_integral_inith(int W, uint32_t* sum, pixel* pix, intptr_t stride)
{ ensure(); int i = integralIdx(W); if (i < 0 || !g_c.integral_inith[i]) return -1; g_c.integral_inith[i](sum, pix, stride); return 0; }
int ref_integral_initv(int H, uint32_t* sum, intptr_t stride)
{ ensure(); int i = integralIdx(H); if (i < 0 || !g_c.integral_initv[i]) return -1; g_c.integral_initv[i](sum, stride); return 0; }
/* the row loop of FrameFilter::computeMEIntegral (framefilter.cpp:770-832) over one padded picture, t = y + padY */
int ref_me_integral(pixel* pix, intptr_t stride, int rows, uint32_t* sums, size_t planePitch)
{
    ensure();
    static const int W[12] = { 32, 32, 32, 24, 16, 16, 16, 12, 8, 8, 4, 4 }, H[12] = { 32, 24, 8, 32, 16, 12, 4, 16, 32, 8, 16, 4 };
    if (!g_c.integral_inith[INTEGRAL_4]) return -1;
    for (int k = 0; k < 12; k++) memset(sums + k * planePitch, 0, stride * sizeof(uint32_t));
    for (int t = 0; t < rows - 1; t++)
        for (int k = 0; k < 12; k++)
        {
            uint32_t* S = sums + k * planePitch;
            g_c.integral_inith[integralIdx(W[k])](S + (intptr_t)(t + 1) * stride, pix + (intptr_t)t * stride, stride);
            if (t >= H[k]) g_c.integral_initv[integralIdx(H[k])](S + (intptr_t)(t + 1 - H[k]) * stride, stride);
        }
    return 0;
}


/* ---- weighted prediction slots and the lookahead's weightCostLuma composition (slicetype.cpp:866-897) ---- */
void ref_weight_pp(const pixel* src, pixel* dst, intptr_t stride, int width, int height, int w0, int round, int shift, int offset)
{ ensure(); g_c.weight_pp(src, dst, stride, width, height, w0, round, shift, offset); }
void ref_weight_sp(const int16_t* src, pixel* dst, intptr_t ss, intptr_t ds, int width, int height, int w0, int round, int shift, int offset)
{ ensure(); g_c.weight_sp(src, dst, ss, ds, width, height, w0, round, shift, offset); }
void ref_weight_cost(const pixel* fenc, const pixel* ref, intptr_t stride, int width, int height, const int32_t* intraCost,
                     const int32_t* weights, int K, uint32_t* cost, pixel* tmp)
{
    ensure();
    int pw = (width + 7) & ~7, ph = (height + 7) & ~7;
    for (int k = 0; k < K; k++)
    {
        const int32_t* w = weights + 4 * k;
        const pixel* src = ref;
        if (w[2] >= 0) { g_c.weight_pp(ref, tmp, stride, pw, ph, w[0], w[1], w[2], w[3]); src = tmp; }
        uint32_t c = 0;
        int mb = 0;
        for (int y = 0; y < height; y += 8)
            for (int x = 0; x < width; x += 8, mb++)
            {
                int satd = g_c.pu[LUMA_8x8].satd(src + y * stride + x, stride, fenc + y * stride + x, stride);
                c += intraCost ? X265_MIN(satd, intraCost[mb]) : satd;
            }
        cost[k] = c;
    }
}


/* ---- copy family through the reference table: luma cu[] slots for square sizes, pu[].copy_pp for any luma PU ---- */
int ref_blockcopy(int kind, int w, int h, void* dst, intptr_t ds, const void* src, intptr_t ss, int param)
{
    ensure();
    if (kind == 0)
    {
        int p = lumaPart(w, h);
        if (p < 0) return -1;
        g_c.pu[p].copy_pp((pixel*)dst, ds, (const pixel*)src, ss);
        return 0;
    }
    int cu = -1;
    for (int i = 0; i < NUM_CU_SIZES; i++) if ((4 << i) == w) cu = i;
    if (w != h || cu < 0) return -1;
    switch (kind)
    {
    case 1: g_c.cu[cu].copy_ss((int16_t*)dst, ds, (const int16_t*)src, ss); break;
    case 2: g_c.cu[cu].copy_sp((pixel*)dst, ds, (const int16_t*)src, ss); break;
    case 3: g_c.cu[cu].copy_ps((int16_t*)dst, ds, (const pixel*)src, ss); break;
    case 4: g_c.cu[cu].blockfill_s[NONALIGNED]((int16_t*)dst, ds, (int16_t)param); break;
    case 5: if (ds == w) g_c.cu[cu].cpy2Dto1D_shl((int16_t*)dst, (const int16_t*)src, ss, param);
            else if (ss == w) g_c.cu[cu].cpy1Dto2D_shl[NONALIGNED]((int16_t*)dst, (const int16_t*)src, ds, param); else return -1; break;
    case 6: if (ds == w) g_c.cu[cu].cpy2Dto1D_shr((int16_t*)dst, (const int16_t*)src, ss, param);
            else if (ss == w) g_c.cu[cu].cpy1Dto2D_shr((int16_t*)dst, (const int16_t*)src, ds, param); else return -1; break;
    default: return -1;
    }
    return 0;
}


/* ---- per-block scalars through the reference table ---- */
static int cuIdx(int size) { for (int i = 0; i < NUM_CU_SIZES; i++) if ((4 << i) == size) return i; return -1; }
uint64_t ref_var(int size, const pixel* pix, intptr_t stride) { ensure(); return g_c.cu[cuIdx(size)].var(pix, stride); }
int ref_psy_cost_pp(int size, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb) { ensure(); return g_c.cu[cuIdx(size)].psy_cost_pp(a, sa, b, sb); }
uint32_t ref_copy_cnt(int size, int16_t* coeff, const int16_t* resi, intptr_t stride)
{
    ensure();
    int i = cuIdx(size);
    if (coeff) return g_c.cu[i].copy_cnt(coeff, resi, stride);
    return (uint32_t)g_c.cu[i].count_nonzero(resi);          /* contiguous size x size input */
}
void ref_denoise_dct(int16_t* dct, uint32_t* resSum, const uint16_t* offset, int numCoeff) { ensure(); g_c.denoiseDct(dct, resSum, offset, numCoeff); }

/* ---- exhaustive integer search with the reference's own sad / sad_x4 slots, in the shape of motion.cpp's X265_FULL_SEARCH loop ---- */
void ref_me_full_search(int w, int h, const pixel* fencIn, intptr_t sf, const pixel* fref, intptr_t stride,
                        const int32_t* range, const int32_t* mvp, const uint16_t* costTab, int32_t* bmvIO, int32_t* bcostIO)
{
    ensure();
    int part = lumaPart(w, h);
    ALIGN_VAR_32(pixel, fenc[64 * 64]);                       /* the encoder keeps fenc at FENC_STRIDE */
    for (int y = 0; y < h; y++) memcpy(fenc + y * FENC_STRIDE, fencIn + y * sf, w * sizeof(pixel));
    pixelcmp_t sad = g_c.pu[part].sad;
    pixelcmp_x4_t sad_x4 = g_c.pu[part].sad_x4;
    const uint16_t* cx = costTab - mvp[0];
    const uint16_t* cy = costTab - mvp[1];
    int bcost = *bcostIO, bx = bmvIO[0], by = bmvIO[1];
    int32_t costs[4];
    for (int ty = range[1]; ty <= range[3]; ty++)
        for (int tx = range[0]; tx <= range[2]; tx++)
        {
            if (tx + 3 <= range[2])
            {
                const pixel* base = fref + (intptr_t)ty * stride + tx;
                sad_x4(fenc, base, base + 1, base + 2, base + 3, stride, costs);
                for (int k = 0; k < 4; k++)
                {
                    int c = costs[k] + (uint16_t)(cx[(tx + k) << 2] + cy[ty << 2]);
                    if (c < bcost) { bcost = c; bx = tx + k; by = ty; }
                }
                tx += 3;
            }
            else
            {
                int c = sad(fenc, FENC_STRIDE, fref + (intptr_t)ty * stride + tx, stride) + (uint16_t)(cx[tx << 2] + cy[ty << 2]);
                if (c < bcost) { bcost = c; bx = tx; by = ty; }
            }
        }
    *bcostIO = bcost; bmvIO[0] = bx; bmvIO[1] = by;
}

/* ---- bi-prediction cost through the reference table, in search.cpp:442-448's slot sequence ---- */
static void ref_mc_luma(int part, int w, const pixel* fref, intptr_t sr, int xFrac, int yFrac, pixel* dst)
{
    if (!(xFrac | yFrac)) g_c.pu[part].copy_pp(dst, 64, fref, sr);
    else if (!yFrac) g_c.pu[part].luma_hpp(fref, sr, dst, 64, xFrac);
    else if (!xFrac) g_c.pu[part].luma_vpp(fref, sr, dst, 64, yFrac);
    else g_c.pu[part].luma_hvpp(fref, sr, dst, 64, xFrac, yFrac);
}
int ref_bidir_satd(int w, int h, const pixel* fenc, intptr_t sf, const pixel* ref0, intptr_t sr0, int frac0,
                   const pixel* ref1, intptr_t sr1, int frac1)
{
    ensure();
    int part = lumaPart(w, h);
    if (part < 0) return -1;
    ALIGN_VAR_32(pixel, p0[64 * 64]); ALIGN_VAR_32(pixel, p1[64 * 64]); ALIGN_VAR_32(pixel, avg[64 * 64]);
    ref_mc_luma(part, w, ref0, sr0, frac0 & 3, (frac0 >> 4) & 3, p0);
    ref_mc_luma(part, w, ref1, sr1, frac1 & 3, (frac1 >> 4) & 3, p1);
    g_c.pu[part].pixelavg_pp[NONALIGNED](avg, 64, p0, 64, p1, 64, 32);
    return g_c.pu[part].satd(fenc, sf, avg, 64);
}

/* ---- intra prediction slots and the lookahead's intra estimate in lowresIntraEstimate's slot sequence (slicetype.cpp:781-841) ---- */
void ref_intra_filter(int N, const pixel* s, pixel* f) { ensure(); g_c.cu[cuIdx(N)].intra_filter(s, f); }
void ref_intra_pred(int N, int mode, const pixel* s, int bFilter, pixel* dst, intptr_t ds) { ensure(); g_c.cu[cuIdx(N)].intra_pred[mode](dst, ds, s, mode, bFilter); }
void ref_intra_allangs(int N, pixel* dest, pixel* refPix, pixel* filtPix, int bLuma) { ensure(); g_c.cu[cuIdx(N)].intra_pred_allangs(dest, refPix, filtPix, bLuma); }
int ref_lowres_intra_cu(const pixel* plane, intptr_t stride, int cuX, int cuY, int penalty, int32_t* modeOut)
{
    ensure();
    const int cuSize = X265_LOWRES_CU_SIZE, cuSize2 = cuSize << 1, sizeIdx = X265_LOWRES_CU_BITS - 2;
    ALIGN_VAR_32(pixel, prediction[X265_LOWRES_CU_SIZE * X265_LOWRES_CU_SIZE]);
    pixel fencIntra[X265_LOWRES_CU_SIZE * X265_LOWRES_CU_SIZE];
    pixel neighbours[2][X265_LOWRES_CU_SIZE * 4 + 1];
    pixel* samples = neighbours[0], *filtered = neighbours[1];
    pixelcmp_t satd = g_c.pu[sizeIdx].satd;
    const pixel* pixCur = plane + cuSize * cuX + (intptr_t)cuSize * cuY * stride;
    g_c.cu[sizeIdx].copy_pp(fencIntra, cuSize, pixCur, stride);
    pixCur -= stride + 1;
    memcpy(samples, pixCur, (2 * cuSize + 1) * sizeof(pixel));
    for (int i = 1; i <= 2 * cuSize; i++) samples[cuSize2 + i] = pixCur[i * stride];
    g_c.cu[sizeIdx].intra_filter(samples, filtered);
    int cost, icost = 1 << 28, ilowmode = 0;
    g_c.cu[sizeIdx].intra_pred[DC_IDX](prediction, cuSize, samples, 0, cuSize <= 16);
    cost = satd(fencIntra, cuSize, prediction, cuSize);
    if (cost < icost) { icost = cost; ilowmode = DC_IDX; }
    g_c.cu[sizeIdx].intra_pred[PLANAR_IDX](prediction, cuSize, neighbours[cuSize >= 8], 0, 0);
    cost = satd(fencIntra, cuSize, prediction, cuSize);
    if (cost < icost) { icost = cost; ilowmode = PLANAR_IDX; }
    int acost = 1 << 28, alowmode = 4, filter;
    for (int mode = 5; mode < 35; mode += 5)
    {
        filter = !!(g_intraFilterFlags[mode] & cuSize);
        g_c.cu[sizeIdx].intra_pred[mode](prediction, cuSize, neighbours[filter], mode, cuSize <= 16);
        cost = satd(fencIntra, cuSize, prediction, cuSize);
        if (cost < acost) { acost = cost; alowmode = mode; }
    }
    for (int dist = 2; dist >= 1; dist--)
    {
        int two[2] = { alowmode - dist, alowmode + dist };
        for (int k = 0; k < 2; k++)
        {
            int mode = two[k];
            filter = !!(g_intraFilterFlags[mode] & cuSize);
            g_c.cu[sizeIdx].intra_pred[mode](prediction, cuSize, neighbours[filter], mode, cuSize <= 16);
            cost = satd(fencIntra, cuSize, prediction, cuSize);
            if (cost < acost) { acost = cost; alowmode = mode; }
        }
    }
    if (acost < icost) { icost = acost; ilowmode = alowmode; }
    *modeOut = ilowmode;
    return icost + penalty;
}

} // extern "C"
