/*
 * harness_main.cpp -- runs the REFERENCE's own TestBench harness classes, unmodified, against the
 * B200 table (test infrastructure; built only where /root/reference exists, oracle/Makefile
 * target `harness`, into oracle/_ref/testbench_b200_<depth>).
 *
 * Mirrors reference source/test/testbench.cpp:224-265: C table = setupCPrimitives +
 * setupAliasPrimitives; table under test = zeroed + setupB200Primitives + setupAliasPrimitives,
 * copied into the global `primitives` because HBD aliases dispatch through it; then
 * harness->testCorrectness(cprim, b200prim) for PixelHarness, MBDstHarness, IPFilterHarness, IntraPredHarness.
 * Also checks the coverage contract: every hot-path slot that is non-NULL in the C table is bound to
 * a B200 thunk (not NULL, not the C function).
 *
 * The B200 filler is loaded with dlopen from ../../x265-mod-by-patman_b200/lib so this binary does
 * not link CUDA itself.  usage: testbench_b200_<depth> [seed] [pixel|transforms|interp|intrapred]
 */
#include "common.h"
#include "primitives.h"
#include "pixelharness.h"
#include "mbdstharness.h"
#include "ipfilterharness.h"
#include "intrapredharness.h"

#include <dlfcn.h>
#include <libgen.h>
#include <unistd.h>
#include <stdio.h>
#include <string.h>
#include <string>

using namespace X265_NS;

/* testbench.cpp:40-82 -- names the harness classes print; required link-time symbols */
const char* lumaPartStr[NUM_PU_SIZES] = {
    "  4x4", "  8x8", "16x16", "32x32", "64x64", "  8x4", "  4x8", " 16x8", " 8x16", "32x16", "16x32", "64x32", "32x64",
    "16x12", "12x16", " 16x4", " 4x16", "32x24", "24x32", " 32x8", " 8x32", "64x48", "48x64", "64x16", "16x64" };
static const char* chroma420Str[NUM_PU_SIZES] = {
    "  2x2", "  4x4", "  8x8", "16x16", "32x32", "  4x2", "  2x4", "  8x4", "  4x8", " 16x8", " 8x16", "32x16", "16x32",
    "  8x6", "  6x8", "  8x2", "  2x8", "16x12", "12x16", " 16x4", " 4x16", "32x24", "24x32", " 32x8", " 8x32" };
static const char* chroma422Str[NUM_PU_SIZES] = {
    "  2x4", "  4x8", " 8x16", "16x32", "32x64", "  4x4", "  2x8", "  8x8", " 4x16", "16x16", " 8x32", "32x32", "16x64",
    " 8x12", " 6x16", "  8x4", " 2x16", "16x24", "12x32", " 16x8", " 4x32", "32x48", "24x64", "32x16", " 8x64" };
const char* const* chromaPartStr[X265_CSP_COUNT] = { lumaPartStr, chroma420Str, chroma422Str, lumaPartStr };

static PixelHarness HPixel;
static MBDstHarness HMBDist;
static IPFilterHarness HIPFilter;
static IntraPredHarness HIPred;

static EncoderPrimitives cprim, b200prim;

#define SLOT(expr) do { const void* c__ = (const void*)cprim.expr; const void* o__ = (const void*)b200prim.expr; \
        if (c__) { want++; if (!o__) { missing++; printf("  missing: %s\n", #expr); } else if (o__ == c__ && !aliasOK) { same++; printf("  still C: %s\n", #expr); } else bound++; } \
        else if (o__) { extra++; printf("  extra: %s\n", #expr); } } while (0)

static int coverage()
{
    int want = 0, bound = 0, missing = 0, same = 0, extra = 0;
    bool aliasOK = false;
    for (int i = 0; i < NUM_PU_SIZES; i++)
    {
        SLOT(pu[i].sad); SLOT(pu[i].sad_x3); SLOT(pu[i].sad_x4); SLOT(pu[i].ads); SLOT(pu[i].satd);
        SLOT(pu[i].luma_hpp); SLOT(pu[i].luma_hps); SLOT(pu[i].luma_vpp); SLOT(pu[i].luma_vps); SLOT(pu[i].luma_vsp);
        SLOT(pu[i].luma_vss); SLOT(pu[i].luma_hvpp); SLOT(pu[i].convert_p2s[0]); SLOT(pu[i].convert_p2s[1]);
        for (int c = 1; c < X265_CSP_COUNT; c++)
        {
            SLOT(chroma[c].pu[i].filter_hpp); SLOT(chroma[c].pu[i].filter_hps); SLOT(chroma[c].pu[i].filter_vpp);
            SLOT(chroma[c].pu[i].filter_vps); SLOT(chroma[c].pu[i].filter_vsp); SLOT(chroma[c].pu[i].filter_vss);
            SLOT(chroma[c].pu[i].p2s[0]); SLOT(chroma[c].pu[i].p2s[1]); SLOT(chroma[c].pu[i].satd);
        }
    }
    for (int i = 0; i < NUM_CU_SIZES; i++)
    {
        SLOT(cu[i].sse_ss); SLOT(cu[i].ssd_s[0]); SLOT(cu[i].ssd_s[1]); SLOT(cu[i].sa8d);
        SLOT(cu[i].dct); SLOT(cu[i].idct); SLOT(cu[i].lowpass_dct);
        for (int c = 1; c < X265_CSP_COUNT; c++) SLOT(chroma[c].cu[i].sa8d);
#if HIGH_BIT_DEPTH
        aliasOK = true;     /* sse_pp is the shared trampoline onto primitives.cu[i].sse_ss (primitives.cpp:98-104) */
#endif
        SLOT(cu[i].sse_pp);
        for (int c = 1; c < X265_CSP_COUNT; c++) SLOT(chroma[c].cu[i].sse_pp);
        aliasOK = false;
    }
    SLOT(dst4x4); SLOT(idst4x4); SLOT(quant); SLOT(nquant); SLOT(dequant_normal); SLOT(dequant_scaling);
    printf("coverage: %d hot-path slots in the C table, %d bound to B200 entries, %d missing, %d still C, %d extra\n",
           want, bound, missing, same, extra);
    int hot = missing + same + extra;
    /* adjacent slots: same chains, not named by the north star */
    want = bound = missing = same = extra = 0;
    for (int i = 0; i < NUM_PU_SIZES; i++)
    {
        SLOT(pu[i].pixelavg_pp[0]); SLOT(pu[i].pixelavg_pp[1]); SLOT(pu[i].addAvg[0]); SLOT(pu[i].addAvg[1]); SLOT(pu[i].copy_pp);
        for (int c = 1; c < X265_CSP_COUNT; c++) { SLOT(chroma[c].pu[i].addAvg[0]); SLOT(chroma[c].pu[i].addAvg[1]); SLOT(chroma[c].pu[i].copy_pp); }
    }
    for (int i = 0; i < NUM_CU_SIZES; i++)
    {
        SLOT(cu[i].sub_ps); SLOT(cu[i].add_ps[0]); SLOT(cu[i].add_ps[1]);
        SLOT(cu[i].var); SLOT(cu[i].psy_cost_pp); SLOT(cu[i].count_nonzero); SLOT(cu[i].copy_cnt);
        SLOT(cu[i].blockfill_s[0]); SLOT(cu[i].blockfill_s[1]); SLOT(cu[i].calcresidual[0]); SLOT(cu[i].calcresidual[1]);
        SLOT(cu[i].cpy2Dto1D_shl); SLOT(cu[i].cpy2Dto1D_shr); SLOT(cu[i].cpy1Dto2D_shl[0]); SLOT(cu[i].cpy1Dto2D_shl[1]); SLOT(cu[i].cpy1Dto2D_shr);
        for (int c = 1; c < X265_CSP_COUNT; c++) { SLOT(chroma[c].cu[i].sub_ps); SLOT(chroma[c].cu[i].add_ps[0]); SLOT(chroma[c].cu[i].add_ps[1]); }
#if HIGH_BIT_DEPTH
        aliasOK = true;     /* copy_ps/sp/ss (and the cu copy_pp alias) are shared trampolines onto pu[].copy_pp at HBD (primitives.cpp:106-168) */
#endif
        SLOT(cu[i].copy_ss); SLOT(cu[i].copy_sp); SLOT(cu[i].copy_ps); SLOT(cu[i].copy_pp);
        for (int c = 1; c < X265_CSP_COUNT; c++) { SLOT(chroma[c].cu[i].copy_ss); SLOT(chroma[c].cu[i].copy_sp); SLOT(chroma[c].cu[i].copy_ps); SLOT(chroma[c].cu[i].copy_pp); }
        aliasOK = false;
    }
    for (int i = 0; i < NUM_CU_SIZES; i++)
    {
        SLOT(cu[i].intra_filter); SLOT(cu[i].intra_pred_allangs);
        for (int m = 0; m < NUM_INTRA_MODE; m++) SLOT(cu[i].intra_pred[m]);
    }
    SLOT(frameInitLowres); SLOT(weight_pp); SLOT(weight_sp); SLOT(denoiseDct);
    for (int i = 0; i < NUM_INTEGRAL_SIZE; i++) { SLOT(integral_inith[i]); SLOT(integral_initv[i]); }
    printf("adjacent: %d slots in the C table, %d bound to B200 entries, %d missing, %d still C, %d extra\n", want, bound, missing, same, extra);
    return hot + missing + same + extra;
}

int main(int argc, char** argv)
{
    unsigned seed = argc > 1 ? (unsigned)strtoul(argv[1], NULL, 0) : 0x265;
    const char* only = argc > 2 ? argv[2] : NULL;
    printf("x265 TestBench harness vs B200 table: %d bit, seed %X\n", X265_DEPTH, seed);
    srand(seed);

    /* C table first: setupCPrimitives rebinds lowpassdct.cpp's static slot pointers */
    memset(&cprim, 0, sizeof(cprim));
    setupCPrimitives(cprim);
    setupAliasPrimitives(cprim);

    char self[4096];
    ssize_t len = readlink("/proc/self/exe", self, sizeof(self) - 1);
    if (len <= 0) return 3;
    self[len] = 0;
    char glue[4200];
    snprintf(glue, sizeof(glue), "%s/../../x265-mod-by-patman_b200/lib/libx265b200_glue_%d.so", dirname(self), X265_DEPTH);
    void* h = dlopen(glue, RTLD_NOW | RTLD_GLOBAL);
    if (!h) { fprintf(stderr, "cannot load %s: %s\n", glue, dlerror()); return 3; }
    typedef int (*setup_t)(void*, int);
    setup_t setup = (setup_t)dlsym(h, "x265b200_setup_primitives");
    if (!setup) { fprintf(stderr, "x265b200_setup_primitives not found\n"); return 3; }

    memset(&b200prim, 0, sizeof(b200prim));
    int r = setup(&b200prim, 0);
    if (r) { fprintf(stderr, "x265b200_setup_primitives failed: %d (no CUDA device? there is no CPU fallback)\n", r); return 4; }
    setupAliasPrimitives(b200prim);
    memcpy(&primitives, &b200prim, sizeof(b200prim));          /* testbench.cpp:246 */

    if (coverage()) { fprintf(stderr, "coverage contract violated\n"); return 2; }

    TestHarness* harness[] = { &HPixel, &HMBDist, &HIPFilter, &HIPred };
    for (size_t i = 0; i < sizeof(harness) / sizeof(harness[0]); i++)
    {
        if (only && strncmp(only, harness[i]->getName(), strlen(only))) continue;
        printf("testCorrectness: %s ...\n", harness[i]->getName());
        fflush(stdout);
        if (!harness[i]->testCorrectness(cprim, b200prim))
        {
            fflush(stdout);
            fprintf(stderr, "\nB200 primitive has failed in harness '%s'\n", harness[i]->getName());
            return 1;
        }
        printf("testCorrectness: %s PASSED\n", harness[i]->getName());
    }
    typedef int (*status_t)(const void*);
    typedef const void* (*ctx_t)(void);
    status_t status = (status_t)dlsym(h, "x265b200_status");
    ctx_t ctx = (ctx_t)dlsym(h, "x265b200_glue_context");
    typedef unsigned long long (*lc_t)(const void*);
    lc_t lc = (lc_t)dlsym(h, "x265b200_launch_count");
    if (status && ctx && status(ctx())) { fprintf(stderr, "sticky CUDA error %d\n", status(ctx())); return 5; }
    printf("ALL PASSED (%llu CUDA kernel launches)\n", lc && ctx ? lc(ctx()) : 0ULL);
    return 0;
}
