/*
 * x265_oracle.c -- CPU restatement of the x265 analysis-primitive hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * leg may load it.  The product (x265-mod-by-patman_b200/) never links or calls it.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit against
 *   (1) the reference's own C primitives compiled from /root/reference into
 *       oracle/_ref/libx265ref_<depth>.so (tests/test_oracle_vs_ref.py, runs wherever
 *       oracle/_ref exists), and
 *   (2) golden vectors generated from that library and committed under
 *       tests/golden/ (tests/test_oracle_golden.py, runs everywhere).
 * The reference ships no golden vectors of its own.
 *
 * Written from the algorithm descriptions, in a deliberately different code shape
 * from the reference (plain int32 Hadamard instead of SWAR lane packing, full-matrix
 * two-stage transforms instead of partial butterflies, generated coefficient
 * tables).  Each function cites the reference file:line whose behaviour it restates
 * (paths relative to /root/reference/source/common).
 *
 * Build: one shared object per bit depth, -DX265_DEPTH={8,10,12}  (oracle/Makefile).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef X265_DEPTH
#error "compile with -DX265_DEPTH=8|10|12"
#endif

#if X265_DEPTH == 8
typedef uint8_t pixel;          /* common.h:127-143 */
typedef uint32_t sse_t;         /* common.h:145-149 */
#else
typedef uint16_t pixel;
typedef uint64_t sse_t;
#endif

#define PIXEL_MAX ((1 << X265_DEPTH) - 1)
#define FENC_STRIDE 64          /* common.h:71 */
#define IF_INTERNAL_PREC 14     /* constants.h:66-70 */
#define IF_FILTER_PREC 6
#define IF_INTERNAL_OFFS (1 << (IF_INTERNAL_PREC - 1))

#define EXPORT __attribute__((visibility("default")))

static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
/* two's-complement wrapping helpers: the reference relies on what gcc does for int
 * overflow / shifts of negatives; state it explicitly here. */
static inline int32_t wrap_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static inline int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wrap_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static inline int32_t wrap_shl(int32_t a, int s) { return (int32_t)((uint32_t)a << s); }

EXPORT int orc_depth(void) { return X265_DEPTH; }
EXPORT int orc_pixel_bytes(void) { return (int)sizeof(pixel); }
EXPORT int orc_sse_bytes(void) { return (int)sizeof(sse_t); }

/* ------------------------------------------------------------------ constants */

/* HEVC core transform: T32[k][n] = C[(k*(2n+1)) mod 128] where C is the integer
 * "cosine" with the 32 unique magnitudes below and the symmetries of cos(m*pi/64).
 * TN[k][n] = T32[k*(32/N)][n].  Equals g_t4/g_t8/g_t16/g_t32, constants.cpp:270-344
 * (checked element-wise against the reference in tests/test_oracle_vs_ref.py). */
static const int16_t k_cosmag[32] = {
    64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67,
    64, 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4 };

static int16_t g_T32[32][32];
static int g_tables_ready;

static int cos128(int m)
{
    m &= 127;
    if (m > 64) m = 128 - m;        /* cos(2pi - x) = cos x */
    if (m == 32) return 0;
    if (m > 32) return -(int)k_cosmag[64 - m]; /* cos(pi - x) = -cos x; m == 64 -> -64 */
    return k_cosmag[m];
}

static void init_tables(void)
{
    if (g_tables_ready) return;
    for (int k = 0; k < 32; k++)
        for (int n = 0; n < 32; n++)
            g_T32[k][n] = (int16_t)cos128(k * (2 * n + 1));
    g_tables_ready = 1;
}

static inline int tcoef(int N, int k, int n) { return g_T32[k * (32 / N)][n]; }

/* 4x4 DST-VII matrix; equivalent to fastForwardDst / inversedst, dct.cpp:43-81 */
static const int16_t k_dst4[4][4] = {
    { 29, 55, 74, 84 }, { 74, 74, 0, -74 }, { 84, -29, -74, 55 }, { 55, -84, 74, -29 } };

/* constants.cpp:250-268 (HEVC spec tables 8-11 / 8-12) */
static const int16_t k_luma_taps[4][8] = {
    { 0, 0, 0, 64, 0, 0, 0, 0 }, { -1, 4, -10, 58, 17, -5, 1, 0 },
    { -1, 4, -11, 40, 40, -11, 4, -1 }, { 0, 1, -5, 17, 58, -10, 4, -1 } };
static const int16_t k_chroma_taps[8][4] = {
    { 0, 64, 0, 0 }, { -2, 58, 10, -2 }, { -4, 54, 16, -2 }, { -6, 46, 28, -4 },
    { -4, 36, 36, -4 }, { -4, 28, 46, -6 }, { -2, 16, 54, -4 }, { -2, 10, 58, -2 } };

EXPORT void orc_get_dct_matrix(int N, int16_t* out /* N*N */)
{
    init_tables();
    for (int k = 0; k < N; k++)
        for (int n = 0; n < N; n++)
            out[k * N + n] = (int16_t)tcoef(N, k, n);
}
EXPORT void orc_get_dst_matrix(int16_t* out) { memcpy(out, k_dst4, sizeof(k_dst4)); }
EXPORT void orc_get_luma_taps(int16_t* out) { memcpy(out, k_luma_taps, sizeof(k_luma_taps)); }
EXPORT void orc_get_chroma_taps(int16_t* out) { memcpy(out, k_chroma_taps, sizeof(k_chroma_taps)); }

/* ------------------------------------------------------------------ pixel metrics */

/* pixel.cpp:40-55  sad<lx,ly> */
EXPORT int orc_sad(int w, int h, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    int sum = 0;
    for (int y = 0; y < h; y++, a += sa, b += sb)
        for (int x = 0; x < w; x++)
            sum += iabs((int)a[x] - (int)b[x]);
    return sum;
}

/* pixel.cpp:74-119  sad_x3 / sad_x4: fenc stride is the constant FENC_STRIDE */
EXPORT void orc_sad_xn(int w, int h, int n, const pixel* fenc, const pixel* const* refs,
                       intptr_t frefstride, int32_t* res)
{
    for (int i = 0; i < n; i++)
        res[i] = orc_sad(w, h, fenc, FENC_STRIDE, refs[i], frefstride);
}
EXPORT void orc_sad_x3(int w, int h, const pixel* fenc, const pixel* r0, const pixel* r1,
                       const pixel* r2, intptr_t frefstride, int32_t* res)
{
    const pixel* r[3] = { r0, r1, r2 };
    orc_sad_xn(w, h, 3, fenc, r, frefstride, res);
}
EXPORT void orc_sad_x4(int w, int h, const pixel* fenc, const pixel* r0, const pixel* r1,
                       const pixel* r2, const pixel* r3, intptr_t frefstride, int32_t* res)
{
    const pixel* r[4] = { r0, r1, r2, r3 };
    orc_sad_xn(w, h, 4, fenc, r, frefstride, res);
}

/* pixel.cpp:121-165 ads_x4/x2/x1 and the per-shape choice, pixel.cpp:1122-1146.
 * Returns which form (4, 2 or 1 DC terms) the reference binds to a w x h PU. */
EXPORT int orc_ads_terms(int w, int h)
{
    static const struct { int w, h, k; } tab[] = {
        {4,4,1},{8,8,1},{8,4,2},{4,8,2},{16,16,4},{16,8,2},{8,16,2},{16,12,1},{12,16,1},
        {16,4,1},{4,16,1},{32,32,4},{32,16,2},{16,32,2},{32,24,4},{24,32,4},{32,8,4},{8,32,4},
        {64,64,4},{64,32,2},{32,64,2},{64,48,4},{48,64,4},{64,16,4},{16,64,4} };
    for (unsigned i = 0; i < sizeof(tab) / sizeof(tab[0]); i++)
        if (tab[i].w == w && tab[i].h == h) return tab[i].k;
    return 0;
}

EXPORT int orc_ads(int w, int h, const int* encDC, const uint32_t* sums, int delta,
                   const uint16_t* costMvX, int16_t* mvs, int width, int thresh)
{
    int k = orc_ads_terms(w, h);
    int nmv = 0;
    int half = w >> 1;
    /* the reference loop index is an int16_t (pixel.cpp:125); width <= 32767 in any caller */
    for (int i = 0; i < width; i++)
    {
        const uint32_t* s = sums + i;
        long ads;
        if (k == 4)
            ads = labs((long)encDC[0] - (long)s[0]) + labs((long)encDC[1] - (long)s[half])
                + labs((long)encDC[2] - (long)s[delta]) + labs((long)encDC[3] - (long)s[delta + half]);
        else if (k == 2)
            ads = labs((long)encDC[0] - (long)s[0]) + labs((long)encDC[1] - (long)s[delta]);
        else
            ads = labs((long)encDC[0] - (long)s[0]);
        /* reference: int ads = <long sum> + costMvX[i]; -> truncation to int */
        int adsi = (int)(ads + costMvX[i]);
        if (adsi < thresh)
            mvs[nmv++] = (int16_t)i;
    }
    return nmv;
}

/* pixel.cpp:167-186  sse<lx,ly,T1,T2>: squares in int, accumulates in sse_t */
EXPORT sse_t orc_sse_pp(int w, int h, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    sse_t sum = 0;
    for (int y = 0; y < h; y++, a += sa, b += sb)
        for (int x = 0; x < w; x++)
        {
            int t = (int)a[x] - (int)b[x];
            sum += (sse_t)(t * t);
        }
    return sum;
}
EXPORT sse_t orc_sse_ss(int w, int h, const int16_t* a, intptr_t sa, const int16_t* b, intptr_t sb)
{
    sse_t sum = 0;
    for (int y = 0; y < h; y++, a += sa, b += sb)
        for (int x = 0; x < w; x++)
        {
            int t = (int)a[x] - (int)b[x];
            /* (tmp * tmp) is an int product in the reference: wraps beyond |t| > 46340,
             * then converts (sign-extending) to sse_t */
            sum += (sse_t)wrap_mul(t, t);
        }
    return sum;
}
/* pixel.cpp:371-383 pixel_ssd_s_c<size> */
EXPORT sse_t orc_ssd_s(int size, const int16_t* a, intptr_t stride)
{
    sse_t sum = 0;
    for (int y = 0; y < size; y++, a += stride)
        for (int x = 0; x < size; x++)
            sum += (sse_t)((int)a[x] * (int)a[x]);
    return sum;
}

/* 4x4 Hadamard SATD tile, raw (un-halved) sum.  pixel.cpp:210-261: satd_4x4 returns
 * raw>>1 and satd_8x4 returns (raw_left + raw_right)>>1; every raw tile sum is even,
 * so a per-tile >>1 is identical. */
static int hadamard4x4_abs(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    int d[4][4], t[4][4];
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++)
            d[y][x] = (int)a[y * sa + x] - (int)b[y * sb + x];
    for (int y = 0; y < 4; y++)
    {
        int s0 = d[y][0] + d[y][1], s1 = d[y][0] - d[y][1];
        int s2 = d[y][2] + d[y][3], s3 = d[y][2] - d[y][3];
        t[y][0] = s0 + s2; t[y][1] = s1 + s3; t[y][2] = s0 - s2; t[y][3] = s1 - s3;
    }
    int sum = 0;
    for (int x = 0; x < 4; x++)
    {
        int s0 = t[0][x] + t[1][x], s1 = t[0][x] - t[1][x];
        int s2 = t[2][x] + t[3][x], s3 = t[2][x] - t[3][x];
        sum += iabs(s0 + s2) + iabs(s1 + s3) + iabs(s0 - s2) + iabs(s1 - s3);
    }
    return sum;
}

/* pixel.cpp:263-289 satd4<w,h> / satd8<w,h>; slot binding pixel.cpp:1148-1172 */
EXPORT int orc_satd(int w, int h, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    int satd = 0;
    for (int y = 0; y < h; y += 4)
        for (int x = 0; x < w; x += 4)
            satd += hadamard4x4_abs(a + y * sa + x, sa, b + y * sb + x, sb) >> 1;
    return satd;
}

/* pixel.cpp:291-334 _sa8d_8x8: raw sum of |H8 * D * H8^T| */
static int hadamard8x8_abs(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    int m[8][8];
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            m[y][x] = (int)a[y * sa + x] - (int)b[y * sb + x];
    /* rows then columns, radix-2 */
    for (int y = 0; y < 8; y++)
        for (int step = 1; step < 8; step <<= 1)
            for (int i = 0; i < 8; i += step << 1)
                for (int j = i; j < i + step; j++)
                {
                    int u = m[y][j], v = m[y][j + step];
                    m[y][j] = u + v; m[y][j + step] = u - v;
                }
    for (int x = 0; x < 8; x++)
        for (int step = 1; step < 8; step <<= 1)
            for (int i = 0; i < 8; i += step << 1)
                for (int j = i; j < i + step; j++)
                {
                    int u = m[j][x], v = m[j + step][x];
                    m[j][x] = u + v; m[j + step][x] = u - v;
                }
    int sum = 0;
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            sum += iabs(m[y][x]);
    return sum;
}

/* pixel.cpp:336-369: sa8d8<w,h> = sum of (raw8x8+2)>>2; sa8d16<w,h> = sum over 16x16 of
 * (four raw 8x8 + 2)>>2.  Slot binding: pixel.cpp:1180-1184,1260-1263,1339-1342:
 * both dims multiple of 16 -> sa8d16 form; both multiple of 8 -> sa8d8 form; else satd. */
EXPORT int orc_sa8d(int w, int h, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    int cost = 0;
    if (w % 16 == 0 && h % 16 == 0)
    {
        for (int y = 0; y < h; y += 16)
            for (int x = 0; x < w; x += 16)
            {
                int raw = 0;
                for (int q = 0; q < 4; q++)
                {
                    int ox = x + (q & 1) * 8, oy = y + (q >> 1) * 8;
                    raw += hadamard8x8_abs(a + oy * sa + ox, sa, b + oy * sb + ox, sb);
                }
                cost += (raw + 2) >> 2;
            }
    }
    else if (w % 8 == 0 && h % 8 == 0)
    {
        for (int y = 0; y < h; y += 8)
            for (int x = 0; x < w; x += 8)
                cost += (hadamard8x8_abs(a + y * sa + x, sa, b + y * sb + x, sb) + 2) >> 2;
    }
    else
        cost = orc_satd(w, h, a, sa, b, sb);
    return cost;
}

/* ------------------------------------------------------------------ transforms */

/* Forward stage in matrix form (restates partialButterflyN,
 * dct.cpp:83-240,418-440): dst[k*n + j] = (int16)((sum_i M[k][i]*src[j*n + i] + add) >> shift) */
static void fwd_stage(int n, int dst4, const int16_t* src, int16_t* dst, int shift)
{
    int add = 1 << (shift - 1);
    for (int j = 0; j < n; j++)
        for (int k = 0; k < n; k++)
        {
            int acc = 0;
            for (int i = 0; i < n; i++)
                acc += (dst4 ? k_dst4[k][i] : tcoef(n, k, i)) * (int)src[j * n + i];
            dst[k * n + j] = (int16_t)((acc + add) >> shift);  /* truncation, dct.cpp:113 */
        }
}

/* Inverse stage (restates partialButterflyInverseN, dct.cpp:242-416):
 * dst[j*n + i] = clip16((sum_k M[k][i]*src[k*n + j] + add) >> shift) */
static void inv_stage(int n, int dst4, const int16_t* src, int16_t* dst, int shift)
{
    int add = 1 << (shift - 1);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++)
        {
            int acc = 0;
            for (int k = 0; k < n; k++)
                acc += (dst4 ? k_dst4[k][i] : tcoef(n, k, i)) * (int)src[k * n + j];
            dst[j * n + i] = (int16_t)clip3(-32768, 32767, (acc + add) >> shift); /* dct.cpp:257 */
        }
}

static int ilog2(int n) { int l = 0; while ((1 << l) < n) l++; return l; }

/* dct.cpp:443-526 dst4_c / dct4_c .. dct32_c */
static void fwd_transform(int n, int dst4, const int16_t* src, int16_t* dst, intptr_t srcStride)
{
    int16_t block[32 * 32], coef[32 * 32];
    init_tables();
    int l2 = ilog2(n);
    for (int y = 0; y < n; y++)
        memcpy(block + y * n, src + y * srcStride, (size_t)n * sizeof(int16_t));
    fwd_stage(n, dst4, block, coef, l2 - 1 + (X265_DEPTH - 8));
    fwd_stage(n, dst4, coef, dst, l2 + 6);
}
/* dct.cpp:528-611 idst4_c / idct4_c .. idct32_c */
static void inv_transform(int n, int dst4, const int16_t* src, int16_t* dst, intptr_t dstStride)
{
    int16_t block[32 * 32], coef[32 * 32];
    init_tables();
    inv_stage(n, dst4, src, coef, 7);
    inv_stage(n, dst4, coef, block, 12 - (X265_DEPTH - 8));
    for (int y = 0; y < n; y++)
        memcpy(dst + y * dstStride, block + y * n, (size_t)n * sizeof(int16_t));
}

EXPORT void orc_dct(int n, const int16_t* src, int16_t* dst, intptr_t srcStride) { fwd_transform(n, 0, src, dst, srcStride); }
EXPORT void orc_idct(int n, const int16_t* src, int16_t* dst, intptr_t dstStride) { inv_transform(n, 0, src, dst, dstStride); }
EXPORT void orc_dst4(const int16_t* src, int16_t* dst, intptr_t srcStride) { fwd_transform(4, 1, src, dst, srcStride); }
EXPORT void orc_idst4(const int16_t* src, int16_t* dst, intptr_t dstStride) { inv_transform(4, 1, src, dst, dstStride); }

/* lowpassdct.cpp:34-116: 2x2-average the block, run the (n/2)-point DCT, place it in
 * the top-left quadrant, zero the rest, overwrite DC with the scaled block sum.
 * Note the int16 truncation of every 2x2 sum, and of the running total for n == 8. */
EXPORT void orc_lowpass_dct(int n, const int16_t* src, int16_t* dst, intptr_t srcStride)
{
    int hn = n / 2;
    int16_t avg[16 * 16], coef[16 * 16];
    int32_t total32 = 0;
    int16_t total16 = 0;
    for (int i = 0; i < hn; i++)
        for (int j = 0; j < hn; j++)
        {
            int16_t sum = (int16_t)(src[2 * i * srcStride + 2 * j] + src[2 * i * srcStride + 2 * j + 1]
                                  + src[(2 * i + 1) * srcStride + 2 * j] + src[(2 * i + 1) * srcStride + 2 * j + 1]);
            avg[i * hn + j] = (int16_t)(sum >> 2);
            total32 += sum;
            total16 = (int16_t)(total16 + sum);
        }
    orc_dct(hn, avg, coef, hn);
    memset(dst, 0, (size_t)n * n * sizeof(int16_t));
    for (int i = 0; i < hn; i++)
        memcpy(dst + i * n, coef + i * hn, (size_t)hn * sizeof(int16_t));
    if (n == 8)
    {
#if X265_DEPTH == 8
        dst[0] = (int16_t)wrap_shl(total16, 1);
#else
        dst[0] = (int16_t)(total16 >> (X265_DEPTH - 9));
#endif
    }
    else if (n == 16)
        dst[0] = (int16_t)(total32 >> (1 + (X265_DEPTH - 8)));
    else
        dst[0] = (int16_t)(total32 >> (3 + (X265_DEPTH - 8)));
}

/* dct.cpp:666-688 quant_c.  int32 products wrap (TestBench feeds negative quantCoeff). */
EXPORT uint32_t orc_quant(const int16_t* coef, const int32_t* quantCoeff, int32_t* deltaU,
                          int16_t* qCoef, int qBits, int add, int numCoeff)
{
    int qBits8 = qBits - 8;
    uint32_t numSig = 0;
    for (int i = 0; i < numCoeff; i++)
    {
        int level = coef[i];
        int sign = level < 0 ? -1 : 1;
        int32_t tmplevel = wrap_mul(iabs(level), quantCoeff[i]);
        level = wrap_add(tmplevel, add) >> qBits;
        deltaU[i] = wrap_sub(tmplevel, wrap_shl(level, qBits)) >> qBits8;
        if (level) ++numSig;
        level = wrap_mul(level, sign);
        qCoef[i] = (int16_t)clip3(-32768, 32767, level);
    }
    return numSig;
}

/* dct.cpp:690-715 nquant_c: stores abs(clip16(level)); -32768 stays -32768 after the cast */
EXPORT uint32_t orc_nquant(const int16_t* coef, const int32_t* quantCoeff, int16_t* qCoef,
                           int qBits, int add, int numCoeff)
{
    uint32_t numSig = 0;
    for (int i = 0; i < numCoeff; i++)
    {
        int level = coef[i];
        int sign = level < 0 ? -1 : 1;
        int32_t tmplevel = wrap_mul(iabs(level), quantCoeff[i]);
        level = wrap_add(tmplevel, add) >> qBits;
        if (level) ++numSig;
        level = wrap_mul(level, sign);
        qCoef[i] = (int16_t)iabs(clip3(-32768, 32767, level));
    }
    return numSig;
}

/* dct.cpp:614-636 dequant_normal_c */
EXPORT void orc_dequant_normal(const int16_t* quantCoef, int16_t* coef, int num, int scale, int shift)
{
    int add = 1 << (shift - 1);
    for (int n = 0; n < num; n++)
    {
        int q = wrap_add(wrap_mul(quantCoef[n], scale), add) >> shift;
        coef[n] = (int16_t)clip3(-32768, 32767, q);
    }
}

/* dct.cpp:638-664 dequant_scaling_c: two regimes on (shift + 4) > per */
EXPORT void orc_dequant_scaling(const int16_t* quantCoef, const int32_t* deQuantCoef, int16_t* coef,
                                int num, int per, int shift)
{
    shift += 4;
    if (shift > per)
    {
        int add = 1 << (shift - per - 1);
        for (int n = 0; n < num; n++)
        {
            int q = wrap_add(wrap_mul(quantCoef[n], deQuantCoef[n]), add) >> (shift - per);
            coef[n] = (int16_t)clip3(-32768, 32767, q);
        }
    }
    else
    {
        for (int n = 0; n < num; n++)
        {
            int q = clip3(-32768, 32767, wrap_mul(quantCoef[n], deQuantCoef[n]));
            coef[n] = (int16_t)clip3(-32768, 32767, wrap_mul(q, 1 << (per - shift)));
        }
    }
}

/* ------------------------------------------------------------------ interpolation */

static inline const int16_t* taps(int N, int idx) { return N == 8 ? k_luma_taps[idx] : k_chroma_taps[idx]; }

/* ipfilter.cpp:40-57 filterPixelToShort_c */
EXPORT void orc_p2s(int w, int h, const pixel* src, intptr_t srcStride, int16_t* dst, intptr_t dstStride)
{
    int shift = IF_INTERNAL_PREC - X265_DEPTH;
    for (int y = 0; y < h; y++, src += srcStride, dst += dstStride)
        for (int x = 0; x < w; x++)
        {
            int16_t val = (int16_t)((int)src[x] << shift);
            dst[x] = (int16_t)(val - (int16_t)IF_INTERNAL_OFFS);
        }
}

/* ipfilter.cpp:79-118 interp_horiz_pp_c / :164-203 interp_vert_pp_c.
 * step = 1 for horizontal, srcStride for vertical. */
static void filt_pp(int N, int w, int h, const pixel* src, intptr_t srcStride, intptr_t step,
                    pixel* dst, intptr_t dstStride, int idx)
{
    const int16_t* c = taps(N, idx);
    src -= (N / 2 - 1) * step;
    for (int y = 0; y < h; y++, src += srcStride, dst += dstStride)
        for (int x = 0; x < w; x++)
        {
            int sum = 0;
            for (int t = 0; t < N; t++) sum += (int)src[x + t * step] * c[t];
            int16_t val = (int16_t)((sum + 32) >> IF_FILTER_PREC);   /* cast before clip */
            if (val < 0) val = 0;
            if (val > PIXEL_MAX) val = PIXEL_MAX;
            dst[x] = (pixel)val;
        }
}
EXPORT void orc_interp_hpp(int N, int w, int h, const pixel* src, intptr_t ss, pixel* dst, intptr_t ds, int idx)
{ filt_pp(N, w, h, src, ss, 1, dst, ds, idx); }
EXPORT void orc_interp_vpp(int N, int w, int h, const pixel* src, intptr_t ss, pixel* dst, intptr_t ds, int idx)
{ filt_pp(N, w, h, src, ss, ss, dst, ds, idx); }

/* ipfilter.cpp:120-162 interp_horiz_ps_c (isRowExt adds N/2-1 rows above, N/2 below)
 * ipfilter.cpp:205-238 interp_vert_ps_c */
EXPORT void orc_interp_hps(int N, int w, int h, const pixel* src, intptr_t ss, int16_t* dst, intptr_t ds,
                           int idx, int isRowExt)
{
    const int16_t* c = taps(N, idx);
    int shift = IF_FILTER_PREC - (IF_INTERNAL_PREC - X265_DEPTH);
    int offset = (int)((unsigned)-IF_INTERNAL_OFFS << shift);
    int rows = h;
    src -= N / 2 - 1;
    if (isRowExt) { src -= (N / 2 - 1) * ss; rows += N - 1; }
    for (int y = 0; y < rows; y++, src += ss, dst += ds)
        for (int x = 0; x < w; x++)
        {
            int sum = 0;
            for (int t = 0; t < N; t++) sum += (int)src[x + t] * c[t];
            dst[x] = (int16_t)((sum + offset) >> shift);
        }
}
EXPORT void orc_interp_vps(int N, int w, int h, const pixel* src, intptr_t ss, int16_t* dst, intptr_t ds, int idx)
{
    const int16_t* c = taps(N, idx);
    int shift = IF_FILTER_PREC - (IF_INTERNAL_PREC - X265_DEPTH);
    int offset = (int)((unsigned)-IF_INTERNAL_OFFS << shift);
    src -= (N / 2 - 1) * ss;
    for (int y = 0; y < h; y++, src += ss, dst += ds)
        for (int x = 0; x < w; x++)
        {
            int sum = 0;
            for (int t = 0; t < N; t++) sum += (int)src[x + t * ss] * c[t];
            dst[x] = (int16_t)((sum + offset) >> shift);
        }
}
/* ipfilter.cpp:240-283 interp_vert_sp_c */
EXPORT void orc_interp_vsp(int N, int w, int h, const int16_t* src, intptr_t ss, pixel* dst, intptr_t ds, int idx)
{
    const int16_t* c = taps(N, idx);
    int shift = IF_FILTER_PREC + (IF_INTERNAL_PREC - X265_DEPTH);
    int offset = (1 << (shift - 1)) + (IF_INTERNAL_OFFS << IF_FILTER_PREC);
    src -= (N / 2 - 1) * ss;
    for (int y = 0; y < h; y++, src += ss, dst += ds)
        for (int x = 0; x < w; x++)
        {
            int sum = 0;
            for (int t = 0; t < N; t++) sum += (int)src[x + t * ss] * c[t];
            int16_t val = (int16_t)((sum + offset) >> shift);
            if (val < 0) val = 0;
            if (val > PIXEL_MAX) val = PIXEL_MAX;
            dst[x] = (pixel)val;
        }
}
/* ipfilter.cpp:285-317 interp_vert_ss_c: >>6, no rounding, int16 wrap */
EXPORT void orc_interp_vss(int N, int w, int h, const int16_t* src, intptr_t ss, int16_t* dst, intptr_t ds, int idx)
{
    const int16_t* c = taps(N, idx);
    src -= (N / 2 - 1) * ss;
    for (int y = 0; y < h; y++, src += ss, dst += ds)
        for (int x = 0; x < w; x++)
        {
            int sum = 0;
            for (int t = 0; t < N; t++) sum += (int)src[x + t * ss] * c[t];
            dst[x] = (int16_t)(sum >> IF_FILTER_PREC);
        }
}
/* ipfilter.cpp:362-369 interp_hv_pp_c: hps(isRowExt=1) into a w-stride scratch, then vertical sp */
EXPORT void orc_interp_hvpp(int N, int w, int h, const pixel* src, intptr_t ss, pixel* dst, intptr_t ds,
                            int idxX, int idxY)
{
    int16_t* immed = (int16_t*)malloc((size_t)w * (h + N - 1) * sizeof(int16_t));
    orc_interp_hps(N, w, h, src, ss, immed, w, idxX, 1);
    orc_interp_vsp(N, w, h, immed + (N / 2 - 1) * w, w, dst, ds, idxY);
    free(immed);
}

/* ------------------------------------------------------------------ batched drivers
 * Descriptor-array loops used to diff whole-frame GPU output element-wise and as the
 * "port" CPU baseline.  Offsets are element offsets from the plane base pointers. */

enum { ORC_SAD = 0, ORC_SATD = 1, ORC_SA8D = 2, ORC_SSE_PP = 3 };

EXPORT void orc_pixelcmp_batch(int op, int w, int h, const pixel* A, intptr_t sa, const pixel* B, intptr_t sb,
                               const int32_t* offA, const int32_t* offB, int n, void* out)
{
    for (int i = 0; i < n; i++)
    {
        const pixel* a = A + offA[i];
        const pixel* b = B + offB[i];
        switch (op)
        {
        case ORC_SAD:  ((int32_t*)out)[i] = orc_sad(w, h, a, sa, b, sb); break;
        case ORC_SATD: ((int32_t*)out)[i] = orc_satd(w, h, a, sa, b, sb); break;
        case ORC_SA8D: ((int32_t*)out)[i] = orc_sa8d(w, h, a, sa, b, sb); break;
        default:       ((uint64_t*)out)[i] = (uint64_t)orc_sse_pp(w, h, a, sa, b, sb); break;
        }
    }
}

/* residual = fenc - pred (pixel.cpp: sub_ps, adjacent slot) for building DCT inputs */
EXPORT void orc_residual_batch(int w, int h, const pixel* A, intptr_t sa, const pixel* B, intptr_t sb,
                               const int32_t* offA, const int32_t* offB, int n, int16_t* out /* n*w*h */)
{
    for (int i = 0; i < n; i++)
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                out[(size_t)i * w * h + y * w + x] = (int16_t)((int)A[offA[i] + y * sa + x] - (int)B[offB[i] + y * sb + x]);
}

/* extendPicBorder (common/pixel.cpp:1044-1061; the row part is extendCURowColBorder, pixel.cpp:1018-1038 without the
 * word-splat shortcuts): marginX copies of the first / last sample of every picture row to its left / right, then
 * marginY copies of the first and the last padded row above / below.  pic points at sample (0, 0). */
EXPORT void orc_extend_pic_border(pixel* pic, intptr_t stride, int width, int height, int marginX, int marginY)
{
    for (int y = 0; y < height; y++)
    {
        pixel* row = pic + (intptr_t)y * stride;
        for (int x = 0; x < marginX; x++)
        {
            row[-marginX + x] = row[0];
            row[width + x] = row[width - 1];
        }
    }
    /* the reference copies `stride` samples per row (pixel.cpp:1052, :1057), i.e. the whole buffer row, including the
     * columns between width + marginX and the row end when the picture is narrower than the CTU-aligned plane */
    for (int y = 1; y <= marginY; y++)
        for (int x = -marginX; x < (int)stride - marginX; x++)
        {
            pic[x - (intptr_t)y * stride] = pic[x];
            pic[x + (intptr_t)(height - 1 + y) * stride] = pic[x + (intptr_t)(height - 1) * stride];
        }
}

/* Inter luma TU reconstruction chain without RDOQ / psy / sign hiding / transform skip, scaling lists off:
 * encoder/search.cpp:5536-5575 (estimateResidualQT) -> common/quant.cpp:397-480 (transformNxN: dct, quant) and
 * quant.cpp:543-605 (invtransformNxN: dequant_normal, DC-only shortcut :588-598, idct), pixel.cpp sub_ps / add_ps
 * (:821-831) and sse_pp.  numSig == 0 leaves the prediction as reconstruction (cbf = 0).
 * Outputs per TU: qCoef[N*N], numSig, recon block, sse(fenc, pred), sse(fenc, recon). */
/* ttype 0: inter luma, and every chroma TU (the chroma loop of estimateResidualQT, search.cpp:5638-5700, runs the same calls on the chroma
 * planes with log2TrSizeC).  ttype 1: intra luma -- for N == 4 Quant::transformNxN / invtransformNxN take the DST-VII pair instead
 * (quant.cpp:430-441 "useDST", :585 `useDST = !sizeIdx && ttype == TEXT_LUMA && bIntra`, :600-603) and the DC-only shortcut is off (:588). */
EXPORT void orc_tu_chain_tt(int N, int ttype, const pixel* fenc, intptr_t sf, const pixel* pred, intptr_t sp,
                            const int32_t* quantCoeff, int qBits, int add, int dqScale, int dqShift,
                            int16_t* qCoef, uint32_t* numSig, pixel* recon, intptr_t sr, uint64_t* sseZero, uint64_t* sseRecon)
{
    const int useDST = ttype == 1 && N == 4;
    int16_t resi[32 * 32], coef[32 * 32], dq[32 * 32], rec[32 * 32];
    int32_t deltaU[32 * 32];
    int nn = N * N;
    for (int y = 0; y < N; y++)
        for (int x = 0; x < N; x++)
            resi[y * N + x] = (int16_t)((int)fenc[y * sf + x] - (int)pred[y * sp + x]);
    if (useDST) orc_dst4(resi, coef, N);
    else orc_dct(N, resi, coef, N);
    uint32_t ns = orc_quant(coef, quantCoeff, deltaU, qCoef, qBits, add, nn);
    *numSig = ns;
    *sseZero = (uint64_t)orc_sse_pp(N, N, fenc, sf, pred, sp);
    if (!ns)
    {
        for (int y = 0; y < N; y++)
            for (int x = 0; x < N; x++) recon[y * sr + x] = pred[y * sp + x];
        *sseRecon = *sseZero;
        return;
    }
    orc_dequant_normal(qCoef, dq, nn, dqScale, dqShift);
    if (ns == 1 && qCoef[0] != 0 && !useDST)
    {
        /* quant.cpp:588-598 */
        const int shift_1st = 7 - 6, add_1st = 1 << (shift_1st - 1);
        const int shift_2nd = 12 - (X265_DEPTH - 8) - 3, add_2nd = 1 << (shift_2nd - 1);
        int dc_val = (((dq[0] * (64 >> 6) + add_1st) >> shift_1st) * (64 >> 3) + add_2nd) >> shift_2nd;
        for (int i = 0; i < nn; i++) rec[i] = (int16_t)dc_val;
    }
    else if (useDST)
        orc_idst4(dq, rec, N);
    else
        orc_idct(N, dq, rec, N);
    for (int y = 0; y < N; y++)
        for (int x = 0; x < N; x++)
            recon[y * sr + x] = (pixel)clip3(0, PIXEL_MAX, (int)pred[y * sp + x] + rec[y * N + x]);   /* add_ps, pixel.cpp:821-831 */
    *sseRecon = (uint64_t)orc_sse_pp(N, N, fenc, sf, recon, sr);
}
EXPORT void orc_tu_chain(int N, const pixel* fenc, intptr_t sf, const pixel* pred, intptr_t sp,
                         const int32_t* quantCoeff, int qBits, int add, int dqScale, int dqShift,
                         int16_t* qCoef, uint32_t* numSig, pixel* recon, intptr_t sr, uint64_t* sseZero, uint64_t* sseRecon)
{
    orc_tu_chain_tt(N, 0, fenc, sf, pred, sp, quantCoeff, qBits, add, dqScale, dqShift, qCoef, numSig, recon, sr, sseZero, sseRecon);
}

EXPORT void orc_tu_chain_batch(int N, const pixel* fenc, intptr_t sf, const pixel* pred, intptr_t sp, const int32_t* offF,
                               const int32_t* offP, int n, const int32_t* quantCoeff, int qBits, int add, int dqScale, int dqShift,
                               int16_t* qCoef, uint32_t* numSig, pixel* recon, intptr_t sr, const int32_t* offR,
                               uint64_t* sseZero, uint64_t* sseRecon)
{
    for (int i = 0; i < n; i++)
        orc_tu_chain(N, fenc + offF[i], sf, pred + offP[i], sp, quantCoeff, qBits, add, dqScale, dqShift,
                     qCoef + (size_t)i * N * N, numSig + i, recon + offR[i], sr, sseZero + i, sseRecon + i);
}
EXPORT void orc_tu_chain_tt_batch(int N, int ttype, const pixel* fenc, intptr_t sf, const pixel* pred, intptr_t sp, const int32_t* offF,
                                  const int32_t* offP, int n, const int32_t* quantCoeff, int qBits, int add, int dqScale, int dqShift,
                                  int16_t* qCoef, uint32_t* numSig, pixel* recon, intptr_t sr, const int32_t* offR,
                                  uint64_t* sseZero, uint64_t* sseRecon)
{
    for (int i = 0; i < n; i++)
        orc_tu_chain_tt(N, ttype, fenc + offF[i], sf, pred + offP[i], sp, quantCoeff, qBits, add, dqScale, dqShift,
                        qCoef + (size_t)i * N * N, numSig + i, recon + offR[i], sr, sseZero + i, sseRecon + i);
}

/* dct over n strided blocks of one int16 plane; output contiguous n x (N*N) */
EXPORT void orc_dct_batch(int N, int dst4, const int16_t* src, intptr_t srcStride, const int32_t* off, int n, int16_t* out)
{
    for (int i = 0; i < n; i++)
        fwd_transform(N, dst4, src + off[i], out + (size_t)i * N * N, srcStride);
}
EXPORT void orc_idct_batch(int N, int dst4, const int16_t* src, int n, int16_t* dst, intptr_t dstStride, const int32_t* off)
{
    for (int i = 0; i < n; i++)
        inv_transform(N, dst4, src + (size_t)i * N * N, dst + off[i], dstStride);
}

/* ------------------------------------------------------------------------------------------------
 * Adjacent slots: two-input block operations and the lookahead's lowres downscale.
 * op: 0 sub_ps (pixel.cpp:806-818), 1 add_ps (pixel.cpp:820-832), 2 pixelavg_pp (pixel.cpp:537-549),
 *     3 addAvg (pixel.cpp:834-855).  Element types follow the op.
 * ------------------------------------------------------------------------------------------------ */
EXPORT void orc_blockop(int op, int w, int h, const void* A, intptr_t sa, const void* B, intptr_t sb, void* D, intptr_t sd)
{
    const int pmax = (1 << X265_DEPTH) - 1;
    const int shift = 14 + 1 - X265_DEPTH, offset = (1 << (shift - 1)) + 2 * 8192;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
        {
            if (op == 0)
                ((int16_t*)D)[y * sd + x] = (int16_t)((int)((const pixel*)A)[y * sa + x] - (int)((const pixel*)B)[y * sb + x]);
            else if (op == 1)
                ((pixel*)D)[y * sd + x] = (pixel)clip3(0, pmax, (int)((const pixel*)A)[y * sa + x] + (int)((const int16_t*)B)[y * sb + x]);
            else if (op == 2)
                ((pixel*)D)[y * sd + x] = (pixel)(((int)((const pixel*)A)[y * sa + x] + (int)((const pixel*)B)[y * sb + x] + 1) >> 1);
            else
                ((pixel*)D)[y * sd + x] = (pixel)clip3(0, pmax, ((int)((const int16_t*)A)[y * sa + x] + (int)((const int16_t*)B)[y * sb + x] + offset) >> shift);
        }
}

EXPORT void orc_blockop_batch(int op, int w, int h, const void* A, intptr_t sa, const int32_t* offA, const void* B, intptr_t sb,
                              const int32_t* offB, void* D, intptr_t sd, const int32_t* offD, int n)
{
    const size_t ea = op == 3 ? 2 : sizeof(pixel), eb = (op == 1 || op == 3) ? 2 : sizeof(pixel), ed = op == 0 ? 2 : sizeof(pixel);
    for (int i = 0; i < n; i++)
        orc_blockop(op, w, h, (const char*)A + (size_t)offA[i] * ea, sa, (const char*)B + (size_t)offB[i] * eb, sb, (char*)D + (size_t)offD[i] * ed, sd);
}

/* frame_init_lowres_core (pixel.cpp:595-620): rounding-average cascade, "slower than naive bilinear, but matches asm" */
static inline int avg2(int a, int b) { return (a + b + 1) >> 1; }
EXPORT void orc_lowres(const pixel* src, intptr_t ss, pixel* d0, pixel* dh, pixel* dv, pixel* dc, intptr_t ds, int width, int height)
{
    for (int y = 0; y < height; y++)
    {
        const pixel* r0 = src + (intptr_t)(2 * y) * ss;
        const pixel* r1 = r0 + ss;
        const pixel* r2 = r1 + ss;
        for (int x = 0; x < width; x++)
        {
            int c = 2 * x;
            d0[y * ds + x] = (pixel)avg2(avg2(r0[c], r1[c]), avg2(r0[c + 1], r1[c + 1]));
            dh[y * ds + x] = (pixel)avg2(avg2(r0[c + 1], r1[c + 1]), avg2(r0[c + 2], r1[c + 2]));
            dv[y * ds + x] = (pixel)avg2(avg2(r1[c], r2[c]), avg2(r1[c + 1], r2[c + 1]));
            dc[y * ds + x] = (pixel)avg2(avg2(r1[c + 1], r2[c + 1]), avg2(r1[c + 2], r2[c + 2]));
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Sub-pel candidate cost: MotionEstimate::subpelCompare, luma part (encoder/motion.cpp:1780-1821):
 * copy / luma_hpp / luma_vpp / luma_hvpp into a width-stride buffer, then sad (op 0) or satd (op 1) against fenc.
 * ------------------------------------------------------------------------------------------------ */
EXPORT int orc_subpel_cmp(int op, int w, int h, const pixel* fenc, intptr_t sf, const pixel* fref, intptr_t sr, int xFrac, int yFrac)
{
    if (!(xFrac | yFrac))
        return op ? orc_satd(w, h, fenc, sf, fref, sr) : orc_sad(w, h, fenc, sf, fref, sr);
    pixel buf[64 * 64];
    if (!yFrac) orc_interp_hpp(8, w, h, fref, sr, buf, w, xFrac);
    else if (!xFrac) orc_interp_vpp(8, w, h, fref, sr, buf, w, yFrac);
    else orc_interp_hvpp(8, w, h, fref, sr, buf, w, xFrac, yFrac);
    return op ? orc_satd(w, h, fenc, sf, buf, w) : orc_sad(w, h, fenc, sf, buf, w);
}

EXPORT void orc_subpel_cmp_batch(int op, int w, int h, const pixel* fenc, intptr_t sf, const pixel* ref, intptr_t sr,
                                 const int32_t* offF, const int32_t* offR, const int32_t* frac, int K, int n, int32_t* cost)
{
    for (int i = 0; i < n * K; i++)
        cost[i] = orc_subpel_cmp(op, w, h, fenc + offF[i / K], sf, ref + offR[i], sr, frac[i] & 3, (frac[i] >> 4) & 3);
}

/* ------------------------------------------------------------------------------------------------
 * Bi-prediction candidate cost (encoder/search.cpp:442-448 in Search::predInterSearch): both motion-compensated luma
 * blocks (Predict::predInterLumaPixel, common/predict.cpp:279-300: copy / luma_hpp / luma_vpp / luma_hvpp by the vector's
 * fraction), their rounded average (pixelavg_pp, pixel.cpp:537-549), SATD against fenc.
 * ------------------------------------------------------------------------------------------------ */
static void mc_luma(int w, int h, const pixel* fref, intptr_t sr, int xFrac, int yFrac, pixel* dst)
{
    if (!(xFrac | yFrac)) { for (int y = 0; y < h; y++) memcpy(dst + y * w, fref + y * sr, w * sizeof(pixel)); }
    else if (!yFrac) orc_interp_hpp(8, w, h, fref, sr, dst, w, xFrac);
    else if (!xFrac) orc_interp_vpp(8, w, h, fref, sr, dst, w, yFrac);
    else orc_interp_hvpp(8, w, h, fref, sr, dst, w, xFrac, yFrac);
}
EXPORT int orc_bidir_satd(int w, int h, const pixel* fenc, intptr_t sf, const pixel* ref0, intptr_t sr0, int frac0,
                          const pixel* ref1, intptr_t sr1, int frac1)
{
    pixel p0[64 * 64], p1[64 * 64];
    mc_luma(w, h, ref0, sr0, frac0 & 3, (frac0 >> 4) & 3, p0);
    mc_luma(w, h, ref1, sr1, frac1 & 3, (frac1 >> 4) & 3, p1);
    for (int i = 0; i < w * h; i++) p0[i] = (pixel)((p0[i] + p1[i] + 1) >> 1);
    return orc_satd(w, h, fenc, sf, p0, w);
}
EXPORT void orc_bidir_satd_batch(int w, int h, const pixel* fenc, intptr_t sf, const int32_t* offF, const pixel* ref0, intptr_t sr0,
                                 const int32_t* off0, const int32_t* frac0, const pixel* ref1, intptr_t sr1, const int32_t* off1,
                                 const int32_t* frac1, int n, int32_t* cost)
{
    for (int i = 0; i < n; i++)
        cost[i] = orc_bidir_satd(w, h, fenc + offF[i], sf, ref0 + off0[i], sr0, frac0[i], ref1 + off1[i], sr1, frac1[i]);
}

/* ------------------------------------------------------------------------------------------------
 * SEA integral planes (encoder/framefilter.cpp:38-140; row loop of FrameFilter::computeMEIntegral, :737-835).
 * ------------------------------------------------------------------------------------------------ */
EXPORT void orc_integral_inith(int W, uint32_t* sum, const pixel* pix, intptr_t stride)
{
    /* framefilter.cpp:39-103: running W-wide horizontal sum added to the row above */
    for (intptr_t x = 0; x < stride - W; x++)
    {
        uint32_t v = 0;
        for (int i = 0; i < W; i++) v += pix[x + i];
        sum[x] = v + sum[x - stride];
    }
}
EXPORT void orc_integral_initv(int H, uint32_t* sum, intptr_t stride)
{
    /* framefilter.cpp:106-140 */
    for (intptr_t x = 0; x < stride; x++) sum[x] = sum[x + H * stride] - sum[x];
}
static const int k_intW[12] = { 32, 32, 32, 24, 16, 16, 16, 12, 8, 8, 4, 4 };   /* framefilter.cpp:776-787 */
static const int k_intH[12] = { 32, 24, 8, 32, 16, 12, 4, 16, 32, 8, 16, 4 };
/* whole padded picture of `rows` rows: row t = y + padY of the picture feeds row t + 1 of every plane (framefilter.cpp:770-832) */
EXPORT void orc_me_integral(const pixel* pix, intptr_t stride, int rows, uint32_t* sums, size_t planePitch)
{
    for (int k = 0; k < 12; k++) memset(sums + k * planePitch, 0, (size_t)stride * sizeof(uint32_t));
    for (int t = 0; t < rows - 1; t++)
        for (int k = 0; k < 12; k++)
        {
            uint32_t* S = sums + k * planePitch;
            orc_integral_inith(k_intW[k], S + (intptr_t)(t + 1) * stride, pix + (intptr_t)t * stride, stride);
            if (t >= k_intH[k]) orc_integral_initv(k_intH[k], S + (intptr_t)(t + 1 - k_intH[k]) * stride, stride);
        }
}

/* ------------------------------------------------------------------------------------------------
 * Weighted prediction (pixel.cpp:485-535) and the lookahead's weighted-prediction cost
 * (slicetype.cpp:866-897 weightCostLuma / weightPrediction.cpp:171-222 luma branch).
 * ------------------------------------------------------------------------------------------------ */
EXPORT void orc_weight_pp(const pixel* src, pixel* dst, intptr_t stride, int width, int height, int w0, int round, int shift, int offset)
{
    const int correction = 14 - X265_DEPTH, pmax = (1 << X265_DEPTH) - 1;
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++)
        {
            int16_t val = (int16_t)(src[y * stride + x] << correction);
            dst[y * stride + x] = (pixel)clip3(0, pmax, ((w0 * val + round) >> shift) + offset);
        }
}
EXPORT void orc_weight_sp(const int16_t* src, pixel* dst, intptr_t ss, intptr_t ds, int width, int height, int w0, int round, int shift, int offset)
{
    const int pmax = (1 << X265_DEPTH) - 1;
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++)
            dst[y * ds + x] = (pixel)clip3(0, pmax, ((w0 * (src[y * ss + x] + 8192) + round) >> shift) + offset);
}
/* weights: K x {w0, round, shift, offset}, shift < 0 = unweighted; tmp: scratch plane like the reference's weighted plane */
EXPORT void orc_weight_cost(const pixel* fenc, const pixel* ref, intptr_t stride, int width, int height, const int32_t* intraCost,
                            const int32_t* weights, int K, uint32_t* cost, pixel* tmp)
{
    int pw = (width + 7) & ~7, ph = (height + 7) & ~7;
    for (int k = 0; k < K; k++)
    {
        const int32_t* w = weights + 4 * k;
        const pixel* src = ref;
        if (w[2] >= 0) { orc_weight_pp(ref, tmp, stride, pw, ph, w[0], w[1], w[2], w[3]); src = tmp; }
        uint32_t c = 0;
        int mb = 0;
        for (int y = 0; y < height; y += 8)
            for (int x = 0; x < width; x += 8, mb++)
            {
                int s = orc_satd(8, 8, src + y * stride + x, stride, fenc + y * stride + x, stride);
                c += (uint32_t)((intraCost && intraCost[mb] < s) ? intraCost[mb] : s);
            }
        cost[k] = c;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Copy family (pixel.cpp:385-461, 751-804).  kind: 0 copy_pp, 1 copy_ss, 2 copy_sp, 3 copy_ps, 4 blockfill_s (param = value),
 * 5 shift left (cpy2Dto1D_shl / cpy1Dto2D_shl), 6 rounding shift right (cpy2Dto1D_shr / cpy1Dto2D_shr).
 * ------------------------------------------------------------------------------------------------ */
EXPORT void orc_blockcopy(int kind, int w, int h, void* dst, intptr_t ds, const void* src, intptr_t ss, int param)
{
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            switch (kind)
            {
            case 0: ((pixel*)dst)[y * ds + x] = ((const pixel*)src)[y * ss + x]; break;
            case 1: ((int16_t*)dst)[y * ds + x] = ((const int16_t*)src)[y * ss + x]; break;
            case 2: ((pixel*)dst)[y * ds + x] = (pixel)((const int16_t*)src)[y * ss + x]; break;
            case 3: ((int16_t*)dst)[y * ds + x] = (int16_t)((const pixel*)src)[y * ss + x]; break;
            case 4: ((int16_t*)dst)[y * ds + x] = (int16_t)param; break;
            case 5: ((int16_t*)dst)[y * ds + x] = (int16_t)((uint32_t)((const int16_t*)src)[y * ss + x] << param); break;
            default: ((int16_t*)dst)[y * ds + x] = (int16_t)((((const int16_t*)src)[y * ss + x] + (int16_t)(1 << (param - 1))) >> param); break;
            }
}
EXPORT void orc_blockcopy_batch(int kind, int w, int h, const void* src, intptr_t ss, const int32_t* offS, void* dst, intptr_t ds,
                                const int32_t* offD, int n, int param)
{
    const size_t es = (kind == 0 || kind == 3) ? sizeof(pixel) : 2, ed = (kind == 0 || kind == 2) ? sizeof(pixel) : 2;
    for (int i = 0; i < n; i++)
        orc_blockcopy(kind, w, h, (char*)dst + (size_t)offD[i] * ed, ds, kind == 4 ? NULL : (const char*)src + (size_t)offS[i] * es, ss, param);
}

/* ------------------------------------------------------------------------------------------------
 * Per-block scalars: var (pixel.cpp:695-712), psy_cost_pp (pixel.cpp:718-749), count_nonzero / copy_cnt (dct.cpp:716-744),
 * denoiseDct (dct.cpp:746-757).
 * ------------------------------------------------------------------------------------------------ */
EXPORT uint64_t orc_var(int size, const pixel* pix, intptr_t stride)
{
    uint32_t sum = 0, sqr = 0;
    for (int y = 0; y < size; y++)
        for (int x = 0; x < size; x++) { uint32_t v = pix[y * stride + x]; sum += v; sqr += v * v; }
    return (uint64_t)sum + ((uint64_t)sqr << 32);
}
static int ac_energy(int n, const pixel* p, intptr_t stride)
{
    /* n x n Hadamard magnitude of the block itself (the reference differences it against a zero row), minus its DC */
    int m[8][8], t[8][8], sad = 0, raw = 0;
    for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) { m[y][x] = p[y * stride + x]; sad += m[y][x]; }
    for (int y = 0; y < n; y++)
        for (int k = 0; k < n; k++)
        {
            int acc = 0;
            for (int x = 0; x < n; x++) acc += (__builtin_popcount(k & x) & 1) ? -m[y][x] : m[y][x];
            t[y][k] = acc;
        }
    for (int k = 0; k < n; k++)
        for (int j = 0; j < n; j++)
        {
            int acc = 0;
            for (int y = 0; y < n; y++) acc += (__builtin_popcount(j & y) & 1) ? -t[y][k] : t[y][k];
            raw += iabs(acc);
        }
    return (n == 8 ? (raw + 2) >> 2 : raw >> 1) - (sad >> 2);
}
EXPORT int orc_psy_cost_pp(int size, const pixel* src, intptr_t ss, const pixel* rec, intptr_t sr)
{
    if (size == 4) return iabs(ac_energy(4, src, ss) - ac_energy(4, rec, sr));
    uint32_t tot = 0;
    for (int i = 0; i < size; i += 8)
        for (int j = 0; j < size; j += 8)
            tot += (uint32_t)iabs(ac_energy(8, src + i * ss + j, ss) - ac_energy(8, rec + i * sr + j, sr));
    return (int)tot;
}
EXPORT uint32_t orc_copy_cnt(int size, int16_t* coeff, const int16_t* resi, intptr_t stride)
{
    uint32_t n = 0;
    for (int k = 0; k < size; k++)
        for (int j = 0; j < size; j++)
        {
            if (coeff) coeff[k * size + j] = resi[k * stride + j];
            n += resi[k * stride + j] != 0;
        }
    return n;
}
EXPORT void orc_denoise_dct(int16_t* dct, uint32_t* resSum, const uint16_t* offset, int numCoeff)
{
    for (int i = 0; i < numCoeff; i++)
    {
        int level = dct[i];
        int sign = level >> 31;
        level = (level + sign) ^ sign;
        resSum[i] += (uint32_t)level;
        level -= offset[i];
        dct[i] = (int16_t)(level < 0 ? 0 : (level ^ sign) - sign);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Exhaustive integer motion search (encoder/motion.cpp:1593-1637, the X265_FULL_SEARCH case of
 * MotionEstimate::motionEstimate; cost of a vector = SAD + mvcost, encoder/bitcost.h:53-56).
 * Candidates are visited in raster order (y outer, x inner) and a candidate replaces the running best only
 * when it is strictly cheaper (COPY2_IF_LT, common.h:193-198), so the initial (bmv, bcost) wins every tie
 * and, among candidates, the first in raster order does.
 *   range = { mvmin.x, mvmin.y, mvmax.x, mvmax.y } in full pels; mvp in quarter pels;
 *   costTab points at the centre of the lambda-scaled table BitCost::setQP builds (bitcost.cpp:44-54):
 *   mvcost(mv) = (uint16_t)(costTab[mv.x - mvp.x] + costTab[mv.y - mvp.y]) with mv = candidate << 2.
 * ------------------------------------------------------------------------------------------------ */
EXPORT void orc_me_full_search(int w, int h, const pixel* fenc, intptr_t sf, const pixel* fref, intptr_t sr,
                               const int32_t* range, const int32_t* mvp, const uint16_t* costTab,
                               int32_t* bmv /* in/out x,y */, int32_t* bcost /* in/out */)
{
    int best = *bcost, bx = bmv[0], by = bmv[1];
    for (int y = range[1]; y <= range[3]; y++)
        for (int x = range[0]; x <= range[2]; x++)
        {
            int cost = orc_sad(w, h, fenc, sf, fref + (intptr_t)y * sr + x, sr);
            cost += (uint16_t)(costTab[(x << 2) - mvp[0]] + costTab[(y << 2) - mvp[1]]);
            if (cost < best) { best = cost; bx = x; by = y; }
        }
    *bcost = best; bmv[0] = bx; bmv[1] = by;
}
EXPORT void orc_me_full_batch(int w, int h, const pixel* fenc, intptr_t sf, const pixel* ref, intptr_t sr,
                              const int32_t* offF, const int32_t* offR, const int32_t* range, const int32_t* mvp,
                              const uint16_t* costTab, int n, int32_t* bmv, int32_t* bcost)
{
    for (int i = 0; i < n; i++)
        orc_me_full_search(w, h, fenc + offF[i], sf, ref + offR[i], sr, range + 4 * i, mvp + 2 * i, costTab, bmv + 2 * i, bcost + i);
}

/* ------------------------------------------------------------------------------------------------
 * MotionEstimate::motionEstimate with searchMethod = DIA, HEX or FULL on full-resolution planes, luma only
 * (encoder/motion.cpp:923-1013 start point, :1016-1138 / :1593-1637 search, :1643-1773 sub-pel refinement; one slice, no vertical
 * restriction, no chroma SATD -- the setSourcePU variant of motion.cpp:166-189).  Returns the cost, writes the qpel vector.
 *   fref: the co-located block (vector 0,0); range: mvmin.x, mvmin.y, mvmax.x, mvmax.y in full pels;
 *   qmvp and mvc[] in quarter pels; costTab as in orc_me_full_search.
 * ------------------------------------------------------------------------------------------------ */
static const int k_subpel_workload[8][5] = {      /* motion.cpp:48-58: hpel_iters, hpel_dirs, qpel_iters, qpel_dirs, hpel_satd */
    { 1, 4, 0, 4, 0 }, { 1, 4, 1, 4, 0 }, { 1, 4, 1, 4, 1 }, { 2, 4, 1, 4, 1 },
    { 2, 4, 2, 4, 1 }, { 1, 8, 1, 8, 1 }, { 2, 8, 1, 8, 1 }, { 2, 8, 2, 8, 1 } };
static const int k_square1[9][2] = { {0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {-1, 1}, {1, -1}, {1, 1} };   /* motion.cpp:67 */

typedef struct { int w, h; const pixel* fenc; intptr_t sf; const pixel* fref; intptr_t sr; const uint16_t* cx; const uint16_t* cy;
                 const pixel* hpel[4]; /* lowres: the four half-pel planes at the co-located block, else NULL */
                 /* chroma residual term of subpelCompare (bChromaSATD): co-located chroma blocks, or chroma == 0 */
                 int chroma, hshift, vshift; const pixel* fencC[2]; intptr_t sfc; const pixel* frefC[2]; intptr_t src;
                 /* SEA: the twelve integral planes at the co-located block (framefilter.cpp:770-832 order), or NULL */
                 const uint32_t* const* integral; const uint16_t* costTab; const int32_t* qmvp; } me_ctx;
/* chroma part of subpelCompare (motion.cpp:1805-1865): the vector in 1/8 chroma samples, 4-tap filters, always SATD */
static int me_chroma_cost(const me_ctx* c, int qx, int qy)
{
    const int mvx = (int32_t)((uint32_t)qx << (1 - c->hshift)), mvy = (int32_t)((uint32_t)qy << (1 - c->vshift));
    const intptr_t off = (mvx >> 3) + (intptr_t)(mvy >> 3) * c->src;
    const int xFrac = mvx & 7, yFrac = mvy & 7, cw = c->w >> c->hshift, ch = c->h >> c->vshift;
    int cost = 0;
    for (int k = 0; k < 2; k++)
    {
        const pixel* r = c->frefC[k] + off;
        if (!(xFrac | yFrac)) { cost += orc_satd(cw, ch, c->fencC[k], c->sfc, r, c->src); continue; }
        pixel buf[64 * 64];
        if (!yFrac) orc_interp_hpp(4, cw, ch, r, c->src, buf, cw, xFrac);
        else if (!xFrac) orc_interp_vpp(4, cw, ch, r, c->src, buf, cw, yFrac);
        else orc_interp_hvpp(4, cw, ch, r, c->src, buf, cw, xFrac, yFrac);
        cost += orc_satd(cw, ch, c->fencC[k], c->sfc, buf, cw);
    }
    return cost;
}
/* ReferencePlanes::lowresQPelCost (common/lowres.h:95-119): a quarter-pel position of a lowres reference is the rounded
 * average of the two nearest half-pel planes (pixelavg_pp, pixel.cpp:586-594), a half / full-pel one is a plane itself */
static int me_lowres_cost(const me_ctx* c, int qx, int qy, int op)
{
    pixel buf[64 * 64];
    const pixel* a = c->hpel[(qy & 2) | ((qx & 2) >> 1)] + (qx >> 2) + (intptr_t)(qy >> 2) * c->sr;
    const pixel* src = a; intptr_t ss = c->sr;
    if ((qx | qy) & 1)
    {
        int bx = qx + (qx & 1), by = qy + (qy & 1);
        const pixel* b = c->hpel[(by & 2) | ((bx & 2) >> 1)] + (bx >> 2) + (intptr_t)(by >> 2) * c->sr;
        for (int y = 0; y < c->h; y++)
            for (int x = 0; x < c->w; x++) buf[y * c->w + x] = (pixel)((a[y * c->sr + x] + b[y * c->sr + x] + 1) >> 1);
        src = buf; ss = c->w;
    }
    return op ? orc_satd(c->w, c->h, c->fenc, c->sf, src, ss) : orc_sad(c->w, c->h, c->fenc, c->sf, src, ss);
}
static int me_subpel(const me_ctx* c, int qx, int qy, int op)   /* subpelCompare, motion.cpp:1775-1803 */
{
    if (c->hpel[0]) return me_lowres_cost(c, qx, qy, op);
    int cost = orc_subpel_cmp(op, c->w, c->h, c->fenc, c->sf, c->fref + (qx >> 2) + (intptr_t)(qy >> 2) * c->sr, c->sr, qx & 3, qy & 3);
    return c->chroma ? cost + me_chroma_cost(c, qx, qy) : cost;
}
static int me_mvcost(const me_ctx* c, int qx, int qy) { return (uint16_t)(c->cx[qx] + c->cy[qy]); }   /* bitcost.h:56 */
static int clip3i(int lo, int hi, int v) { return v < lo ? lo : v > hi ? hi : v; }

/* full-pel candidate: SAD + mvcost(mv << 2), the COST_MV family of macros (motion.cpp:263-330) */
static int me_fpel(const me_ctx* c, int x, int y)
{
    return orc_sad(c->w, c->h, c->fenc, c->sf, c->fref + x + (intptr_t)y * c->sr, c->sr) + me_mvcost(c, x * 4, y * 4);
}
static int me_in_range(const int32_t* range, int x, int y) { return x >= range[0] && x <= range[2] && y >= range[1] && y <= range[3]; }

/* Diamond search, radius 1 (motion.cpp:1016-1039).  The reference packs "which neighbour won" into the low bits of the
 * cost (tags 1, 3, 4, 12), which also fixes the tie-break: a neighbour replaces the centre only when strictly cheaper,
 * and among equal neighbours the earlier one in the order up, down, left, right stays.  Only the candidate's row is
 * range-checked; a step may leave the window horizontally, which then ends the walk. */
static void me_dia_search(const me_ctx* c, const int32_t* range, int merange, int* bx, int* by, int* bcost)
{
    static const int nb[4][2] = { {0, -1}, {0, 1}, {-1, 0}, {1, 0} };
    int x = *bx, y = *by, best = *bcost, i = merange;
    do
    {
        int win = -1;
        for (int k = 0; k < 4; k++)
        {
            int cy = y + nb[k][1];
            if (k < 2 && (cy < range[1] || cy > range[3])) continue;
            int cost = me_fpel(c, x + nb[k][0], cy);
            if (cost < best) { best = cost; win = k; }
        }
        if (win < 0) break;
        x += nb[win][0]; y += nb[win][1];
    }
    while (--i && me_in_range(range, x, y));
    *bx = x; *by = y; *bcost = best;
}

/* Hexagon search, radius 2, then a one-step square refinement (motion.cpp:1041-1138).  First the six corners of the
 * hexagon around the start, in the order of k_hex2[1..6]; then up to merange/2 - 1 steps that test only the three corners
 * not covered by the previous hexagon; candidates outside the vertical range are ignored, the walk stops when the centre
 * leaves the window.  Equal costs keep the earlier candidate (the reference's tag-in-low-bits comparison). */
static const int k_hex2[8][2] = { {-1, -2}, {-2, 0}, {-1, 2}, {1, 2}, {2, 0}, {1, -2}, {-1, -2}, {-2, 0} };   /* motion.cpp:65 */
static void me_hex_search(const me_ctx* c, const int32_t* range, int merange, int* bx, int* by, int* bcost)
{
    int x = *bx, y = *by, best = *bcost, win = 0;
    for (int t = 2; t <= 7; t++)
    {
        int cx = x + k_hex2[t - 1][0], cy = y + k_hex2[t - 1][1];
        if (cy < range[1] || cy > range[3]) continue;
        int cost = me_fpel(c, cx, cy);
        if (cost < best) { best = cost; win = t; }
    }
    if (win)
    {
        int dir = win - 2;
        x += k_hex2[dir + 1][0]; y += k_hex2[dir + 1][1];       /* the winner passed the row check above */
        for (int i = (merange >> 1) - 1; i > 0 && me_in_range(range, x, y); i--)
        {
            win = 0;
            for (int t = 1; t <= 3; t++)
            {
                int cx = x + k_hex2[dir + t - 1][0], cy = y + k_hex2[dir + t - 1][1];
                if (cy < range[1] || cy > range[3]) continue;
                int cost = me_fpel(c, cx, cy);
                if (cost < best) { best = cost; win = t; }
            }
            if (!win) break;
            dir = (dir + win - 2 + 6) % 6;                      /* mod6m1[dir + 1], motion.cpp:66 */
            x += k_hex2[dir + 1][0]; y += k_hex2[dir + 1][1];
        }
    }
    /* square refine: all eight neighbours of the final centre, first the cross then the corners (square1 order) */
    win = 0;
    for (int d = 1; d <= 8; d++)
    {
        int cx = x + k_square1[d][0], cy = y + k_square1[d][1];
        if (k_square1[d][1] && (cy < range[1] || cy > range[3])) continue;
        int cost = me_fpel(c, cx, cy);
        if (cost < best) { best = cost; win = d; }
    }
    *bx = x + k_square1[win][0]; *by = y + k_square1[win][1]; *bcost = best;
}

/* Star search (motion.cpp:386-630 StarPatternSearch, :1327-1435 the X265_STAR_SEARCH case; adapted from HM).
 * One pattern pass tests rings around a fixed centre: distance 1 (4 points), 2 / 4 / 8 (8 points: the four axis points at
 * the distance, the four diagonal ones at half of it), 16 .. merange (16 points on a diamond).  The reference has a fast
 * path when the whole ring is inside the window and a per-point checked path otherwise; both visit the same points in
 * the same order and a point is measured exactly when it lies inside the window, which is what is restated here.  A pass
 * ends early after `earlyExit` consecutive rings without improvement.  Point numbers (1..8, 0 for the big rings) and the
 * distance of the best point steer what follows. */
typedef struct { int x, y, cost, point, dist; } me_star_t;
static void me_star_try(const me_ctx* c, const int32_t* range, me_star_t* b, int x, int y, int point, int dist)
{
    if (!me_in_range(range, x, y)) return;
    int cost = me_fpel(c, x, y);
    if (cost < b->cost) { b->cost = cost; b->x = x; b->y = y; b->point = point; b->dist = dist; }
}
static void me_star_pattern(const me_ctx* c, const int32_t* range, me_star_t* b, int earlyExit, int merange)
{
    const int ox = b->x, oy = b->y;
    int rounds = 0, saved = b->cost;
    for (int dist = 1; dist <= 8 || dist <= (int16_t)merange; dist <<= 1)
    {
        if (dist > 8 && dist > (int16_t)merange) break;
        if (dist > 1) saved = b->cost;
        if (dist == 1)
        {
            me_star_try(c, range, b, ox, oy - 1, 2, 1); me_star_try(c, range, b, ox - 1, oy, 4, 1);
            me_star_try(c, range, b, ox + 1, oy, 5, 1); me_star_try(c, range, b, ox, oy + 1, 7, 1);
        }
        else if (dist <= 8)
        {
            const int h = dist >> 1;
            me_star_try(c, range, b, ox, oy - dist, 2, dist); me_star_try(c, range, b, ox - h, oy - h, 1, h);
            me_star_try(c, range, b, ox + h, oy - h, 3, h);   me_star_try(c, range, b, ox - dist, oy, 4, dist);
            me_star_try(c, range, b, ox + dist, oy, 5, dist); me_star_try(c, range, b, ox - h, oy + h, 6, h);
            me_star_try(c, range, b, ox + h, oy + h, 8, h);   me_star_try(c, range, b, ox, oy + dist, 7, dist);
        }
        else
        {
            const int q = dist >> 2;
            me_star_try(c, range, b, ox, oy - dist, 0, dist); me_star_try(c, range, b, ox - dist, oy, 0, dist);
            me_star_try(c, range, b, ox + dist, oy, 0, dist); me_star_try(c, range, b, ox, oy + dist, 0, dist);
            for (int i = 1; i < 4; i++)
            {
                me_star_try(c, range, b, ox - q * i, oy - dist + q * i, 0, dist); me_star_try(c, range, b, ox + q * i, oy - dist + q * i, 0, dist);
                me_star_try(c, range, b, ox - q * i, oy + dist - q * i, 0, dist); me_star_try(c, range, b, ox + q * i, oy + dist - q * i, 0, dist);
            }
        }
        if (b->cost < saved) rounds = 0;
        else if (++rounds >= earlyExit) return;
    }
}
/* the two outer neighbours of a distance-1 winner, indexed by its point number (motion.cpp:76-86 `offsets`) */
static const int k_two_point[16][2] = { {-1, 0}, {0, -1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {-1, -1},
                                        {1, -1}, {1, 1}, {-1, 0}, {0, 1}, {-1, 1}, {1, 1}, {1, 0}, {0, 1} };
static void me_two_point(const me_ctx* c, const int32_t* range, me_star_t* b)
{
    const int x = b->x, y = b->y, p = (b->point - 1) * 2;      /* both neighbours are taken around the winner as it was */
    for (int k = 0; k < 2; k++)
    {
        int cx = x + k_two_point[p + k][0], cy = y + k_two_point[p + k][1];
        if (!me_in_range(range, cx, cy)) continue;
        int cost = me_fpel(c, cx, cy);
        if (cost < b->cost) { b->cost = cost; b->x = cx; b->y = cy; }
    }
}
static void me_star_search(const me_ctx* c, const int32_t* range, int merange, int* bx, int* by, int* bcost)
{
    me_star_t b = { *bx, *by, *bcost, 0, 0 };
    me_star_pattern(c, range, &b, 3, merange);
    int done = 0;
    if (b.dist == 1)
    {   /* :1336-1362 */
        if (!b.point) done = 1;
        else
        {
            int saved = b.cost;
            me_two_point(c, range, &b);
            if (b.cost == saved) done = 1;
        }
    }
    if (!done)
    {
        if (b.dist > 5)
        {   /* :1364-1399 raster over the window in steps of 5; the reference measures four columns per sad_x4 call and
             * charges the fourth one mvcost(mv << 3) instead of mv << 2 -- reproduced, it decides vectors */
            for (int y = range[1]; y <= range[3]; y += 5)
                for (int x = range[0]; x <= range[2]; x += 5)
                {
                    if (x + 15 <= range[2])
                    {
                        for (int k = 0; k < 4; k++, x += (k < 4 ? 5 : 0))
                        {
                            int cost = orc_sad(c->w, c->h, c->fenc, c->sf, c->fref + x + (intptr_t)y * c->sr, c->sr)
                                     + (k < 3 ? me_mvcost(c, x * 4, y * 4) : me_mvcost(c, x * 8, y * 8));
                            if (cost < b.cost) { b.cost = cost; b.x = x; b.y = y; }
                        }
                    }
                    else
                    {
                        int cost = me_fpel(c, x, y);
                        if (cost < b.cost) { b.cost = cost; b.x = x; b.y = y; }
                    }
                }
        }
        while (b.dist > 0)
        {   /* :1401-1433 re-centred passes until one brings nothing */
            b.dist = 0; b.point = 0;
            me_star_pattern(c, range, &b, 32, merange);
            if (b.dist == 1)
            {
                if (b.point) me_two_point(c, range, &b);
                break;
            }
        }
    }
    *bx = b.x; *by = b.y; *bcost = b.cost;
}

/* Uneven multi-hexagon search (motion.cpp:1142-1324, from x264): small diamonds around the predictor, the zero vector and
 * the running best; an early-termination ladder driven by SAD thresholds scaled by the PU height (sizeScale = H*H >> 4,
 * motion.cpp:124-152); a cross whose reach adapts to how much the neighbour vectors disagree; a 5x5 corner check; rings
 * of the 16-point hexagon at radius i = 1 .. merange/4; and finally the plain hexagon search from wherever that ended.
 * `x4` candidates are judged around a fixed origin with only their row range-checked; single candidates of the cross
 * are checked on the side they move to.  Written with small helpers instead of the reference's macro cascade. */
typedef struct { const me_ctx* c; const int32_t* range; int x, y, cost; } me_umh_t;
static void umh_try(me_umh_t* u, int x, int y)                 /* COST_MV */
{
    int cost = me_fpel(u->c, x, y);
    if (cost < u->cost) { u->cost = cost; u->x = x; u->y = y; }
}
static void umh_x4(me_umh_t* u, int ox, int oy, const int (*d)[2]) /* COST_MV_X4 around (ox, oy) */
{
    for (int k = 0; k < 4; k++)
    {
        int y = oy + d[k][1];
        if (y < u->range[1] || y > u->range[3]) continue;
        umh_try(u, ox + d[k][0], y);
    }
}
static void umh_cross(me_umh_t* u, int ox, int oy, int start, int x_max, int y_max)   /* CROSS, motion.cpp:359-385 */
{
    const int32_t* r = u->range;
    int i = start;
    int roomx = r[2] - ox < ox - r[0] ? r[2] - ox : ox - r[0];
    if (x_max <= roomx)
        for (; i < x_max - 2; i += 4)
        {
            const int d[4][2] = { {i, 0}, {-i, 0}, {i + 2, 0}, {-i - 2, 0} };
            umh_x4(u, ox, oy, d);
        }
    for (; i < x_max; i += 2)
    {
        if (ox + i <= r[2]) umh_try(u, ox + i, oy);
        if (ox - i >= r[0]) umh_try(u, ox - i, oy);
    }
    i = start;
    int roomy = r[3] - oy < oy - r[1] ? r[3] - oy : oy - r[1];
    if (y_max <= roomy)
        for (; i < y_max - 2; i += 4)
        {
            const int d[4][2] = { {0, i}, {0, -i}, {0, i + 2}, {0, -i - 2} };
            umh_x4(u, ox, oy, d);
        }
    for (; i < y_max; i += 2)
    {
        if (oy + i <= r[3]) umh_try(u, ox, oy + i);
        if (oy - i >= r[1]) umh_try(u, ox, oy - i);
    }
}
static const int k_hex4[16][2] = { {0, -4}, {0, 4}, {-2, -3}, {2, -3}, {-4, -2}, {4, -2}, {-4, -1}, {4, -1},
                                   {-4, 0}, {4, 0}, {-4, 1}, {4, 1}, {-4, 2}, {4, 2}, {-2, 3}, {2, 3} };   /* motion.cpp:68-74 */
static const int k_dia1[4][2] = { {0, -1}, {0, 1}, {-1, 0}, {1, 0} };
/* returns 1 when the search goes on into the hexagon stage (me_hex2), 0 when it ended inside UMH */
static int me_umh_search(const me_ctx* c, const int32_t* range, int* merangeIO, int pmvx, int pmvy /* full pel */,
                         const int32_t* qmvp, int numCand, const int32_t* mvc, int* bx, int* by, int* bcost)
{
    me_umh_t u = { c, range, *bx, *by, *bcost };
    int merange = *merangeIO;
    const int scale = (c->h * c->h) >> 4;
#define UMH_THRESH(v) (u.cost < (((v) >> 4) * scale))
    int cross_start = 1;
    const int ucost1 = u.cost;
    umh_x4(&u, pmvx, pmvy, k_dia1);
    if (pmvx | pmvy) umh_x4(&u, 0, 0, k_dia1);
    const int ucost2 = u.cost;
    if ((u.x | u.y) && (u.x != pmvx || u.y != pmvy)) umh_x4(&u, u.x, u.y, k_dia1);
    if (u.cost == ucost2) cross_start = 3;

    int ox = u.x, oy = u.y;
    if (u.cost == ucost2 && UMH_THRESH(2000))
    {
        static const int oct_a[4][2] = { {0, -2}, {-1, -1}, {1, -1}, {-2, 0} }, oct_b[4][2] = { {2, 0}, {-1, 1}, {1, 1}, {0, 2} };
        umh_x4(&u, ox, oy, oct_a); umh_x4(&u, ox, oy, oct_b);
        if (u.cost == ucost1 && UMH_THRESH(500)) { *bx = u.x; *by = u.y; *bcost = u.cost; return 0; }
        if (u.cost == ucost2)
        {
            static const int k2a[4][2] = { {-1, -2}, {1, -2}, {-2, -1}, {2, -1} }, k2b[4][2] = { {-2, 1}, {2, 1}, {-1, 2}, {1, 2} };
            const int reach = (int16_t)(merange >> 1) | 1;
            umh_cross(&u, ox, oy, 3, reach, reach);
            umh_x4(&u, ox, oy, k2a); umh_x4(&u, ox, oy, k2b);
            if (u.cost == ucost2) { *bx = u.x; *by = u.y; *bcost = u.cost; return 0; }
            cross_start = reach + 2;
        }
    }
    if (numCand)
    {   /* search range scaled by the disagreement of the predictors and by how good the match already is */
        static const uint8_t range_mul[4][4] = { { 3, 3, 4, 4 }, { 3, 4, 4, 4 }, { 4, 4, 4, 5 }, { 4, 4, 5, 6 } };
        const int is64 = c->w == 64 && c->h == 64;
        int mvd, denom = 1;
        if (numCand == 1)
            mvd = is64 ? 25 : iabs(qmvp[0] - mvc[0]) + iabs(qmvp[1] - mvc[1]);
        else
        {
            denom = numCand - 1;
            mvd = 0;
            if (!is64) { mvd = iabs(qmvp[0] - mvc[0]) + iabs(qmvp[1] - mvc[1]); denom++; }
            for (int i = 0; i < numCand - 1; i++)
                mvd += iabs(mvc[2 * i] - mvc[2 * i + 2]) + iabs(mvc[2 * i + 1] - mvc[2 * i + 3]);
        }
        const int sad_ctx = UMH_THRESH(1000) ? 0 : UMH_THRESH(2000) ? 1 : UMH_THRESH(4000) ? 2 : 3;
        const int mvd_ctx = mvd < 10 * denom ? 0 : mvd < 20 * denom ? 1 : mvd < 40 * denom ? 2 : 3;
        merange = (merange * range_mul[mvd_ctx][sad_ctx]) >> 2;
    }
    {   /* the cross and the corners stay centred where the diamonds ended (the reference's FIXME) */
        static const int corners[4][2] = { {-2, -2}, {-2, 2}, {2, -2}, {2, 2} };
        umh_cross(&u, ox, oy, cross_start, merange, merange >> 1);
        umh_x4(&u, ox, oy, corners);
    }
    /* hexagon grid around the new best */
    ox = u.x; oy = u.y;
    uint16_t i = 1;
    do
    {
        int room = range[2] - ox;
        if (ox - range[0] < room) room = ox - range[0];
        if (range[3] - oy < room) room = range[3] - oy;
        if (oy - range[1] < room) room = oy - range[1];
        if (4 * i > room)
        {
            for (int j = 0; j < 16; j++)
            {
                int x = ox + k_hex4[j][0] * i, y = oy + k_hex4[j][1] * i;
                if (me_in_range(range, x, y)) umh_try(&u, x, y);
            }
        }
        else
        {   /* whole ring inside the window: all sixteen judged against the best before the ring, first-best wins */
            int best = u.cost, dir = -1;
            for (int j = 0; j < 16; j++)
            {
                int x = ox + k_hex4[j][0] * i, y = oy + k_hex4[j][1] * i;
                int cost = orc_sad(c->w, c->h, c->fenc, c->sf, c->fref + x + (intptr_t)y * c->sr, c->sr) + c->cx[x * 4] + c->cy[y * 4];
                if (cost < best) { best = cost; dir = j; }
            }
            if (dir >= 0) { u.cost = best; u.x = ox + k_hex4[dir][0] * i; u.y = oy + k_hex4[dir][1] * i; }
        }
    }
    while (++i <= merange >> 2);
#undef UMH_THRESH
    *bx = u.x; *by = u.y; *bcost = u.cost; *merangeIO = merange;
    return me_in_range(range, u.x, u.y);
}

/* Successive elimination (motion.cpp:1438-1591): every row of the window (start +- merange, clipped) is pre-filtered with
 * `ads` -- |DC of the PU's sub-blocks - box sums of the reference| + the row's x-cost against a threshold -- and only the
 * survivors get a real SAD.  Restated with the reference's cost bookkeeping as it is: the row cost is taken from the table
 * shifted by the predictor a second time and indexed by the full-pel y, then << 2; candidates measured three at a time
 * are charged only that shifted x-cost while the row cost is temporarily subtracted from the running best; the window
 * width is rounded up to a multiple of 4, so up to three columns right of it are examined too. */
static void me_sea_search(const me_ctx* c, const int32_t* range, int merange, int* bx, int* by, int* bcostIO)
{
    const int w = c->w, h = c->h;
    const int ox = *bx, oy = *by;
    const int minX = ox - merange > range[0] ? ox - merange : range[0], minY = oy - merange > range[1] ? oy - merange : range[1];
    const int maxX = ox + merange < range[2] ? ox + merange : range[2], maxY = oy + merange < range[3] ? oy + merange : range[3];
    const uint16_t* pcx = c->cx - c->qmvp[0];
    const uint16_t* pcy = c->cy - c->qmvp[1];
    const int width = (maxX - minX + 3) & ~3;
    int deltaX = w <= 8 ? w : w >> 1, deltaY = h <= 8 ? h : h >> 1;
    const int vertical = (w == 32 && h == 64) || (w == 16 && h == 32) || (w == 8 && h == 16) || (w == 4 && h == 8);
    const int horizontal = (w == 64 && h == 32) || (w == 32 && h == 16) || (w == 16 && h == 8) || (w == 8 && h == 4);
    const int smallRect = (w == 4 && h == 4) || (w == 16 && h == 12) || (w == 12 && h == 16) || (w == 16 && h == 4) || (w == 4 && h == 16);
    const int asym = (w == 12 && h == 16) || (w == 4 && h == 16) || (w == 24 && h == 32) || (w == 8 && h == 32) || (w == 48 && h == 64) ||
                     (w == 16 && h == 64) || (w == 16 && h == 12) || (w == 16 && h == 4) || (w == 32 && h == 24) || (w == 32 && h == 8) ||
                     (w == 64 && h == 48) || (w == 64 && h == 16);
    int tw, th;                                                 /* the block whose DC each of the four sad_x4 references takes */
    if (vertical) { tw = w; th = h >> 1; }
    else if (horizontal) { tw = w >> 1; th = h; }
    else if (asym) { tw = smallRect ? w : w >> 1; th = smallRect ? h : h >> 1; }
    else { tw = w <= 8 ? w : w >> 1; th = w <= 8 ? h : h >> 1; }
    int encDC[4];
    {
        const int offs[4][2] = { {0, 0}, {deltaX, 0}, {0, deltaY}, {deltaX, deltaY} };
        for (int k = 0; k < 4; k++)
        {   /* sad against a zero block = sum of the samples; the reference reads past the PU inside its 64-stride cache,
               where the rest of a previous, larger PU may still lie -- only in-block sums are ever used by ads */
            int sum = 0;
            for (int y = 0; y < th; y++)
                for (int x = 0; x < tw; x++)
                {
                    int yy = offs[k][1] + y, xx = offs[k][0] + x;
                    sum += (yy < h && xx < w) ? c->fenc[yy * c->sf + xx] : 0;
                }
            encDC[k] = sum;
        }
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
    const uint32_t* sumsBase = c->integral[plane];
    if ((w == h && w >= 16) || vertical || (w == 12 && h == 16) || (w == 4 && h == 16) || (w == 24 && h == 32) || (w == 8 && h == 32) ||
        (w == 48 && h == 64) || (w == 16 && h == 64))
        deltaY *= (int)c->sr;
    if (vertical) encDC[1] = encDC[2];
    if (horizontal) deltaY = deltaX;

    int x = *bx, y = *by, bcost = *bcostIO;
    uint16_t* rowCost = (uint16_t*)malloc((size_t)(width + 4) * sizeof(uint16_t));
    int16_t* mvs = (int16_t*)malloc((size_t)(2 * merange + 8 + width) * sizeof(int16_t));
    for (int i = 0; i < width; i++) rowCost[i] = c->cx[(minX + i) * 4];       /* m_fpelMvCosts: the x-cost at full-pel columns */
    for (int ty = minY; ty <= maxY; ty++)
    {
        const int ycost = pcy[ty] << 2;
        if (bcost <= ycost) continue;
        bcost -= ycost;
        const int xn = orc_ads(w, h, encDC, sumsBase + minX + (intptr_t)ty * c->sr, deltaY, rowCost, mvs, width, bcost);
        int i = 0;
        for (; i < xn - 2; i += 3)
            for (int k = 0; k < 3; k++)
            {
                const int mx = minX + mvs[i + k];
                int cost = orc_sad(w, h, c->fenc, c->sf, c->fref + mx + (intptr_t)ty * c->sr, c->sr) + pcx[mx * 4];
                if (cost < bcost) { bcost = cost; x = mx; y = ty; }
            }
        bcost += ycost;
        for (; i < xn; i++)
        {
            const int mx = minX + mvs[i];
            int cost = me_fpel(c, mx, ty);
            if (cost < bcost) { bcost = cost; x = mx; y = ty; }
        }
    }
    free(rowCost); free(mvs);
    *bx = x; *by = y; *bcostIO = bcost;
}

/* method: X265_DIA_SEARCH 0, X265_HEX_SEARCH 1, X265_UMH_SEARCH 2, X265_STAR_SEARCH 3, X265_SEA 4 (needs the integral planes), X265_FULL_SEARCH 5 (x265.h:511-519); others return -1 */
static int me_estimate(me_ctx cc, int method, int merange, int subme, const int32_t* range, const int32_t* qmvp, int numCand, const int32_t* mvc,
                       const uint16_t* costTab, int32_t* outQMv)
{
    if (method < 0 || method > 5 || (method == 4 && !cc.integral)) return -1;
    const me_ctx c = cc;
    const int w = c.w, h = c.h; const pixel* fenc = c.fenc; const pixel* fref = c.fref; const intptr_t sf = c.sf, sr = c.sr;
    const int lowres = c.hpel[0] != 0;
    const int qminx = range[0] * 4, qminy = range[1] * 4, qmaxx = range[2] * 4, qmaxy = range[3] * 4;

    /* :954-975 SAD at the clipped predictor, then at its full-pel rounding */
    int pmvx = clip3i(qminx, qmaxx, qmvp[0]), pmvy = clip3i(qminy, qmaxy, qmvp[1]);
    int bestprex = pmvx, bestprey = pmvy;
    int bprecost = me_subpel(&c, pmvx, pmvy, 0);
    int bmvx = (pmvx + 2) >> 2, bmvy = (pmvy + 2) >> 2;
    int bcost = bprecost;
    if ((pmvx | pmvy) & 3)
        bcost = orc_sad(w, h, fenc, sf, fref + bmvx + (intptr_t)bmvy * sr, sr) + me_mvcost(&c, bmvx * 4, bmvy * 4);
    /* :978-988 the zero vector */
    if (pmvx | pmvy)
    {
        int cost = orc_sad(w, h, fenc, sf, fref, sr) + me_mvcost(&c, 0, 0);
        if (cost < bcost)
        {
            bcost = cost; bmvx = 0;
            bmvy = range[3] < 0 ? range[3] : 0;
            if (bmvy < range[1]) bmvy = range[1];
        }
    }
    /* :992-1004 neighbour candidates compete for the sub-pel starting point only */
    for (int i = 0; i < numCand; i++)
    {
        int mx = clip3i(qminx, qmaxx, mvc[2 * i]), my = clip3i(qminy, qmaxy, mvc[2 * i + 1]);
        if ((mx | my) && (mx != pmvx || my != pmvy) && (mx != bestprex || my != bestprey))
        {
            int cost = me_subpel(&c, mx, my, 0) + me_mvcost(&c, mx, my);
            if (cost < bprecost) { bprecost = cost; bestprex = mx; bestprey = my; }
        }
    }
    if (bcost == 0)
    {   /* :1008-1012 */
        outQMv[0] = bmvx * 4; outQMv[1] = bmvy * 4;
        return me_mvcost(&c, bmvx * 4, bmvy * 4);
    }
    if (method == 2)
    {   /* :1142-1324, falling through to the hexagon search when it ends inside the window */
        int mr = merange;
        if (me_umh_search(&c, range, &mr, (pmvx + 2) >> 2, (pmvy + 2) >> 2, qmvp, numCand, mvc, &bmvx, &bmvy, &bcost))
            me_hex_search(&c, range, mr, &bmvx, &bmvy, &bcost);
    }
    else if (method == 4) me_sea_search(&c, range, merange, &bmvx, &bmvy, &bcost);
    else if (method == 0) me_dia_search(&c, range, merange, &bmvx, &bmvy, &bcost);
    else if (method == 1) me_hex_search(&c, range, merange, &bmvx, &bmvy, &bcost);
    else if (method == 3) me_star_search(&c, range, merange, &bmvx, &bmvy, &bcost);
    else
    {   /* :1593-1637 */
        int32_t mv[2] = { bmvx, bmvy }, bc = bcost;
        orc_me_full_search(w, h, fenc, sf, fref, sr, range, qmvp, costTab, mv, &bc);
        bmvx = mv[0]; bmvy = mv[1]; bcost = bc;
    }
    /* :1643-1650 */
    if (bprecost < bcost) { bmvx = bestprex; bmvy = bestprey; bcost = bprecost; }
    else { bmvx *= 4; bmvy *= 4; }

    const int* wl = k_subpel_workload[subme];
    if (!bcost)
        bcost = me_mvcost(&c, bmvx, bmvy);              /* :1661-1666 */
    else if (lowres)
    {   /* :1667-1698: one half-pel step judged by SAD, re-measured with SATD, one quarter-pel step */
        int bdir = 0;
        for (int i = 1; i <= wl[1]; i++)
        {
            int qx = bmvx + k_square1[i][0] * 2, qy = bmvy + k_square1[i][1] * 2;
            if (qy < qminy || qy > qmaxy) continue;
            int cost = me_subpel(&c, qx, qy, 0) + me_mvcost(&c, qx, qy);
            if (cost < bcost) { bcost = cost; bdir = i; }
        }
        bmvx += k_square1[bdir][0] * 2; bmvy += k_square1[bdir][1] * 2;
        bcost = me_subpel(&c, bmvx, bmvy, 1) + me_mvcost(&c, bmvx, bmvy);
        bdir = 0;
        for (int i = 1; i <= wl[3]; i++)
        {
            int qx = bmvx + k_square1[i][0], qy = bmvy + k_square1[i][1];
            if (qy < qminy || qy > qmaxy) continue;
            int cost = me_subpel(&c, qx, qy, 1) + me_mvcost(&c, qx, qy);
            if (cost < bcost) { bcost = cost; bdir = i; }
        }
        bmvx += k_square1[bdir][0]; bmvy += k_square1[bdir][1];
    }
    else
    {   /* :1700-1757 */
        int hpelop = 0;
        if (wl[4]) { bcost = me_subpel(&c, bmvx, bmvy, 1) + me_mvcost(&c, bmvx, bmvy); hpelop = 1; }
        for (int iter = 0; iter < wl[0]; iter++)
        {
            int bdir = 0;
            for (int i = 1; i <= wl[1]; i++)
            {
                int qx = bmvx + k_square1[i][0] * 2, qy = bmvy + k_square1[i][1] * 2;
                if (qy < qminy || qy > qmaxy) continue;
                int cost = me_subpel(&c, qx, qy, hpelop) + me_mvcost(&c, qx, qy);
                if (cost < bcost) { bcost = cost; bdir = i; }
            }
            if (!bdir) break;
            bmvx += k_square1[bdir][0] * 2; bmvy += k_square1[bdir][1] * 2;
        }
        if (!wl[4]) bcost = me_subpel(&c, bmvx, bmvy, 1) + me_mvcost(&c, bmvx, bmvy);
        for (int iter = 0; iter < wl[2]; iter++)
        {
            int bdir = 0;
            for (int i = 1; i <= wl[3]; i++)
            {
                int qx = bmvx + k_square1[i][0], qy = bmvy + k_square1[i][1];
                if (qy < qminy || qy > qmaxy) continue;
                int cost = me_subpel(&c, qx, qy, 1) + me_mvcost(&c, qx, qy);
                if (cost < bcost) { bcost = cost; bdir = i; }
            }
            if (!bdir) break;
            bmvx += k_square1[bdir][0]; bmvy += k_square1[bdir][1];
        }
    }
    /* :1762-1768 the zero vector gets a last chance; the returned cost stays the winner's */
    if (bmvx | bmvy)
    {
        int cost = me_subpel(&c, 0, 0, 1) + me_mvcost(&c, 0, 0);
        if (cost <= bcost) { bmvx = 0; bmvy = 0; }
    }
    outQMv[0] = bmvx; outQMv[1] = bmvy;
    return bcost;
}

EXPORT int orc_motion_estimate(int method, int merange, int subme, int w, int h, const pixel* fenc, intptr_t sf, const pixel* fref, intptr_t sr,
                               const int32_t* range, const int32_t* qmvp, int numCand, const int32_t* mvc,
                               const uint16_t* costTab, int32_t* outQMv)
{
    me_ctx c = { w, h, fenc, sf, fref, sr, costTab - qmvp[0], costTab - qmvp[1], { 0, 0, 0, 0 }, 0, 0, 0, { 0, 0 }, 0, { 0, 0 }, 0 };
    return me_estimate(c, method, merange, subme, range, qmvp, numCand, mvc, costTab, outQMv);
}
/* one chroma block of that term, for the kernel test: SATD(fenc, 4-tap interpolation of fref at xFrac, yFrac eighths) */
EXPORT int orc_subpel_cmp_chroma(int w, int h, const pixel* fenc, intptr_t sf, const pixel* fref, intptr_t sr, int xFrac, int yFrac)
{
    if (!(xFrac | yFrac)) return orc_satd(w, h, fenc, sf, fref, sr);
    pixel buf[64 * 64];
    if (!yFrac) orc_interp_hpp(4, w, h, fref, sr, buf, w, xFrac);
    else if (!xFrac) orc_interp_vpp(4, w, h, fref, sr, buf, w, yFrac);
    else orc_interp_hvpp(4, w, h, fref, sr, buf, w, xFrac, yFrac);
    return orc_satd(w, h, fenc, sf, buf, w);
}
/* the encoder's call (setSourcePU of motion.cpp:222-247 with bChroma): from subme 3 on every subpelCompare also charges the
 * SATD of both chroma blocks, when the chroma block is a multiple of 4x4 (non-NULL chroma satd slot, pixel.cpp:1217-1243).
 * fencC / frefC: co-located Cb and Cr blocks; hshift / vshift: 1,1 for 4:2:0, 0,0 for 4:4:4. */
EXPORT int orc_motion_estimate_chroma(int method, int merange, int subme, int w, int h, const pixel* fenc, intptr_t sf, const pixel* fref, intptr_t sr,
                                      const pixel* fencCb, const pixel* fencCr, intptr_t sfc, const pixel* frefCb, const pixel* frefCr, intptr_t src,
                                      int hshift, int vshift, const int32_t* range, const int32_t* qmvp, int numCand, const int32_t* mvc,
                                      const uint16_t* costTab, int32_t* outQMv)
{
    const int on = subme > 2 && !((w >> hshift) & 3) && !((h >> vshift) & 3);
    me_ctx c = { w, h, fenc, sf, fref, sr, costTab - qmvp[0], costTab - qmvp[1], { 0, 0, 0, 0 },
                 on, hshift, vshift, { fencCb, fencCr }, sfc, { frefCb, frefCr }, src };
    return me_estimate(c, method, merange, subme, range, qmvp, numCand, mvc, costTab, outQMv);
}
/* the lookahead's call (encoder/slicetype.cpp:4484-4566): lowres reference = four half-pel planes `pitch` samples apart
 * (frameInitLowres), 8x8 blocks, no neighbour candidates; refBlock points at the co-located block in plane 0 */
EXPORT int orc_lowres_motion_estimate(int method, int merange, int subme, int w, int h, const pixel* fenc, intptr_t sf,
                                      const pixel* refBlock, intptr_t sr, size_t pitch, const int32_t* range, const int32_t* qmvp,
                                      const uint16_t* costTab, int32_t* outQMv)
{
    me_ctx c = { w, h, fenc, sf, refBlock, sr, costTab - qmvp[0], costTab - qmvp[1],
                 { refBlock, refBlock + pitch, refBlock + 2 * pitch, refBlock + 3 * pitch }, 0, 0, 0, { 0, 0 }, 0, { 0, 0 }, 0 };
    return me_estimate(c, method, merange, subme, range, qmvp, 0, 0, costTab, outQMv);
}

/* ------------------------------------------------------------------------------------------------
 * Intra prediction (common/intrapred.cpp:31-234): reference-sample smoothing, DC, planar and the 33 angular modes, and
 * the lookahead's intra cost estimate built on them (encoder/slicetype.cpp:755-864 LookaheadTLD::lowresIntraEstimate).
 * Neighbour array layout (intrapred.cpp:36-50): [0] top-left, [1 .. 2N] above and above-right, [2N+1 .. 4N] left and
 * below-left.  Written per output sample instead of per row / with a flip at the end: a horizontal mode (< 18) is the
 * vertical mode 36 - m... i.e. the same angular rule with above and left swapped and the block transposed.
 * ------------------------------------------------------------------------------------------------ */
static const uint8_t k_intra_filter_flags[35] = {       /* constants.cpp:561-567 */
    0x38, 0x00,
    0x38, 0x30, 0x30, 0x30, 0x30, 0x30, 0x30, 0x20, 0x00, 0x20, 0x30, 0x30, 0x30, 0x30, 0x30, 0x30,
    0x38, 0x30, 0x30, 0x30, 0x30, 0x30, 0x30, 0x20, 0x00, 0x20, 0x30, 0x30, 0x30, 0x30, 0x30, 0x30,
    0x38 };
static const int k_intra_angle[17] = { -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32 };   /* intrapred.cpp:122 */
static const int k_intra_inv_angle[8] = { 4096, 1638, 910, 630, 482, 390, 315, 256 };                         /* intrapred.cpp:123 */

EXPORT void orc_intra_filter(int N, const pixel* s, pixel* f)       /* intrapred.cpp:31-51 */
{
    const int n2 = 2 * N;
    for (int i = 1; i < 2 * n2; i++)
        f[i] = (pixel)((2 * s[i] + s[i - 1] + s[i + 1] + 2) >> 2);
    /* the two chains meet at the corner: top-left sees above[0] and left[0], left[0] sees the top-left */
    f[0] = (pixel)((2 * s[0] + s[1] + s[n2 + 1] + 2) >> 2);
    f[n2 + 1] = (pixel)((2 * s[n2 + 1] + s[0] + s[n2 + 2] + 2) >> 2);
    f[n2] = s[n2]; f[2 * n2] = s[2 * n2];
}
/* neighbour j as a horizontal mode sees it: above and left exchanged (intrapred.cpp:111-119) */
static int intra_nb(const pixel* s, int N, int hor, int j)
{
    if (!hor || j == 0) return s[j];
    return j <= 2 * N ? s[2 * N + j] : s[j - 2 * N];
}
/* reference sample i of the angular rule (intrapred.cpp:147-171): the above row, extended to the left by the left
 * column projected along the inverse angle when the angle is negative */
static int intra_ref(const pixel* s, int N, int hor, int angleOffset, int i)
{
    if (i >= -1) return intra_nb(s, N, hor, i + 1);
    const int k = -2 - i;
    return intra_nb(s, N, hor, 2 * N + ((128 + (k + 1) * k_intra_inv_angle[-angleOffset - 1]) >> 8));
}
EXPORT void orc_intra_pred(int N, int mode, const pixel* s, int bFilter, pixel* dst, intptr_t ds)
{
    if (mode == 0)
    {   /* planar, intrapred.cpp:87-100 */
        int lg = 0; while ((1 << lg) < N) lg++;
        const pixel* above = s + 1; const pixel* left = s + 2 * N + 1;
        for (int y = 0; y < N; y++)
            for (int x = 0; x < N; x++)
                dst[y * ds + x] = (pixel)(((N - 1 - x) * left[y] + (N - 1 - y) * above[x] + (x + 1) * above[N] + (y + 1) * left[N] + N) >> (lg + 1));
        return;
    }
    if (mode == 1)
    {   /* DC with optional edge smoothing, intrapred.cpp:53-85 */
        int dc = N;
        for (int i = 0; i < N; i++) dc += s[1 + i] + s[2 * N + 1 + i];
        dc /= 2 * N;
        for (int y = 0; y < N; y++)
            for (int x = 0; x < N; x++)
            {
                int v = dc;
                if (bFilter)
                {
                    if (!x && !y) v = (s[1] + s[2 * N + 1] + 2 * dc + 2) >> 2;
                    else if (!y) v = (s[1 + x] + 3 * dc + 2) >> 2;
                    else if (!x) v = (s[2 * N + 1 + y] + 3 * dc + 2) >> 2;
                }
                dst[y * ds + x] = (pixel)v;
            }
        return;
    }
    /* angular, intrapred.cpp:102-204 */
    const int hor = mode < 18;
    const int angleOffset = hor ? 10 - mode : mode - 26;
    const int angle = k_intra_angle[8 + angleOffset];
    for (int r = 0; r < N; r++)
        for (int c = 0; c < N; c++)
        {
            const int y = hor ? c : r, x = hor ? r : c;         /* position in the un-flipped (vertical) frame */
            int v;
            if (!angle)
            {
                v = intra_nb(s, N, hor, 1 + x);
                if (bFilter && x == 0)
                {
                    v = (int16_t)(intra_nb(s, N, hor, 1) + ((intra_nb(s, N, hor, 2 * N + 1 + y) - intra_nb(s, N, hor, 0)) >> 1));
                    v = v < 0 ? 0 : v > PIXEL_MAX ? PIXEL_MAX : v;
                }
            }
            else
            {
                const int sum = (y + 1) * angle, off = sum >> 5, frac = sum & 31;
                v = intra_ref(s, N, hor, angleOffset, off + x);
                if (frac) v = ((32 - frac) * v + frac * intra_ref(s, N, hor, angleOffset, off + x + 1) + 16) >> 5;
            }
            dst[r * ds + c] = (pixel)v;
        }
}

/* One lowres CU of LookaheadTLD::lowresIntraEstimate (slicetype.cpp:781-841): neighbours from the plane itself, DC, planar,
 * then the angular modes coarse to fine (5, 10, .. 30; best +-2; best +-1), SATD cost, first-best on ties.
 * Returns icost (with the signalling penalty added) and writes the mode. */
EXPORT int orc_lowres_intra_cu(const pixel* plane, intptr_t stride, int cuX, int cuY, int penalty, int32_t* modeOut)
{
    enum { N = 8 };
    pixel nb[2][4 * N + 1], pred[N * N];
    const pixel* cur = plane + (intptr_t)N * cuY * stride + N * cuX;
    const pixel* p = cur - stride - 1;
    memcpy(nb[0], p, (2 * N + 1) * sizeof(pixel));
    for (int i = 1; i <= 2 * N; i++) nb[0][2 * N + i] = p[i * stride];
    orc_intra_filter(N, nb[0], nb[1]);
    int icost = 1 << 28, imode = 0, cost;                       /* MotionEstimate::COST_MAX, motion.h:68 */
    orc_intra_pred(N, 1, nb[0], 1, pred, N); cost = orc_satd(N, N, cur, stride, pred, N);
    if (cost < icost) { icost = cost; imode = 1; }
    orc_intra_pred(N, 0, nb[1], 0, pred, N); cost = orc_satd(N, N, cur, stride, pred, N);
    if (cost < icost) { icost = cost; imode = 0; }
    int acost = 1 << 28, amode = 4;
#define ORC_TRY_ANG(m) do { int m_ = (m); orc_intra_pred(N, m_, nb[!!(k_intra_filter_flags[m_] & N)], 1, pred, N); \
                            cost = orc_satd(N, N, cur, stride, pred, N); if (cost < acost) { acost = cost; amode = m_; } } while (0)
    for (int m = 5; m < 35; m += 5) ORC_TRY_ANG(m);
    for (int dist = 2; dist >= 1; dist--)
    {
        const int minus = amode - dist, plus = amode + dist;    /* both around the best BEFORE this round */
        ORC_TRY_ANG(minus);
        ORC_TRY_ANG(plus);
    }
#undef ORC_TRY_ANG
    if (acost < icost) { icost = acost; imode = amode; }
    *modeOut = imode;
    return icost + penalty;
}
EXPORT void orc_lowres_intra_frame(const pixel* plane, intptr_t stride, int widthInCU, int heightInCU, int penalty, int32_t* cost, int32_t* mode)
{
    for (int y = 0; y < heightInCU; y++)
        for (int x = 0; x < widthInCU; x++)
            cost[y * widthInCU + x] = orc_lowres_intra_cu(plane, stride, x, y, penalty, mode + y * widthInCU + x);
}

/* All 35 luma predictions of one TU the way the analysis forms them (encoder/search.cpp:1703-1727): DC from the unfiltered
 * neighbours with edge smoothing for N <= 16, planar from the smoothed ones for N >= 8, the angular modes from smoothed or
 * unfiltered neighbours per g_intraFilterFlags with edge filtering for N <= 16.  dst: 35 blocks of N x N, mode-major,
 * every mode in picture orientation (the reference's all-angles slot leaves modes < 18 transposed and compares them with a
 * transposed fenc; the Hadamard costs are the same). */
EXPORT void orc_intra_pred_all(int N, const pixel* s, pixel* dst)
{
    pixel f[4 * 32 + 1];
    orc_intra_filter(N, s, f);
    orc_intra_pred(N, 1, s, N <= 16, dst + N * N, N);
    orc_intra_pred(N, 0, N >= 8 ? f : s, 0, dst, N);
    for (int m = 2; m < 35; m++)
        orc_intra_pred(N, m, (k_intra_filter_flags[m] & N) ? f : s, N <= 16, dst + m * N * N, N);
}

/* SEA through the whole motionEstimate: integral = twelve planes as orc_me_integral writes them, each addressed at the PU's
 * co-located block (plane k of the padded reference picture + the block's element offset) */
EXPORT int orc_motion_estimate_sea(int merange, int subme, int w, int h, const pixel* fenc, intptr_t sf, const pixel* fref, intptr_t sr,
                                   const uint32_t* sums, size_t planePitch, intptr_t blockOffset, const int32_t* range, const int32_t* qmvp,
                                   int numCand, const int32_t* mvc, const uint16_t* costTab, int32_t* outQMv)
{
    const uint32_t* planes[12];
    for (int k = 0; k < 12; k++) planes[k] = sums + k * planePitch + blockOffset;
    me_ctx c = { w, h, fenc, sf, fref, sr, costTab - qmvp[0], costTab - qmvp[1], { 0, 0, 0, 0 }, 0, 0, 0, { 0, 0 }, 0, { 0, 0 }, 0,
                 planes, costTab, qmvp };
    return me_estimate(c, 4, merange, subme, range, qmvp, numCand, mvc, costTab, outQMv);
}

/* ------------------------------------------------------------------ lookahead cost assembly (encoder/slicetype.cpp:4520-4620)
 * lowresMC (common/lowres.h:74-93): the block a quarter-pel vector addresses in a lowres reference = the half-pel plane itself when the
 * vector is half / full pel, else the rounded average of the two nearest planes.  planes = the four half-pel planes, `pitch` apart, each
 * addressed at the CU's co-located block. */
static void orc_lowres_mc(const pixel* planes, intptr_t stride, size_t pitch, int qx, int qy, pixel* out /* 8 x 8, stride 8 */)
{
    const pixel* a = planes + (size_t)((qy & 2) | ((qx & 2) >> 1)) * pitch + (qx >> 2) + (intptr_t)(qy >> 2) * stride;
    if ((qx | qy) & 1)
    {
        const int bx = qx + (qx & 1), by = qy + (qy & 1);
        const pixel* b = planes + (size_t)((by & 2) | ((bx & 2) >> 1)) * pitch + (bx >> 2) + (intptr_t)(by >> 2) * stride;
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) out[y * 8 + x] = (pixel)((a[y * stride + x] + b[y * stride + x] + 1) >> 1);
    }
    else
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) out[y * 8 + x] = a[y * stride + x];
}

/* predictor selection (slicetype.cpp:4520-4558): every neighbour vector is costed by the 8x8 SATD of its motion-compensated block; the
 * cheapest (first on ties, COPY2_IF_LT) becomes the predictor; while the running predictor is the zero vector and the frame is a B frame,
 * skipCost follows the cost of the candidate just measured (the reference's own rule, kept as it is).  numc == 0: mvp = 0, nothing measured.
 * out: mvp[2], mvpCost (COST_MAX = 1 << 28 when numc == 0), skipCost (INT_MAX when never set). */
EXPORT void orc_lowres_mvp(const pixel* fenc, intptr_t sf, const pixel* planes, intptr_t sr, size_t pitch, const int32_t* mvc, int numc, int bBidir,
                           int32_t* out /* mvpx, mvpy, mvpCost, skipCost */)
{
    int mvpx = 0, mvpy = 0, mvpcost = 1 << 28, skip = 0x7fffffff;
    for (int i = 0; i < numc; i++)
    {
        pixel buf[64];
        orc_lowres_mc(planes, sr, pitch, mvc[2 * i], mvc[2 * i + 1], buf);
        int cost = orc_satd(8, 8, fenc, sf, buf, 8);
        if (cost < mvpcost) { mvpcost = cost; mvpx = mvc[2 * i]; mvpy = mvc[2 * i + 1]; }
        if (!(mvpx | mvpy) && bBidir) skip = cost;
    }
    out[0] = mvpx; out[1] = mvpy; out[2] = mvpcost; out[3] = skip;
}

/* bi-directional candidates of a B-frame CU (slicetype.cpp:4577-4596): SATD against the average of both lists' motion-compensated blocks,
 * and against the average of the two co-located full-pel blocks.  out: bidir cost, co-located cost. */
EXPORT void orc_lowres_bidir(const pixel* fenc, intptr_t sf, const pixel* planes0, intptr_t s0, size_t pitch0, const pixel* planes1, intptr_t s1,
                             size_t pitch1, const int32_t* mv0, const int32_t* mv1, int32_t* out)
{
    pixel a[64], b[64], r[64];
    orc_lowres_mc(planes0, s0, pitch0, mv0[0], mv0[1], a);
    orc_lowres_mc(planes1, s1, pitch1, mv1[0], mv1[1], b);
    for (int i = 0; i < 64; i++) r[i] = (pixel)((a[i] + b[i] + 1) >> 1);
    out[0] = orc_satd(8, 8, fenc, sf, r, 8);
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) r[y * 8 + x] = (pixel)((planes0[y * s0 + x] + planes1[y * s1 + x] + 1) >> 1);
    out[1] = orc_satd(8, 8, fenc, sf, r, 8);
}
