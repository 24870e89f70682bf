/* frame_job_example.c -- a plain C caller of the host-buffer layer (include/x265b200.h): host pictures in, host results out.
 *
 * What a maintainer's integration inside the encoder looks like, reduced to one file: pinned picture buffers with PicYuv's
 * geometry (reference common/picyuv.cpp:86-118), one job with the SATD passes of a few PU shapes and a DCT+quant pass
 * registered once (descriptors from the CTU grid plus one motion vector per block, as ThreadedME's row tasks have them,
 * reference encoder/threadedme.cpp:207-261), frames submitted back to back with three in flight.  The results are checked
 * against the one-block host slots of the same library (x265b200_satd, x265b200_sub_ps / x265b200_dct / x265b200_quant), which
 * the reference's own TestBench verifies -- so this file needs neither CUDA headers nor the oracle.
 *
 *   cc -std=c99 -O2 -Iinclude tools/frame_job_example.c -Lx265-mod-by-patman_b200/lib -lx265b200 -o frame_job_example
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "x265b200.h"

#define W 416
#define H 240
#define CTU 64
#define FRAMES 7
#define SLOTS 3

static uint32_t rnd(uint32_t* s) { *s = *s * 1664525u + 1013904223u; return *s >> 8; }

#define CHECK(call) do { int rc_ = (call); if (rc_ < 0) { fprintf(stderr, "%s failed: %d (%s)\n", #call, rc_, x265b200_last_error(ctx)); return 1; } } while (0)

int main(void)
{
    x265b200_ctx* ctx = NULL;
    if (x265b200_open(0, 10, &ctx) != X265B200_OK) { fprintf(stderr, "no usable sm_100 device (there is no CPU fallback)\n"); return 2; }

    x265b200_plane *fenc[SLOTS], *ref[SLOTS];
    intptr_t stride; int rows; int32_t origin; size_t elems;
    for (int k = 0; k < SLOTS; k++)
    {
        CHECK(x265b200_plane_create(ctx, W, H, CTU, 0, 0, &fenc[k]));
        CHECK(x265b200_plane_create(ctx, W, H, CTU, 0, 0, &ref[k]));
    }
    x265b200_plane_info(fenc[0], &stride, &rows, &origin, &elems, NULL);

    /* host pictures: pinned, padded like PicYuv::m_picBuf */
    uint16_t* hF[FRAMES]; uint16_t* hR[FRAMES];
    uint32_t seed = 265;
    for (int f = 0; f < FRAMES; f++)
    {
        hF[f] = (uint16_t*)x265b200_host_alloc(ctx, elems * 2);
        hR[f] = (uint16_t*)x265b200_host_alloc(ctx, elems * 2);
        if (!hF[f] || !hR[f]) return 1;
        for (size_t i = 0; i < elems; i++)
        {
            int base = 400 + (int)((i % (size_t)stride) / 3) + (int)(i / (size_t)stride);
            hF[f][i] = (uint16_t)((base + (int)(rnd(&seed) & 31)) & 1023);
            hR[f][i] = (uint16_t)((base + (int)(rnd(&seed) & 31) + f) & 1023);
        }
    }

    /* descriptors: every block of the CTU-aligned picture, one motion vector each */
    const int cw = (W + CTU - 1) / CTU * CTU, ch = (H + CTU - 1) / CTU * CTU;
    const int shapes[3][2] = { { 16, 16 }, { 8, 4 }, { 32, 32 } };
    int32_t *offF[4], *offR[4]; int n[4];
    x265b200_frame_job* job = NULL;
    CHECK(x265b200_frame_job_create(ctx, W, H, CTU, SLOTS, &job));
    for (int s = 0; s < 4; s++)
    {
        int bw = s < 3 ? shapes[s][0] : 8, bh = s < 3 ? shapes[s][1] : 8;
        n[s] = (cw / bw) * (ch / bh);
        offF[s] = (int32_t*)malloc(sizeof(int32_t) * n[s]); offR[s] = (int32_t*)malloc(sizeof(int32_t) * n[s]);
        int i = 0;
        for (int y = 0; y < ch; y += bh)
            for (int x = 0; x < cw; x += bw, i++)
            {
                int mvx = (int)(rnd(&seed) % 41) - 20, mvy = (int)(rnd(&seed) % 41) - 20;
                offF[s][i] = origin + y * (int32_t)stride + x;
                offR[s][i] = origin + (y + mvy) * (int32_t)stride + x + mvx;
            }
    }
    for (int s = 0; s < 3; s++) CHECK(x265b200_frame_job_add_cmp(job, X265B200_SATD, shapes[s][0], shapes[s][1], offF[s], offR[s], n[s]));
    /* 8x8 TUs at QP 30: flat table 26214 (qp % 6 == 0), qBits = 14 + 5 + (15 - 10 - 3), rounding 171/512 (quant.cpp:465-466) */
    int32_t qc[64];
    for (int i = 0; i < 64; i++) qc[i] = 26214;
    const int qBits = 14 + 5 + 2, add = 171 << (qBits - 9);
    CHECK(x265b200_frame_job_add_transform(job, X265B200_PASS_LEVELS, 8, offF[3], offR[3], n[3], qc, qBits, add));

    x265b200_pass_result res[4];
    int slotOf[FRAMES];
    long checked = 0, levels = 0;
    for (int f = 0; f < FRAMES + SLOTS; f++)
    {
        if (f >= SLOTS)
        {
            int g = f - SLOTS;                                  /* oldest frame in flight */
            CHECK(x265b200_frame_job_wait(job, slotOf[g], res, 4));
            for (int s = 0; s < 3; s++)
                for (int i = 0; i < n[s]; i += 7)               /* a sample of the blocks through the one-block slot */
                {
                    int want = x265b200_satd(ctx, shapes[s][0], shapes[s][1], hF[g] + offF[s][i], stride, hR[g] + offR[s][i], stride);
                    if (res[s].cost[i] != want) { fprintf(stderr, "frame %d shape %d block %d: %d != %d\n", g, s, i, res[s].cost[i], want); return 1; }
                    checked++;
                }
            uint32_t pos = 0;
            for (int i = 0; i < n[3]; i++)
            {
                int16_t resi[64], coef[64], q[64];
                int32_t deltaU[64];
                uint32_t ns = res[3].numSig[i];
                if (i % 5 == 0)
                {
                    x265b200_sub_ps(ctx, 8, 8, resi, 8, hF[g] + offF[3][i], hR[g] + offR[3][i], stride, stride);
                    x265b200_dct(ctx, X265B200_TR_DCT, 8, resi, coef, 8);
                    uint32_t wantNs = x265b200_quant(ctx, coef, qc, deltaU, q, qBits, add, 64);
                    if (wantNs != ns) { fprintf(stderr, "frame %d TU %d: numSig %u != %u\n", g, i, ns, wantNs); return 1; }
                    uint32_t p = pos;
                    for (int c = 0; c < 64; c++)
                    {
                        uint32_t flat = (uint32_t)i * 64 + c;
                        int sig = (res[3].sigMap[flat >> 5] >> (flat & 31)) & 1;
                        int16_t got = sig ? res[3].levels[p++] : 0;
                        if (got != q[c]) { fprintf(stderr, "frame %d TU %d coef %d: %d != %d\n", g, i, c, got, q[c]); return 1; }
                    }
                    checked++;
                }
                pos += ns;
            }
            if (pos != res[3].nlevels) { fprintf(stderr, "level stream length %u != %u\n", pos, res[3].nlevels); return 1; }
            levels += pos;
        }
        if (f < FRAMES)
        {
            int k = f % SLOTS;
            CHECK(x265b200_plane_upload_padded(fenc[k], hF[f]));
            CHECK(x265b200_plane_upload_padded(ref[k], hR[f]));
            slotOf[f] = x265b200_frame_job_submit(job, fenc[k], ref[k]);
            CHECK(slotOf[f]);
        }
    }
    uint64_t h2d, d2h;
    x265b200_transfer_stats(ctx, &h2d, &d2h);
    if (x265b200_status(ctx) != X265B200_OK) { fprintf(stderr, "sticky error: %s\n", x265b200_last_error(ctx)); return 1; }
    printf("frame job example ok: %d frames of %dx%d, %ld results checked against the one-block slots, %ld non-zero levels, "
           "%.1f MB up, %.1f MB down, %llu kernel launches\n", FRAMES, W, H, checked, levels, h2d / 1e6, d2h / 1e6,
           (unsigned long long)x265b200_launch_count(ctx));
    x265b200_frame_job_destroy(job);
    for (int k = 0; k < SLOTS; k++) { x265b200_plane_destroy(fenc[k]); x265b200_plane_destroy(ref[k]); }
    for (int f = 0; f < FRAMES; f++) { x265b200_host_free(ctx, hF[f]); x265b200_host_free(ctx, hR[f]); }
    for (int s = 0; s < 4; s++) { free(offF[s]); free(offR[s]); }
    x265b200_close(ctx);
    return 0;
}
