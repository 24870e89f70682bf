// satd_lab.cu -- tuning lab for the SATD throughput kernel: instantiates variants of
// csrc/tile_kernels.cuh side by side on 2160p10-shaped planes, checks that they agree, prints ms.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I x265-mod-by-patman_b200/csrc tools/satd_lab.cu -o tools/satd_lab
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "tile_kernels.cuh"
#include "tma_cmp.cuh"

using namespace b200;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static uint64_t splitmix(uint64_t z) { z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

__global__ void fill(uint16_t* p, size_t n, uint64_t seed)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t z = i + seed; z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    p[i] = (uint16_t)(z & 1023);
}

typedef void (*kern_t)(const uint16_t*, intptr_t, const uint16_t*, intptr_t, const int32_t*, const int32_t*, int, int, int, int, int, int32_t*);

struct Variant { const char* name; kern_t k; int threads; int gdiv; };

int main(int argc, char** argv)
{
    const int F = argc > 1 ? atoi(argv[1]) : 8, W = 3840, H = 2176, stride = 4032, rows = 2336, mx = 96, my = 80;
    const size_t pe = (size_t)stride * rows;
    uint16_t *A, *B;
    CK(cudaMalloc(&A, F * pe * 2)); CK(cudaMalloc(&B, F * pe * 2));
    fill<<<(F * pe + 255) / 256, 256>>>(A, F * pe, 1); fill<<<(F * pe + 255) / 256, 256>>>(B, F * pe, 77);
    CK(cudaDeviceSynchronize());
    Variant vars[] = {
        { "t4 u1 128 thr (library rule)", tile4_fast_kernel<uint16_t, OP_SATD, int, int32_t, 1, 1>, 128, 0 },
        { "t4 u1 256 thr", tile4_fast_kernel<uint16_t, OP_SATD, int, int32_t, 1, 1>, 256, 0 },
        { "t4 u1 128 thr, full-warp groups", tile4_fast_kernel<uint16_t, OP_SATD, int, int32_t, 1, 1>, 128, 1 },
        { "t4 u2 128 thr (2 tiles loaded together)", tile4_fast_kernel<uint16_t, OP_SATD, int, int32_t, 2, 1>, 128, 0 },
    };
    const int NV = sizeof(vars) / sizeof(vars[0]);
    for (int v = 0; v < NV; v++)
    {
        cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, (const void*)vars[v].k));
        printf("variant %d: %-28s regs %d\n", v, vars[v].name, fa.numRegs);
    }
    int shapes[][2] = { {64, 64}, {32, 32}, {16, 32}, {16, 16}, {16, 8}, {8, 16}, {8, 8}, {8, 4}, {4, 8} };
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (auto& sh : shapes)
    {
        int w = sh[0], h = sh[1];
        std::vector<int32_t> oa, ob;
        for (int f = 0; f < F; f++)
            for (int y = 0; y < H; y += h)
                for (int x = 0; x < W; x += w)
                {
                    uint64_t r = splitmix(oa.size() * 31 + w * 7 + h);
                    int mvx = (int)(r % 115) - 57, mvy = (int)((r >> 20) % 115) - 57;
                    int rx = x + mvx, ry = y + mvy;
                    if (argc > 3) rx &= ~7;                      // lab: 16-byte aligned reference blocks, the only ones a TMA box can start at
                    if (rx < -mx + 8) rx = -mx + 8; if (rx > W + mx - w - 8) rx = W + mx - w - 8;
                    if (ry < -my + 8) ry = -my + 8; if (ry > H + my - h - 8) ry = H + my - h - 8;
                    oa.push_back((int32_t)(f * pe + (size_t)(my + y) * stride + mx + x));
                    ob.push_back((int32_t)(f * pe + (size_t)(my + ry) * stride + mx + rx));
                }
        int n = (int)oa.size();
        int32_t *dOA, *dOB, *out0, *out;
        CK(cudaMalloc(&dOA, n * 4)); CK(cudaMalloc(&dOB, n * 4)); CK(cudaMalloc(&out0, n * 4)); CK(cudaMalloc(&out, n * 4));
        CK(cudaMemcpy(dOA, oa.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dOB, ob.data(), n * 4, cudaMemcpyHostToDevice));
        int T4 = (w / 4) * (h / 4), G0 = 1; while (G0 < T4 && G0 < 32) G0 <<= 1;
        double bytes = (double)n * w * h * 4 + n * 4.0;
        int G = G0;
        printf("shape %dx%d  n=%d  G=%d  (roofline 6534.8 GB/s -> %.4f ms)\n", w, h, n, G, bytes / 6534.8e9 * 1e3);
        std::vector<int32_t> h0(n), h1(n);
        for (int v = 0; v < NV; v++)
        {
            int32_t* o = v == 0 ? out0 : out;
            if (vars[v].gdiv == 0)
            {   // library rule for 4x4 tiles
                int per = T4 >= 16 ? 4 : 2; G = 1; while (G * 2 * per <= T4 && G < 32) G <<= 1;
            }
            else { G = G0 / vars[v].gdiv; if (G < 1) G = 1; }
            long long threads = (long long)n * G;
            int grid = (int)((threads + vars[v].threads - 1) / vars[v].threads);
            float best = 1e9f;
            for (int rep = 0; rep < 3; rep++)
            {   // 6 launches back to back: the queue stays full, so launch latency is not in the number
                vars[v].k<<<grid, vars[v].threads>>>(A, stride, B, stride, dOA, dOB, 1, n, w, h, G, o);
                CK(cudaEventRecord(e0));
                for (int k = 0; k < 6; k++) vars[v].k<<<grid, vars[v].threads>>>(A, stride, B, stride, dOA, dOB, 1, n, w, h, G, o);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                ms /= 6;
                if (ms < best) best = ms;
            }
            CK(cudaGetLastError());
            bool same = true;
            if (v > 0)
            {
                CK(cudaMemcpy(h0.data(), out0, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h1.data(), out, n * 4, cudaMemcpyDeviceToHost));
                for (int i = 0; i < n; i++) if (h0[i] != h1[i]) { same = false; break; }
            }
            printf("   v%d %-28s %.4f ms  %.0f GB/s  %.2f of roofline %s\n", v, vars[v].name, best, bytes / best / 1e6, bytes / best / 1e6 / 6534.8, same ? "" : "MISMATCH");
        }

        // 8x4-strip variants with 16-byte chunk loads (narrow blocks)
        if (w <= 16)
        {
            int S = (w / 8) * (h / 4);
            for (int Gs = 1; Gs <= S && Gs <= 32; Gs <<= 1)
            {
                long long threads = (long long)n * Gs;
                int grid = (int)((threads + 127) / 128);
                float best = 1e9f;
                for (int rep = 0; rep < 3; rep++)
                {
                    strip8_fast_kernel<OP_SATD, int, int32_t><<<grid, 128>>>(A, stride, B, stride, dOA, dOB, 1, n, w, h, Gs, out);
                    CK(cudaEventRecord(e0));
                    for (int k = 0; k < 6; k++) strip8_fast_kernel<OP_SATD, int, int32_t><<<grid, 128>>>(A, stride, B, stride, dOA, dOB, 1, n, w, h, Gs, out);
                    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    ms /= 6;
                    if (ms < best) best = ms;
                }
                CK(cudaGetLastError());
                CK(cudaMemcpy(h0.data(), out0, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h1.data(), out, n * 4, cudaMemcpyDeviceToHost));
                int bad = 0;
                for (int i = 0; i < n; i++) if (h0[i] != h1[i]) bad++;
                printf("   strip8x4 v16 loads, G = %2d lanes/block  %.4f ms  %.0f GB/s  %.2f of roofline %s\n", Gs, best, bytes / best / 1e6, bytes / best / 1e6 / 6534.8, bad ? "MISMATCH" : "");
            }
        }

        // TMA-staged variants: jobBytes per plane per stage, stages, warps per CTA, L2 promotion
        struct TC { int job, stages, warps, promo; };
        TC tcs[] = { {2048, 4, 8, 0}, {2048, 4, 16, 0}, {4096, 3, 8, 0}, {1024, 6, 16, 0}, {2048, 6, 8, 0}, {2048, 4, 8, 2}, {4096, 4, 12, 0}, {2048, 3, 16, 0} };
        for (auto& tc : tcs)
        {
            if (argc < 3) break;       // TMA variants only on request: unaligned box origins fault (see profiles/r1_tma_probe_notes.md)
            cudaError_t err;
            CK(cudaMemset(out, 0xff, n * 4));
            bool ok = launch_cmp_tma<uint16_t, OP_SATD, int, int32_t>(A, stride, B, stride, dOA, dOB, 1, n, w, h, out, 0, 148, tc.job, tc.stages, tc.warps, tc.promo, &err);
            if (!ok) { printf("   tma job %d S %d W %d promo %d: not applicable (%s)\n", tc.job, tc.stages, tc.warps, tc.promo, cudaGetErrorString(err)); continue; }
            CK(cudaDeviceSynchronize());
            float best = 1e9f;
            for (int rep = 0; rep < 3; rep++)
            {
                CK(cudaEventRecord(e0));
                for (int k = 0; k < 6; k++) launch_cmp_tma<uint16_t, OP_SATD, int, int32_t>(A, stride, B, stride, dOA, dOB, 1, n, w, h, out, 0, 148, tc.job, tc.stages, tc.warps, tc.promo, &err);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                ms /= 6;
                if (ms < best) best = ms;
            }
            CK(cudaGetLastError());
            CK(cudaMemcpy(h0.data(), out0, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h1.data(), out, n * 4, cudaMemcpyDeviceToHost));
            int bad = 0, first = -1;
            for (int i = 0; i < n; i++) if (h0[i] != h1[i]) { if (first < 0) first = i; bad++; }
            printf("   tma job %4d S %d W %2d promo %d        %.4f ms  %.0f GB/s  %.2f of roofline", tc.job, tc.stages, tc.warps, tc.promo, best, bytes / best / 1e6, bytes / best / 1e6 / 6534.8);
            if (bad) printf("  MISMATCH x%d (first %d: %d vs %d)", bad, first, h0[first], h1[first]);
            printf("\n");
        }
        cudaFree(dOA); cudaFree(dOB); cudaFree(out0); cudaFree(out);
    }
    return 0;
}
