// mma_probe.cu -- measures issue throughput of the legacy warp-level tensor instructions on sm_100a
// (IMMA.16832 s8, HMMA.16816 bf16) to size the tensor-core DCT.  Prints MMA/clk/SM and dense TOPS.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template<int KIND, int CHAINS>
__global__ void __launch_bounds__(256) probe(int iters, int* sink)
{
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = threadIdx.x ^ 5, b1 = 11;
    int c[CHAINS][4];
    float f[CHAINS][4];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) for (int j = 0; j < 4; j++) { c[i][j] = 0; f[i][j] = 0.f; }
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int i = 0; i < CHAINS; i++)
        {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                    : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                    : "+f"(f[i][0]), "+f"(f[i][1]), "+f"(f[i][2]), "+f"(f[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                    : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3]) : "r"(a0), "r"(a1), "r"(b0));
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) for (int j = 0; j < 4; j++) s += c[i][j] + (int)f[i][j];
    if (s == 0x7fffffff) sink[0] = s;
}

template<int KIND> int run(const char* name, double macs_per_mma, int* sink, int sms, double clk_ghz)
{
    const int iters = 4096, CH = 8;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int ctas = 1; ctas <= 4; ctas *= 2)
    {
        float best = 1e9f;
        for (int rep = 0; rep < 4; rep++)
        {
            CK(cudaEventRecord(e0));
            probe<KIND, CH><<<sms * ctas, 256>>>(iters, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
        }
        double mmas = (double)sms * ctas * 8 * iters * CH;
        printf("%-28s %d CTA/SM (8 warps each): %.3f ms  %.1f G mma/s  %.3f mma/clk/SM @%.2f GHz  %.0f dense TOPS\n",
               name, ctas, best, mmas / best / 1e6, mmas / (best * 1e-3) / sms / (clk_ghz * 1e9), clk_ghz, 2 * macs_per_mma * mmas / best / 1e9);
    }
    return 0;
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int* sink; CK(cudaMalloc(&sink, 4));
    double clk = p.clockRate / 1e6;
    printf("%s, %d SMs, %.2f GHz max\n", p.name, p.multiProcessorCount, clk);
    run<0>("IMMA m16n8k32 s8.u8", 16.0 * 8 * 32, sink, p.multiProcessorCount, clk);
    run<2>("IMMA m16n8k16 s8.u8", 16.0 * 8 * 16, sink, p.multiProcessorCount, clk);
    run<1>("HMMA m16n8k16 bf16", 16.0 * 8 * 16, sink, p.multiProcessorCount, clk);
    return 0;
}
