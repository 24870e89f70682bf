// issue_probe.cu -- measured per-SM issue rate of the SAD instructions the motion-search kernel is built from
// (VABSDIFF.U32 with accumulate, VABSDIFF4.U8.ACC) next to a plain IADD3 chain: the denominators for mesearch.cu's
// "fraction of issue peak".  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/issue_probe.cu -o tools/issue_probe
#include <cstdio>
#include <cuda_runtime.h>

template<int OP>
__global__ void __launch_bounds__(1024) probe(unsigned* out, unsigned seed, int iters)
{
    unsigned a[8], x = seed + threadIdx.x, y = seed * 3 + blockIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = i;
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                if (OP == 0) a[i] = __usad(x, y + i, a[i]);
                else if (OP == 1) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(x), "r"(y + i));
                else a[i] = a[i] + x + (y ^ i);
            }
        x += a[0] & 1;
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 0x12345678) out[0] = s;
}

template<int OP> static void run(const char* name, int sms, int khz)
{
    unsigned* out; cudaMalloc(&out, 4);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<OP><<<sms * 2, 1024>>>(out, 1, 100);
    cudaEventRecord(e0);
    probe<OP><<<sms * 2, 1024>>>(out, 1, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)sms * 2 * 1024 * iters * 32.0;
    printf("%-22s %.2f T lane-ops/s  = %.1f lanes/clk/SM at the max SM clock %d MHz\n", name, ops / ms / 1e9, ops / ms / sms / khz, khz / 1000);
    cudaFree(out);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    run<0>("VABSDIFF.U32 (+acc)", p.multiProcessorCount, khz);
    run<1>("VABSDIFF4.U8.ACC", p.multiProcessorCount, khz);
    run<2>("IADD3", p.multiProcessorCount, khz);
    return 0;
}
