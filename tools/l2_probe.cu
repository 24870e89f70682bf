// l2_probe.cu -- L2 -> SM read bandwidth on this GPU for an L2-resident buffer (ld.global.cg, 16 bytes per lane, every sector fully used):
// the ceiling of kernels whose reference blocks overlap in L2 but are fetched sector by sector (the fused per-CU SATD).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/l2_probe.cu -o tools/l2_probe_bin
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) rd(const uint4* __restrict__ p, size_t n16, int reps, unsigned* sink)
{
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; r++)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
        {
            uint4 v = __ldcg(p + i);
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x12345678u) *sink = acc;
}
int main()
{
    for (size_t mb : { 16, 32, 64, 96, 512 })
    {
        size_t bytes = mb << 20;
        uint4* p; unsigned* sink;
        cudaMalloc(&p, bytes); cudaMalloc(&sink, 4); cudaMemset(p, 1, bytes);
        int reps = mb >= 512 ? 4 : 64;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        rd<<<148 * 8, 256>>>(p, bytes / 16, 2, sink);
        cudaEventRecord(e0);
        rd<<<148 * 8, 256>>>(p, bytes / 16, reps, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%4zu MB buffer: %.0f GB/s\n", mb, (double)bytes * reps / ms / 1e6);
        cudaFree(p); cudaFree(sink);
    }
    return 0;
}
