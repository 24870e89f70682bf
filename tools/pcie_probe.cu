// pcie_probe.cu -- host <-> device copy rates of this box: each direction alone and both at once, pinned memory,
// chunk sizes of the e2e path (one padded 2160p10 plane = 18.8 MB).  Explains bench.py's e2e.per_rank_GBps.
//   nvcc -O2 -o tools/pcie_probe tools/pcie_probe.cu && tools/pcie_probe [device]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main(int argc, char** argv)
{
    int dev = argc > 1 ? atoi(argv[1]) : 0;
    CK(cudaSetDevice(dev));
    const size_t chunk = 18837504, n = 16;
    char *h0, *h1, *d0, *d1;
    CK(cudaHostAlloc((void**)&h0, chunk * n, cudaHostAllocPortable)); CK(cudaHostAlloc((void**)&h1, chunk * n, cudaHostAllocPortable));
    CK(cudaMalloc((void**)&d0, chunk * n)); CK(cudaMalloc((void**)&d1, chunk * n));
    for (size_t i = 0; i < chunk * n; i += 4096) { h0[i] = 1; h1[i] = 2; }
    cudaStream_t s0, s1; CK(cudaStreamCreate(&s0)); CK(cudaStreamCreate(&s1));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int mode = 0; mode < 3; mode++)
        for (int rep = 0; rep < 3; rep++)
        {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(a, 0));
            CK(cudaStreamWaitEvent(s0, a, 0)); CK(cudaStreamWaitEvent(s1, a, 0));
            for (size_t k = 0; k < n; k++)
            {
                if (mode != 1) CK(cudaMemcpyAsync(d0 + k * chunk, h0 + k * chunk, chunk, cudaMemcpyHostToDevice, s0));
                if (mode != 0) CK(cudaMemcpyAsync(h1 + k * chunk, d1 + k * chunk, chunk, cudaMemcpyDeviceToHost, s1));
            }
            CK(cudaEventRecord(b, s0)); CK(cudaStreamWaitEvent(0, b, 0));
            CK(cudaEventRecord(b, s1)); CK(cudaStreamWaitEvent(0, b, 0));
            CK(cudaEventRecord(b, 0));
            CK(cudaEventSynchronize(b));
            float ms; CK(cudaEventElapsedTime(&ms, a, b));
            double gb = chunk * n / 1e9;
            if (rep == 2)
                printf("%s: %.1f GB/s per direction%s\n", mode == 0 ? "H2D alone" : mode == 1 ? "D2H alone" : "H2D + D2H together", gb / (ms * 1e-3),
                       mode == 2 ? " (each)" : "");
        }
    return 0;
}
