// umma_probe.cu -- lab probe for tcgen05.mma kind::i8 on sm_100a: which shared-memory matrix-descriptor fields address the
// no-swizzle canonical layouts (K-major and MN-major), and how the accumulator comes back through tcgen05.ld.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu && tools/umma_probe
// D[M x N] (int32, TMEM) = A[M x K] (int8, smem) * B[N x K]^T (int8, smem), M = 128, K = 32, N = 32 or 128.
// Every hypothesis is run and compared with the host result; the matching ones are printed as OK.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

struct Cfg
{
    int N;              // 32 or 128
    int aMN, bMN;       // 0 = K-major operand, 1 = MN-major
    int swapA, swapB;   // 1 = exchange the LBO / SBO fields of that descriptor
    int aSigned, bSigned;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
    return d;                          // layout type 0 = no swizzle, base offset 0
}

// canonical byte offsets (no swizzle): core matrix = 8 rows x 16 bytes, 128 bytes contiguous
__host__ __device__ inline int off_kmajor(int mn, int k, int lbo, int sbo) { return (mn >> 3) * sbo + (k >> 4) * lbo + (mn & 7) * 16 + (k & 15); }
__host__ __device__ inline int off_mnmajor(int mn, int k, int lbo, int sbo) { return (mn >> 4) * sbo + (k >> 3) * lbo + (k & 7) * 16 + (mn & 15); }

__global__ void __launch_bounds__(128) probe(Cfg c, const int8_t* A, const int8_t* B, int32_t* D)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmemBase;
    const int M = 128, K = 32, N = c.N;
    uint8_t* sA = smem;                 // 4096 bytes
    uint8_t* sB = smem + 4096;          // N * 32 bytes
    // K-major: LBO = 128 (next 16-byte K chunk), SBO = 256 (next 8-row group).  MN-major: LBO = distance between 8-row K groups,
    // SBO = distance between 16-element MN chunks.
    const int aL = c.aMN ? (M / 16) * 128 : 128, aS = c.aMN ? 128 : 256;
    const int bL = c.bMN ? (N / 16) * 128 : 128, bS = c.bMN ? 128 : 256;
    for (int i = threadIdx.x; i < M * K; i += 128)
    {
        int m = i / K, k = i % K;
        sA[c.aMN ? off_mnmajor(m, k, aL, aS) : off_kmajor(m, k, aL, aS)] = (uint8_t)A[i];
    }
    for (int i = threadIdx.x; i < N * K; i += 128)
    {
        int n = i / K, k = i % K;
        sB[c.bMN ? off_mnmajor(n, k, bL, bS) : off_kmajor(n, k, bL, bS)] = (uint8_t)B[i];
    }
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;");          // generic-proxy smem writes -> visible to the tensor core's async proxy
    __syncthreads();
    if (threadIdx.x < 32)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" :: "r"(smem_u32(&tmemBase)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmemBase;
    if (threadIdx.x == 0)
    {
        uint64_t da = c.swapA ? make_desc(smem_u32(sA), aS, aL) : make_desc(smem_u32(sA), aL, aS);
        uint64_t db = c.swapB ? make_desc(smem_u32(sB), bS, bL) : make_desc(smem_u32(sB), bL, bS);
        uint32_t idesc = (2u << 4) | ((uint32_t)c.aSigned << 7) | ((uint32_t)c.bSigned << 10) | ((uint32_t)c.aMN << 15) | ((uint32_t)c.bMN << 16) |
                         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                     :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0));
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    // wait for the MMA
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    const int warp = threadIdx.x >> 5;
    for (int c0 = 0; c0 < N; c0 += 32)
    {
        uint32_t r[32];
        uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                       "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                       "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int j = 0; j < 32; j++) D[(size_t)threadIdx.x * N + c0 + j] = (int32_t)r[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(tmem));
}

int main()
{
    const int M = 128, K = 32;
    int8_t hA[M * K], hB[128 * K];
    srand(265);
    for (int i = 0; i < M * K; i++) hA[i] = (int8_t)(rand() % 255 - 127);
    for (int i = 0; i < 128 * K; i++) hB[i] = (int8_t)(rand() % 181 - 90);
    int8_t *dA, *dB; int32_t* dD;
    CK(cudaMalloc(&dA, sizeof(hA))); CK(cudaMalloc(&dB, sizeof(hB))); CK(cudaMalloc(&dD, M * 128 * 4));
    CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
    static int32_t hD[M * 128];
    int found = 0;
    for (int N : { 32, 128 })
        for (int aMN = 0; aMN < 2; aMN++)
            for (int bMN = 0; bMN < 2; bMN++)
                for (int swapA = 0; swapA < 2; swapA++)
                    for (int swapB = 0; swapB < 2; swapB++)
                        for (int sgn = 0; sgn < 2; sgn++)      // sgn = 1: A unsigned (the low byte of a hi / lo split)
                        {
                            Cfg c = { N, aMN, bMN, swapA, swapB, sgn ? 0 : 1, 1 };
                            CK(cudaMemset(dD, 0xff, M * 128 * 4));
                            probe<<<1, 128, 4096 + 128 * 32 + 1024>>>(c, dA, dB, dD);
                            cudaError_t e = cudaDeviceSynchronize();
                            if (e != cudaSuccess) { printf("N=%d aMN=%d bMN=%d swapA=%d swapB=%d uA=%d: %s\n", N, aMN, bMN, swapA, swapB, sgn, cudaGetErrorString(e)); return 2; }
                            CK(cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost));
                            long bad = 0;
                            for (int m = 0; m < M; m++)
                                for (int n = 0; n < N; n++)
                                {
                                    int s = 0;
                                    for (int k = 0; k < K; k++) s += (sgn ? (int)(uint8_t)hA[m * K + k] : (int)hA[m * K + k]) * (int)hB[n * K + k];
                                    bad += s != hD[m * N + n];
                                }
                            printf("N=%3d A %s-major B %s-major swapA=%d swapB=%d A %s: %s (%ld wrong)\n", N, aMN ? "MN" : "K", bMN ? "MN" : "K", swapA, swapB,
                                   sgn ? "u8" : "s8", bad ? "--" : "OK", bad);
                            found += !bad;
                        }
    printf("%d matching configurations\n", found);
    return 0;
}
