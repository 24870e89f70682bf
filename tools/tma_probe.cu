// tma_probe.cu -- isolates the TMA box-load mechanics used by tools/tma_cmp.cuh (one experiment per process).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I x265-mod-by-patman_b200/csrc tools/tma_probe.cu -o tools/tma_probe
// mode bits: 1 mbarrier_init fence, 2 proxy fence, 4 exact row count in the map, 8 descriptor in global memory,
//            16 aligned coordinates (64, 32), 32 int32 elements, 64 libcu++ barrier + cp_async_bulk_tensor wrappers, 128 static smem
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda/barrier>
#include "tma_cmp.cuh"
using namespace b200;
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void probe(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, int x, int y, int boxBytes, uint16_t* out, int mode)
{
    extern __shared__ __align__(1024) uint8_t dsmem[];
    __shared__ alignas(1024) uint8_t ssmem[8192];
    __shared__ alignas(8) uint64_t sbar;
    uint8_t* smem = (mode & 128) ? ssmem : dsmem;
    uint64_t* bar = (mode & 128) ? &sbar : (uint64_t*)(dsmem + 16384);
    const CUtensorMap* mp = (mode & 8) ? gmap : &pmap;
    int lane = threadIdx.x;
    if (mode & 64)
    {
#pragma nv_diag_suppress static_var_with_dynamic_init
        __shared__ barrier cbar;
        if (lane == 0) { init(&cbar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
        __syncthreads();
        barrier::arrival_token token;
        if (lane == 0)
        {
            cde::cp_async_bulk_tensor_2d_global_to_shared(smem, mp, x, y, cbar);
            token = cuda::device::barrier_arrive_tx(cbar, 1, boxBytes);
        }
        else token = cbar.arrive();
        cbar.wait(std::move(token));
    }
    else
    {
        if (lane == 0) mbar_init(smem_u32(bar), 1);
        if (mode & 1) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (mode & 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (lane == 0)
        {
            mbar_arrive_expect_tx(smem_u32(bar), boxBytes);
            tma_load_2d(smem_u32(smem), mp, smem_u32(bar), x, y);
        }
        __syncwarp();
        mbar_wait(smem_u32(bar), 0);
    }
    for (int i = lane; i < boxBytes / 2; i += 32) out[i] = ((uint16_t*)smem)[i];
}

int main(int argc, char** argv)
{
    int mode = argc > 1 ? atoi(argv[1]) : 1;
    int w = argc > 2 ? atoi(argv[2]) : 64, h = argc > 3 ? atoi(argv[3]) : 16;
    const int stride = 4032, rows = 2336;
    std::vector<uint16_t> hp((size_t)stride * rows);
    for (size_t i = 0; i < hp.size(); i++) hp[i] = (uint16_t)(i * 2654435761u >> 20);
    uint16_t *P, *out;
    CK(cudaMalloc(&P, hp.size() * 2)); CK(cudaMalloc(&out, 65536));
    CK(cudaMemcpy(P, hp.data(), hp.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap m;
    int eb = (mode & 32) ? 4 : 2;
    cuuint64_t gdim[2] = { (cuuint64_t)stride * 2 / eb, (mode & 4) ? (cuuint64_t)rows : (cuuint64_t)(0x80000000ull / stride + 64) };
    cuuint64_t gstr[1] = { (cuuint64_t)stride * 2 };
    cuuint32_t box[2] = { (cuuint32_t)(w * 2 / eb), (cuuint32_t)h }, est[2] = { 1, 1 };
    CUresult r = tmap_encoder()(&m, (mode & 32) ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, P, gdim, gstr, box, est,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int x = (mode & 16) ? 64 : 101, y = (mode & 16) ? 32 : 77;
    int cx = (mode & 32) ? x / 2 : x;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    CUtensorMap* gm; CK(cudaMalloc(&gm, 256)); CK(cudaMemcpy(gm, &m, sizeof(m), cudaMemcpyHostToDevice));
    probe<<<1, 32, 32768>>>(m, gm, cx, y, w * h * 2, out, mode);
    CK(cudaGetLastError());
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %3d box %dx%d encode %d: %s\n", mode, w, h, (int)r, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint16_t> ho(w * h);
    CK(cudaMemcpy(ho.data(), out, w * h * 2, cudaMemcpyDeviceToHost));
    int bad = 0;
    x = (mode & 32) ? cx * 2 : x;
    for (int rr = 0; rr < h; rr++) for (int c = 0; c < w; c++) if (ho[rr * w + c] != hp[(size_t)(y + rr) * stride + x + c]) bad++;
    printf("   mismatches %d of %d\n", bad, w * h);
    return 0;
}
