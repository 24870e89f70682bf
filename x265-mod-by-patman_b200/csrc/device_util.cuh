// device_util.cuh -- load/reduce helpers shared by the kernels
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

// Four horizontally adjacent samples as ints.  The address is only guaranteed to be aligned to the
// element size (reference blocks sit at arbitrary motion-vector offsets), so the widest load the
// actual alignment allows is chosen at run time; the branch is warp-uniform for the usual case of
// a warp walking 4-sample tiles of one plane.  Never touches bytes outside the four samples.
__device__ __forceinline__ void load4(const uint16_t* p, int (&v)[4])
{
    uintptr_t a = (uintptr_t)p;
    if ((a & 7) == 0)
    {
        uint2 q = __ldg((const uint2*)p);
        v[0] = q.x & 0xffff; v[1] = q.x >> 16; v[2] = q.y & 0xffff; v[3] = q.y >> 16;
    }
    else if ((a & 3) == 0)
    {
        uint32_t x = __ldg((const uint32_t*)p), y = __ldg((const uint32_t*)p + 1);
        v[0] = x & 0xffff; v[1] = x >> 16; v[2] = y & 0xffff; v[3] = y >> 16;
    }
    else
    {
        uint32_t m = __ldg((const uint32_t*)(p + 1));
        v[0] = __ldg(p); v[1] = m & 0xffff; v[2] = m >> 16; v[3] = __ldg(p + 3);
    }
}

__device__ __forceinline__ void load4(const uint8_t* p, int (&v)[4])
{
    uintptr_t a = (uintptr_t)p;
    uint32_t x;
    if ((a & 3) == 0)
        x = __ldg((const uint32_t*)p);
    else if ((a & 1) == 0)
        x = (uint32_t)__ldg((const uint16_t*)p) | ((uint32_t)__ldg((const uint16_t*)p + 1) << 16);
    else
        x = (uint32_t)__ldg(p) | ((uint32_t)__ldg((const uint16_t*)(p + 1)) << 8) | ((uint32_t)__ldg(p + 3) << 24);
    v[0] = x & 0xff; v[1] = (x >> 8) & 0xff; v[2] = (x >> 16) & 0xff; v[3] = x >> 24;
}

__device__ __forceinline__ void load4(const int16_t* p, int (&v)[4])
{
    uintptr_t a = (uintptr_t)p;
    if ((a & 7) == 0)
    {
        uint2 q = __ldg((const uint2*)p);
        v[0] = (int16_t)(q.x & 0xffff); v[1] = (int32_t)q.x >> 16; v[2] = (int16_t)(q.y & 0xffff); v[3] = (int32_t)q.y >> 16;
    }
    else if ((a & 3) == 0)
    {
        uint32_t x = __ldg((const uint32_t*)p), y = __ldg((const uint32_t*)p + 1);
        v[0] = (int16_t)(x & 0xffff); v[1] = (int32_t)x >> 16; v[2] = (int16_t)(y & 0xffff); v[3] = (int32_t)y >> 16;
    }
    else
    {
        uint32_t m = __ldg((const uint32_t*)(p + 1));
        v[0] = __ldg(p); v[1] = (int16_t)(m & 0xffff); v[2] = (int32_t)m >> 16; v[3] = __ldg(p + 3);
    }
}

// butterfly sum over the low `lanes` (power of two <= 32) lanes of each aligned lane group
template<typename T>
__device__ __forceinline__ T group_sum(T v, int lanes)
{
    for (int m = lanes >> 1; m > 0; m >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

__host__ __device__ inline int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

} // namespace b200
