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

// ---- fast tile loads (throughput kernels) -------------------------------------------------------
// Four rows of four horizontally adjacent samples, returned as packed 16-bit pairs
// (lo[r] = samples 0,1 ; hi[r] = samples 2,3 of row r).  The row stride must be a multiple of four
// samples (every x265 plane and every staging tile is), so the misalignment of the tile against
// the 8-byte (16-bit samples) / 4-byte (8-bit samples) grid is the same for all rows: the kernel
// issues ALL aligned chunk loads of the tile back to back (no data-dependent code in between, so
// they overlap in flight) and then realigns with funnel shifts.  Only chunks that contain at least
// one requested sample are touched.
__device__ __forceinline__ void load_tile4x4(const uint16_t* p, intptr_t stride, uint32_t (&lo)[4], uint32_t (&hi)[4])
{
    uintptr_t a = (uintptr_t)p;
    int s = (int)(a >> 1) & 3;                              // misalignment in samples
    const uint2* base = (const uint2*)(a & ~(uintptr_t)7);
    intptr_t cs = stride >> 2;                              // row stride in 8-byte chunks
    uint2 q0[4], q1[4];
#pragma unroll
    for (int r = 0; r < 4; r++) q0[r] = __ldg(base + r * cs);
    if (s)
    {
#pragma unroll
        for (int r = 0; r < 4; r++) q1[r] = __ldg(base + r * cs + 1);
    }
    else
    {
#pragma unroll
        for (int r = 0; r < 4; r++) q1[r] = make_uint2(0, 0);
    }
    int sh = (s & 1) << 4;
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        uint32_t w0 = q0[r].x, w1 = q0[r].y, w2 = q1[r].x, w3 = q1[r].y;
        if (s & 2) { w0 = w1; w1 = w2; w2 = w3; }
        lo[r] = __funnelshift_r(w0, w1, sh);
        hi[r] = __funnelshift_r(w1, w2, sh);
    }
}

__device__ __forceinline__ void load_tile4x4(const uint8_t* p, intptr_t stride, uint32_t (&lo)[4], uint32_t (&hi)[4])
{
    uintptr_t a = (uintptr_t)p;
    int s = (int)a & 3;
    const uint32_t* base = (const uint32_t*)(a & ~(uintptr_t)3);
    intptr_t cs = stride >> 2;
    uint32_t c0[4], c1[4];
#pragma unroll
    for (int r = 0; r < 4; r++) c0[r] = __ldg(base + r * cs);
    if (s)
    {
#pragma unroll
        for (int r = 0; r < 4; r++) c1[r] = __ldg(base + r * cs + 1);
    }
    else
    {
#pragma unroll
        for (int r = 0; r < 4; r++) c1[r] = 0;
    }
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        uint32_t w = __funnelshift_r(c0[r], c1[r], s << 3);
        lo[r] = __byte_perm(w, 0, 0x4140);                 // bytes 0,1 -> 16-bit lanes
        hi[r] = __byte_perm(w, 0, 0x4342);                 // bytes 2,3
    }
}

// NQ consecutive quads (4 samples each) of one row starting at p (aligned to the sample size only):
// aligned 8-byte chunk loads, all issued first, then funnel-shift realignment.  16-bit samples.
template<int NQ>
__device__ __forceinline__ void load_row_quads(const uint16_t* p, uint32_t (&w)[2 * NQ])
{
    uintptr_t a = (uintptr_t)p;
    int s = (int)(a >> 1) & 3;
    const uint2* base = (const uint2*)(a & ~(uintptr_t)7);
    uint32_t c[2 * NQ + 2];
#pragma unroll
    for (int i = 0; i < NQ; i++) { uint2 q = __ldg(base + i); c[2 * i] = q.x; c[2 * i + 1] = q.y; }
    if (s) { uint2 q = __ldg(base + NQ); c[2 * NQ] = q.x; c[2 * NQ + 1] = q.y; }
    else { c[2 * NQ] = 0; c[2 * NQ + 1] = 0; }
    int sh = (s & 1) << 4;
#pragma unroll
    for (int i = 0; i < 2 * NQ; i++)
    {
        uint32_t lo = (s & 2) ? c[i + 1] : c[i];
        uint32_t hi = (s & 2) ? c[(i + 2 <= 2 * NQ + 1) ? i + 2 : 2 * NQ + 1] : c[i + 1];
        w[i] = __funnelshift_r(lo, hi, sh);
    }
}
template<int NQ>
__device__ __forceinline__ void load_row_quads(const int16_t* p, uint32_t (&w)[2 * NQ]) { load_row_quads<NQ>((const uint16_t*)p, w); }

// 8-bit samples: 4-byte chunks, result widened to packed 16-bit pairs
template<int NQ>
__device__ __forceinline__ void load_row_quads(const uint8_t* p, uint32_t (&w)[2 * NQ])
{
    uintptr_t a = (uintptr_t)p;
    int s = (int)a & 3;
    const uint32_t* base = (const uint32_t*)(a & ~(uintptr_t)3);
    uint32_t c[NQ + 1];
#pragma unroll
    for (int i = 0; i < NQ; i++) c[i] = __ldg(base + i);
    c[NQ] = s ? __ldg(base + NQ) : 0;
#pragma unroll
    for (int i = 0; i < NQ; i++)
    {
        uint32_t v = __funnelshift_r(c[i], c[i + 1], s << 3);
        w[2 * i] = __byte_perm(v, 0, 0x4140);
        w[2 * i + 1] = __byte_perm(v, 0, 0x4342);
    }
}

// ROWS rows of NQ quads each (row r starts at p + r * stride), like load_row_quads but with ONE misalignment class for
// the whole strip (stride % 4 == 0): the base pointer, the shift and the word selection are computed once, all chunk loads
// are issued before the first use, and the (s & 2) word selection is a branch around two straight-line realign blocks
// instead of per-word selects.  w[r] gets 2 * NQ packed words (+ one zero word of padding for the FIR's odd-phase shifts).
template<int NQ, int ROWS>
__device__ __forceinline__ void load_rows_quads(const uint16_t* p, intptr_t stride, uint32_t (&w)[ROWS][2 * NQ + 1])
{
    uintptr_t a = (uintptr_t)p;
    int s = (int)(a >> 1) & 3;
    const uint2* base = (const uint2*)(a & ~(uintptr_t)7);
    intptr_t cs = stride >> 2;
    uint32_t c[ROWS][2 * NQ + 2];
#pragma unroll
    for (int r = 0; r < ROWS; r++)
    {
#pragma unroll
        for (int i = 0; i < NQ; i++) { uint2 q = __ldg(base + r * cs + i); c[r][2 * i] = q.x; c[r][2 * i + 1] = q.y; }
    }
    if (s)
    {
#pragma unroll
        for (int r = 0; r < ROWS; r++) { uint2 q = __ldg(base + r * cs + NQ); c[r][2 * NQ] = q.x; c[r][2 * NQ + 1] = q.y; }
    }
    else
    {
#pragma unroll
        for (int r = 0; r < ROWS; r++) { c[r][2 * NQ] = 0; c[r][2 * NQ + 1] = 0; }
    }
    int sh = (s & 1) << 4;
    if (s & 2)
    {
#pragma unroll
        for (int r = 0; r < ROWS; r++)
        {
#pragma unroll
            for (int i = 0; i < 2 * NQ; i++) w[r][i] = __funnelshift_r(c[r][i + 1], c[r][(i + 2 <= 2 * NQ + 1) ? i + 2 : 2 * NQ + 1], sh);
            w[r][2 * NQ] = 0;
        }
    }
    else
    {
#pragma unroll
        for (int r = 0; r < ROWS; r++)
        {
#pragma unroll
            for (int i = 0; i < 2 * NQ; i++) w[r][i] = __funnelshift_r(c[r][i], c[r][i + 1], sh);
            w[r][2 * NQ] = 0;
        }
    }
}
template<int NQ, int ROWS>
__device__ __forceinline__ void load_rows_quads(const int16_t* p, intptr_t stride, uint32_t (&w)[ROWS][2 * NQ + 1])
{ load_rows_quads<NQ, ROWS>((const uint16_t*)p, stride, w); }

template<int NQ, int ROWS>
__device__ __forceinline__ void load_rows_quads(const uint8_t* p, intptr_t stride, uint32_t (&w)[ROWS][2 * NQ + 1])
{
    uintptr_t a = (uintptr_t)p;
    int s = (int)a & 3;
    const uint32_t* base = (const uint32_t*)(a & ~(uintptr_t)3);
    intptr_t cs = stride >> 2;
    uint32_t c[ROWS][NQ + 1];
#pragma unroll
    for (int r = 0; r < ROWS; r++)
    {
#pragma unroll
        for (int i = 0; i < NQ; i++) c[r][i] = __ldg(base + r * cs + i);
    }
    if (s)
    {
#pragma unroll
        for (int r = 0; r < ROWS; r++) c[r][NQ] = __ldg(base + r * cs + NQ);
    }
    else
    {
#pragma unroll
        for (int r = 0; r < ROWS; r++) c[r][NQ] = 0;
    }
#pragma unroll
    for (int r = 0; r < ROWS; r++)
    {
#pragma unroll
        for (int i = 0; i < NQ; i++)
        {
            uint32_t v = __funnelshift_r(c[r][i], c[r][i + 1], s << 3);
            w[r][2 * i] = __byte_perm(v, 0, 0x4140);
            w[r][2 * i + 1] = __byte_perm(v, 0, 0x4342);
        }
        w[r][2 * NQ] = 0;
    }
}

// ROWS rows of eight horizontally adjacent samples as packed 16-bit pairs (w[r][i] = samples 2i, 2i+1 of row r).
// The row stride must be a multiple of four samples, so every row shares one misalignment class: all aligned
// chunk loads of the strip are issued before the first use (ROWS * 16 bytes per thread in flight), then realigned.
template<int ROWS>
__device__ __forceinline__ void load_rows8(const uint16_t* p, intptr_t stride, uint32_t (&w)[ROWS][4])
{
    uintptr_t a = (uintptr_t)p;
    int s = (int)(a >> 1) & 3;
    const uint2* base = (const uint2*)(a & ~(uintptr_t)7);
    intptr_t cs = stride >> 2;
    uint2 q0[ROWS], q1[ROWS], q2[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) { q0[r] = __ldg(base + r * cs); q1[r] = __ldg(base + r * cs + 1); }
    if (s)
    {
#pragma unroll
        for (int r = 0; r < ROWS; r++) q2[r] = __ldg(base + r * cs + 2);
    }
    else
    {
#pragma unroll
        for (int r = 0; r < ROWS; r++) q2[r] = make_uint2(0, 0);
    }
    int sh = (s & 1) << 4;
#pragma unroll
    for (int r = 0; r < ROWS; r++)
    {
        uint32_t c0 = q0[r].x, c1 = q0[r].y, c2 = q1[r].x, c3 = q1[r].y, c4 = q2[r].x, c5 = q2[r].y;
        if (s & 2) { c0 = c1; c1 = c2; c2 = c3; c3 = c4; c4 = c5; }
        w[r][0] = __funnelshift_r(c0, c1, sh);
        w[r][1] = __funnelshift_r(c1, c2, sh);
        w[r][2] = __funnelshift_r(c2, c3, sh);
        w[r][3] = __funnelshift_r(c3, c4, sh);
    }
}
template<int ROWS>
__device__ __forceinline__ void load_rows8(const int16_t* p, intptr_t stride, uint32_t (&w)[ROWS][4]) { load_rows8<ROWS>((const uint16_t*)p, stride, w); }

// 8-bit samples: 4-byte chunks, widened to packed 16-bit pairs
template<int ROWS>
__device__ __forceinline__ void load_rows8(const uint8_t* p, intptr_t stride, uint32_t (&w)[ROWS][4])
{
    uintptr_t a = (uintptr_t)p;
    int s = (int)a & 3;
    const uint32_t* base = (const uint32_t*)(a & ~(uintptr_t)3);
    intptr_t cs = stride >> 2;
    uint32_t c0[ROWS], c1[ROWS], c2[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) { c0[r] = __ldg(base + r * cs); c1[r] = __ldg(base + r * cs + 1); }
    if (s)
    {
#pragma unroll
        for (int r = 0; r < ROWS; r++) c2[r] = __ldg(base + r * cs + 2);
    }
    else
    {
#pragma unroll
        for (int r = 0; r < ROWS; r++) c2[r] = 0;
    }
#pragma unroll
    for (int r = 0; r < ROWS; r++)
    {
        uint32_t v0 = __funnelshift_r(c0[r], c1[r], s << 3), v1 = __funnelshift_r(c1[r], c2[r], s << 3);
        w[r][0] = __byte_perm(v0, 0, 0x4140); w[r][1] = __byte_perm(v0, 0, 0x4342);
        w[r][2] = __byte_perm(v1, 0, 0x4140); w[r][3] = __byte_perm(v1, 0, 0x4342);
    }
}

// Four rows of eight samples through 16-byte aligned chunk loads (narrow blocks: a row of an 8-wide block is 16 bytes at an
// arbitrary 2-byte offset, i.e. it lies inside two aligned 16-byte chunks = at most two 32-byte sectors; 8-byte chunk loads
// touch the same sectors with three instructions).  stride % 8 == 0, 16-bit samples.  The per-lane sample offset (0..7)
// is resolved with two select stages (word offset bits 1 and 0) and one funnel shift per word.
__device__ __forceinline__ void load_rows8_v16(const uint16_t* p, intptr_t stride, uint32_t (&w)[4][4])
{
    uintptr_t a = (uintptr_t)p;
    int s = (int)(a >> 1) & 7;
    const uint4* base = (const uint4*)(a & ~(uintptr_t)15);
    intptr_t cs = stride >> 3;
    uint4 q0[4], q1[4];
#pragma unroll
    for (int r = 0; r < 4; r++) q0[r] = __ldg(base + r * cs);
    if (s)
    {
#pragma unroll
        for (int r = 0; r < 4; r++) q1[r] = __ldg(base + r * cs + 1);
    }
    else
    {
#pragma unroll
        for (int r = 0; r < 4; r++) q1[r] = make_uint4(0, 0, 0, 0);
    }
    int sh = (s & 1) << 4;
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        uint32_t c0 = q0[r].x, c1 = q0[r].y, c2 = q0[r].z, c3 = q0[r].w, c4 = q1[r].x, c5 = q1[r].y, c6 = q1[r].z, c7 = q1[r].w;
        if (s & 4) { c0 = c2; c1 = c3; c2 = c4; c3 = c5; c4 = c6; c5 = c7; }
        if (s & 2) { c0 = c1; c1 = c2; c2 = c3; c3 = c4; c4 = c5; }
        w[r][0] = __funnelshift_r(c0, c1, sh);
        w[r][1] = __funnelshift_r(c1, c2, sh);
        w[r][2] = __funnelshift_r(c2, c3, sh);
        w[r][3] = __funnelshift_r(c3, c4, sh);
    }
}

// eight int16 (four packed words) to `d`, with the widest stores its alignment allows
__device__ __forceinline__ void store8_s16(int16_t* d, const uint32_t (&w)[4])
{
    uintptr_t a = (uintptr_t)d;
    if ((a & 15) == 0) *(uint4*)d = make_uint4(w[0], w[1], w[2], w[3]);
    else if ((a & 7) == 0) { *(uint2*)d = make_uint2(w[0], w[1]); *(uint2*)(d + 4) = make_uint2(w[2], w[3]); }
    else if ((a & 3) == 0) { uint32_t* q = (uint32_t*)d; q[0] = w[0]; q[1] = w[1]; q[2] = w[2]; q[3] = w[3]; }
    else
    {
#pragma unroll
        for (int i = 0; i < 4; i++) { d[2 * i] = (int16_t)(w[i] & 0xffff); d[2 * i + 1] = (int16_t)(w[i] >> 16); }
    }
}

// lane-wise a - b of packed 16-bit pairs when every a, b is in [0, 32767]: setting bit 15 of each a-lane
// keeps the low lane from borrowing into the high one; the final xor removes it again (mod 2^16 per lane)
__device__ __forceinline__ uint32_t psub16(uint32_t a, uint32_t b) { return ((a | 0x80008000u) - b) ^ 0x80008000u; }

template<typename T> struct SampleTraits;
template<> struct SampleTraits<uint8_t>  { static __device__ __forceinline__ void unpack(uint32_t w, int& x, int& y) { x = w & 0xffff; y = w >> 16; } };
template<> struct SampleTraits<uint16_t> { static __device__ __forceinline__ void unpack(uint32_t w, int& x, int& y) { x = w & 0xffff; y = w >> 16; } };
template<> struct SampleTraits<int16_t>  { static __device__ __forceinline__ void unpack(uint32_t w, int& x, int& y) { x = (int)(int16_t)(w & 0xffff); y = (int)w >> 16; } };

// packed signed pair x + (y << 16) (as produced by subtracting two packed unsigned pairs) -> x, y
__device__ __forceinline__ void unpack_s16x2(uint32_t w, int& x, int& y)
{
    x = (int)(int16_t)(w & 0xffff);
    y = (int)(w + 0x8000u) >> 16;                          // +0x8000 undoes the borrow a negative x took
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
