// integral.cu -- SEA motion search support: the integral planes `ads` consumes.
//
// Reference: encoder/framefilter.cpp:38-140 (integral_init{4,8,12,16,24,32}h_c / ..v_c) driven row by row by
// FrameFilter::computeMEIntegral (framefilter.cpp:737-835).  With T padded rows, row h of a plane first receives the column
// prefix of the W-wide horizontal sums of pixel rows 0..h-1 (inith), and H rows later (initv) the difference of two
// prefixes, so that finally
//     sum[r][x] = sum of the W x H pixel box whose top-left sample is (x, r)      for 1 <= r <= T - 1 - H, x < stride - W
// with row 0 all zero (the reference memsets it and never converts it).  All arithmetic is uint32 modulo 2^32.
// Cells outside that range hold prefix values / uninitialised memory in the reference and are never read by the search;
// the whole-plane entry writes 0 there.
//
// Whole-plane kernel: a CTA owns a 64 x 32 tile of outputs, builds the 2-D inclusive prefix sum of the 96 x 64 pixels the
// tile's largest box (32 x 32) can reach in shared memory (row and column scans by warp shuffles),
// and then emits all TWELVE planes from it with four shared-memory reads per value: the picture is read once.
#include "internal.h"

namespace b200 {

__constant__ int c_intW[12] = { 32, 32, 32, 24, 16, 16, 16, 12, 8, 8, 4, 4 };     // framefilter.cpp:776-787 plane order
__constant__ int c_intH[12] = { 32, 24, 8, 32, 16, 12, 4, 16, 32, 8, 16, 4 };

template<typename T>
__global__ void __launch_bounds__(256)
me_integral_kernel(const T* __restrict__ pix, intptr_t stride, int rows, size_t framePixels,
                   uint32_t* __restrict__ sums, size_t planePitch, size_t frameSums)
{
    constexpr int TX = 64, TY = 32, PW = TX + 32, PH = TY + 32;
    constexpr int PP = PW + 4;                        // row pitch: a multiple of four words so that quads of prefix values load as one LDS.128
    __shared__ __align__(16) uint32_t P[PH + 1][PP];
    const int x0 = blockIdx.x * TX, r0 = blockIdx.y * TY;
    const T* src = pix + (size_t)blockIdx.z * framePixels;
    uint32_t* dst = sums + (size_t)blockIdx.z * frameSums;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < PW + 1; i += 256) P[0][i] = 0;
    for (int i = tid; i < PH; i += 256) P[i + 1][0] = 0;
    for (int i = tid; i < PW * PH; i += 256)
    {
        int j = i / PW, c = i - j * PW;
        int gx = x0 + c, gr = r0 + j;
        P[j + 1][c + 1] = (gx < stride && gr < rows) ? (uint32_t)src[(size_t)gr * stride + gx] : 0u;
    }
    __syncthreads();
    // row-wise inclusive scans: warp w takes rows w, w + 8, ...; 96 columns = three 32-lane chunks with a carry
    for (int j = warp; j < PH; j += 8)
    {
        uint32_t carry = 0;
#pragma unroll
        for (int c = 0; c < PW; c += 32)
        {
            uint32_t v = P[j + 1][c + 1 + lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                uint32_t u = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v += u;
            }
            v += carry;
            P[j + 1][c + 1 + lane] = v;
            carry = __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    // column-wise inclusive scans, again by warp shuffles: lane = row inside a 32-row chunk, warp w takes columns w, w + 8, ...
    for (int c = warp; c < PW; c += 8)
    {
        uint32_t carry = 0;
#pragma unroll
        for (int j0 = 0; j0 < PH; j0 += 32)
        {
            uint32_t v = P[j0 + 1 + lane][c + 1];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                uint32_t u = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v += u;
            }
            v += carry;
            P[j0 + 1 + lane][c + 1] = v;
            carry = __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    for (int k = 0; k < 12; k++)
    {
        const int W = c_intW[k], H = c_intH[k];
        uint32_t* out = dst + (size_t)k * planePitch;
        // four adjacent outputs per thread: box(x) = P[y+H][x+W] - P[y][x+W] - P[y+H][x] + P[y][x], x and x + W multiples of 4
        for (int i = tid; i < (TX / 4) * TY; i += 256)
        {
            int y = i / (TX / 4), x = (i - y * (TX / 4)) * 4;
            int gx = x0 + x, gr = r0 + y;
            if (gx >= stride || gr >= rows) continue;
            uint4 a = *(const uint4*)&P[y + H][x + W], b = *(const uint4*)&P[y][x + W];
            uint4 c = *(const uint4*)&P[y + H][x], d = *(const uint4*)&P[y][x];
            bool rowOk = gr >= 1 && gr <= rows - 1 - H;
            int lim = (int)stride - W;                   // columns gx < lim are defined
            uint4 v;
            v.x = rowOk && gx + 0 < lim ? a.x - b.x - c.x + d.x : 0u;
            v.y = rowOk && gx + 1 < lim ? a.y - b.y - c.y + d.y : 0u;
            v.z = rowOk && gx + 2 < lim ? a.z - b.z - c.z + d.z : 0u;
            v.w = rowOk && gx + 3 < lim ? a.w - b.w - c.w + d.w : 0u;
            uint32_t* o = out + (size_t)gr * stride + gx;
            if (gx + 3 < stride && (((uintptr_t)o) & 15) == 0) *(uint4*)o = v;
            else
            {
                o[0] = v.x;
                if (gx + 1 < stride) o[1] = v.y;
                if (gx + 2 < stride) o[2] = v.z;
                if (gx + 3 < stride) o[3] = v.w;
            }
        }
    }
}

// row primitives (the slots themselves): sum[x] = hsum_W(pix, x) + above[x], and sum[x] = below[x] - sum[x]
template<typename T>
__global__ void __launch_bounds__(256)
integral_h_kernel(const T* __restrict__ pix, const uint32_t* __restrict__ above, uint32_t* __restrict__ sum, int W, int count)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= count) return;
    uint32_t v = 0;
    for (int i = 0; i < W; i++) v += pix[x + i];
    sum[x] = v + above[x];
}
__global__ void __launch_bounds__(256)
integral_v_kernel(const uint32_t* __restrict__ top, const uint32_t* __restrict__ bottom, uint32_t* __restrict__ out, int count)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x < count) out[x] = bottom[x] - top[x];
}

} // namespace b200

using namespace b200;

extern "C" int x265b200_me_integral_batch(x265b200_ctx* ctx, const void* pix, intptr_t stride, int rows, int nframes,
                                          uint32_t* sums, size_t planePitch, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (stride < 33 || rows < 34 || nframes < 0 || planePitch < (size_t)stride * rows)
        return fail(ctx, X265B200_ERR_ARG, "me_integral: bad geometry");
    if (nframes == 0) return X265B200_OK;
    dim3 grid((unsigned)ceil_div(stride, 64), (unsigned)ceil_div(rows, 32), (unsigned)nframes);
    cudaStream_t st = (cudaStream_t)stream;
    size_t fp = (size_t)stride * rows;
    if (ctx->pixbytes == 1)
        me_integral_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)pix, stride, rows, fp, sums, planePitch, 12 * planePitch);
    else
        me_integral_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)pix, stride, rows, fp, sums, planePitch, 12 * planePitch);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}

extern "C" int x265b200_integral_row_batch(x265b200_ctx* ctx, int vertical, int size, const void* pix, const uint32_t* a, const uint32_t* b,
                                           uint32_t* out, int count, x265b200_stream stream)
{
    if (!ctx) return X265B200_ERR_ARG;
    if (count < 0 || size < 1 || size > 64) return fail(ctx, X265B200_ERR_ARG, "integral_row: bad size");
    if (count == 0) return X265B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (vertical)
        integral_v_kernel<<<ceil_div(count, 256), 256, 0, st>>>(a, b, out, count);
    else if (ctx->pixbytes == 1)
        integral_h_kernel<uint8_t><<<ceil_div(count, 256), 256, 0, st>>>((const uint8_t*)pix, a, out, size, count);
    else
        integral_h_kernel<uint16_t><<<ceil_div(count, 256), 256, 0, st>>>((const uint16_t*)pix, a, out, size, count);
    B200_LAUNCH_CHECK(ctx);
    return X265B200_OK;
}
