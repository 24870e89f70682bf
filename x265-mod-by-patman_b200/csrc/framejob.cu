// framejob.cu -- the host-buffer layer of the C ABI: planes resident in HBM and frame jobs.
//
// The batched entries take device pointers; the one-block host slots stage a block per call.  A caller inside the encoder
// (ThreadedME, reference encoder/threadedme.cpp:207-261; the lookahead, encoder/slicetype.cpp:4467) has HOST pictures and
// wants HOST results for every block of a frame.  This file owns everything in between: plane memory with the reference's
// geometry (common/picyuv.cpp:86-118), border extension on the device (common/pixel.cpp:1044-1061), the upload / kernel /
// download pipeline over several frames in flight, and a compact return format for quantised levels so that the device ->
// host direction carries what the entropy coder consumes (significance bits + non-zero levels) instead of 2 bytes per
// coefficient.
#include "internal.h"

#include <stdio.h>
#include <string.h>

struct x265b200_plane
{
    x265b200_ctx* ctx = nullptr;
    int width = 0, height = 0, ctu = 0, hshift = 0, vshift = 0;
    int marginX = 0, marginY = 0, rows = 0;
    intptr_t stride = 0;
    size_t elems = 0;
    void* d = nullptr;
    cudaStream_t stream = nullptr;          // uploads of this plane
    cudaEvent_t ready = nullptr;            // last upload complete
    std::vector<cudaEvent_t> readers;       // job slots that read the plane since the last upload
};

namespace b200 {

// ---- border extension (extendPicBorder, reference common/pixel.cpp:1044-1061) --------------------------------------------
// phase 0: left / right margins of the `height` picture rows; phase 1: marginY copies of the first and last padded row.
template<typename T>
__global__ void extend_rows_kernel(T* pic, intptr_t stride, int width, int height, int marginX)
{
    int y = blockIdx.x;
    T* row = pic + (intptr_t)y * stride;
    T l = row[0], r = row[width - 1];
    for (int x = threadIdx.x; x < marginX; x += blockDim.x) { row[-marginX + x] = l; row[width + x] = r; }
}
template<typename T>
__global__ void extend_cols_kernel(T* pic, intptr_t stride, int width, int height, int marginX, int marginY)
{
    // blockIdx.y: 0 .. 2 * marginY - 1 (top copies, then bottom copies); x over the whole buffer row: the reference copies
    // `stride` samples (pixel.cpp:1052, :1057), including the columns right of width + marginX of a picture narrower than the plane
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= stride) return;
    int k = blockIdx.y;
    const T* top = pic - marginX;
    if (k < marginY) pic[-marginX + x - (intptr_t)(k + 1) * stride] = top[x];
    else
    {
        const T* bot = top + (intptr_t)(height - 1) * stride;
        pic[-marginX + x + (intptr_t)(height - 1 + (k - marginY + 1)) * stride] = bot[x];
    }
}

// ---- sparse return format for quantised levels ---------------------------------------------------------------------------
constexpr int PACK_THREADS = 256;
constexpr int PACK_CHUNK = 8192;            // coefficients per CTA: 256 threads x 8 x 4 rounds; every TU size divides it

// partial[c] = number of non-zero levels in chunk c (from the per-TU counts), numSig16 = the counts as uint16
__global__ void __launch_bounds__(PACK_THREADS)
levels_partial_kernel(const uint32_t* __restrict__ numSig, int n, int tusPerChunk, uint32_t* __restrict__ partial, uint16_t* __restrict__ numSig16)
{
    __shared__ uint32_t red[PACK_THREADS / 32];
    int c = blockIdx.x;
    uint32_t s = 0;
    for (int i = threadIdx.x; i < tusPerChunk; i += PACK_THREADS)
    {
        int t = c * tusPerChunk + i;
        if (t < n) { uint32_t v = numSig[t]; numSig16[t] = (uint16_t)v; s += v; }
    }
    s = __reduce_add_sync(0xffffffffu, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint32_t t = 0;
        for (int w = 0; w < PACK_THREADS / 32; w++) t += red[w];
        partial[c] = t;
    }
}

// flat ordered compaction: levels[] receives the non-zero coefficients of q[0 .. total) in order, sigMap one bit each
__global__ void __launch_bounds__(PACK_THREADS)
levels_pack_kernel(const int16_t* __restrict__ q, long long total, const uint32_t* __restrict__ partial, int nchunks,
                   uint32_t* __restrict__ sigMap, int16_t* __restrict__ levels, uint32_t* __restrict__ totalOut)
{
    __shared__ uint32_t red[PACK_THREADS / 32];
    __shared__ uint32_t sBase;
    const int c = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // base = sum of the partial counts of the chunks before this one
    uint32_t s = 0;
    for (int i = threadIdx.x; i < c; i += PACK_THREADS) s += partial[i];
    s = __reduce_add_sync(0xffffffffu, s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint32_t t = 0;
        for (int w = 0; w < PACK_THREADS / 32; w++) t += red[w];
        sBase = t;
        if (c == nchunks - 1) *totalOut = t + partial[c];
    }
    __syncthreads();
    uint32_t base = sBase;
    for (int round = 0; round < PACK_CHUNK / (PACK_THREADS * 8); round++)
    {
        long long idx = (long long)c * PACK_CHUNK + (long long)round * (PACK_THREADS * 8) + threadIdx.x * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (idx < total) v = *(const uint4*)(q + idx);
        const uint32_t w[4] = { v.x, v.y, v.z, v.w };
        uint32_t mask = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            if (w[k] & 0xffffu) mask |= 1u << (2 * k);
            if (w[k] >> 16) mask |= 2u << (2 * k);
        }
        // significance word of four neighbouring lanes (32 coefficients)
        uint32_t m1 = __shfl_down_sync(0xffffffffu, mask, 1), m2 = __shfl_down_sync(0xffffffffu, mask, 2), m3 = __shfl_down_sync(0xffffffffu, mask, 3);
        if ((lane & 3) == 0 && idx < total) sigMap[idx >> 5] = mask | (m1 << 8) | (m2 << 16) | (m3 << 24);
        // exclusive prefix of the counts over the CTA
        uint32_t cnt = __popc(mask), incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        __syncthreads();                        // red[] of the previous round / of the base sum has been consumed
        if (lane == 31) red[warp] = incl;
        __syncthreads();
        uint32_t before = 0, all = 0;
#pragma unroll
        for (int wv = 0; wv < PACK_THREADS / 32; wv++) { uint32_t t = red[wv]; if (wv < warp) before += t; all += t; }
        uint32_t pos = base + before + incl - cnt;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            if (w[k] & 0xffffu) levels[pos++] = (int16_t)(w[k] & 0xffffu);
            if (w[k] >> 16) levels[pos++] = (int16_t)(w[k] >> 16);
        }
        base += all;
    }
}

bool launch_tu_forward(x265b200_ctx* ctx, int N, const void* fenc, intptr_t sf, const void* pred, intptr_t sp,
                       const int32_t* offF, const int32_t* offP, int n, const int32_t* quantCoeff, int qBits, int qAdd,
                       int16_t* qCoef, uint32_t* numSig, uint64_t* sseZero, cudaStream_t st, int dst4);        // tu_fused.cuh

struct Pass
{
    int kind = 0, op = 0, w = 0, h = 0, N = 0, n = 0;
    int qBits = 0, add = 0;
    int32_t *dOffF = nullptr, *dOffR = nullptr, *dQuant = nullptr;
    size_t hostOff = 0;         // byte offset of this pass's results inside a slot's pinned / device result buffers
    size_t fixedBytes = 0;      // bytes copied unconditionally (costs / coefficients / numSig16 + sigMap + total)
    size_t levelsOff = 0;       // LEVELS: byte offset of the level stream (worst case n * N * N * 2 bytes reserved)
    int nchunks = 0;
};

struct Slot
{
    cudaStream_t stream = nullptr;
    cudaEvent_t sizes = nullptr, done = nullptr;
    uint8_t *dOut = nullptr, *hOut = nullptr;       // results, same layout on both sides
    int16_t* dScratch = nullptr;                    // residual / dense qCoef of the pass in flight
    uint32_t* dNumSig = nullptr;                    // per-TU counts of the pass in flight
    uint32_t* dPartial = nullptr;
    int state = 0;                                  // 0 free, 1 submitted (fixed-size results on their way), 2 draining (level streams enqueued)
    x265b200_plane *fenc = nullptr, *ref = nullptr;
};

} // namespace b200

using namespace b200;

struct x265b200_frame_job
{
    x265b200_ctx* ctx = nullptr;
    int width = 0, height = 0, ctu = 0;
    intptr_t stride = 0;
    size_t elems = 0;
    std::vector<Pass> passes;
    std::vector<Slot> slots;
    size_t outBytes = 0, scratchElems = 0;
    int maxTUs = 0, maxChunks = 0;
    bool sealed = false;
    int next = 0;
    bool anyLevels = false;
};

// ------------------------------------------------------------------ pinned host memory

extern "C" void* x265b200_host_alloc(x265b200_ctx* ctx, size_t bytes)
{
    if (!ctx || !bytes) return nullptr;
    void* p = nullptr;
    if (cudaSetDevice(ctx->device) != cudaSuccess) { fail(ctx, X265B200_ERR_CUDA, "cudaSetDevice"); return nullptr; }
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { fail(ctx, X265B200_ERR_CUDA, "host_alloc", e); return nullptr; }
    return p;
}
extern "C" void x265b200_host_free(x265b200_ctx* ctx, void* p) { (void)ctx; if (p) cudaFreeHost(p); }
extern "C" int x265b200_host_register(x265b200_ctx* ctx, void* p, size_t bytes)
{
    if (!ctx || !p || !bytes) return X265B200_ERR_ARG;
    B200_CUDA(ctx, cudaSetDevice(ctx->device));
    B200_CUDA(ctx, cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return X265B200_OK;
}
extern "C" int x265b200_host_unregister(x265b200_ctx* ctx, void* p)
{
    if (!ctx || !p) return X265B200_ERR_ARG;
    B200_CUDA(ctx, cudaHostUnregister(p));
    return X265B200_OK;
}
extern "C" void x265b200_transfer_stats(const x265b200_ctx* ctx, uint64_t* h2d, uint64_t* d2h)
{
    if (h2d) *h2d = ctx ? ctx->h2d_bytes.load() : 0;
    if (d2h) *d2h = ctx ? ctx->d2h_bytes.load() : 0;
}

// ------------------------------------------------------------------ planes

extern "C" int x265b200_plane_create(x265b200_ctx* ctx, int width, int height, int ctu, int hshift, int vshift, x265b200_plane** out)
{
    if (!ctx || !out) return X265B200_ERR_ARG;
    *out = nullptr;
    if (width < 1 || height < 1 || (ctu != 16 && ctu != 32 && ctu != 64) || hshift < 0 || hshift > 1 || vshift < 0 || vshift > 1)
        return fail(ctx, X265B200_ERR_ARG, "plane_create: bad geometry");
    B200_CUDA(ctx, cudaSetDevice(ctx->device));
    x265b200_plane* p = new x265b200_plane();
    p->ctx = ctx; p->width = width >> hshift; p->height = height >> vshift; p->ctu = ctu; p->hshift = hshift; p->vshift = vshift;
    const int cuW = (width + ctu - 1) / ctu, cuH = (height + ctu - 1) / ctu;
    p->marginX = ctu + 32;                                      // picyuv.cpp:89 and :106 (chroma keeps the luma margin)
    p->marginY = (ctu + 16) >> vshift;                          // picyuv.cpp:90 and :107
    p->stride = ((intptr_t)(cuW * ctu) >> hshift) + 2 * p->marginX;
    p->rows = ((cuH * ctu) >> vshift) + 2 * p->marginY;
    p->elems = (size_t)p->stride * p->rows;
    cudaError_t e = cudaMalloc(&p->d, p->elems * ctx->pixbytes);
    if (e == cudaSuccess) e = cudaMemset(p->d, 0, p->elems * ctx->pixbytes);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(p->ready, p->stream);
    if (e != cudaSuccess)
    {
        int rc = fail(ctx, X265B200_ERR_CUDA, "plane_create", e);
        x265b200_plane_destroy(p);
        return rc;
    }
    *out = p;
    return X265B200_OK;
}

extern "C" void x265b200_plane_destroy(x265b200_plane* p)
{
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    for (cudaEvent_t e : p->readers) { cudaEventSynchronize(e); cudaEventDestroy(e); }
    if (p->ready) cudaEventDestroy(p->ready);
    if (p->stream) cudaStreamDestroy(p->stream);
    if (p->d) cudaFree(p->d);
    delete p;
}

extern "C" int x265b200_plane_info(const x265b200_plane* p, intptr_t* stride, int* rows, int32_t* origin, size_t* elems, void** device)
{
    if (!p) return X265B200_ERR_ARG;
    if (stride) *stride = p->stride;
    if (rows) *rows = p->rows;
    if (origin) *origin = (int32_t)((intptr_t)p->marginY * p->stride + p->marginX);
    if (elems) *elems = p->elems;
    if (device) *device = p->d;
    return X265B200_OK;
}

// an upload must not overtake jobs that still read the previous contents
static int plane_begin_write(x265b200_plane* p)
{
    x265b200_ctx* ctx = p->ctx;
    B200_CUDA(ctx, cudaSetDevice(ctx->device));
    for (cudaEvent_t e : p->readers)
    {
        B200_CUDA(ctx, cudaStreamWaitEvent(p->stream, e, 0));
        cudaEventDestroy(e);            // released once the work captured above has completed
    }
    p->readers.clear();
    return X265B200_OK;
}

extern "C" int x265b200_plane_upload_padded(x265b200_plane* p, const void* host)
{
    if (!p || !host) return X265B200_ERR_ARG;
    x265b200_ctx* ctx = p->ctx;
    int rc = plane_begin_write(p);
    if (rc != X265B200_OK) return rc;
    size_t bytes = p->elems * ctx->pixbytes;
    B200_CUDA(ctx, cudaMemcpyAsync(p->d, host, bytes, cudaMemcpyHostToDevice, p->stream));
    B200_CUDA(ctx, cudaEventRecord(p->ready, p->stream));
    ctx->h2d_bytes.fetch_add(bytes, std::memory_order_relaxed);
    return X265B200_OK;
}

// extendPicBorder on the plane's stream
static int extend_on_device(x265b200_plane* p)
{
    x265b200_ctx* ctx = p->ctx;
    const size_t pb = ctx->pixbytes;
    uint8_t* org = (uint8_t*)p->d + ((size_t)p->marginY * p->stride + p->marginX) * pb;
    dim3 g2(ceil_div(p->stride, 256), 2 * p->marginY);
    if (pb == 1)
    {
        extend_rows_kernel<uint8_t><<<p->height, 128, 0, p->stream>>>((uint8_t*)org, p->stride, p->width, p->height, p->marginX);
        extend_cols_kernel<uint8_t><<<g2, 256, 0, p->stream>>>((uint8_t*)org, p->stride, p->width, p->height, p->marginX, p->marginY);
    }
    else
    {
        extend_rows_kernel<uint16_t><<<p->height, 128, 0, p->stream>>>((uint16_t*)org, p->stride, p->width, p->height, p->marginX);
        extend_cols_kernel<uint16_t><<<g2, 256, 0, p->stream>>>((uint16_t*)org, p->stride, p->width, p->height, p->marginX, p->marginY);
    }
    ctx->launches.fetch_add(2, std::memory_order_relaxed);
    if (cudaGetLastError() != cudaSuccess) return fail(ctx, X265B200_ERR_CUDA, "border extension launch");
    return X265B200_OK;
}

extern "C" int x265b200_plane_upload_picture(x265b200_plane* p, const void* host, intptr_t hostStride)
{
    if (!p || !host || hostStride < p->width) return X265B200_ERR_ARG;
    x265b200_ctx* ctx = p->ctx;
    int rc = plane_begin_write(p);
    if (rc != X265B200_OK) return rc;
    const size_t pb = ctx->pixbytes;
    uint8_t* org = (uint8_t*)p->d + ((size_t)p->marginY * p->stride + p->marginX) * pb;
    B200_CUDA(ctx, cudaMemcpy2DAsync(org, p->stride * pb, host, hostStride * pb, p->width * pb, p->height, cudaMemcpyHostToDevice, p->stream));
    rc = extend_on_device(p);
    if (rc != X265B200_OK) return rc;
    B200_CUDA(ctx, cudaEventRecord(p->ready, p->stream));
    ctx->h2d_bytes.fetch_add((size_t)p->width * p->height * pb, std::memory_order_relaxed);
    return X265B200_OK;
}

extern "C" int x265b200_plane_upload_rows(x265b200_plane* p, const void* hostPlane)
{
    if (!p || !hostPlane) return X265B200_ERR_ARG;
    x265b200_ctx* ctx = p->ctx;
    int rc = plane_begin_write(p);
    if (rc != X265B200_OK) return rc;
    const size_t pb = ctx->pixbytes;
    // the `height` picture rows as ONE linear copy (whole buffer rows, stride samples each): a strided 2-D copy of the picture
    // alone moves 5 % fewer bytes but runs at a lower DMA rate; the horizontal margins that come along are overwritten below
    const size_t first = (size_t)p->marginY * p->stride, bytes = (size_t)p->height * p->stride * pb;
    B200_CUDA(ctx, cudaMemcpyAsync((uint8_t*)p->d + first * pb, (const uint8_t*)hostPlane + first * pb, bytes, cudaMemcpyHostToDevice, p->stream));
    rc = extend_on_device(p);
    if (rc != X265B200_OK) return rc;
    B200_CUDA(ctx, cudaEventRecord(p->ready, p->stream));
    ctx->h2d_bytes.fetch_add(bytes, std::memory_order_relaxed);
    return X265B200_OK;
}

extern "C" int x265b200_plane_download_padded(x265b200_plane* p, void* host)
{
    if (!p || !host) return X265B200_ERR_ARG;
    x265b200_ctx* ctx = p->ctx;
    B200_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t bytes = p->elems * ctx->pixbytes;
    B200_CUDA(ctx, cudaMemcpyAsync(host, p->d, bytes, cudaMemcpyDeviceToHost, p->stream));
    B200_CUDA(ctx, cudaStreamSynchronize(p->stream));
    ctx->d2h_bytes.fetch_add(bytes, std::memory_order_relaxed);
    return X265B200_OK;
}

// ------------------------------------------------------------------ frame jobs

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" int x265b200_frame_job_create(x265b200_ctx* ctx, int width, int height, int ctu, int slots, x265b200_frame_job** out)
{
    if (!ctx || !out) return X265B200_ERR_ARG;
    *out = nullptr;
    if (width < 1 || height < 1 || (ctu != 16 && ctu != 32 && ctu != 64) || slots < 1 || slots > 8)
        return fail(ctx, X265B200_ERR_ARG, "frame_job_create: bad geometry / slot count");
    x265b200_frame_job* j = new x265b200_frame_job();
    j->ctx = ctx; j->width = width; j->height = height; j->ctu = ctu;
    const int cuW = (width + ctu - 1) / ctu, cuH = (height + ctu - 1) / ctu;
    j->stride = (intptr_t)cuW * ctu + 2 * (ctu + 32);
    j->elems = (size_t)j->stride * (cuH * ctu + 2 * (ctu + 16));
    j->slots.resize(slots);
    *out = j;
    return X265B200_OK;
}

static void free_slot(Slot& s)
{
    if (s.stream) cudaStreamSynchronize(s.stream);
    if (s.dOut) cudaFree(s.dOut);
    if (s.hOut) cudaFreeHost(s.hOut);
    if (s.dScratch) cudaFree(s.dScratch);
    if (s.dNumSig) cudaFree(s.dNumSig);
    if (s.dPartial) cudaFree(s.dPartial);
    if (s.sizes) cudaEventDestroy(s.sizes);
    if (s.done) cudaEventDestroy(s.done);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = Slot();
}

extern "C" void x265b200_frame_job_destroy(x265b200_frame_job* j)
{
    if (!j) return;
    cudaSetDevice(j->ctx->device);
    for (Slot& s : j->slots) free_slot(s);
    for (Pass& p : j->passes) { cudaFree(p.dOffF); cudaFree(p.dOffR); if (p.dQuant) cudaFree(p.dQuant); }
    delete j;
}

extern "C" int x265b200_frame_job_pass_count(const x265b200_frame_job* j) { return j ? (int)j->passes.size() : 0; }

static int upload_blocks(x265b200_frame_job* j, Pass& p, const int32_t* offF, const int32_t* offR)
{
    x265b200_ctx* ctx = j->ctx;
    // a block must lie inside the padded plane: the kernels trust the descriptors
    const long long last = (long long)(p.h - 1) * j->stride + p.w;
    for (int i = 0; i < p.n; i++)
        if (offF[i] < 0 || offR[i] < 0 || offF[i] + last > (long long)j->elems || offR[i] + last > (long long)j->elems)
            return fail(ctx, X265B200_ERR_ARG, "frame job: block descriptor outside the padded plane");
    B200_CUDA(ctx, cudaMemcpy(p.dOffF, offF, (size_t)p.n * 4, cudaMemcpyHostToDevice));
    B200_CUDA(ctx, cudaMemcpy(p.dOffR, offR, (size_t)p.n * 4, cudaMemcpyHostToDevice));
    ctx->h2d_bytes.fetch_add((size_t)p.n * 8, std::memory_order_relaxed);
    return X265B200_OK;
}

static int add_pass(x265b200_frame_job* j, Pass p, const int32_t* offF, const int32_t* offR, const int32_t* quantCoeff)
{
    x265b200_ctx* ctx = j->ctx;
    if (j->sealed) return fail(ctx, X265B200_ERR_ARG, "frame job: passes must be registered before the first submit");
    if (p.n < 1 || !offF || !offR) return fail(ctx, X265B200_ERR_ARG, "frame job: empty pass");
    B200_CUDA(ctx, cudaSetDevice(ctx->device));
    B200_CUDA(ctx, cudaMalloc((void**)&p.dOffF, (size_t)p.n * 4));
    B200_CUDA(ctx, cudaMalloc((void**)&p.dOffR, (size_t)p.n * 4));
    int rc = upload_blocks(j, p, offF, offR);
    if (rc != X265B200_OK) { cudaFree(p.dOffF); cudaFree(p.dOffR); return rc; }
    if (quantCoeff)
    {
        B200_CUDA(ctx, cudaMalloc((void**)&p.dQuant, (size_t)p.N * p.N * 4));
        B200_CUDA(ctx, cudaMemcpy(p.dQuant, quantCoeff, (size_t)p.N * p.N * 4, cudaMemcpyHostToDevice));
    }
    const size_t coefs = (size_t)p.n * p.N * p.N;
    p.hostOff = j->outBytes;
    if (p.kind == X265B200_PASS_CMP) p.fixedBytes = (size_t)p.n * 4;
    else if (p.kind == X265B200_PASS_COEF) p.fixedBytes = coefs * 2;
    else
    {
        // [total u32 | pad to 64][numSig16 n][sigMap coefs/32 words], then the level stream (worst case every coefficient)
        p.nchunks = ceil_div((long long)coefs, PACK_CHUNK);
        p.fixedBytes = align_up(64 + (size_t)p.n * 2, 64) + align_up(coefs / 8, 64);
        p.levelsOff = p.hostOff + align_up(p.fixedBytes, 256);
        j->anyLevels = true;
        if (p.nchunks > j->maxChunks) j->maxChunks = p.nchunks;
    }
    j->outBytes = align_up((p.kind == X265B200_PASS_LEVELS ? p.levelsOff + coefs * 2 : p.hostOff + p.fixedBytes), 256);
    if (p.kind != X265B200_PASS_CMP)
    {
        if (coefs > j->scratchElems) j->scratchElems = coefs;
        if (p.n > j->maxTUs) j->maxTUs = p.n;
    }
    j->passes.push_back(p);
    return (int)j->passes.size() - 1;
}

extern "C" int x265b200_frame_job_add_cmp(x265b200_frame_job* j, int op, int w, int h, const int32_t* offF, const int32_t* offR, int n)
{
    if (!j) return X265B200_ERR_ARG;
    if ((op != X265B200_SAD && op != X265B200_SATD && op != X265B200_SA8D) || w < 4 || h < 4 || (w & 3) || (h & 3) || w > 64 || h > 64)
        return fail(j->ctx, X265B200_ERR_ARG, "frame_job_add_cmp: bad op / shape");
    Pass p; p.kind = X265B200_PASS_CMP; p.op = op; p.w = w; p.h = h; p.n = n;
    return add_pass(j, p, offF, offR, nullptr);
}

extern "C" int x265b200_frame_job_add_transform(x265b200_frame_job* j, int kind, int N, const int32_t* offF, const int32_t* offR, int n,
                                                const int32_t* quantCoeff, int qBits, int add)
{
    if (!j) return X265B200_ERR_ARG;
    if ((kind != X265B200_PASS_COEF && kind != X265B200_PASS_LEVELS) || (N != 4 && N != 8 && N != 16 && N != 32))
        return fail(j->ctx, X265B200_ERR_ARG, "frame_job_add_transform: bad kind / size");
    if (kind == X265B200_PASS_LEVELS && (!quantCoeff || qBits < 8)) return fail(j->ctx, X265B200_ERR_ARG, "frame_job_add_transform: LEVELS needs a quant table");
    Pass p; p.kind = kind; p.N = N; p.w = N; p.h = N; p.n = n; p.qBits = qBits; p.add = add;
    return add_pass(j, p, offF, offR, kind == X265B200_PASS_LEVELS ? quantCoeff : nullptr);
}

extern "C" int x265b200_frame_job_set_blocks(x265b200_frame_job* j, int pass, const int32_t* offF, const int32_t* offR)
{
    if (!j || pass < 0 || pass >= (int)j->passes.size() || !offF || !offR) return X265B200_ERR_ARG;
    x265b200_ctx* ctx = j->ctx;
    B200_CUDA(ctx, cudaSetDevice(ctx->device));
    // the descriptors are shared by all slots: frames still in flight must finish with the old ones first
    for (Slot& s : j->slots) if (s.stream) B200_CUDA(ctx, cudaStreamSynchronize(s.stream));
    return upload_blocks(j, j->passes[pass], offF, offR);
}

static int seal(x265b200_frame_job* j)
{
    x265b200_ctx* ctx = j->ctx;
    if (j->passes.empty()) return fail(ctx, X265B200_ERR_ARG, "frame job: no passes registered");
    for (Slot& s : j->slots)
    {
        B200_CUDA(ctx, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        B200_CUDA(ctx, cudaEventCreateWithFlags(&s.sizes, cudaEventDisableTiming));
        B200_CUDA(ctx, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        B200_CUDA(ctx, cudaMalloc((void**)&s.dOut, j->outBytes));
        B200_CUDA(ctx, cudaHostAlloc((void**)&s.hOut, j->outBytes, cudaHostAllocPortable));
        if (j->scratchElems)
        {
            B200_CUDA(ctx, cudaMalloc((void**)&s.dScratch, j->scratchElems * 2));
            B200_CUDA(ctx, cudaMalloc((void**)&s.dNumSig, (size_t)j->maxTUs * 4));
            B200_CUDA(ctx, cudaMalloc((void**)&s.dPartial, (size_t)(j->maxChunks > 0 ? j->maxChunks : 1) * 4));
        }
    }
    j->sealed = true;
    return X265B200_OK;
}

// level streams have a data-dependent size: once a slot's counts are on the host, enqueue exactly that many bytes
static int drain(x265b200_frame_job* j, Slot& s, bool block)
{
    x265b200_ctx* ctx = j->ctx;
    if (s.state != 1) return X265B200_OK;
    if (block) B200_CUDA(ctx, cudaEventSynchronize(s.sizes));
    else
    {
        cudaError_t q = cudaEventQuery(s.sizes);
        if (q == cudaErrorNotReady) return X265B200_OK;
        if (q != cudaSuccess) return fail(ctx, X265B200_ERR_CUDA, "frame job: size event", q);
    }
    for (const Pass& p : j->passes)
    {
        if (p.kind != X265B200_PASS_LEVELS) continue;
        uint32_t total = *(const uint32_t*)(s.hOut + p.hostOff);
        if (total > (uint64_t)p.n * p.N * p.N) return fail(ctx, X265B200_ERR_CUDA, "frame job: level count out of range");
        size_t bytes = (size_t)total * 2;
        if (bytes)
        {
            B200_CUDA(ctx, cudaMemcpyAsync(s.hOut + p.levelsOff, s.dOut + p.levelsOff, bytes, cudaMemcpyDeviceToHost, s.stream));
            ctx->d2h_bytes.fetch_add(bytes, std::memory_order_relaxed);
        }
    }
    B200_CUDA(ctx, cudaEventRecord(s.done, s.stream));
    s.state = 2;
    return X265B200_OK;
}

extern "C" int x265b200_frame_job_submit(x265b200_frame_job* j, x265b200_plane* fenc, x265b200_plane* ref)
{
    if (!j || !fenc || !ref) return X265B200_ERR_ARG;
    x265b200_ctx* ctx = j->ctx;
    if (fenc->elems != j->elems || ref->elems != j->elems || fenc->stride != j->stride || ref->stride != j->stride || fenc->hshift || ref->hshift)
        return fail(ctx, X265B200_ERR_ARG, "frame_job_submit: plane geometry differs from the job's");
    B200_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!j->sealed) { int rc = seal(j); if (rc != X265B200_OK) return rc; }
    const int si = j->next;
    Slot& s = j->slots[si];
    if (s.state != 0) return fail(ctx, X265B200_ERR_ARG, "frame_job_submit: the next slot still holds results that were not waited for");
    // keep the other slots' level streams moving
    for (Slot& o : j->slots) { int rc = drain(j, o, false); if (rc != X265B200_OK) return rc; }

    cudaStream_t st = s.stream;
    B200_CUDA(ctx, cudaStreamWaitEvent(st, fenc->ready, 0));
    B200_CUDA(ctx, cudaStreamWaitEvent(st, ref->ready, 0));
    const intptr_t stride = j->stride;
    size_t d2h = 0, runStart = (size_t)-1;
    // metric passes first (their costs are one contiguous region when registered first), then the transforms
    for (const Pass& p : j->passes)
    {
        uint8_t* dres = s.dOut + p.hostOff;
        int rc = X265B200_OK;
        if (p.kind == X265B200_PASS_CMP)
            rc = x265b200_pixelcmp_batch(ctx, p.op, p.w, p.h, fenc->d, stride, ref->d, stride, p.dOffF, p.dOffR, p.n, dres, st);
        else if (p.kind == X265B200_PASS_COEF)
        {
            rc = x265b200_residual_batch(ctx, p.N, p.N, fenc->d, stride, ref->d, stride, p.dOffF, p.dOffR, p.n, s.dScratch, st);
            if (rc == X265B200_OK) rc = x265b200_dct_batch(ctx, X265B200_TR_DCT, p.N, s.dScratch, p.N, nullptr, p.n, (int16_t*)dres, st);
        }
        else
        {
            const long long coefs = (long long)p.n * p.N * p.N;
            rc = x265b200_tu_forward_batch(ctx, p.N, fenc->d, stride, ref->d, stride, p.dOffF, p.dOffR, p.n, p.dQuant, p.qBits, p.add,
                                           s.dScratch, s.dNumSig, nullptr, st);
            if (rc == X265B200_OK)
            {
                uint16_t* ns16 = (uint16_t*)(dres + 64);
                uint32_t* sig = (uint32_t*)(dres + align_up(64 + (size_t)p.n * 2, 64));
                levels_partial_kernel<<<p.nchunks, PACK_THREADS, 0, st>>>(s.dNumSig, p.n, PACK_CHUNK / (p.N * p.N), s.dPartial, ns16);
                levels_pack_kernel<<<p.nchunks, PACK_THREADS, 0, st>>>(s.dScratch, coefs, s.dPartial, p.nchunks, sig,
                                                                      (int16_t*)(s.dOut + p.levelsOff), (uint32_t*)dres);
                ctx->launches.fetch_add(2, std::memory_order_relaxed);
                if (cudaGetLastError() != cudaSuccess) rc = fail(ctx, X265B200_ERR_CUDA, "level packing launch");
            }
        }
        if (rc != X265B200_OK) return rc;
        // results go home as soon as they exist; the costs of consecutive metric passes (a few hundred KB each) travel as ONE copy
        const size_t pi = &p - &j->passes[0];
        if (p.kind == X265B200_PASS_CMP)
        {
            if (runStart == (size_t)-1) runStart = p.hostOff;
            const bool runEnds = pi + 1 == j->passes.size() || j->passes[pi + 1].kind != X265B200_PASS_CMP;
            if (runEnds)
            {
                size_t bytes = p.hostOff + p.fixedBytes - runStart;
                B200_CUDA(ctx, cudaMemcpyAsync(s.hOut + runStart, s.dOut + runStart, bytes, cudaMemcpyDeviceToHost, st));
                d2h += bytes;
                runStart = (size_t)-1;
            }
        }
        else
        {
            B200_CUDA(ctx, cudaMemcpyAsync(s.hOut + p.hostOff, dres, p.fixedBytes, cudaMemcpyDeviceToHost, st));
            d2h += p.fixedBytes;
        }
    }
    ctx->d2h_bytes.fetch_add(d2h, std::memory_order_relaxed);
    B200_CUDA(ctx, cudaEventRecord(s.sizes, st));
    // the planes may be overwritten once this slot's kernels are done
    for (x265b200_plane* pl : { fenc, ref })
    {
        if (pl->readers.size() >= 16)
        {   // a plane that is never uploaded again (a resident reference) must not collect events for ever
            size_t keep = 0;
            for (cudaEvent_t old : pl->readers)
                if (cudaEventQuery(old) == cudaSuccess) cudaEventDestroy(old); else pl->readers[keep++] = old;
            pl->readers.resize(keep);
        }
        cudaEvent_t e;
        B200_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        B200_CUDA(ctx, cudaEventRecord(e, st));
        pl->readers.push_back(e);
    }
    s.fenc = fenc; s.ref = ref;
    s.state = 1;
    j->next = (si + 1) % (int)j->slots.size();
    return si;
}

extern "C" int x265b200_frame_job_wait(x265b200_frame_job* j, int slot, x265b200_pass_result* results, int maxPasses)
{
    if (!j || slot < 0 || slot >= (int)j->slots.size()) return X265B200_ERR_ARG;
    x265b200_ctx* ctx = j->ctx;
    Slot& s = j->slots[slot];
    if (s.state == 0) return fail(ctx, X265B200_ERR_ARG, "frame_job_wait: slot was not submitted");
    B200_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = drain(j, s, true);
    if (rc != X265B200_OK) return rc;
    B200_CUDA(ctx, cudaEventSynchronize(s.done));
    s.state = 0;
    for (int i = 0; i < (int)j->passes.size() && i < maxPasses && results; i++)
    {
        const Pass& p = j->passes[i];
        x265b200_pass_result r;
        memset(&r, 0, sizeof(r));
        r.kind = p.kind; r.n = p.n;
        const uint8_t* h = s.hOut + p.hostOff;
        if (p.kind == X265B200_PASS_CMP) r.cost = (const int32_t*)h;
        else if (p.kind == X265B200_PASS_COEF) r.coef = (const int16_t*)h;
        else
        {
            r.nlevels = *(const uint32_t*)h;
            r.numSig = (const uint16_t*)(h + 64);
            r.sigMap = (const uint32_t*)(h + align_up(64 + (size_t)p.n * 2, 64));
            r.levels = (const int16_t*)(s.hOut + p.levelsOff);
        }
        results[i] = r;
    }
    return X265B200_OK;
}
