// context.cu -- context lifecycle and the per-call HOST entries (drop-in slot bodies).
//
// A host entry receives the raw host pointers/strides an EncoderPrimitives slot receives
// (reference primitives.h:133-182), packs exactly the footprint the reference function would read
// into pinned staging, copies it to the device, runs the SAME batched kernel the device entries
// use with n = 1, copies the result back and writes exactly the cells the reference would write.
// No CPU arithmetic on sample data happens here: a missing/failed device makes the call a no-op
// that records a sticky error (x265b200_status), never a CPU fallback.
#include "internal.h"

#include <stdio.h>
#include <string.h>

namespace b200 {

int upload_transform_tables(x265b200_ctx* ctx);   // transform.cu
int upload_filter_tables(x265b200_ctx* ctx);      // ipfilter.cu
int upload_mma_tables(x265b200_ctx* ctx);         // transform_mma.cu

int fail(x265b200_ctx* ctx, int code, const char* what, cudaError_t e)
{
    if (!ctx) return code;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (ctx->status.load() == 0)
    {
        if (e != cudaSuccess) snprintf(ctx->err, sizeof(ctx->err), "%s: %s", what, cudaGetErrorString(e));
        else snprintf(ctx->err, sizeof(ctx->err), "%s", what);
        ctx->status.store(code);        // published after the text: a reader that sees the status sees the whole message
    }
    return code;
}

static Lane* lane_acquire(x265b200_ctx* ctx)
{
    if (cudaSetDevice(ctx->device) != cudaSuccess) { fail(ctx, X265B200_ERR_CUDA, "cudaSetDevice"); return nullptr; }
    {
        std::lock_guard<std::mutex> g(ctx->mu);
        while (!ctx->free_lanes.empty())
        {
            Lane* l = ctx->free_lanes.back();
            ctx->free_lanes.pop_back();
            if (l) return l;                    // the pool never hands out a null lane
        }
    }
    Lane* l = new Lane();
    cudaError_t e = cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&l->h, LANE_BYTES);
    if (e == cudaSuccess) e = cudaMalloc((void**)&l->d, LANE_BYTES);
    if (e != cudaSuccess)
    {
        fail(ctx, X265B200_ERR_CUDA, "lane allocation", e);
        if (l->h) cudaFreeHost(l->h);
        if (l->stream) cudaStreamDestroy(l->stream);
        delete l;
        return nullptr;
    }
    l->cap = LANE_BYTES;
    std::lock_guard<std::mutex> g(ctx->mu);
    ctx->all_lanes.push_back(l);
    return l;
}

static void lane_release(x265b200_ctx* ctx, Lane* l)
{
    if (!l) return;
    std::lock_guard<std::mutex> g(ctx->mu);
    ctx->free_lanes.push_back(l);
}

// Whole-plane slots (frameInitLowres, weight_pp on a padded lowres plane: reference common/lowres.cpp:385,
// encoder/slicetype.cpp:880) stage far more than one block: the lane's two buffers are replaced by larger ones.  Only called
// before anything of the current call has been written into the lane (every entry allocates first, then packs).
static bool lane_grow(x265b200_ctx* ctx, Lane* l, size_t need)
{
    if (need > LANE_MAX_BYTES) { fail(ctx, X265B200_ERR_ARG, "host call exceeds the maximum staging size (2 GiB)"); return false; }
    size_t cap = l->cap;
    while (cap < need) cap <<= 1;
    uint8_t *h = nullptr, *d = nullptr;
    cudaError_t e = cudaStreamSynchronize(l->stream);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&h, cap);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d, cap);
    if (e != cudaSuccess)
    {
        if (h) cudaFreeHost(h);
        fail(ctx, X265B200_ERR_CUDA, "lane growth", e);
        return false;
    }
    cudaFreeHost(l->h); cudaFree(l->d);
    l->h = h; l->d = d; l->cap = cap;
    return true;
}

// One host call: bump-allocates matching host/device staging, then upload -> kernels -> download.
struct Call
{
    x265b200_ctx* ctx;
    Lane* lane;
    size_t used = 0;
    explicit Call(x265b200_ctx* c) : ctx(c), lane(c ? lane_acquire(c) : nullptr) {}
    ~Call() { if (lane) lane_release(ctx, lane); }
    bool ok() const { return lane != nullptr; }
    size_t alloc(size_t bytes)
    {
        if (!lane) return 0;                // an earlier alloc of this call already failed
        size_t off = (used + 63) & ~(size_t)63;
        used = off + bytes;
        if (used > lane->cap && !lane_grow(ctx, lane, used)) { lane_release(ctx, lane); lane = nullptr; return 0; }
        return off;
    }
    template<typename T> T* h(size_t off) { return (T*)(lane->h + off); }
    template<typename T> T* d(size_t off) { return (T*)(lane->d + off); }
    cudaStream_t st() const { return lane->stream; }
    bool upload(size_t off, size_t bytes)
    {
        cudaError_t e = cudaMemcpyAsync(lane->d + off, lane->h + off, bytes, cudaMemcpyHostToDevice, lane->stream);
        if (e != cudaSuccess) { fail(ctx, X265B200_ERR_CUDA, "H2D", e); return false; }
        return true;
    }
    bool download(size_t off, size_t bytes)
    {
        cudaError_t e = cudaMemcpyAsync(lane->h + off, lane->d + off, bytes, cudaMemcpyDeviceToHost, lane->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(lane->stream);
        if (e != cudaSuccess) { fail(ctx, X265B200_ERR_CUDA, "D2H", e); return false; }
        return true;
    }
};

// copy `rows` rows of `rowBytes` from a strided host block into packed staging
static void pack(void* dst, const void* src, int rows, size_t rowBytes, intptr_t strideBytes)
{
    uint8_t* d = (uint8_t*)dst;
    const uint8_t* s = (const uint8_t*)src;
    for (int r = 0; r < rows; r++) memcpy(d + r * rowBytes, s + r * strideBytes, rowBytes);
}
static void unpack(void* dst, const void* src, int rows, size_t rowBytes, intptr_t strideBytes)
{
    uint8_t* d = (uint8_t*)dst;
    const uint8_t* s = (const uint8_t*)src;
    for (int r = 0; r < rows; r++) memcpy(d + r * strideBytes, s + r * rowBytes, rowBytes);
}

} // namespace b200

using namespace b200;

// ------------------------------------------------------------------ lifecycle

extern "C" int x265b200_open(int device, int bit_depth, x265b200_ctx** out)
{
    if (!out || (bit_depth != 8 && bit_depth != 10 && bit_depth != 12)) return X265B200_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return X265B200_ERR_NO_DEVICE;
    if (device < 0 || device >= count) return X265B200_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return X265B200_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return X265B200_ERR_CUDA;
    if (prop.major != 10)
    {
        fprintf(stderr, "x265b200: device %d is sm_%d%d; this library is built for sm_100a only\n", device, prop.major, prop.minor);
        return X265B200_ERR_NO_DEVICE;
    }
    x265b200_ctx* ctx = new x265b200_ctx();
    ctx->device = device;
    ctx->depth = bit_depth;
    ctx->pixbytes = bit_depth == 8 ? 1 : 2;
    ctx->sm_count = prop.multiProcessorCount;
    // stream-ordered scratch (x265b200_tu_chain_batch) must not go back to the OS at every synchronisation
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
    {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    int r = upload_transform_tables(ctx);
    if (r == X265B200_OK) r = upload_filter_tables(ctx);
    if (r == X265B200_OK) r = upload_mma_tables(ctx);
    if (r != X265B200_OK) { fprintf(stderr, "x265b200: %s\n", ctx->err); delete ctx; return r; }
    *out = ctx;
    return X265B200_OK;
}

extern "C" void x265b200_close(x265b200_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (Lane* l : ctx->all_lanes)
    {
        cudaStreamSynchronize(l->stream);
        cudaFree(l->d);
        cudaFreeHost(l->h);
        cudaStreamDestroy(l->stream);
        delete l;
    }
    delete ctx;
}

extern "C" int x265b200_set_dct_path(x265b200_ctx* ctx, int path)
{
    if (!ctx || path < 0 || path > 3) return X265B200_ERR_ARG;
    ctx->dct_path = path;
    return X265B200_OK;
}

extern "C" int x265b200_bit_depth(const x265b200_ctx* ctx) { return ctx ? ctx->depth : 0; }
extern "C" int x265b200_sm_count(const x265b200_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
extern "C" int x265b200_status(const x265b200_ctx* ctx) { return ctx ? ctx->status.load() : X265B200_ERR_ARG; }
extern "C" const char* x265b200_last_error(const x265b200_ctx* ctx) { return ctx ? ctx->err : "null context"; }
extern "C" uint64_t x265b200_launch_count(const x265b200_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

// ------------------------------------------------------------------ host entries: metrics

// shared body of sad / satd / sa8d / sse_pp
static uint64_t host_pixelcmp(x265b200_ctx* ctx, int op, int w, int h, const void* a, intptr_t sa, const void* b, intptr_t sb)
{
    Call c(ctx);
    if (!c.ok()) return 0;
    size_t pb = ctx->pixbytes, blk = (size_t)w * h * pb;
    size_t oA = c.alloc(blk), oB = c.alloc(blk), oOff = c.alloc(8), oOut = c.alloc(8);
    if (!c.ok()) return 0;
    pack(c.h<void>(oA), a, h, w * pb, sa * pb);
    pack(c.h<void>(oB), b, h, w * pb, sb * pb);
    c.h<int32_t>(oOff)[0] = 0; c.h<int32_t>(oOff)[1] = 0;
    if (!c.upload(0, oOff + 8)) return 0;
    if (x265b200_pixelcmp_batch(ctx, op, w, h, c.d<void>(oA), w, c.d<void>(oB), w, c.d<int32_t>(oOff), c.d<int32_t>(oOff) + 1,
                                1, c.d<void>(oOut), c.st()) != X265B200_OK) return 0;
    if (!c.download(oOut, 8)) return 0;
    return op == X265B200_SSE_PP ? *c.h<uint64_t>(oOut) : (uint64_t)(int64_t)*c.h<int32_t>(oOut);
}

extern "C" int x265b200_sad(x265b200_ctx* ctx, int w, int h, const void* a, intptr_t sa, const void* b, intptr_t sb)
{ return (int)host_pixelcmp(ctx, X265B200_SAD, w, h, a, sa, b, sb); }
extern "C" int x265b200_satd(x265b200_ctx* ctx, int w, int h, const void* a, intptr_t sa, const void* b, intptr_t sb)
{ return (int)host_pixelcmp(ctx, X265B200_SATD, w, h, a, sa, b, sb); }
extern "C" int x265b200_sa8d(x265b200_ctx* ctx, int w, int h, const void* a, intptr_t sa, const void* b, intptr_t sb)
{ return (int)host_pixelcmp(ctx, X265B200_SA8D, w, h, a, sa, b, sb); }
extern "C" uint64_t x265b200_sse_pp(x265b200_ctx* ctx, int w, int h, const void* a, intptr_t sa, const void* b, intptr_t sb)
{ return host_pixelcmp(ctx, X265B200_SSE_PP, w, h, a, sa, b, sb); }

extern "C" uint64_t x265b200_sse_ss(x265b200_ctx* ctx, int w, int h, const int16_t* a, intptr_t sa, const int16_t* b, intptr_t sb)
{
    Call c(ctx);
    if (!c.ok()) return 0;
    size_t blk = (size_t)w * h * 2;
    size_t oA = c.alloc(blk), oB = c.alloc(blk), oOff = c.alloc(8), oOut = c.alloc(8);
    if (!c.ok()) return 0;
    pack(c.h<void>(oA), a, h, w * 2, sa * 2);
    pack(c.h<void>(oB), b, h, w * 2, sb * 2);
    c.h<int32_t>(oOff)[0] = 0; c.h<int32_t>(oOff)[1] = 0;
    if (!c.upload(0, oOff + 8)) return 0;
    if (x265b200_sse_ss_batch(ctx, w, h, c.d<int16_t>(oA), w, c.d<int16_t>(oB), w, c.d<int32_t>(oOff), c.d<int32_t>(oOff) + 1,
                              1, c.d<uint64_t>(oOut), c.st()) != X265B200_OK) return 0;
    if (!c.download(oOut, 8)) return 0;
    return *c.h<uint64_t>(oOut);
}

extern "C" uint64_t x265b200_ssd_s(x265b200_ctx* ctx, int size, const int16_t* a, intptr_t sa)
{
    Call c(ctx);
    if (!c.ok()) return 0;
    size_t oA = c.alloc((size_t)size * size * 2), oOff = c.alloc(8), oOut = c.alloc(8);
    if (!c.ok()) return 0;
    pack(c.h<void>(oA), a, size, size * 2, sa * 2);
    c.h<int32_t>(oOff)[0] = 0;
    if (!c.upload(0, oOff + 8)) return 0;
    if (x265b200_ssd_s_batch(ctx, size, c.d<int16_t>(oA), size, c.d<int32_t>(oOff), 1, c.d<uint64_t>(oOut), c.st()) != X265B200_OK) return 0;
    if (!c.download(oOut, 8)) return 0;
    return *c.h<uint64_t>(oOut);
}

static void host_sad_xn(x265b200_ctx* ctx, int w, int h, int K, const void* fenc, const void* const* refs, intptr_t rs, int32_t* res)
{
    Call c(ctx);
    if (!c.ok()) return;
    size_t pb = ctx->pixbytes, blk = (size_t)w * h * pb;
    size_t oF = c.alloc(blk), oR = c.alloc(blk * K), oOff = c.alloc(4 * (1 + K)), oOut = c.alloc(4 * K);
    if (!c.ok()) return;
    pack(c.h<void>(oF), fenc, h, w * pb, FENC_STRIDE * pb);                     // pixel.cpp:86 / :110
    for (int k = 0; k < K; k++) pack(c.h<uint8_t>(oR) + k * blk, refs[k], h, w * pb, rs * pb);
    int32_t* off = c.h<int32_t>(oOff);
    off[0] = 0;
    for (int k = 0; k < K; k++) off[1 + k] = k * w * h;
    if (!c.upload(0, oOff + 4 * (1 + K))) return;
    if (x265b200_sad_multi_batch(ctx, w, h, c.d<void>(oF), w, c.d<void>(oR), w, c.d<int32_t>(oOff), c.d<int32_t>(oOff) + 1,
                                 K, 1, c.d<int32_t>(oOut), c.st()) != X265B200_OK) return;
    if (!c.download(oOut, 4 * K)) return;
    memcpy(res, c.h<int32_t>(oOut), 4 * K);
}

extern "C" void x265b200_sad_x3(x265b200_ctx* ctx, int w, int h, const void* fenc, const void* r0, const void* r1, const void* r2,
                                intptr_t rs, int32_t* res)
{ const void* r[3] = { r0, r1, r2 }; host_sad_xn(ctx, w, h, 3, fenc, r, rs, res); }
extern "C" void x265b200_sad_x4(x265b200_ctx* ctx, int w, int h, const void* fenc, const void* r0, const void* r1, const void* r2,
                                const void* r3, intptr_t rs, int32_t* res)
{ const void* r[4] = { r0, r1, r2, r3 }; host_sad_xn(ctx, w, h, 4, fenc, r, rs, res); }

// number of DC terms the reference binds to a w x h PU (pixel.cpp:1122-1146)
static int ads_terms(int w, int h)
{
    static const struct { int w, h, k; } tab[] = {
        {4,4,1},{8,8,1},{8,4,2},{4,8,2},{16,16,4},{16,8,2},{8,16,2},{16,12,1},{12,16,1},
        {16,4,1},{4,16,1},{32,32,4},{32,16,2},{16,32,2},{32,24,4},{24,32,4},{32,8,4},{8,32,4},
        {64,64,4},{64,32,2},{32,64,2},{64,48,4},{48,64,4},{64,16,4},{16,64,4} };
    for (const auto& t : tab) if (t.w == w && t.h == h) return t.k;
    return 0;
}

extern "C" int x265b200_ads(x265b200_ctx* ctx, int w, int h, const int* encDC, const uint32_t* sums, int delta,
                            const uint16_t* costMvX, int16_t* mvs, int width, int thresh)
{
    int terms = ads_terms(w, h);
    if (!terms || width <= 0) { if (!terms) fail(ctx, X265B200_ERR_ARG, "ads: not a PU size"); return 0; }
    Call c(ctx);
    if (!c.ok()) return 0;
    int half = terms == 4 ? w >> 1 : 0;
    int seg = width + half;                                   // samples read from each of the two sum rows
    size_t oEnc = c.alloc(16), oSum = c.alloc((size_t)seg * 8), oCost = c.alloc((size_t)width * 2), oJob = c.alloc(5 * 4);
    size_t oMvs = c.alloc((size_t)width * 2), oCnt = c.alloc(4);
    if (!c.ok()) return 0;
    int32_t* e = c.h<int32_t>(oEnc);
    for (int i = 0; i < 4; i++) e[i] = i < terms ? encDC[i] : 0;
    memcpy(c.h<uint32_t>(oSum), sums, (size_t)seg * 4);
    if (terms > 1) memcpy(c.h<uint32_t>(oSum) + seg, sums + delta, (size_t)seg * 4);
    memcpy(c.h<void>(oCost), costMvX, (size_t)width * 2);
    int32_t* job = c.h<int32_t>(oJob);
    job[0] = 0; job[1] = seg; job[2] = 0; job[3] = width; job[4] = thresh;     // sumOff, delta, costOff, width, thresh
    if (!c.upload(0, oJob + 20)) return 0;
    int32_t* dj = c.d<int32_t>(oJob);
    if (x265b200_ads_batch(ctx, terms, half, c.d<int32_t>(oEnc), c.d<uint32_t>(oSum), dj, dj + 1, c.d<uint16_t>(oCost), dj + 2,
                           dj + 3, dj + 4, 1, c.d<int16_t>(oMvs), width, c.d<int32_t>(oCnt), c.st()) != X265B200_OK) return 0;
    if (!c.download(oMvs, oCnt + 4 - oMvs)) return 0;
    int n = *c.h<int32_t>(oCnt);
    memcpy(mvs, c.h<int16_t>(oMvs), (size_t)n * 2);
    return n;
}

// ------------------------------------------------------------------ host entries: transforms

extern "C" void x265b200_dct(x265b200_ctx* ctx, int kind, int N, const int16_t* src, int16_t* dst, intptr_t srcStride)
{
    Call c(ctx);
    if (!c.ok()) return;
    size_t blk = (size_t)N * N * 2;
    size_t oS = c.alloc(blk), oOff = c.alloc(4), oD = c.alloc(blk);
    if (!c.ok()) return;
    pack(c.h<void>(oS), src, N, N * 2, srcStride * 2);
    *c.h<int32_t>(oOff) = 0;
    if (!c.upload(0, oOff + 4)) return;
    if (x265b200_dct_batch(ctx, kind, N, c.d<int16_t>(oS), N, c.d<int32_t>(oOff), 1, c.d<int16_t>(oD), c.st()) != X265B200_OK) return;
    if (!c.download(oD, blk)) return;
    memcpy(dst, c.h<void>(oD), blk);
}

extern "C" void x265b200_idct(x265b200_ctx* ctx, int kind, int N, const int16_t* src, int16_t* dst, intptr_t dstStride)
{
    Call c(ctx);
    if (!c.ok()) return;
    size_t blk = (size_t)N * N * 2;
    size_t oS = c.alloc(blk), oOff = c.alloc(4), oD = c.alloc(blk);
    if (!c.ok()) return;
    memcpy(c.h<void>(oS), src, blk);
    *c.h<int32_t>(oOff) = 0;
    if (!c.upload(0, oOff + 4)) return;
    if (x265b200_idct_batch(ctx, kind, N, c.d<int16_t>(oS), 1, c.d<int16_t>(oD), N, c.d<int32_t>(oOff), c.st()) != X265B200_OK) return;
    if (!c.download(oD, blk)) return;
    unpack(dst, c.h<void>(oD), N, N * 2, dstStride * 2);
}

static uint32_t host_quant(x265b200_ctx* ctx, const int16_t* coef, const int32_t* qc, int32_t* deltaU, int16_t* qCoef,
                           int qBits, int add, int numCoeff)
{
    Call c(ctx);
    if (!c.ok()) return 0;
    size_t n = (size_t)numCoeff;
    size_t oC = c.alloc(n * 2), oQ = c.alloc(n * 4), oIn = c.used;
    size_t oOut = c.alloc(n * 2), oDu = c.alloc(n * 4), oSig = c.alloc(4);
    if (!c.ok()) return 0;
    memcpy(c.h<void>(oC), coef, n * 2);
    memcpy(c.h<void>(oQ), qc, n * 4);
    if (!c.upload(0, oIn)) return 0;
    if (x265b200_quant_batch(ctx, c.d<int16_t>(oC), c.d<int32_t>(oQ), deltaU ? c.d<int32_t>(oDu) : nullptr, c.d<int16_t>(oOut),
                             qBits, add, numCoeff, 1, c.d<uint32_t>(oSig), c.st()) != X265B200_OK) return 0;
    if (!c.download(oOut, oSig + 4 - oOut)) return 0;
    memcpy(qCoef, c.h<void>(oOut), n * 2);
    if (deltaU) memcpy(deltaU, c.h<void>(oDu), n * 4);
    return *c.h<uint32_t>(oSig);
}

extern "C" uint32_t x265b200_quant(x265b200_ctx* ctx, const int16_t* coef, const int32_t* qc, int32_t* deltaU, int16_t* qCoef,
                                   int qBits, int add, int numCoeff)
{
    if (!deltaU) { fail(ctx, X265B200_ERR_ARG, "quant: deltaU is NULL"); return 0; }
    return host_quant(ctx, coef, qc, deltaU, qCoef, qBits, add, numCoeff);
}
extern "C" uint32_t x265b200_nquant(x265b200_ctx* ctx, const int16_t* coef, const int32_t* qc, int16_t* qCoef, int qBits, int add, int numCoeff)
{ return host_quant(ctx, coef, qc, nullptr, qCoef, qBits, add, numCoeff); }

extern "C" void x265b200_dequant_normal(x265b200_ctx* ctx, const int16_t* q, int16_t* coef, int num, int scale, int shift)
{
    Call c(ctx);
    if (!c.ok()) return;
    size_t oS = c.alloc((size_t)num * 2), oD = c.alloc((size_t)num * 2);
    if (!c.ok()) return;
    memcpy(c.h<void>(oS), q, (size_t)num * 2);
    if (!c.upload(0, (size_t)num * 2)) return;
    if (x265b200_dequant_normal_batch(ctx, c.d<int16_t>(oS), c.d<int16_t>(oD), num, scale, shift, c.st()) != X265B200_OK) return;
    if (!c.download(oD, (size_t)num * 2)) return;
    memcpy(coef, c.h<void>(oD), (size_t)num * 2);
}

extern "C" void x265b200_dequant_scaling(x265b200_ctx* ctx, const int16_t* q, const int32_t* dq, int16_t* coef, int num, int per, int shift)
{
    Call c(ctx);
    if (!c.ok()) return;
    size_t oS = c.alloc((size_t)num * 2), oT = c.alloc((size_t)num * 4), oIn = c.used, oD = c.alloc((size_t)num * 2);
    if (!c.ok()) return;
    memcpy(c.h<void>(oS), q, (size_t)num * 2);
    memcpy(c.h<void>(oT), dq, (size_t)num * 4);
    if (!c.upload(0, oIn)) return;
    if (x265b200_dequant_scaling_batch(ctx, c.d<int16_t>(oS), c.d<int32_t>(oT), c.d<int16_t>(oD), num, 1, per, shift, c.st()) != X265B200_OK) return;
    if (!c.download(oD, (size_t)num * 2)) return;
    memcpy(coef, c.h<void>(oD), (size_t)num * 2);
}

// ------------------------------------------------------------------ host entries: interpolation

extern "C" void x265b200_interp(x265b200_ctx* ctx, int kind, int taps, int w, int h, const void* src, intptr_t srcStride,
                                void* dst, intptr_t dstStride, int coeffIdx, int extra)
{
    Call c(ctx);
    if (!c.ok()) return;
    const bool srcShort = kind == X265B200_IP_VSP || kind == X265B200_IP_VSS;
    const bool dstShort = kind == X265B200_IP_HPS || kind == X265B200_IP_VPS || kind == X265B200_IP_VSS || kind == X265B200_IP_P2S;
    const size_t sb = srcShort ? 2 : ctx->pixbytes, db = dstShort ? 2 : ctx->pixbytes;
    const bool horiz = kind == X265B200_IP_HPP || kind == X265B200_IP_HPS || kind == X265B200_IP_HVPP;
    const bool rowExt = kind == X265B200_IP_HVPP || (kind == X265B200_IP_HPS && extra);
    const bool vert = kind == X265B200_IP_VPP || kind == X265B200_IP_VPS || kind == X265B200_IP_VSP || kind == X265B200_IP_VSS;
    // footprint the reference reads: taps/2-1 before and taps/2 after along each filtered direction
    int left = horiz ? taps / 2 - 1 : 0, right = horiz ? taps / 2 : 0;
    int top = (vert || rowExt) ? taps / 2 - 1 : 0, bottom = (vert || rowExt) ? taps / 2 : 0;
    int pw = w + left + right, ph = h + top + bottom;
    int outRows = (kind == X265B200_IP_HPS && extra) ? h + taps - 1 : h;
    size_t oS = c.alloc((size_t)pw * ph * sb), oJob = c.alloc(12), oD = c.alloc((size_t)w * outRows * db);
    if (!c.ok()) return;
    pack(c.h<void>(oS), (const uint8_t*)src - ((intptr_t)top * srcStride + left) * (intptr_t)sb, ph, pw * sb, srcStride * sb);
    int32_t* job = c.h<int32_t>(oJob);
    job[0] = top * pw + left;                                   // offSrc: the block origin inside the packed tile
    job[1] = 0;                                                 // offDst
    job[2] = kind == X265B200_IP_HVPP ? (coeffIdx | extra << 4) : (coeffIdx | (extra ? 1 << 8 : 0));
    if (!c.upload(0, oJob + 12)) return;
    int32_t* dj = c.d<int32_t>(oJob);
    if (x265b200_interp_batch(ctx, kind, taps, w, h, c.d<void>(oS), pw, dj, c.d<void>(oD), w, dj + 1, dj + 2, 1, c.st()) != X265B200_OK) return;
    if (!c.download(oD, (size_t)w * outRows * db)) return;
    unpack(dst, c.h<void>(oD), outRows, w * db, dstStride * db);
}

// ------------------------------------------------------------------ host entries: adjacent slots (blockops.cu)

static void host_blockop(x265b200_ctx* ctx, int op, int w, int h, const void* a, intptr_t sa, size_t ab, const void* b, intptr_t sb, size_t bb,
                         void* d, intptr_t sd, size_t db)
{
    Call c(ctx);
    if (!c.ok()) return;
    size_t oA = c.alloc((size_t)w * h * ab), oB = c.alloc((size_t)w * h * bb);
    size_t inEnd = c.used;
    size_t oD = c.alloc((size_t)w * h * db);
    if (!c.ok()) return;
    pack(c.h<void>(oA), a, h, w * ab, sa * (intptr_t)ab);
    pack(c.h<void>(oB), b, h, w * bb, sb * (intptr_t)bb);
    if (!c.upload(0, inEnd)) return;
    if (x265b200_blockop_batch(ctx, op, w, h, c.d<void>(oA), w, nullptr, c.d<void>(oB), w, nullptr, c.d<void>(oD), w, nullptr, 1, c.st()) != X265B200_OK) return;
    if (!c.download(oD, (size_t)w * h * db)) return;
    unpack(d, c.h<void>(oD), h, w * db, sd * (intptr_t)db);
}

extern "C" void x265b200_sub_ps(x265b200_ctx* ctx, int w, int h, int16_t* dst, intptr_t dstride, const void* src0, const void* src1, intptr_t sstride0, intptr_t sstride1)
{ if (ctx) host_blockop(ctx, X265B200_BOP_SUB_PS, w, h, src0, sstride0, ctx->pixbytes, src1, sstride1, ctx->pixbytes, dst, dstride, 2); }
extern "C" void x265b200_add_ps(x265b200_ctx* ctx, int w, int h, void* dst, intptr_t dstride, const void* src0, const int16_t* src1, intptr_t sstride0, intptr_t sstride1)
{ if (ctx) host_blockop(ctx, X265B200_BOP_ADD_PS, w, h, src0, sstride0, ctx->pixbytes, src1, sstride1, 2, dst, dstride, ctx->pixbytes); }
extern "C" void x265b200_pixelavg_pp(x265b200_ctx* ctx, int w, int h, void* dst, intptr_t dstride, const void* src0, intptr_t sstride0, const void* src1, intptr_t sstride1, int)
{ if (ctx) host_blockop(ctx, X265B200_BOP_PIXELAVG, w, h, src0, sstride0, ctx->pixbytes, src1, sstride1, ctx->pixbytes, dst, dstride, ctx->pixbytes); }
extern "C" void x265b200_addAvg(x265b200_ctx* ctx, int w, int h, const int16_t* src0, const int16_t* src1, void* dst, intptr_t src0Stride, intptr_t src1Stride, intptr_t dstStride)
{ if (ctx) host_blockop(ctx, X265B200_BOP_ADDAVG, w, h, src0, src0Stride, 2, src1, src1Stride, 2, dst, dstStride, ctx->pixbytes); }

extern "C" void x265b200_frame_init_lowres(x265b200_ctx* ctx, const void* src0, void* dst0, void* dsth, void* dstv, void* dstc,
                                           intptr_t srcStride, intptr_t dstStride, int width, int height)
{
    Call c(ctx);
    if (!c.ok()) return;
    const size_t pb = ctx->pixbytes;
    // the reference reads 2 * width + 1 columns of 2 * height + 1 rows (pixel.cpp:600-612)
    const int pw = 2 * width + 1, ph = 2 * height + 1;
    size_t oS = c.alloc((size_t)pw * ph * pb);
    size_t inEnd = c.used;
    size_t plane = (size_t)width * height * pb;
    size_t oD = c.alloc(4 * plane);
    if (!c.ok()) return;
    pack(c.h<void>(oS), src0, ph, pw * pb, srcStride * (intptr_t)pb);
    if (!c.upload(0, inEnd)) return;
    uint8_t* dd = c.d<uint8_t>(oD);
    if (x265b200_lowres_batch(ctx, c.d<void>(oS), pw, dd, dd + plane, dd + 2 * plane, dd + 3 * plane, width, width, height, c.st()) != X265B200_OK) return;
    if (!c.download(oD, 4 * plane)) return;
    uint8_t* hd = c.h<uint8_t>(oD);
    unpack(dst0, hd, height, width * pb, dstStride * (intptr_t)pb);
    unpack(dsth, hd + plane, height, width * pb, dstStride * (intptr_t)pb);
    unpack(dstv, hd + 2 * plane, height, width * pb, dstStride * (intptr_t)pb);
    unpack(dstc, hd + 3 * plane, height, width * pb, dstStride * (intptr_t)pb);
}

// ------------------------------------------------------------------ host entries: SEA integral rows (integral.cu)

// integralh_t (framefilter.cpp:39-103): sum[x] = hsum_W(pix, x) + sum[x - stride] for x < stride - W
extern "C" void x265b200_integral_inith(x265b200_ctx* ctx, int W, uint32_t* sum, const void* pix, intptr_t stride)
{
    Call c(ctx);
    if (!c.ok()) return;
    const int count = (int)stride - W;
    if (count <= 0) return;
    const size_t pb = ctx->pixbytes;
    size_t oP = c.alloc((size_t)stride * pb), oA = c.alloc((size_t)count * 4);
    size_t inEnd = c.used;
    size_t oD = c.alloc((size_t)count * 4);
    if (!c.ok()) return;
    memcpy(c.h<void>(oP), pix, (size_t)stride * pb);
    memcpy(c.h<void>(oA), sum - stride, (size_t)count * 4);
    if (!c.upload(0, inEnd)) return;
    if (x265b200_integral_row_batch(ctx, 0, W, c.d<void>(oP), c.d<uint32_t>(oA), nullptr, c.d<uint32_t>(oD), count, c.st()) != X265B200_OK) return;
    if (!c.download(oD, (size_t)count * 4)) return;
    memcpy(sum, c.h<void>(oD), (size_t)count * 4);
}

// integralv_t (framefilter.cpp:106-140): sum[x] = sum[x + H * stride] - sum[x] for x < stride
extern "C" void x265b200_integral_initv(x265b200_ctx* ctx, int H, uint32_t* sum, intptr_t stride)
{
    Call c(ctx);
    if (!c.ok()) return;
    const int count = (int)stride;
    if (count <= 0) return;
    size_t oA = c.alloc((size_t)count * 4), oB = c.alloc((size_t)count * 4);
    size_t inEnd = c.used;
    size_t oD = c.alloc((size_t)count * 4);
    if (!c.ok()) return;
    memcpy(c.h<void>(oA), sum, (size_t)count * 4);
    memcpy(c.h<void>(oB), sum + (intptr_t)H * stride, (size_t)count * 4);
    if (!c.upload(0, inEnd)) return;
    if (x265b200_integral_row_batch(ctx, 1, H, nullptr, c.d<uint32_t>(oA), c.d<uint32_t>(oB), c.d<uint32_t>(oD), count, c.st()) != X265B200_OK) return;
    if (!c.download(oD, (size_t)count * 4)) return;
    memcpy(sum, c.h<void>(oD), (size_t)count * 4);
}

// ------------------------------------------------------------------ host entries: weighted prediction (weight.cu)

static void host_weight(x265b200_ctx* ctx, int sp, const void* src, void* dst, intptr_t srcStride, intptr_t dstStride, int width, int height,
                        int w0, int round, int shift, int offset)
{
    Call c(ctx);
    if (!c.ok()) return;
    const size_t sb = sp ? 2 : ctx->pixbytes, db = ctx->pixbytes;
    size_t oS = c.alloc((size_t)width * height * sb);
    size_t inEnd = c.used;
    size_t oD = c.alloc((size_t)width * height * db);
    if (!c.ok()) return;
    pack(c.h<void>(oS), src, height, width * sb, srcStride * (intptr_t)sb);
    if (!c.upload(0, inEnd)) return;
    if (x265b200_weight_batch(ctx, sp, c.d<void>(oS), width, c.d<void>(oD), width, width, height, w0, round, shift, offset, c.st()) != X265B200_OK) return;
    if (!c.download(oD, (size_t)width * height * db)) return;
    unpack(dst, c.h<void>(oD), height, width * db, dstStride * (intptr_t)db);
}
extern "C" void x265b200_weight_pp(x265b200_ctx* ctx, const void* src, void* dst, intptr_t stride, int width, int height, int w0, int round, int shift, int offset)
{ if (ctx) host_weight(ctx, 0, src, dst, stride, stride, width, height, w0, round, shift, offset); }
extern "C" void x265b200_weight_sp(x265b200_ctx* ctx, const int16_t* src, void* dst, intptr_t srcStride, intptr_t dstStride, int width, int height, int w0, int round, int shift, int offset)
{ if (ctx) host_weight(ctx, 1, src, dst, srcStride, dstStride, width, height, w0, round, shift, offset); }

// ------------------------------------------------------------------ host entries: copy family (blockops.cu)

extern "C" void x265b200_blockcopy(x265b200_ctx* ctx, int kind, int w, int h, void* dst, intptr_t dstStride, const void* src, intptr_t srcStride, int param)
{
    Call c(ctx);
    if (!c.ok()) return;
    const size_t pb = ctx->pixbytes;
    const size_t sb = (kind == 0 || kind == 3) ? pb : 2, db = (kind == 0 || kind == 2) ? pb : 2;
    size_t oS = c.alloc((size_t)w * h * sb);
    size_t inEnd = c.used;
    size_t oD = c.alloc((size_t)w * h * db);
    if (!c.ok()) return;
    if (kind != 4)
    {
        pack(c.h<void>(oS), src, h, w * sb, srcStride * (intptr_t)sb);
        if (!c.upload(0, inEnd)) return;
    }
    if (x265b200_blockcopy_batch(ctx, kind, w, h, c.d<void>(oS), w, nullptr, c.d<void>(oD), w, nullptr, 1, param, c.st()) != X265B200_OK) return;
    if (!c.download(oD, (size_t)w * h * db)) return;
    unpack(dst, c.h<void>(oD), h, w * db, dstStride * (intptr_t)db);
}

// ------------------------------------------------------------------ host entries: per-block scalars (blockstats.cu)

extern "C" uint64_t x265b200_var(x265b200_ctx* ctx, int size, const void* pix, intptr_t stride)
{
    Call c(ctx);
    if (!c.ok()) return 0;
    const size_t pb = ctx->pixbytes;
    size_t oS = c.alloc((size_t)size * size * pb);
    size_t inEnd = c.used;
    size_t oD = c.alloc(8);
    if (!c.ok()) return 0;
    pack(c.h<void>(oS), pix, size, size * pb, stride * (intptr_t)pb);
    if (!c.upload(0, inEnd)) return 0;
    if (x265b200_var_batch(ctx, size, c.d<void>(oS), size, nullptr, 1, c.d<uint64_t>(oD), c.st()) != X265B200_OK) return 0;
    if (!c.download(oD, 8)) return 0;
    return *c.h<uint64_t>(oD);
}

extern "C" int x265b200_psy_cost_pp(x265b200_ctx* ctx, int size, const void* source, intptr_t sstride, const void* recon, intptr_t rstride)
{
    Call c(ctx);
    if (!c.ok()) return 0;
    const size_t pb = ctx->pixbytes;
    size_t oS = c.alloc((size_t)size * size * pb), oR = c.alloc((size_t)size * size * pb);
    size_t inEnd = c.used;
    size_t oD = c.alloc(4);
    if (!c.ok()) return 0;
    pack(c.h<void>(oS), source, size, size * pb, sstride * (intptr_t)pb);
    pack(c.h<void>(oR), recon, size, size * pb, rstride * (intptr_t)pb);
    if (!c.upload(0, inEnd)) return 0;
    if (x265b200_psy_cost_batch(ctx, size, c.d<void>(oS), size, nullptr, c.d<void>(oR), size, nullptr, 1, c.d<int32_t>(oD), c.st()) != X265B200_OK) return 0;
    if (!c.download(oD, 4)) return 0;
    return *c.h<int32_t>(oD);
}

// count_nonzero (coeff == NULL, contiguous size x size input) and copy_cnt (strided residual -> contiguous coeff)
extern "C" uint32_t x265b200_copy_cnt(x265b200_ctx* ctx, int size, int16_t* coeff, const int16_t* residual, intptr_t resiStride)
{
    Call c(ctx);
    if (!c.ok()) return 0;
    size_t bytes = (size_t)size * size * 2;
    size_t oS = c.alloc(bytes);
    size_t inEnd = c.used;
    size_t oC = c.alloc(bytes), oD = c.alloc(4);
    if (!c.ok()) return 0;
    pack(c.h<void>(oS), residual, size, (size_t)size * 2, resiStride * 2);
    if (!c.upload(0, inEnd)) return 0;
    if (x265b200_count_nonzero_batch(ctx, size, c.d<int16_t>(oS), size, nullptr, 1, coeff ? c.d<int16_t>(oC) : nullptr, c.d<uint32_t>(oD), c.st()) != X265B200_OK) return 0;
    if (!c.download(oC, (oD + 4) - oC)) return 0;
    if (coeff) memcpy(coeff, c.h<void>(oC), bytes);
    return *c.h<uint32_t>(oD);
}

extern "C" void x265b200_denoise_dct(x265b200_ctx* ctx, int16_t* dctCoef, uint32_t* resSum, const uint16_t* offset, int numCoeff)
{
    Call c(ctx);
    if (!c.ok()) return;
    size_t oC = c.alloc((size_t)numCoeff * 2), oR = c.alloc((size_t)numCoeff * 4), oO = c.alloc((size_t)numCoeff * 2);
    if (!c.ok()) return;
    memcpy(c.h<void>(oC), dctCoef, (size_t)numCoeff * 2);
    memcpy(c.h<void>(oR), resSum, (size_t)numCoeff * 4);
    memcpy(c.h<void>(oO), offset, (size_t)numCoeff * 2);
    if (!c.upload(0, c.used)) return;
    if (x265b200_denoise_dct_batch(ctx, c.d<int16_t>(oC), c.d<uint32_t>(oR), c.d<uint16_t>(oO), numCoeff, 1, c.st()) != X265B200_OK) return;
    if (!c.download(0, oO)) return;
    memcpy(dctCoef, c.h<void>(oC), (size_t)numCoeff * 2);
    memcpy(resSum, c.h<void>(oR), (size_t)numCoeff * 4);
}

// ------------------------------------------------------------------ host entries: intra prediction slots (intra.cu)

static void host_intra(x265b200_ctx* ctx, int kind, int N, int mode, int bFilter, const void* src, const void* filt, void* dst, intptr_t dstStride)
{
    Call c(ctx);
    if (!c.ok()) return;
    const size_t pb = ctx->pixbytes, L = (size_t)(4 * N + 1) * pb;
    const size_t outElems = kind == 1 ? (size_t)(4 * N + 1) : kind == 0 ? (size_t)N * N : (size_t)33 * N * N;
    size_t oS = c.alloc(L), oF = c.alloc(L);
    size_t inEnd = c.used;
    size_t oD = c.alloc(outElems * pb);
    if (!c.ok()) return;
    memcpy(c.h<void>(oS), src, L);
    if (filt) memcpy(c.h<void>(oF), filt, L);
    if (!c.upload(0, inEnd)) return;
    if (x265b200_intra_slot_batch(ctx, kind, N, mode, bFilter, c.d<void>(oS), filt ? c.d<void>(oF) : nullptr, 1, c.d<void>(oD), c.st()) != X265B200_OK) return;
    if (!c.download(oD, outElems * pb)) return;
    if (kind == 0) unpack(dst, c.h<void>(oD), N, N * pb, dstStride * (intptr_t)pb);
    else memcpy(dst, c.h<void>(oD), outElems * pb);
}
extern "C" void x265b200_intra_pred(x265b200_ctx* ctx, int N, int mode, void* dst, intptr_t dstStride, const void* srcPix, int bFilter)
{ if (ctx) host_intra(ctx, 0, N, mode, bFilter, srcPix, nullptr, dst, dstStride); }
extern "C" void x265b200_intra_filter(x265b200_ctx* ctx, int N, const void* samples, void* filtered)
{ if (ctx) host_intra(ctx, 1, N, 0, 0, samples, nullptr, filtered, 0); }
extern "C" void x265b200_intra_pred_allangs(x265b200_ctx* ctx, int N, void* dst, const void* refPix, const void* filtPix, int bLuma)
{ if (ctx) host_intra(ctx, 2, N, 0, bLuma, refPix, filtPix, dst, 0); }
