// tu_umma.cuh -- the inter-luma TU reconstruction chain as ONE kernel on the 5th-generation tensor cores (tcgen05.mma kind::i8,
// accumulators in tensor memory), N = 32 and 16.  Included at the end of transform_mma.cu, inside namespace b200, after tu_fused.cuh
// (it reuses that file's quant / dequant / DC-fill helpers).  Reference call sequence: encoder/search.cpp:5536-5575 through
// common/quant.cpp:397-480 (transformNxN) and :543-605 (invtransformNxN); transforms common/dct.cpp:83-611.
//
// Why tcgen05 here.  The mma.sync version (tu_fused.cuh) keeps every accumulator and operand fragment in registers: 106-128 registers per
// thread, 4 CTAs per SM, and it needs two kernels (qCoef goes to HBM and comes back, fenc / pred are read twice: 14 bytes per sample for an
// 8-byte algorithm).  With the accumulators in TMEM a thread only ever holds ONE row of ONE matrix, so all four transform stages, quant,
// dequant and the reconstruction fit one kernel that reads fenc + pred once and writes qCoef + recon once.
//
// Shape of the work.  A CTA (128 threads) owns 128 / N TUs at a time; thread t = row (t % N) of TU (t / N).  Each transform stage is
//      D[128 x N] = A[128 x 32] * B[N x 32]^T        (tcgen05.mma cta_group::1 kind::i8, M = 128, K = 32, one instruction per byte plane)
// with the int16 operand split into a signed high-byte plane and an unsigned low-byte plane (two MMAs, exact: |sum| < 2^31), A written
// to shared memory by the threads in the canonical no-swizzle layout (8 x 16-byte core matrices), B the constant transform matrix.
//   forward 1:  A = residual rows            (K-major: a thread writes its own row contiguously)      B = T      -> Z[j][k]   row j
//   forward 2:  A = Z^T                      (MN-major: the same contiguous write, read transposed)    B = T      -> C[k2][k]  COLUMN k
//   inverse 1:  A = dequantised C^T          (K-major: the thread's column is a row of C^T)            B = T^T    -> tmp[j][n] row j
//   inverse 2:  A = tmp^T                    (MN-major)                                                B = T^T    -> resi[a][b] row a
// so every stage hands each thread exactly the vector the next stage wants it to write, the final residual row meets the prediction row the
// thread loaded at the start, and no shuffle or transpose instruction is issued.  tcgen05.ld (32 lanes x 32 bit, one row per thread)
// brings the accumulators back; rounding shifts, saturation, quant and dequant run on them in registers.
//
// Bit-exactness: same integer arithmetic as dct.cpp (full-matrix form of the partial butterflies, rounding shift per stage, int16
// truncation after the forward stages, clip3(-32768, 32767) after the inverse stages), quant / dequant as dct.cpp:614-688, the cbf == 0
// and DC-only reconstruction shortcuts of quant.cpp:543-605.

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle (tools/umma_probe.cu verified the field meaning for K-major and MN-major operands)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor: D = S32, A / B = 8-bit (signed flags), majors, N, M = 128
__device__ __forceinline__ uint32_t make_idesc(int N, int aSigned, int aMN)
{
    return (2u << 4) | ((uint32_t)aSigned << 7) | (1u << 10) | ((uint32_t)aMN << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t phase)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

} // namespace umma

// c_ummaB (transform_mma.cu): constant B tiles in the canonical K-major layout, [size: 0 = 32, 1 = 16][0 = T (forward), 1 = T^T (inverse)][1024 bytes],
// filled by upload_mma_tables (rows n >= N and columns k >= N are zero)

// One transform stage for the CTA's 128 rows: the threads have written both byte planes of A; run the two MMAs, wait, and hand every
// thread its row of the accumulator as N int32 values (high plane * 256 + low plane).
template<int N, int LO>
__device__ __forceinline__ void umma_stage(uint32_t aHi, uint32_t aLo, int aMN, uint32_t bAddr, uint32_t tmem, uint32_t bar, uint32_t& phase, int (&v)[N])
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // the A planes were written through the generic proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");      // earlier tcgen05.ld of this accumulator are done
    __syncthreads();
    if (threadIdx.x < 32)
    {
        if (threadIdx.x == 0)
        {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // K-major: 16-byte K chunks 128 B apart, 8-row groups 256 B apart.  MN-major: 8-row K groups 1024 B apart, 16-element MN chunks 128 B apart.
            const uint32_t lbo = aMN ? 1024 : 128, sbo = aMN ? 128 : 256;
            const uint64_t db = umma::make_desc(bAddr, 128, 256);
            umma::mma_i8(tmem, umma::make_desc(aHi, lbo, sbo), db, umma::make_idesc(N, 1, aMN));
            umma::mma_i8(tmem + LO, umma::make_desc(aLo, lbo, sbo), db, umma::make_idesc(N, 0, aMN));
            umma::commit(bar);
        }
        // the issuing lane's 31 neighbours must not start spinning on the barrier while it is still issuing: a spin loop in the same warp
        // can starve the divergent lane for ever (seen as a hang with several CTAs per SM)
        __syncwarp();
    }
    umma::wait_bar(bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t lane0 = tmem + ((uint32_t)(threadIdx.x & ~31) << 16);     // a warp reads its own 32 lanes
#pragma unroll
    for (int c = 0; c < N; c += 16)
    {
        uint32_t h[16], l[16];
        umma::ld16(lane0 + c, h);
        umma::ld16(lane0 + LO + c, l);
        umma::ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i++) v[c + i] = ((int)h[i] << 8) + (int)l[i];
    }
}

// write N int16 (packed pairs w[N / 2]) as the thread's line of both byte planes: K-major -> row `row` of A, MN-major -> K index `row`, MN run
// starting at 16-element chunk `chunk0`
template<int N>
__device__ __forceinline__ void umma_write_kmajor(uint8_t* hi, uint8_t* lo, int row, const uint32_t (&w)[N / 2])
{
    const int base = (row >> 3) * 256 + (row & 7) * 16;
#pragma unroll
    for (int c = 0; c < N / 16; c++)
    {
        uint32_t l[4], h[4];
#pragma unroll
        for (int q = 0; q < 4; q++) split4(make_uint2(w[8 * c + 2 * q], w[8 * c + 2 * q + 1]), l[q], h[q]);
        *(uint4*)(lo + base + c * 128) = make_uint4(l[0], l[1], l[2], l[3]);
        *(uint4*)(hi + base + c * 128) = make_uint4(h[0], h[1], h[2], h[3]);
    }
}
template<int N>
__device__ __forceinline__ void umma_write_mnmajor(uint8_t* hi, uint8_t* lo, int k, int chunk0, const uint32_t (&w)[N / 2])
{
    const int base = (k >> 3) * 1024 + (k & 7) * 16 + chunk0 * 128;
#pragma unroll
    for (int c = 0; c < N / 16; c++)
    {
        uint32_t l[4], h[4];
#pragma unroll
        for (int q = 0; q < 4; q++) split4(make_uint2(w[8 * c + 2 * q], w[8 * c + 2 * q + 1]), l[q], h[q]);
        *(uint4*)(lo + base + c * 128) = make_uint4(l[0], l[1], l[2], l[3]);
        *(uint4*)(hi + base + c * 128) = make_uint4(h[0], h[1], h[2], h[3]);
    }
}

// sum over the N lanes that share a TU (N = 16 or 32)
template<int N> __device__ __forceinline__ unsigned long long tu_sum64(unsigned long long s)
{
#pragma unroll
    for (int m = N >> 1; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    return s;
}
template<int N> __device__ __forceinline__ int tu_sum32(int s)
{
#pragma unroll
    for (int m = N >> 1; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    return s;
}

template<typename T, int N, int MINB, int COLS>
__global__ void __launch_bounds__(128, MINB)
tu_umma_kernel(const T* __restrict__ fenc, intptr_t sf, const T* __restrict__ pred, intptr_t sp, const int32_t* __restrict__ offF,
               const int32_t* __restrict__ offP, int n, const int32_t* __restrict__ quantCoeff, QuantP P, int fshift1, int fshift2, int ishift2,
               int depth, int16_t* __restrict__ qCoef, uint32_t* __restrict__ numSig, T* __restrict__ recon, intptr_t sr,
               const int32_t* __restrict__ offR, unsigned long long* __restrict__ sseZero, unsigned long long* __restrict__ sseRecon, int lab)
{
    constexpr int NT = 128 / N, H = N / 2, NN = N * N;
    __shared__ __align__(128) uint8_t sAk[2][4096];         // K-major A planes: [0] high bytes (s8), [1] low bytes (u8)
    __shared__ __align__(128) uint8_t sAm[2][4096];         // MN-major A planes
    __shared__ __align__(128) uint8_t sB[2][1024];          // T, T^T
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmemBase;
    const int tid = threadIdx.x;

    for (int i = tid; i < 2 * 4096 / 16; i += 128) { ((uint4*)sAk)[i] = make_uint4(0, 0, 0, 0); ((uint4*)sAm)[i] = make_uint4(0, 0, 0, 0); }
    for (int i = tid; i < 2 * 1024 / 16; i += 128) ((uint4*)sB)[i] = ((const uint4*)c_ummaB[N == 32 ? 0 : 1])[i];
    if (tid == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(umma::smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(umma::smem_u32(&tmemBase)), "n"(COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmemBase, barA = umma::smem_u32(&bar);
    const uint32_t aK[2] = { umma::smem_u32(sAk[0]), umma::smem_u32(sAk[1]) }, aM[2] = { umma::smem_u32(sAm[0]), umma::smem_u32(sAm[1]) };
    const uint32_t bF = umma::smem_u32(sB[0]), bI = umma::smem_u32(sB[1]);
    uint32_t phase = 0;

    const int row = tid % N, tl = tid / N;                  // this thread's line inside its TU, the TU's slot in the group
    const int chunk0 = tl * (N / 16);                       // first 16-element MN chunk of the TU's rows
    const uint32_t mx = ((uint32_t)((1 << depth) - 1)) * 0x10001u, negmx = (0u - (uint32_t)((1 << depth) - 1)) & 0xffffu;
    const uint32_t negmx2 = negmx | (negmx << 16);
    const int ngroups = (n + NT - 1) / NT;

    for (int g = blockIdx.x; g < ngroups; g += gridDim.x)
    {
        const int tu = g * NT + tl;
        const bool live = tu < n;
        // ---- this thread's fenc and prediction rows (packed sample pairs), the residual row, sse(fenc, pred)
        uint32_t f[H], p[H], r[H];
        if (live && !(lab & 1))
        {
            load_row_quads<N / 4>(fenc + offF[tu] + (intptr_t)row * sf, f);
            load_row_quads<N / 4>(pred + offP[tu] + (intptr_t)row * sp, p);
        }
        else
        {
#pragma unroll
            for (int i = 0; i < H; i++) { f[i] = 0; p[i] = 0; }
        }
        uint32_t z32 = 0;
#pragma unroll
        for (int i = 0; i < H; i += 2)
        {
            r[i] = psub16(f[i], p[i]); r[i + 1] = psub16(f[i + 1], p[i + 1]);
            z32 += sumsq4(make_uint2(r[i], r[i + 1]));
        }
        // ---- forward stage 1: Z[j][k] = (sum_x resi[j][x] T[k][x] + add) >> shift1
        int v[N];
        umma_write_kmajor<N>(sAk[0], sAk[1], tid, r);
        umma_stage<N, COLS / 2>(aK[0], aK[1], 0, bF, tmem, barA, phase, v);
        uint32_t w[H];
        {
            const int add = 1 << (fshift1 - 1);
#pragma unroll
            for (int i = 0; i < H; i++) w[i] = __byte_perm((v[2 * i] + add) >> fshift1, (v[2 * i + 1] + add) >> fshift1, 0x5410);
        }
        // ---- forward stage 2: C[k2][k] = (sum_j T[k2][j] Z[j][k] + add) >> shift2; this thread receives COLUMN k = row
        umma_write_mnmajor<N>(sAm[0], sAm[1], row, chunk0, w);
        umma_stage<N, COLS / 2>(aM[0], aM[1], 1, bF, tmem, barA, phase, v);
        // ---- quant (dct.cpp:666-688 without deltaU) and dequant_normal (dct.cpp:614-636) down the column
        int sig = 0, lvDC = 0;
        {
            const int add = 1 << (fshift2 - 1);
            int16_t* qTu = qCoef + (size_t)tu * NN + row;
#pragma unroll
            for (int i = 0; i < H; i++)
            {
                int lv0, lv1;
                const int c0 = (int)(int16_t)((v[2 * i] + add) >> fshift2), c1 = (int)(int16_t)((v[2 * i + 1] + add) >> fshift2);
                const int q0 = (live && !(lab & 16)) ? __ldg(quantCoeff + (2 * i) * N + row) : 0, q1 = (live && !(lab & 16)) ? __ldg(quantCoeff + (2 * i + 1) * N + row) : 0;
                sig += quant_one(c0, q0, P, lv0) + quant_one(c1, q1, P, lv1);
                if (i == 0) lvDC = lv0;                     // level (0, row): the DC level in the TU's first thread
                const uint32_t lvp = pack_sat_s16(lv0, lv1);
                if (live && !(lab & 2)) { qTu[(2 * i) * N] = (int16_t)(lvp & 0xffff); qTu[(2 * i + 1) * N] = (int16_t)(lvp >> 16); }
                w[i] = dequant_pair(lvp, P);
            }
        }
        const int ns = tu_sum32<N>(sig);
        // the DC coefficient is (k2 = 0, k = 0): the TU's first thread holds its level in lvDC and its dequantised value in w[0]'s low half
        const int first = (tid & 31) & ~(N - 1);
        const int dq0 = (int)(int16_t)(__shfl_sync(0xffffffffu, w[0], first) & 0xffff);
        const bool dcOnly = ns == 1 && __shfl_sync(0xffffffffu, lvDC, first) != 0;
        // ---- inverse stage 1: tmp[j][n] = clip16((sum_k C[k][j] T[k][n] + 64) >> 7); the thread's column j of C is row j of C^T
        umma_write_kmajor<N>(sAk[0], sAk[1], tid, w);
        umma_stage<N, COLS / 2>(aK[0], aK[1], 0, bI, tmem, barA, phase, v);
#pragma unroll
        for (int i = 0; i < H; i++) w[i] = pack_sat_s16((v[2 * i] + 64) >> 7, (v[2 * i + 1] + 64) >> 7);
        // ---- inverse stage 2: resi[a][b] = clip16((sum_k tmp[k][a] T[k][b] + add) >> shift2); this thread receives row a = row
        umma_write_mnmajor<N>(sAm[0], sAm[1], row, chunk0, w);
        umma_stage<N, COLS / 2>(aM[0], aM[1], 1, bI, tmem, barA, phase, v);
        // ---- reconstruction: cbf == 0 -> prediction; DC only -> flat residual (quant.cpp:588-598); else the inverse transform's row
        uint32_t d32 = 0;
        {
            const int add = 1 << (ishift2 - 1);
            const int dcv = dc_fill_value(dq0, depth);
            T* out = recon + (live ? offR[tu] : 0) + (intptr_t)row * sr;
#pragma unroll
            for (int i = 0; i < H; i += 2)
            {
                uint32_t r0, r1;
                if (dcOnly) { r0 = r1 = __byte_perm(dcv, dcv, 0x5410); }
                else
                {
                    r0 = pack_sat_s16((v[2 * i] + add) >> ishift2, (v[2 * i + 1] + add) >> ishift2);
                    r1 = pack_sat_s16((v[2 * i + 2] + add) >> ishift2, (v[2 * i + 3] + add) >> ishift2);
                }
                uint32_t o0 = p[i], o1 = p[i + 1];
                if (ns)
                {
                    o0 = __viaddmin_s16x2_relu(p[i], __vmins2(__vmaxs2(r0, negmx2), mx), mx);
                    o1 = __viaddmin_s16x2_relu(p[i + 1], __vmins2(__vmaxs2(r1, negmx2), mx), mx);
                }
                if (live && !(lab & 4)) store_pix4(out + 2 * i, o0, o1);
                d32 += sumsq4(make_uint2(psub16(f[i], o0), psub16(f[i + 1], o1)));
            }
        }
        const unsigned long long zs = tu_sum64<N>(z32), ds = tu_sum64<N>(d32);
        if (live && row == 0 && !(lab & 8))
        {
            numSig[tu] = (uint32_t)ns;
            if (sseZero) sseZero[tu] = zs;
            sseRecon[tu] = ds;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(COLS));
}

// the whole chain over n TUs of size N (32 or 16) in one launch.  Returns false when nothing was launched (alignment / size not covered).
bool launch_tu_umma(x265b200_ctx* ctx, int N, const void* fenc, intptr_t sf, const void* pred, intptr_t sp,
                    const int32_t* offF, const int32_t* offP, int n, const int32_t* quantCoeff, int qBits, int qAdd,
                    int dqScale, int dqShift, int16_t* qCoef, uint32_t* numSig, void* recon, intptr_t sr,
                    const int32_t* offR, uint64_t* sseZero, uint64_t* sseRecon, cudaStream_t st)
{
    if ((N != 32 && N != 16) || ((sf | sp) & 3)) return false;
    if (N == 16 && !getenv("X265B200_UMMA_LAB")) return false;      // N = 16 is held back until its multi-CTA hang is understood (tools/umma_debug.py)
    const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
    const int lg = N == 32 ? 5 : 4, d8 = ctx->depth - 8;
    QuantP P; P.qBits = qBits; P.qAdd = qAdd; P.dqScale = dqScale; P.dqAdd = 1 << (dqShift - 1); P.dqShift = dqShift;
    const int ngroups = (n + 128 / N - 1) / (128 / N);
    constexpr int RES = 4;                                    // resident CTAs per SM the launch bounds ask for
    int grid = sms * RES;
    int pad = 0, lab = 0;                                     // lab: X265B200_UMMA_LAB="ctas_per_sm,dynamic_smem_bytes,skip_mask"
    if (const char* e = getenv("X265B200_UMMA_LAB")) { int c = RES; sscanf(e, "%d,%d,%d", &c, &pad, &lab); grid = sms * c; }
    if (grid > ngroups) grid = ngroups;
#define UM(T, N_) tu_umma_kernel<T, N_, RES, 64><<<grid, 128, pad, st>>>((const T*)fenc, sf, (const T*)pred, sp, offF, offP, n, quantCoeff, P, lg - 1 + d8, lg + 6, 12 - d8, \
                      ctx->depth, qCoef, numSig, (T*)recon, sr, offR, (unsigned long long*)sseZero, (unsigned long long*)sseRecon, lab)
    if (ctx->pixbytes == 1) { if (N == 32) UM(uint8_t, 32); else UM(uint8_t, 16); }
    else { if (N == 32) UM(uint16_t, 32); else UM(uint16_t, 16); }
#undef UM
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError() == cudaSuccess;
}
