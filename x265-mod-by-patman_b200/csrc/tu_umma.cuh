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
//   forward 2:  A = diag(T, T, ..)  K = 128  (constant, one 128 x 32 tile per K step)   B = the Z rows of all TUs, MN-major   -> C[k2][k] row k2
//   inverse 1:  A = dequantised C^T          (MN-major: the thread's row of C, read transposed)        B = T^T    -> tmp[j][n] row j
//   inverse 2:  A = tmp^T                    (MN-major)                                                B = T^T    -> resi[a][b] row a
// so every stage hands each thread a ROW: of Z, of the coefficient block (quant table, qCoef store and dequant run along contiguous rows:
// 16-byte loads and stores), of the intermediate and finally of the residual, which meets the prediction row the thread loaded at the
// start.  No shuffle or transpose instruction is issued; the transposes are the operand layouts.  Forward stage 2 contracts over the rows
// of Z, which belong to different TUs in different lanes: the block-diagonal constant A keeps the TUs apart (4 accumulating MMAs).  tcgen05.ld (32 lanes x 32 bit, one row per thread)
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
// instruction descriptor: D = S32, A / B = 8-bit (signed flags), majors (0 = K, 1 = MN), N, M = 128
__device__ __forceinline__ uint32_t make_idesc(int N, int aSigned, int bSigned, int aMN, int bMN)
{
    return (2u << 4) | ((uint32_t)aSigned << 7) | ((uint32_t)bSigned << 10) | ((uint32_t)aMN << 15) | ((uint32_t)bMN << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, int accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
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
// lab (-DB200_UMMA_TRACE): per-warp progress markers in mapped host memory, readable while a kernel hangs (tools/umma_debug.py found the
// half-warp shuffle deadlock of the first N = 16 version with them)
#ifdef B200_UMMA_TRACE
__device__ int* g_ummaTrace;
#define UMMA_TRACE(code) do { int* t__ = g_ummaTrace; if (t__ && (threadIdx.x & 31) == 0) { ((volatile int*)t__)[blockIdx.x * 4 + (threadIdx.x >> 5)] = (code); __threadfence_system(); } } while (0)
#else
#define UMMA_TRACE(code) do { } while (0)
#endif

// One transform stage for the CTA's 128 rows: the threads have written both byte planes of the data operand; issue the MMAs, wait, and
// hand every thread its row of the accumulator as N int32 values (high plane * 256 + low plane).
//   mode 0: A = data, K-major      B = constant tile bAddr            (forward 1)
//   mode 1: A = data, MN-major     B = constant tile bAddr            (inverse 1 and 2)
//   mode 2: A = constant block-diagonal tiles at aConst (4 K steps), B = data, MN-major (forward 2)
template<int N, int LO>
__device__ __forceinline__ void umma_stage(int stageId, int mode, uint32_t dHi, uint32_t dLo, uint32_t cAddr, uint32_t tmem, uint32_t bar, uint32_t& phase, int (&v)[N])
{
    UMMA_TRACE(stageId * 10 + 1);
    // every tcgen05 instruction below is .sync.aligned: the warp must arrive converged
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // the data planes were written through the generic proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");      // earlier tcgen05.ld of this accumulator are done
    __syncthreads();
    UMMA_TRACE(stageId * 10 + 2);
    if (threadIdx.x < 32)
    {
        if (threadIdx.x == 0)
        {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (mode != 2)
            {
                // K-major: 16-byte K chunks 128 B apart, 8-row groups 256 B apart.  MN-major: 8-row K groups 1024 B apart, 16-element MN chunks 128 B apart.
                const uint32_t lbo = mode ? 1024 : 128, sbo = mode ? 128 : 256;
                // the inverse stages want B[n][k] = T[k][n]: the forward tile (T, K-major) read as an MN-major operand -- byte (k, n) of the K-major
                // layout, (k / 8) * 256 + (n / 16) * 128 + (k % 8) * 16 + n % 16, is the MN-major address with K groups 256 B and MN chunks 128 B apart
                const uint64_t db = mode ? umma::make_desc(cAddr, 256, 128) : umma::make_desc(cAddr, 128, 256);
                umma::mma_i8(tmem, umma::make_desc(dHi, lbo, sbo), db, umma::make_idesc(N, 1, 1, mode, mode), 0);
                umma::mma_i8(tmem + LO, umma::make_desc(dLo, lbo, sbo), db, umma::make_idesc(N, 0, 1, mode, mode), 0);
            }
            else
            {
                // data = B[n = column k][K = the CTA's 128 Z rows], MN-major: K groups of 8 rows (N / 16) * 128 B apart, 16-column chunks 128 B apart
                constexpr uint32_t lbo = (N / 16) * 128;
#pragma unroll
                for (int ks = 0; ks < 4; ks++)
                {
                    const uint64_t da = umma::make_desc(cAddr + (12 - 4 * ks) * 256, 128, 256);
                    umma::mma_i8(tmem, da, umma::make_desc(dHi + ks * 4 * lbo, lbo, 128), umma::make_idesc(N, 1, 1, 0, 1), ks);
                    umma::mma_i8(tmem + LO, da, umma::make_desc(dLo + ks * 4 * lbo, lbo, 128), umma::make_idesc(N, 1, 0, 0, 1), ks);
                }
            }
            umma::commit(bar);
        }
        // the issuing lane's neighbours wait here, not in the barrier's spin loop
        __syncwarp();
    }
    UMMA_TRACE(stageId * 10 + 3);
    umma::wait_bar(bar, phase);
    UMMA_TRACE(stageId * 10 + 4);
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
    UMMA_TRACE(stageId * 10 + 5);
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

// forward stage 2's data operand: B[n = 0 .. N-1][K = tid], MN-major with K groups (N / 16) * 128 bytes apart
template<int N>
__device__ __forceinline__ void umma_write_bdata(uint8_t* hi, uint8_t* lo, int tid, const uint32_t (&w)[N / 2])
{
    const int base = (tid >> 3) * ((N / 16) * 128) + (tid & 7) * 16;
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

// ---- coalesced I/O for the row-per-thread kernel (16-bit samples) -------------------------------------------------------------------
// A thread owns a row, and rows of a picture lie a whole stride apart: a warp-level load of "my row, bytes 0-15" touches 32 different
// 32-byte sectors and uses a third of each (ncu on the first version: 28.6 sectors per request, LSU wavefronts 83 % busy).  So the warp
// moves its 32 rows cooperatively -- lane = (row, 16-byte piece), PPR pieces per row, consecutive lanes on consecutive pieces -- through
// a tile in shared memory, and each thread then reads / writes its own row there.  Piece c of row r sits at r * RB + (c ^ swz(r)) * 16:
// the xor keeps both the row owners (stride RB) and the cooperative lanes (consecutive pieces) free of bank conflicts.
template<int N> struct WarpTile
{
    static constexpr int RB = 2 * N;                // bytes per row of 16-bit samples
    static constexpr int PPR = RB / 16;             // 16-byte pieces per row: 4 (N = 32) or 2 (N = 16)
    static constexpr int RPI = 32 / PPR;            // rows per cooperative instruction
    static constexpr int NI = 32 / RPI;             // cooperative instructions per tile
    static __device__ __forceinline__ int swz(int r) { return N == 32 ? (r >> 1) & 3 : (r >> 2) & 1; }
    static __device__ __forceinline__ uint4* piece(uint8_t* tile, int r, int c) { return (uint4*)(tile + r * RB + ((c ^ swz(r)) << 4)); }
};

template<typename T, int N, int MINB, int COLS>
__global__ void __launch_bounds__(128, MINB)
tu_umma_kernel(const T* __restrict__ fenc, intptr_t sf, const T* __restrict__ pred, intptr_t sp, const int32_t* __restrict__ offF,
               const int32_t* __restrict__ offP, int n, const int32_t* __restrict__ quantCoeff, QuantP P, int fshift1, int fshift2, int ishift2,
               int depth, int16_t* __restrict__ qCoef, uint32_t* __restrict__ numSig, T* __restrict__ recon, intptr_t sr,
               const int32_t* __restrict__ offR, unsigned long long* __restrict__ sseZero, unsigned long long* __restrict__ sseRecon, int lab)
{
    constexpr int NT = 128 / N, H = N / 2, NN = N * N;
    // dynamic shared memory (more than the 48 KB a static allocation may take), carved by hand; every piece is 128-byte aligned
    extern __shared__ __align__(1024) uint8_t umma_smem[];
    uint8_t (*sAk)[4096] = (uint8_t (*)[4096])umma_smem;                        // [2] K-major data planes (forward 1): [0] high bytes (s8), [1] low bytes (u8)
    uint8_t (*sAm)[4096] = (uint8_t (*)[4096])(umma_smem + 8192);               // [2] MN-major data planes (forward 2's B, inverse 1 / 2's A)
    uint8_t* sB = umma_smem + 16384;                                            // T as the B operand: K-major for the forward, read MN-major for the inverse
    // forward 2's A operand diag(T, T, ..): tile ks (128 rows x 32 K bytes, K-major) is zero except for the four 8-row groups of the rows whose
    // TUs own K step ks.  One buffer [12 zero groups][4 groups: the non-zero block][12 zero groups] serves all four tiles: tile ks starts
    // (12 - 4 ks) groups into it.
    uint8_t* sTd = umma_smem + 17408;                                           // 28 groups x 256 bytes = 7168
    int32_t* sQ = (int32_t*)(umma_smem + 24576);                                // quant table rows, 16-byte pieces xor-swizzled by row (4 KB)
    uint8_t (*sTile)[2][64 * N] = (uint8_t (*)[2][64 * N])(umma_smem + 28672);  // [4 warps][fenc, prediction] rows of 16-bit samples
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmemBase;
    const int tid = threadIdx.x;

    for (int i = tid; i < 2 * 4096 / 16; i += 128) { ((uint4*)sAk)[i] = make_uint4(0, 0, 0, 0); ((uint4*)sAm)[i] = make_uint4(0, 0, 0, 0); }
    for (int i = tid; i < 1024 / 16; i += 128) ((uint4*)sB)[i] = ((const uint4*)c_ummaB[N == 32 ? 0 : 1][0])[i];
    for (int i = tid; i < 7168 / 16; i += 128) ((uint4*)sTd)[i] = ((const uint4*)c_ummaAD[N == 32 ? 0 : 1])[i];
    for (int i = tid; i < N * N / 4; i += 128)
    {   // quant table row r = i / (N / 4), piece c: stored at piece c ^ (r & (N / 4 - 1)) so that the row owners' 16-byte reads spread over the banks
        const int r = i / (N / 4), c = i % (N / 4);
        ((int4*)sQ)[r * (N / 4) + (c ^ (r & (N / 4 - 1)))] = __ldg((const int4*)quantCoeff + i);
    }
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
    const uint32_t bF = umma::smem_u32(sB), bI = bF, aD = umma::smem_u32(sTd);
    uint32_t phase = 0;

    const int row = tid % N, tl = tid / N;                  // this thread's line inside its TU, the TU's slot in the group
    const int chunk0 = tl * (N / 16);                       // first 16-element MN chunk of the TU's rows
    const uint32_t mx = ((uint32_t)((1 << depth) - 1)) * 0x10001u, negmx = (0u - (uint32_t)((1 << depth) - 1)) & 0xffffu;
    const uint32_t negmx2 = negmx | (negmx << 16);
    const int ngroups = (n + NT - 1) / NT;

    UMMA_TRACE(1);
    for (int g = blockIdx.x; g < ngroups; g += gridDim.x)
    {
        const int tu = g * NT + tl;
        const bool live = tu < n;
        // ---- this thread's fenc and prediction rows (packed sample pairs), the residual row, sse(fenc, pred)
        uint32_t f[H], p[H], w[H];
        const int lane = tid & 31, warp = tid >> 5;
        const int myOffF = live ? offF[tu] : 0, myOffP = live ? offP[tu] : 0;
        if constexpr (sizeof(T) == 2)
        {
            typedef WarpTile<N> WT;
            uint8_t* tF = sTile[warp][0];
            uint8_t* tP = sTile[warp][1];
            __syncwarp();                                   // the previous group's readers of these tiles are done
#pragma unroll
            for (int it = 0; it < WT::NI; it++)
            {
                const int R = it * WT::RPI + lane / WT::PPR, c = lane % WT::PPR;       // the row this lane fetches a piece of, and which piece
                const int oF = __shfl_sync(0xffffffffu, myOffF, R), oP = __shfl_sync(0xffffffffu, myOffP, R);
                const bool rl = __shfl_sync(0xffffffffu, (int)live, R) != 0;
                uint4 vf = make_uint4(0, 0, 0, 0);
                uint32_t vp[4] = { 0, 0, 0, 0 };
                if (rl && !(lab & 1))
                {
                    const int rr = R % N;
                    vf = __ldg((const uint4*)(fenc + oF + (intptr_t)rr * sf + 8 * c));      // fenc TUs sit on the TU grid: 16-byte aligned
                    load_row_quads<2>(pred + oP + (intptr_t)rr * sp + 8 * c, vp);         // prediction rows: any 2-byte alignment
                }
                *WT::piece(tF, R, c) = vf;
                *WT::piece(tP, R, c) = make_uint4(vp[0], vp[1], vp[2], vp[3]);
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < WT::PPR; c++)
            {
                const uint4 a = *WT::piece(tF, lane, c), b = *WT::piece(tP, lane, c);
                f[4 * c] = a.x; f[4 * c + 1] = a.y; f[4 * c + 2] = a.z; f[4 * c + 3] = a.w;
                p[4 * c] = b.x; p[4 * c + 1] = b.y; p[4 * c + 2] = b.z; p[4 * c + 3] = b.w;
            }
        }
        else
        {
            if (live && !(lab & 1))
            {
                load_row_quads<N / 4>(fenc + myOffF + (intptr_t)row * sf, f);
                load_row_quads<N / 4>(pred + myOffP + (intptr_t)row * sp, p);
            }
            else
            {
#pragma unroll
                for (int i = 0; i < H; i++) { f[i] = 0; p[i] = 0; }
            }
        }
        uint32_t z32 = 0;
#pragma unroll
        for (int i = 0; i < H; i += 2)
        {
            w[i] = psub16(f[i], p[i]); w[i + 1] = psub16(f[i + 1], p[i + 1]);
            z32 += sumsq4(make_uint2(w[i], w[i + 1]));
        }
        // ---- forward stage 1: Z[j][k] = (sum_x resi[j][x] T[k][x] + add) >> shift1, row j = this thread's row
        int v[N];
        umma_write_kmajor<N>(sAk[0], sAk[1], tid, w);
        umma_stage<N, COLS / 2>(1, 0, aK[0], aK[1], bF, tmem, barA, phase, v);
        {
            const int add = 1 << (fshift1 - 1);
#pragma unroll
            for (int i = 0; i < H; i++) w[i] = __byte_perm((v[2 * i] + add) >> fshift1, (v[2 * i + 1] + add) >> fshift1, 0x5410);
        }
        // ---- forward stage 2: C[k2][k] = (sum_j T[k2][j] Z[j][k] + add) >> shift2, row k2 = this thread's row
        umma_write_bdata<N>(sAm[0], sAm[1], tid, w);
        umma_stage<N, COLS / 2>(2, 2, aM[0], aM[1], aD, tmem, barA, phase, v);
        // ---- quant (dct.cpp:666-688 without deltaU) and dequant_normal (dct.cpp:614-636) along the row
        int sig = 0, lvDC = 0;
        {
            const int add = 1 << (fshift2 - 1);
            const int4* qRow = (const int4*)sQ + row * (N / 4);
            const int qs = row & (N / 4 - 1);
            uint4* qOut = (uint4*)(qCoef + (size_t)(live ? tu : 0) * NN + row * N);
#pragma unroll
            for (int c = 0; c < N / 8; c++)
            {
                const int4 qa = qRow[(2 * c) ^ qs], qb = qRow[(2 * c + 1) ^ qs];
                const int qq[8] = { qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w };
                uint32_t lvp[4];
#pragma unroll
                for (int e = 0; e < 4; e++)
                {
                    int lv0, lv1;
                    const int c0 = (int)(int16_t)((v[8 * c + 2 * e] + add) >> fshift2), c1 = (int)(int16_t)((v[8 * c + 2 * e + 1] + add) >> fshift2);
                    sig += quant_one(c0, qq[2 * e], P, lv0) + quant_one(c1, qq[2 * e + 1], P, lv1);
                    if (c == 0 && e == 0) lvDC = lv0;        // level (row, 0): the DC level in the TU's first thread
                    lvp[e] = pack_sat_s16(lv0, lv1);
                    w[4 * c + e] = dequant_pair(lvp[e], P);
                }
                if constexpr (sizeof(T) == 2)
                    *WarpTile<N>::piece(sAk[0] + warp * (64 * N), lane, c) = make_uint4(lvp[0], lvp[1], lvp[2], lvp[3]);   // staged, see below
                else if (live && !(lab & 2)) qOut[c] = make_uint4(lvp[0], lvp[1], lvp[2], lvp[3]);
            }
            if constexpr (sizeof(T) == 2)
            {
                // the forward-1 planes are idle until the next group: the warp's level rows go through them and leave as whole 16-byte
                // pieces of consecutive rows (a TU's levels are contiguous in qCoef: row r of the warp is 2N bytes after row r - 1)
                typedef WarpTile<N> WT;
                uint8_t* tQ = sAk[0] + warp * (64 * N);         // sAk is 8 KB = 4 warps x 32 rows x 64 bytes (N = 32); half of it for N = 16
                __syncwarp();
                const int tu0 = g * NT + warp * (32 / N);       // first TU of this warp
#pragma unroll
                for (int it = 0; it < WT::NI; it++)
                {
                    const int R = it * WT::RPI + lane / WT::PPR, c = lane % WT::PPR;
                    const int tuR = tu0 + R / N;
                    if (tuR < n && !(lab & 2))
                        *(uint4*)(qCoef + (size_t)tuR * NN + (R % N) * N + 8 * c) = *WT::piece(tQ, R, c);
                }
                (void)qOut;
            }
        }
        const int ns = tu_sum32<N>(sig);
        // the DC coefficient is (0, 0): the TU's first thread holds its level in lvDC and its dequantised value in w[0]'s low half
        const int first = (tid & 31) & ~(N - 1);
        const int dq0 = (int)(int16_t)(__shfl_sync(0xffffffffu, w[0], first) & 0xffff);
        const int dcLevel = __shfl_sync(0xffffffffu, lvDC, first);        // outside the && : with two TUs per warp only one half may have ns == 1,
        const bool dcOnly = ns == 1 && dcLevel != 0;                      // and a shuffle that half the warp skips never completes
        // ---- inverse stage 1: tmp[j][n] = clip16((sum_k C[k][j] T[k][n] + 64) >> 7); the thread's row k of C is column k of C^T
        umma_write_mnmajor<N>(sAm[0], sAm[1], row, chunk0, w);
        umma_stage<N, COLS / 2>(3, 1, aM[0], aM[1], bI, tmem, barA, phase, v);
#pragma unroll
        for (int i = 0; i < H; i++) w[i] = pack_sat_s16((v[2 * i] + 64) >> 7, (v[2 * i + 1] + 64) >> 7);
        // ---- inverse stage 2: resi[a][b] = clip16((sum_k tmp[k][a] T[k][b] + add) >> shift2); this thread receives row a = row
        umma_write_mnmajor<N>(sAm[0], sAm[1], row, chunk0, w);
        umma_stage<N, COLS / 2>(4, 1, aM[0], aM[1], bI, tmem, barA, phase, v);
        // ---- reconstruction: cbf == 0 -> prediction; DC only -> flat residual (quant.cpp:588-598); else the inverse transform's row
        uint32_t d32 = 0;
        if constexpr (sizeof(T) == 2)
        {   // the thread's fenc / prediction rows again, from the warp's tiles (not held in registers across the four stages)
            typedef WarpTile<N> WT;
#pragma unroll
            for (int c = 0; c < WT::PPR; c++)
            {
                const uint4 a = *WT::piece(sTile[warp][0], lane, c), b = *WT::piece(sTile[warp][1], lane, c);
                f[4 * c] = a.x; f[4 * c + 1] = a.y; f[4 * c + 2] = a.z; f[4 * c + 3] = a.w;
                p[4 * c] = b.x; p[4 * c + 1] = b.y; p[4 * c + 2] = b.z; p[4 * c + 3] = b.w;
            }
        }
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
                if constexpr (sizeof(T) == 2) { p[i] = o0; p[i + 1] = o1; }
                else if (live && !(lab & 4)) store_pix4(out + 2 * i, o0, o1);
                d32 += sumsq4(make_uint2(psub16(f[i], o0), psub16(f[i + 1], o1)));
            }
            if constexpr (sizeof(T) == 2)
            {
                // reconstructed rows: through the prediction tile and out as coalesced 16-byte pieces when the TU's rows are 16-byte aligned
                // (TUs on the grid of an aligned plane); otherwise every thread stores its own row
                typedef WarpTile<N> WT;
                const bool al = !(((uintptr_t)out | (uintptr_t)(sr * 2)) & 15);
                if (__all_sync(0xffffffffu, al || !live))
                {
                    uint8_t* tP = sTile[warp][1];
#pragma unroll
                    for (int c = 0; c < WT::PPR; c++) *WT::piece(tP, lane, c) = make_uint4(p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]);
                    __syncwarp();
                    const int myOffR = live ? offR[tu] : 0;
#pragma unroll
                    for (int it = 0; it < WT::NI; it++)
                    {
                        const int R = it * WT::RPI + lane / WT::PPR, c = lane % WT::PPR;
                        const int oR = __shfl_sync(0xffffffffu, myOffR, R);
                        const bool rl = __shfl_sync(0xffffffffu, (int)live, R) != 0;
                        if (rl && !(lab & 4)) *(uint4*)(recon + oR + (intptr_t)(R % N) * sr + 8 * c) = *WT::piece(tP, R, c);
                    }
                }
                else if (live && !(lab & 4))
                {
#pragma unroll
                    for (int i = 0; i < H; i += 2) store_pix4(out + 2 * i, p[i], p[i + 1]);
                }
            }
        }
        __syncwarp();
        const unsigned long long zs = tu_sum64<N>(z32), ds = tu_sum64<N>(d32);
        if (live && row == 0 && !(lab & 8))
        {
            numSig[tu] = (uint32_t)ns;
            if (sseZero) sseZero[tu] = zs;
            sseRecon[tu] = ds;
        }
    }
    UMMA_TRACE(99);
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
    if ((N != 32 && N != 16) || ((sf | sp) & 3) || ((uintptr_t)qCoef & 15) || ((uintptr_t)quantCoeff & 15)) return false;
    const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
    const int lg = N == 32 ? 5 : 4, d8 = ctx->depth - 8;
    QuantP P; P.qBits = qBits; P.qAdd = qAdd; P.dqScale = dqScale; P.dqAdd = 1 << (dqShift - 1); P.dqShift = dqShift;
    const int ngroups = (n + 128 / N - 1) / (128 / N);
    // resident CTAs per SM the launch bounds ask for: 4 at N = 32 (115 registers; capping at 96 for 5 CTAs spills and is 35 % slower), 5 at N = 16
    constexpr int RES32 = 4, RES16 = 5;
    const int RES = N == 32 ? RES32 : RES16;
    int grid = sms * RES;
    int pad = 0, lab = 0;                                     // lab: X265B200_UMMA_LAB="ctas_per_sm,dynamic_smem_bytes,skip_mask"
    if (const char* e = getenv("X265B200_UMMA_LAB")) { int c = RES; sscanf(e, "%d,%d,%d", &c, &pad, &lab); grid = sms * c; }
#ifdef B200_UMMA_TRACE
    if (const char* e = getenv("X265B200_UMMA_TRACE")) { int* tp = (int*)strtoull(e, nullptr, 0); cudaMemcpyToSymbol(g_ummaTrace, &tp, sizeof(tp)); }
#endif
    if (grid > ngroups) grid = ngroups;
#define UM(T, N_) cudaFuncSetAttribute(tu_umma_kernel<T, N_, N_ == 32 ? RES32 : RES16, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 28672 + 4 * 2 * 64 * N_ + pad); \
                  tu_umma_kernel<T, N_, N_ == 32 ? RES32 : RES16, 64><<<grid, 128, 28672 + 4 * 2 * 64 * N_ + pad, st>>>((const T*)fenc, sf, (const T*)pred, sp, offF, offP, n, quantCoeff, P, lg - 1 + d8, lg + 6, 12 - d8, \
                      ctx->depth, qCoef, numSig, (T*)recon, sr, offR, (unsigned long long*)sseZero, (unsigned long long*)sseRecon, lab)
    if (ctx->pixbytes == 1) { if (N == 32) { UM(uint8_t, 32); } else { UM(uint8_t, 16); } }
    else { if (N == 32) { UM(uint16_t, 32); } else { UM(uint16_t, 16); } }
#undef UM
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError() == cudaSuccess;
}
