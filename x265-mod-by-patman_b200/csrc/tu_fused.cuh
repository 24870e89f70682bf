// tu_fused.cuh -- the inter-luma TU reconstruction chain as TWO tensor-core kernels, no scratch buffer (included at the
// end of transform_mma.cu, inside namespace b200, because it reuses that file's IMMA fragments and operand tables):
//
//   tu_fwd*_kernel : resi = fenc - pred (formed in the DCT's operand layout straight from the planes), sse(fenc, pred)
//                    -> forward DCT (both IMMA stages, as dct*_imma_kernel)
//                    -> quant (dct.cpp:666-688 without deltaU) on the accumulators -> qCoef, numSig
//   tu_inv*_kernel : qCoef -> dequant_normal (dct.cpp:614-636) while building the inverse transform's operand
//                    -> inverse DCT (as idct*_imma_kernel; skipped for cbf == 0, replaced by the DC fill of
//                    quant.cpp:588-598 for DC-only TUs) -> recon = clip(pred + resi') (pixel.cpp:821-831)
//                    -> sse(fenc, recon) (pixel.cpp:167-186)
//
// The MMA accumulator layout gives a lane two adjacent columns of a row; lanes t and t ^ 1 swap one pair each
// (pair_to_quad) so that every lane owns FOUR adjacent columns: 8-byte qCoef / pixel accesses and 16-byte table loads
// instead of twice as many half-sized ones.  Reconstruction and clipping run on packed sample pairs
// (VIADDMNMX.S16x2.RELU).  HBM traffic: fenc + pred in, qCoef out, then qCoef + fenc + pred in, recon out = 7b + 4 bytes per
// sample against the chain's algorithmic 3b + 2.  Reference call sequence: encoder/search.cpp:5536-5575.

template<typename T> __device__ __forceinline__ uint2 res_quad(const T* f, const T* p)
{
    // four horizontally adjacent fenc - pred as lane-exact packed int16 pairs (any sample alignment)
    uint32_t wf[2], wp[2];
    load_row_quads<1>(f, wf);
    load_row_quads<1>(p, wp);
    return make_uint2(psub16(wf[0], wp[0]), psub16(wf[1], wp[1]));
}

// ROWS quads, `rowStep` rows apart (strip loader: one misalignment class, all loads issued first)
template<typename T, int ROWS>
__device__ __forceinline__ void res_quads(const T* f, intptr_t fStep, const T* p, intptr_t pStep, uint2 (&x)[ROWS])
{
    uint32_t wf[ROWS][3], wp[ROWS][3];
    load_rows_quads<1, ROWS>(f, fStep, wf);
    load_rows_quads<1, ROWS>(p, pStep, wp);
#pragma unroll
    for (int r = 0; r < ROWS; r++) x[r] = make_uint2(psub16(wf[r][0], wp[r][0]), psub16(wf[r][1], wp[r][1]));
}

// sum of squares of the four lane-exact int16 in q
__device__ __forceinline__ uint32_t sumsq4(uint2 q)
{
    int a = (int16_t)(q.x & 0xffff), b = (int)q.x >> 16, c = (int16_t)(q.y & 0xffff), d = (int)q.y >> 16;
    return (uint32_t)(a * a) + (uint32_t)(b * b) + (uint32_t)(c * c) + (uint32_t)(d * d);
}

// lanes t and t ^ 1: each holds column pair `a` of slot 0 and column pair `b` of slot 1 (columns 2t, 2t+1);
// afterwards the even lane holds columns 4(t>>1)..+3 of slot 0 and the odd lane those of slot 1
__device__ __forceinline__ uint2 pair_to_quad(uint32_t a, uint32_t b, int t)
{
    uint32_t recv = __shfl_xor_sync(0xffffffffu, (t & 1) ? a : b, 1);
    return (t & 1) ? make_uint2(recv, b) : make_uint2(a, recv);
}
__device__ __forceinline__ uint32_t pack2(int v0, int v1) { return __byte_perm((uint32_t)v0, (uint32_t)v1, 0x5410); }

struct QuantP { int qBits, qAdd, dqScale, dqAdd, dqShift; };

// quant of one coefficient (the low 16 bits of c are the DCT output); returns level != 0.  The clip3(-32768, 32767, .)
// of dct.cpp:683 happens when two levels are packed (pack_sat_s16 = I2IP.S16.S32.SAT).
__device__ __forceinline__ int quant_one(int c, int q, const QuantP& P, int& level)
{
    int sign = c < 0 ? -1 : 1;
    int tmplevel = (int)((unsigned)abs(c) * (unsigned)q);                       // int32 wrap, dct.cpp:678
    int lv = (int)((unsigned)tmplevel + (unsigned)P.qAdd) >> P.qBits;
    level = (int)((unsigned)lv * (unsigned)sign);
    return lv != 0;
}
// four horizontally adjacent coefficients at TU position pos (a multiple of 4): quantise, store the levels
__device__ __forceinline__ int quant_quad_store(uint2 c, const int32_t* __restrict__ quantCoeff, int pos, const QuantP& P, int16_t* __restrict__ qTu)
{
    int4 q = __ldg((const int4*)(quantCoeff + pos));
    int l0, l1, l2, l3;
    int nz = quant_one((int)(int16_t)(c.x & 0xffff), q.x, P, l0) + quant_one((int)c.x >> 16, q.y, P, l1)
           + quant_one((int)(int16_t)(c.y & 0xffff), q.z, P, l2) + quant_one((int)c.y >> 16, q.w, P, l3);
    *(uint2*)(qTu + pos) = make_uint2(pack_sat_s16(l0, l1), pack_sat_s16(l2, l3));
    return nz;
}
__device__ __forceinline__ int dequant_raw(int lv, const QuantP& P)
{
    return (int)((unsigned)lv * (unsigned)P.dqScale + (unsigned)P.dqAdd) >> P.dqShift;
}
__device__ __forceinline__ int dequant_one(int lv, const QuantP& P) { return min(32767, max(-32768, dequant_raw(lv, P))); }
__device__ __forceinline__ uint32_t dequant_pair(uint32_t w, const QuantP& P)
{
    return pack_sat_s16(dequant_raw((int)(int16_t)(w & 0xffff), P), dequant_raw((int)w >> 16, P));
}

// DC-only reconstruction value (quant.cpp:588-598)
__device__ __forceinline__ int dc_fill_value(int dq0, int depth)
{
    const int shift_2nd = 12 - (depth - 8) - 3, add_2nd = 1 << (shift_2nd - 1);
    return (int)(int16_t)((((dq0 + 1) >> 1) * 8 + add_2nd) >> shift_2nd);
}

__device__ __forceinline__ void store_pix4(uint16_t* d, uint32_t p01, uint32_t p23)
{
    uintptr_t a = (uintptr_t)d;
    if ((a & 7) == 0) *(uint2*)d = make_uint2(p01, p23);
    else if ((a & 3) == 0) { ((uint32_t*)d)[0] = p01; ((uint32_t*)d)[1] = p23; }
    else { d[0] = (uint16_t)(p01 & 0xffff); d[1] = (uint16_t)(p01 >> 16); d[2] = (uint16_t)(p23 & 0xffff); d[3] = (uint16_t)(p23 >> 16); }
}
__device__ __forceinline__ void store_pix4(uint8_t* d, uint32_t p01, uint32_t p23)
{
    uint32_t b = __byte_perm(p01, p23, 0x6420);
    if (((uintptr_t)d & 3) == 0) *(uint32_t*)d = b;
    else { d[0] = (uint8_t)(b & 0xff); d[1] = (uint8_t)((b >> 8) & 0xff); d[2] = (uint8_t)((b >> 16) & 0xff); d[3] = (uint8_t)(b >> 24); }
}

// four adjacent samples: recon = ns ? clip(pred + r) : pred, stored; d += sum (fenc - recon)^2.  r = packed int16 pairs.
// clip(p + r, 0, max) == clip(p + clamp(r, -max, max), 0, max) for p in [0, max], which keeps the packed add inside int16.
template<typename T>
__device__ __forceinline__ void recon_quad(const T* __restrict__ pf, const T* __restrict__ pp, T* __restrict__ pr, int mode, uint2 r,
                                           uint32_t mx, uint32_t negmx, uint32_t& d)
{
    uint32_t wf[2], wp[2];
    load_row_quads<1>(pf, wf);
    load_row_quads<1>(pp, wp);
    uint32_t o0 = wp[0], o1 = wp[1];
    if (mode)
    {
        o0 = __viaddmin_s16x2_relu(wp[0], __vmins2(__vmaxs2(r.x, negmx), mx), mx);
        o1 = __viaddmin_s16x2_relu(wp[1], __vmins2(__vmaxs2(r.y, negmx), mx), mx);
    }
    store_pix4(pr, o0, o1);
    d += sumsq4(make_uint2(psub16(wf[0], o0), psub16(wf[1], o1)));
}

__device__ __forceinline__ unsigned long long warp_sum64(uint32_t v)
{
    unsigned long long s = v;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    return s;
}

// ------------------------------------------------------------------------------------------------ N = 32
template<typename T, int MINB>
__global__ void __launch_bounds__(128, MINB)
tu_fwd32_kernel(const T* __restrict__ fenc, intptr_t sf, const T* __restrict__ pred, intptr_t sp,
                const int32_t* __restrict__ offF, const int32_t* __restrict__ offP, int n,
                const int32_t* __restrict__ quantCoeff, QuantP P, int shift1, int shift2,
                int16_t* __restrict__ qCoef, uint32_t* __restrict__ numSig, unsigned long long* __restrict__ sseZero)
{
    int lane = threadIdx.x & 31;
    int warp = blockIdx.x * 4 + (threadIdx.x >> 5);
    int nwarps = gridDim.x * 4;
    if (warp >= n) return;
    int g = lane >> 2, t = lane & 3;
    uint32_t a1[2][4], a2[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int r = 0; r < 4; r++) { a1[mt][r] = c_A32[0][mt][r][lane]; a2[mt][r] = c_A32[1][mt][r][lane]; }
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);
    const intptr_t lf = (intptr_t)g * sf + 4 * t, lp = (intptr_t)g * sp + 4 * t;

    uint2 x[4][2];
    {
        const T* f = fenc + offF[warp] + lf;
        const T* p = pred + offP[warp] + lp;
        uint2 xa[4], xb[4];
        res_quads<T, 4>(f, 8 * sf, p, 8 * sp, xa);
        res_quads<T, 4>(f + 16, 8 * sf, p + 16, 8 * sp, xb);
#pragma unroll
        for (int jt = 0; jt < 4; jt++) { x[jt][0] = xa[jt]; x[jt][1] = xb[jt]; }
    }
    for (int tu = warp; tu < n; tu += nwarps)
    {
        uint32_t blo[4][2], bhi[4][2];
        uint32_t z = 0;
#pragma unroll
        for (int jt = 0; jt < 4; jt++)
        {
            split4(x[jt][0], blo[jt][0], bhi[jt][0]);
            split4(x[jt][1], blo[jt][1], bhi[jt][1]);
            z += sumsq4(x[jt][0]) + sumsq4(x[jt][1]);
        }
        int nxt = tu + nwarps;
        if (nxt < n)
        {
            const T* f = fenc + offF[nxt] + lf;
            const T* p = pred + offP[nxt] + lp;
            uint2 xa[4], xb[4];
            res_quads<T, 4>(f, 8 * sf, p, 8 * sp, xa);
            res_quads<T, 4>(f + 16, 8 * sf, p + 16, 8 * sp, xb);
#pragma unroll
            for (int jt = 0; jt < 4; jt++) { x[jt][0] = xa[jt]; x[jt][1] = xb[jt]; }
        }
        uint32_t b2lo[4][2], b2hi[4][2];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
        {
            int v[4][4];
#pragma unroll
            for (int jt = 0; jt < 4; jt++)
            {
                int chi[4] = { 0, 0, 0, 0 }, clo[4] = { add1, add1, add1, add1 };
                imma32_ss(chi, a1[mt], bhi[jt][0], bhi[jt][1]);
                imma32_su(clo, a1[mt], blo[jt][0], blo[jt][1]);
#pragma unroll
                for (int r = 0; r < 4; r++) v[jt][r] = recombine(chi[r], clo[r], shift1);
            }
            pack4(v[0][0], v[0][1], v[1][0], v[1][1], b2lo[2 * mt][0], b2hi[2 * mt][0]);
            pack4(v[2][0], v[2][1], v[3][0], v[3][1], b2lo[2 * mt][1], b2hi[2 * mt][1]);
            pack4(v[0][2], v[0][3], v[1][2], v[1][3], b2lo[2 * mt + 1][0], b2hi[2 * mt + 1][0]);
            pack4(v[2][2], v[2][3], v[3][2], v[3][3], b2lo[2 * mt + 1][1], b2hi[2 * mt + 1][1]);
        }
        int16_t* qTu = qCoef + (size_t)tu * 1024;
        int sig = 0;
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
        {
#pragma unroll
            for (int np = 0; np < 2; np++)
            {
                uint32_t pa[2], pb[2];                                   // [row g | row g + 8] of n-tiles 2np / 2np + 1
#pragma unroll
                for (int k = 0; k < 2; k++)
                {
                    int chi[4] = { 0, 0, 0, 0 }, clo[4] = { add2, add2, add2, add2 };
                    imma32_ss(chi, a2[mt], b2hi[2 * np + k][0], b2hi[2 * np + k][1]);
                    imma32_su(clo, a2[mt], b2lo[2 * np + k][0], b2lo[2 * np + k][1]);
                    uint32_t top = pack2(recombine(chi[0], clo[0], shift2), recombine(chi[1], clo[1], shift2));
                    uint32_t bot = pack2(recombine(chi[2], clo[2], shift2), recombine(chi[3], clo[3], shift2));
                    if (k == 0) { pa[0] = top; pa[1] = bot; } else { pb[0] = top; pb[1] = bot; }
                }
                int pos = (mt * 16 + g) * 32 + (2 * np + (t & 1)) * 8 + 4 * (t >> 1);
                sig += quant_quad_store(pair_to_quad(pa[0], pb[0], t), quantCoeff, pos, P, qTu);
                sig += quant_quad_store(pair_to_quad(pa[1], pb[1], t), quantCoeff, pos + 256, P, qTu);
            }
        }
        sig = __reduce_add_sync(0xffffffffu, sig);
        unsigned long long zs = warp_sum64(z);
        if (lane == 0) { numSig[tu] = (uint32_t)sig; if (sseZero) sseZero[tu] = zs; }
    }
}

template<typename T, int MINB>
__global__ void __launch_bounds__(128, MINB)
tu_inv32_kernel(const int16_t* __restrict__ qCoef, const uint32_t* __restrict__ numSig, int n, QuantP P,
                const T* __restrict__ fenc, intptr_t sf, const T* __restrict__ pred, intptr_t sp,
                const int32_t* __restrict__ offF, const int32_t* __restrict__ offP,
                T* __restrict__ recon, intptr_t sr, const int32_t* __restrict__ offR,
                unsigned long long* __restrict__ sseRecon, int shift1, int shift2, int depth)
{
    int lane = threadIdx.x & 31;
    int warp = blockIdx.x * 4 + (threadIdx.x >> 5);
    int nwarps = gridDim.x * 4;
    if (warp >= n) return;
    int g = lane >> 2, t = lane & 3;
    uint32_t a1[2][4], b2[4][2];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int r = 0; r < 4; r++) a1[mt][r] = c_IA32[mt][r][lane];
#pragma unroll
    for (int it = 0; it < 4; it++) { b2[it][0] = c_IB32[it][0][lane]; b2[it][1] = c_IB32[it][1][lane]; }
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);
    const uint32_t mx = (uint32_t)((1 << depth) - 1) * 0x10001u, negmx = (uint32_t)(-((1 << depth) - 1) & 0xffff) * 0x10001u;

    uint32_t x[2][2][4];
    {
        const int16_t* q = qCoef + (size_t)warp * 1024 + t * 32 + 2 * g;
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int half = 0; half < 2; half++)
#pragma unroll
                for (int e = 0; e < 4; e++) x[p][half][e] = __ldg((const uint32_t*)(q + ((half * 4 + e) * 4) * 32 + p * 16));
    }
    for (int tu = warp; tu < n; tu += nwarps)
    {
        uint32_t ns = numSig[tu];
        int q0 = qCoef[(size_t)tu * 1024];
        bool dcOnly = ns == 1 && q0 != 0;
        int dcv = dcOnly ? dc_fill_value(dequant_one(q0, P), depth) : 0;
        bool full = ns != 0 && !dcOnly;
        uint32_t blo[4][2], bhi[4][2];
        if (full)
        {
#pragma unroll
            for (int p = 0; p < 2; p++)
#pragma unroll
                for (int half = 0; half < 2; half++)
                {
                    uint32_t y0 = dequant_pair(x[p][half][0], P), y1 = dequant_pair(x[p][half][1], P);
                    uint32_t y2 = dequant_pair(x[p][half][2], P), y3 = dequant_pair(x[p][half][3], P);
                    uint32_t e01 = __byte_perm(y0, y1, 0x5140), e23 = __byte_perm(y2, y3, 0x5140);
                    uint32_t o01 = __byte_perm(y0, y1, 0x7362), o23 = __byte_perm(y2, y3, 0x7362);
                    blo[2 * p][half] = __byte_perm(e01, e23, 0x5410); bhi[2 * p][half] = __byte_perm(e01, e23, 0x7632);
                    blo[2 * p + 1][half] = __byte_perm(o01, o23, 0x5410); bhi[2 * p + 1][half] = __byte_perm(o01, o23, 0x7632);
                }
        }
        int nxt = tu + nwarps;
        if (nxt < n)
        {
            const int16_t* q = qCoef + (size_t)nxt * 1024 + t * 32 + 2 * g;
#pragma unroll
            for (int p = 0; p < 2; p++)
#pragma unroll
                for (int half = 0; half < 2; half++)
#pragma unroll
                    for (int e = 0; e < 4; e++) x[p][half][e] = __ldg((const uint32_t*)(q + ((half * 4 + e) * 4) * 32 + p * 16));
        }
        const T* pf = fenc + offF[tu] + (intptr_t)g * sf + 4 * (t >> 1);
        const T* pp = pred + offP[tu] + (intptr_t)g * sp + 4 * (t >> 1);
        T* pr = recon + offR[tu] + (intptr_t)g * sr + 4 * (t >> 1);
        const uint32_t dcp = pack2(dcv, dcv);
        uint32_t d = 0;
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
        {
            uint32_t alo[4], ahi[4];
            if (full)
            {
                int v[4][4];
#pragma unroll
                for (int nt = 0; nt < 4; nt++)
                {
                    int chi[4] = { 0, 0, 0, 0 }, clo[4] = { add1, add1, add1, add1 };
                    imma32_ss(chi, a1[mt], bhi[nt][0], bhi[nt][1]);
                    imma32_su(clo, a1[mt], blo[nt][0], blo[nt][1]);
#pragma unroll
                    for (int r = 0; r < 4; r++) v[nt][r] = recombine(chi[r], clo[r], shift1);
                }
                pack4_sat(v[0][0], v[0][1], v[1][0], v[1][1], alo[0], ahi[0]);
                pack4_sat(v[0][2], v[0][3], v[1][2], v[1][3], alo[1], ahi[1]);
                pack4_sat(v[2][0], v[2][1], v[3][0], v[3][1], alo[2], ahi[2]);
                pack4_sat(v[2][2], v[2][3], v[3][2], v[3][3], alo[3], ahi[3]);
            }
#pragma unroll
            for (int ip = 0; ip < 2; ip++)
            {
                uint2 qa = make_uint2(dcp, dcp), qb = make_uint2(dcp, dcp);       // rows g / g + 8, four adjacent columns
                if (full)
                {
                    uint32_t pa[2], pb[2];
#pragma unroll
                    for (int k = 0; k < 2; k++)
                    {
                        int dhi[4] = { 0, 0, 0, 0 }, dlo[4] = { add2, add2, add2, add2 };
                        imma32_ss(dhi, ahi, b2[2 * ip + k][0], b2[2 * ip + k][1]);
                        imma32_us(dlo, alo, b2[2 * ip + k][0], b2[2 * ip + k][1]);
                        uint32_t top = recombine_sat2(dhi[0], dlo[0], dhi[1], dlo[1], shift2);
                        uint32_t bot = recombine_sat2(dhi[2], dlo[2], dhi[3], dlo[3], shift2);
                        if (k == 0) { pa[0] = top; pa[1] = bot; } else { pb[0] = top; pb[1] = bot; }
                    }
                    qa = pair_to_quad(pa[0], pb[0], t);
                    qb = pair_to_quad(pa[1], pb[1], t);
                }
                int col = (2 * ip + (t & 1)) * 8;
                intptr_t ra = mt * 16, rb = mt * 16 + 8;
                recon_quad(pf + ra * sf + col, pp + ra * sp + col, pr + ra * sr + col, (int)ns, qa, mx, negmx, d);
                recon_quad(pf + rb * sf + col, pp + rb * sp + col, pr + rb * sr + col, (int)ns, qb, mx, negmx, d);
            }
        }
        unsigned long long ds = warp_sum64(d);
        if (lane == 0) sseRecon[tu] = ds;
    }
}

// ------------------------------------------------------------------------------------------------ N = 16
template<typename T, int TPW>
__global__ void __launch_bounds__(128)
tu_fwd16_kernel(const T* __restrict__ fenc, intptr_t sf, const T* __restrict__ pred, intptr_t sp,
                const int32_t* __restrict__ offF, const int32_t* __restrict__ offP, int n,
                const int32_t* __restrict__ quantCoeff, QuantP P, int shift1, int shift2,
                int16_t* __restrict__ qCoef, uint32_t* __restrict__ numSig, unsigned long long* __restrict__ sseZero)
{
    int lane = threadIdx.x & 31;
    int warp = blockIdx.x * 4 + (threadIdx.x >> 5);
    int nwarps = gridDim.x * 4;
    int ngroups = (n + TPW - 1) / TPW;
    if (warp >= ngroups) return;
    int g = lane >> 2, t = lane & 3;
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);
    uint32_t a1[2] = { c_A16[0][0][lane], c_A16[0][1][lane] };
    uint32_t a2[2] = { c_A16[1][0][lane], c_A16[1][1][lane] };
    const intptr_t lf = (intptr_t)g * sf + 4 * t, lp = (intptr_t)g * sp + 4 * t;
    uint2 x[TPW][2];
#pragma unroll
    for (int u = 0; u < TPW; u++)
    {
        int tu = min(warp * TPW + u, n - 1);
        const T* f = fenc + offF[tu] + lf;
        const T* p = pred + offP[tu] + lp;
        res_quads<T, 2>(f, 8 * sf, p, 8 * sp, x[u]);
    }
    for (int grp = warp; grp < ngroups; grp += nwarps)
    {
        uint32_t blo[TPW][2], bhi[TPW][2], z[TPW];
#pragma unroll
        for (int u = 0; u < TPW; u++)
        {
            split4(x[u][0], blo[u][0], bhi[u][0]);
            split4(x[u][1], blo[u][1], bhi[u][1]);
            z[u] = sumsq4(x[u][0]) + sumsq4(x[u][1]);
        }
        int nxt = grp + nwarps;
        if (nxt < ngroups)
        {
#pragma unroll
            for (int u = 0; u < TPW; u++)
            {
                int tu = min(nxt * TPW + u, n - 1);
                const T* f = fenc + offF[tu] + lf;
                const T* p = pred + offP[tu] + lp;
                res_quads<T, 2>(f, 8 * sf, p, 8 * sp, x[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < TPW; u++)
        {
            int tu = grp * TPW + u;
            if (tu >= n) break;
            int v[2][4];
#pragma unroll
            for (int jt = 0; jt < 2; jt++)
            {
                int chi[4] = { 0, 0, 0, 0 }, clo[4] = { add1, add1, add1, add1 };
                imma16_ss(chi, a1, bhi[u][jt]);
                imma16_su(clo, a1, blo[u][jt]);
#pragma unroll
                for (int r = 0; r < 4; r++) v[jt][r] = recombine(chi[r], clo[r], shift1);
            }
            uint32_t b2lo[2], b2hi[2];
            pack4(v[0][0], v[0][1], v[1][0], v[1][1], b2lo[0], b2hi[0]);
            pack4(v[0][2], v[0][3], v[1][2], v[1][3], b2lo[1], b2hi[1]);
            int16_t* qTu = qCoef + (size_t)tu * 256;
            uint32_t pa[2], pb[2];
#pragma unroll
            for (int k = 0; k < 2; k++)
            {
                int chi[4] = { 0, 0, 0, 0 }, clo[4] = { add2, add2, add2, add2 };
                imma16_ss(chi, a2, b2hi[k]);
                imma16_su(clo, a2, b2lo[k]);
                uint32_t top = pack2(recombine(chi[0], clo[0], shift2), recombine(chi[1], clo[1], shift2));
                uint32_t bot = pack2(recombine(chi[2], clo[2], shift2), recombine(chi[3], clo[3], shift2));
                if (k == 0) { pa[0] = top; pa[1] = bot; } else { pb[0] = top; pb[1] = bot; }
            }
            int pos = g * 16 + (t & 1) * 8 + 4 * (t >> 1);
            int sig = quant_quad_store(pair_to_quad(pa[0], pb[0], t), quantCoeff, pos, P, qTu)
                    + quant_quad_store(pair_to_quad(pa[1], pb[1], t), quantCoeff, pos + 128, P, qTu);
            sig = __reduce_add_sync(0xffffffffu, sig);
            unsigned long long zs = warp_sum64(z[u]);
            if (lane == 0) { numSig[tu] = (uint32_t)sig; if (sseZero) sseZero[tu] = zs; }
        }
    }
}

template<typename T, int TPW>
__global__ void __launch_bounds__(128)
tu_inv16_kernel(const int16_t* __restrict__ qCoef, const uint32_t* __restrict__ numSig, int n, QuantP P,
                const T* __restrict__ fenc, intptr_t sf, const T* __restrict__ pred, intptr_t sp,
                const int32_t* __restrict__ offF, const int32_t* __restrict__ offP,
                T* __restrict__ recon, intptr_t sr, const int32_t* __restrict__ offR,
                unsigned long long* __restrict__ sseRecon, int shift1, int shift2, int depth)
{
    int lane = threadIdx.x & 31;
    int warp = blockIdx.x * 4 + (threadIdx.x >> 5);
    int nwarps = gridDim.x * 4;
    int ngroups = (n + TPW - 1) / TPW;
    if (warp >= ngroups) return;
    int g = lane >> 2, t = lane & 3;
    uint32_t a1[2] = { c_IA16[0][lane], c_IA16[1][lane] };
    uint32_t b2[2] = { c_IB16[0][lane], c_IB16[1][lane] };
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);
    const uint32_t mx = (uint32_t)((1 << depth) - 1) * 0x10001u, negmx = (uint32_t)(-((1 << depth) - 1) & 0xffff) * 0x10001u;
    uint32_t x[TPW][4];
#pragma unroll
    for (int u = 0; u < TPW; u++)
    {
        const int16_t* q = qCoef + (size_t)min(warp * TPW + u, n - 1) * 256 + t * 16 + 2 * g;
#pragma unroll
        for (int e = 0; e < 4; e++) x[u][e] = __ldg((const uint32_t*)(q + e * 4 * 16));
    }
    for (int grp = warp; grp < ngroups; grp += nwarps)
    {
        uint32_t blo[TPW][2], bhi[TPW][2];
#pragma unroll
        for (int u = 0; u < TPW; u++)
        {
            uint32_t y0 = dequant_pair(x[u][0], P), y1 = dequant_pair(x[u][1], P), y2 = dequant_pair(x[u][2], P), y3 = dequant_pair(x[u][3], P);
            uint32_t e01 = __byte_perm(y0, y1, 0x5140), e23 = __byte_perm(y2, y3, 0x5140);
            uint32_t o01 = __byte_perm(y0, y1, 0x7362), o23 = __byte_perm(y2, y3, 0x7362);
            blo[u][0] = __byte_perm(e01, e23, 0x5410); bhi[u][0] = __byte_perm(e01, e23, 0x7632);
            blo[u][1] = __byte_perm(o01, o23, 0x5410); bhi[u][1] = __byte_perm(o01, o23, 0x7632);
        }
        int nxt = grp + nwarps;
        if (nxt < ngroups)
        {
#pragma unroll
            for (int u = 0; u < TPW; u++)
            {
                const int16_t* q = qCoef + (size_t)min(nxt * TPW + u, n - 1) * 256 + t * 16 + 2 * g;
#pragma unroll
                for (int e = 0; e < 4; e++) x[u][e] = __ldg((const uint32_t*)(q + e * 4 * 16));
            }
        }
#pragma unroll
        for (int u = 0; u < TPW; u++)
        {
            int tu = grp * TPW + u;
            if (tu >= n) break;
            uint32_t ns = numSig[tu];
            int q0 = qCoef[(size_t)tu * 256];
            bool dcOnly = ns == 1 && q0 != 0;
            int dcv = dcOnly ? dc_fill_value(dequant_one(q0, P), depth) : 0;
            bool full = ns != 0 && !dcOnly;
            const uint32_t dcp = pack2(dcv, dcv);
            uint2 qa = make_uint2(dcp, dcp), qb = make_uint2(dcp, dcp);
            if (full)
            {
                int v[2][4];
#pragma unroll
                for (int nt = 0; nt < 2; nt++)
                {
                    int chi[4] = { 0, 0, 0, 0 }, clo[4] = { add1, add1, add1, add1 };
                    imma16_ss(chi, a1, bhi[u][nt]);
                    imma16_su(clo, a1, blo[u][nt]);
#pragma unroll
                    for (int r = 0; r < 4; r++) v[nt][r] = recombine(chi[r], clo[r], shift1);
                }
                uint32_t alo[2], ahi[2];
                pack4_sat(v[0][0], v[0][1], v[1][0], v[1][1], alo[0], ahi[0]);
                pack4_sat(v[0][2], v[0][3], v[1][2], v[1][3], alo[1], ahi[1]);
                uint32_t pa[2], pb[2];
#pragma unroll
                for (int k = 0; k < 2; k++)
                {
                    int dhi[4] = { 0, 0, 0, 0 }, dlo[4] = { add2, add2, add2, add2 };
                    imma16_ss(dhi, ahi, b2[k]);
                    imma16_us(dlo, alo, b2[k]);
                    uint32_t top = recombine_sat2(dhi[0], dlo[0], dhi[1], dlo[1], shift2);
                    uint32_t bot = recombine_sat2(dhi[2], dlo[2], dhi[3], dlo[3], shift2);
                    if (k == 0) { pa[0] = top; pa[1] = bot; } else { pb[0] = top; pb[1] = bot; }
                }
                qa = pair_to_quad(pa[0], pb[0], t);
                qb = pair_to_quad(pa[1], pb[1], t);
            }
            int col = (t & 1) * 8 + 4 * (t >> 1);
            const T* pf = fenc + offF[tu] + (intptr_t)g * sf + col;
            const T* pp = pred + offP[tu] + (intptr_t)g * sp + col;
            T* pr = recon + offR[tu] + (intptr_t)g * sr + col;
            uint32_t d = 0;
            recon_quad(pf, pp, pr, (int)ns, qa, mx, negmx, d);
            recon_quad(pf + 8 * sf, pp + 8 * sp, pr + 8 * sr, (int)ns, qb, mx, negmx, d);
            unsigned long long ds = warp_sum64(d);
            if (lane == 0) sseRecon[tu] = ds;
        }
    }
}

// ------------------------------------------------------------------------------------------------ N = 8, N = 4
// Lanes of one MMA group belong to different TUs: per-TU sums run over the lane bits that stay inside the TU.
// mask bit i set = lane bit i varies inside the TU.
template<int MASK> __device__ __forceinline__ uint32_t lane_bits_sum(uint32_t v)
{
#pragma unroll
    for (int m = 1; m < 32; m <<= 1)
        if (MASK & m) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

template<typename T, int SMALL, int UN>
__global__ void __launch_bounds__(128)
tu_fwd_small_kernel(const T* __restrict__ fenc, intptr_t sf, const T* __restrict__ pred, intptr_t sp,
                    const int32_t* __restrict__ offF, const int32_t* __restrict__ offP, int n,
                    const int32_t* __restrict__ quantCoeff, QuantP P, int shift1, int shift2,
                    int16_t* __restrict__ qCoef, uint32_t* __restrict__ numSig, unsigned long long* __restrict__ sseZero, int kind)
{
    constexpr int TPG = SMALL == 8 ? 2 : 8;
    constexpr int NN = SMALL * SMALL;
    // kind 1 (SMALL == 4 only): DST-VII instead of the DCT (intra luma 4x4 TUs).  lane = 4g + t.  loads: N = 8 TU t>>1 (lane bits 0,2,3,4 vary inside the TU), N = 4 TU (g>>2)*4 + t (bits 2,3 vary).
    // stores after pair_to_quad: N = 8 TU t&1 (bits 1,2,3,4 vary), N = 4 TU 4(t&1) + (t>>1) + 2(g>>2) (bits 2,3 vary).
    constexpr int LD_MASK = SMALL == 8 ? 0x1d : 0x0c;
    constexpr int ST_MASK = SMALL == 8 ? 0x1e : 0x0c;
    int lane = threadIdx.x & 31;
    int warp = blockIdx.x * 4 + (threadIdx.x >> 5);
    int nwarps = gridDim.x * 4;
    int ngroups = (n + TPG * UN - 1) / (TPG * UN);
    if (warp >= ngroups) return;
    int g = lane >> 2, t = lane & 3;
    const uint32_t (*A)[2][32] = SMALL == 8 ? c_A8 : c_A4[kind];
    uint32_t a1[2] = { A[0][0][lane], A[0][1][lane] };
    uint32_t a2[2] = { A[1][0][lane], A[1][1][lane] };
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);
    const int ld_tu = SMALL == 8 ? (t >> 1) : ((g >> 2) * 4 + t);
    const int ld_row = SMALL == 8 ? g : (g & 3), ld_col = SMALL == 8 ? 4 * (t & 1) : 0;
    const int st_tu = SMALL == 8 ? (t & 1) : 4 * (t & 1) + (t >> 1) + 2 * (g >> 2);
    const int st_pos = SMALL == 8 ? g * 8 + 4 * (t >> 1) : (g & 3) * 4;

    uint2 x[UN];
#pragma unroll
    for (int u = 0; u < UN; u++)
    {
        int tu = min((warp * UN + u) * TPG + ld_tu, n - 1);
        x[u] = res_quad(fenc + offF[tu] + (intptr_t)ld_row * sf + ld_col, pred + offP[tu] + (intptr_t)ld_row * sp + ld_col);
    }
    for (int grp = warp; grp < ngroups; grp += nwarps)
    {
        uint32_t blo[UN], bhi[UN], z[UN];
#pragma unroll
        for (int u = 0; u < UN; u++) { split4(x[u], blo[u], bhi[u]); z[u] = sumsq4(x[u]); }
        int nxt = grp + nwarps;
        if (nxt < ngroups)
        {
#pragma unroll
            for (int u = 0; u < UN; u++)
            {
                int tu = min((nxt * UN + u) * TPG + ld_tu, n - 1);
                x[u] = res_quad(fenc + offF[tu] + (intptr_t)ld_row * sf + ld_col, pred + offP[tu] + (intptr_t)ld_row * sp + ld_col);
            }
        }
#pragma unroll
        for (int u = 0; u < UN; u++)
        {
            int base = (grp * UN + u) * TPG;
            if (base >= n) break;
            int chi[4] = { 0, 0, 0, 0 }, clo[4] = { add1, add1, add1, add1 };
            imma16_ss(chi, a1, bhi[u]);
            imma16_su(clo, a1, blo[u]);
            uint32_t b2lo, b2hi;
            pack4(recombine(chi[0], clo[0], shift1), recombine(chi[1], clo[1], shift1),
                  recombine(chi[2], clo[2], shift1), recombine(chi[3], clo[3], shift1), b2lo, b2hi);
            int dhi[4] = { 0, 0, 0, 0 }, dlo[4] = { add2, add2, add2, add2 };
            imma16_ss(dhi, a2, b2hi);
            imma16_su(dlo, a2, b2lo);
            uint2 quad = pair_to_quad(pack2(recombine(dhi[0], dlo[0], shift2), recombine(dhi[1], dlo[1], shift2)),
                                      pack2(recombine(dhi[2], dlo[2], shift2), recombine(dhi[3], dlo[3], shift2)), t);
            int tuS = base + st_tu;
            uint32_t sig = 0;
            if (tuS < n) sig = (uint32_t)quant_quad_store(quad, quantCoeff, st_pos, P, qCoef + (size_t)tuS * NN);
            sig = lane_bits_sum<ST_MASK>(sig);
            if (tuS < n && (lane & ST_MASK) == 0) numSig[tuS] = sig;
            uint32_t zs = lane_bits_sum<LD_MASK>(z[u]);
            int tuL = base + ld_tu;
            if (sseZero && tuL < n && (lane & LD_MASK) == 0) sseZero[tuL] = zs;
        }
    }
}

template<typename T, int SMALL, int UN>
__global__ void __launch_bounds__(128)
tu_inv_small_kernel(const int16_t* __restrict__ qCoef, const uint32_t* __restrict__ numSig, int n, QuantP P,
                    const T* __restrict__ fenc, intptr_t sf, const T* __restrict__ pred, intptr_t sp,
                    const int32_t* __restrict__ offF, const int32_t* __restrict__ offP,
                    T* __restrict__ recon, intptr_t sr, const int32_t* __restrict__ offR,
                    unsigned long long* __restrict__ sseRecon, int shift1, int shift2, int depth, int kind)
{
    constexpr int TPG = SMALL == 8 ? 2 : 8;
    constexpr int NN = SMALL * SMALL;
    // kind 1 (SMALL == 4 only): inverse DST-VII, and no DC-only shortcut (quant.cpp:585-588).  outputs after pair_to_quad: N = 8 TU t&1, row g, columns 4(t>>1)..+3 (lane bits 1,2,3,4 vary inside the TU);
    // N = 4 TU (t>>1)*4 + (g>>2) + 2(t&1), row g&3 (bits 2,3 vary)
    constexpr int ST_MASK = SMALL == 8 ? 0x1e : 0x0c;
    int lane = threadIdx.x & 31;
    int warp = blockIdx.x * 4 + (threadIdx.x >> 5);
    int nwarps = gridDim.x * 4;
    int ngroups = (n + TPG * UN - 1) / (TPG * UN);
    if (warp >= ngroups) return;
    int g = lane >> 2, t = lane & 3;
    uint32_t a1[2], b2;
    if (SMALL == 8) { a1[0] = c_IA8[0][lane]; a1[1] = c_IA8[1][lane]; b2 = c_IB8[lane]; }
    else { a1[0] = c_IA4[kind][0][lane]; a1[1] = c_IA4[kind][1][lane]; b2 = c_IB4[kind][lane]; }
    const int add1 = 1 << (shift1 - 1), add2 = 1 << (shift2 - 1);
    const uint32_t mx = (uint32_t)((1 << depth) - 1) * 0x10001u, negmx = (uint32_t)(-((1 << depth) - 1) & 0xffff) * 0x10001u;
    const int ld_tu = SMALL == 8 ? (t >> 1) : ((g >> 2) * 4 + t);
    const int ld_off = SMALL == 8 ? (2 * (g & 3) + (t & 1)) * 8 + 4 * (g >> 2) : (g & 3) * 4;
    const int st_tu = SMALL == 8 ? (t & 1) : (t >> 1) * 4 + (g >> 2) + 2 * (t & 1);
    const int st_row = SMALL == 8 ? g : (g & 3);
    const int st_col = SMALL == 8 ? 4 * (t >> 1) : 0;

    uint2 x[UN];
#pragma unroll
    for (int u = 0; u < UN; u++)
        x[u] = __ldg((const uint2*)(qCoef + (size_t)min((warp * UN + u) * TPG + ld_tu, n - 1) * NN + ld_off));
    for (int grp = warp; grp < ngroups; grp += nwarps)
    {
        uint32_t blo[UN], bhi[UN];
#pragma unroll
        for (int u = 0; u < UN; u++)
        {
            uint2 y = make_uint2(dequant_pair(x[u].x, P), dequant_pair(x[u].y, P));
            split4(transpose4x4_s16(y, g & 3), blo[u], bhi[u]);
        }
        int nxt = grp + nwarps;
        if (nxt < ngroups)
        {
#pragma unroll
            for (int u = 0; u < UN; u++)
                x[u] = __ldg((const uint2*)(qCoef + (size_t)min((nxt * UN + u) * TPG + ld_tu, n - 1) * NN + ld_off));
        }
#pragma unroll
        for (int u = 0; u < UN; u++)
        {
            int base = (grp * UN + u) * TPG;
            if (base >= n) break;
            int chi[4] = { 0, 0, 0, 0 }, clo[4] = { add1, add1, add1, add1 };
            imma16_ss(chi, a1, bhi[u]);
            imma16_su(clo, a1, blo[u]);
            uint32_t alo[2], ahi[2];
            split4(make_uint2(recombine_sat2(chi[0], clo[0], chi[1], clo[1], shift1), 0u), alo[0], ahi[0]);
            split4(make_uint2(recombine_sat2(chi[2], clo[2], chi[3], clo[3], shift1), 0u), alo[1], ahi[1]);
            int dhi[4] = { 0, 0, 0, 0 }, dlo[4] = { add2, add2, add2, add2 };
            imma16_ss(dhi, ahi, b2);
            imma16_us(dlo, alo, b2);
            uint2 quad = pair_to_quad(recombine_sat2(dhi[0], dlo[0], dhi[1], dlo[1], shift2),
                                      recombine_sat2(dhi[2], dlo[2], dhi[3], dlo[3], shift2), t);
            int tu = base + st_tu;
            uint32_t d = 0;
            if (tu < n)
            {
                uint32_t ns = numSig[tu];
                int q0 = qCoef[(size_t)tu * NN];
                if (ns == 1 && q0 != 0 && !kind)
                {
                    int dcv = dc_fill_value(dequant_one(q0, P), depth);
                    quad = make_uint2(pack2(dcv, dcv), pack2(dcv, dcv));
                }
                recon_quad(fenc + offF[tu] + (intptr_t)st_row * sf + st_col, pred + offP[tu] + (intptr_t)st_row * sp + st_col,
                           recon + offR[tu] + (intptr_t)st_row * sr + st_col, (int)ns, quad, mx, negmx, d);
            }
            d = lane_bits_sum<ST_MASK>(d);
            if (tu < n && (lane & ST_MASK) == 0) sseRecon[tu] = d;
        }
    }
}

// ------------------------------------------------------------------------------------------------ launchers
template<typename T>
static void launch_tu_fwd(int sms, int N, const T* fenc, intptr_t sf, const T* pred, intptr_t sp, const int32_t* offF, const int32_t* offP, int n,
                          const int32_t* quantCoeff, QuantP P, int shift1, int shift2, int16_t* qCoef, uint32_t* numSig,
                          unsigned long long* sseZero, cudaStream_t st, int dst4)
{
    int grid;
    if (N == 32)
    {
        // 4 resident CTAs per SM (106 registers): forcing 5 or 6 spills and measured 20-45 % slower
        grid = PGRID((tu_fwd32_kernel<T, 4>));
        if (grid > ceil_div(n, 4)) grid = ceil_div(n, 4);
        tu_fwd32_kernel<T, 4><<<grid, 128, 0, st>>>(fenc, sf, pred, sp, offF, offP, n, quantCoeff, P, shift1, shift2, qCoef, numSig, sseZero);
    }
    else if (N == 16)
    {
        int need = ceil_div(ceil_div(n, 4), 4);
        grid = PGRID((tu_fwd16_kernel<T, 4>));
        if (grid > need) grid = need;
        tu_fwd16_kernel<T, 4><<<grid, 128, 0, st>>>(fenc, sf, pred, sp, offF, offP, n, quantCoeff, P, shift1, shift2, qCoef, numSig, sseZero);
    }
    else if (N == 8)
    {
        int need = ceil_div(ceil_div(n, 2 * 4), 4);
        grid = PGRID((tu_fwd_small_kernel<T, 8, 4>));
        if (grid > need) grid = need;
        tu_fwd_small_kernel<T, 8, 4><<<grid, 128, 0, st>>>(fenc, sf, pred, sp, offF, offP, n, quantCoeff, P, shift1, shift2, qCoef, numSig, sseZero, 0);
    }
    else
    {
        int need = ceil_div(ceil_div(n, 8 * 4), 4);
        grid = PGRID((tu_fwd_small_kernel<T, 4, 4>));
        if (grid > need) grid = need;
        tu_fwd_small_kernel<T, 4, 4><<<grid, 128, 0, st>>>(fenc, sf, pred, sp, offF, offP, n, quantCoeff, P, shift1, shift2, qCoef, numSig, sseZero, dst4 ? 1 : 0);
    }
}

template<typename T>
static void launch_tu_inv(int sms, int N, const int16_t* qCoef, const uint32_t* numSig, int n, QuantP P, const T* fenc, intptr_t sf,
                          const T* pred, intptr_t sp, const int32_t* offF, const int32_t* offP, T* recon, intptr_t sr, const int32_t* offR,
                          unsigned long long* sseRecon, int shift1, int shift2, int depth, cudaStream_t st, int dst4)
{
    int grid;
    if (N == 32)
    {
        grid = PGRID((tu_inv32_kernel<T, 4>));
        if (grid > ceil_div(n, 4)) grid = ceil_div(n, 4);
        tu_inv32_kernel<T, 4><<<grid, 128, 0, st>>>(qCoef, numSig, n, P, fenc, sf, pred, sp, offF, offP, recon, sr, offR, sseRecon, shift1, shift2, depth);
    }
    else if (N == 16)
    {
        int need = ceil_div(ceil_div(n, 4), 4);
        grid = PGRID((tu_inv16_kernel<T, 4>));
        if (grid > need) grid = need;
        tu_inv16_kernel<T, 4><<<grid, 128, 0, st>>>(qCoef, numSig, n, P, fenc, sf, pred, sp, offF, offP, recon, sr, offR, sseRecon, shift1, shift2, depth);
    }
    else if (N == 8)
    {
        int need = ceil_div(ceil_div(n, 2 * 4), 4);
        grid = PGRID((tu_inv_small_kernel<T, 8, 4>));
        if (grid > need) grid = need;
        tu_inv_small_kernel<T, 8, 4><<<grid, 128, 0, st>>>(qCoef, numSig, n, P, fenc, sf, pred, sp, offF, offP, recon, sr, offR, sseRecon, shift1, shift2, depth, 0);
    }
    else
    {
        int need = ceil_div(ceil_div(n, 8 * 4), 4);
        grid = PGRID((tu_inv_small_kernel<T, 4, 4>));
        if (grid > need) grid = need;
        tu_inv_small_kernel<T, 4, 4><<<grid, 128, 0, st>>>(qCoef, numSig, n, P, fenc, sf, pred, sp, offF, offP, recon, sr, offR, sseRecon, shift1, shift2, depth, dst4 ? 1 : 0);
    }
}

// the fused chain over n TUs of size N.  Returns false if the operands do not meet the alignment the tensor-core path
// needs (nothing launched; the caller falls back to the stage kernels) or if a launch failed.
bool launch_tu_fused(x265b200_ctx* ctx, int N, const void* fenc, intptr_t sf, const void* pred, intptr_t sp,
                     const int32_t* offF, const int32_t* offP, int n, const int32_t* quantCoeff, int qBits, int qAdd,
                     int dqScale, int dqShift, int16_t* qCoef, uint32_t* numSig, void* recon, intptr_t sr,
                     const int32_t* offR, uint64_t* sseZero, uint64_t* sseRecon, cudaStream_t st, int dst4)
{
    if (((uintptr_t)qCoef & 7) || ((uintptr_t)quantCoeff & 15) || ((sf | sp) & 3)) return false;
    int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
    int lg = N == 32 ? 5 : N == 16 ? 4 : N == 8 ? 3 : 2;
    int d8 = ctx->depth - 8;
    QuantP P; P.qBits = qBits; P.qAdd = qAdd; P.dqScale = dqScale; P.dqAdd = 1 << (dqShift - 1); P.dqShift = dqShift;
    if (ctx->pixbytes == 1)
    {
        launch_tu_fwd<uint8_t>(sms, N, (const uint8_t*)fenc, sf, (const uint8_t*)pred, sp, offF, offP, n, quantCoeff, P, lg - 1 + d8, lg + 6, qCoef, numSig,
                               (unsigned long long*)sseZero, st, dst4);
        launch_tu_inv<uint8_t>(sms, N, qCoef, numSig, n, P, (const uint8_t*)fenc, sf, (const uint8_t*)pred, sp, offF, offP, (uint8_t*)recon, sr, offR,
                               (unsigned long long*)sseRecon, 7, 12 - d8, ctx->depth, st, dst4);
    }
    else
    {
        launch_tu_fwd<uint16_t>(sms, N, (const uint16_t*)fenc, sf, (const uint16_t*)pred, sp, offF, offP, n, quantCoeff, P, lg - 1 + d8, lg + 6, qCoef, numSig,
                                (unsigned long long*)sseZero, st, dst4);
        launch_tu_inv<uint16_t>(sms, N, qCoef, numSig, n, P, (const uint16_t*)fenc, sf, (const uint16_t*)pred, sp, offF, offP, (uint16_t*)recon, sr, offR,
                                (unsigned long long*)sseRecon, 7, 12 - d8, ctx->depth, st, dst4);
    }
    ctx->launches.fetch_add(2, std::memory_order_relaxed);
    return cudaGetLastError() == cudaSuccess;
}
// forward half alone: residual + DCT + quant -> qCoef / numSig (/ sseZero); same preconditions as launch_tu_fused
bool launch_tu_forward(x265b200_ctx* ctx, int N, const void* fenc, intptr_t sf, const void* pred, intptr_t sp,
                       const int32_t* offF, const int32_t* offP, int n, const int32_t* quantCoeff, int qBits, int qAdd,
                       int16_t* qCoef, uint32_t* numSig, uint64_t* sseZero, cudaStream_t st, int dst4)
{
    if (((uintptr_t)qCoef & 7) || ((uintptr_t)quantCoeff & 15) || ((sf | sp) & 3)) return false;
    int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
    int lg = N == 32 ? 5 : N == 16 ? 4 : N == 8 ? 3 : 2;
    int d8 = ctx->depth - 8;
    QuantP P; P.qBits = qBits; P.qAdd = qAdd; P.dqScale = 0; P.dqAdd = 0; P.dqShift = 1;
    if (ctx->pixbytes == 1)
        launch_tu_fwd<uint8_t>(sms, N, (const uint8_t*)fenc, sf, (const uint8_t*)pred, sp, offF, offP, n, quantCoeff, P, lg - 1 + d8, lg + 6, qCoef, numSig,
                               (unsigned long long*)sseZero, st, dst4);
    else
        launch_tu_fwd<uint16_t>(sms, N, (const uint16_t*)fenc, sf, (const uint16_t*)pred, sp, offF, offP, n, quantCoeff, P, lg - 1 + d8, lg + 6, qCoef, numSig,
                                (unsigned long long*)sseZero, st, dst4);
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError() == cudaSuccess;
}
