// idwt8.cuh -- the wide streaming synthesis kernel: 2*NP columns per lane (NP = kWideNP pairs).
//
// Same streaming scheme as idwt_stream_kernel (idwt.cu), but every lane owns NP column pairs:
// each step fetches one 64/128-bit vector per subband (NP consecutive coefficients), lifts
// 2*NP columns vertically in registers, lifts the finished rows horizontally mostly inside the
// thread (only the samples next to the lane boundary come from the neighbour lanes by shuffle)
// and stores 2*NP packed uint16 samples with one vector store per row.  This cuts the per-sample
// instruction count (address math, loads, shuffles, stores) against the two-column kernel; it
// needs subband widths that are multiples of NP and vector-aligned rows, otherwise the launcher
// falls back to the two-column kernel.  Raw row pairs are staged in shared memory by cp.async
// (4 pairs ahead, lane-private slots), which keeps the register ring down to the live window of
// the vertical lifting and the unrolled step block under the 32 KB instruction cache
// (B300_MICROARCH.md, I-cache) -- with the look-ahead held in registers the NP = 4 block did not
// fit and the kernel stalled on instruction fetch.
#pragma once

#include "wide_common.cuh"

namespace vc2b {

template <int FV, int FH, int STEPS_ = 64>
struct Strip8Cfg {
    static constexpr int NP = kWideNP;
    static constexpr StreamPlan P = make_stream_plan(FV);
    static constexpr StreamPlan PH = make_stream_plan(FH);
    // raw pairs are staged in shared memory by cp.async, so the register ring only holds the live
    // window of the vertical lifting
    // (each row parity enters the ring only when a stage first touches it: P.entry)
    static constexpr int RING = P.window <= 4 ? 4 : (P.window <= 8 ? 8 : 16);
    static constexpr int EMAX = P.entry[0] > P.entry[1] ? P.entry[0] : P.entry[1];
    static constexpr int STAGES = 4 + EMAX;   // shared-memory slots: 4 pairs of look-ahead + late entries
    static constexpr int STEPS = STEPS_;                                       // steps per strip (multiple of RING)
    static constexpr int RSP = STEPS - P.warm - P.maxlag;                      // output row pairs per strip
    static constexpr int HLQ = (get_filter(FH).halo + 2 * NP - 1) / (2 * NP);   // halo lanes left / right
    static constexpr int WOUT = 2 * NP * (32 - 2 * HLQ);                        // output columns per strip
    static constexpr int hmin = PH.dmin[0] < PH.dmin[1] ? PH.dmin[0] : PH.dmin[1];
    static constexpr int hmax = PH.dmax[0] > PH.dmax[1] ? PH.dmax[0] : PH.dmax[1];
    // vertical ring small enough for registers; horizontal neighbours all in the adjacent lane
    static constexpr bool kSupported = (RING <= 8) && (-hmin <= NP) && (hmax <= NP) && (PH.ns <= 2 || FH == 6);
};

struct Strip8Ctx {
    const WideVec *ll, *b0, *b1, *b2;  // row-major, w/NP vectors per row
    int wq, h, jq, jqc, lane, lane_lo, lane_hi;
    bool lane_ok, hedge, ld_hints;
    int shift, cadd, vmax, vmaxs;      // final: v = clamp((x + cadd) >> shift, 0, vmax); vmaxs = (vmax << shift) | (2^shift - 1)
    int rnd;
    uint16_t *pic;
    int cw, ch;
    bool vec_store;
    int32_t *out;
    int W2;
    uint64_t pol_ll, pol_hb, pol_out;  // L2 eviction policies of the LL input, the high-pass inputs and the output
    mutable int lo16, hi16;            // kModeOut16: smallest / largest value this lane stored
};

enum { kModeHb16 = 1, kModeLl16 = 2, kModeOut16 = 4 };
#ifndef VC2B_IDWT16_MINB
#define VC2B_IDWT16_MINB 16   // resident warps per SM the all-int16 levels are compiled for
#endif

struct RawPair {
    int ll[kWideNP], b0[kWideNP], b1[kWideNP], b2[kWideNP];
};

// Copies one vector of NP int16 coefficients (half a WideVec) into the low half of a stage slot.
__device__ __forceinline__ void cp_async_half(WideVec *smem_dst, const void *gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (sizeof(WideVec) == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}

// NP int16 coefficients (the low half of a stage slot) widened to int32
__device__ __forceinline__ void half_to_array(const WideVec *slot, int (&a)[kWideNP])
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(slot);
#pragma unroll
    for (int i = 0; i < kWideNP / 2; i++) {
        const uint32_t v = w[i];
        a[2 * i] = (int)(short)(v & 0xffffu);
        a[2 * i + 1] = (int)v >> 16;
    }
}

// a 16-byte slot holding two vectors of NP int16 coefficients each, widened to int32
__device__ __forceinline__ void int16x8_to_arrays(const int4 &v, int (&a)[4], int (&b)[4])
{
    a[0] = (int)(short)(v.x & 0xffff); a[1] = v.x >> 16; a[2] = (int)(short)(v.y & 0xffff); a[3] = v.y >> 16;
    b[0] = (int)(short)(v.z & 0xffff); b[1] = v.z >> 16; b[2] = (int)(short)(v.w & 0xffff); b[3] = v.w >> 16;
}
__device__ __forceinline__ void int16x8_to_arrays(const int2 &v, int (&a)[2], int (&b)[2])
{
    a[0] = (int)(short)(v.x & 0xffff); a[1] = v.x >> 16;
    b[0] = (int)(short)(v.y & 0xffff); b[1] = v.y >> 16;
}

// issue the asynchronous copy of pair m (one commit group per pair); NS = number of stage slots.
// MODE (the int16 planes of vc2b_idwt16): kModeHb16 -- the high-pass bands are int16 (the coefficient plane
// written by vc2b_unpack16), X.b0..b2 point at int16 rows and a vector is NP * 2 bytes; kModeLl16 -- so is the
// LL input (the previous level's output); kModeOut16 -- this level writes its output as int16.
template <int NS, int MODE>
__device__ __forceinline__ void stage_pair8(const Strip8Ctx &X, WideVec *smem, int m)
{
    const int sy = min(max(m, 0), X.h - 1);
    const int idx = sy * X.wq + X.jqc;
    constexpr int SLOTS = (MODE & kModeLl16) ? 2 : 4;  // vectors per lane and stage
    WideVec *st = smem + ((((m % NS) + NS) % NS) * SLOTS) * 32 + X.lane;
    if constexpr ((MODE & kModeLl16) != 0) {
        // all four inputs are int16: two 16-byte slots per lane -- {LL, HL} feed the even row, {LH, HH} the odd one
        constexpr int HB = kWideNP * 2;  // bytes per int16 vector
        char *lo = reinterpret_cast<char *>(st), *hi = reinterpret_cast<char *>(st + 32);
        cp_async_half(reinterpret_cast<WideVec *>(lo), reinterpret_cast<const char *>(X.ll) + (long long)idx * HB);
        cp_async_half(reinterpret_cast<WideVec *>(lo + HB), reinterpret_cast<const char *>(X.b0) + (long long)idx * HB);
        cp_async_half(reinterpret_cast<WideVec *>(hi), reinterpret_cast<const char *>(X.b1) + (long long)idx * HB);
        cp_async_half(reinterpret_cast<WideVec *>(hi + HB), reinterpret_cast<const char *>(X.b2) + (long long)idx * HB);
    } else if constexpr ((MODE & kModeHb16) != 0) {
        constexpr int HB = kWideNP * 2;
        cp_async_vec(st, X.ll + idx);
        cp_async_half(st + 32, reinterpret_cast<const char *>(X.b0) + (long long)idx * HB);
        cp_async_half(st + 64, reinterpret_cast<const char *>(X.b1) + (long long)idx * HB);
        cp_async_half(st + 96, reinterpret_cast<const char *>(X.b2) + (long long)idx * HB);
    } else if (X.ld_hints) {
        cp_async_vec(st, X.ll + idx, X.pol_ll);
        cp_async_vec(st + 32, X.b0 + idx, X.pol_hb);
        cp_async_vec(st + 64, X.b1 + idx, X.pol_hb);
        cp_async_vec(st + 96, X.b2 + idx, X.pol_hb);
    } else {
        cp_async_vec(st, X.ll + idx);
        cp_async_vec(st + 32, X.b0 + idx);
        cp_async_vec(st + 64, X.b1 + idx);
        cp_async_vec(st + 96, X.b2 + idx);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// At step t the even row of pair t - E0 and the odd row of pair t - E1 enter the register ring.
// The youngest of them was issued LOOK - 1 commit groups before the newest group in flight.
template <int NS, int E0, int E1, int MODE, int LOOK = 4>
__device__ __forceinline__ void load_pair8(const Strip8Ctx &X, const WideVec *smem, int t, RawPair &r)
{
    asm volatile("cp.async.wait_group %0;" ::"n"(LOOK - 1) : "memory");
    const int me = t - E0, mo = t - E1;
    constexpr int SLOTS = (MODE & kModeLl16) ? 2 : 4;
    const WideVec *se = smem + ((((me % NS) + NS) % NS) * SLOTS) * 32 + X.lane;
    const WideVec *so = smem + ((((mo % NS) + NS) % NS) * SLOTS) * 32 + X.lane;
    if constexpr ((MODE & kModeLl16) != 0) {
        int16x8_to_arrays(se[0], r.ll, r.b0);   // one 16-byte shared load per row
        int16x8_to_arrays(so[32], r.b1, r.b2);
    } else {
        vec_to_array(se[0], r.ll);
        if constexpr ((MODE & kModeHb16) != 0) {
            half_to_array(se + 32, r.b0);
            half_to_array(so + 64, r.b1);
            half_to_array(so + 96, r.b2);
        } else {
            vec_to_array(se[32], r.b0);
            vec_to_array(so[64], r.b1);
            vec_to_array(so[96], r.b2);
        }
    }
}

// LL/HL feed the even row (even/odd columns), LH/HH the odd row (vh_synthesis :164-169)
template <int RING>
__device__ __forceinline__ void put_pair8(PairRing<RING, kWideNC> &R, int slot_even, int slot_odd, const RawPair &r)
{
    // the slots must be compile-time constants at every call site (the call is force-inlined)
#pragma unroll
    for (int q = 0; q < kWideNP; q++) {
        R.v[0][2 * q][slot_even] = r.ll[q];
        R.v[0][2 * q + 1][slot_even] = r.b0[q];
        R.v[1][2 * q][slot_odd] = r.b1[q];
        R.v[1][2 * q + 1][slot_odd] = r.b2[q];
    }
}

template <int FH, bool FINAL, bool OUT16 = false>
__device__ __forceinline__ void finish_row8(const Strip8Ctx &X, int gy, int (&E)[kWideNP], int (&O)[kWideNP])
{
    constexpr int NP = kWideNP, NC = kWideNC;
    hlift8<FH, false, Strip8Ctx>(E, O, X);
    if (FINAL) {
        if (X.lane_ok && gy < X.ch) {
            uint32_t pk[NP];
#pragma unroll
            for (int i = 0; i < NP; i++) {
                // clamp((x + cadd) >> s, 0, vmax) == clamp(x + cadd, 0, (vmax << s) | (2^s - 1)) >> s: one
                // add-min-relu instruction (VIADDMNMX.RELU) and a shift per sample, one PRMT per pair
                const int v0 = __viaddmin_s32_relu(E[i], X.cadd, X.vmaxs) >> X.shift;
                const int v1 = __viaddmin_s32_relu(O[i], X.cadd, X.vmaxs) >> X.shift;
                pk[i] = __byte_perm((uint32_t)v0, (uint32_t)v1, 0x5410);
            }
            uint16_t *d = X.pic + (long long)gy * X.cw + NC * X.jq;
            if (X.vec_store && NC * X.jq + NC <= X.cw) {
                st_hint(d, array_to_uvec(pk), X.pol_out);
            } else {
#pragma unroll
                for (int i = 0; i < NP; i++) {
                    if (NC * X.jq + 2 * i < X.cw) d[2 * i] = (uint16_t)(pk[i] & 0xffff);
                    if (NC * X.jq + 2 * i + 1 < X.cw) d[2 * i + 1] = (uint16_t)(pk[i] >> 16);
                }
            }
        }
    } else if (OUT16) {
        if (X.lane_ok) {  // int16 output (vc2b_idwt16): one 16-byte store per row, range tracked per lane
            uint32_t pk[NP];
#pragma unroll
            for (int i = 0; i < NP; i++) {
                const int v0 = (E[i] + X.rnd) >> X.shift, v1 = (O[i] + X.rnd) >> X.shift;
                X.lo16 = min(X.lo16, min(v0, v1));
                X.hi16 = max(X.hi16, max(v0, v1));
                pk[i] = ((uint32_t)v0 & 0xffffu) | ((uint32_t)v1 << 16);
            }
            int16_t *d = reinterpret_cast<int16_t *>(X.out) + (long long)gy * X.W2 + NC * X.jq;
            st_hint(d, array_to_uvec(pk), X.pol_out);
        }
    } else {
        if (X.lane_ok) {
            int lo[NP], hi[NP];  // interleave E/O back into column order
#pragma unroll
            for (int i = 0; i < NP; i++) {
                const int c0 = 2 * i, c1 = 2 * i + 1;
                const int v0 = (E[i] + X.rnd) >> X.shift, v1 = (O[i] + X.rnd) >> X.shift;
                if (c0 < NP) lo[c0] = v0; else hi[c0 - NP] = v0;
                if (c1 < NP) lo[c1] = v1; else hi[c1 - NP] = v1;
            }
            WideVec *d = reinterpret_cast<WideVec *>(X.out + (long long)gy * X.W2 + NC * X.jq);
            st_hint(d, array_to_vec(lo), X.pol_out);
            st_hint(d + 1, array_to_vec(hi), X.pol_out);
        }
    }
}

template <int FV, int FH, bool FINAL, int U, int MODE, int LOOK = 4>
__device__ __forceinline__ void strip8_step_fast(PairRing<Strip8Cfg<FV, FH>::RING, kWideNC> &R, const Strip8Ctx &X,
                                                 WideVec *smem, int t, int K0, int K1)
{
    using C = Strip8Cfg<FV, FH>;
    constexpr int RING = C::RING, M = RING - 1;
    constexpr int E0 = C::P.entry[0], E1 = C::P.entry[1], NS = LOOK + C::EMAX;
    {
        RawPair r;
        load_pair8<NS, E0, E1, MODE, LOOK>(X, smem, t, r);
        put_pair8<RING>(R, (U - E0) & M, (U - E1) & M, r);
    }
    vstep_fast<FV, U, RING, kWideNC>(R);
    stage_pair8<NS, MODE>(X, smem, t - C::EMAX + NS);  // the slot of pair t - EMAX is free again
    const int kf = t - C::P.maxlag;
    if (kf >= K0 && kf < K1) {
        constexpr int sl = (U - C::P.maxlag) & M;
#pragma unroll
        for (int par = 0; par < 2; par++) {
            int E[kWideNP], O[kWideNP];
#pragma unroll
            for (int i = 0; i < kWideNP; i++) { E[i] = R.v[par][2 * i][sl]; O[i] = R.v[par][2 * i + 1][sl]; }
            finish_row8<FH, FINAL, (MODE & kModeOut16) != 0>(X, 2 * kf + par, E, O);
        }
    }
}

template <int FV, int FH, bool FINAL, int U, int MODE, int LOOK = 4>
__device__ __forceinline__ void strip8_block_fast(PairRing<Strip8Cfg<FV, FH>::RING, kWideNC> &R, const Strip8Ctx &X,
                                                  WideVec *smem, int t0, int K0, int K1)
{
    strip8_step_fast<FV, FH, FINAL, U, MODE, LOOK>(R, X, smem, t0 + U, K0, K1);
    if constexpr (U + 1 < Strip8Cfg<FV, FH>::RING)
        strip8_block_fast<FV, FH, FINAL, U + 1, MODE, LOOK>(R, X, smem, t0, K0, K1);
}

template <int FV, int FH, bool FINAL, int MODE, int LOOK = 4>
__device__ __forceinline__ void strip8_step_slow(PairRing<Strip8Cfg<FV, FH>::RING, kWideNC> &R, const Strip8Ctx &X,
                                                 WideVec *smem, int t, int ks, int last, int K0, int K1)
{
    using C = Strip8Cfg<FV, FH>;
    constexpr int RING = C::RING, M = RING - 1;
    constexpr int E0 = C::P.entry[0], E1 = C::P.entry[1], NS = LOOK + C::EMAX;
    RawPair r;
    load_pair8<NS, E0, E1, MODE, LOOK>(X, smem, t, r);
    const int le = (t - E0 - ks) & M, lo = (t - E1 - ks) & M;
#pragma unroll
    for (int i = 0; i < RING; i++) {
#pragma unroll
        for (int q = 0; q < kWideNP; q++) {
            if (i == le) { R.v[0][2 * q][i] = r.ll[q]; R.v[0][2 * q + 1][i] = r.b0[q]; }
            if (i == lo) { R.v[1][2 * q][i] = r.b1[q]; R.v[1][2 * q + 1][i] = r.b2[q]; }
        }
    }
    vstep_slow<FV, RING, kWideNC>(R, t, ks, last);
    stage_pair8<NS, MODE>(X, smem, t - C::EMAX + NS);
    const int kf = t - C::P.maxlag;
    if (kf >= K0 && kf < K1) {
        const int sl = (kf - ks) & M;
#pragma unroll
        for (int par = 0; par < 2; par++) {
            int E[kWideNP], O[kWideNP];
#pragma unroll
            for (int i = 0; i < kWideNP; i++) {
                E[i] = ring_get_dyn<RING>(R.v[par][2 * i], sl);
                O[i] = ring_get_dyn<RING>(R.v[par][2 * i + 1], sl);
            }
            finish_row8<FH, FINAL, (MODE & kModeOut16) != 0>(X, 2 * kf + par, E, O);
        }
    }
}

template <int FV, int FH, bool FINAL, int STEPS, int MODE>
__device__ __forceinline__ void run_strip8(const Strip8Ctx &X, WideVec *smem, int ty)
{
    using C = Strip8Cfg<FV, FH, STEPS>;
    constexpr int RING = C::RING;
    PairRing<RING, kWideNC> R;
#pragma unroll
    for (int i = 0; i < RING; i++)
#pragma unroll
        for (int c = 0; c < kWideNC; c++) R.v[0][c][i] = R.v[1][c][i] = 0;
    const int last = X.h - 1;
    const int K0 = ty * C::RSP, K1 = min(K0 + C::RSP, X.h);
    const int ks = K0 - C::P.warm;
#pragma unroll
    for (int u = 0; u < C::STAGES; u++) stage_pair8<C::STAGES, MODE>(X, smem, ks - C::EMAX + u);
    const int t_end = K1 - 1 + C::P.maxlag;
    for (int t0 = ks; t0 <= t_end; t0 += RING) {
        if (t0 >= C::P.back && t0 + RING - 1 <= last) {
            strip8_block_fast<FV, FH, FINAL, 0, MODE>(R, X, smem, t0, K0, K1);
        } else {
#pragma unroll 1
            for (int u = 0; u < RING; u++) strip8_step_slow<FV, FH, FINAL, MODE>(R, X, smem, t0 + u, ks, last, K0, K1);
        }
    }
}

template <int FV, int FH, int STEPS, int MODE>
__global__ void __launch_bounds__(32, (MODE & kModeLl16) ? VC2B_IDWT16_MINB : ((MODE & kModeHb16) ? 16 : 1))
idwt_stream8_kernel(LevelArgs A)
{
    using C = Strip8Cfg<FV, FH, STEPS>;
    constexpr int NP = kWideNP, NC = kWideNC;
    __shared__ __align__(16) WideVec smem[C::STAGES * ((MODE & kModeLl16) ? 2 : 4) * 32];
    const int c = blockIdx.y;
    const int pic = blockIdx.z;
    const int tx = blockIdx.x % A.tiles_x[c];
    const int ty = blockIdx.x / A.tiles_x[c];
    if (ty >= A.tiles_y[c]) return;

    Strip8Ctx X;
    X.lane = threadIdx.x;
    X.wq = A.w[c] / NP;
    X.h = A.h[c];
    X.jq = tx * (C::WOUT / NC) - C::HLQ + X.lane;
    X.jqc = min(max(X.jq, 0), X.wq - 1);
    X.lane_lo = C::HLQ - tx * (C::WOUT / NC);
    X.lane_hi = X.lane_lo + X.wq - 1;
    X.lane_ok = (X.lane >= C::HLQ) && (X.lane < 32 - C::HLQ) && (X.jq >= 0) && (X.jq < X.wq);
    if constexpr ((MODE & kModeLl16) != 0)
        X.ll = reinterpret_cast<const WideVec *>(reinterpret_cast<const int16_t *>(A.ll[c]) + (long long)pic * A.ll_pic_stride[c]);
    else
        X.ll = reinterpret_cast<const WideVec *>(A.ll[c] + (long long)pic * A.ll_pic_stride[c]);
    if constexpr ((MODE & kModeHb16) != 0) {  // A.hb* address int16 rows
        const int16_t *h0 = reinterpret_cast<const int16_t *>(A.hb[c][0]), *h1 = reinterpret_cast<const int16_t *>(A.hb[c][1]);
        const int16_t *h2 = reinterpret_cast<const int16_t *>(A.hb[c][2]);
        X.b0 = reinterpret_cast<const WideVec *>(h0 + (long long)pic * A.hb_pic_stride);
        X.b1 = reinterpret_cast<const WideVec *>(h1 + (long long)pic * A.hb_pic_stride);
        X.b2 = reinterpret_cast<const WideVec *>(h2 + (long long)pic * A.hb_pic_stride);
    } else {
        X.b0 = reinterpret_cast<const WideVec *>(A.hb[c][0] + (long long)pic * A.hb_pic_stride);
        X.b1 = reinterpret_cast<const WideVec *>(A.hb[c][1] + (long long)pic * A.hb_pic_stride);
        X.b2 = reinterpret_cast<const WideVec *>(A.hb[c][2] + (long long)pic * A.hb_pic_stride);
    }
    X.shift = A.shift;
    X.rnd = A.shift > 0 ? (1 << (A.shift - 1)) : 0;
    const int half = 1 << (A.depth[c] - 1);
    X.cadd = X.rnd + (half << A.shift);   // ((x + rnd) >> s) + half == (x + rnd + (half << s)) >> s
    X.vmax = 2 * half - 1;
    X.vmaxs = (X.vmax << A.shift) | ((1 << A.shift) - 1);
    X.pic = A.pic[c] + (long long)pic * A.pic_pic_stride;
    X.cw = A.crop_w[c];
    X.ch = A.crop_h[c];
    X.vec_store = ((X.cw % NC) == 0) && ((reinterpret_cast<uintptr_t>(X.pic) % (2 * NC)) == 0);
    if constexpr ((MODE & kModeOut16) != 0)
        X.out = reinterpret_cast<int32_t *>(reinterpret_cast<int16_t *>(A.out[c]) + (long long)pic * A.out_pic_stride[c]);
    else
        X.out = A.out[c] + (long long)pic * A.out_pic_stride[c];
    X.lo16 = X.hi16 = 0;
    X.W2 = 2 * A.w[c];
    X.hedge = (X.lane_lo > 0) || (X.lane_hi < 31);
    X.ld_hints = A.hint_loads != 0;
    X.pol_ll = l2_policy(A.hint_ll);
    X.pol_hb = l2_policy(A.hint_hb);
    X.pol_out = l2_policy(A.hint_out);
    if constexpr ((MODE & kModeOut16) != 0) {
        run_strip8<FV, FH, false, STEPS, MODE>(X, smem, ty);
        // a value outside int16: the picture has to be decoded again through the int32 planes (never over a
        // status the unpack set)
        if (__any_sync(0xffffffffu, X.hi16 > 32767 || X.lo16 < -32768) && threadIdx.x == 0)
            atomicCAS(A.status + 4 * (long long)pic, VC2B_ST_OK, VC2B_ST_COEFF16_RANGE);
    } else if constexpr (MODE != 0) {
        run_strip8<FV, FH, true, STEPS, MODE>(X, smem, ty);   // int16 inputs: only vc2b_idwt16's last level
    } else {
        if (A.final_u16) run_strip8<FV, FH, true, STEPS, MODE>(X, smem, ty); else run_strip8<FV, FH, false, STEPS, MODE>(X, smem, ty);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");  // drain look-ahead copies before exit
}

// Whether the wide kernel can run this level: every pointer it dereferences as a vector must be
// vector aligned and every subband row a whole number of vectors.
static bool level_is_wide_ok(const LevelArgs &A, int npic)
{
    constexpr int NP = kWideNP;
    constexpr uintptr_t amask = 4 * NP - 1;  // int32 vectors; int16 vectors need half of it
    const uintptr_t m_ll = A.ll16 ? amask >> 1 : amask, m_hb = A.hb16 ? amask >> 1 : amask;
    for (int c = 0; c < A.ncomp; c++) {
        if (A.w[c] % NP) return false;
        if (reinterpret_cast<uintptr_t>(A.ll[c]) & m_ll) return false;
        if (npic > 1 && (A.ll_pic_stride[c] % NP)) return false;
        for (int k = 0; k < 3; k++)
            if (reinterpret_cast<uintptr_t>(A.hb[c][k]) & m_hb) return false;
        if (!A.final_u16) {
            if (reinterpret_cast<uintptr_t>(A.out[c]) & amask) return false;  // 16-byte stores in both modes
            if (npic > 1 && (A.out_pic_stride[c] % (A.out16 ? 2 * NP : NP))) return false;
        }
    }
    if (npic > 1 && (A.hb_pic_stride % NP)) return false;  // elements: int32 or int16 (hb16)
    return true;
}

// Launches the wide kernel with the strip height that gives the GPU enough warps: a strip is a
// serial chain of STEPS steps, so small levels use short strips (more, shorter chains) and pay a
// few more warm-up rows.
template <int FV, int FH, int STEPS>
static int launch_wide_steps(LevelArgs A, int npic, cudaStream_t s)
{
    using C8 = Strip8Cfg<FV, FH, STEPS>;
    int maxtiles = 0;
    for (int c = 0; c < A.ncomp; c++) {
        A.tiles_x[c] = (2 * A.w[c] + C8::WOUT - 1) / C8::WOUT;
        A.tiles_y[c] = (A.h[c] + C8::RSP - 1) / C8::RSP;
        maxtiles = max(maxtiles, A.tiles_x[c] * A.tiles_y[c]);
    }
    dim3 grid(maxtiles, A.ncomp, npic);
    const int mode = (A.hb16 ? kModeHb16 : 0) | (A.ll16 ? kModeLl16 : 0) | (A.out16 ? kModeOut16 : 0);
    if (mode == 0) {
        idwt_stream8_kernel<FV, FH, STEPS, 0><<<grid, 32, 0, s>>>(A);
    } else if constexpr (FV == FH) {  // the int16 planes need one filter for both directions (coeff16_transform_ok)
        switch (mode) {
        case kModeHb16: idwt_stream8_kernel<FV, FH, STEPS, kModeHb16><<<grid, 32, 0, s>>>(A); break;
        case kModeHb16 | kModeLl16: idwt_stream8_kernel<FV, FH, STEPS, kModeHb16 | kModeLl16><<<grid, 32, 0, s>>>(A); break;
        case kModeHb16 | kModeOut16: idwt_stream8_kernel<FV, FH, STEPS, kModeHb16 | kModeOut16><<<grid, 32, 0, s>>>(A); break;
        case kModeHb16 | kModeLl16 | kModeOut16:
            idwt_stream8_kernel<FV, FH, STEPS, kModeHb16 | kModeLl16 | kModeOut16><<<grid, 32, 0, s>>>(A);
            break;
        default: return VC2B_E_UNSUPPORTED;
        }
    } else {
        return VC2B_E_UNSUPPORTED;
    }
    VC2B_CHECK_LAUNCH();
    return 0;
}

template <int FV, int FH>
static int launch_wide(const LevelArgs &A, int npic, cudaStream_t s)
{
    // A strip is a serial chain of STEPS steps of which RSP produce output (the rest is warm-up and
    // pipeline drain: 14 of them for Fidelity, 4 for LeGall).  Estimated duration of the launch with
    // strip height S: (waves of resident warps, at least one, plus half a wave of tail) x S steps;
    // short strips win only while the launch is a few waves and their RSP / STEPS is still decent.
    const double slots = 148.0 * (A.hb16 ? 16 : 14);  // resident warps (the int16 modes are compiled for 16 per SM)
    double best = 0;
    int pick = 64;
    const int rsp[3] = {Strip8Cfg<FV, FH, 64>::RSP, Strip8Cfg<FV, FH, 32>::RSP, Strip8Cfg<FV, FH, 16>::RSP};
    const int steps[3] = {64, 32, 16};
    for (int k = 0; k < 3; k++) {
        if (rsp[k] < 1) continue;
        double warps = 0;
        for (int c = 0; c < A.ncomp; c++) {
            const long long tx = (2 * A.w[c] + Strip8Cfg<FV, FH, 64>::WOUT - 1) / Strip8Cfg<FV, FH, 64>::WOUT;
            warps += (double)tx * ((A.h[c] + rsp[k] - 1) / rsp[k]) * npic;
        }
        const double waves = warps / slots;
        const double cost = ((waves > 1 ? waves : 1) + 0.5) * steps[k];
        if (best == 0 || cost < best) { best = cost; pick = steps[k]; }
    }
    if (pick == 64) return launch_wide_steps<FV, FH, 64>(A, npic, s);
    if constexpr (Strip8Cfg<FV, FH, 32>::RSP >= 1) {
        if (pick == 32) return launch_wide_steps<FV, FH, 32>(A, npic, s);
    }
    if constexpr (Strip8Cfg<FV, FH, 16>::RSP >= 1) {
        if (pick == 16) return launch_wide_steps<FV, FH, 16>(A, npic, s);
    }
    return launch_wide_steps<FV, FH, 64>(A, npic, s);
}

}  // namespace vc2b
