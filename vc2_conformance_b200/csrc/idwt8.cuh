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
    int shift, cadd, vmax;             // final: v = clamp((x + cadd) >> shift, 0, vmax)
    int rnd;
    uint16_t *pic;
    int cw, ch;
    bool vec_store;
    int32_t *out;
    int W2;
    uint64_t pol_ll, pol_hb, pol_out;  // L2 eviction policies of the LL input, the high-pass inputs and the output
    int4 *ring;                        // fused pair of levels: shared LL ring (see idwt_fused2_kernel)
    int ring_q;                        // consumer: this lane's group of four LL columns inside the ring row
};

// Roles of a warp in the fused two-level kernel (0: the stand-alone per-level kernel).
enum { kRolePlain = 0, kRoleProducer = 1, kRoleConsumer = 2 };
constexpr int kRingRows = 4;        // LL rows held by the ring: the pair being read and the pair being written
constexpr int kRingPlane = 36;      // int4 per plane: 32 lanes + 64 bytes so that the two planes use different banks
constexpr int kRingRowVecs = 2 * kRingPlane;
constexpr int kFusedThreads = 96;

__device__ __forceinline__ void fused_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kFusedThreads) : "memory"); }

struct RawPair {
    int ll[kWideNP], b0[kWideNP], b1[kWideNP], b2[kWideNP];
};

// issue the asynchronous copy of pair m (one commit group per pair); NS = number of stage slots
template <int NS, int ROLE = kRolePlain>
__device__ __forceinline__ void stage_pair8(const Strip8Ctx &X, WideVec *smem, int m)
{
    const int sy = min(max(m, 0), X.h - 1);
    const int idx = sy * X.wq + X.jqc;
    WideVec *st = smem + ((((m % NS) + NS) % NS) * 4) * 32 + X.lane;
    if constexpr (ROLE == kRoleConsumer) {  // the LL rows come from the producer warp through the ring
        cp_async_vec(st + 32, X.b0 + idx);
        cp_async_vec(st + 64, X.b1 + idx);
        cp_async_vec(st + 96, X.b2 + idx);
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
template <int NS, int E0, int E1, int ROLE = kRolePlain, int LOOK = 4>
__device__ __forceinline__ void load_pair8(const Strip8Ctx &X, const WideVec *smem, int t, RawPair &r)
{
    asm volatile("cp.async.wait_group %0;" ::"n"(LOOK - 1) : "memory");
    const int me = t - E0, mo = t - E1;
    const WideVec *se = smem + ((((me % NS) + NS) % NS) * 4) * 32 + X.lane;
    const WideVec *so = smem + ((((mo % NS) + NS) % NS) * 4) * 32 + X.lane;
    if constexpr (ROLE == kRoleConsumer)
        vec_to_array(X.ring[(me & (kRingRows - 1)) * kRingRowVecs + (X.ring_q & 1) * kRingPlane + (X.ring_q >> 1)], r.ll);
    else
        vec_to_array(se[0], r.ll);
    vec_to_array(se[32], r.b0);
    vec_to_array(so[64], r.b1);
    vec_to_array(so[96], r.b2);
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

template <int FH, bool FINAL, int ROLE = kRolePlain>
__device__ __forceinline__ void finish_row8(const Strip8Ctx &X, int gy, int (&E)[kWideNP], int (&O)[kWideNP])
{
    constexpr int NP = kWideNP, NC = kWideNC;
    hlift8<FH, false, Strip8Ctx>(E, O, X);
    if constexpr (ROLE == kRoleProducer) {
        // the finished LL row of the next level goes to the shared ring: plane 0 holds the first four of
        // the lane's eight columns, plane 1 the last four (both stores conflict-free)
        int lo[NP], hi[NP];
#pragma unroll
        for (int i = 0; i < NP; i++) {
            const int c0 = 2 * i, c1 = 2 * i + 1;
            const int v0 = (E[i] + X.rnd) >> X.shift, v1 = (O[i] + X.rnd) >> X.shift;
            if (c0 < NP) lo[c0] = v0; else hi[c0 - NP] = v0;
            if (c1 < NP) lo[c1] = v1; else hi[c1 - NP] = v1;
        }
        int4 *row = X.ring + (gy & (kRingRows - 1)) * kRingRowVecs;
        row[X.lane] = array_to_vec(lo);
        row[kRingPlane + X.lane] = array_to_vec(hi);
    } else if (FINAL) {
        if (X.lane_ok && gy < X.ch) {
            uint32_t pk[NP];
#pragma unroll
            for (int i = 0; i < NP; i++) {
                const int v0 = min(max((E[i] + X.cadd) >> X.shift, 0), X.vmax);
                const int v1 = min(max((O[i] + X.cadd) >> X.shift, 0), X.vmax);
                pk[i] = (uint32_t)v0 | ((uint32_t)v1 << 16);
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

template <int FV, int FH, bool FINAL, int U, int ROLE = kRolePlain, int LOOK = 4>
__device__ __forceinline__ void strip8_step_fast(PairRing<Strip8Cfg<FV, FH>::RING, kWideNC> &R, const Strip8Ctx &X,
                                                 WideVec *smem, int t, int K0, int K1)
{
    using C = Strip8Cfg<FV, FH>;
    constexpr int RING = C::RING, M = RING - 1;
    constexpr int E0 = C::P.entry[0], E1 = C::P.entry[1], NS = LOOK + C::EMAX;
    {
        RawPair r;
        load_pair8<NS, E0, E1, ROLE, LOOK>(X, smem, t, r);
        put_pair8<RING>(R, (U - E0) & M, (U - E1) & M, r);
    }
    vstep_fast<FV, U, RING, kWideNC>(R);
    stage_pair8<NS, ROLE>(X, smem, t - C::EMAX + NS);  // the slot of pair t - EMAX is free again
    const int kf = t - C::P.maxlag;
    if (kf >= K0 && kf < K1) {
        constexpr int sl = (U - C::P.maxlag) & M;
#pragma unroll
        for (int par = 0; par < 2; par++) {
            int E[kWideNP], O[kWideNP];
#pragma unroll
            for (int i = 0; i < kWideNP; i++) { E[i] = R.v[par][2 * i][sl]; O[i] = R.v[par][2 * i + 1][sl]; }
            finish_row8<FH, FINAL, ROLE>(X, 2 * kf + par, E, O);
        }
    }
    // fused pair of levels: the producer publishes an LL row pair per step, a consumer uses one per two steps
    if constexpr (ROLE == kRoleProducer) fused_barrier();
    if constexpr (ROLE == kRoleConsumer && (U & 1)) fused_barrier();
}

template <int FV, int FH, bool FINAL, int U, int ROLE = kRolePlain, int LOOK = 4>
__device__ __forceinline__ void strip8_block_fast(PairRing<Strip8Cfg<FV, FH>::RING, kWideNC> &R, const Strip8Ctx &X,
                                                  WideVec *smem, int t0, int K0, int K1)
{
    strip8_step_fast<FV, FH, FINAL, U, ROLE, LOOK>(R, X, smem, t0 + U, K0, K1);
    if constexpr (U + 1 < Strip8Cfg<FV, FH>::RING)
        strip8_block_fast<FV, FH, FINAL, U + 1, ROLE, LOOK>(R, X, smem, t0, K0, K1);
}

template <int FV, int FH, bool FINAL, int ROLE = kRolePlain, int LOOK = 4>
__device__ __forceinline__ void strip8_step_slow(PairRing<Strip8Cfg<FV, FH>::RING, kWideNC> &R, const Strip8Ctx &X,
                                                 WideVec *smem, int t, int ks, int last, int K0, int K1)
{
    using C = Strip8Cfg<FV, FH>;
    constexpr int RING = C::RING, M = RING - 1;
    constexpr int E0 = C::P.entry[0], E1 = C::P.entry[1], NS = LOOK + C::EMAX;
    RawPair r;
    load_pair8<NS, E0, E1, ROLE, LOOK>(X, smem, t, r);
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
    stage_pair8<NS, ROLE>(X, smem, t - C::EMAX + NS);
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
            finish_row8<FH, FINAL, ROLE>(X, 2 * kf + par, E, O);
        }
    }
    if constexpr (ROLE == kRoleProducer) fused_barrier();
    if constexpr (ROLE == kRoleConsumer) {
        if ((t - ks) & 1) fused_barrier();
    }
}

template <int FV, int FH, bool FINAL, int STEPS>
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
    for (int u = 0; u < C::STAGES; u++) stage_pair8<C::STAGES>(X, smem, ks - C::EMAX + u);
    const int t_end = K1 - 1 + C::P.maxlag;
    for (int t0 = ks; t0 <= t_end; t0 += RING) {
        if (t0 >= C::P.back && t0 + RING - 1 <= last) {
            strip8_block_fast<FV, FH, FINAL, 0>(R, X, smem, t0, K0, K1);
        } else {
#pragma unroll 1
            for (int u = 0; u < RING; u++) strip8_step_slow<FV, FH, FINAL>(R, X, smem, t0 + u, ks, last, K0, K1);
        }
    }
}

template <int FV, int FH, int STEPS>
__global__ void __launch_bounds__(32)
idwt_stream8_kernel(LevelArgs A)
{
    using C = Strip8Cfg<FV, FH, STEPS>;
    constexpr int NP = kWideNP, NC = kWideNC;
    __shared__ __align__(16) WideVec smem[C::STAGES * 4 * 32];
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
    X.ll = reinterpret_cast<const WideVec *>(A.ll[c] + (long long)pic * A.ll_pic_stride[c]);
    X.b0 = reinterpret_cast<const WideVec *>(A.hb[c][0] + (long long)pic * A.hb_pic_stride);
    X.b1 = reinterpret_cast<const WideVec *>(A.hb[c][1] + (long long)pic * A.hb_pic_stride);
    X.b2 = reinterpret_cast<const WideVec *>(A.hb[c][2] + (long long)pic * A.hb_pic_stride);
    X.shift = A.shift;
    X.rnd = A.shift > 0 ? (1 << (A.shift - 1)) : 0;
    const int half = 1 << (A.depth[c] - 1);
    X.cadd = X.rnd + (half << A.shift);   // ((x + rnd) >> s) + half == (x + rnd + (half << s)) >> s
    X.vmax = 2 * half - 1;
    X.pic = A.pic[c] + (long long)pic * A.pic_pic_stride;
    X.cw = A.crop_w[c];
    X.ch = A.crop_h[c];
    X.vec_store = ((X.cw % NC) == 0) && ((reinterpret_cast<uintptr_t>(X.pic) % (2 * NC)) == 0);
    X.out = A.out[c] + (long long)pic * A.out_pic_stride[c];
    X.W2 = 2 * A.w[c];
    X.hedge = (X.lane_lo > 0) || (X.lane_hi < 31);
    X.ld_hints = A.hint_loads != 0;
    X.pol_ll = l2_policy(A.hint_ll);
    X.pol_hb = l2_policy(A.hint_hb);
    X.pol_out = l2_policy(A.hint_out);
    if (A.final_u16) run_strip8<FV, FH, true, STEPS>(X, smem, ty); else run_strip8<FV, FH, false, STEPS>(X, smem, ty);
    asm volatile("cp.async.wait_group 0;" ::: "memory");  // drain look-ahead copies before exit
}

// ---- two levels in one kernel ------------------------------------------------------------------------
// The last two levels of a transform move the most data, and the LL band between them (a quarter of
// the picture, int32) would otherwise be written to HBM and read straight back.  idwt_fused2_kernel
// keeps it on chip: a CTA is three warps -- a PRODUCER running the strip scheme on level top-1 and two
// CONSUMERS running it on level top, side by side.  Every producer step finishes one LL row pair of the
// last level and stores it into a four-row shared-memory ring instead of global memory; the consumers
// take their LL vectors from the ring (their three high-pass bands still arrive by cp.async) and use
// one row pair per two steps.  One named barrier per producer step keeps the three warps in step: the
// pair written at step s is read after barrier s+1 and overwritten after barrier s+2.  Horizontally the
// producer's 30 valid lanes (240 LL columns) cover the two consumers' 2 x 28 output lanes (224 LL
// columns = 448 picture columns) plus their halo lanes; vertically the producer starts early enough to
// deliver the consumers' warm-up rows.  Warm-up rows and halo lanes are recomputed, not exchanged.
template <int FV, int FH>
struct FusedCfg {
    using C = Strip8Cfg<FV, FH, 64>;
    static constexpr int STEPS = 64;                                            // consumer steps per CTA
    static constexpr int RSPF = (STEPS - C::P.warm - C::P.maxlag - 1) & ~1;     // output row pairs of the last level per CTA
    static constexpr int TLL = 224;                                             // LL columns of the last level per CTA
    static constexpr bool kSupported = C::kSupported && C::RING == 4 && C::HLQ == 1 && kWideNP == 4;
};

#ifndef VC2B_FUSED_LOOK
#define VC2B_FUSED_LOOK 4     // row pairs staged ahead per warp of the fused kernel
#endif
#ifndef VC2B_FUSED_MINB
#define VC2B_FUSED_MINB 4     // resident CTAs per SM the fused kernel is compiled for
#endif
constexpr int kFusedLook = VC2B_FUSED_LOOK;

template <int FV, int FH, bool FINAL, int ROLE>
__device__ __forceinline__ int run_strip8_role(const Strip8Ctx &X, WideVec *smem, int K0, int K1, int ks, int nblocks)
{
    using C = Strip8Cfg<FV, FH, 64>;
    constexpr int RING = C::RING, NS = kFusedLook + C::EMAX;
    PairRing<RING, kWideNC> R;
#pragma unroll
    for (int i = 0; i < RING; i++)
#pragma unroll
        for (int c = 0; c < kWideNC; c++) R.v[0][c][i] = R.v[1][c][i] = 0;
    const int last = X.h - 1;
#pragma unroll
    for (int u = 0; u < NS; u++) stage_pair8<NS, ROLE>(X, smem, ks - C::EMAX + u);
    for (int b = 0; b < nblocks; b++) {
        const int t0 = ks + b * RING;
        if (t0 >= C::P.back && t0 + RING - 1 <= last) {
            strip8_block_fast<FV, FH, FINAL, 0, ROLE, kFusedLook>(R, X, smem, t0, K0, K1);
        } else {
#pragma unroll 1
            for (int u = 0; u < RING; u++)
                strip8_step_slow<FV, FH, FINAL, ROLE, kFusedLook>(R, X, smem, t0 + u, ks, last, K0, K1);
        }
    }
    return nblocks * RING;
}

template <int FV, int FH>
__global__ void __launch_bounds__(kFusedThreads, VC2B_FUSED_MINB)
idwt_fused2_kernel(LevelArgs A3, LevelArgs A4)
{
    using C = Strip8Cfg<FV, FH, 64>;
    using F = FusedCfg<FV, FH>;
    constexpr int NP = kWideNP, NC = kWideNC;
    constexpr int E0 = C::P.entry[0];
    extern __shared__ __align__(16) WideVec fused_smem[];
    constexpr int kStageVecs = (kFusedLook + C::EMAX) * 4 * 32;
    WideVec(*stage)[kStageVecs] = reinterpret_cast<WideVec(*)[kStageVecs]>(fused_smem);
    int4 *ring = reinterpret_cast<int4 *>(fused_smem + 3 * kStageVecs);
    const int c = blockIdx.y;
    const int pic = blockIdx.z;
    const int tx = blockIdx.x % A4.tiles_x[c];
    const int ty = blockIdx.x / A4.tiles_x[c];
    if (ty >= A4.tiles_y[c]) return;  // the whole CTA leaves together
    const int role = threadIdx.x >> 5;  // 0: producer (level top-1); 1, 2: consumers (level top)
    const int lane = threadIdx.x & 31;

    // consumer steps: output row pairs [K0, K1) of the last level; (ks4 - E0) even so that LL row pairs
    // are used by steps (even, odd)
    const int h4 = A4.h[c];
    const int K0 = ty * F::RSPF, K1 = min(K0 + F::RSPF, h4);
    int ks4 = K0 - C::P.warm;
    ks4 -= (ks4 - E0) & 1;
    const int tend4 = K1 - 1 + C::P.maxlag;
    const int blocks4 = (tend4 - ks4 + 1 + C::RING - 1) / C::RING;
    // producer steps: LL rows (ks4 - E0) .. (tend4 - E0) of the last level = its output pairs [J0, J1)
    const int h3 = A3.h[c];
    const int J0 = (ks4 - E0) >> 1;
    const int J1 = min(((tend4 - E0) >> 1) + 1, h3);
    const int ks3 = J0 - C::P.warm;
    const int blocks3 = (J1 - 1 + C::P.maxlag - ks3 + 1 + C::RING - 1) / C::RING;
    // barrier schedule: the producer arrives once per step; a consumer waits until pair J0 is out
    // (warm + maxlag + 1 producer steps), then arrives once per two steps
    constexpr int pre = C::P.warm + C::P.maxlag;
    const int nb3 = blocks3 * C::RING;
    const int nb4 = pre + 1 + blocks4 * (C::RING / 2);
    const int nb = max(nb3, nb4);

    Strip8Ctx X;
    X.lane = lane;
    X.ring = ring;
    X.ring_q = 0;
    X.ld_hints = false;
    X.pol_ll = X.pol_hb = 0;
    int done;
    if (role == 0) {
        X.wq = A3.w[c] / NP;
        X.h = h3;
        X.jq = tx * (F::TLL / NC) - 2 + lane;  // lane 1 holds the eight LL columns left of the tile (consumer halo)
        X.jqc = min(max(X.jq, 0), X.wq - 1);
        X.lane_lo = 2 - tx * (F::TLL / NC);
        X.lane_hi = X.lane_lo + X.wq - 1;
        X.lane_ok = true;
        X.ll = reinterpret_cast<const WideVec *>(A3.ll[c] + (long long)pic * A3.ll_pic_stride[c]);
        X.b0 = reinterpret_cast<const WideVec *>(A3.hb[c][0] + (long long)pic * A3.hb_pic_stride);
        X.b1 = reinterpret_cast<const WideVec *>(A3.hb[c][1] + (long long)pic * A3.hb_pic_stride);
        X.b2 = reinterpret_cast<const WideVec *>(A3.hb[c][2] + (long long)pic * A3.hb_pic_stride);
        X.shift = A3.shift;
        X.rnd = A3.shift > 0 ? (1 << (A3.shift - 1)) : 0;
        X.cadd = 0;
        X.vmax = 0;
        X.pic = nullptr;
        X.cw = X.ch = 0;
        X.vec_store = false;
        X.out = nullptr;
        X.W2 = 0;
        X.pol_out = 0;
        X.hedge = (X.lane_lo > 0) || (X.lane_hi < 31);
        done = run_strip8_role<FV, FH, false, kRoleProducer>(X, stage[0], J0, J1, ks3, blocks3);
    } else {
        const int half_tile = (role - 1) * (F::TLL / 2 / NP);  // second consumer: 28 lanes (112 LL columns) further right
        X.wq = A4.w[c] / NP;
        X.h = h4;
        X.jq = tx * (F::TLL / NP) - 1 + half_tile + lane;  // lane 0 is the left halo lane
        X.jqc = min(max(X.jq, 0), X.wq - 1);
        X.lane_lo = 1 - tx * (F::TLL / NP) - half_tile;
        X.lane_hi = X.lane_lo + X.wq - 1;
        X.lane_ok = (lane >= 1) && (lane <= 28) && (X.jq >= 0) && (X.jq < X.wq);
        X.ring_q = 3 + half_tile + lane;  // the producer's lane 1 (ring groups 2, 3) starts eight LL columns left of the tile
        X.ll = nullptr;
        X.b0 = reinterpret_cast<const WideVec *>(A4.hb[c][0] + (long long)pic * A4.hb_pic_stride);
        X.b1 = reinterpret_cast<const WideVec *>(A4.hb[c][1] + (long long)pic * A4.hb_pic_stride);
        X.b2 = reinterpret_cast<const WideVec *>(A4.hb[c][2] + (long long)pic * A4.hb_pic_stride);
        X.shift = A4.shift;
        X.rnd = A4.shift > 0 ? (1 << (A4.shift - 1)) : 0;
        const int half = 1 << (A4.depth[c] - 1);
        X.cadd = X.rnd + (half << A4.shift);
        X.vmax = 2 * half - 1;
        X.pic = A4.pic[c] + (long long)pic * A4.pic_pic_stride;
        X.cw = A4.crop_w[c];
        X.ch = A4.crop_h[c];
        X.vec_store = ((X.cw % NC) == 0) && ((reinterpret_cast<uintptr_t>(X.pic) % (2 * NC)) == 0);
        X.out = nullptr;
        X.W2 = 2 * A4.w[c];
        X.pol_out = l2_policy(A4.hint_out);
        X.hedge = (X.lane_lo > 0) || (X.lane_hi < 31);
        for (int k = 0; k < pre + 1; k++) fused_barrier();
        done = pre + 1 + run_strip8_role<FV, FH, true, kRoleConsumer>(X, stage[role], K0, K1, ks4, blocks4) / 2;
    }
    for (; done < nb; done++) fused_barrier();
    asm volatile("cp.async.wait_group 0;" ::: "memory");  // drain look-ahead copies before exit
}

// Whether the fused kernel can run the last two levels (given that each of them passes level_is_wide_ok).
template <int FV, int FH>
static int launch_fused2(LevelArgs A3, LevelArgs A4, int npic, cudaStream_t s)
{
    using F = FusedCfg<FV, FH>;
    int maxtiles = 0;
    for (int c = 0; c < A4.ncomp; c++) {
        A4.tiles_x[c] = (A4.w[c] + F::TLL - 1) / F::TLL;
        A4.tiles_y[c] = (A4.h[c] + F::RSPF - 1) / F::RSPF;
        maxtiles = max(maxtiles, A4.tiles_x[c] * A4.tiles_y[c]);
    }
    dim3 grid(maxtiles, A4.ncomp, npic);
    using C = Strip8Cfg<FV, FH, 64>;
    const size_t smem = (3 * (kFusedLook + C::EMAX) * 4 * 32 + kRingRows * kRingRowVecs) * sizeof(WideVec);
    const cudaError_t attr = cudaFuncSetAttribute(idwt_fused2_kernel<FV, FH>,
                                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (attr != cudaSuccess) return (int)attr;
    idwt_fused2_kernel<FV, FH><<<grid, kFusedThreads, smem, s>>>(A3, A4);
    VC2B_CHECK_LAUNCH();
    return 0;
}

// Whether the wide kernel can run this level: every pointer it dereferences as a vector must be
// vector aligned and every subband row a whole number of vectors.
static bool level_is_wide_ok(const LevelArgs &A, int npic)
{
    constexpr int NP = kWideNP;
    constexpr uintptr_t amask = 4 * NP - 1;
    for (int c = 0; c < A.ncomp; c++) {
        if (A.w[c] % NP) return false;
        if (reinterpret_cast<uintptr_t>(A.ll[c]) & amask) return false;
        if (npic > 1 && (A.ll_pic_stride[c] % NP)) return false;
        for (int k = 0; k < 3; k++)
            if (reinterpret_cast<uintptr_t>(A.hb[c][k]) & amask) return false;
        if (!A.final_u16) {
            if (reinterpret_cast<uintptr_t>(A.out[c]) & amask) return false;
            if (npic > 1 && (A.out_pic_stride[c] % NP)) return false;
        }
    }
    if (npic > 1 && (A.hb_pic_stride % NP)) return false;
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
    idwt_stream8_kernel<FV, FH, STEPS><<<grid, 32, 0, s>>>(A);
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
    const double slots = 148.0 * 14;
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
