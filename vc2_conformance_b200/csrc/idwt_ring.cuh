// idwt_ring.cuh -- synthesis level with the vertical lifting window in SHARED memory.
//
// The wide register kernel (idwt8.cuh) unrolls the step loop by the size of its register ring so that
// every slot is a compile-time register.  For the Fidelity filter (two stages of eight taps, 64-bit
// sums, a window of twelve row pairs) that unrolled block is several hundred KB of code: the kernel
// ran out of the instruction cache (ncu: stall_no_instruction 7.2 per issue, 255 registers with
// spills; profiles/r1_v8_idwt_cfg4_summary.txt) at 5 % of the HBM roofline.  Here the window lives in a
// ring of row pairs in shared memory: cp.async writes the raw subband vectors of pair m straight into
// slot m mod RS, the lifting stages update the slots in place, and finished pairs are read back into
// registers for the horizontal lifting (hlift8) and the store (finish_row8), both shared with the wide
// kernel.  Slots are addressed at run time, so there is ONE rolled step loop, the edge clamps of
// lift1-4 (picture_decoding.py:210-212, :240-242) are a min / max on the pair index instead of a
// separate slow path, and the code is a few KB.  Every lane only ever touches its own eight columns
// (vertical lifting is per column), so the ring needs no synchronisation at all.
#pragma once

#include "idwt8.cuh"

namespace vc2b {

template <int FV, int FH, int STEPS_ = 64>
struct RingCfg {
    static constexpr int NP = kWideNP;
    static constexpr StreamPlan P = make_stream_plan(FV);
    static constexpr StreamPlan PH = make_stream_plan(FH);
    static constexpr int LOOK = 4;                                  // pairs staged ahead of the step that needs them
    static constexpr int LIVE = P.back + 1;                         // pairs t - back .. t are read at step t
    static constexpr int RS = (LIVE + LOOK <= 8) ? 8 : ((LIVE + LOOK <= 16) ? 16 : 32);  // ring slots
    static constexpr int STEPS = STEPS_;
    static constexpr int RSP = STEPS - P.warm - P.maxlag;           // output row pairs per strip
    static constexpr int HLQ = (get_filter(FH).halo + 2 * NP - 1) / (2 * NP);
    static constexpr int WOUT = 2 * NP * (32 - 2 * HLQ);
    static constexpr int hmin = PH.dmin[0] < PH.dmin[1] ? PH.dmin[0] : PH.dmin[1];
    static constexpr int hmax = PH.dmax[0] > PH.dmax[1] ? PH.dmax[0] : PH.dmax[1];
    static constexpr bool kSupported = (RS <= 16) && (-hmin <= NP) && (hmax <= NP) && (PH.ns <= 2 || FH == 6) && NP == 4;
    static constexpr int kSlotVecs = 4 * 32;                        // LL | HL | LH | HH vectors of the 32 lanes
    static constexpr size_t kSmemBytes = (size_t)RS * kSlotVecs * sizeof(int4);
};

// slot of pair m: [band 0..3][lane]; bands 0, 1 (LL, HL) are the even / odd columns of the even row,
// bands 2, 3 (LH, HH) those of the odd row (vh_synthesis :164-169)
template <int RS>
__device__ __forceinline__ int4 *ring_slot(int4 *ring, int m, int lane)
{
    return ring + ((m & (RS - 1)) * 4) * 32 + lane;
}

template <int RS, bool COMMIT = true>
__device__ __forceinline__ void ring_stage_pair(const Strip8Ctx &X, int4 *ring, int m)
{
    const int sy = min(max(m, 0), X.h - 1);
    const int idx = sy * X.wq + X.jqc;
    int4 *st = ring_slot<RS>(ring, m, X.lane);
    cp_async_vec(st, X.ll + idx);
    cp_async_vec(st + 32, X.b0 + idx);
    cp_async_vec(st + 64, X.b1 + idx);
    cp_async_vec(st + 96, X.b2 + idx);
    if (COMMIT) asm volatile("cp.async.commit_group;" ::: "memory");
}

// One lifting stage of one step: updates the row of parity `par` of pair k from the rows of the other
// parity of pairs k + dmin .. k + dmax, source pairs clamped to the picture.
template <int FV, int SIDX, int RS>
__device__ __forceinline__ void ring_vstage(int4 *ring, int lane, int t, int last)
{
    constexpr StreamPlan P = make_stream_plan(FV);
    constexpr StageDef st = stream_stage(FV, SIDX, false);
    constexpr int par = P.parity[SIDX];
    constexpr bool add = (st.type == 1 || st.type == 3);
    using Acc = typename std::conditional<(st.S >= kWideAccS), long long, int>::type;
    const int k = t - P.lag[SIDX];
    if (k < 0 || k > last) return;
#pragma unroll
    for (int v = 0; v < 2; v++) {  // the two vectors (even / odd columns) of the row
        constexpr Acc rnd = st.S > 0 ? (Acc)1 << (st.S > 0 ? st.S - 1 : 0) : (Acc)0;
        Acc sum[4] = {rnd, rnd, rnd, rnd};
        auto src = [&](int i) -> int4 {
            int m = k + P.dmin[SIDX] + i;
            m = m < 0 ? 0 : (m > last ? last : m);
            return ring_slot<RS>(ring, m, lane)[(2 * (1 - par) + v) * 32];
        };
        if (sizeof(Acc) == 8) {  // one IMAD.WIDE per tap and column (see folded_taps)
#pragma unroll
            for (int i = 0; i < st.L; i++) {
                const int4 a = src(i);
                sum[0] += (Acc)st.taps[i] * (Acc)a.x;
                sum[1] += (Acc)st.taps[i] * (Acc)a.y;
                sum[2] += (Acc)st.taps[i] * (Acc)a.z;
                sum[3] += (Acc)st.taps[i] * (Acc)a.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < st.L / 2; i++) {
                const int4 a = src(i), b = src(st.L - 1 - i);
                if (st.taps[i] == st.taps[st.L - 1 - i]) {  // symmetric taps: tap * (a + b)
                    sum[0] += (Acc)st.taps[i] * ((Acc)a.x + (Acc)b.x);
                    sum[1] += (Acc)st.taps[i] * ((Acc)a.y + (Acc)b.y);
                    sum[2] += (Acc)st.taps[i] * ((Acc)a.z + (Acc)b.z);
                    sum[3] += (Acc)st.taps[i] * ((Acc)a.w + (Acc)b.w);
                } else {
                    sum[0] += (Acc)st.taps[i] * a.x + (Acc)st.taps[st.L - 1 - i] * b.x;
                    sum[1] += (Acc)st.taps[i] * a.y + (Acc)st.taps[st.L - 1 - i] * b.y;
                    sum[2] += (Acc)st.taps[i] * a.z + (Acc)st.taps[st.L - 1 - i] * b.z;
                    sum[3] += (Acc)st.taps[i] * a.w + (Acc)st.taps[st.L - 1 - i] * b.w;
                }
            }
            if (st.L & 1) {
                const int4 a = src(st.L / 2);
                sum[0] += (Acc)st.taps[st.L / 2] * a.x;
                sum[1] += (Acc)st.taps[st.L / 2] * a.y;
                sum[2] += (Acc)st.taps[st.L / 2] * a.z;
                sum[3] += (Acc)st.taps[st.L / 2] * a.w;
            }
        }
        int d[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            sum[q] >>= st.S;
            d[q] = add ? (int)sum[q] : -(int)sum[q];
        }
        int4 *tp = ring_slot<RS>(ring, k, lane) + (2 * par + v) * 32;
        int4 tv = *tp;
        tv.x += d[0];
        tv.y += d[1];
        tv.z += d[2];
        tv.w += d[3];
        *tp = tv;
    }
}

// The same stage for TWO consecutive steps (pairs k and k + 1) at once.  Within a stage the updates of
// different pairs are independent (a stage only reads the other row parity), and running stage s of step
// t + 1 before stage s + 1 of step t changes nothing either: stage s + 1 at step t reads pairs that stage
// s finished at step t or earlier, and writes a pair that stage s has already read (that is what the
// lags of the streaming plan guarantee).  The two windows overlap in L - 1 of L source rows, so a block
// of two steps loads L + 1 rows instead of 2 L -- the kernel is bound by shared-memory traffic and by
// the latency of these loads -- and the two accumulations interleave.
template <int FV, int SIDX, int RS>
__device__ __forceinline__ void ring_vstage2(int4 *ring, int lane, int t, int last)
{
    constexpr StreamPlan P = make_stream_plan(FV);
    constexpr StageDef st = stream_stage(FV, SIDX, false);
    constexpr int par = P.parity[SIDX];
    constexpr bool add = (st.type == 1 || st.type == 3);
    constexpr int L = st.L;
    using Acc = typename std::conditional<(st.S >= kWideAccS), long long, int>::type;
    const int k = t - P.lag[SIDX];
    if (k < 0 || k + 1 > last || k + P.dmin[SIDX] < 0 || k + 1 + P.dmax[SIDX] > last) {
        // a clamped window (picture top / bottom) or a pair outside the picture: one step at a time
        ring_vstage<FV, SIDX, RS>(ring, lane, t, last);
        ring_vstage<FV, SIDX, RS>(ring, lane, t + 1, last);
        return;
    }
    constexpr Acc rnd = st.S > 0 ? (Acc)1 << (st.S > 0 ? st.S - 1 : 0) : (Acc)0;
#pragma unroll
    for (int v = 0; v < 2; v++) {
        int4 a[L + 1];
#pragma unroll
        for (int i = 0; i <= L; i++) a[i] = ring_slot<RS>(ring, k + P.dmin[SIDX] + i, lane)[(2 * (1 - par) + v) * 32];
        Acc s0[4] = {rnd, rnd, rnd, rnd}, s1[4] = {rnd, rnd, rnd, rnd};
#pragma unroll
        for (int i = 0; i < L; i++) {
            s0[0] += (Acc)st.taps[i] * (Acc)a[i].x;
            s0[1] += (Acc)st.taps[i] * (Acc)a[i].y;
            s0[2] += (Acc)st.taps[i] * (Acc)a[i].z;
            s0[3] += (Acc)st.taps[i] * (Acc)a[i].w;
            s1[0] += (Acc)st.taps[i] * (Acc)a[i + 1].x;
            s1[1] += (Acc)st.taps[i] * (Acc)a[i + 1].y;
            s1[2] += (Acc)st.taps[i] * (Acc)a[i + 1].z;
            s1[3] += (Acc)st.taps[i] * (Acc)a[i + 1].w;
        }
        int4 *tp0 = ring_slot<RS>(ring, k, lane) + (2 * par + v) * 32;
        int4 *tp1 = ring_slot<RS>(ring, k + 1, lane) + (2 * par + v) * 32;
        int4 v0 = *tp0, v1 = *tp1;
        const int d0[4] = {(int)(s0[0] >> st.S), (int)(s0[1] >> st.S), (int)(s0[2] >> st.S), (int)(s0[3] >> st.S)};
        const int d1[4] = {(int)(s1[0] >> st.S), (int)(s1[1] >> st.S), (int)(s1[2] >> st.S), (int)(s1[3] >> st.S)};
        if (add) {
            v0.x += d0[0]; v0.y += d0[1]; v0.z += d0[2]; v0.w += d0[3];
            v1.x += d1[0]; v1.y += d1[1]; v1.z += d1[2]; v1.w += d1[3];
        } else {
            v0.x -= d0[0]; v0.y -= d0[1]; v0.z -= d0[2]; v0.w -= d0[3];
            v1.x -= d1[0]; v1.y -= d1[1]; v1.z -= d1[2]; v1.w -= d1[3];
        }
        *tp0 = v0;
        *tp1 = v1;
    }
}

template <int FV, int FH, int STEPS>
__global__ void __launch_bounds__(32)
idwt_ring_kernel(LevelArgs A)
{
    using C = RingCfg<FV, FH, STEPS>;
    constexpr int NP = kWideNP, NC = kWideNC, RS = C::RS;
    constexpr int ns = get_filter(FV).nstages;
    extern __shared__ __align__(16) int4 ring_smem[];
    int4 *ring = ring_smem;
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
    X.cadd = X.rnd + (half << A.shift);
    X.vmax = 2 * half - 1;
    X.vmaxs = (X.vmax << A.shift) | ((1 << A.shift) - 1);
    X.pic = A.pic[c] + (long long)pic * A.pic_pic_stride;
    X.cw = A.crop_w[c];
    X.ch = A.crop_h[c];
    X.vec_store = ((X.cw % NC) == 0) && ((reinterpret_cast<uintptr_t>(X.pic) % (2 * NC)) == 0);
    X.out = A.out[c] + (long long)pic * A.out_pic_stride[c];
    X.W2 = 2 * A.w[c];
    X.hedge = (X.lane_lo > 0) || (X.lane_hi < 31);
    X.ld_hints = false;
    X.pol_ll = X.pol_hb = 0;
    X.pol_out = l2_policy(A.hint_out);

    const int last = X.h - 1;
    const int K0 = ty * C::RSP, K1 = min(K0 + C::RSP, X.h);
    const int ks = K0 - C::P.warm;
    const int t_end = K1 - 1 + C::P.maxlag;
    // Pairs older than ks are never staged: whatever the slots hold only reaches rows of the warm-up
    // region (that is what `warm` is for), exactly as the zero-initialised register ring of the wide kernel.
    // Blocks of two steps.  A block first needs its own two pairs (staged two blocks earlier), runs every
    // stage for both steps, emits the pairs that became final, and only then stages the pairs of the block
    // after next: their slots belong to pairs t - 12 and t - 11, which this block was the last to read.
    static_assert(C::LOOK == 4 && (C::LIVE + C::LOOK + 1) <= RS + 1, "two-step blocks: look-ahead of two blocks");
#pragma unroll
    for (int u = 0; u < C::LOOK; u += 2) {
        ring_stage_pair<RS, false>(X, ring, ks + u);
        ring_stage_pair<RS, true>(X, ring, ks + u + 1);
    }
#pragma unroll 1
    for (int t = ks; t <= t_end; t += 2) {
        asm volatile("cp.async.wait_group %0;" ::"n"(C::LOOK / 2 - 1) : "memory");  // pairs t, t + 1 have landed
        ring_vstage2<FV, 0, RS>(ring, X.lane, t, last);
        if constexpr (ns > 1) ring_vstage2<FV, (ns > 1 ? 1 : 0), RS>(ring, X.lane, t, last);
        if constexpr (ns > 2) ring_vstage2<FV, (ns > 2 ? 2 : 0), RS>(ring, X.lane, t, last);
        if constexpr (ns > 3) ring_vstage2<FV, (ns > 3 ? 3 : 0), RS>(ring, X.lane, t, last);
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int kf = t + u - C::P.maxlag;
            if (kf >= K0 && kf < K1) {
                const int4 *sl = ring_slot<RS>(ring, kf, X.lane);
#pragma unroll
                for (int par = 0; par < 2; par++) {
                    int E[NP], O[NP];
                    vec_to_array(sl[(2 * par) * 32], E);
                    vec_to_array(sl[(2 * par + 1) * 32], O);
                    if (A.final_u16) finish_row8<FH, true>(X, 2 * kf + par, E, O);
                    else finish_row8<FH, false>(X, 2 * kf + par, E, O);
                }
            }
        }
        ring_stage_pair<RS, false>(X, ring, t + C::LOOK);
        ring_stage_pair<RS, true>(X, ring, t + C::LOOK + 1);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");  // drain look-ahead copies before exit
}

template <int FV, int FH, int STEPS>
static int launch_ring_steps(LevelArgs A, int npic, cudaStream_t s)
{
    using C = RingCfg<FV, FH, STEPS>;
    int maxtiles = 0;
    for (int c = 0; c < A.ncomp; c++) {
        A.tiles_x[c] = (2 * A.w[c] + C::WOUT - 1) / C::WOUT;
        A.tiles_y[c] = (A.h[c] + C::RSP - 1) / C::RSP;
        maxtiles = max(maxtiles, A.tiles_x[c] * A.tiles_y[c]);
    }
    const cudaError_t attr = cudaFuncSetAttribute(idwt_ring_kernel<FV, FH, STEPS>,
                                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes);
    if (attr != cudaSuccess) return (int)attr;
    dim3 grid(maxtiles, A.ncomp, npic);
    idwt_ring_kernel<FV, FH, STEPS><<<grid, 32, C::kSmemBytes, s>>>(A);
    VC2B_CHECK_LAUNCH();
    return 0;
}

// strip height by the cost model of launch_wide (idwt8.cuh); resident warps are bounded by the ring
template <int FV, int FH>
static int launch_ring(const LevelArgs &A, int npic, cudaStream_t s)
{
    const double slots = 148.0 * (double)((227 << 10) / RingCfg<FV, FH, 64>::kSmemBytes);
    double best = 0;
    int pick = 64;
    const int rsp[2] = {RingCfg<FV, FH, 64>::RSP, RingCfg<FV, FH, 32>::RSP};
    const int steps[2] = {64, 32};
    for (int k = 0; k < 2; k++) {
        if (rsp[k] < 1) continue;
        double warps = 0;
        for (int c = 0; c < A.ncomp; c++) {
            const long long tx = (2 * A.w[c] + RingCfg<FV, FH, 64>::WOUT - 1) / RingCfg<FV, FH, 64>::WOUT;
            warps += (double)tx * ((A.h[c] + rsp[k] - 1) / rsp[k]) * npic;
        }
        const double waves = warps / slots;
        const double cost = ((waves > 1 ? waves : 1) + 0.5) * steps[k];
        if (best == 0 || cost < best) { best = cost; pick = steps[k]; }
    }
    if constexpr (RingCfg<FV, FH, 32>::RSP >= 1) {
        if (pick == 32) return launch_ring_steps<FV, FH, 32>(A, npic, s);
    }
    return launch_ring_steps<FV, FH, 64>(A, npic, s);
}

#ifdef VC2B_DEV_RING  // developer builds: force one instantiation for a quick compile check
template __global__ void idwt_ring_kernel<5, 6, 64>(LevelArgs);
#endif

}  // namespace vc2b
