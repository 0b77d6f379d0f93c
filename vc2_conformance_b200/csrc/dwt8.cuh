// dwt8.cuh -- wide streaming ANALYSIS kernel, the mirror of idwt8.cuh.
//
// A warp owns a strip of the level's input (interleaved domain): lane l holds 2*NP consecutive
// columns.  Per step it takes one input row pair from the cp.async-staged shared memory (first
// level: the uint16 picture with the offset removed (picture_encoding.py:264) and the last
// column/row replicated into the padding (:235-256)), pre-shifts by the filter bit shift
// (:176-181), lifts both rows horizontally with wavelet_index_ho across lanes, pushes them into the
// register ring, advances the streaming vertical lifting with wavelet_index (stages reversed,
// add/subtract swapped :203-223) and writes the finished pair de-interleaved: even row -> LL, HL;
// odd row -> LH, HH (:188-199), one 128-bit store per subband.
#pragma once

#include "wide_common.cuh"

namespace vc2b {

template <int FV, int FH, int STEPS_ = 64>
struct Ana8Cfg {
    static constexpr int NP = kWideNP;
    static constexpr StreamPlan P = make_stream_plan(FV, true);
    static constexpr StreamPlan PH = make_stream_plan(FH, true);
    static constexpr int RING = P.window <= 4 ? 4 : (P.window <= 8 ? 8 : 16);
    static constexpr int EMAX = P.entry[0] > P.entry[1] ? P.entry[0] : P.entry[1];
    static constexpr int STAGES = 4 + EMAX;
    static constexpr int STEPS = STEPS_;
    static constexpr int RSP = STEPS - P.warm - P.maxlag;
    static constexpr int HLQ = (get_filter(FH).halo + 2 * NP - 1) / (2 * NP);
    static constexpr int WOUT = 2 * NP * (32 - 2 * HLQ);  // input columns per strip
    static constexpr int hmin = PH.dmin[0] < PH.dmin[1] ? PH.dmin[0] : PH.dmin[1];
    static constexpr int hmax = PH.dmax[0] > PH.dmax[1] ? PH.dmax[0] : PH.dmax[1];
    static constexpr bool kSupported = (RING <= 8) && (-hmin <= NP) && (hmax <= NP) && (PH.ns <= 2 || FH == 6);
};

struct Ana8Ctx {
    const int32_t *in;       // int32 source, row stride W
    const uint16_t *pic;     // uint16 source (first level), row stride pic_w
    int from_u16, pic_w, pic_h, half;
    int W, h2;               // input width, number of input row pairs (H / 2)
    int jq, jqc, lane, lane_lo, lane_hi;
    bool lane_ok, hedge, pad_lane;  // pad_lane: this lane touches columns >= pic_w (edge replication)
    int shift;
    int32_t *ll, *b0, *b1, *b2;  // outputs, row stride W / 2
};

// stage pair m: rows 2m, 2m+1 of this lane's 2*NP columns (NS stage slots)
template <int NS>
__device__ __forceinline__ void ana_stage_pair(const Ana8Ctx &X, WideVec *smem, int m)
{
    constexpr int NC = kWideNC;
    WideVec *st = smem + ((((m % NS) + NS) % NS) * 4) * 32 + X.lane;
    const int r0 = min(max(2 * m, 0), 2 * X.h2 - 1), r1 = min(max(2 * m + 1, 0), 2 * X.h2 - 1);
    if (X.from_u16) {
        // 2*NP uint16 = one vector per row; rows beyond the picture replicate the last row
        const int y0 = min(r0, X.pic_h - 1), y1 = min(r1, X.pic_h - 1);
        const int vx = min(X.jqc, X.pic_w / NC - 1);  // lanes beyond the edge are fixed up at load time
        const WideVec *s0 = reinterpret_cast<const WideVec *>(X.pic + (long long)y0 * X.pic_w) + vx / 2;
        const WideVec *s1 = reinterpret_cast<const WideVec *>(X.pic + (long long)y1 * X.pic_w) + vx / 2;
        // a WideVec (16 bytes) holds 8 uint16 = one lane's row when NP = 4
        static_assert(kWideNP == 4, "uint16 staging assumes 8 samples = 16 bytes per lane");
        cp_async_vec(st, reinterpret_cast<const WideVec *>(X.pic + (long long)y0 * X.pic_w) + vx);
        cp_async_vec(st + 32, reinterpret_cast<const WideVec *>(X.pic + (long long)y1 * X.pic_w) + vx);
        (void)s0; (void)s1;
    } else {
        const WideVec *s0 = reinterpret_cast<const WideVec *>(X.in + (long long)r0 * X.W) + 2 * X.jqc;
        const WideVec *s1 = reinterpret_cast<const WideVec *>(X.in + (long long)r1 * X.W) + 2 * X.jqc;
        cp_async_vec(st, s0);
        cp_async_vec(st + 32, s0 + 1);
        cp_async_vec(st + 64, s1);
        cp_async_vec(st + 96, s1 + 1);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// At step t: the even row of pair t - EV0 and the odd row of pair t - EV1 as (E, O) column-parity
// arrays, offset removed and pre-shifted
template <int NS, int EV0, int EV1>
__device__ __forceinline__ void ana_load_pair(const Ana8Ctx &X, const WideVec *smem, int t, int (&E0)[kWideNP],
                                              int (&O0)[kWideNP], int (&E1)[kWideNP], int (&O1)[kWideNP])
{
    constexpr int NP = kWideNP, NC = kWideNC;
    asm volatile("cp.async.wait_group %0;" ::"n"(3) : "memory");
    const int me = t - EV0, mo = t - EV1;
    const WideVec *se = smem + ((((me % NS) + NS) % NS) * 4) * 32 + X.lane;
    const WideVec *so = smem + ((((mo % NS) + NS) % NS) * 4) * 32 + X.lane;
    int row0[NC], row1[NC];
    if (X.from_u16) {
        const WideVec a = se[0], b = so[32];
        const uint32_t wa[4] = {(uint32_t)a.x, (uint32_t)a.y, (uint32_t)a.z, (uint32_t)a.w};
        const uint32_t wb[4] = {(uint32_t)b.x, (uint32_t)b.y, (uint32_t)b.z, (uint32_t)b.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            row0[2 * i] = (int)(wa[i] & 0xffff) - X.half;
            row0[2 * i + 1] = (int)(wa[i] >> 16) - X.half;
            row1[2 * i] = (int)(wb[i] & 0xffff) - X.half;
            row1[2 * i + 1] = (int)(wb[i] >> 16) - X.half;
        }
        if (X.pad_lane) {  // dwt_pad_addition: columns past the picture replicate its last column
            const int r0 = min(max(2 * me, 0), 2 * X.h2 - 1), r1 = min(max(2 * mo + 1, 0), 2 * X.h2 - 1);
            const int y0 = min(r0, X.pic_h - 1), y1 = min(r1, X.pic_h - 1);
#pragma unroll
            for (int i = 0; i < NC; i++) {
                const int cx = min(max(NC * X.jq + i, 0), X.pic_w - 1);
                row0[i] = (int)X.pic[(long long)y0 * X.pic_w + cx] - X.half;
                row1[i] = (int)X.pic[(long long)y1 * X.pic_w + cx] - X.half;
            }
        }
    } else {
        int tmp[NP];
        vec_to_array(se[0], tmp);
#pragma unroll
        for (int i = 0; i < NP; i++) row0[i] = tmp[i];
        vec_to_array(se[32], tmp);
#pragma unroll
        for (int i = 0; i < NP; i++) row0[NP + i] = tmp[i];
        vec_to_array(so[64], tmp);
#pragma unroll
        for (int i = 0; i < NP; i++) row1[i] = tmp[i];
        vec_to_array(so[96], tmp);
#pragma unroll
        for (int i = 0; i < NP; i++) row1[NP + i] = tmp[i];
    }
#pragma unroll
    for (int i = 0; i < NP; i++) {
        E0[i] = row0[2 * i] << X.shift;
        O0[i] = row0[2 * i + 1] << X.shift;
        E1[i] = row1[2 * i] << X.shift;
        O1[i] = row1[2 * i + 1] << X.shift;
    }
}

template <int RING>
__device__ __forceinline__ void ana_put(PairRing<RING, kWideNC> &R, int slot_even, int slot_odd,
                                        const int (&E0)[kWideNP], const int (&O0)[kWideNP], const int (&E1)[kWideNP],
                                        const int (&O1)[kWideNP])
{
#pragma unroll
    for (int i = 0; i < kWideNP; i++) {
        R.v[0][2 * i][slot_even] = E0[i];
        R.v[0][2 * i + 1][slot_even] = O0[i];
        R.v[1][2 * i][slot_odd] = E1[i];
        R.v[1][2 * i + 1][slot_odd] = O1[i];
    }
}

__device__ __forceinline__ void ana_store(const Ana8Ctx &X, int kf, const int (&ll)[kWideNP], const int (&hl)[kWideNP],
                                          const int (&lh)[kWideNP], const int (&hh)[kWideNP])
{
    if (X.lane_ok) {
        const long long g = (long long)kf * (X.W >> 1) + kWideNP * X.jq;
        *reinterpret_cast<WideVec *>(X.ll + g) = array_to_vec(ll);
        *reinterpret_cast<WideVec *>(X.b0 + g) = array_to_vec(hl);
        *reinterpret_cast<WideVec *>(X.b1 + g) = array_to_vec(lh);
        *reinterpret_cast<WideVec *>(X.b2 + g) = array_to_vec(hh);
    }
}

template <int FV, int FH, int U>
__device__ __forceinline__ void ana_step_fast(PairRing<Ana8Cfg<FV, FH>::RING, kWideNC> &R, const Ana8Ctx &X,
                                              WideVec *smem, int t, int K0, int K1)
{
    using C = Ana8Cfg<FV, FH>;
    constexpr int RING = C::RING, M = RING - 1;
    constexpr int EV0 = C::P.entry[0], EV1 = C::P.entry[1], NS = C::STAGES;
    {
        int E0[kWideNP], O0[kWideNP], E1[kWideNP], O1[kWideNP];
        ana_load_pair<NS, EV0, EV1>(X, smem, t, E0, O0, E1, O1);
        hlift8<FH, true, Ana8Ctx>(E0, O0, X);
        hlift8<FH, true, Ana8Ctx>(E1, O1, X);
        ana_put<RING>(R, (U - EV0) & M, (U - EV1) & M, E0, O0, E1, O1);
    }
    vstep_fast<FV, U, RING, kWideNC, true>(R);
    ana_stage_pair<NS>(X, smem, t - C::EMAX + NS);
    const int kf = t - C::P.maxlag;
    if (kf >= K0 && kf < K1) {
        constexpr int sl = (U - C::P.maxlag) & M;
        int ll[kWideNP], hl[kWideNP], lh[kWideNP], hh[kWideNP];
#pragma unroll
        for (int i = 0; i < kWideNP; i++) {
            ll[i] = R.v[0][2 * i][sl];
            hl[i] = R.v[0][2 * i + 1][sl];
            lh[i] = R.v[1][2 * i][sl];
            hh[i] = R.v[1][2 * i + 1][sl];
        }
        ana_store(X, kf, ll, hl, lh, hh);
    }
}

template <int FV, int FH, int U>
__device__ __forceinline__ void ana_block_fast(PairRing<Ana8Cfg<FV, FH>::RING, kWideNC> &R, const Ana8Ctx &X,
                                               WideVec *smem, int t0, int K0, int K1)
{
    ana_step_fast<FV, FH, U>(R, X, smem, t0 + U, K0, K1);
    if constexpr (U + 1 < Ana8Cfg<FV, FH>::RING) ana_block_fast<FV, FH, U + 1>(R, X, smem, t0, K0, K1);
}

template <int FV, int FH>
__device__ __forceinline__ void ana_step_slow(PairRing<Ana8Cfg<FV, FH>::RING, kWideNC> &R, const Ana8Ctx &X,
                                              WideVec *smem, int t, int ks, int last, int K0, int K1)
{
    using C = Ana8Cfg<FV, FH>;
    constexpr int RING = C::RING, M = RING - 1;
    constexpr int EV0 = C::P.entry[0], EV1 = C::P.entry[1], NS = C::STAGES;
    int E0[kWideNP], O0[kWideNP], E1[kWideNP], O1[kWideNP];
    ana_load_pair<NS, EV0, EV1>(X, smem, t, E0, O0, E1, O1);
    hlift8<FH, true, Ana8Ctx>(E0, O0, X);
    hlift8<FH, true, Ana8Ctx>(E1, O1, X);
    const int le = (t - EV0 - ks) & M, lo = (t - EV1 - ks) & M;
#pragma unroll
    for (int i = 0; i < RING; i++) {
#pragma unroll
        for (int q = 0; q < kWideNP; q++) {
            if (i == le) { R.v[0][2 * q][i] = E0[q]; R.v[0][2 * q + 1][i] = O0[q]; }
            if (i == lo) { R.v[1][2 * q][i] = E1[q]; R.v[1][2 * q + 1][i] = O1[q]; }
        }
    }
    vstep_slow<FV, RING, kWideNC, true>(R, t, ks, last);
    ana_stage_pair<NS>(X, smem, t - C::EMAX + NS);
    const int kf = t - C::P.maxlag;
    if (kf >= K0 && kf < K1) {
        const int sl = (kf - ks) & M;
        int ll[kWideNP], hl[kWideNP], lh[kWideNP], hh[kWideNP];
#pragma unroll
        for (int i = 0; i < kWideNP; i++) {
            ll[i] = ring_get_dyn<RING>(R.v[0][2 * i], sl);
            hl[i] = ring_get_dyn<RING>(R.v[0][2 * i + 1], sl);
            lh[i] = ring_get_dyn<RING>(R.v[1][2 * i], sl);
            hh[i] = ring_get_dyn<RING>(R.v[1][2 * i + 1], sl);
        }
        ana_store(X, kf, ll, hl, lh, hh);
    }
}

template <int FV, int FH, int STEPS>
__global__ void __launch_bounds__(32)
dwt_stream8_kernel(AnaArgs A)
{
    using C = Ana8Cfg<FV, FH, STEPS>;
    constexpr int NP = kWideNP, NC = kWideNC, RING = C::RING;
    __shared__ __align__(16) WideVec smem[C::STAGES * 4 * 32];
    const int c = blockIdx.y;
    const int pic = blockIdx.z;
    const int tx = blockIdx.x % A.tiles_x[c];
    const int ty = blockIdx.x / A.tiles_x[c];
    if (ty >= A.tiles_y[c]) return;

    Ana8Ctx X;
    X.lane = threadIdx.x;
    X.W = A.W[c];
    X.h2 = A.H[c] >> 1;
    const int wq = X.W / NC;  // lanes' worth of columns
    X.jq = tx * (C::WOUT / NC) - C::HLQ + X.lane;
    X.jqc = min(max(X.jq, 0), wq - 1);
    X.lane_lo = C::HLQ - tx * (C::WOUT / NC);
    X.lane_hi = X.lane_lo + wq - 1;
    X.lane_ok = (X.lane >= C::HLQ) && (X.lane < 32 - C::HLQ) && (X.jq >= 0) && (X.jq < wq);
    X.hedge = (X.lane_lo > 0) || (X.lane_hi < 31);
    X.from_u16 = A.from_u16;
    X.in = A.from_u16 ? nullptr : A.in[c] + (long long)pic * A.in_pic_stride[c];
    X.pic = A.from_u16 ? A.pic[c] + (long long)pic * A.pic_pic_stride : nullptr;
    X.pic_w = A.pic_w[c];
    X.pic_h = A.pic_h[c];
    X.half = 1 << (A.depth[c] - 1);
    X.pad_lane = A.from_u16 && (NC * X.jqc + NC > X.pic_w);
    X.shift = A.shift;
    X.ll = A.ll[c] + (long long)pic * A.ll_pic_stride[c];
    X.b0 = A.hb[c][0] + (long long)pic * A.hb_pic_stride;
    X.b1 = A.hb[c][1] + (long long)pic * A.hb_pic_stride;
    X.b2 = A.hb[c][2] + (long long)pic * A.hb_pic_stride;

    PairRing<RING, NC> R;
#pragma unroll
    for (int i = 0; i < RING; i++)
#pragma unroll
        for (int k = 0; k < NC; k++) R.v[0][k][i] = R.v[1][k][i] = 0;
    const int last = X.h2 - 1;
    const int K0 = ty * C::RSP, K1 = min(K0 + C::RSP, X.h2);
    const int ks = K0 - C::P.warm;
#pragma unroll
    for (int u = 0; u < C::STAGES; u++) ana_stage_pair<C::STAGES>(X, smem, ks - C::EMAX + u);
    const int t_end = K1 - 1 + C::P.maxlag;
    for (int t0 = ks; t0 <= t_end; t0 += RING) {
        if (t0 >= C::P.back && t0 + RING - 1 <= last) {
            ana_block_fast<FV, FH, 0>(R, X, smem, t0, K0, K1);
        } else {
#pragma unroll 1
            for (int u = 0; u < RING; u++) ana_step_slow<FV, FH>(R, X, smem, t0 + u, ks, last, K0, K1);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// alignment / divisibility the wide analysis kernel needs
static bool ana_is_wide_ok(const AnaArgs &A, int npic)
{
    constexpr int NC = kWideNC;
    for (int c = 0; c < A.ncomp; c++) {
        if (A.W[c] % NC || (A.H[c] & 1)) return false;
        if (A.from_u16) {
            if (A.pic_w[c] % NC) return false;
            if (reinterpret_cast<uintptr_t>(A.pic[c]) & 15) return false;
            if (npic > 1 && (A.pic_pic_stride % 8)) return false;
        } else {
            if (reinterpret_cast<uintptr_t>(A.in[c]) & 15) return false;
            if (npic > 1 && (A.in_pic_stride[c] % 4)) return false;
        }
        if (reinterpret_cast<uintptr_t>(A.ll[c]) & 15) return false;
        if (npic > 1 && (A.ll_pic_stride[c] % 4)) return false;
        for (int k = 0; k < 3; k++)
            if (reinterpret_cast<uintptr_t>(A.hb[c][k]) & 15) return false;
    }
    if (npic > 1 && (A.hb_pic_stride % 4)) return false;
    return true;
}

template <int FV, int FH, int STEPS>
static int launch_ana_wide_steps(AnaArgs A, int npic, cudaStream_t s)
{
    using C = Ana8Cfg<FV, FH, STEPS>;
    int maxtiles = 0;
    for (int c = 0; c < A.ncomp; c++) {
        A.tiles_x[c] = (A.W[c] + C::WOUT - 1) / C::WOUT;
        A.tiles_y[c] = ((A.H[c] >> 1) + C::RSP - 1) / C::RSP;
        maxtiles = max(maxtiles, A.tiles_x[c] * A.tiles_y[c]);
    }
    dim3 grid(maxtiles, A.ncomp, npic);
    dwt_stream8_kernel<FV, FH, STEPS><<<grid, 32, 0, s>>>(A);
    VC2B_CHECK_LAUNCH();
    return 0;
}

template <int FV, int FH>
static int launch_ana_wide(const AnaArgs &A, int npic, cudaStream_t s)
{
    // strip height by the same cost model as the synthesis launcher (idwt8.cuh launch_wide)
    const double slots = 148.0 * 14;
    double best = 0;
    int pick = 64;
    const int rsp[3] = {Ana8Cfg<FV, FH, 64>::RSP, Ana8Cfg<FV, FH, 32>::RSP, Ana8Cfg<FV, FH, 16>::RSP};
    const int steps[3] = {64, 32, 16};
    for (int k = 0; k < 3; k++) {
        if (rsp[k] < 1) continue;
        double warps = 0;
        for (int c = 0; c < A.ncomp; c++) {
            const long long tx = (A.W[c] + Ana8Cfg<FV, FH, 64>::WOUT - 1) / Ana8Cfg<FV, FH, 64>::WOUT;
            warps += (double)tx * (((A.H[c] >> 1) + rsp[k] - 1) / rsp[k]) * npic;
        }
        const double waves = warps / slots;
        const double cost = ((waves > 1 ? waves : 1) + 0.5) * steps[k];
        if (best == 0 || cost < best) { best = cost; pick = steps[k]; }
    }
    if (pick == 64) return launch_ana_wide_steps<FV, FH, 64>(A, npic, s);
    if constexpr (Ana8Cfg<FV, FH, 32>::RSP >= 1) {
        if (pick == 32) return launch_ana_wide_steps<FV, FH, 32>(A, npic, s);
    }
    if constexpr (Ana8Cfg<FV, FH, 16>::RSP >= 1) {
        if (pick == 16) return launch_ana_wide_steps<FV, FH, 16>(A, npic, s);
    }
    return launch_ana_wide_steps<FV, FH, 64>(A, npic, s);
}

}  // namespace vc2b
