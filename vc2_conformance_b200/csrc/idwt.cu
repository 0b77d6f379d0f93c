// idwt.cu -- inverse wavelet transform (synthesis) for all seven VC-2 filters, symmetric and
// asymmetric, fused with pad removal, clip and offset on the last level.
//
// Replaces picture_decode / inverse_wavelet_transform / idwt / h_synthesis / vh_synthesis /
// oned_synthesis / filter_bit_shift / lift1-4 / idwt_pad_removal / clip_picture /
// offset_picture (pseudocode/picture_decoding.py:45-349).
//
// Design (B200): the stage is HBM-bound only if the lifting itself is nearly free, so the
// lifting lives entirely in REGISTERS.  One launch per transform level covers all three
// components and a group of pictures.  A WARP owns a strip of the level's output and walks down
// it one row pair per step: columns are lifted with wavelet_index by the streaming scheme of
// stream_lift.cuh (static register slots after unrolling), rows with wavelet_index_ho by
// exchanging neighbours with warp shuffles (:172-175).  Halo rows / lanes are recomputed by
// neighbouring strips instead of being exchanged; no block-level barrier.  The edge clamps of
// lift1-4 (:210-212, :240-242) are reproduced by clamped (dynamic-slot) steps at the top / bottom
// of the picture and by refreshing out-of-range lanes from the boundary lane before every
// horizontal stage -- only in strips that touch an edge.  The last level applies the rounding
// shift (:177-182, shift from wavelet_index_ho :198-201), crops (:282), clips (:327), offsets
// (:301) and stores packed uint16.  Kernels: idwt_stream8_kernel (idwt8.cuh: 8 columns per lane,
// cp.async staging; the fast path), idwt_stream_kernel (2 columns per lane; any width / filter
// pair), idwt_rows_kernel (horizontal-only levels).
#include <stdlib.h>

#include "stream_lift.cuh"

namespace vc2b {

constexpr int kIdwtWarps = 4;  // warps per CTA (independent of each other)

struct LevelArgs {
    // input LL band and the level's high-pass bands, per component
    const int32_t *ll[3];
    long long ll_pic_stride[3];     // elements between consecutive pictures
    const int32_t *hb[3][3];        // HL, LH, HH (or H in slot 0 for horizontal-only levels)
    long long hb_pic_stride;        // elements (int32, or int16 when hb16)
    int32_t hb16;                   // 1: hb[][] address int16 coefficients (vc2b_unpack16's plane; wide kernel only)
    int32_t ll16, out16;            // 1: ll[] / out[] address int16 samples (vc2b_idwt16's intermediate bands)
    int32_t *status;                // out16: per-picture status words (4 per picture) for VC2B_ST_COEFF16_RANGE
    int32_t w[3], h[3];             // input LL dimensions
    // output
    int32_t *out[3];                // int32 output (intermediate LL or raw final), row stride = 2w
    long long out_pic_stride[3];
    uint16_t *pic[3];               // final picture planes
    long long pic_pic_stride;
    int32_t crop_w[3], crop_h[3], depth[3];
    int32_t final_u16;              // 1: write pic (crop+clip+offset); 0: write out (int32)
    int32_t shift;                  // filter_bit_shift(wavelet_index_ho)
    int32_t tiles_x[3], tiles_y[3];
    int32_t ncomp;
    int32_t hint_loads;                  // 1: the cp.async loads carry hint_ll / hint_hb
    int32_t hint_ll, hint_hb, hint_out;  // L2Hint of the LL input, the high-pass inputs and the output (wide kernel)
};

// ---- horizontal lifting across lanes ---------------------------------------------------------

// One stage on one row: lane l holds E = sample 2j, O = sample 2j+1 (j = j0 + l).
template <int TYPE, int S, int L, int D, int T0, int T1, int T2, int T3, int T4, int T5, int T6, int T7>
__device__ __forceinline__ void hstage(int &E, int &O, int lane, bool edge, int lane_lo, int lane_hi)
{
    constexpr int taps[8] = {T0, T1, T2, T3, T4, T5, T6, T7};
    constexpr bool kEven = (TYPE == 1 || TYPE == 2);
    constexpr bool kAdd = (TYPE == 1 || TYPE == 3);
    int src = kEven ? O : E;
    if (edge) {  // replicate the boundary column of the source parity into out-of-range lanes
        const int lo = __shfl_sync(0xffffffffu, src, lane_lo & 31);
        const int hi = __shfl_sync(0xffffffffu, src, lane_hi & 31);
        if (lane < lane_lo) src = lo;
        if (lane > lane_hi) src = hi;
    }
    typename std::conditional<(S >= kWideAccS), long long, int>::type sum = 0;
#pragma unroll
    for (int i = D; i < L + D; i++) {
        // even update reads odd sample index n+i-1; odd update reads even sample index n+i
        const int delta = kEven ? i - 1 : i;
        const int v = (delta == 0) ? src : __shfl_sync(0xffffffffu, src, (lane + delta) & 31);
        sum += (decltype(sum))taps[i - D] * v;
    }
    if (S > 0) sum += (decltype(sum))1 << (S > 0 ? S - 1 : 0);
    sum >>= S;
    if (kEven) { if (kAdd) E += (int)sum; else E -= (int)sum; }
    else { if (kAdd) O += (int)sum; else O -= (int)sum; }
}

template <int F, int SIDX>
__device__ __forceinline__ void hrun_stage(int &E, int &O, int lane, bool edge, int lane_lo, int lane_hi)
{
    constexpr FilterDef f = get_filter(F);
    constexpr StageDef s = f.st[SIDX];
    hstage<s.type, s.S, s.L, s.D, s.taps[0], s.taps[1], s.taps[2], s.taps[3], s.taps[4], s.taps[5], s.taps[6],
           s.taps[7]>(E, O, lane, edge, lane_lo, lane_hi);
}

template <int F>
__device__ __forceinline__ void hlift(int &E, int &O, int lane, bool edge, int lane_lo, int lane_hi)
{
    constexpr int ns = get_filter(F).nstages;
    hrun_stage<F, 0>(E, O, lane, edge, lane_lo, lane_hi);
    if constexpr (ns > 1) hrun_stage<F, (ns > 1 ? 1 : 0)>(E, O, lane, edge, lane_lo, lane_hi);
    if constexpr (ns > 2) hrun_stage<F, (ns > 2 ? 2 : 0)>(E, O, lane, edge, lane_lo, lane_hi);
    if constexpr (ns > 3) hrun_stage<F, (ns > 3 ? 3 : 0)>(E, O, lane, edge, lane_lo, lane_hi);
}

// ---- horizontal-only level kernel (h_synthesis, picture_decoding.py:134) --------------------------

template <int FH>
struct RowCfg {
    static constexpr int HL = get_filter(FH).halo / 2;  // halo lanes left / right
    static constexpr int R = 16;                         // rows per warp tile
    static constexpr int WOUT = 2 * (32 - 2 * HL);       // output columns per warp tile
};

template <int FH>
__global__ void __launch_bounds__(kIdwtWarps * 32)
idwt_rows_kernel(LevelArgs A)
{
    using C = RowCfg<FH>;
    constexpr int HL = C::HL, R = C::R, WOUT = C::WOUT;

    const int lane = threadIdx.x & 31;
    const int c = blockIdx.y;
    const int pic = blockIdx.z;
    const int tile = blockIdx.x * kIdwtWarps + (threadIdx.x >> 5);
    const int tx = tile % A.tiles_x[c];
    const int ty = tile / A.tiles_x[c];
    if (ty >= A.tiles_y[c]) return;

    const int w = A.w[c], h = A.h[c];
    const int W2 = 2 * w;
    const int X0 = tx * WOUT, Y0 = ty * R;
    const int j = (X0 >> 1) - HL + lane;          // column pair owned by this lane
    const int jc = min(max(j, 0), w - 1);
    const int lane_lo = HL - (X0 >> 1);           // lane of column pair 0 (<= 0 away from the edge)
    const int lane_hi = lane_lo + w - 1;          // lane of the last column pair
    const bool hedge = (lane_lo > 0) || (lane_hi < 31);

    const int32_t *ll = A.ll[c] + (long long)pic * A.ll_pic_stride[c];
    const int32_t *b0 = A.hb[c][0] + (long long)pic * A.hb_pic_stride;
    const int shift = A.shift;
    const int rnd = shift > 0 ? (1 << (shift - 1)) : 0;
    const bool lane_ok = (lane >= HL) && (lane < 32 - HL) && (j >= 0) && (j < w);
    const int cw = A.crop_w[c], ch = A.crop_h[c];
    const int half = 1 << (A.depth[c] - 1);
    uint16_t *pdst = A.final_u16 ? A.pic[c] + (long long)pic * A.pic_pic_stride : nullptr;
    int32_t *idst = A.final_u16 ? nullptr : A.out[c] + (long long)pic * A.out_pic_stride[c];
#pragma unroll 4
    for (int i = 0; i < R; i++) {
        const int gy = Y0 + i;
        const int sy = min(gy, h - 1);
        const long long g = (long long)sy * w + jc;
        int E = __ldg(ll + g), O = __ldg(b0 + g);   // L -> even column, H -> odd column (:140-143)
        hlift<FH>(E, O, lane, hedge, lane_lo, lane_hi);
        if (!lane_ok || gy >= h) continue;
        if (A.final_u16) {
            if (gy < ch) {
                int v0 = (E + rnd) >> shift, v1 = (O + rnd) >> shift;
                v0 = min(max(v0, -half), half - 1) + half;
                v1 = min(max(v1, -half), half - 1) + half;
                const long long idx = (long long)gy * cw + 2 * j;
                if (2 * j < cw) pdst[idx] = (uint16_t)v0;
                if (2 * j + 1 < cw) pdst[idx + 1] = (uint16_t)v1;
            }
        } else {
            *reinterpret_cast<int2 *>(idst + (long long)gy * W2 + 2 * j) =
                make_int2((E + rnd) >> shift, (O + rnd) >> shift);
        }
    }
}

// ---- streaming 2-D level kernel ------------------------------------------------------------------
//
// One warp (one CTA) owns a strip: WOUT output columns wide, RSP row pairs tall.  It walks down
// the strip one row pair per step: prefetch the raw pair P steps ahead (LL/HL -> even row,
// LH/HH -> odd row), advance every vertical lifting stage by one pair (stream_lift.cuh), then
// lift the finished pair's two rows horizontally across lanes, shift, clip, offset and store.
// No vertical halo is recomputed inside a strip (only `warm` pairs at its top).

template <int FV, int FH>
struct StripCfg {
    static constexpr StreamPlan P = make_stream_plan(FV);
    static constexpr int RING = P.ring;
    static constexpr int STEPS = 64;                           // steps per strip (multiple of RING)
    static constexpr int RSP = STEPS - P.warm - P.maxlag;      // output row pairs per strip
    static constexpr int HL = get_filter(FH).halo / 2;         // halo lanes left / right
    static constexpr int WOUT = 2 * (32 - 2 * HL);
};

struct StripCtx {
    const int32_t *ll, *b0, *b1, *b2;
    int w, h, jc, j, lane, lane_lo, lane_hi;
    bool lane_ok, hedge;
    int shift, rnd;
    // final output
    uint16_t *pic;
    int cw, ch, half;
    int32_t *out;
    int W2;
};

template <int RING>
__device__ __forceinline__ void load_pair(const StripCtx &X, int m, int &ee, int &eo, int &oe, int &oo)
{
    const int sy = min(max(m, 0), X.h - 1);
    const int idx = sy * X.w + X.jc;
    ee = __ldg(X.ll + idx);
    eo = __ldg(X.b0 + idx);
    oe = __ldg(X.b1 + idx);
    oo = __ldg(X.b2 + idx);
}

template <int FH, bool FINAL>
__device__ __forceinline__ void finish_row(const StripCtx &X, int gy, int E, int O)
{
    hlift<FH>(E, O, X.lane, X.hedge, X.lane_lo, X.lane_hi);
    if (FINAL) {
        if (X.lane_ok && gy < X.ch) {
            int v0 = (E + X.rnd) >> X.shift, v1 = (O + X.rnd) >> X.shift;
            v0 = min(max(v0, -X.half), X.half - 1) + X.half;
            v1 = min(max(v1, -X.half), X.half - 1) + X.half;
            uint16_t *d = X.pic + (long long)gy * X.cw + 2 * X.j;
            if (2 * X.j + 1 < X.cw && !(reinterpret_cast<uintptr_t>(d) & 3)) {
                *reinterpret_cast<uint32_t *>(d) = (uint32_t)v0 | ((uint32_t)v1 << 16);
            } else {
                if (2 * X.j < X.cw) d[0] = (uint16_t)v0;
                if (2 * X.j + 1 < X.cw) d[1] = (uint16_t)v1;
            }
        }
    } else {
        if (X.lane_ok) {
            int2 v = make_int2((E + X.rnd) >> X.shift, (O + X.rnd) >> X.shift);
            *reinterpret_cast<int2 *>(X.out + (long long)gy * X.W2 + 2 * X.j) = v;
        }
    }
}

template <int FV, int FH, bool FINAL, int U>
__device__ __forceinline__ void strip_step_fast(PairRing<StripCfg<FV, FH>::RING> &R, const StripCtx &X, int t,
                                                int K0, int K1)
{
    using C = StripCfg<FV, FH>;
    constexpr int RING = C::RING, M = RING - 1;
    constexpr int PF = C::P.prefetch;
    load_pair<RING>(X, t + PF, R.v[0][0][(U + PF) & M], R.v[0][1][(U + PF) & M], R.v[1][0][(U + PF) & M],
                    R.v[1][1][(U + PF) & M]);
    vstep_fast<FV, U, RING, 2>(R);
    const int kf = t - C::P.maxlag;
    if (kf >= K0 && kf < K1) {
        constexpr int sl = (U - C::P.maxlag) & M;
        finish_row<FH, FINAL>(X, 2 * kf, R.v[0][0][sl], R.v[0][1][sl]);
        finish_row<FH, FINAL>(X, 2 * kf + 1, R.v[1][0][sl], R.v[1][1][sl]);
    }
}

template <int FV, int FH, bool FINAL, int U>
__device__ __forceinline__ void strip_block_fast(PairRing<StripCfg<FV, FH>::RING> &R, const StripCtx &X, int t0,
                                                 int K0, int K1)
{
    strip_step_fast<FV, FH, FINAL, U>(R, X, t0 + U, K0, K1);
    if constexpr (U + 1 < StripCfg<FV, FH>::RING) strip_block_fast<FV, FH, FINAL, U + 1>(R, X, t0, K0, K1);
}

template <int FV, int FH, bool FINAL>
__device__ __forceinline__ void strip_step_slow(PairRing<StripCfg<FV, FH>::RING> &R, const StripCtx &X, int t,
                                                int ks, int last, int K0, int K1)
{
    using C = StripCfg<FV, FH>;
    constexpr int RING = C::RING, M = RING - 1;
    int ee, eo, oe, oo;
    load_pair<RING>(X, t + C::P.prefetch, ee, eo, oe, oo);
    const int ls = (t + C::P.prefetch - ks) & M;
#pragma unroll
    for (int i = 0; i < RING; i++)
        if (i == ls) { R.v[0][0][i] = ee; R.v[0][1][i] = eo; R.v[1][0][i] = oe; R.v[1][1][i] = oo; }
    vstep_slow<FV, RING, 2>(R, t, ks, last);
    const int kf = t - C::P.maxlag;
    if (kf >= K0 && kf < K1) {
        const int sl = (kf - ks) & M;
        finish_row<FH, FINAL>(X, 2 * kf, ring_get_dyn<RING>(R.v[0][0], sl), ring_get_dyn<RING>(R.v[0][1], sl));
        finish_row<FH, FINAL>(X, 2 * kf + 1, ring_get_dyn<RING>(R.v[1][0], sl), ring_get_dyn<RING>(R.v[1][1], sl));
    }
}

template <int FV, int FH, bool FINAL>
__device__ __forceinline__ void run_strip(const StripCtx &X, int ty)
{
    using C = StripCfg<FV, FH>;
    constexpr int RING = C::RING;
    PairRing<RING> R;
#pragma unroll
    for (int i = 0; i < RING; i++) R.v[0][0][i] = R.v[0][1][i] = R.v[1][0][i] = R.v[1][1][i] = 0;
    const int last = X.h - 1;
    const int K0 = ty * C::RSP, K1 = min(K0 + C::RSP, X.h);
    const int ks = K0 - C::P.warm;
#pragma unroll
    for (int u = 0; u < C::P.prefetch; u++)
        load_pair<RING>(X, ks + u, R.v[0][0][u], R.v[0][1][u], R.v[1][0][u], R.v[1][1][u]);
    const int t_end = K1 - 1 + C::P.maxlag;  // last step that finishes a pair of this strip
    for (int t0 = ks; t0 <= t_end; t0 += RING) {
        if (t0 >= C::P.back && t0 + RING - 1 <= last) {
            strip_block_fast<FV, FH, FINAL, 0>(R, X, t0, K0, K1);
        } else {
#pragma unroll 1
            for (int u = 0; u < RING; u++) strip_step_slow<FV, FH, FINAL>(R, X, t0 + u, ks, last, K0, K1);
        }
    }
}

template <int FV, int FH>
__global__ void __launch_bounds__(32)
idwt_stream_kernel(LevelArgs A)
{
    using C = StripCfg<FV, FH>;
    const int c = blockIdx.y;
    const int pic = blockIdx.z;
    const int tx = blockIdx.x % A.tiles_x[c];
    const int ty = blockIdx.x / A.tiles_x[c];
    if (ty >= A.tiles_y[c]) return;

    StripCtx X;
    X.lane = threadIdx.x;
    X.w = A.w[c];
    X.h = A.h[c];
    const int X0 = tx * C::WOUT;
    X.j = (X0 >> 1) - C::HL + X.lane;
    X.jc = min(max(X.j, 0), X.w - 1);
    X.lane_lo = C::HL - (X0 >> 1);
    X.lane_hi = X.lane_lo + X.w - 1;
    X.lane_ok = (X.lane >= C::HL) && (X.lane < 32 - C::HL) && (X.j >= 0) && (X.j < X.w);
    X.ll = A.ll[c] + (long long)pic * A.ll_pic_stride[c];
    X.b0 = A.hb[c][0] + (long long)pic * A.hb_pic_stride;
    X.b1 = A.hb[c][1] + (long long)pic * A.hb_pic_stride;
    X.b2 = A.hb[c][2] + (long long)pic * A.hb_pic_stride;
    X.shift = A.shift;
    X.rnd = A.shift > 0 ? (1 << (A.shift - 1)) : 0;
    X.pic = A.pic[c] + (long long)pic * A.pic_pic_stride;
    X.cw = A.crop_w[c];
    X.ch = A.crop_h[c];
    X.half = 1 << (A.depth[c] - 1);
    X.out = A.out[c] + (long long)pic * A.out_pic_stride[c];
    X.W2 = 2 * X.w;
    X.hedge = (X.lane_lo > 0) || (X.lane_hi < 31);
    if (A.final_u16) run_strip<FV, FH, true>(X, ty); else run_strip<FV, FH, false>(X, ty);
}

}  // namespace vc2b

#include "idwt8.cuh"
#include "idwt_ring.cuh"

namespace vc2b {

// No transform at all (dwt_depth == dwt_depth_ho == 0): picture = level-0 band.
__global__ void idwt_copy_kernel(LevelArgs A)
{
    const int c = blockIdx.y, pic = blockIdx.z;
    const int w = A.w[c], h = A.h[c];
    const int32_t *ll = A.ll[c] + (long long)pic * A.ll_pic_stride[c];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)w * h;
         i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / w), x = (int)(i - (long long)y * w);
        int v = ll[i];
        if (A.final_u16) {
            if (x < A.crop_w[c] && y < A.crop_h[c]) {
                const int half = 1 << (A.depth[c] - 1);
                v = min(max(v, -half), half - 1) + half;
                A.pic[c][(long long)pic * A.pic_pic_stride + (long long)y * A.crop_w[c] + x] = (uint16_t)v;
            }
        } else {
            A.out[c][(long long)pic * A.out_pic_stride[c] + i] = v;
        }
    }
}

template <int FV, int FH, bool VERT>
static int launch_level_t(const LevelArgs &A_in, int npic, cudaStream_t s)
{
    LevelArgs A = A_in;
    int maxtiles = 0;
    if constexpr (VERT && FV == 5 && RingCfg<FV, FH>::kSupported) {  // Fidelity: vertical window in shared memory
        if (level_is_wide_ok(A, npic)) return launch_ring<FV, FH>(A, npic, s);
    }
    if constexpr (VERT && FV != 5 && Strip8Cfg<FV, FH>::kSupported && FV == FH) {
        if (level_is_wide_ok(A, npic)) return launch_wide<FV, FH>(A, npic, s);
    }
    if (A.hb16) return VC2B_E_UNSUPPORTED;  // only the wide kernel reads int16 high-pass bands
    if constexpr (VERT) {
        using C = StripCfg<FV, FH>;
        for (int c = 0; c < A.ncomp; c++) {
            A.tiles_x[c] = (2 * A.w[c] + C::WOUT - 1) / C::WOUT;
            A.tiles_y[c] = (A.h[c] + C::RSP - 1) / C::RSP;
            maxtiles = max(maxtiles, A.tiles_x[c] * A.tiles_y[c]);
        }
        dim3 grid(maxtiles, A.ncomp, npic);
        idwt_stream_kernel<FV, FH><<<grid, 32, 0, s>>>(A);
    } else {
        using C = RowCfg<FH>;
        for (int c = 0; c < A.ncomp; c++) {
            A.tiles_x[c] = (2 * A.w[c] + C::WOUT - 1) / C::WOUT;
            A.tiles_y[c] = (A.h[c] + C::R - 1) / C::R;
            maxtiles = max(maxtiles, A.tiles_x[c] * A.tiles_y[c]);
        }
        dim3 grid((maxtiles + kIdwtWarps - 1) / kIdwtWarps, A.ncomp, npic);
        idwt_rows_kernel<FH><<<grid, kIdwtWarps * 32, 0, s>>>(A);
    }
    VC2B_CHECK_LAUNCH();
    return 0;
}

template <int FV, bool VERT>
static int launch_level_h(int fh, const LevelArgs &A, int npic, cudaStream_t s)
{
#ifdef VC2B_DEV_BUILD
    return VC2B_E_UNSUPPORTED;
#else
    switch (fh) {
    case 0: return launch_level_t<FV, 0, VERT>(A, npic, s);
    case 1: return launch_level_t<FV, 1, VERT>(A, npic, s);
    case 2: return launch_level_t<FV, 2, VERT>(A, npic, s);
    case 3: return launch_level_t<FV, 3, VERT>(A, npic, s);
    case 5: return launch_level_t<FV, 5, VERT>(A, npic, s);
    default: return launch_level_t<FV, 6, VERT>(A, npic, s);
    }
#endif
}

static int launch_level(int fv, int fh, bool vert, const LevelArgs &A, int npic, cudaStream_t s)
{
    // Haar with and without shift share their lifting stages (the shift is a runtime argument)
    if (fv == 4) fv = 3;
    if (fh == 4) fh = 3;
#ifdef VC2B_DEV_BUILD  // quick developer builds: Deslauriers-Dubuc (9,7) only
    if (fh != 0 || (vert && fv != 0)) return VC2B_E_UNSUPPORTED;
    return vert ? launch_level_t<0, 0, true>(A, npic, s) : launch_level_t<0, 0, false>(A, npic, s);
#endif
    if (!vert) return launch_level_h<0, false>(fh, A, npic, s);
    switch (fv) {
    case 0: return launch_level_h<0, true>(fh, A, npic, s);
    case 1: return launch_level_h<1, true>(fh, A, npic, s);
    case 2: return launch_level_h<2, true>(fh, A, npic, s);
    case 3: return launch_level_h<3, true>(fh, A, npic, s);
    case 5: return launch_level_h<5, true>(fh, A, npic, s);
    default: return launch_level_h<6, true>(fh, A, npic, s);
    }
}

// first high-pass band index of level n (1-based)
static int level_first_band(const DevParams &P, int n)
{
    return (n <= P.dwt_depth_ho) ? n : P.dwt_depth_ho + 1 + 3 * (n - P.dwt_depth_ho - 1);
}

// Scratch layout per picture: two ping-pong regions holding the int32 output of the
// non-final levels of all (selected) components.
struct ScratchPlan {
    long long size_a, size_b;  // elements
    long long comp_off_a[3], comp_off_b[3];
};

static void plan_scratch(const DevParams &P, int c0, int c1, ScratchPlan *sp)
{
    const int top = P.dwt_depth + P.dwt_depth_ho;
    sp->size_a = sp->size_b = 0;
    for (int c = c0; c < c1; c++) {
        // output of level n (1-based) has the dimensions of the level n+1 input LL band
        long long a = 0, b = 0;
        for (int n = top - 1; n >= 1; n--) {
            const Band &B = P.band[c][level_first_band(P, n + 1)];  // same dims as level n+1's LL
            long long sz = (long long)B.w * B.h;
            if (((top - 1) - n) % 2 == 0) a = max(a, sz); else b = max(b, sz);
        }
        sp->comp_off_a[c] = sp->size_a;
        sp->comp_off_b[c] = sp->size_b;
        // every region a multiple of 8 elements: 16-byte aligned vector stores for int32 and for int16 bands
        // (vc2b_idwt16), whatever the picture's position in the group
        sp->size_a += (a + 7) & ~7LL;
        sp->size_b += (b + 7) & ~7LL;
    }
}

// Runs all levels for components [c0, c1) of `npic` pictures, one launch per level over the whole group.
// The final level's stores carry an L2 evict_first policy (the picture is written once and never read
// back by this library: +3 % on B200).  `coeffs16`: when not null, the high-pass bands are read from this
// int16 plane (same element offsets as `coeffs`, which then only supplies the level-0 band).
//
// Measured and removed: keeping the band between the last two levels in L2 by running them
// over L2-sized sub-groups of pictures, a ticket-ordered two-level launch, and a producer/consumer kernel
// fusing the last two levels through a shared-memory ring -- all bit-exact, all slower than one launch per
// level on B200.  What does cut the DRAM traffic of this stage is the int16 coefficient plane.
static int run_idwt(const DevParams &P, int c0, int c1, const int32_t *coeffs, const int16_t *coeffs16,
                    long long coeff_pic_stride, int npic, uint16_t *pictures, int32_t *raw_out, int32_t *scratch,
                    const ScratchPlan &sp, int32_t *status, cudaStream_t s)
{
    // with the int16 coefficient plane the intermediate bands are int16 too (same element offsets in the
    // scratch buffer); a level that produces a value outside int16 flags the picture in `status`
    int16_t *scratch16 = reinterpret_cast<int16_t *>(scratch);
    const bool s16 = coeffs16 != nullptr;
    const int top = P.dwt_depth + P.dwt_depth_ho;
    const int ncomp = c1 - c0;
    const long long scratch_pic = sp.size_a + sp.size_b;
    LevelArgs A;
    memset(&A, 0, sizeof(A));
    A.ncomp = ncomp;
    A.hb_pic_stride = coeff_pic_stride;
    A.hb16 = s16 ? 1 : 0;
    A.status = status;
    A.pic_pic_stride = P.pic_samples;
    A.shift = get_filter(P.wavelet_index_ho).shift;
    for (int i = 0; i < ncomp; i++) {
        const int c = c0 + i;
        A.crop_w[i] = P.width[c];
        A.crop_h[i] = P.height[c];
        A.depth[i] = P.depth[c];
    }
    if (top == 0) {
        for (int i = 0; i < ncomp; i++) {
            const int c = c0 + i;
            if (pictures) A.pic[i] = pictures + P.pic_off[c];
            A.ll[i] = coeffs + P.comp_off[c] + P.band[c][0].off;
            A.ll_pic_stride[i] = coeff_pic_stride;
            A.w[i] = P.band[c][0].w;
            A.h[i] = P.band[c][0].h;
            A.out[i] = raw_out;
            A.out_pic_stride[i] = 0;
        }
        A.final_u16 = pictures ? 1 : 0;
        dim3 grid(148 * 4, ncomp, npic);
        idwt_copy_kernel<<<grid, 256, 0, s>>>(A);
        VC2B_CHECK_LAUNCH();
        return 0;
    }
    for (int n = 1; n <= top; n++) {
        // Level n: its LL input is what level n-1 wrote (the coefficients for n == 1); ping-pong: level n
        // writes scratch region ((top-1)-n)%2 (0 = a)
        const bool vert = n > P.dwt_depth_ho;
        const bool last = n == top;
        const int fb = level_first_band(P, n);
        const int wsel = ((top - 1) - n) & 1;
        for (int i = 0; i < ncomp; i++) {
            const int c = c0 + i;
            const Band &B = P.band[c][fb];
            A.w[i] = B.w;
            A.h[i] = B.h;
            for (int k = 0; k < (vert ? 3 : 1); k++) {
                const long long off = (long long)P.comp_off[c] + P.band[c][fb + k].off;
                A.hb[i][k] = coeffs16 ? reinterpret_cast<const int32_t *>(coeffs16 + off) : coeffs + off;
            }
            A.pic[i] = pictures ? pictures + P.pic_off[c] : nullptr;
            if (n == 1) {
                A.ll[i] = coeffs + P.comp_off[c] + P.band[c][0].off;
                A.ll_pic_stride[i] = coeff_pic_stride;
            } else {
                const int rsel = ((top - 1) - (n - 1)) & 1;
                const long long off = rsel == 0 ? sp.comp_off_a[c] : sp.size_a + sp.comp_off_b[c];
                A.ll[i] = s16 ? reinterpret_cast<const int32_t *>(scratch16 + off) : scratch + off;
                A.ll_pic_stride[i] = scratch_pic;
            }
            if (!last) {
                const long long off = wsel == 0 ? sp.comp_off_a[c] : sp.size_a + sp.comp_off_b[c];
                A.out[i] = s16 ? reinterpret_cast<int32_t *>(scratch16 + off) : scratch + off;
                A.out_pic_stride[i] = scratch_pic;
            } else {
                A.out[i] = raw_out;
                A.out_pic_stride[i] = 0;
            }
        }
        A.ll16 = (s16 && n > 1) ? 1 : 0;
        A.out16 = (s16 && !last) ? 1 : 0;
        A.final_u16 = (last && pictures) ? 1 : 0;
        A.hint_loads = 0;
        A.hint_hb = A.hint_ll = kL2Normal;
        A.hint_out = (last && npic > 1) ? kL2EvictFirst : kL2Normal;
        const int st = launch_level(P.wavelet_index, P.wavelet_index_ho, vert, A, npic, s);
        if (st) return st;
    }
    return 0;
}

// Whether every level of this parameter set runs on the wide kernel with vector-aligned bands: the
// condition for the int16 coefficient plane (vc2b_unpack16 / vc2b_idwt16).
static bool coeff16_transform_ok(const DevParams &P)
{
    if (P.dwt_depth_ho != 0 || P.dwt_depth < 1) return false;
    int f = P.wavelet_index, fh = P.wavelet_index_ho;
    if (f == 4) f = 3;
    if (fh == 4) fh = 3;
    if (f != fh) return false;
    bool wide = false;
    switch (f) {
    case 0: wide = Strip8Cfg<0, 0>::kSupported; break;
#ifndef VC2B_DEV_BUILD
    case 1: wide = Strip8Cfg<1, 1>::kSupported; break;
    case 2: wide = Strip8Cfg<2, 2>::kSupported; break;
    case 3: wide = Strip8Cfg<3, 3>::kSupported; break;
    case 6: wide = Strip8Cfg<6, 6>::kSupported; break;
#endif
    default: wide = false;
    }
    if (!wide) return false;
    if (P.pic_coeffs % kWideNP) return false;
    for (int c = 0; c < 3; c++) {
        if (P.comp_off[c] % kWideNP) return false;
        for (int b = 0; b < P.nb; b++)
            if (P.band[c][b].w % kWideNP || P.band[c][b].off % kWideNP) return false;
    }
    return true;
}

}  // namespace vc2b

using namespace vc2b;

extern "C" {

int64_t vc2b_transform_scratch_bytes(const vc2b_params *p)
{
    DevParams P;
    int st = make_dev_params(p, &P);
    if (st) return st;
    ScratchPlan sp;
    plan_scratch(P, 0, 3, &sp);
    long long n = sp.size_a + sp.size_b;
    // regions are 16-byte aligned so that paired stores stay aligned
    return ((n > 0 ? n : 1) * (long long)sizeof(int32_t) + 15) & ~15LL;
}

int vc2b_idwt(const vc2b_params *p, const int32_t *d_coeffs, int npictures, uint16_t *d_pictures,
              void *d_scratch, int64_t scratch_bytes, void *stream)
{
    DevParams P;
    int st = make_dev_params(p, &P);
    if (st) return st;
    if (npictures <= 0) return 0;
    ScratchPlan sp;
    plan_scratch(P, 0, 3, &sp);
    const long long per_pic = (sp.size_a + sp.size_b) * (long long)sizeof(int32_t);
    int group = per_pic > 0 ? (int)(scratch_bytes / per_pic) : npictures;
    if (group < 1) return VC2B_E_SCRATCH_TOO_SMALL;
    if (group > npictures) group = npictures;
    for (int i = 0; i < npictures; i += group) {
        const int n = (npictures - i < group) ? npictures - i : group;
        st = run_idwt(P, 0, 3, d_coeffs + (long long)i * P.pic_coeffs, nullptr, P.pic_coeffs, n,
                      d_pictures + (long long)i * P.pic_samples, nullptr, (int32_t *)d_scratch, sp, nullptr,
                      (cudaStream_t)stream);
        if (st) return st;
    }
    return 0;
}

int vc2b_coeff16_supported(const vc2b_params *p)
{
    DevParams P;
    if (make_dev_params(p, &P)) return 0;
    return coeff16_transform_ok(P) && lanes_unpack_ok(P) ? 1 : 0;
}

int vc2b_idwt16(const vc2b_params *p, const int32_t *d_coeffs, const int16_t *d_coeffs16, int npictures,
                uint16_t *d_pictures, void *d_scratch, int64_t scratch_bytes, int32_t *d_status, void *stream)
{
    DevParams P;
    int st = make_dev_params(p, &P);
    if (st) return st;
    if (!coeff16_transform_ok(P)) return VC2B_E_UNSUPPORTED;
    if (!d_coeffs16 || !d_status) return VC2B_E_BAD_PARAMS;
    if ((reinterpret_cast<uintptr_t>(d_coeffs) & 15) || (reinterpret_cast<uintptr_t>(d_coeffs16) & 15) ||
        (reinterpret_cast<uintptr_t>(d_scratch) & 15))
        return VC2B_E_BAD_PARAMS;
    if (npictures <= 0) return 0;
    ScratchPlan sp;
    plan_scratch(P, 0, 3, &sp);
    const long long per_pic = (sp.size_a + sp.size_b) * (long long)sizeof(int16_t);  // int16 intermediate bands
    int group = per_pic > 0 ? (int)(scratch_bytes / per_pic) : npictures;
    if (group < 1) return VC2B_E_SCRATCH_TOO_SMALL;
    if (group > npictures) group = npictures;
    for (int i = 0; i < npictures; i += group) {
        const int n = (npictures - i < group) ? npictures - i : group;
        st = run_idwt(P, 0, 3, d_coeffs + (long long)i * P.pic_coeffs, d_coeffs16 + (long long)i * P.pic_coeffs,
                      P.pic_coeffs, n, d_pictures + (long long)i * P.pic_samples, nullptr, (int32_t *)d_scratch, sp,
                      d_status + 4LL * i, (cudaStream_t)stream);
        if (st) return st;
    }
    return 0;
}

int vc2b_idwt_raw(const vc2b_params *p, int comp, const int32_t *d_coeffs, int32_t *d_out, void *d_scratch,
                  int64_t scratch_bytes, void *stream)
{
    DevParams P;
    int st = make_dev_params(p, &P);
    if (st) return st;
    if (comp < 0 || comp > 2) return VC2B_E_BAD_PARAMS;
    ScratchPlan sp;
    plan_scratch(P, comp, comp + 1, &sp);
    if ((sp.size_a + sp.size_b) * (long long)sizeof(int32_t) > scratch_bytes) return VC2B_E_SCRATCH_TOO_SMALL;
    // d_coeffs points at the component's own coefficients
    return run_idwt(P, comp, comp + 1, d_coeffs - P.comp_off[comp], nullptr, 0, 1, nullptr, d_out,
                    (int32_t *)d_scratch, sp, nullptr, (cudaStream_t)stream);
}

}  // extern "C"
