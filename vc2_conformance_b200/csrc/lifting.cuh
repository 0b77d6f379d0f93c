// lifting.cuh -- the seven VC-2 lifting filters as compile-time tables and the
// tile-level lifting pass shared by the synthesis (idwt.cu) and analysis
// (dwt.cu) kernels.
//
// Tables: vc2_data_tables.LIFTING_FILTERS (un-vendored dependency of the
// reference, SMPTE ST 2042-1 Tables 15.1-15.7), consumed by oned_synthesis
// (pseudocode/picture_decoding.py:188) and oned_analysis
// (pseudocode/picture_encoding.py:203).  lift1..lift4: picture_decoding.py:205-261.
#pragma once

#include <type_traits>

#include "common.cuh"

namespace vc2b {

// Lifting sums accumulate in 32 bits except for stages with S >= kWideAccS (Fidelity, S = 8, eight
// taps up to 161; Daubechies (9,7), S = 12, taps up to 6497), which accumulate in 64 bits.  The int32
// envelope check (params.cu, pass_gain) bounds every 32-bit accumulator.  (With 32-bit sums Fidelity
// runs 1.6x faster but a depth-4 Fidelity transform would only be safe for |coefficient| <= 533.)
constexpr int kWideAccS = 8;

struct StageDef {
    int type;  // 1 even += odd, 2 even -= odd, 3 odd += even, 4 odd -= even
    int S, L, D;
    int taps[8];
};

struct FilterDef {
    int nstages;
    StageDef st[4];
    int shift;  // filter_bit_shift
    int halo;   // samples of context needed on each side of a tile (even)
};

__host__ __device__ constexpr FilterDef get_filter(int f)
{
    switch (f) {
    case 0:  // Deslauriers-Dubuc (9,7)
        return {2, {{2, 2, 2, 0, {1, 1}}, {3, 4, 4, -1, {-1, 9, 9, -1}}}, 1, 4};
    case 1:  // LeGall (5,3)
        return {2, {{2, 2, 2, 0, {1, 1}}, {3, 1, 2, 0, {1, 1}}}, 1, 2};
    case 2:  // Deslauriers-Dubuc (13,7)
        return {2, {{2, 5, 4, -1, {-1, 9, 9, -1}}, {3, 4, 4, -1, {-1, 9, 9, -1}}}, 1, 6};
    case 3:  // Haar, no shift
        return {2, {{2, 1, 1, 1, {1}}, {3, 0, 1, 0, {1}}}, 0, 0};
    case 4:  // Haar, with shift
        return {2, {{2, 1, 1, 1, {1}}, {3, 0, 1, 0, {1}}}, 1, 0};
    case 5:  // Fidelity
        return {2,
                {{3, 8, 8, -3, {-2, 10, -25, 81, 81, -25, 10, -2}},
                 {2, 8, 8, -3, {-8, 21, -46, 161, 161, -46, 21, -8}}},
                0,
                14};
    default:  // 6: Daubechies (9,7)
        return {4,
                {{2, 12, 2, 0, {1817, 1817}},
                 {4, 12, 2, 0, {3616, 3616}},
                 {1, 12, 2, 0, {217, 217}},
                 {3, 12, 2, 0, {6497, 6497}}},
                1,
                4};
    }
}

inline int filter_shift_host(int f) { return get_filter(f).shift; }
inline int filter_halo_host(int f) { return get_filter(f).halo; }

// One lifting stage over a shared-memory tile.
//   tile/tstride : tile storage; element (r, c) at tile[r * tstride + c]
//   rw, rh       : region extent held in the tile
//   g0           : global coordinate (along the filtered axis) of tile index 0; even
//   glen         : global length of the filtered axis (array bound for the clamps)
// VERT filters columns (axis = rows), otherwise rows.  Values near tile edges that are not
// picture edges come out wrong by construction; callers only consume the interior that
// lies `halo` samples inside the region.
template <int TYPE, int S, int L, int D, int T0, int T1, int T2, int T3, int T4, int T5, int T6, int T7,
          bool VERT>
__device__ __forceinline__ void lift_stage(int32_t *__restrict__ tile, int tstride, int rw, int rh, int g0,
                                           int glen)
{
    constexpr int taps[8] = {T0, T1, T2, T3, T4, T5, T6, T7};
    constexpr bool kEven = (TYPE == 1 || TYPE == 2);  // updates even positions from odd ones
    constexpr bool kAdd = (TYPE == 1 || TYPE == 3);
    const int len = VERT ? rh : rw;    // region length along the filtered axis
    const int other = VERT ? rw : rh;
    const int npairs = len >> 1;
    const int gn0 = g0 >> 1;
    for (int item = threadIdx.x; item < npairs * other; item += blockDim.x) {
        int n, o;
        if (VERT) { n = item / other; o = item - n * other; }
        else { o = item / npairs; n = item - o * npairs; }
        const int gn = gn0 + n;
        // large-tap filters (Fidelity, Daubechies) accumulate in 64 bits
        typename std::conditional<(S >= kWideAccS), long long, int>::type sum = 0;
#pragma unroll
        for (int i = D; i < L + D; i++) {
            int pos;
            if (kEven) {
                pos = 2 * (gn + i) - 1;
                pos = min(pos, glen - 1);
                pos = max(pos, 1);
            } else {
                pos = 2 * (gn + i);
                pos = min(pos, glen - 2);
                pos = max(pos, 0);
            }
            int lp = pos - g0;
            lp = max(0, min(lp, len - 1));  // keep garbage-margin reads inside the tile
            const int v = VERT ? tile[lp * tstride + o] : tile[o * tstride + lp];
            sum += (decltype(sum))taps[i - D] * v;
        }
        if (S > 0) sum += (decltype(sum))1 << (S > 0 ? S - 1 : 0);
        sum >>= S;
        const int tp = kEven ? 2 * n : 2 * n + 1;
        int32_t *dst = VERT ? &tile[tp * tstride + o] : &tile[o * tstride + tp];
        if (kAdd) *dst += (int)sum; else *dst -= (int)sum;
    }
    __syncthreads();
}

// add <-> subtract swap of the analysis direction (pseudocode/picture_encoding.py:214-223)
__host__ __device__ constexpr int analysis_type(int t) { return t == 1 ? 2 : t == 2 ? 1 : t == 3 ? 4 : 3; }

template <int F, int SIDX, bool VERT, bool ANALYSIS>
__device__ __forceinline__ void run_stage(int32_t *tile, int tstride, int rw, int rh, int g0, int glen)
{
    constexpr FilterDef f = get_filter(F);
    constexpr StageDef s = f.st[SIDX];
    constexpr int type = ANALYSIS ? analysis_type(s.type) : s.type;
    lift_stage<type, s.S, s.L, s.D, s.taps[0], s.taps[1], s.taps[2], s.taps[3], s.taps[4], s.taps[5], s.taps[6],
               s.taps[7], VERT>(tile, tstride, rw, rh, g0, glen);
}

// oned_synthesis (picture_decoding.py:188) / oned_analysis (picture_encoding.py:203) over every
// row or column of the tile.
template <int F, bool VERT, bool ANALYSIS>
__device__ __forceinline__ void lift_tile(int32_t *tile, int tstride, int rw, int rh, int g0, int glen)
{
    constexpr int ns = get_filter(F).nstages;
    if (!ANALYSIS) {
        run_stage<F, 0, VERT, false>(tile, tstride, rw, rh, g0, glen);
        if constexpr (ns > 1) run_stage<F, (ns > 1 ? 1 : 0), VERT, false>(tile, tstride, rw, rh, g0, glen);
        if constexpr (ns > 2) run_stage<F, (ns > 2 ? 2 : 0), VERT, false>(tile, tstride, rw, rh, g0, glen);
        if constexpr (ns > 3) run_stage<F, (ns > 3 ? 3 : 0), VERT, false>(tile, tstride, rw, rh, g0, glen);
    } else {
        if constexpr (ns > 3) run_stage<F, (ns > 3 ? 3 : 0), VERT, true>(tile, tstride, rw, rh, g0, glen);
        if constexpr (ns > 2) run_stage<F, (ns > 2 ? 2 : 0), VERT, true>(tile, tstride, rw, rh, g0, glen);
        if constexpr (ns > 1) run_stage<F, (ns > 1 ? 1 : 0), VERT, true>(tile, tstride, rw, rh, g0, glen);
        run_stage<F, 0, VERT, true>(tile, tstride, rw, rh, g0, glen);
    }
}

}  // namespace vc2b
