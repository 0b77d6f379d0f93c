// stream_lift.cuh -- streaming (sliding-window) vertical lifting in registers.
//
// A thread walks DOWN its two columns one row pair per step.  Pair m = rows (2m, 2m+1).  Each
// lifting stage s of oned_synthesis (pseudocode/picture_decoding.py:188-261) updates, at step
// t, the pair t - lag[s]; the lags are chosen (at compile time, from the stage tables) so that
// every source sample a stage reads has already passed all earlier stages that write it and has
// not yet been touched by a later one -- exactly the values the reference's whole-array,
// stage-after-stage order sees.  The live pairs sit in a power-of-two register ring addressed
// with compile-time slots (the step loop is unrolled by the ring size), together with the
// prefetched raw pairs of the next few steps.
#pragma once

#include "lifting.cuh"

namespace vc2b {

struct StreamPlan {
    int ns;
    int parity[4];   // parity (0 even row, 1 odd row) the stage writes
    int dmin[4], dmax[4];  // source pair offsets relative to the pair being updated
    int lag[4];
    int maxlag;      // pair t - maxlag is final at the end of step t
    int back;        // oldest pair read at step t is t - back
    int entry[2];    // row parity q of pair m is first touched at step m + entry[q]: it may enter the ring late
    int backq[2];    // oldest pair of parity q still needed at step t is t - backq[q]
    int window;      // ring slots needed when each parity enters at its entry lag
    int warm;        // pairs of warm-up before the first exact output
    int prefetch;    // raw pairs loaded ahead of the step that first needs them
    int ring;        // power of two >= back + 1 + prefetch
};

// Stage `sidx` in order of application: synthesis applies the table order; analysis applies the
// stages reversed with add <-> subtract swapped (pseudocode/picture_encoding.py:203-223).
__host__ __device__ constexpr StageDef stream_stage(int filter, int sidx, bool analysis)
{
    const FilterDef f = get_filter(filter);
    StageDef st = f.st[analysis ? f.nstages - 1 - sidx : sidx];
    if (analysis) st.type = analysis_type(st.type);
    return st;
}

__host__ __device__ constexpr StreamPlan make_stream_plan(int filter, bool analysis = false)
{
    const FilterDef f = get_filter(filter);
    StreamPlan p = {};
    p.ns = f.nstages;
    for (int s = 0; s < f.nstages; s++) {
        const StageDef st = stream_stage(filter, s, analysis);
        const bool even = (st.type == 1 || st.type == 2);
        p.parity[s] = even ? 0 : 1;
        // even update reads odd sample n+i-1, odd update reads even sample n+i (i in [D, L+D))
        p.dmin[s] = even ? st.D - 1 : st.D;
        p.dmax[s] = even ? st.L + st.D - 2 : st.L + st.D - 1;
        int l = p.dmax[s] > 0 ? p.dmax[s] : 0;
        for (int q = 0; q < s; q++) {
            if (p.parity[q] == 1 - p.parity[s] && p.lag[q] + p.dmax[s] > l) l = p.lag[q] + p.dmax[s];
            if (1 - p.parity[q] == p.parity[s] && p.lag[q] - p.dmin[q] > l) l = p.lag[q] - p.dmin[q];
            if (p.lag[q] > l) l = p.lag[q];
        }
        p.lag[s] = l;
    }
    p.maxlag = p.lag[p.ns - 1];
    p.back = p.maxlag;
    for (int s = 0; s < p.ns; s++)
        if (p.lag[s] - p.dmin[s] > p.back) p.back = p.lag[s] - p.dmin[s];
    p.entry[0] = p.entry[1] = 1 << 20;
    p.backq[0] = p.backq[1] = p.maxlag;
    for (int s = 0; s < p.ns; s++) {
        const int w = p.parity[s], r = 1 - p.parity[s];
        if (p.lag[s] < p.entry[w]) p.entry[w] = p.lag[s];
        if (p.lag[s] - p.dmax[s] < p.entry[r]) p.entry[r] = p.lag[s] - p.dmax[s];
        if (p.lag[s] - p.dmin[s] > p.backq[r]) p.backq[r] = p.lag[s] - p.dmin[s];
    }
    for (int q = 0; q < 2; q++)
        if (p.entry[q] == (1 << 20) || p.entry[q] < 0) p.entry[q] = 0;
    p.window = 1;
    for (int q = 0; q < 2; q++)
        if (p.backq[q] - p.entry[q] + 1 > p.window) p.window = p.backq[q] - p.entry[q] + 1;
    p.warm = (f.halo + 1) / 2;
    p.prefetch = 4;
    int r = 8;
    while (r < p.back + 1 + p.prefetch) r *= 2;
    p.ring = r;
    return p;
}

// Register ring: v[row parity][column][slot]
template <int RING, int NC = 2>
struct PairRing {
    int v[2][NC][RING];
};

// Weighted sum of L taps plus `init` (the rounding constant).  32-bit sums: every VC-2 filter has
// symmetric taps, so equal outer pairs are folded (tap * (a + b)) to halve the multiplies.  64-bit sums
// (Fidelity, Daubechies): NOT folded -- a + b needs a 64-bit add before a 64 x 64 multiply, while
// tap * a accumulates in a single IMAD.WIDE.  get(i) returns the i-th source sample.
template <int SIDX_L, typename Acc, typename Get>
__device__ __forceinline__ Acc folded_taps(const int (&taps)[8], Get get, Acc init = 0)
{
    constexpr int L = SIDX_L;
    Acc sum = init;
    if (sizeof(Acc) == 8) {
#pragma unroll
        for (int i = 0; i < L; i++) sum += (Acc)taps[i] * (Acc)get(i);
        return sum;
    }
#pragma unroll
    for (int i = 0; i < L / 2; i++) {
        if (taps[i] == taps[L - 1 - i]) {
            sum += (Acc)taps[i] * ((Acc)get(i) + (Acc)get(L - 1 - i));
        } else {
            sum += (Acc)taps[i] * get(i) + (Acc)taps[L - 1 - i] * get(L - 1 - i);
        }
    }
    if (L & 1) sum += (Acc)taps[L / 2] * get(L / 2);
    return sum;
}

template <int RING>
__device__ __forceinline__ int ring_get_dyn(const int (&a)[RING], int slot)
{
    int r = 0;
#pragma unroll
    for (int i = 0; i < RING; i++)
        if (i == slot) r = a[i];
    return r;
}

template <int RING>
__device__ __forceinline__ void ring_add_dyn(int (&a)[RING], int slot, int d)
{
#pragma unroll
    for (int i = 0; i < RING; i++)
        if (i == slot) a[i] += d;
}

// One stage of one step, compile-time slots.  U = (t - ks) mod RING.
template <int F, int SIDX, int U, int RING, int NC, bool ANA = false>
__device__ __forceinline__ void vstage_fast(PairRing<RING, NC> &R)
{
    constexpr StreamPlan P = make_stream_plan(F, ANA);
    constexpr StageDef st = stream_stage(F, SIDX, ANA);
    constexpr int par = P.parity[SIDX];
    constexpr int k = U - P.lag[SIDX];
    constexpr bool add = (st.type == 1 || st.type == 3);
    using Acc = typename std::conditional<(st.S >= kWideAccS), long long, int>::type;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        Acc sum = folded_taps<st.L, Acc>(st.taps, [&](int i) {
            return R.v[1 - par][c][(k + P.dmin[SIDX] + i) & (RING - 1)];
        }, st.S > 0 ? (Acc)1 << (st.S > 0 ? st.S - 1 : 0) : (Acc)0);
        sum >>= st.S;
        if (add) R.v[par][c][k & (RING - 1)] += (int)sum;
        else R.v[par][c][k & (RING - 1)] -= (int)sum;
    }
}

// Same with run-time pair indices and the edge clamps of lift1-4 (picture_decoding.py:210-212,
// :240-242): a source pair outside [0, last] reads pair 0 / pair `last`.  t is the absolute
// step; ks the pair stored in slot 0 (slots are (m - ks) mod RING).
template <int F, int SIDX, int RING, int NC, bool ANA = false>
__device__ __forceinline__ void vstage_slow(PairRing<RING, NC> &R, int t, int ks, int last)
{
    constexpr StreamPlan P = make_stream_plan(F, ANA);
    constexpr StageDef st = stream_stage(F, SIDX, ANA);
    constexpr int par = P.parity[SIDX];
    constexpr bool add = (st.type == 1 || st.type == 3);
    const int k = t - P.lag[SIDX];
    if (k < 0 || k > last) return;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        typename std::conditional<(st.S >= kWideAccS), long long, int>::type sum = 0;
#pragma unroll
        for (int i = 0; i < st.L; i++) {
            int m = k + P.dmin[SIDX] + i;
            m = m < 0 ? 0 : (m > last ? last : m);
            sum += (decltype(sum))st.taps[i] * ring_get_dyn<RING>(R.v[1 - par][c], (m - ks) & (RING - 1));
        }
        if (st.S > 0) sum += (decltype(sum))1 << (st.S > 0 ? st.S - 1 : 0);
        sum >>= st.S;
        ring_add_dyn<RING>(R.v[par][c], (k - ks) & (RING - 1), add ? (int)sum : -(int)sum);
    }
}

template <int F, int U, int RING, int NC, bool ANA = false>
__device__ __forceinline__ void vstep_fast(PairRing<RING, NC> &R)
{
    constexpr int ns = get_filter(F).nstages;
    vstage_fast<F, 0, U, RING, NC, ANA>(R);
    if constexpr (ns > 1) vstage_fast<F, (ns > 1 ? 1 : 0), U, RING, NC, ANA>(R);
    if constexpr (ns > 2) vstage_fast<F, (ns > 2 ? 2 : 0), U, RING, NC, ANA>(R);
    if constexpr (ns > 3) vstage_fast<F, (ns > 3 ? 3 : 0), U, RING, NC, ANA>(R);
}

template <int F, int RING, int NC, bool ANA = false>
__device__ __forceinline__ void vstep_slow(PairRing<RING, NC> &R, int t, int ks, int last)
{
    constexpr int ns = get_filter(F).nstages;
    vstage_slow<F, 0, RING, NC, ANA>(R, t, ks, last);
    if constexpr (ns > 1) vstage_slow<F, (ns > 1 ? 1 : 0), RING, NC, ANA>(R, t, ks, last);
    if constexpr (ns > 2) vstage_slow<F, (ns > 2 ? 2 : 0), RING, NC, ANA>(R, t, ks, last);
    if constexpr (ns > 3) vstage_slow<F, (ns > 3 ? 3 : 0), RING, NC, ANA>(R, t, ks, last);
}

}  // namespace vc2b
