// wide_common.cuh -- pieces shared by the wide streaming synthesis (idwt8.cuh) and analysis
// (dwt8.cuh) kernels: vector types, the cp.async staging primitive and the horizontal lifting of
// a row held as column pairs across the lanes of a warp.
#pragma once

#include "stream_lift.cuh"

namespace vc2b {

constexpr int kWideNP = 4;
constexpr int kWideNC = 2 * kWideNP;

template <int NP> struct VecOf;
template <> struct VecOf<2> { using type = int2; using utype = uint2; };
template <> struct VecOf<4> { using type = int4; using utype = uint4; };
using WideVec = VecOf<kWideNP>::type;
using WideUVec = VecOf<kWideNP>::utype;

__device__ __forceinline__ void vec_to_array(const int2 &v, int (&a)[2]) { a[0] = v.x; a[1] = v.y; }
__device__ __forceinline__ void vec_to_array(const int4 &v, int (&a)[4]) { a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w; }
__device__ __forceinline__ int2 array_to_vec(const int (&a)[2]) { return make_int2(a[0], a[1]); }
__device__ __forceinline__ int4 array_to_vec(const int (&a)[4]) { return make_int4(a[0], a[1], a[2], a[3]); }
__device__ __forceinline__ uint2 array_to_uvec(const uint32_t (&a)[2]) { return make_uint2(a[0], a[1]); }
__device__ __forceinline__ uint4 array_to_uvec(const uint32_t (&a)[4]) { return make_uint4(a[0], a[1], a[2], a[3]); }

// Shared-memory staging of raw row pairs: stage slot (m mod STAGES) holds the four vectors of pair m
// for every lane (lane-private slots, so no warp synchronisation is needed); STAGES = 4 pairs of
// look-ahead plus the largest entry lag of the vertical filter (stream_lift.cuh).

__device__ __forceinline__ void cp_async_vec(WideVec *smem_dst, const WideVec *gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (sizeof(WideVec) == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}

// One horizontal lifting stage on a row held as E[i] = column 2*NP*l + 2i, O[i] = the next one.
// ---- L2 residency control -------------------------------------------------------------------------
// The levels of a transform run back to back over a group of pictures small enough for the
// intermediate LL bands to stay in the 126 MB L2 between the level that writes them and the level
// that reads them.  That only works if the one-pass streams (coefficients in, pictures out) do not
// push them out: streams are tagged evict_first, the intermediate bands evict_last.
enum L2Hint : int { kL2Normal = 0, kL2EvictFirst = 1, kL2EvictLast = 2 };

__device__ __forceinline__ uint64_t l2_policy(int hint)
{
    uint64_t pn, pf, pl;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pn));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pf));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pl));
    return hint == kL2EvictFirst ? pf : (hint == kL2EvictLast ? pl : pn);
}

__device__ __forceinline__ void cp_async_vec(WideVec *smem_dst, const WideVec *gsrc, uint64_t pol)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    // keep the shared address in one plain register: with the [R + UR + imm] form ptxas (12.9) put the
    // policy descriptor of LDGSTS in an odd uniform register and the kernel died with an illegal instruction
    asm volatile("mov.u32 %0, %0;" : "+r"(sa));
    if (sizeof(WideVec) == 8)
        asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(sa), "l"(gsrc), "l"(pol) : "memory");
    else
        asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(sa), "l"(gsrc), "l"(pol) : "memory");
}

__device__ __forceinline__ void st_hint(void *p, const int4 &v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(void *p, const uint4 &v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(void *p, const int2 &v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1, %2}, %3;" ::"l"(p), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(void *p, const uint2 &v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1, %2}, %3;" ::"l"(p), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
}

template <int F, int SIDX, bool ANA, typename Ctx>
__device__ __forceinline__ void hstage8(int (&E)[kWideNP], int (&O)[kWideNP], const Ctx &X)
{
    constexpr int NP = kWideNP;
    constexpr StreamPlan P = make_stream_plan(F, ANA);
    constexpr StageDef st = stream_stage(F, SIDX, ANA);
    constexpr bool even = (P.parity[SIDX] == 0);
    constexpr bool add = (st.type == 1 || st.type == 3);
    constexpr int dmin = P.dmin[SIDX], dmax = P.dmax[SIDX];
    constexpr int NL = dmin < 0 ? -dmin : 0, NR = dmax > 0 ? dmax : 0;
    static_assert(NL <= NP && NR <= NP, "neighbour lanes only");
    using Acc = typename std::conditional<(st.S >= kWideAccS), long long, int>::type;
    int src[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) src[i] = even ? O[i] : E[i];
    if (X.hedge) {  // replicate the boundary column of the source parity into out-of-range lanes
        const int lo = __shfl_sync(0xffffffffu, src[0], X.lane_lo & 31);
        const int hi = __shfl_sync(0xffffffffu, src[NP - 1], X.lane_hi & 31);
#pragma unroll
        for (int i = 0; i < NP; i++) {
            if (X.lane < X.lane_lo) src[i] = lo;
            if (X.lane > X.lane_hi) src[i] = hi;
        }
    }
    int ext[3 * NP];  // [0,NP) left lane, [NP,2NP) own, [2NP,3NP) right lane
#pragma unroll
    for (int i = 0; i < NP; i++) {
        ext[NP + i] = src[i];
        ext[i] = (i >= NP - NL) ? __shfl_sync(0xffffffffu, src[i], (X.lane + 31) & 31) : 0;
        ext[2 * NP + i] = (i < NR) ? __shfl_sync(0xffffffffu, src[i], (X.lane + 1) & 31) : 0;
    }
#pragma unroll
    for (int i = 0; i < NP; i++) {
        Acc sum = folded_taps<st.L, Acc>(st.taps, [&](int t) { return ext[NP + i + dmin + t]; },
                                         st.S > 0 ? (Acc)1 << (st.S > 0 ? st.S - 1 : 0) : (Acc)0);
        sum >>= st.S;
        if (even) { if (add) E[i] += (int)sum; else E[i] -= (int)sum; }
        else { if (add) O[i] += (int)sum; else O[i] -= (int)sum; }
    }
}

template <int F, bool ANA, typename Ctx>
__device__ __forceinline__ void hlift8(int (&E)[kWideNP], int (&O)[kWideNP], const Ctx &X)
{
    constexpr int ns = get_filter(F).nstages;
    hstage8<F, 0, ANA, Ctx>(E, O, X);
    if constexpr (ns > 1) hstage8<F, (ns > 1 ? 1 : 0), ANA, Ctx>(E, O, X);
    if constexpr (ns > 2) hstage8<F, (ns > 2 ? 2 : 0), ANA, Ctx>(E, O, X);
    if constexpr (ns > 3) hstage8<F, (ns > 3 ? 3 : 0), ANA, Ctx>(E, O, X);
}


}  // namespace vc2b
