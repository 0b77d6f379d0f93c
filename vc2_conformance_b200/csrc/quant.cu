// quant.cu -- quantiser kernels: element-wise inverse/forward quantisation, the per-slice
// quantisation-index search of the encoder, and slice bit-packing.
//
// Replaces inverse_quant / forward_quant / quant_factor / quant_offset
// (pseudocode/quantization.py:18-62), clip/offset components
// (pseudocode/picture_decoding.py:301-349), quantize_coeffs / calculate_coeffs_bits /
// quantize_to_fit (encoder/pictures.py:217-347) as driven by make_transform_data_hq_lossy
// (:617) and make_transform_data_ld_lossy (:724), and the serialisation of hq_slice / ld_slice
// (bitstream/vc2.py:727-839) with write_sint (bitstream/io.py:557-583).
#include <string.h>

#include "common.cuh"

namespace vc2b {

constexpr int kQWarps = 4;

// ---------------------------------------------------------------------------
// element-wise
// ---------------------------------------------------------------------------

__device__ __forceinline__ int32_t inverse_quant_dev(int32_t q, int qi, bool *overflow)
{
    if (q == 0) return 0;
    const uint32_t m = q < 0 ? (uint32_t)(-(long long)q) : (uint32_t)q;
    const unsigned long long t = ((unsigned long long)m * quant_factor_u64(qi) + quant_offset_u64(qi) + 2) >> 2;
    if (t > 0x7fffffffull) *overflow = true;
    return q < 0 ? -(int32_t)t : (int32_t)t;
}

// forward_quant (quantization.py:30): sign(c) * ((4*|c|) // quant_factor)
__device__ __forceinline__ uint32_t forward_quant_mag(uint32_t m, unsigned long long qf)
{
    const unsigned long long num = 4ull * m;
    if (((num | qf) >> 32) == 0) return (uint32_t)num / (uint32_t)qf;
    return (uint32_t)(num / qf);
}

__global__ void inverse_quant_kernel(const int32_t *__restrict__ in, const int32_t *__restrict__ qi,
                                     int32_t *__restrict__ out, long long n)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        bool ov = false;
        out[i] = inverse_quant_dev(in[i], max(qi[i], 0), &ov);
    }
}

__global__ void forward_quant_kernel(const int32_t *__restrict__ in, const int32_t *__restrict__ qi,
                                     int32_t *__restrict__ out, long long n)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int32_t c = in[i];
        const uint32_t m = c < 0 ? (uint32_t)(-(long long)c) : (uint32_t)c;
        const uint32_t r = forward_quant_mag(m, quant_factor_u64(max(qi[i], 0)));
        out[i] = c < 0 ? -(int32_t)r : (int32_t)r;
    }
}

__global__ void clip_offset_kernel(int32_t *__restrict__ data, long long n, int depth, int mode)
{
    const int half = 1 << (depth - 1);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        int v = data[i];
        if (mode & 1) v = min(max(v, -half), half - 1);
        if (mode & 2) v += half;
        if (mode & 4) v -= half;
        data[i] = v;
    }
}

// ---------------------------------------------------------------------------
// per-slice tables shared by the qindex search and the packer
// ---------------------------------------------------------------------------

struct SliceBands {
    int32_t cum[kMaxBands + 1];
    int32_t w[kMaxBands];
    int32_t stride[kMaxBands];
    int32_t base[kMaxBands];
    int32_t qm[kMaxBands];
};

__device__ __forceinline__ int build_slice_bands(const DevParams &P, int comp, int sx, int sy, SliceBands &T,
                                                 int lane)
{
    const int nb = P.nb;
    int cnt = 0;
    if (lane < nb) {
        const Band &B = P.band[comp][lane];
        int x1, x2, y1, y2;  // slice_left / right / top / bottom (pseudocode/slice_sizes.py:19-50)
        if ((long long)B.w * (sx + 1) < 0x7fffffffLL && (long long)B.h * (sy + 1) < 0x7fffffffLL) {
            x1 = (int)(((unsigned)B.w * (unsigned)sx) / (unsigned)P.slices_x);
            x2 = (int)(((unsigned)B.w * (unsigned)(sx + 1)) / (unsigned)P.slices_x);
            y1 = (int)(((unsigned)B.h * (unsigned)sy) / (unsigned)P.slices_y);
            y2 = (int)(((unsigned)B.h * (unsigned)(sy + 1)) / (unsigned)P.slices_y);
        } else {
            x1 = (int)(((long long)B.w * sx) / P.slices_x);
            x2 = (int)(((long long)B.w * (sx + 1)) / P.slices_x);
            y1 = (int)(((long long)B.h * sy) / P.slices_y);
            y2 = (int)(((long long)B.h * (sy + 1)) / P.slices_y);
        }
        T.w[lane] = max(x2 - x1, 1);
        T.stride[lane] = B.w;
        T.base[lane] = B.off + y1 * B.w + x1;
        T.qm[lane] = B.qm;
        cnt = (x2 - x1) * (y2 - y1);
    }
    int pre = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, pre, d);
        if (lane >= d) pre += o;
    }
    if (lane < nb) T.cum[lane + 1] = pre;
    if (lane == 0) T.cum[0] = 0;
    __syncwarp();
    return __shfl_sync(0xffffffffu, pre, 31);
}

// Band and offset (inside the component) of coefficient j of the slice component: bisection over the
// cumulative band sizes (branch-free, so that callers can keep several loads in flight).
__device__ __forceinline__ int locate_coeff(const SliceBands &T, int nb, int j, int *band)
{
    int b = 0;
#pragma unroll
    for (int s = kMaxBands / 2; s > 0; s >>= 1) {
        const int t = b + s;
        if (t < nb && T.cum[t] <= j) b = t;  // largest b with cum[b] <= j: the (non-empty) band holding j
    }
    const int local = j - T.cum[b];
    const int w = T.w[b];
    int y, x;
    if (local < (1 << 24)) {
        y = (int)(((float)local + 0.5f) * (1.0f / (float)w));  // off by at most one, fixed below
        x = local - y * w;
        if (x < 0) { y--; x += w; }
        else if (x >= w) { y++; x -= w; }
    } else {
        y = local / w;
        x = local - y * w;
    }
    *band = b;
    return T.base[b] + y * T.stride[b] + x;
}

// coefficient j (bitstream order, encoder/pictures.py:418-451) of the slice component
__device__ __forceinline__ int32_t slice_coeff(const SliceBands &T, int nb, const int32_t *__restrict__ comp,
                                               int j, int *band)
{
    return comp[locate_coeff(T, nb, j, band)];
}

// signed_exp_golomb_length (bitstream/exp_golomb.py:28)
__device__ __forceinline__ int sint_length(uint32_t mag)
{
    if (mag == 0) return 1;
    const int k = 31 - __clz(mag + 1u);  // mag <= 2^31: no wrap
    return 2 * k + 2;
}

// calculate_coeffs_bits (encoder/pictures.py:217) of one slice component quantised at qindex
// (quantize_coeffs :251).  INTERLEAVED: C1/C2 interleaved per coefficient (:712).
template <bool INTERLEAVED>
__device__ int slice_bits(const SliceBands &T, int nb, const int32_t *__restrict__ c1,
                          const int32_t *__restrict__ c2, int ncoef, int qindex, int lane)
{
    int sum = 0, last = -1;
    for (int idx = lane; idx < ncoef; idx += 32) {
        const int j = INTERLEAVED ? (idx >> 1) : idx;
        int b;
        const int32_t c = slice_coeff(T, nb, (INTERLEAVED && (idx & 1)) ? c2 : c1, j, &b);
        const uint32_t m = c < 0 ? (uint32_t)(-(long long)c) : (uint32_t)c;
        const uint32_t q = forward_quant_mag(m, quant_factor_u64(max(qindex - T.qm[b], 0)));
        sum += sint_length(q);
        if (q != 0) last = idx;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
        last = max(last, __shfl_xor_sync(0xffffffffu, last, d));
    }
    return sum - (ncoef - 1 - last);  // trailing zeros are free
}

// ---- slice staged in shared memory ------------------------------------------------------------
// The search probes every coefficient of the slice several times and the packer needs the values in
// bitstream order; gathering the slice once (band by band: every band rectangle is a handful of
// contiguous row segments) makes a probe a pure shared-memory pass.
constexpr int kStageMax = 2048;  // coefficients (all components) of a slice that fit the staging buffer

// Per-warp staging area carved out of dynamic shared memory: `cap` = the largest slice of the
// parameter set (rounded up), so small slices leave room for more resident warps.
struct SliceStage {
    uint2 *qtab;     // [kMaxBands] per probe and band: {min(qf, 2^30), floor(log2 of it)}
    int32_t *coef;   // [cap] bitstream order: Y, then C1, C2 (HQ) or C1/C2 interleaved (LD); 4*|c|
    uint8_t *band;   // [cap]
};

__host__ __device__ inline size_t slice_stage_bytes(int cap) { return kMaxBands * sizeof(uint2) + 5 * (size_t)cap; }

__device__ __forceinline__ SliceStage slice_stage_at(unsigned char *base, int warp, int cap)
{
    unsigned char *p = base + warp * slice_stage_bytes(cap);
    SliceStage S;
    S.qtab = reinterpret_cast<uint2 *>(p);
    S.coef = reinterpret_cast<int32_t *>(p + kMaxBands * sizeof(uint2));
    S.band = p + kMaxBands * sizeof(uint2) + 4 * (size_t)cap;
    return S;
}

// Gathers the slice rectangle of one component into dst[j * step] (step 2: C1/C2 interleaved, ld_slice),
// eight independent loads per lane in flight.  ABS4: store 4*|c| (saturated at 2^32-1 for |c| >= 2^30);
// returns the lane's largest |c|.
template <bool ABS4>
__device__ __forceinline__ uint32_t stage_component(const SliceBands &T, int nb, const int32_t *__restrict__ comp,
                                                    int n, int32_t *dst, uint8_t *dband, int step, int lane)
{
    constexpr int U = 8;
    uint32_t mx = 0;
    for (int j0 = 0; j0 < n; j0 += 32 * U) {
        int32_t v[U];
        int bb[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = j0 + 32 * u + lane;
            v[u] = 0;
            bb[u] = 0;
            if (j < n) v[u] = __ldg(comp + locate_coeff(T, nb, j, &bb[u]));
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = j0 + 32 * u + lane;
            if (j < n) {
                const int32_t c = v[u];
                if (ABS4) {
                    const uint32_t m = c < 0 ? (uint32_t)(-(long long)c) : (uint32_t)c;
                    mx = max(mx, m);
                    dst[j * step] = (int32_t)(m >= 0x40000000u ? 0xffffffffu : 4u * m);
                } else {
                    dst[j * step] = c;
                }
                dband[j * step] = (uint8_t)bb[u];
            }
        }
    }
    return mx;
}

template <bool ABS4>
__device__ __forceinline__ uint32_t stage_slice(const DevParams &P, const SliceBands &TY, const SliceBands &TC,
                                                const int32_t *__restrict__ pc, int ny, int nc, const SliceStage &S, int lane)
{
    const int32_t *cy = pc + P.comp_off[0], *c1 = pc + P.comp_off[1], *c2 = pc + P.comp_off[2];
    uint32_t mx = stage_component<ABS4>(TY, P.nb, cy, ny, S.coef, S.band, 1, lane);
    if (P.is_ld) {
        mx = max(mx, stage_component<ABS4>(TC, P.nb, c1, nc, S.coef + ny, S.band + ny, 2, lane));
        mx = max(mx, stage_component<ABS4>(TC, P.nb, c2, nc, S.coef + ny + 1, S.band + ny + 1, 2, lane));
    } else {
        mx = max(mx, stage_component<ABS4>(TC, P.nb, c1, nc, S.coef + ny, S.band + ny, 1, lane));
        mx = max(mx, stage_component<ABS4>(TC, P.nb, c2, nc, S.coef + ny + nc, S.band + ny + nc, 1, lane));
    }
    __syncwarp();
    return __reduce_max_sync(0xffffffffu, mx);
}

// ---- quantiser search probe -------------------------------------------------------------------
// The coded length of a value only needs k = floor(log2(q + 1)), q = floor(n / qf), n = 4|c|
// (signed_exp_golomb_length: 1 bit for q = 0, else 2k + 2).  q + 1 = floor((n + qf) / qf), so with
// x = n + qf and d = floor(log2 x) - floor(log2 qf):  k = d if (qf << d) <= x, else d - 1 -- no
// division.  Valid for n < 2^30 with qf clamped to 2^30 (a larger factor quantises every such n to
// zero, and so does 2^30); slices holding larger magnitudes take the generic path.
constexpr uint32_t kFastMagLimit = 0x10000000u;  // |c| < 2^28

__device__ __forceinline__ void stage_search_quantisers(const SliceBands &T, int nb, int qindex, const SliceStage &S, int lane)
{
    __syncwarp();
    if (lane < nb) {
        const uint64_t qf = quant_factor_u64(max(qindex - T.qm[lane], 0));
        const uint32_t y = qf > 0x40000000ull ? 0x40000000u : (uint32_t)qf;
        S.qtab[lane] = make_uint2(y, 31 - __clz(y));
    }
    __syncwarp();
}

// calculate_coeffs_bits (encoder/pictures.py:217) of the staged block [j0, j0 + cnt) at the staged quantisers
__device__ __forceinline__ int staged_block_bits(const SliceStage &S, int j0, int cnt, int lane)
{
    int acc = 0, last = -1;  // acc = sum of (2k + [k > 0])
#pragma unroll 4
    for (int j = lane; j < cnt; j += 32) {
        const uint32_t n = (uint32_t)S.coef[j0 + j];
        const uint2 t = S.qtab[S.band[j0 + j]];
        const uint32_t x = n + t.x;
        const int d = (31 - __clz(x)) - (int)t.y;
        const int k = d - (((t.x << d) > x) ? 1 : 0);
        acc += 2 * k;
        if (k > 0) { acc++; last = j; }
    }
    acc = __reduce_add_sync(0xffffffffu, acc);
    last = __reduce_max_sync(0xffffffffu, last);
    // sum of lengths = acc + cnt; trailing zeros (cnt - 1 - last) are free
    return acc + last + 1;
}

// ---- packer quantisation ----------------------------------------------------------------------
__device__ __forceinline__ void stage_pack_quantisers(const SliceBands &T, int nb, int qindex, uint2 *qtab, int lane)
{
    __syncwarp();
    if (lane < nb) {
        const uint64_t qf = quant_factor_u64(max(qindex - T.qm[lane], 0));
        const uint32_t q32 = qf > 0xfffffffeull ? 0xffffffffu : (uint32_t)qf;
        qtab[lane] = make_uint2(q32, 0xffffffffu / q32);
    }
    __syncwarp();
}

// forward_quant magnitude with the staged reciprocal: exact floor(4m / qf)
__device__ __forceinline__ uint32_t staged_quant_mag(const uint2 *qtab, int b, int32_t c)
{
    const uint32_t m = c < 0 ? (uint32_t)(-(long long)c) : (uint32_t)c;
    const uint2 t = qtab[b];
    const uint32_t qf = t.x;
    if (m >= 0x40000000u) return qf == 0xffffffffu ? 0u : (uint32_t)((4ull * m) / qf);  // 4m overflows 32 bits
    if (qf == 0xffffffffu) return 0u;  // quant_factor >= 2^32 > 4m
    const uint32_t n = 4u * m;
    uint32_t q = __umulhi(n, t.y);
    uint32_t r = n - q * qf;
    while (r >= qf) { q++; r -= qf; }  // at most two corrections
    return q;
}

struct FitArgs {
    long long total_coeff_bytes;  // HQ: picture_bytes - 4*slices
    int minimum_qindex;
    int max_qm;
    int stage_cap;                // coefficients the per-warp staging area holds (quantize_fit_kernel)
    int use_map;                  // every band fits the 24-bit offsets of FitMap
    int uniform;                  // every band divides evenly between the slices: all slices have the same rectangles
    int fixed_qindex;             // >= 0: no search, only the block sizes at this index (vc2b_slice_bits)
    const int32_t *fixed_per_slice;  // != nullptr: no search, block sizes at the slice's own index
    const uint32_t *lossless_offset; // packer, make_transform_data_hq_lossless (:528): [npictures * (slices + 1)]
                                     // byte offset of every slice inside its picture; minimal length fields
};

__device__ __forceinline__ long long hq_total_length(const DevParams &P, const FitArgs &F, long long n)
{
    // slice_bytes (slice_sizes.py:95) re-purposed by make_transform_data_hq_lossy (:672-690)
    const long long den = (long long)P.slices_x * P.slices_y * P.slice_size_scaler;
    return ((n + 1) * F.total_coeff_bytes) / den - (n * F.total_coeff_bytes) / den;
}

// last qindex found for a picture: the search hint of a warp's FIRST slice (shared by the warps of that
// picture; stale or racy values only cost probes).  Later slices start from the warp's own previous answer
// -- the slice next door -- which costs no dependent global load.
constexpr int kHintSlots = 1024;
__device__ int g_qindex_hint[kHintSlots];

// Coefficient -> (band, offset from the slice's corner of that band) map of the CTA's reference slice:
// when every slice of the geometry has the same rectangles (all BASELINE configurations) staging a
// coefficient is one table look-up instead of a band search and a division.
struct FitMap {
    uint32_t e[kStageMax];            // band << 24 | (y * band width + x); luma entries, then chroma
    int32_t ref_w[2][kMaxBands];
    int32_t ref_cnt[2][kMaxBands];
    int32_t ny, nc, ok;
};

// Builds the map from slice (0, 0) with all threads of the CTA (tabs: two scratch SliceBands of warp 0).
// `limit`: the map is only used when the reference slice has at most that many coefficients per picture.
__device__ __forceinline__ void build_fit_map(const DevParams &P, FitMap &M, SliceBands *tabs0, bool enable, int limit)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        const int ny0 = build_slice_bands(P, 0, 0, 0, tabs0[0], lane);
        const int nc0 = build_slice_bands(P, 1, 0, 0, tabs0[1], lane);
        if (lane < P.nb) {
            for (int k = 0; k < 2; k++) {
                M.ref_w[k][lane] = tabs0[k].w[lane];
                M.ref_cnt[k][lane] = tabs0[k].cum[lane + 1] - tabs0[k].cum[lane];
            }
        }
        if (lane == 0) {
            M.ny = ny0;
            M.nc = nc0;
            M.ok = enable && (ny0 + nc0 <= kStageMax) && (ny0 + 2 * nc0 <= limit);
        }
    }
    __syncthreads();
    if (M.ok) {
        for (int j = threadIdx.x; j < M.ny + M.nc; j += blockDim.x) {
            const SliceBands &T = j < M.ny ? tabs0[0] : tabs0[1];
            int b;
            const int off = locate_coeff(T, P.nb, j < M.ny ? j : j - M.ny, &b);
            M.e[j] = ((uint32_t)b << 24) | (uint32_t)(off - T.base[b]);
        }
    }
    __syncthreads();
}

// whether the slice described by TY / TC has the rectangles of the reference slice (warp-uniform)
__device__ __forceinline__ bool matches_fit_map(const DevParams &P, const FitMap &M, const SliceBands &TY,
                                                const SliceBands &TC, int lane)
{
    bool same = M.ok != 0;
    if (same && lane < P.nb)
        same = TY.w[lane] == M.ref_w[0][lane] && TY.cum[lane + 1] - TY.cum[lane] == M.ref_cnt[0][lane] &&
               TC.w[lane] == M.ref_w[1][lane] && TC.cum[lane + 1] - TC.cum[lane] == M.ref_cnt[1][lane];
    return __all_sync(0xffffffffu, same);
}

// stage one (NSRC = 1) or two (C1 and C2, same geometry) components through the map: dst[j * step] for
// j < n.  Sixteen loads per lane are issued before the first one is used: a slice is a few hundred
// scattered 8..64-byte row segments, and the stage is bound by their latency, not by bandwidth.
template <int NSRC>
__device__ __forceinline__ uint32_t stage_mapped(const uint32_t *__restrict__ map, const SliceBands &T,
                                                 const int32_t *__restrict__ c1, const int32_t *__restrict__ c2, int n,
                                                 int32_t *d1, uint8_t *b1, int32_t *d2, uint8_t *b2, int step, int lane)
{
    constexpr int U = 16 / NSRC;
    uint32_t mx = 0;
    for (int j0 = 0; j0 < n; j0 += 32 * U) {
        int32_t v1[U], v2[U];
        int bb[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = j0 + 32 * u + lane;
            v1[u] = v2[u] = 0;
            bb[u] = 0;
            if (j < n) {
                const uint32_t e = map[j];
                bb[u] = (int)(e >> 24);
                const int off = T.base[bb[u]] + (int)(e & 0xffffffu);
                v1[u] = __ldg(c1 + off);
                if (NSRC == 2) v2[u] = __ldg(c2 + off);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = j0 + 32 * u + lane;
            if (j < n) {
                const uint32_t m1 = v1[u] < 0 ? (uint32_t)(-(long long)v1[u]) : (uint32_t)v1[u];
                mx = max(mx, m1);
                d1[j * step] = (int32_t)(m1 >= 0x40000000u ? 0xffffffffu : 4u * m1);
                b1[j * step] = (uint8_t)bb[u];
                if (NSRC == 2) {
                    const uint32_t m2 = v2[u] < 0 ? (uint32_t)(-(long long)v2[u]) : (uint32_t)v2[u];
                    mx = max(mx, m2);
                    d2[j * step] = (int32_t)(m2 >= 0x40000000u ? 0xffffffffu : 4u * m2);
                    b2[j * step] = (uint8_t)bb[u];
                }
            }
        }
    }
    return mx;
}

__global__ void __launch_bounds__(kQWarps * 32)
quantize_fit_kernel(DevParams P, FitArgs F, const int32_t *__restrict__ coeffs, int npictures,
                    int32_t *__restrict__ qindex_out, int32_t *__restrict__ bits_out)
{
    __shared__ SliceBands tabs[kQWarps][2];
    __shared__ FitMap M;
    extern __shared__ __align__(16) unsigned char stage_mem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nslices = P.slices_x * P.slices_y;
    const long long total_slices = (long long)npictures * nslices;

    build_fit_map(P, M, tabs[0], F.use_map != 0, F.stage_cap);

  // A warp walks a contiguous run of slices, so its previous answer is the answer of the slice next door.
  // (Fetching the next slice's coefficients by cp.async behind the current search was tried and lost:
  // 163 against 150 us per 2160p picture -- the extra staging buffer costs a resident CTA per SM.)
  int own_hint = -1;
  const long long nwarps_total = (long long)gridDim.x * kQWarps;
  const long long run = (total_slices + nwarps_total - 1) / nwarps_total;
  const long long gw0 = ((long long)blockIdx.x * kQWarps + warp) * run;
  SliceBands &TY = tabs[warp][0];
  SliceBands &TC = tabs[warp][1];
  // Uniform geometry (every BASELINE configuration): the slice tables are the same for every slice but for the
  // corners of the rectangles, and slice number, position and budget advance by additions -- the divisions of
  // build_slice_bands, gw / nslices and slice_bytes were a third of this kernel's instructions.
  const bool uni = F.uniform != 0;
  int u_ny = 0, u_nc = 0, rw_y = 0, rh_y = 0, rw_c = 0, rh_c = 0;
  if (uni) {
      __syncwarp();
      u_ny = build_slice_bands(P, 0, 0, 0, TY, lane);
      u_nc = build_slice_bands(P, 1, 0, 0, TC, lane);
      if (lane < P.nb) {
          rw_y = P.band[0][lane].w / P.slices_x;
          rh_y = P.band[0][lane].h / P.slices_y;
          rw_c = P.band[1][lane].w / P.slices_x;
          rh_c = P.band[1][lane].h / P.slices_y;
      }
  }
  // running slice position and budget: floor((n+1) T / D) - floor(n T / D) = floor((n T mod D + T) / D)
  const unsigned long long bud_t = P.is_ld ? (unsigned long long)P.sb_num : (unsigned long long)F.total_coeff_bytes;
  const unsigned long long bud_d = P.is_ld ? (unsigned long long)P.sb_den
                                           : (unsigned long long)nslices * (unsigned long long)P.slice_size_scaler;
  int cur_pic = (int)(gw0 / nslices);
  int cur_n = (int)(gw0 - (long long)cur_pic * nslices);
  int cur_sy = cur_n / P.slices_x, cur_sx = cur_n - cur_sy * P.slices_x;
  unsigned long long bud_rem = uni ? ((unsigned long long)cur_n * bud_t) % bud_d : 0ull;
  for (long long gw = gw0; gw < min(gw0 + run, total_slices); gw++) {
    int pic, n, sy, sx;
    if (uni) {
        pic = cur_pic; n = cur_n; sy = cur_sy; sx = cur_sx;
    } else {
        pic = (int)(gw / nslices);
        n = (int)(gw - (long long)pic * nslices);
        sy = n / P.slices_x;
        sx = n - sy * P.slices_x;
    }
    // issued before the staging loads so that its latency hides behind them
    const int shared_hint = own_hint < 0 ? g_qindex_hint[pic & (kHintSlots - 1)] : 0;
    const int32_t *pc = coeffs + (size_t)pic * P.pic_coeffs;
    const SliceStage S = slice_stage_at(stage_mem, warp, F.stage_cap);
    __syncwarp();
    int ny, nc;
    long long uni_len = 0;   // hq_total_length / slice_bytes of this slice
    if (uni) {
        if (lane < P.nb) {
            TY.base[lane] = P.band[0][lane].off + (rh_y * sy) * P.band[0][lane].w + rw_y * sx;
            TC.base[lane] = P.band[1][lane].off + (rh_c * sy) * P.band[1][lane].w + rw_c * sx;
        }
        __syncwarp();
        ny = u_ny;
        nc = u_nc;
        const unsigned long long ssum = bud_rem + bud_t;
        const unsigned long long dq = (ssum >> 32) == 0 && (bud_d >> 32) == 0 ? (unsigned long long)((uint32_t)ssum / (uint32_t)bud_d)
                                                                           : ssum / bud_d;
        uni_len = (long long)dq;
        bud_rem = ssum - dq * bud_d;
        // the next slice of the run
        if (++cur_sx == P.slices_x) { cur_sx = 0; cur_sy++; }
        if (++cur_n == nslices) { cur_n = 0; cur_sx = 0; cur_sy = 0; cur_pic++; bud_rem = 0; }
    } else {
        ny = build_slice_bands(P, 0, sx, sy, TY, lane);
        nc = build_slice_bands(P, 1, sx, sy, TC, lane);
    }
    // fast path: slice staged as 4|c| in shared memory, every magnitude below kFastMagLimit
    bool staged = (ny + 2 * nc) <= F.stage_cap;
    if (staged) {
        uint32_t mx;
        if (uni ? (M.ok != 0) : matches_fit_map(P, M, TY, TC, lane)) {
            const int32_t *c0 = pc + P.comp_off[0], *c1 = pc + P.comp_off[1], *c2 = pc + P.comp_off[2];
            mx = stage_mapped<1>(M.e, TY, c0, nullptr, ny, S.coef, S.band, nullptr, nullptr, 1, lane);
            if (P.is_ld)
                mx = max(mx, stage_mapped<2>(M.e + ny, TC, c1, c2, nc, S.coef + ny, S.band + ny, S.coef + ny + 1,
                                             S.band + ny + 1, 2, lane));
            else
                mx = max(mx, stage_mapped<2>(M.e + ny, TC, c1, c2, nc, S.coef + ny, S.band + ny, S.coef + ny + nc,
                                             S.band + ny + nc, 1, lane));
            __syncwarp();
            mx = __reduce_max_sync(0xffffffffu, mx);
        } else {
            mx = stage_slice<true>(P, TY, TC, pc, ny, nc, S, lane);
        }
        staged = mx < kFastMagLimit;
    }

    long long target, align;
    if (P.is_ld) {
        const long long sb = uni ? uni_len : ld_slice_offset(P, n + 1) - ld_slice_offset(P, n);
        target = 8 * sb - 7;
        target -= ilog2_ceil(target);
        align = 1;
    } else {
        align = 8LL * P.slice_size_scaler;
        target = align * (uni ? uni_len : hq_total_length(P, F, n));
    }
    const int ialign = (int)align;

    int by = 0, bc1 = 0, bc2 = 0;   // bits of the last probe
    int ok_q = -1, oy = 0, o1 = 0, o2 = 0;  // last probe that fitted
    auto fits = [&](int q) -> bool {
        long long total;
        if (staged) {  // quant matrices are per band, identical for the three components
            stage_search_quantisers(TY, P.nb, q, S, lane);
            by = staged_block_bits(S, 0, ny, lane);
            if (P.is_ld) {
                bc1 = staged_block_bits(S, ny, 2 * nc, lane);
                bc2 = 0;
                total = (long long)by + bc1;
            } else {
                bc1 = staged_block_bits(S, ny, nc, lane);
                bc2 = staged_block_bits(S, ny + nc, nc, lane);
                total = (long long)((by + ialign - 1) / ialign + (bc1 + ialign - 1) / ialign + (bc2 + ialign - 1) / ialign) *
                        ialign;
            }
        } else {
            by = slice_bits<false>(TY, P.nb, pc + P.comp_off[0], nullptr, ny, q, lane);
            if (P.is_ld) {
                bc1 = slice_bits<true>(TC, P.nb, pc + P.comp_off[1], pc + P.comp_off[2], 2 * nc, q, lane);
                bc2 = 0;
                total = (long long)by + bc1;
            } else {
                bc1 = slice_bits<false>(TC, P.nb, pc + P.comp_off[1], nullptr, nc, q, lane);
                bc2 = slice_bits<false>(TC, P.nb, pc + P.comp_off[2], nullptr, nc, q, lane);
                total = ((by + align - 1) / align + (bc1 + align - 1) / align + (bc2 + align - 1) / align) * align;
            }
        }
        const bool ok = total <= target;
        if (ok) { ok_q = q; oy = by; o1 = bc1; o2 = bc2; }
        return ok;
    };

    // The total is non-increasing in qindex (|forward_quant| is non-increasing in the quantiser),
    // so the reference's linear search (quantize_to_fit :329) equals this bisection.
    int lo = F.minimum_qindex;
    int hi = max(lo, 128 + F.max_qm + 4);  // everything quantises to zero here (int32 envelope)
    // Neighbouring slices of a picture usually need almost the same index: gallop outwards from the last
    // answer (a hint only -- every bound used below is established by an actual probe), then bisect the
    // bracket, then fetch the block sizes at the answer unless the last probe that fitted was the answer.
    // One loop with ONE probe site: with the probe inlined at every step of the gallops and the bisection the
    // kernel spent a stall_no_instruction per issue waiting for its own code (ncu, r2_v11).
    enum { kHint = 0, kDown = 1, kUp = 2, kBisect = 3, kSizes = 4 };
    const bool searching = !(F.fixed_qindex >= 0 || F.fixed_per_slice) && target >= 0;
    int phase = kSizes;
    const int hint = own_hint < 0 ? shared_hint : own_hint;
    if (F.fixed_qindex >= 0 || F.fixed_per_slice) {
        lo = hi = F.fixed_per_slice ? F.fixed_per_slice[gw] : F.fixed_qindex;  // sizes only, no search
    } else if (target < 0) {
        lo = hi;  // cannot fit: reported as the saturated index (the reference would not terminate)
    } else {
        phase = (hint > lo && hint < hi) ? kHint : kBisect;
    }
    int step = 1;
#pragma unroll 1
    while (true) {
        int q;
        if (phase == kHint) {
            q = hint;
        } else if (phase == kDown) {
            if (!(hi - step > lo)) { phase = kBisect; continue; }
            q = hi - step;
        } else if (phase == kUp) {
            if (!(lo + step < hi)) { phase = kBisect; continue; }
            q = lo + step - 1;
        } else if (phase == kBisect) {
            if (!(lo < hi)) { phase = kSizes; continue; }
            q = (lo + hi) >> 1;
        } else {
            if (ok_q == lo) { by = oy; bc1 = o1; bc2 = o2; break; }  // the answer is the last probe that fitted
            q = lo;
        }
        const bool ok = fits(q);
        if (phase == kHint) {
            if (ok) { hi = hint; phase = kDown; } else { lo = hint + 1; phase = kUp; }
            step = 1;
        } else if (phase == kDown) {
            if (ok) { hi -= step; step *= 2; } else { lo = hi - step + 1; phase = kBisect; }
        } else if (phase == kUp) {
            if (!ok) { lo += step; step *= 2; } else { hi = lo + step - 1; phase = kBisect; }
        } else if (phase == kBisect) {
            if (ok) hi = q; else lo = q + 1;
        } else {
            break;  // block sizes at the answer
        }
    }
    if (searching) {
        if (lane == 0 && own_hint < 0) g_qindex_hint[pic & (kHintSlots - 1)] = lo;
        own_hint = lo;
    }
    if (lane == 0) {
        qindex_out[gw] = lo;
        if (bits_out) {
            bits_out[3 * gw + 0] = by;
            bits_out[3 * gw + 1] = bc1;
            bits_out[3 * gw + 2] = bc2;
        }
    }
  }  // slices of this warp
}

// ---------------------------------------------------------------------------
// slice packing
// ---------------------------------------------------------------------------

// OR `len` (<= 64) bits of `code` (right-aligned) into the big-endian bit string `out32` at
// absolute bit position `pos`; bits at or beyond `end` are dropped (1-bits past the end of a
// bounded block are not written: bitstream/io.py:457-470).
__device__ __forceinline__ void put_bits(uint32_t *__restrict__ out32, long long pos, unsigned long long code,
                                         int len, long long end)
{
    if (pos + len > end) {
        const long long drop = pos + len - end;
        if (drop >= len) return;
        code >>= drop;
        len -= (int)drop;
    }
    if (len <= 0) return;
    const long long w0 = pos >> 5;
    const int o = (int)(pos & 31);
    if (len <= 32) {  // the usual case: at most two words
        const unsigned long long t = (code & 0xffffffffull) << (64 - o - len);
        const uint32_t q0 = (uint32_t)(t >> 32), q1 = (uint32_t)t;
        if (q0) atomicOr(&out32[w0], __byte_perm(q0, 0, 0x0123));
        if (q1) atomicOr(&out32[w0 + 1], __byte_perm(q1, 0, 0x0123));
        return;
    }
    // 96-bit window [w0, w0+3): place the code so that its MSB sits at bit offset o
    const int sh = 96 - o - len;  // >= 0
    uint32_t p0 = 0, p1 = 0, p2 = 0;
    // value = code << sh as three 32-bit words (p0 most significant)
    if (sh >= 64) {
        p0 = (uint32_t)(code << (sh - 64));
    } else if (sh >= 32) {
        const unsigned long long t = code << (sh - 32);  // len + sh - 32 <= 64 bits
        p1 = (uint32_t)t;
        p0 = (uint32_t)(t >> 32);
    } else {
        p2 = (uint32_t)(code << sh);
        p1 = (uint32_t)(sh == 0 ? (code >> 32) : (code >> (32 - sh)));
        p0 = (uint32_t)(sh == 0 ? 0 : (code >> (64 - sh)));
    }
    if (p0) atomicOr(&out32[w0], __byte_perm(p0, 0, 0x0123));
    if (p1) atomicOr(&out32[w0 + 1], __byte_perm(p1, 0, 0x0123));
    if (p2) atomicOr(&out32[w0 + 2], __byte_perm(p2, 0, 0x0123));
}

// write_sint (bitstream/io.py:575): code bits right-aligned, and its length
__device__ __forceinline__ unsigned long long sint_code(int32_t v, int *len)
{
    if (v == 0) { *len = 1; return 1ull; }
    const uint32_t mag = v < 0 ? (uint32_t)(-(long long)v) : (uint32_t)v;
    if (mag < 0x7fffu) {  // at most 14 data bits: the code fits 30 bits, all in 32-bit arithmetic
        const uint32_t m32 = mag + 1;
        const int k32 = 31 - __clz(m32);
        uint32_t d32 = m32 - (1u << k32);
        d32 = (d32 | (d32 << 8)) & 0x00ff00ffu;
        d32 = (d32 | (d32 << 4)) & 0x0f0f0f0fu;
        d32 = (d32 | (d32 << 2)) & 0x33333333u;
        d32 = (d32 | (d32 << 1)) & 0x55555555u;
        *len = 2 * k32 + 2;
        return (unsigned long long)((((d32 << 1) | 1u) << 1) | (v < 0 ? 1u : 0u));
    }
    const unsigned long long m = (unsigned long long)mag + 1;
    const int k = 63 - __clzll((long long)m);
    unsigned long long d = m - (1ull << k);  // k data bits
    // spread: bit i of d -> bit 2i
    d = (d | (d << 16)) & 0x0000ffff0000ffffull;
    d = (d | (d << 8)) & 0x00ff00ff00ff00ffull;
    d = (d | (d << 4)) & 0x0f0f0f0f0f0f0f0full;
    d = (d | (d << 2)) & 0x3333333333333333ull;
    d = (d | (d << 1)) & 0x5555555555555555ull;
    *len = 2 * k + 2;
    return (((d << 1) | 1ull) << 1) | (v < 0 ? 1ull : 0ull);
}

// Packs one bounded block: values in bitstream order at absolute bit `start`, block end `end`.
// The coefficients of the next 32 values are fetched while the current ones are coded.
template <bool INTERLEAVED>
__device__ void pack_block(const SliceBands &T, int nb, const int32_t *__restrict__ c1,
                           const int32_t *__restrict__ c2, int ncoef, const uint2 *qtab, uint32_t *__restrict__ out32,
                           long long start, long long end, int lane, const uint32_t *map)
{
    // map: the CTA's coefficient -> (band, offset) table for this component when the slice has the
    // reference rectangles (build_fit_map), else nullptr (band search and division per coefficient)
    auto fetch = [&](int idx, int *band) -> int32_t {
        if (idx >= ncoef) { *band = 0; return 0; }
        const int j = INTERLEAVED ? (idx >> 1) : idx;
        int off;
        if (map) {
            const uint32_t e = map[j];
            *band = (int)(e >> 24);
            off = T.base[*band] + (int)(e & 0xffffffu);
        } else {
            off = locate_coeff(T, nb, j, band);
        }
        return __ldg(((INTERLEAVED && (idx & 1)) ? c2 : c1) + off);
    };
    long long pos = start;
    int bn;
    int32_t cn = fetch(lane, &bn);
    for (int i0 = 0; i0 < ncoef && pos < end; i0 += 32) {
        const int idx = i0 + lane;
        const int32_t c = cn;
        const int b = bn;
        cn = fetch(idx + 32, &bn);
        int len = 0;
        unsigned long long code = 0;
        if (idx < ncoef) {
            const uint32_t q = staged_quant_mag(qtab, b, c);
            code = sint_code(c < 0 ? -(int32_t)q : (int32_t)q, &len);
        }
        int pre = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int o = __shfl_up_sync(0xffffffffu, pre, d);
            if (lane >= d) pre += o;
        }
        if (len) put_bits(out32, pos + pre - len, code, len, end);
        pos += __shfl_sync(0xffffffffu, pre, 31);
    }
}

__global__ void __launch_bounds__(kQWarps * 32)
pack_slices_kernel(DevParams P, FitArgs F, const int32_t *__restrict__ coeffs, int npictures,
                   const int32_t *__restrict__ qindex_in, const int32_t *__restrict__ bits_in,
                   uint8_t *__restrict__ out, long long out_stride)
{
    __shared__ SliceBands tabs[kQWarps][2];
    __shared__ uint2 qtabs[kQWarps][kMaxBands];  // {qf saturated to 2^32-1, floor((2^32-1) / qf)} per band
    __shared__ FitMap M;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nslices = P.slices_x * P.slices_y;
    const long long total_slices = (long long)npictures * nslices;
    build_fit_map(P, M, tabs[0], F.use_map != 0, kStageMax);

  // persistent CTAs, a contiguous run of slices per warp (as in quantize_fit_kernel)
  const long long nwarps_total = (long long)gridDim.x * kQWarps;
  const long long run = (total_slices + nwarps_total - 1) / nwarps_total;
  const long long gw0 = ((long long)blockIdx.x * kQWarps + warp) * run;
  for (long long gw = gw0; gw < min(gw0 + run, total_slices); gw++) {
    const int pic = (int)(gw / nslices);
    const int n = (int)(gw - (long long)pic * nslices);
    const int sy = n / P.slices_x, sx = n - sy * P.slices_x;
    const int32_t *pc = coeffs + (size_t)pic * P.pic_coeffs;
    SliceBands &TY = tabs[warp][0];
    SliceBands &TC = tabs[warp][1];
    const uint2 *qtab = qtabs[warp];
    __syncwarp();
    const int ny = build_slice_bands(P, 0, sx, sy, TY, lane);
    const int nc = build_slice_bands(P, 1, sx, sy, TC, lane);
    const bool mapped = matches_fit_map(P, M, TY, TC, lane);
    const uint32_t *map_y = mapped ? M.e : nullptr, *map_c = mapped ? M.e + ny : nullptr;
    const int qindex = qindex_in[gw];
    stage_pack_quantisers(TY, P.nb, qindex, qtabs[warp], lane);  // quant matrices are identical for the components
    uint32_t *out32 = reinterpret_cast<uint32_t *>(out);  // out is 4-byte aligned; positions absolute
    const long long pic_bit0 = 8 * (long long)pic * out_stride;

    if (P.is_ld) {
        // ld_slice (bitstream/vc2.py:727): 7-bit qindex, slice_y_length, luma block, chroma block
        const long long s0 = ld_slice_offset(P, n), s1 = ld_slice_offset(P, n + 1);
        const long long sb = s1 - s0;
        const int length_bits = ilog2_ceil(8 * sb - 7);
        const long long y_len = bits_in[3 * gw + 0];
        long long pos = pic_bit0 + 8 * s0;
        const long long end = pic_bit0 + 8 * s1;
        if (lane == 0) {
            put_bits(out32, pos, (unsigned long long)qindex, 7, end);
            put_bits(out32, pos + 7, (unsigned long long)y_len, length_bits, end);
        }
        pos += 7 + length_bits;
        pack_block<false>(TY, P.nb, pc + P.comp_off[0], nullptr, ny, qtab, out32, pos, pos + y_len, lane, map_y);
        pos += y_len;
        pack_block<true>(TC, P.nb, pc + P.comp_off[1], pc + P.comp_off[2], 2 * nc, qtab, out32, pos, end, lane, map_c);
    } else {
        // hq_slice (bitstream/vc2.py:798): prefix bytes (zero), qindex, then per component a
        // length byte and a bounded block of 8*scaler*length bits (make_hq_slice :456)
        const long long scaler = P.slice_size_scaler;
        const long long den = (long long)nslices * scaler;
        const long long align = 8 * scaler;
        const long long ly = (bits_in[3 * gw + 0] + align - 1) / align;
        const long long l1 = (bits_in[3 * gw + 1] + align - 1) / align;
        long long off, l2;
        if (F.lossless_offset) {  // every length field minimal (make_hq_slice total_length=None, :456)
            off = F.lossless_offset[(long long)pic * (nslices + 1) + n];
            l2 = (bits_in[3 * gw + 2] + align - 1) / align;
        } else {
            off = (long long)n * (P.slice_prefix_bytes + 4) + scaler * ((n * F.total_coeff_bytes) / den);
            l2 = hq_total_length(P, F, n) - ly - l1;
        }
        long long pos = pic_bit0 + 8 * (off + P.slice_prefix_bytes);
        const long long big = 0x7fffffffffffffffLL;
        if (lane == 0) {
            put_bits(out32, pos, (unsigned long long)qindex, 8, big);
            put_bits(out32, pos + 8, (unsigned long long)ly, 8, big);
        }
        pos += 16;
        pack_block<false>(TY, P.nb, pc + P.comp_off[0], nullptr, ny, qtab, out32, pos, pos + align * ly, lane, map_y);
        pos += align * ly;
        if (lane == 0) put_bits(out32, pos, (unsigned long long)l1, 8, big);
        pos += 8;
        pack_block<false>(TC, P.nb, pc + P.comp_off[1], nullptr, nc, qtab, out32, pos, pos + align * l1, lane, map_c);
        pos += align * l1;
        if (lane == 0) put_bits(out32, pos, (unsigned long long)l2, 8, big);
        pos += 8;
        pack_block<false>(TC, P.nb, pc + P.comp_off[2], nullptr, nc, qtab, out32, pos, pos + align * l2, lane, map_c);
    }
  }  // slices of this warp
}

}  // namespace vc2b

#include "pack_lanes.cuh"

namespace vc2b {

// ---------------------------------------------------------------------------
// whole-picture quantize_coeffs (encoder/pictures.py:251): forward_quant of every coefficient at
// max(qindex of its slice - quant matrix value, 0); output in the coefficient layout
// ---------------------------------------------------------------------------
__global__ void quantize_slices_kernel(DevParams P, const int32_t *__restrict__ coeffs, int npictures,
                                       const int32_t *__restrict__ qindex, int fixed_qindex, int32_t *__restrict__ out)
{
    const int c = blockIdx.y;
    const int pic = blockIdx.z;
    const int32_t *src = coeffs + (size_t)pic * P.pic_coeffs + P.comp_off[c];
    int32_t *dst = out + (size_t)pic * P.pic_coeffs + P.comp_off[c];
    const int32_t *qi = qindex ? qindex + (size_t)pic * P.slices_x * P.slices_y : nullptr;
    for (int b = 0; b < P.nb; b++) {
        const Band B = P.band[c][b];
        const long long n = (long long)B.w * B.h;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
            const int y = (int)(i / B.w), x = (int)(i - (long long)y * B.w);
            // the slice whose rectangle holds (x, y): largest s with (w * s) // slices <= x (slice_sizes.py:108)
            const int sx = (int)((((long long)x + 1) * P.slices_x - 1) / B.w);
            const int sy = (int)((((long long)y + 1) * P.slices_y - 1) / B.h);
            const int q = qi ? qi[sy * P.slices_x + sx] : fixed_qindex;
            const int32_t v = src[B.off + i];
            const uint32_t m = v < 0 ? (uint32_t)(-(long long)v) : (uint32_t)v;
            const uint32_t r = forward_quant_mag(m, quant_factor_u64(max(q - B.qm, 0)));
            dst[B.off + i] = v < 0 ? -(int32_t)r : (int32_t)r;
        }
    }
}

// ---------------------------------------------------------------------------
// make_transform_data_hq_lossless (encoder/pictures.py:528): the slices of a picture follow each other
// with minimal length fields, so a slice's offset is a prefix sum over the sizes of the slices before it
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
hq_lossless_layout_kernel(DevParams P, const int32_t *__restrict__ bits, int npictures, uint32_t *__restrict__ offset,
                          int32_t *__restrict__ max_length)
{
    __shared__ unsigned long long warp_sum[32];
    __shared__ unsigned long long carry;
    __shared__ int smax;
    const int pic = blockIdx.x;
    const int nslices = P.slices_x * P.slices_y;
    const long long align = 8LL * P.slice_size_scaler;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry = 0; smax = 0; }
    __syncthreads();
    int mymax = 0;
    for (int n0 = 0; n0 < nslices; n0 += blockDim.x) {
        const int n = n0 + threadIdx.x;
        unsigned long long sz = 0;
        if (n < nslices) {
            const int32_t *b = bits + 3 * ((size_t)pic * nslices + n);
            long long len = 0;
            for (int k = 0; k < 3; k++) {
                const long long l = (b[k] + align - 1) / align;
                mymax = max(mymax, (int)min(l, 0x7fffffffLL));
                len += l;
            }
            sz = (unsigned long long)(P.slice_prefix_bytes + 4 + P.slice_size_scaler * len);
        }
        unsigned long long pre = sz;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, pre, d);
            if (lane >= d) pre += o;
        }
        if (lane == 31) warp_sum[warp] = pre;
        __syncthreads();
        unsigned long long base = carry;
        for (int w = 0; w < warp; w++) base += warp_sum[w];
        if (n < nslices) offset[(size_t)pic * (nslices + 1) + n] = (uint32_t)(base + pre - sz);
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = base + pre;
        __syncthreads();
    }
    if (threadIdx.x == 0) offset[(size_t)pic * (nslices + 1) + nslices] = (uint32_t)carry;
    atomicMax(&smax, mymax);
    __syncthreads();
    if (threadIdx.x == 0 && max_length) max_length[pic] = smax;
}

// ---------------------------------------------------------------------------
// quantize_to_fit (encoder/pictures.py:295) on explicit coefficient lists: the function-level drop-in
// (one CTA; the lists of one slice).  result[0] = qindex, result[1 + s] = calculate_coeffs_bits of set s.
// ---------------------------------------------------------------------------
constexpr int kListSets = 8;
struct ListArgs {
    int nsets;
    int set_off[kListSets + 1];
    long long target, align;
    int minimum_qindex, max_qindex;
};

__global__ void __launch_bounds__(256)
quantize_fit_lists_kernel(ListArgs A, const int32_t *__restrict__ values, const int32_t *__restrict__ qm,
                          int32_t *__restrict__ result, int32_t *__restrict__ quantised)
{
    __shared__ long long s_bits[kListSets];
    __shared__ int s_last[kListSets];
    __shared__ int s_done;
    const int n = A.set_off[A.nsets];
    for (int q = A.minimum_qindex;; q++) {
        if (threadIdx.x < kListSets) { s_bits[threadIdx.x] = 0; s_last[threadIdx.x] = -1; }
        __syncthreads();
        for (int s = 0; s < A.nsets; s++) {
            long long sum = 0;
            int last = -1;
            for (int i = A.set_off[s] + threadIdx.x; i < A.set_off[s + 1]; i += blockDim.x) {
                const int32_t v = values[i];
                const uint32_t m = v < 0 ? (uint32_t)(-(long long)v) : (uint32_t)v;
                const uint32_t r = forward_quant_mag(m, quant_factor_u64(max(q - qm[i], 0)));
                sum += sint_length(r);
                if (r) last = i;
            }
            atomicAdd((unsigned long long *)&s_bits[s], (unsigned long long)sum);
            atomicMax(&s_last[s], last);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long total = 0;
            for (int s = 0; s < A.nsets; s++) {
                // trailing zeros are free (calculate_coeffs_bits :217)
                const long long b = s_last[s] < 0 ? 0 : s_bits[s] - (A.set_off[s + 1] - 1 - s_last[s]);
                s_bits[s] = b;
                total += ((b + A.align - 1) / A.align) * A.align;
            }
            s_done = (total <= A.target || q >= A.max_qindex) ? 1 : 0;
        }
        __syncthreads();
        if (s_done) {
            if (threadIdx.x == 0) {
                result[0] = q;
                for (int s = 0; s < A.nsets; s++) result[1 + s] = (int32_t)s_bits[s];
            }
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int32_t v = values[i];
                const uint32_t m = v < 0 ? (uint32_t)(-(long long)v) : (uint32_t)v;
                const uint32_t r = forward_quant_mag(m, quant_factor_u64(max(q - qm[i], 0)));
                quantised[i] = v < 0 ? -(int32_t)r : (int32_t)r;
            }
            return;
        }
        __syncthreads();
    }
}

static int max_qm(const DevParams &P)
{
    int m = 0;
    for (int b = 0; b < P.nb; b++) m = max(m, P.band[0][b].qm);
    return m;
}

}  // namespace vc2b

using namespace vc2b;

static inline unsigned grid_for(long long n, int threads)
{
    long long b = (n + threads - 1) / threads;
    const long long cap = 148LL * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" {

int vc2b_inverse_quant(const int32_t *d_in, const int32_t *d_qi, int32_t *d_out, int64_t n, void *stream)
{
    if (n <= 0) return 0;
    inverse_quant_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(d_in, d_qi, d_out, n);
    VC2B_CHECK_LAUNCH();
    return 0;
}

int vc2b_forward_quant(const int32_t *d_in, const int32_t *d_qi, int32_t *d_out, int64_t n, void *stream)
{
    if (n <= 0) return 0;
    forward_quant_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(d_in, d_qi, d_out, n);
    VC2B_CHECK_LAUNCH();
    return 0;
}

int vc2b_clip_offset(int32_t *d_data, int64_t n, int depth, int mode, void *stream)
{
    if (n <= 0) return 0;
    if (depth < 1 || depth > 30) return VC2B_E_BAD_PARAMS;
    clip_offset_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(d_data, n, depth, mode);
    VC2B_CHECK_LAUNCH();
    return 0;
}

// shared by vc2b_quantize_to_fit (search) and vc2b_slice_bits (sizes at a given index)
static int launch_fit(const vc2b_params *p, const int32_t *d_coeffs, int npictures, int64_t picture_bytes,
                      int minimum_qindex, int fixed_qindex, const int32_t *d_fixed, int32_t *d_qindex,
                      int32_t *d_slice_bits, void *stream)
{
    DevParams P;
    int st = make_dev_params(p, &P);
    if (st) return st;
    if (npictures <= 0) return 0;
    FitArgs F;
    memset(&F, 0, sizeof(F));
    const bool search = fixed_qindex < 0 && !d_fixed;
    const long long nslices = (long long)P.slices_x * P.slices_y;
    F.total_coeff_bytes = picture_bytes - 4 * nslices;
    if (search && !P.is_ld && F.total_coeff_bytes < 0) return VC2B_E_BAD_PARAMS;  // InsufficientHQPictureBytesError
    if (search && !P.is_ld) {
        // every length field must fit 8 bits (get_safe_lossy_hq_slice_size_scaler, encoder/pictures.py:586)
        const long long den = nslices * P.slice_size_scaler;
        if ((F.total_coeff_bytes + den - 1) / den > 255) return VC2B_E_BAD_PARAMS;
    }
    F.minimum_qindex = minimum_qindex < 0 ? 0 : minimum_qindex;
    F.fixed_qindex = search ? -1 : (fixed_qindex < 0 ? 0 : fixed_qindex);
    F.fixed_per_slice = d_fixed;
    F.max_qm = max_qm(P);
    // staging area: the largest slice (every band rectangle is at most ceil(w / slices_x) x ceil(h / slices_y))
    long long cap = 0;
    for (int c = 0; c < 3; c++)
        for (int b = 0; b < P.nb; b++)
            cap += (long long)((P.band[c][b].w + P.slices_x - 1) / P.slices_x) *
                   ((P.band[c][b].h + P.slices_y - 1) / P.slices_y);
    F.stage_cap = cap <= kStageMax ? (int)((cap + 15) & ~15LL) : 16;  // larger slices are searched unstaged
    F.use_map = 1;
    F.uniform = 1;
    for (int c = 0; c < 3; c++)
        for (int b = 0; b < P.nb; b++) {
            if ((long long)P.band[c][b].w * P.band[c][b].h >= (1 << 24)) F.use_map = 0;
            if (P.band[c][b].w % P.slices_x || P.band[c][b].h % P.slices_y || P.band[c][b].w < P.slices_x ||
                P.band[c][b].h < P.slices_y)
                F.uniform = 0;
        }
    if (F.total_coeff_bytes < 0 || (P.is_ld && (P.sb_num < 0 || P.sb_den <= 0))) F.uniform = 0;
    const size_t smem = kQWarps * slice_stage_bytes(F.stage_cap);
    {   // static + dynamic shared memory can exceed the 48 KB default for the largest staged slices;
        // the attribute is per device, so it is set on every call (cheap) rather than once per process
        const cudaError_t attr = cudaFuncSetAttribute(quantize_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      (int)(kQWarps * slice_stage_bytes(kStageMax)));
        if (attr != cudaSuccess) return (int)attr;
    }
    const long long warps = npictures * nslices;
    // persistent CTAs: each warp walks a contiguous run of slices (the reference map is built once per CTA)
    long long nblk = (warps + kQWarps - 1) / kQWarps;
    if (nblk > 148LL * 16) nblk = 148LL * 16;
    quantize_fit_kernel<<<(unsigned)nblk, kQWarps * 32, smem, (cudaStream_t)stream>>>(
        P, F, d_coeffs, npictures, d_qindex, d_slice_bits);
    VC2B_CHECK_LAUNCH();
    return 0;
}

int vc2b_quantize_to_fit(const vc2b_params *p, const int32_t *d_coeffs, int npictures, int64_t picture_bytes,
                         int minimum_qindex, int32_t *d_qindex, int32_t *d_slice_bits, void *stream)
{
    return launch_fit(p, d_coeffs, npictures, picture_bytes, minimum_qindex, -1, nullptr, d_qindex, d_slice_bits, stream);
}

int vc2b_slice_bits(const vc2b_params *p, const int32_t *d_coeffs, int npictures, int qindex,
                    const int32_t *d_qindex_in, int32_t *d_qindex_out, int32_t *d_slice_bits, void *stream)
{
    if (!d_slice_bits || !d_qindex_out) return VC2B_E_BAD_PARAMS;
    return launch_fit(p, d_coeffs, npictures, 0, 0, qindex < 0 ? 0 : qindex, d_qindex_in, d_qindex_out, d_slice_bits,
                      stream);
}

int vc2b_quantize_slices(const vc2b_params *p, const int32_t *d_coeffs, int npictures, const int32_t *d_qindex,
                         int qindex, int32_t *d_out, void *stream)
{
    DevParams P;
    int st = make_dev_params(p, &P);
    if (st) return st;
    if (npictures <= 0) return 0;
    long long biggest = 1;
    for (int b = 0; b < P.nb; b++) biggest = max(biggest, (long long)P.band[0][b].w * P.band[0][b].h);
    dim3 grid(grid_for(biggest, 256), 3, npictures);
    quantize_slices_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P, d_coeffs, npictures, d_qindex, qindex < 0 ? 0 : qindex,
                                                                   d_out);
    VC2B_CHECK_LAUNCH();
    return 0;
}

int vc2b_hq_lossless_layout(const vc2b_params *p, const int32_t *d_slice_bits, int npictures,
                            uint32_t *d_slice_offset, int32_t *d_max_length, void *stream)
{
    DevParams P;
    int st = make_dev_params(p, &P);
    if (st) return st;
    if (P.is_ld) return VC2B_E_BAD_PARAMS;
    if (npictures <= 0) return 0;
    hq_lossless_layout_kernel<<<npictures, 1024, 0, (cudaStream_t)stream>>>(P, d_slice_bits, npictures, d_slice_offset,
                                                                           d_max_length);
    VC2B_CHECK_LAUNCH();
    return 0;
}

int vc2b_quantize_to_fit_lists(const int32_t *d_values, const int32_t *d_quant_matrix_values, const int32_t *set_offsets,
                               int nsets, int64_t target_size, int64_t align_bits, int minimum_qindex,
                               int32_t *d_result, int32_t *d_quantised, void *stream)
{
    if (nsets < 1 || nsets > kListSets || !set_offsets || target_size < 0 || align_bits < 1) return VC2B_E_BAD_PARAMS;
    ListArgs A;
    memset(&A, 0, sizeof(A));
    A.nsets = nsets;
    for (int i = 0; i <= nsets; i++) {
        if (set_offsets[i] < 0 || (i && set_offsets[i] < set_offsets[i - 1])) return VC2B_E_BAD_PARAMS;
        A.set_off[i] = set_offsets[i];
    }
    A.target = target_size;
    A.align = align_bits;
    A.minimum_qindex = minimum_qindex < 0 ? 0 : minimum_qindex;
    // at this index every int32 value quantises to zero whatever the matrix value (quant_factor >= 2^33)
    A.max_qindex = A.minimum_qindex > 4 * 33 + 1024 ? A.minimum_qindex : 4 * 33 + 1024;
    quantize_fit_lists_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(A, d_values, d_quant_matrix_values, d_result,
                                                                   d_quantised);
    VC2B_CHECK_LAUNCH();
    return 0;
}

int64_t vc2b_packed_bytes(const vc2b_params *p, int64_t picture_bytes)
{
    DevParams P;
    int st = make_dev_params(p, &P);
    if (st) return st;
    const long long nslices = (long long)P.slices_x * P.slices_y;
    long long n;
    if (P.is_ld) {
        n = ld_slice_offset(P, nslices);
    } else {
        const long long tcb = picture_bytes - 4 * nslices;
        if (tcb < 0) return VC2B_E_BAD_PARAMS;
        const long long den = nslices * P.slice_size_scaler;
        n = nslices * (P.slice_prefix_bytes + 4) + P.slice_size_scaler * ((nslices * tcb) / den);
    }
    return n;
}

static int launch_pack(const vc2b_params *p, const int32_t *d_coeffs, int npictures, int64_t picture_bytes,
                       const int32_t *d_qindex, const int32_t *d_slice_bits, const uint32_t *d_lossless_offset,
                       uint8_t *d_out, int64_t out_stride, void *stream)
{
    DevParams P;
    int st = make_dev_params(p, &P);
    if (st) return st;
    if (npictures <= 0) return 0;
    if (d_lossless_offset) {
        if (P.is_ld || out_stride < 4 || (out_stride & 3)) return VC2B_E_BAD_PARAMS;
    } else {
        const int64_t need = vc2b_packed_bytes(p, picture_bytes);
        if (need < 0) return (int)need;
        if (out_stride < need || (out_stride & 3)) return VC2B_E_BAD_PARAMS;
    }
    FitArgs F;
    memset(&F, 0, sizeof(F));
    const long long nslices = (long long)P.slices_x * P.slices_y;
    F.total_coeff_bytes = picture_bytes - 4 * nslices;
    F.minimum_qindex = 0;
    F.max_qm = max_qm(P);
    F.stage_cap = 0;
    F.use_map = 1;
    F.fixed_qindex = -1;
    F.lossless_offset = d_lossless_offset;
    for (int c = 0; c < 3; c++)
        for (int b = 0; b < P.nb; b++)
            if ((long long)P.band[c][b].w * P.band[c][b].h >= (1 << 24)) F.use_map = 0;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_out, 0, (size_t)out_stride * npictures, s);
    if (e != cudaSuccess) return (int)e;
    if (pack_lanes_eligible(P)) {
        // small slices that all have the same rectangles (every BASELINE configuration): one LANE per slice
        const long long tasks = (long long)npictures * ((nslices + 31) / 32);
        long long nblk = (tasks + kPkWarps - 1) / kPkWarps;
        if (nblk > 148LL * 2) nblk = 148LL * 2;  // persistent: two CTAs of 16 warps per SM
        e = cudaFuncSetAttribute(pack_lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPkSmemBytes);
        if (e != cudaSuccess) return (int)e;
        pack_lanes_kernel<<<(unsigned)nblk, kPkWarps * 32, kPkSmemBytes, s>>>(
            P, F, d_coeffs, npictures, d_qindex, d_slice_bits, d_out, out_stride);
        VC2B_CHECK_LAUNCH();
        return 0;
    }
    const long long warps = npictures * nslices;
    long long nblk = (warps + kQWarps - 1) / kQWarps;
    if (nblk > 148LL * 16) nblk = 148LL * 16;  // persistent CTAs (the reference map is built once per CTA)
    pack_slices_kernel<<<(unsigned)nblk, kQWarps * 32, 0, s>>>(
        P, F, d_coeffs, npictures, d_qindex, d_slice_bits, d_out, out_stride);
    VC2B_CHECK_LAUNCH();
    return 0;
}

int vc2b_pack_slices(const vc2b_params *p, const int32_t *d_coeffs, int npictures, int64_t picture_bytes,
                     const int32_t *d_qindex, const int32_t *d_slice_bits, uint8_t *d_out, int64_t out_stride,
                     void *stream)
{
    return launch_pack(p, d_coeffs, npictures, picture_bytes, d_qindex, d_slice_bits, nullptr, d_out, out_stride, stream);
}

int vc2b_pack_slices_lossless(const vc2b_params *p, const int32_t *d_coeffs, int npictures, const int32_t *d_qindex,
                              const int32_t *d_slice_bits, const uint32_t *d_slice_offset, uint8_t *d_out,
                              int64_t out_stride, void *stream)
{
    if (!d_slice_offset) return VC2B_E_BAD_PARAMS;
    return launch_pack(p, d_coeffs, npictures, 0, d_qindex, d_slice_bits, d_slice_offset, d_out, out_stride, stream);
}

}  // extern "C"
