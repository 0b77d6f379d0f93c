// params.cu -- host-side geometry and introspection entry points of libvc2b200.
//
// Restates pseudocode/slice_sizes.py:53-92 (subband_width / subband_height) and
// the subband ordering of decoder/transform_data_syntax.py:176-189,230-243.
#include <map>
#include <mutex>

#include "lifting.cuh"

namespace vc2b {

unsigned long long g_launch_count = 0;

static int32_t subband_width(const vc2b_params *p, int level, int comp)
{
    int32_t w = comp == 0 ? p->luma_width : p->color_diff_width;
    int32_t scale_w = 1 << (p->dwt_depth_ho + p->dwt_depth);
    int32_t pw = scale_w * ((w + scale_w - 1) / scale_w);
    if (level == 0) return pw / scale_w;
    return pw / (1 << (p->dwt_depth_ho + p->dwt_depth - level + 1));
}

static int32_t subband_height(const vc2b_params *p, int level, int comp)
{
    int32_t h = comp == 0 ? p->luma_height : p->color_diff_height;
    int32_t scale_h = 1 << p->dwt_depth;
    int32_t ph = scale_h * ((h + scale_h - 1) / scale_h);
    if (level == 0 || level <= p->dwt_depth_ho) return ph / scale_h;
    return ph / (1 << (p->dwt_depth_ho + p->dwt_depth - level + 1));
}

int make_dev_params(const vc2b_params *p, DevParams *d)
{
    if (!p || !d) return VC2B_E_BAD_PARAMS;
    if (p->luma_width < 1 || p->luma_height < 1 || p->color_diff_width < 1 ||
        p->color_diff_height < 1)
        return VC2B_E_BAD_PARAMS;
    if (p->luma_depth < 1 || p->luma_depth > 16 || p->color_diff_depth < 1 ||
        p->color_diff_depth > 16)
        return VC2B_E_UNSUPPORTED;
    if (p->wavelet_index < 0 || p->wavelet_index > 6 || p->wavelet_index_ho < 0 ||
        p->wavelet_index_ho > 6)
        return VC2B_E_BAD_PARAMS;
    if (p->dwt_depth < 0 || p->dwt_depth_ho < 0 || p->dwt_depth + p->dwt_depth_ho + 1 > VC2B_MAX_LEVELS)
        return VC2B_E_BAD_PARAMS;
    int nb = 1 + p->dwt_depth_ho + 3 * p->dwt_depth;
    if (nb > kMaxBands) return VC2B_E_UNSUPPORTED;
    if (p->slices_x < 1 || p->slices_y < 1) return VC2B_E_BAD_PARAMS;
    if (p->is_ld) {
        if (p->slice_bytes_denominator < 1 || p->slice_bytes_numerator < 0) return VC2B_E_BAD_PARAMS;
    } else {
        if (p->slice_prefix_bytes < 0 || p->slice_size_scaler < 1) return VC2B_E_BAD_PARAMS;
    }

    memset(d, 0, sizeof(*d));
    d->wavelet_index = p->wavelet_index;
    d->wavelet_index_ho = p->wavelet_index_ho;
    d->dwt_depth = p->dwt_depth;
    d->dwt_depth_ho = p->dwt_depth_ho;
    d->slices_x = p->slices_x;
    d->slices_y = p->slices_y;
    d->is_ld = p->is_ld ? 1 : 0;
    d->slice_prefix_bytes = p->slice_prefix_bytes;
    d->slice_size_scaler = p->slice_size_scaler;
    d->sb_num = p->slice_bytes_numerator;
    d->sb_den = p->slice_bytes_denominator;
    d->nb = nb;
    int top = p->dwt_depth + p->dwt_depth_ho;
    int64_t coeff_total = 0, sample_total = 0;
    for (int c = 0; c < 3; c++) {
        d->width[c] = c == 0 ? p->luma_width : p->color_diff_width;
        d->height[c] = c == 0 ? p->luma_height : p->color_diff_height;
        d->depth[c] = c == 0 ? p->luma_depth : p->color_diff_depth;
        int b = 0;
        int64_t off = 0;
        for (int lv = 0; lv <= top; lv++) {
            int no = (lv == 0) ? 1 : (lv <= p->dwt_depth_ho ? 1 : 3);
            for (int o = 0; o < no; o++, b++) {
                Band &B = d->band[c][b];
                B.w = subband_width(p, lv, c);
                B.h = subband_height(p, lv, c);
                B.level = lv;
                B.slot = (lv == 0) ? 0 : o + 1;
                B.off = (int32_t)off;
                B.qm = p->quant_matrix[lv][B.slot];
                off += (int64_t)B.w * B.h;
            }
        }
        d->pw[c] = (top == 0) ? d->band[c][0].w : 2 * d->band[c][nb - 1].w;
        d->ph[c] = (p->dwt_depth == 0) ? d->band[c][0].h : 2 * d->band[c][nb - 1].h;
        if (off != (int64_t)d->pw[c] * d->ph[c]) return VC2B_E_BAD_PARAMS;
        d->comp_off[c] = (int32_t)coeff_total;
        d->comp_count[c] = (int32_t)off;
        coeff_total += off;
        d->pic_off[c] = (int32_t)sample_total;
        sample_total += (int64_t)d->width[c] * d->height[c];
        if (coeff_total > 0x7fffffffLL || sample_total > 0x7fffffffLL) return VC2B_E_UNSUPPORTED;
    }
    d->pic_coeffs = (int32_t)coeff_total;
    d->pic_samples = (int32_t)sample_total;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// int32 envelope of the transforms.  The reference computes in unbounded integers; the kernels
// hold samples in int32 (accumulating in int64 only for the S >= 8 stages).  These bounds
// propagate the largest possible magnitude through every level using the L-infinity operator
// norms of the lifting passes and report the worst value that must fit 31 bits.
// ---------------------------------------------------------------------------------------------
struct PassGain {
    double out;    // max abs row sum of the whole 1-D pass (L-infinity operator norm)
    double worst;  // largest norm reached after any stage, or by an int32 accumulator
};

// Norms of the linear part of oned_synthesis / oned_analysis, measured on the actual lifting
// operator (including the edge clamps) for several array lengths.
static PassGain pass_gain(int filter, bool analysis)
{
    const FilterDef f = get_filter(filter);
    PassGain g = {1.0, 1.0};
    const int lens[5] = {2, 4, 8, 16, 64};
    for (int li = 0; li < 5; li++) {
        const int N = lens[li];
        double M[64][64];  // a local: host threads may call concurrently (ctypes releases the GIL)
        for (int r = 0; r < N; r++)
            for (int c = 0; c < N; c++) M[r][c] = r == c ? 1.0 : 0.0;
        for (int k = 0; k < f.nstages; k++) {
            const StageDef &st = f.st[analysis ? f.nstages - 1 - k : k];
            const int type = analysis ? analysis_type(st.type) : st.type;
            const bool even = (type == 1 || type == 2);
            const double sign = (type == 1 || type == 3) ? 1.0 : -1.0;
            for (int n = 0; n < N / 2; n++) {
                double acc[64];
                for (int c = 0; c < N; c++) acc[c] = 0.0;
                for (int i = st.D; i < st.L + st.D; i++) {
                    int pos = even ? 2 * (n + i) - 1 : 2 * (n + i);
                    if (even) { if (pos > N - 1) pos = N - 1; if (pos < 1) pos = 1; }
                    else { if (pos > N - 2) pos = N - 2; if (pos < 0) pos = 0; }
                    for (int c = 0; c < N; c++) acc[c] += st.taps[i - st.D] * M[pos][c];
                }
                double accnorm = 0.0;
                for (int c = 0; c < N; c++) accnorm += acc[c] < 0 ? -acc[c] : acc[c];
                if (st.S < kWideAccS && accnorm > g.worst) g.worst = accnorm;  // int32 accumulator
                const int t = even ? 2 * n : 2 * n + 1;
                double rown = 0.0;
                for (int c = 0; c < N; c++) {
                    M[t][c] += sign * acc[c] / (double)(1LL << st.S);
                    rown += M[t][c] < 0 ? -M[t][c] : M[t][c];
                }
                if (rown > g.worst) g.worst = rown;
                if (k == f.nstages - 1 && rown > g.out) g.out = rown;
            }
            if (k == f.nstages - 1) {  // rows the last stage did not touch
                for (int r = 0; r < N; r++) {
                    double rown = 0.0;
                    for (int c = 0; c < N; c++) rown += M[r][c] < 0 ? -M[r][c] : M[r][c];
                    if (rown > g.out) g.out = rown;
                }
            }
        }
    }
    if (g.out > g.worst) g.worst = g.out;
    return g;
}

// one 1-D pass over samples bounded by v: returns the new bound; rounding adds < 1 per stage,
// amplified by at most the pass's worst norm
static double apply_pass(const PassGain &g, double v, double *worst)
{
    const double slack = 4.0 * g.worst + 4.0;
    if (g.worst * v + slack > *worst) *worst = g.worst * v + slack;
    return g.out * v + slack;
}

// worst intermediate magnitude of the synthesis when every coefficient is bounded by m
static double idwt_worst(const vc2b_params *p, double m)
{
    const PassGain gv = pass_gain(p->wavelet_index, false), gh = pass_gain(p->wavelet_index_ho, false);
    double worst = m, ll = m;
    const int shift = get_filter(p->wavelet_index_ho).shift;
    const int top = p->dwt_depth + p->dwt_depth_ho;
    for (int n = 1; n <= top; n++) {
        double v = ll > m ? ll : m;
        if (n > p->dwt_depth_ho) v = apply_pass(gv, v, &worst);
        v = apply_pass(gh, v, &worst);
        if (shift > 0) v = (v + (double)(1 << (shift - 1))) / (double)(1 << shift) + 1.0;
        ll = v;
    }
    return worst;
}

// worst intermediate magnitude of the analysis of samples bounded by m
static double dwt_worst(const vc2b_params *p, double m)
{
    const PassGain gv = pass_gain(p->wavelet_index, true), gh = pass_gain(p->wavelet_index_ho, true);
    double worst = m, ll = m;
    const int shift = get_filter(p->wavelet_index_ho).shift;
    const int top = p->dwt_depth + p->dwt_depth_ho;
    for (int n = top; n >= 1; n--) {
        double v = ll * (double)(1 << shift);
        if (v > worst) worst = v;
        v = apply_pass(gh, v, &worst);
        if (n > p->dwt_depth_ho) v = apply_pass(gv, v, &worst);
        ll = v;
    }
    return worst;
}

// The bound only depends on the filters and depths; the bisection costs a millisecond or two of host
// time, which vc2b_unpack would otherwise pay on every call: cache it per transform (mutex: ctypes
// callers may be on several host threads).
static std::mutex g_safe_mutex;
static std::map<uint32_t, long long> g_safe_cache;

long long idwt_max_safe_coeff(const vc2b_params *p)
{
    const uint32_t key = ((uint32_t)(p->wavelet_index & 0xff) << 24) | ((uint32_t)(p->wavelet_index_ho & 0xff) << 16) |
                         ((uint32_t)(p->dwt_depth & 0xff) << 8) | (uint32_t)(p->dwt_depth_ho & 0xff);
    {
        std::lock_guard<std::mutex> lock(g_safe_mutex);
        auto it = g_safe_cache.find(key);
        if (it != g_safe_cache.end()) return it->second;
    }
    const double limit = 2147483647.0;
    long long lo = 0, hi = 0x7fffffffLL;  // largest m with idwt_worst(m) <= limit
    while (lo < hi) {
        long long mid = (lo + hi + 1) >> 1;
        if (idwt_worst(p, (double)mid) <= limit) lo = mid; else hi = mid - 1;
    }
    std::lock_guard<std::mutex> lock(g_safe_mutex);
    g_safe_cache[key] = lo;
    return lo;
}

bool dwt_is_safe(const vc2b_params *p)
{
    int depth = p->luma_depth > p->color_diff_depth ? p->luma_depth : p->color_diff_depth;
    return dwt_worst(p, (double)(1LL << (depth - 1))) <= 2147483647.0;
}

}  // namespace vc2b

using namespace vc2b;

extern "C" {

const char *vc2b_version(void) { return "vc2b200 0.1 (sm_100a)"; }

int64_t vc2b_launch_count(void) { return (int64_t)__atomic_load_n(&vc2b::g_launch_count, __ATOMIC_RELAXED); }

int vc2b_check_params(const vc2b_params *p)
{
    DevParams d;
    return make_dev_params(p, &d);
}

int vc2b_num_subbands(const vc2b_params *p)
{
    DevParams d;
    int st = make_dev_params(p, &d);
    return st ? st : d.nb;
}

int vc2b_subband_info(const vc2b_params *p, int comp, int band, int32_t *level, int32_t *slot,
                      int32_t *width, int32_t *height, int64_t *offset)
{
    DevParams d;
    int st = make_dev_params(p, &d);
    if (st) return st;
    if (comp < 0 || comp > 2 || band < 0 || band >= d.nb) return VC2B_E_BAD_PARAMS;
    const Band &B = d.band[comp][band];
    if (level) *level = B.level;
    if (slot) *slot = B.slot;
    if (width) *width = B.w;
    if (height) *height = B.h;
    if (offset) *offset = B.off;
    return 0;
}

int64_t vc2b_component_coeff_count(const vc2b_params *p, int comp)
{
    DevParams d;
    int st = make_dev_params(p, &d);
    if (st) return st;
    if (comp < 0 || comp > 2) return VC2B_E_BAD_PARAMS;
    return d.comp_count[comp];
}

int64_t vc2b_picture_coeff_count(const vc2b_params *p)
{
    DevParams d;
    int st = make_dev_params(p, &d);
    return st ? st : d.pic_coeffs;
}

int vc2b_padded_dims(const vc2b_params *p, int comp, int32_t *width, int32_t *height)
{
    DevParams d;
    int st = make_dev_params(p, &d);
    if (st) return st;
    if (comp < 0 || comp > 2) return VC2B_E_BAD_PARAMS;
    if (width) *width = d.pw[comp];
    if (height) *height = d.ph[comp];
    return 0;
}

int64_t vc2b_idwt_max_safe_coeff(const vc2b_params *p)
{
    DevParams d;
    int st = make_dev_params(p, &d);
    if (st) return st;
    return idwt_max_safe_coeff(p);
}

int64_t vc2b_picture_sample_count(const vc2b_params *p)
{
    DevParams d;
    int st = make_dev_params(p, &d);
    return st ? st : d.pic_samples;
}

}  // extern "C"
