// dwt.cu -- forward wavelet transform (analysis), the encoder mirror of idwt.cu.
//
// Replaces picture_encode / remove_offset_picture / forward_wavelet_transform /
// dwt_pad_addition / dwt / h_analysis / vh_analysis / oned_analysis
// (pseudocode/picture_encoding.py:45-286).
//
// One launch per level, top level first (picture_encoding.py:118-126).  A CTA owns a
// TW x TH tile of the level's INPUT; it stages tile + halo in shared memory -- on the first
// level straight from the uint16 picture with the offset removed (:264) and the last
// column/row replicated into the padding (:235-256) -- pre-shifts left by the filter bit
// shift (:176-181), lifts rows with wavelet_index_ho then columns with wavelet_index
// (:183-186, stages reversed and add/subtract swapped :203-223) and writes the de-interleaved
// LL / HL / LH / HH samples (:188-199).
#include "lifting.cuh"

namespace vc2b {

constexpr int kDwtThreads = 256;

struct AnaArgs {
    // input: int32 array of W x H (row stride W), or the uint16 picture on the first level
    const int32_t *in[3];
    long long in_pic_stride[3];
    const uint16_t *pic[3];
    long long pic_pic_stride;
    int32_t pic_w[3], pic_h[3], depth[3];
    int32_t from_u16;
    int32_t W[3], H[3];             // input (interleaved) dimensions of this level
    // output
    int32_t *ll[3];
    long long ll_pic_stride[3];
    int32_t *hb[3][3];
    long long hb_pic_stride;
    int32_t shift;
    int32_t tiles_x[3], tiles_y[3];
    int32_t ncomp;
};

}  // namespace vc2b

#include "dwt8.cuh"

namespace vc2b {

template <int FV, int FH, bool VERT, int TW, int TH>
__global__ void __launch_bounds__(kDwtThreads)
dwt_level_kernel(AnaArgs A)
{
    constexpr int HX = get_filter(FH).halo;
    constexpr int HY = VERT ? get_filter(FV).halo : 0;
    constexpr int TS = TW + 2 * HX + 1;
    extern __shared__ int32_t tile[];

    const int c = blockIdx.y;
    const int pic = blockIdx.z;
    const int tx = blockIdx.x % A.tiles_x[c];
    const int ty = blockIdx.x / A.tiles_x[c];
    if (ty >= A.tiles_y[c]) return;

    const int W = A.W[c], H = A.H[c];
    const int X0 = tx * TW, Y0 = ty * TH;
    const int gx0 = max(0, X0 - HX), gx1 = min(W, X0 + TW + HX);
    const int gy0 = max(0, Y0 - HY), gy1 = min(H, Y0 + TH + HY);
    const int rw = gx1 - gx0, rh = gy1 - gy0;
    const int shift = A.shift;

    if (A.from_u16) {
        const uint16_t *src = A.pic[c] + (long long)pic * A.pic_pic_stride;
        const int pw = A.pic_w[c], ph = A.pic_h[c];
        const int half = 1 << (A.depth[c] - 1);
        for (int i = threadIdx.x; i < rw * rh; i += kDwtThreads) {
            const int y = i / rw, x = i - y * rw;
            const int sy = min(gy0 + y, ph - 1), sx = min(gx0 + x, pw - 1);
            const int v = (int)__ldg(src + (long long)sy * pw + sx) - half;
            tile[y * TS + x] = v * (1 << shift);
        }
    } else {
        const int32_t *src = A.in[c] + (long long)pic * A.in_pic_stride[c];
        for (int i = threadIdx.x; i < rw * rh; i += kDwtThreads) {
            const int y = i / rw, x = i - y * rw;
            tile[y * TS + x] = __ldg(src + (long long)(gy0 + y) * W + gx0 + x) * (1 << shift);
        }
    }
    __syncthreads();

    lift_tile<FH, false, true>(tile, TS, rw, rh, gx0, W);
    if (VERT) lift_tile<FV, true, true>(tile, TS, rw, rh, gy0, H);

    const int ox1 = min(W, X0 + TW), oy1 = min(H, Y0 + TH);
    const int w2 = W >> 1;
    int32_t *ll = A.ll[c] + (long long)pic * A.ll_pic_stride[c];
    int32_t *b0 = A.hb[c][0] + (long long)pic * A.hb_pic_stride;
    if (VERT) {
        int32_t *b1 = A.hb[c][1] + (long long)pic * A.hb_pic_stride;
        int32_t *b2 = A.hb[c][2] + (long long)pic * A.hb_pic_stride;
        const int sw = (ox1 - X0) >> 1, sh = (oy1 - Y0) >> 1;
        for (int i = threadIdx.x; i < sw * sh; i += kDwtThreads) {
            const int sy = i / sw, sx = i - sy * sw;
            const int32_t *t = &tile[(Y0 - gy0 + 2 * sy) * TS + (X0 - gx0 + 2 * sx)];
            const long long g = (long long)((Y0 >> 1) + sy) * w2 + (X0 >> 1) + sx;
            ll[g] = t[0];
            b0[g] = t[1];
            b1[g] = t[TS];
            b2[g] = t[TS + 1];
        }
    } else {
        const int sw = (ox1 - X0) >> 1, sh = oy1 - Y0;
        for (int i = threadIdx.x; i < sw * sh; i += kDwtThreads) {
            const int sy = i / sw, sx = i - sy * sw;
            const int32_t *t = &tile[(Y0 - gy0 + sy) * TS + (X0 - gx0 + 2 * sx)];
            const long long g = (long long)(Y0 + sy) * w2 + (X0 >> 1) + sx;
            ll[g] = t[0];
            b0[g] = t[1];
        }
    }
}

// No transform: coefficients = picture - offset (padded).
__global__ void dwt_copy_kernel(AnaArgs A)
{
    const int c = blockIdx.y, pic = blockIdx.z;
    const int W = A.W[c], H = A.H[c];
    int32_t *ll = A.ll[c] + (long long)pic * A.ll_pic_stride[c];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)W * H;
         i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i - (long long)y * W);
        int v;
        if (A.from_u16) {
            const int sy = min(y, A.pic_h[c] - 1), sx = min(x, A.pic_w[c] - 1);
            v = (int)A.pic[c][(long long)pic * A.pic_pic_stride + (long long)sy * A.pic_w[c] + sx] -
                (1 << (A.depth[c] - 1));
        } else {
            v = A.in[c][(long long)pic * A.in_pic_stride[c] + i];
        }
        ll[i] = v;
    }
}

template <int FV, int FH, bool VERT>
static int launch_ana_t(const AnaArgs &A_in, int npic, cudaStream_t s)
{
    if constexpr (VERT && Ana8Cfg<FV, FH>::kSupported && (FV == FH)) {
        if (ana_is_wide_ok(A_in, npic)) return launch_ana_wide<FV, FH>(A_in, npic, s);
    }
    constexpr int HX = get_filter(FH).halo;
    constexpr int HY = VERT ? get_filter(FV).halo : 0;
    constexpr int TW = 128;
    constexpr int TH = (HY >= 8) ? 64 : 32;
    constexpr int TS = TW + 2 * HX + 1;
    constexpr size_t smem = sizeof(int32_t) * (size_t)TS * (TH + 2 * HY);
    AnaArgs A = A_in;
    int maxtiles = 0;
    for (int c = 0; c < A.ncomp; c++) {
        A.tiles_x[c] = (A.W[c] + TW - 1) / TW;
        A.tiles_y[c] = (A.H[c] + TH - 1) / TH;
        maxtiles = max(maxtiles, A.tiles_x[c] * A.tiles_y[c]);
    }
    auto kern = dwt_level_kernel<FV, FH, VERT, TW, TH>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid(maxtiles, A.ncomp, npic);
    kern<<<grid, kDwtThreads, smem, s>>>(A);
    VC2B_CHECK_LAUNCH();
    return 0;
}

template <int FV, bool VERT>
static int launch_ana_h(int fh, const AnaArgs &A, int npic, cudaStream_t s)
{
#ifdef VC2B_DEV_BUILD
    return VC2B_E_UNSUPPORTED;
#else
    switch (fh) {
    case 0: return launch_ana_t<FV, 0, VERT>(A, npic, s);
    case 1: return launch_ana_t<FV, 1, VERT>(A, npic, s);
    case 2: return launch_ana_t<FV, 2, VERT>(A, npic, s);
    case 3: return launch_ana_t<FV, 3, VERT>(A, npic, s);
    case 5: return launch_ana_t<FV, 5, VERT>(A, npic, s);
    default: return launch_ana_t<FV, 6, VERT>(A, npic, s);
    }
#endif
}

static int launch_ana(int fv, int fh, bool vert, const AnaArgs &A, int npic, cudaStream_t s)
{
    if (fv == 4) fv = 3;  // Haar with/without shift: same stages, shift is a runtime argument
    if (fh == 4) fh = 3;
#ifdef VC2B_DEV_BUILD
    if (fh != 0 || (vert && fv != 0)) return VC2B_E_UNSUPPORTED;
    return vert ? launch_ana_t<0, 0, true>(A, npic, s) : launch_ana_t<0, 0, false>(A, npic, s);
#endif
    if (!vert) return launch_ana_h<0, false>(fh, A, npic, s);
    switch (fv) {
    case 0: return launch_ana_h<0, true>(fh, A, npic, s);
    case 1: return launch_ana_h<1, true>(fh, A, npic, s);
    case 2: return launch_ana_h<2, true>(fh, A, npic, s);
    case 3: return launch_ana_h<3, true>(fh, A, npic, s);
    case 5: return launch_ana_h<5, true>(fh, A, npic, s);
    default: return launch_ana_h<6, true>(fh, A, npic, s);
    }
}

static int level_first_band(const DevParams &P, int n)
{
    return (n <= P.dwt_depth_ho) ? n : P.dwt_depth_ho + 1 + 3 * (n - P.dwt_depth_ho - 1);
}

// Scratch: the LL output of level n (n = top .. 2) feeds level n-1; ping-pong as in idwt.cu
// (region a holds the largest one: the LL band of the top level).
struct AnaScratch {
    long long size_a, size_b;
    long long comp_off_a[3], comp_off_b[3];
};

static void plan_ana_scratch(const DevParams &P, int c0, int c1, AnaScratch *sp)
{
    const int top = P.dwt_depth + P.dwt_depth_ho;
    sp->size_a = sp->size_b = 0;
    for (int c = c0; c < c1; c++) {
        long long a = 0, b = 0;
        for (int n = top; n >= 2; n--) {  // LL output of level n has the dims of its own bands
            const Band &B = P.band[c][level_first_band(P, n)];
            long long sz = (long long)B.w * B.h;
            if ((top - n) % 2 == 0) a = max(a, sz); else b = max(b, sz);
        }
        sp->comp_off_a[c] = sp->size_a;
        sp->comp_off_b[c] = sp->size_b;
        sp->size_a += a;
        sp->size_b += b;
    }
}

static int run_dwt(const DevParams &P, int c0, int c1, const uint16_t *pictures, const int32_t *raw_in,
                   int npic, int32_t *coeffs, long long coeff_pic_stride, int32_t *scratch,
                   const AnaScratch &sp, cudaStream_t s)
{
    const int top = P.dwt_depth + P.dwt_depth_ho;
    const int ncomp = c1 - c0;
    const long long scratch_pic = sp.size_a + sp.size_b;
    AnaArgs A;
    memset(&A, 0, sizeof(A));
    A.ncomp = ncomp;
    A.hb_pic_stride = coeff_pic_stride;
    A.pic_pic_stride = P.pic_samples;
    A.shift = get_filter(P.wavelet_index_ho).shift;
    for (int i = 0; i < ncomp; i++) {
        const int c = c0 + i;
        A.pic_w[i] = P.width[c];
        A.pic_h[i] = P.height[c];
        A.depth[i] = P.depth[c];
        if (pictures) A.pic[i] = pictures + P.pic_off[c];
    }
    if (top == 0) {
        for (int i = 0; i < ncomp; i++) {
            const int c = c0 + i;
            A.W[i] = P.pw[c];
            A.H[i] = P.ph[c];
            A.in[i] = raw_in;
            A.in_pic_stride[i] = 0;
            A.ll[i] = coeffs + P.comp_off[c] + P.band[c][0].off;
            A.ll_pic_stride[i] = coeff_pic_stride;
        }
        A.from_u16 = pictures ? 1 : 0;
        dim3 grid(148 * 4, ncomp, npic);
        dwt_copy_kernel<<<grid, 256, 0, s>>>(A);
        VC2B_CHECK_LAUNCH();
        return 0;
    }
    for (int n = top; n >= 1; n--) {
        const bool vert = n > P.dwt_depth_ho;
        const int fb = level_first_band(P, n);
        const int wsel = (top - n) & 1;  // region this level's LL goes to
        for (int i = 0; i < ncomp; i++) {
            const int c = c0 + i;
            const Band &B = P.band[c][fb];
            A.W[i] = 2 * B.w;
            A.H[i] = vert ? 2 * B.h : B.h;
            for (int k = 0; k < (vert ? 3 : 1); k++)
                A.hb[i][k] = coeffs + P.comp_off[c] + P.band[c][fb + k].off;
            if (n == top) {
                A.in[i] = raw_in;
                A.in_pic_stride[i] = 0;
            } else {
                const int rsel = (top - (n + 1)) & 1;
                A.in[i] = scratch + (rsel == 0 ? sp.comp_off_a[c] : sp.size_a + sp.comp_off_b[c]);
                A.in_pic_stride[i] = scratch_pic;
            }
            if (n == 1) {
                A.ll[i] = coeffs + P.comp_off[c] + P.band[c][0].off;
                A.ll_pic_stride[i] = coeff_pic_stride;
            } else {
                A.ll[i] = scratch + (wsel == 0 ? sp.comp_off_a[c] : sp.size_a + sp.comp_off_b[c]);
                A.ll_pic_stride[i] = scratch_pic;
            }
        }
        A.from_u16 = (n == top && pictures) ? 1 : 0;
        int st = launch_ana(P.wavelet_index, P.wavelet_index_ho, vert, A, npic, s);
        if (st) return st;
    }
    return 0;
}

}  // namespace vc2b

using namespace vc2b;

extern "C" {

int vc2b_dwt(const vc2b_params *p, const uint16_t *d_pictures, int npictures, int32_t *d_coeffs,
             void *d_scratch, int64_t scratch_bytes, void *stream)
{
    DevParams P;
    int st = make_dev_params(p, &P);
    if (st) return st;
    if (npictures <= 0) return 0;
    if (!dwt_is_safe(p)) return VC2B_E_UNSUPPORTED;  // outside the int32 envelope
    AnaScratch sp;
    plan_ana_scratch(P, 0, 3, &sp);
    const long long per_pic = (sp.size_a + sp.size_b) * (long long)sizeof(int32_t);
    int group = per_pic > 0 ? (int)(scratch_bytes / per_pic) : npictures;
    if (group < 1) return VC2B_E_SCRATCH_TOO_SMALL;
    if (group > npictures) group = npictures;
    for (int i = 0; i < npictures; i += group) {
        const int n = (npictures - i < group) ? npictures - i : group;
        st = run_dwt(P, 0, 3, d_pictures + (long long)i * P.pic_samples, nullptr, n,
                     d_coeffs + (long long)i * P.pic_coeffs, P.pic_coeffs, (int32_t *)d_scratch, sp,
                     (cudaStream_t)stream);
        if (st) return st;
    }
    return 0;
}

int vc2b_dwt_raw(const vc2b_params *p, int comp, const int32_t *d_in, int32_t *d_coeffs, void *d_scratch,
                 int64_t scratch_bytes, void *stream)
{
    DevParams P;
    int st = make_dev_params(p, &P);
    if (st) return st;
    if (comp < 0 || comp > 2) return VC2B_E_BAD_PARAMS;
    AnaScratch sp;
    plan_ana_scratch(P, comp, comp + 1, &sp);
    if ((sp.size_a + sp.size_b) * (long long)sizeof(int32_t) > scratch_bytes) return VC2B_E_SCRATCH_TOO_SMALL;
    return run_dwt(P, comp, comp + 1, nullptr, d_in, 1, d_coeffs - P.comp_off[comp], 0, (int32_t *)d_scratch, sp,
                   (cudaStream_t)stream);
}

}  // extern "C"
