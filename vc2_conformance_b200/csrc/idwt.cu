// idwt.cu -- inverse wavelet transform (synthesis) for all seven VC-2 filters, symmetric and
// asymmetric, fused with pad removal, clip and offset on the last level.
//
// Replaces picture_decode / inverse_wavelet_transform / idwt / h_synthesis / vh_synthesis /
// oned_synthesis / filter_bit_shift / lift1-4 / idwt_pad_removal / clip_picture /
// offset_picture (pseudocode/picture_decoding.py:45-349).
//
// One kernel launch per transform level, covering all three components and a group of
// pictures (grid = tiles x components x pictures).  A CTA owns a TW x TH tile of the level's
// OUTPUT, stages tile + halo of the four interleaved subbands in shared memory
// (vh_synthesis :164-169), lifts columns with wavelet_index then rows with wavelet_index_ho
// (:172-175), applies the level's rounding shift (:177-182, shift taken from wavelet_index_ho
// :198-201) and writes either the next level's LL band (int32) or -- on the last level -- the
// cropped, clipped, offset uint16 picture.
#include "lifting.cuh"

namespace vc2b {

constexpr int kIdwtThreads = 256;

struct LevelArgs {
    // input LL band and the level's high-pass bands, per component
    const int32_t *ll[3];
    long long ll_pic_stride[3];     // elements between consecutive pictures
    const int32_t *hb[3][3];        // HL, LH, HH (or H in slot 0 for horizontal-only levels)
    long long hb_pic_stride;
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
};

template <int FV, int FH, bool VERT, int TW, int TH>
__global__ void __launch_bounds__(kIdwtThreads)
idwt_level_kernel(LevelArgs A)
{
    constexpr int HX = get_filter(FH).halo;
    constexpr int HY = VERT ? get_filter(FV).halo : 0;
    constexpr int TS = TW + 2 * HX + 1;  // +1: odd stride keeps the row pass off one bank
    extern __shared__ int32_t tile[];

    const int c = blockIdx.y;
    const int pic = blockIdx.z;
    const int tx = blockIdx.x % A.tiles_x[c];
    const int ty = blockIdx.x / A.tiles_x[c];
    if (ty >= A.tiles_y[c]) return;

    const int w = A.w[c], h = A.h[c];
    const int W2 = 2 * w, H2 = VERT ? 2 * h : h;
    const int X0 = tx * TW, Y0 = ty * TH;
    const int gx0 = max(0, X0 - HX), gx1 = min(W2, X0 + TW + HX);
    const int gy0 = max(0, Y0 - HY), gy1 = min(H2, Y0 + TH + HY);
    const int rw = gx1 - gx0, rh = gy1 - gy0;

    // ---- stage the interleaved subbands ------------------------------------------------
    const int32_t *ll = A.ll[c] + (long long)pic * A.ll_pic_stride[c];
    const int32_t *b0 = A.hb[c][0] + (long long)pic * A.hb_pic_stride;
    if (VERT) {
        const int32_t *b1 = A.hb[c][1] + (long long)pic * A.hb_pic_stride;
        const int32_t *b2 = A.hb[c][2] + (long long)pic * A.hb_pic_stride;
        const int sw = rw >> 1, sh = rh >> 1, sx0 = gx0 >> 1, sy0 = gy0 >> 1;
        for (int i = threadIdx.x; i < sw * sh; i += kIdwtThreads) {
            const int sy = i / sw, sx = i - sy * sw;
            const long long g = (long long)(sy0 + sy) * w + sx0 + sx;
            int32_t *t = &tile[(2 * sy) * TS + 2 * sx];
            t[0] = __ldg(ll + g);
            t[1] = __ldg(b0 + g);
            t[TS] = __ldg(b1 + g);
            t[TS + 1] = __ldg(b2 + g);
        }
    } else {
        const int sw = rw >> 1, sx0 = gx0 >> 1;
        for (int i = threadIdx.x; i < sw * rh; i += kIdwtThreads) {
            const int sy = i / sw, sx = i - sy * sw;
            const long long g = (long long)(gy0 + sy) * w + sx0 + sx;
            int32_t *t = &tile[sy * TS + 2 * sx];
            t[0] = __ldg(ll + g);
            t[1] = __ldg(b0 + g);
        }
    }
    __syncthreads();

    // ---- synthesis: columns first, then rows (picture_decoding.py:172-175) ---------------
    if (VERT) lift_tile<FV, true, false>(tile, TS, rw, rh, gy0, H2);
    lift_tile<FH, false, false>(tile, TS, rw, rh, gx0, W2);

    // ---- rounding shift and store (interior of the tile only) ----------------------------
    const int ox1 = min(W2, X0 + TW), oy1 = min(H2, Y0 + TH);
    const int ow = ox1 - X0, oh = oy1 - Y0;
    const int shift = A.shift;
    const int rnd = shift > 0 ? (1 << (shift - 1)) : 0;
    if (A.final_u16) {
        // idwt_pad_removal :282, clip_picture :327, offset_picture :301
        const int cw = A.crop_w[c], ch = A.crop_h[c];
        const int half = 1 << (A.depth[c] - 1);
        uint16_t *dst = A.pic[c] + (long long)pic * A.pic_pic_stride;
        for (int i = threadIdx.x; i < ow * oh; i += kIdwtThreads) {
            const int y = i / ow, x = i - y * ow;
            const int gx = X0 + x, gy = Y0 + y;
            if (gx < cw && gy < ch) {
                int v = (tile[(gy - gy0) * TS + (gx - gx0)] + rnd) >> shift;
                v = min(max(v, -half), half - 1) + half;
                dst[(long long)gy * cw + gx] = (uint16_t)v;
            }
        }
    } else {
        int32_t *dst = A.out[c] + (long long)pic * A.out_pic_stride[c];
        for (int i = threadIdx.x; i < ow * oh; i += kIdwtThreads) {
            const int y = i / ow, x = i - y * ow;
            const int gx = X0 + x, gy = Y0 + y;
            dst[(long long)gy * W2 + gx] = (tile[(gy - gy0) * TS + (gx - gx0)] + rnd) >> shift;
        }
    }
}

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
    constexpr int HX = get_filter(FH).halo;
    constexpr int HY = VERT ? get_filter(FV).halo : 0;
    constexpr int TW = 128;
    constexpr int TH = (HY >= 8) ? 64 : 32;
    constexpr int TS = TW + 2 * HX + 1;
    constexpr size_t smem = sizeof(int32_t) * (size_t)TS * (TH + 2 * HY);
    LevelArgs A = A_in;
    int maxtiles = 0;
    for (int c = 0; c < A.ncomp; c++) {
        A.tiles_x[c] = (2 * A.w[c] + TW - 1) / TW;
        A.tiles_y[c] = ((VERT ? 2 : 1) * A.h[c] + TH - 1) / TH;
        maxtiles = max(maxtiles, A.tiles_x[c] * A.tiles_y[c]);
    }
    auto kern = idwt_level_kernel<FV, FH, VERT, TW, TH>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid(maxtiles, A.ncomp, npic);
    kern<<<grid, kIdwtThreads, smem, s>>>(A);
    VC2B_CHECK_LAUNCH();
    return 0;
}

template <int FV, bool VERT>
static int launch_level_h(int fh, const LevelArgs &A, int npic, cudaStream_t s)
{
    switch (fh) {
    case 0: return launch_level_t<FV, 0, VERT>(A, npic, s);
    case 1: return launch_level_t<FV, 1, VERT>(A, npic, s);
    case 2: return launch_level_t<FV, 2, VERT>(A, npic, s);
    case 3: return launch_level_t<FV, 3, VERT>(A, npic, s);
    case 5: return launch_level_t<FV, 5, VERT>(A, npic, s);
    default: return launch_level_t<FV, 6, VERT>(A, npic, s);
    }
}

static int launch_level(int fv, int fh, bool vert, const LevelArgs &A, int npic, cudaStream_t s)
{
    // Haar with and without shift share their lifting stages (the shift is a runtime argument)
    if (fv == 4) fv = 3;
    if (fh == 4) fh = 3;
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
        sp->size_a += a;
        sp->size_b += b;
    }
}

// Runs all levels for components [c0, c1) of `npic` pictures.
static int run_idwt(const DevParams &P, int c0, int c1, const int32_t *coeffs, long long coeff_pic_stride,
                    int npic, uint16_t *pictures, int32_t *raw_out, int32_t *scratch, const ScratchPlan &sp,
                    cudaStream_t s)
{
    const int top = P.dwt_depth + P.dwt_depth_ho;
    const int ncomp = c1 - c0;
    const long long scratch_pic = sp.size_a + sp.size_b;
    LevelArgs A;
    memset(&A, 0, sizeof(A));
    A.ncomp = ncomp;
    A.hb_pic_stride = coeff_pic_stride;
    A.pic_pic_stride = P.pic_samples;
    A.shift = get_filter(P.wavelet_index_ho).shift;
    for (int i = 0; i < ncomp; i++) {
        const int c = c0 + i;
        A.crop_w[i] = P.width[c];
        A.crop_h[i] = P.height[c];
        A.depth[i] = P.depth[c];
        if (pictures) A.pic[i] = pictures + P.pic_off[c];
    }
    if (top == 0) {
        for (int i = 0; i < ncomp; i++) {
            const int c = c0 + i;
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
        const bool vert = n > P.dwt_depth_ho;
        const bool last = n == top;
        const int fb = level_first_band(P, n);
        // ping-pong: level n writes region ((top-1)-n)%2 (0 = a), reads what level n-1 wrote
        const int wsel = ((top - 1) - n) & 1;
        for (int i = 0; i < ncomp; i++) {
            const int c = c0 + i;
            const Band &B = P.band[c][fb];
            A.w[i] = B.w;
            A.h[i] = B.h;
            for (int k = 0; k < (vert ? 3 : 1); k++)
                A.hb[i][k] = coeffs + P.comp_off[c] + P.band[c][fb + k].off;
            if (n == 1) {
                A.ll[i] = coeffs + P.comp_off[c] + P.band[c][0].off;
                A.ll_pic_stride[i] = coeff_pic_stride;
            } else {
                const int rsel = ((top - 1) - (n - 1)) & 1;
                A.ll[i] = scratch + (rsel == 0 ? sp.comp_off_a[c] : sp.size_a + sp.comp_off_b[c]);
                A.ll_pic_stride[i] = scratch_pic;
            }
            if (!last) {
                A.out[i] = scratch + (wsel == 0 ? sp.comp_off_a[c] : sp.size_a + sp.comp_off_b[c]);
                A.out_pic_stride[i] = scratch_pic;
            } else {
                A.out[i] = raw_out;
                A.out_pic_stride[i] = 0;
            }
        }
        A.final_u16 = (last && pictures) ? 1 : 0;
        int st = launch_level(P.wavelet_index, P.wavelet_index_ho, vert, A, npic, s);
        if (st) return st;
    }
    return 0;
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
    return (n > 0 ? n : 1) * (long long)sizeof(int32_t);
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
        st = run_idwt(P, 0, 3, d_coeffs + (long long)i * P.pic_coeffs, P.pic_coeffs, n,
                      d_pictures + (long long)i * P.pic_samples, nullptr, (int32_t *)d_scratch, sp,
                      (cudaStream_t)stream);
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
    return run_idwt(P, comp, comp + 1, d_coeffs - P.comp_off[comp], 0, 1, nullptr, d_out, (int32_t *)d_scratch,
                    sp, (cudaStream_t)stream);
}

}  // extern "C"
