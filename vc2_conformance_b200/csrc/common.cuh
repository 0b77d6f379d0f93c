// common.cuh -- shared device/host structures of libvc2b200 (sm_100a).
//
// Geometry restates pseudocode/slice_sizes.py:53-127 of the reference
// (/root/reference/vc2_conformance/): subband dimensions with padding, slice
// rectangles.  Everything here is integer index math that becomes kernel
// parameters.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vc2_b200.h"

namespace vc2b {

constexpr int kMaxBands = VC2B_MAX_BANDS;

// One subband of one component inside a picture's flat int32 coefficient block.
struct Band {
    int32_t w, h;      // subband_width / subband_height (slice_sizes.py:53,74)
    int32_t level;     // transform level
    int32_t slot;      // quant_matrix slot: 0 LL/L, 1 HL/H, 2 LH, 3 HH
    int32_t off;       // offset (coefficients) from the start of the component
    int32_t qm;        // quant_matrix[level][slot]
};

// Kernel-side view of vc2b_params (passed by value: lives in the constant bank).
struct DevParams {
    int32_t width[3], height[3];    // picture component dimensions
    int32_t pw[3], ph[3];           // padded (transform) dimensions
    int32_t depth[3];               // bits per sample
    int32_t wavelet_index, wavelet_index_ho;
    int32_t dwt_depth, dwt_depth_ho;
    int32_t slices_x, slices_y;
    int32_t is_ld;
    int32_t slice_prefix_bytes, slice_size_scaler;
    int64_t sb_num, sb_den;         // slice_bytes_numerator / denominator
    int32_t nb;                     // subbands per component
    int32_t comp_off[3];            // offset of each component in a picture's coefficients
    int32_t comp_count[3];          // coefficients per component
    int32_t pic_coeffs;             // coefficients per picture
    int32_t pic_off[3];             // offset (samples) of each plane in a decoded picture
    int32_t pic_samples;            // samples per decoded picture
    Band band[3][kMaxBands];
};

// Builds DevParams; returns 0 or VC2B_E_*.
int make_dev_params(const vc2b_params *p, DevParams *d);
// Largest |coefficient| for which the int32 synthesis provably cannot overflow (params.cu).
long long idwt_max_safe_coeff(const vc2b_params *p);
// Whether vc2b_unpack16 (lane-per-slice kernel, int16 coefficient plane) supports this parameter set (unpack.cu).
bool lanes_unpack_ok(const DevParams &P);
// Whether the int32 analysis of samples of the given depth provably cannot overflow.
bool dwt_is_safe(const vc2b_params *p);

__host__ __device__ inline int32_t ilog2_ceil(int64_t n)
{
    // pseudocode/vc2_math.py:32 intlog2(n) = (n-1).bit_length()
    int32_t bits = 0;
    if (n <= 1) return 0;
    uint64_t v = (uint64_t)(n - 1);
    while (v) { bits++; v >>= 1; }
    return bits;
}

// pseudocode/quantization.py:40 quant_factor, saturating at 2^40 (any non-zero
// magnitude times such a factor leaves the int32 envelope).
__host__ __device__ inline uint64_t quant_factor_u64(int index)
{
    int k = index >> 2;
    if (k > 36) return (uint64_t)1 << 40;
    uint64_t base = (uint64_t)1 << k;
    switch (index & 3) {
    case 0: return 4 * base;
    case 1: return (503829 * base + 52958) / 105917;
    case 2: return (665857 * base + 58854) / 117708;
    default: return (440253 * base + 32722) / 65444;
    }
}

// pseudocode/quantization.py:54 quant_offset
__host__ __device__ inline uint64_t quant_offset_u64(int index)
{
    if (index == 0) return 1;
    if (index == 1) return 2;
    return (quant_factor_u64(index) + 1) / 2;
}

// pseudocode/slice_sizes.py:95 slice_bytes: byte offset of LD slice n is (n*num)/den.
__host__ __device__ inline int64_t ld_slice_offset(const DevParams &P, int64_t n)
{
    return (n * P.sb_num) / P.sb_den;
}

// kernels launched by the library (vc2b_launch_count); defined in params.cu
extern unsigned long long g_launch_count;

#define VC2B_CHECK_LAUNCH()                          \
    do {                                             \
        __atomic_add_fetch(&vc2b::g_launch_count, 1ull, __ATOMIC_RELAXED); \
        cudaError_t e__ = cudaGetLastError();        \
        if (e__ != cudaSuccess) return (int)e__;     \
    } while (0)

}  // namespace vc2b
