// rawio.cu -- raw planar picture files and picture comparison on the device.
//
// Replaces the per-sample Python loops of file_format.write_picture / read_picture
// (file_format.py:158-194, :263-314; sample sizes from dimensions_and_depths.py:46-86) and the
// difference measurement of vc2-picture-compare (scripts/vc2_picture_compare.py:370-420: count of
// differing samples and mean square error per component; PSNR itself is two logarithms on the host).
//
// Raw file layout: components Y, C1, C2 one after another, each `height` rows of `width` samples,
// little endian, bytes per sample = ceil(depth / 8) rounded up to a power of two.  The device picture
// layout is the same planes as uint16, so a component of 9..16 bits is a straight copy and
// one of up to 8 bits a narrowing; depths above 16 bits are outside the library (vc2b_check_params).
#include "common.cuh"

namespace vc2b {

struct RawPlan {
    long long pic_off[3];   // samples: offset of the plane in a device picture
    long long raw_off[3];   // bytes: offset of the plane in a raw picture
    long long count[3];     // samples per plane
    int bps[3];             // bytes per sample in the file
    int depth[3];
    long long pic_samples, raw_bytes;
};

static void make_raw_plan(const DevParams &P, RawPlan *R)
{
    long long off = 0;
    for (int c = 0; c < 3; c++) {
        R->pic_off[c] = P.pic_off[c];
        R->count[c] = (long long)P.width[c] * P.height[c];
        R->depth[c] = P.depth[c];
        R->bps[c] = P.depth[c] <= 8 ? 1 : 2;
        R->raw_off[c] = off;
        off += R->count[c] * R->bps[c];
    }
    R->pic_samples = P.pic_samples;
    R->raw_bytes = off;
}

// to_raw = 1: device pictures -> file bytes; 0: file bytes -> device pictures, bits above `depth` masked
// off as read_picture does (file_format.py:289-297)
__global__ void raw_convert_kernel(RawPlan R, uint16_t *__restrict__ pictures, uint8_t *__restrict__ raw,
                                   long long raw_stride, int to_raw)
{
    const int pic = blockIdx.z, c = blockIdx.y;
    uint16_t *plane = pictures + (long long)pic * R.pic_samples + R.pic_off[c];
    uint8_t *bytes = raw + (long long)pic * raw_stride + R.raw_off[c];
    const unsigned mask = (1u << R.depth[c]) - 1u;
    const long long n = R.count[c];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (R.bps[c] == 2) {
            if (to_raw) {
                const unsigned v = plane[i];
                bytes[2 * i] = (uint8_t)(v & 0xff);
                bytes[2 * i + 1] = (uint8_t)(v >> 8);
            } else {
                plane[i] = (uint16_t)(((unsigned)bytes[2 * i] | ((unsigned)bytes[2 * i + 1] << 8)) & mask);
            }
        } else {
            if (to_raw) bytes[i] = (uint8_t)(plane[i] & 0xff);
            else plane[i] = (uint16_t)(bytes[i] & mask);
        }
    }
}

// out[(pic * 3 + c) * 2 + {0, 1}] += {samples that differ, sum of squared differences}
__global__ void picture_compare_kernel(RawPlan R, const uint16_t *__restrict__ a, const uint16_t *__restrict__ b,
                                       unsigned long long *__restrict__ out)
{
    const int pic = blockIdx.z, c = blockIdx.y;
    const uint16_t *pa = a + (long long)pic * R.pic_samples + R.pic_off[c];
    const uint16_t *pb = b + (long long)pic * R.pic_samples + R.pic_off[c];
    const long long n = R.count[c];
    unsigned long long cnt = 0, sse = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)pa[i] - (int)pb[i];
        cnt += d != 0;
        sse += (unsigned long long)((long long)d * d);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
        sse += __shfl_xor_sync(0xffffffffu, sse, s);
    }
    if ((threadIdx.x & 31) == 0 && (cnt | sse)) {
        atomicAdd(&out[(pic * 3 + c) * 2 + 0], cnt);
        atomicAdd(&out[(pic * 3 + c) * 2 + 1], sse);
    }
}

}  // namespace vc2b

using namespace vc2b;

extern "C" {

int64_t vc2b_raw_picture_bytes(const vc2b_params *p)
{
    DevParams P;
    int st = make_dev_params(p, &P);
    if (st) return st;
    RawPlan R;
    make_raw_plan(P, &R);
    return R.raw_bytes;
}

static int raw_convert(const vc2b_params *p, uint16_t *d_pictures, int npictures, uint8_t *d_raw, int64_t raw_stride,
                       int to_raw, void *stream)
{
    DevParams P;
    int st = make_dev_params(p, &P);
    if (st) return st;
    if (npictures <= 0) return 0;
    RawPlan R;
    make_raw_plan(P, &R);
    if (raw_stride < R.raw_bytes) return VC2B_E_BAD_PARAMS;
    dim3 grid(148 * 2, 3, npictures);
    raw_convert_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(R, d_pictures, d_raw, raw_stride, to_raw);
    VC2B_CHECK_LAUNCH();
    return 0;
}

int vc2b_pictures_to_raw(const vc2b_params *p, const uint16_t *d_pictures, int npictures, uint8_t *d_raw,
                         int64_t raw_stride, void *stream)
{
    return raw_convert(p, const_cast<uint16_t *>(d_pictures), npictures, d_raw, raw_stride, 1, stream);
}

int vc2b_raw_to_pictures(const vc2b_params *p, const uint8_t *d_raw, int64_t raw_stride, int npictures,
                         uint16_t *d_pictures, void *stream)
{
    return raw_convert(p, d_pictures, npictures, const_cast<uint8_t *>(d_raw), raw_stride, 0, stream);
}

int vc2b_picture_compare(const vc2b_params *p, const uint16_t *d_a, const uint16_t *d_b, int npictures,
                         uint64_t *d_result, void *stream)
{
    DevParams P;
    int st = make_dev_params(p, &P);
    if (st) return st;
    if (npictures <= 0) return 0;
    RawPlan R;
    make_raw_plan(P, &R);
    cudaError_t e = cudaMemsetAsync(d_result, 0, sizeof(uint64_t) * 6 * (size_t)npictures, (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    dim3 grid(148, 3, npictures);
    picture_compare_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(R, d_a, d_b,
                                                                   reinterpret_cast<unsigned long long *>(d_result));
    VC2B_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
