// unpack.cu -- slice coefficient unpacking fused with inverse quantisation.
//
// Replaces, for a whole batch of pictures at once, the reference's sequential
//   transform_data / hq_slice / ld_slice / slice_band / color_diff_slice_band
//   (decoder/transform_data_syntax.py:116-335), the bounded-block bit reader
//   read_bitb / read_uintb / read_sintb / flush_inputb (decoder/io.py:246-313),
//   inverse_quant / quant_factor / quant_offset (pseudocode/quantization.py:18-62),
//   slice_quantizers (:255) and dc_prediction (:63).
//
// Design (B200): one WARP decodes one slice.  Interleaved exp-Golomb is a
// 4-state machine over the bit string (code start / data bit / follow bit of a
// non-zero code / sign bit).  Each lane takes one 32-bit word of the bounded
// block, evaluates the machine for all four entry states with a byte LUT in
// shared memory, a warp scan composes the 32 transition maps, and each lane
// then decodes exactly the codes that START in its word (a code may run on
// into the following two words).  Bits past the end of the bounded block read
// as 1 (decoder/io.py:246-252), which is implemented by OR-ing ones into the
// fetched words, so dangling values come out exactly as in the reference.
// Consecutive lanes hold consecutive coefficients, so the scatter into the
// subband rectangles is coalesced per rectangle row.
#include "common.cuh"

namespace vc2b {

constexpr int kWarpsPerBlock = 4;
constexpr int kErrOk = 0x7fffffff;

// ---------------------------------------------------------------------------
// HQ slice-offset walk (K7).  One CTA per picture; the stream is staged through
// shared memory in chunks by all threads, thread 0 chases the three length
// bytes of each slice (decoder/transform_data_syntax.py:216-244).
// ---------------------------------------------------------------------------
constexpr int kWalkChunk = 32768;
constexpr int kWalkThreads = 256;

__global__ void __launch_bounds__(kWalkThreads)
hq_walk_kernel(DevParams P, const uint8_t *__restrict__ data, const int64_t *__restrict__ pic_offset,
               int npictures, uint32_t *__restrict__ slice_offset, int32_t *__restrict__ status)
{
    __shared__ __align__(16) uint8_t buf[kWalkChunk];
    __shared__ long long s_pos;
    __shared__ int s_n;
    __shared__ int s_err;

    const int pic = blockIdx.x;
    const long long pic_base = pic_offset[pic];
    const long long nbytes = pic_offset[pic + 1] - pic_base;
    const long long data_limit = (pic_offset[npictures] + 15) & ~15LL;  // caller pads >= 16 bytes
    const uint8_t *base = data + pic_base;
    const int nslices = P.slices_x * P.slices_y;
    const int scaler = P.slice_size_scaler;
    const int prefix = P.slice_prefix_bytes;
    uint32_t *so = slice_offset + (size_t)pic * nslices;

    if (threadIdx.x == 0) { s_pos = 0; s_n = 0; s_err = kErrOk; }
    __syncthreads();

    while (true) {
        long long pos = s_pos;
        int n = s_n;
        if (n >= nslices || s_err != kErrOk) break;
        // chunk covers absolute bytes [a0, a0 + kWalkChunk), a0 16-byte aligned
        const long long abs0 = pic_base + pos;
        const long long a0 = abs0 & ~15LL;
        for (int i = threadIdx.x; i < kWalkChunk / 16; i += kWalkThreads) {
            long long a = a0 + 16LL * i;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (a < data_limit) v = __ldg(reinterpret_cast<const uint4 *>(data + a));
            reinterpret_cast<uint4 *>(buf)[i] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const long long c0 = a0 - pic_base;  // chunk start in picture coordinates
            const long long c1 = c0 + kWalkChunk;
            auto rd = [&](long long q) -> int {
                if (q >= nbytes) return 0;
                if (q >= c0 && q < c1) return buf[q - c0];
                return __ldg(base + q);
            };
            int err = kErrOk;
            while (n < nslices) {
                long long q = pos + prefix;  // qindex byte
                long long p1 = q + 2 + (long long)scaler * rd(q + 1);
                long long p2 = p1 + 1 + (long long)scaler * rd(p1);
                long long nx = p2 + 1 + (long long)scaler * rd(p2);
                if (nx > nbytes) {  // some bit of this slice lies past the end: decoder/io.py:197
                    err = n * 8 + VC2B_ST_UNEXPECTED_END_OF_STREAM;
                    break;
                }
                so[n] = (uint32_t)pos;
                n++;
                pos = nx;
                if (pos + 1024 > c1) break;  // refill (slow global reads still correct otherwise)
            }
            if (err != kErrOk) {
                for (int k = n; k < nslices; k++) so[k] = 0xffffffffu;
            }
            s_pos = pos;
            s_n = n;
            s_err = err;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        status[4 * pic + 0] = s_err;
        status[4 * pic + 1] = 0;
        status[4 * pic + 2] = (int32_t)s_pos;
        status[4 * pic + 3] = 0;
    }
}

__global__ void ld_init_status_kernel(DevParams P, const int64_t *__restrict__ pic_offset, int npictures,
                                      int32_t *__restrict__ status)
{
    int pic = blockIdx.x * blockDim.x + threadIdx.x;
    if (pic >= npictures) return;
    const long long nbytes = pic_offset[pic + 1] - pic_offset[pic];
    const int nslices = P.slices_x * P.slices_y;
    // first slice whose last byte lies past the end of the data (decoder/io.py:197)
    int err = kErrOk;
    long long total = ld_slice_offset(P, nslices);
    if (total > nbytes) {
        int lo = 0, hi = nslices - 1;  // smallest n with offset(n+1) > nbytes
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (ld_slice_offset(P, mid + 1) > nbytes) hi = mid; else lo = mid + 1;
        }
        err = lo * 8 + VC2B_ST_UNEXPECTED_END_OF_STREAM;
    }
    status[4 * pic + 0] = err;
    status[4 * pic + 1] = 0;
    status[4 * pic + 2] = (int32_t)(total <= nbytes ? total : nbytes);
    status[4 * pic + 3] = 0;
}

// status[0] is kept as an atomicMin key (slice*8 + code) while kernels run; this turns it
// into {status, first bad slice}.
__global__ void finalize_status_kernel(int npictures, int32_t *__restrict__ status, int max_safe_coeff)
{
    int pic = blockIdx.x * blockDim.x + threadIdx.x;
    if (pic >= npictures) return;
    int key = status[4 * pic + 0];
    if (key == kErrOk) {
        // a value left the int32 envelope (the reference, in unbounded integers, would carry on):
        // reported only when the stream has no error the reference itself would raise
        status[4 * pic + 0] = (status[4 * pic + 3] > max_safe_coeff) ? VC2B_ST_INT32_ENVELOPE : VC2B_ST_OK;
        status[4 * pic + 1] = -1;
    } else {
        status[4 * pic + 0] = key & 7;
        status[4 * pic + 1] = key >> 3;
    }
}

// ---------------------------------------------------------------------------
// warp-level bounded-block decoder
// ---------------------------------------------------------------------------

struct WarpBands {        // per-warp table in shared memory, rebuilt per slice component
    int32_t cum[kMaxBands + 1];   // first coefficient index of each band in the slice
    int32_t w[kMaxBands];         // rectangle width
    int32_t stride[kMaxBands];    // subband row stride
    int32_t base[kMaxBands];      // offset of the rectangle's first coefficient in the component
    uint64_t qf[kMaxBands];       // quant_factor(quantizer)
    uint64_t qo[kMaxBands];       // quant_offset(quantizer) + 2
};

// State machine over the bit string.  0: at a code start, 1: next bit is a data bit,
// 2: next bit is a follow bit and the code already has >= 1 data bit, 3: next bit is the sign.
__device__ __forceinline__ void build_lut(uint8_t *lut, int tid, int nthreads)
{
    for (int e = tid; e < 1024; e += nthreads) {
        int s = e >> 8, byte = e & 255, cnt = 0;
        for (int i = 7; i >= 0; i--) {
            int bit = (byte >> i) & 1;
            if (s == 0) { cnt++; s = bit ? 0 : 1; }
            else if (s == 1) s = 2;
            else if (s == 2) s = bit ? 3 : 1;
            else s = 0;
        }
        lut[e] = (uint8_t)(s | (cnt << 2));
    }
}

// 32 bits of the bounded block starting at block bit `bit` (may be >= nbits): stream bits,
// with ones beyond the end of the block (decoder/io.py:246-252).
__device__ __forceinline__ uint32_t fetch_word(const uint8_t *__restrict__ base, long long start_bit,
                                               long long bit, long long nbits)
{
    long long remaining = nbits - bit;
    if (remaining <= 0) return 0xffffffffu;
    long long pos = start_bit + bit;
    const uint8_t *ptr = base + (pos >> 3);
    uintptr_t addr = reinterpret_cast<uintptr_t>(ptr);
    const uint32_t *ap = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
    uint32_t w0 = __byte_perm(__ldg(ap), 0, 0x0123);
    uint32_t w1 = __byte_perm(__ldg(ap + 1), 0, 0x0123);
    uint32_t sh = (uint32_t)((addr & 3) * 8 + (pos & 7));
    uint32_t w = __funnelshift_l(w1, w0, sh);
    if (remaining < 32) w |= 0xffffffffu >> (uint32_t)remaining;
    return w;
}

__device__ __forceinline__ uint32_t compose_maps(uint32_t a, uint32_t b)
{
    // apply a, then b; maps are 4 x 2 bits
    uint32_t r = 0;
#pragma unroll
    for (int s = 0; s < 4; s++) r |= ((b >> (2 * ((a >> (2 * s)) & 3))) & 3) << (2 * s);
    return r;
}

// Decodes `ncoef` values of the bounded block [start_bit, start_bit+nbits) of `base` and
// scatters the dequantised coefficients.  INTERLEAVED: values alternate C1, C2
// (color_diff_slice_band, decoder/transform_data_syntax.py:315); `out2` is then the C2 base.
// Returns (warp-uniform) the largest |coefficient| stored, or -1 if a value left the
// int32 envelope.
template <bool INTERLEAVED>
__device__ int decode_block(const uint8_t *__restrict__ base, long long start_bit, long long nbits,
                            int ncoef, const WarpBands &T, int nb, int32_t *__restrict__ out,
                            int32_t *__restrict__ out2, const uint8_t *__restrict__ lut, int lane)
{
    int state_in = 0;
    int produced = 0;
    int maxabs = 0;
    bool overflow = false;

    for (long long rbit = 0; produced < ncoef && rbit < nbits; rbit += 1024) {
        const long long mybit = rbit + 32 * lane;
        const uint32_t w0 = fetch_word(base, start_bit, mybit, nbits);
        const uint32_t w1 = fetch_word(base, start_bit, mybit + 32, nbits);
        const uint32_t w2 = fetch_word(base, start_bit, mybit + 64, nbits);

        // transition map and code-start counts of this word for the four entry states
        uint32_t map = 0, cnts = 0;
#pragma unroll
        for (int s0 = 0; s0 < 4; s0++) {
            int s = s0, c = 0;
#pragma unroll
            for (int b = 3; b >= 0; b--) {
                uint32_t e = lut[(s << 8) | ((w0 >> (8 * b)) & 255)];
                s = e & 3;
                c += e >> 2;
            }
            map |= s << (2 * s0);
            cnts |= c << (8 * s0);
        }
        // inclusive scan of the maps
        uint32_t inc = map;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc = compose_maps(o, inc);
        }
        uint32_t exc = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) exc = 0xE4;  // identity
        const int s_start = (exc >> (2 * state_in)) & 3;
        const int mycnt = (cnts >> (8 * s_start)) & 255;
        int pre = mycnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int o = __shfl_up_sync(0xffffffffu, pre, d);
            if (lane >= d) pre += o;
        }
        const int total = __shfl_sync(0xffffffffu, pre, 31);
        const int last_map = __shfl_sync(0xffffffffu, inc, 31);
        int idx = produced + pre - mycnt;
        state_in = (last_map >> (2 * state_in)) & 3;
        produced += total;

        // first code start inside this word
        int p;
        if (s_start == 0) p = 0;
        else if (s_start == 3) p = 1;
        else {
            uint32_t t = w0 & (0xAAAAAAAAu >> (s_start == 1 ? 1 : 0));
            p = t ? __clz(t) + 2 : 32;
        }

        // destination cursor for coefficient `idx`
        int jj = INTERLEAVED ? (idx >> 1) : idx;
        int b = 0, x = 0, y = 0, bw = 1, next_cum = 0;
        if (p < 32 && idx < ncoef) {
            while (b < nb - 1 && jj >= T.cum[b + 1]) b++;
            bw = T.w[b];
            int local = jj - T.cum[b];
            y = local / bw;
            x = local - y * bw;
            next_cum = T.cum[b + 1];
        }

        while (p < 32 && idx < ncoef) {
            const uint32_t hi = __funnelshift_l(w1, w0, p);
            const uint32_t lo = __funnelshift_l(w2, w1, p);
            const unsigned long long win = ((unsigned long long)hi << 32) | lo;
            const unsigned long long f = win & 0xAAAAAAAAAAAAAAAAull;
            int32_t value = 0;
            int len = 1;
            if (f == 0) {  // 32 or more data bits: outside the int32 envelope
                overflow = true;
                break;
            }
            const int k = __clzll((long long)f) >> 1;
            if (k != 0) {
                uint32_t d;
                if (k <= 8) {
                    uint32_t v = (uint32_t)(win >> (64 - 2 * k));
                    v = (v | (v >> 1)) & 0x3333u;
                    v = (v | (v >> 2)) & 0x0f0fu;
                    v = (v | (v >> 4)) & 0x00ffu;
                    d = v;
                } else {
                    unsigned long long v = win >> (64 - 2 * k);
                    v = (v | (v >> 1)) & 0x3333333333333333ull;
                    v = (v | (v >> 2)) & 0x0f0f0f0f0f0f0f0full;
                    v = (v | (v >> 4)) & 0x00ff00ff00ff00ffull;
                    v = (v | (v >> 8)) & 0x0000ffff0000ffffull;
                    v = (v | (v >> 16)) & 0x00000000ffffffffull;
                    d = (uint32_t)v;
                }
                const uint32_t mag = ((1u << k) | d) - 1u;
                const int neg = (int)((win >> (62 - 2 * k)) & 1);
                // inverse_quant (pseudocode/quantization.py:18): (m*qf + qo + 2) // 4
                const unsigned long long t = ((unsigned long long)mag * T.qf[b] + T.qo[b]) >> 2;
                if (t > 0x7fffffffull) overflow = true;
                value = neg ? -(int32_t)t : (int32_t)t;
                maxabs = max(maxabs, (int)t);
                len = 2 * k + 2;
            }
            int32_t *dst = (INTERLEAVED && (idx & 1)) ? out2 : out;
            dst[T.base[b] + y * T.stride[b] + x] = value;
            p += len;
            idx++;
            if (!INTERLEAVED || !(idx & 1)) {
                jj++;
                x++;
                if (x == bw) { x = 0; y++; }
                if (jj == next_cum && idx < ncoef) {
                    do { b++; } while (b < nb - 1 && jj >= T.cum[b + 1]);
                    bw = T.w[b];
                    next_cum = T.cum[b + 1];
                    x = 0;
                    y = 0;
                }
            }
        }
    }
    if (produced > ncoef) produced = ncoef;
    // everything not reached before the block ran out is zero (code "1" from the 1-extension)
    for (int idx = produced + lane; idx < ncoef; idx += 32) {
        int jj = INTERLEAVED ? (idx >> 1) : idx;
        int b = 0;
        while (b < nb - 1 && jj >= T.cum[b + 1]) b++;
        int local = jj - T.cum[b];
        int y = local / T.w[b];
        int x = local - y * T.w[b];
        int32_t *dst = (INTERLEAVED && (idx & 1)) ? out2 : out;
        dst[T.base[b] + y * T.stride[b] + x] = 0;
    }
    overflow = __any_sync(0xffffffffu, overflow);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) maxabs = max(maxabs, __shfl_xor_sync(0xffffffffu, maxabs, d));
    return overflow ? -1 : maxabs;
}

// slice rectangle + quantiser table for component `comp` of slice (sx, sy)
// (slice_left/right/top/bottom pseudocode/slice_sizes.py:108-127, slice_quantizers
// decoder/transform_data_syntax.py:255).  Returns the number of coefficients.
__device__ __forceinline__ int build_bands(const DevParams &P, int comp, int sx, int sy, int qindex,
                                           WarpBands &T, int lane)
{
    const int nb = P.nb;
    int cnt = 0;
    if (lane < nb) {
        const Band &B = P.band[comp][lane];
        int x1 = (int)(((long long)B.w * sx) / P.slices_x);
        int x2 = (int)(((long long)B.w * (sx + 1)) / P.slices_x);
        int y1 = (int)(((long long)B.h * sy) / P.slices_y);
        int y2 = (int)(((long long)B.h * (sy + 1)) / P.slices_y);
        int q = max(qindex - B.qm, 0);
        T.w[lane] = max(x2 - x1, 1);
        T.stride[lane] = B.w;
        T.base[lane] = B.off + y1 * B.w + x1;
        T.qf[lane] = quant_factor_u64(q);
        T.qo[lane] = quant_offset_u64(q) + 2;
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

__device__ __forceinline__ void report(int32_t *status, int pic, int slice, int code)
{
    atomicMin(&status[4 * pic + 0], slice * 8 + code);
}

// One warp per slice, HQ (decoder/transform_data_syntax.py:212 hq_slice).
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
unpack_hq_kernel(DevParams P, const uint8_t *__restrict__ data, const int64_t *__restrict__ pic_offset,
                 int npictures, const uint32_t *__restrict__ slice_offset, int32_t *__restrict__ coeffs,
                 int32_t *__restrict__ qindex_out, int32_t *__restrict__ status)
{
    __shared__ uint8_t lut[1024];
    __shared__ WarpBands tabs[kWarpsPerBlock];
    build_lut(lut, threadIdx.x, blockDim.x);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nslices = P.slices_x * P.slices_y;
    const long long gw = (long long)blockIdx.x * kWarpsPerBlock + warp;
    if (gw >= (long long)npictures * nslices) return;
    const int pic = (int)(gw / nslices);
    const int n = (int)(gw - (long long)pic * nslices);
    const int sy = n / P.slices_x, sx = n - sy * P.slices_x;
    const uint32_t off = slice_offset[gw];
    if (off == 0xffffffffu) return;  // at or after the slice where the stream ended
    const uint8_t *base = data + pic_offset[pic];
    WarpBands &T = tabs[warp];
    int32_t *pc = coeffs + (size_t)pic * P.pic_coeffs;

    long long pos = (long long)off + P.slice_prefix_bytes;
    const int qindex = base[pos];
    pos += 1;
    if (qindex_out && lane == 0) qindex_out[gw] = qindex;
    int maxabs = 0;
    bool bad = false;
    for (int c = 0; c < 3; c++) {
        const long long length = (long long)P.slice_size_scaler * base[pos];
        pos += 1;
        __syncwarp();
        const int ncoef = build_bands(P, c, sx, sy, qindex, T, lane);
        int m = decode_block<false>(base, pos * 8, 8 * length, ncoef, T, P.nb, pc + P.comp_off[c], nullptr,
                                    lut, lane);
        if (m < 0) bad = true; else maxabs = max(maxabs, m);
        pos += length;
    }
    if (lane == 0) {
        atomicMax(&status[4 * pic + 3], bad ? 0x7fffffff : maxabs);
    }
}

// One warp per slice, LD (decoder/transform_data_syntax.py:145 ld_slice).
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
unpack_ld_kernel(DevParams P, const uint8_t *__restrict__ data, const int64_t *__restrict__ pic_offset,
                 int npictures, int32_t *__restrict__ coeffs, int32_t *__restrict__ qindex_out,
                 int32_t *__restrict__ status)
{
    __shared__ uint8_t lut[1024];
    __shared__ WarpBands tabs[kWarpsPerBlock];
    build_lut(lut, threadIdx.x, blockDim.x);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nslices = P.slices_x * P.slices_y;
    const long long gw = (long long)blockIdx.x * kWarpsPerBlock + warp;
    if (gw >= (long long)npictures * nslices) return;
    const int pic = (int)(gw / nslices);
    const int n = (int)(gw - (long long)pic * nslices);
    const int sy = n / P.slices_x, sx = n - sy * P.slices_x;
    const long long nbytes = pic_offset[pic + 1] - pic_offset[pic];
    const uint8_t *base = data + pic_offset[pic];
    const long long s0 = ld_slice_offset(P, n), s1 = ld_slice_offset(P, n + 1);
    if (s1 > nbytes) return;  // stream ended inside this slice (reported by ld_init_status_kernel)
    const long long sbytes = s1 - s0;
    WarpBands &T = tabs[warp];
    int32_t *pc = coeffs + (size_t)pic * P.pic_coeffs;

    // 7-bit qindex, then intlog2(8*slice_bytes-7)-bit slice_y_length (read_nbits, decoder/io.py:230)
    const int length_bits = ilog2_ceil(8 * sbytes - 7);
    long long hdr = 0;  // up to 7 + 32 header bits, read bytewise
    const int hbytes = (7 + length_bits + 7) >> 3;
    for (int i = 0; i < hbytes; i++) hdr = (hdr << 8) | base[s0 + i];
    hdr >>= (8 * hbytes - 7 - length_bits);
    const int qindex = (int)(hdr >> length_bits);
    const long long slice_y_length = hdr & ((1LL << length_bits) - 1);
    long long bits_left = 8 * sbytes - 7 - length_bits;
    if (qindex_out && lane == 0) qindex_out[gw] = qindex;
    if (slice_y_length > bits_left) {  // InvalidSliceYLength, transform_data_syntax.py:166
        if (lane == 0) report(status, pic, n, VC2B_ST_INVALID_SLICE_Y_LENGTH);
        return;
    }
    const long long ybit = s0 * 8 + 7 + length_bits;
    int maxabs = 0;
    bool bad = false;
    int ncoef = build_bands(P, 0, sx, sy, qindex, T, lane);
    int m = decode_block<false>(base, ybit, slice_y_length, ncoef, T, P.nb, pc + P.comp_off[0], nullptr, lut,
                                lane);
    if (m < 0) bad = true; else maxabs = max(maxabs, m);
    __syncwarp();
    ncoef = build_bands(P, 1, sx, sy, qindex, T, lane);
    m = decode_block<true>(base, ybit + slice_y_length, bits_left - slice_y_length, 2 * ncoef, T, P.nb,
                           pc + P.comp_off[1], pc + P.comp_off[2], lut, lane);
    if (m < 0) bad = true; else maxabs = max(maxabs, m);
    if (lane == 0) {
        atomicMax(&status[4 * pic + 3], bad ? 0x7fffffff : maxabs);
    }
}

// ---------------------------------------------------------------------------
// DC prediction (decoder/transform_data_syntax.py:63 dc_prediction and its inverse
// encoder/pictures.py:202 apply_dc_prediction).  One CTA per (picture, component); thread y
// owns row y of a strip of rows and runs skewed by one column per row, so the three
// neighbours (left: own register; up and up-left: the row above, two steps earlier) are
// always final.  mean(a,b,c) = (a+b+c+1)//3 with floor division (pseudocode/vc2_math.py:57).
// ---------------------------------------------------------------------------
constexpr int kDcThreads = 256;

__device__ __forceinline__ long long floordiv3(long long v)
{
    long long q = v / 3;
    return (v - q * 3 < 0) ? q - 1 : q;
}

__device__ __forceinline__ int narrow_or_flag(long long v, int32_t *status_word)
{
    if (v > 0x7fffffffLL || v < -0x7fffffffLL) {  // left the int32 envelope
        if (status_word) atomicMax(status_word, 0x7fffffff);
        return 0;
    }
    if (status_word) {
        const int a = (int)(v < 0 ? -v : v);
        if (a > *status_word) atomicMax(status_word, a);
    }
    return (int)v;
}

__global__ void __launch_bounds__(kDcThreads)
dc_prediction_kernel(DevParams P, int32_t *__restrict__ coeffs, int inverse, int32_t *__restrict__ status)
{
    __shared__ int32_t up[2][kDcThreads + 1];
    const int pic = blockIdx.x / 3, comp = blockIdx.x % 3;
    const Band &B = P.band[comp][0];
    const int w = B.w, h = B.h;
    int32_t *band = coeffs + (size_t)pic * P.pic_coeffs + P.comp_off[comp] + B.off;
    const int t = threadIdx.x;
    int32_t *sw = status ? &status[4 * pic + 3] : nullptr;  // running max |coefficient|

    if (!inverse) {
        // forward recurrence: strips of rows top to bottom, values final in global memory
        for (int y0 = 0; y0 < h; y0 += kDcThreads) {
            const int y = y0 + t;
            const bool active = y < h;
            int left = 0;     // band[y][x-1] (final)
            int upleft = 0;   // band[y-1][x-1] (final)
            const int rows = min(kDcThreads, h - y0);
            for (int step = 0; step < w + rows - 1; step++) {
                const int x = step - t;
                int value = 0;
                if (active && x >= 0 && x < w) {
                    int upv = 0;
                    if (y > 0) upv = (t == 0) ? band[(size_t)(y - 1) * w + x] : up[(step + 1) & 1][t];
                    long long pred;
                    if (x > 0 && y > 0) pred = floordiv3((long long)left + upleft + upv + 1);
                    else if (x > 0) pred = left;
                    else if (y > 0) pred = upv;
                    else pred = 0;
                    value = narrow_or_flag((long long)band[(size_t)y * w + x] + pred, sw);
                    band[(size_t)y * w + x] = value;
                    left = value;
                    upleft = upv;
                }
                up[step & 1][t + 1] = value;  // row t's value at column x, read by row t+1 next step
                __syncthreads();
            }
            __threadfence_block();
            __syncthreads();
        }
    } else {
        // inverse (encoder/pictures.py:202): the reference walks in REVERSE raster order, so every
        // prediction is formed from neighbours that are still unmodified: out = in - pred(in).
        // In place: tiles are visited bottom-to-top, right-to-left; a tile first stages itself
        // plus the row above and the column to its left (both still original) in shared memory.
        constexpr int TR = 16, TC = 64;
        __shared__ int32_t tile[TR + 1][TC + 1];
        for (int y0 = ((h - 1) / TR) * TR; y0 >= 0; y0 -= TR) {
            for (int x0 = ((w - 1) / TC) * TC; x0 >= 0; x0 -= TC) {
                for (int i = t; i < (TR + 1) * (TC + 1); i += kDcThreads) {
                    int ty = i / (TC + 1), tx = i - ty * (TC + 1);
                    int y = y0 - 1 + ty, x = x0 - 1 + tx;
                    tile[ty][tx] = (y >= 0 && y < h && x >= 0 && x < w) ? band[(size_t)y * w + x] : 0;
                }
                __syncthreads();
                for (int i = t; i < TR * TC; i += kDcThreads) {
                    int ty = i / TC + 1, tx = i - (ty - 1) * TC + 1;
                    int y = y0 - 1 + ty, x = x0 - 1 + tx;
                    if (y < h && x < w) {
                        long long pred;
                        if (x > 0 && y > 0)
                            pred = floordiv3((long long)tile[ty][tx - 1] + tile[ty - 1][tx - 1] + tile[ty - 1][tx] + 1);
                        else if (x > 0) pred = tile[ty][tx - 1];
                        else if (y > 0) pred = tile[ty - 1][tx];
                        else pred = 0;
                        band[(size_t)y * w + x] = narrow_or_flag((long long)tile[ty][tx] - pred, sw);
                    }
                }
                __syncthreads();
            }
        }
    }
}

}  // namespace vc2b

using namespace vc2b;

extern "C" {

int vc2b_hq_slice_offsets(const vc2b_params *p, const uint8_t *d_data, const int64_t *d_pic_offset,
                          int npictures, uint32_t *d_slice_offset, int32_t *d_status, void *stream)
{
    DevParams P;
    int st = make_dev_params(p, &P);
    if (st) return st;
    if (P.is_ld) return VC2B_E_BAD_PARAMS;
    if (npictures <= 0) return 0;
    hq_walk_kernel<<<npictures, kWalkThreads, 0, (cudaStream_t)stream>>>(P, d_data, d_pic_offset, npictures,
                                                                          d_slice_offset, d_status);
    VC2B_CHECK_LAUNCH();
    return 0;
}

int vc2b_dc_prediction(const vc2b_params *p, int32_t *d_coeffs, int npictures, int inverse, void *stream)
{
    DevParams P;
    int st = make_dev_params(p, &P);
    if (st) return st;
    if (npictures <= 0) return 0;
    dc_prediction_kernel<<<npictures * 3, kDcThreads, 0, (cudaStream_t)stream>>>(P, d_coeffs, inverse, nullptr);
    VC2B_CHECK_LAUNCH();
    return 0;
}

int vc2b_unpack(const vc2b_params *p, const uint8_t *d_data, const int64_t *d_pic_offset, int npictures,
                const uint32_t *d_slice_offset, int32_t *d_coeffs, int32_t *d_qindex, int32_t *d_status,
                void *stream)
{
    DevParams P;
    int st = make_dev_params(p, &P);
    if (st) return st;
    if (npictures <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const long long warps = (long long)npictures * P.slices_x * P.slices_y;
    const unsigned blocks = (unsigned)((warps + kWarpsPerBlock - 1) / kWarpsPerBlock);
    if (P.is_ld) {
        ld_init_status_kernel<<<(npictures + 127) / 128, 128, 0, s>>>(P, d_pic_offset, npictures, d_status);
        VC2B_CHECK_LAUNCH();
        unpack_ld_kernel<<<blocks, kWarpsPerBlock * 32, 0, s>>>(P, d_data, d_pic_offset, npictures, d_coeffs,
                                                                d_qindex, d_status);
        VC2B_CHECK_LAUNCH();
        // using_dc_prediction (pseudocode/parse_code_functions.py:70): LD pictures only
        dc_prediction_kernel<<<npictures * 3, kDcThreads, 0, s>>>(P, d_coeffs, 0, d_status);
        VC2B_CHECK_LAUNCH();
    } else {
        if (!d_slice_offset) return VC2B_E_BAD_PARAMS;
        unpack_hq_kernel<<<blocks, kWarpsPerBlock * 32, 0, s>>>(P, d_data, d_pic_offset, npictures,
                                                                d_slice_offset, d_coeffs, d_qindex, d_status);
        VC2B_CHECK_LAUNCH();
    }
    long long safe = idwt_max_safe_coeff(p);
    if (safe > 0x7ffffffeLL) safe = 0x7ffffffeLL;  // 0x7fffffff marks a value that did not fit int32
    finalize_status_kernel<<<(npictures + 127) / 128, 128, 0, s>>>(npictures, d_status, (int)safe);
    VC2B_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
