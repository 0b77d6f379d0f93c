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
#include <stdlib.h>
#include <string.h>

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
constexpr int kWalkThreads = 1024;

// Where slice n starts depends on the three length bytes of every earlier slice, so the walk is a
// dependent chain (3 loads per slice).  It is broken by SPECULATION: all 1024 threads guess the
// starts of the next 1024 slices -- from the closed form of a constant-bit-rate encoder
// (make_transform_data_hq_lossy, encoder/pictures.py:617: slice n carries
// scaler * (floor((n+1)T/N) - floor(nT/N)) coefficient bytes) or from a constant stride -- read the
// length bytes at the guessed positions in parallel and keep the longest verified prefix (a guess is
// confirmed when the previous slice, itself confirmed, ends exactly there).  Streams that defeat
// both predictors fall back to the serial chase through shared-memory chunks.  The result is always
// the true chain; speculation only changes the speed.
__global__ void __launch_bounds__(kWalkThreads)
hq_walk_kernel(DevParams P, const uint8_t *__restrict__ data, const int64_t *__restrict__ pic_offset,
               int npictures, uint32_t *__restrict__ slice_offset, int32_t *__restrict__ status)
{
    __shared__ __align__(16) uint8_t buf[kWalkChunk];
    __shared__ long long s_pos;
    __shared__ int s_n;
    __shared__ int s_err;
    __shared__ int s_first_bad;
    __shared__ long long s_last_end, s_last_size;
    __shared__ int s_eof_slice;

    const int pic = blockIdx.x;
    const int tid = threadIdx.x;
    const long long pic_base = pic_offset[pic];
    const long long nbytes = pic_offset[pic + 1] - pic_base;
    const long long data_limit = (pic_offset[npictures] + 15) & ~15LL;  // caller pads >= 16 bytes
    const uint8_t *base = data + pic_base;
    const int nslices = P.slices_x * P.slices_y;
    const int scaler = P.slice_size_scaler;
    const int prefix = P.slice_prefix_bytes;
    const int hdr = prefix + 4;
    uint32_t *so = slice_offset + (size_t)pic * nslices;

    if (tid == 0) { s_pos = 0; s_n = 0; s_err = kErrOk; }
    __syncthreads();

    // ---- speculative phase --------------------------------------------------------------------
    const long long T = nbytes - (long long)nslices * hdr;
    const bool closed_ok = (T >= 0) && (T % scaler == 0);
    const long long Tq = closed_ok ? T / scaler : 0;
    int mode = closed_ok ? 0 : 1;   // 0: closed form, 1: constant stride
    long long stride = 0;           // size of the last confirmed slice
    int strikes = 0;
    while (true) {
        const int n0 = s_n;
        const long long pos0 = s_pos;
        if (n0 >= nslices || s_err != kErrOk || strikes >= 2) break;
        __syncthreads();  // everybody has read s_n / s_pos
        if (tid == 0) { s_first_bad = kWalkThreads; s_eof_slice = 0x7fffffff; }
        __syncthreads();
        const int m = n0 + tid;
        long long guess, guess_next;
        if (mode == 0) {
            const long long shift = pos0 - ((long long)n0 * hdr + scaler * (((long long)n0 * Tq) / nslices));
            guess = (long long)m * hdr + scaler * (((long long)m * Tq) / nslices) + shift;
            guess_next = (long long)(m + 1) * hdr + scaler * (((long long)(m + 1) * Tq) / nslices) + shift;
        } else {
            guess = pos0 + (long long)tid * stride;
            guess_next = guess + stride;
        }
        long long end = 0;
        bool link_ok = false;
        if (m < nslices) {
            auto rd = [&](long long q) -> long long { return (q >= 0 && q < nbytes) ? (long long)__ldg(base + q) : 0; };
            const long long p1 = guess + prefix + 2 + scaler * rd(guess + prefix + 1);
            const long long p2 = p1 + 1 + scaler * rd(p1);
            end = p2 + 1 + scaler * rd(p2);
            // my start is confirmed iff every earlier link of this round holds; my own link holds
            // when I end exactly where the next guess starts (the last slice has no successor)
            link_ok = (end == guess_next) || (m == nslices - 1);
            if (mode == 1 && stride == 0) link_ok = false;  // no stride learnt yet: only thread 0 is real
            if (!link_ok) atomicMin(&s_first_bad, tid);
        }
        __syncthreads();
        // slices n0 .. n0+fb have true starts and ends
        const int fb = min(min(s_first_bad, kWalkThreads - 1), nslices - 1 - n0);
        if (m < nslices && tid <= fb && end > nbytes) atomicMin(&s_eof_slice, m);  // decoder/io.py:197
        __syncthreads();
        const int eof = s_eof_slice;
        if (m < nslices && tid <= fb && m < eof) so[m] = (uint32_t)guess;
        if (eof != 0x7fffffff) {
            if (m == eof) {
                s_err = eof * 8 + VC2B_ST_UNEXPECTED_END_OF_STREAM;
                s_n = eof;
                s_pos = guess;
            }
            __syncthreads();
            break;
        }
        if (tid == fb) { s_last_end = end; s_last_size = end - guess; }
        __syncthreads();
        const bool had_stride = !(mode == 1 && stride == 0);
        stride = s_last_size;
        const int confirmed = fb + 1;
        if (tid == 0) { s_n = n0 + confirmed; s_pos = s_last_end; }
        if (had_stride && confirmed < 16 && n0 + confirmed < nslices) {
            strikes++;
            mode = (mode == 0) ? 1 : (closed_ok ? 0 : 1);
        } else if (had_stride) {
            strikes = 0;
        }
        __syncthreads();
    }
    __syncthreads();

    // ---- serial fallback: chase through shared-memory chunks ---------------------------------------
    while (true) {
        long long pos = s_pos;
        int n = s_n;
        if (n >= nslices || s_err != kErrOk) break;
        // chunk covers absolute bytes [a0, a0 + kWalkChunk), a0 16-byte aligned
        const long long abs0 = pic_base + pos;
        const long long a0 = abs0 & ~15LL;
        for (int i = tid; i < kWalkChunk / 16; i += kWalkThreads) {
            long long a = a0 + 16LL * i;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (a < data_limit) v = __ldg(reinterpret_cast<const uint4 *>(data + a));
            reinterpret_cast<uint4 *>(buf)[i] = v;
        }
        __syncthreads();
        if (tid == 0) {
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
            s_pos = pos;
            s_n = n;
            s_err = err;
        }
        __syncthreads();
    }
    // slices at or after a stream end are marked so that the unpack kernel skips them
    if (s_err != kErrOk) {
        for (int k = s_n + tid; k < nslices; k += kWalkThreads) so[k] = 0xffffffffu;
    }
    if (tid == 0) {
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
        // reported only when the stream has no error the reference itself would raise.  Word 1 is the
        // int16-plane overflow flag of unpack_lanes_kernel<.., C16>.
        // (an int16-plane overflow comes first: decoding the picture again through the int32 plane then stores
        // every value exactly and reports the int32 envelope itself)
        status[4 * pic + 0] = status[4 * pic + 1] != 0 ? VC2B_ST_COEFF16_RANGE
                              : (status[4 * pic + 3] > max_safe_coeff ? VC2B_ST_INT32_ENVELOPE : VC2B_ST_OK);
        status[4 * pic + 1] = -1;
    } else {
        status[4 * pic + 0] = key & 7;
        status[4 * pic + 1] = key >> 3;
    }
}

// ---------------------------------------------------------------------------
// warp-level bounded-block decoder
// ---------------------------------------------------------------------------
//
// Round = 1024 bits of the block = one 32-bit word per lane.
//  A. lane <-> word: a byte LUT gives, for each of the four entry states of the exp-Golomb state
//     machine, the exit state and the bitmask of positions where a code STARTS; a warp scan of the
//     32 transition maps fixes every lane's entry state, a prefix sum of the popcounts numbers
//     the codes, and each lane lists the start positions of its codes in shared memory.
//  B. lane <-> code: lane c takes code c, c+32, ...: it reads the code's bits from the words
//     staged in shared memory, decodes magnitude and sign, dequantises and stores.  All lanes
//     work (no per-word imbalance) and consecutive lanes write consecutive coefficients, so the
//     scatter into the subband rectangles is coalesced per rectangle row.

constexpr int kMapMax = 512;        // coefficients of a slice component with a cached destination map
constexpr int kRoundWords = 32;

struct WarpBands {        // per-warp tables in shared memory, rebuilt per slice component
    int32_t cum[kMaxBands + 1];   // first coefficient index of each band in the slice
    int32_t w[kMaxBands];         // rectangle width
    int32_t h[kMaxBands];         // rectangle height
    int32_t stride[kMaxBands];    // subband row stride
    int32_t base[kMaxBands];      // offset of the rectangle's first coefficient in the component
    uint32_t qf[kMaxBands];       // quant_factor(quantizer), saturated to 2^32-1
    uint32_t qo[kMaxBands];       // quant_offset(quantizer) + 2
    int32_t map_w[2][kMaxBands];  // rectangle dims the cached maps (0 luma, 1 chroma) were built for
    int32_t map_h[2][kMaxBands];
    int32_t map_valid[2];
};

struct WarpScratch {
    uint32_t words[kRoundWords + 4];     // the round's words (+ look-ahead for codes that run on)
    uint32_t pos[kRoundWords * 8];       // non-zero codes of the round: start bit | index in the round << 16
    uint32_t map[2][kMapMax];            // [luma, chroma]: coefficient index -> band << 26 | y * stride + x
};

// State machine over the bit string.  0: at a code start, 1: next bit is a data bit,
// 2: next bit is a follow bit and the code already has >= 1 data bit, 3: next bit is the sign.
// LUT entry: exit state | start mask << 8 (bit 7 = first bit of the byte).
__device__ __forceinline__ void build_lut(uint16_t *lut, int tid, int nthreads)
{
    for (int e = tid; e < 1024; e += nthreads) {
        int s = e >> 8, byte = e & 255, mask = 0;
        for (int i = 7; i >= 0; i--) {
            int bit = (byte >> i) & 1;
            if (s == 0) { mask |= 1 << i; s = bit ? 0 : 1; }
            else if (s == 1) s = 2;
            else if (s == 2) s = bit ? 3 : 1;
            else s = 0;
        }
        lut[e] = (uint16_t)(s | (mask << 8));
    }
}

// Short codes: the first 8 bits of a code window -> the signed value if the whole code (sign bit
// included) fits in those 8 bits, else -128.  (0 -> "1"; +-1..+-2 -> 4 bits; +-3..+-6 -> 6 bits;
// +-7..+-14 -> 8 bits: A.4.4 / decoder/io.py:285-313.)
__device__ __forceinline__ void build_short_lut(int8_t *slut, int tid, int nthreads)
{
    for (int e = tid; e < 256; e += nthreads) {
        int value = 1, pos = 7, result = -128;
        bool done = false;
        while (pos >= 0) {
            const int follow = (e >> pos) & 1;
            pos--;
            if (follow) { done = true; break; }
            if (pos < 0) break;
            value = (value << 1) | ((e >> pos) & 1);
            pos--;
        }
        if (done) {
            value -= 1;
            if (value == 0) result = 0;
            else if (pos >= 0) result = ((e >> pos) & 1) ? -value : value;
        }
        slut[e] = (int8_t)result;
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

// destination of coefficient jj of the slice component: band and offset from the rectangle origin
__device__ __forceinline__ uint32_t locate(const WarpBands &T, int nb, int jj)
{
    int b = 0;
    while (b < nb - 1 && jj >= T.cum[b + 1]) b++;
    const int local = jj - T.cum[b];
    const int y = local / T.w[b];
    const int x = local - y * T.w[b];
    return ((uint32_t)b << 26) | (uint32_t)(y * T.stride[b] + x);
}

// Decodes `ncoef` values of the bounded block [start_bit, start_bit+nbits) of `base` and
// scatters the dequantised coefficients.  INTERLEAVED: values alternate C1, C2
// (color_diff_slice_band, decoder/transform_data_syntax.py:315); `out2` is then the C2 base.
// Returns (warp-uniform) the largest |coefficient| stored, or -1 if a value left the
// int32 envelope.
template <bool INTERLEAVED>
__device__ int decode_block(const uint8_t *__restrict__ base, long long start_bit, long long nbits,
                            int ncoef, const WarpBands &T, WarpScratch &W, const uint32_t *__restrict__ cmap, int nb,
                            int32_t *__restrict__ out, int32_t *__restrict__ out2,
                            const uint16_t *__restrict__ lut, const int8_t *__restrict__ slut, int lane)
{
    int state_in = 0;
    int produced = 0;
    int maxabs = 0;
    bool overflow = false;

    // Most coefficients are zero (a single '1' bit, or nothing at all once the block has run out:
    // the 1-extension).  Zero the slice's rectangles first with coalesced stores, then only the
    // non-zero codes are decoded, dequantised and scattered.
    for (int idx = lane; idx < ncoef; idx += 32) {
        const int jj = INTERLEAVED ? (idx >> 1) : idx;
        const uint32_t loc = cmap ? cmap[jj] : locate(T, nb, jj);
        int32_t *dst = (INTERLEAVED && (idx & 1)) ? out2 : out;
        dst[T.base[loc >> 26] + (loc & 0x3ffffffu)] = 0;
    }
    __syncwarp();

    for (long long rbit = 0; produced < ncoef && rbit < nbits; rbit += 32 * kRoundWords) {
        // ---- A: stage the round's words, find every code start ------------------------------
        const uint32_t w0 = fetch_word(base, start_bit, rbit + 32 * lane, nbits);
        __syncwarp();
        W.words[lane] = w0;
        if (lane < 3) W.words[kRoundWords + lane] = fetch_word(base, start_bit, rbit + 32 * (kRoundWords + lane), nbits);
        uint32_t map = 0, mask_sel[4];
#pragma unroll
        for (int s0 = 0; s0 < 4; s0++) {
            int s = s0;
            uint32_t m = 0;
#pragma unroll
            for (int b = 3; b >= 0; b--) {
                const uint32_t e = lut[(s << 8) | ((w0 >> (8 * b)) & 255)];
                s = e & 3;
                m |= (e >> 8) << (8 * b);
            }
            map |= s << (2 * s0);
            mask_sel[s0] = m;
        }
        uint32_t inc = map;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc = compose_maps(o, inc);
        }
        uint32_t exc = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) exc = 0xE4;  // identity
        const int s_start = (exc >> (2 * state_in)) & 3;
        uint32_t mask = mask_sel[0];
        if (s_start == 1) mask = mask_sel[1];
        if (s_start == 2) mask = mask_sel[2];
        if (s_start == 3) mask = mask_sel[3];
        // a code whose first bit is 1 is the value zero: only the others need decoding
        const uint32_t nzmask = mask & ~w0;
        const int mycnt = __popc(mask), mynz = __popc(nzmask);
        int pre = mycnt | (mynz << 16);  // both prefix sums in one scan (counts <= 1024)
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int o = __shfl_up_sync(0xffffffffu, pre, d);
            if (lane >= d) pre += o;
        }
        const int tot = __shfl_sync(0xffffffffu, pre, 31);
        const int total = tot & 0xffff, total_nz = tot >> 16;
        const int last_map = __shfl_sync(0xffffffffu, inc, 31);
        state_in = (last_map >> (2 * state_in)) & 3;
        {
            const int c0 = (pre & 0xffff) - mycnt;   // index (in the round) of this word's first code
            int z = (pre >> 16) - mynz;              // slot of this word's first non-zero code
            uint32_t m = nzmask;
            while (m) {
                const int bitpos = __clz(m);  // bit 31 = first bit of the word
                const int rank = bitpos ? __popc(mask >> (32 - bitpos)) : 0;  // code starts before it
                W.pos[z++] = (uint32_t)(32 * lane + bitpos) | ((uint32_t)(c0 + rank) << 16);
                m &= ~(0x80000000u >> bitpos);
            }
        }
        __syncwarp();

        // ---- B: one non-zero code per lane ---------------------------------------------------
        const int limit = ncoef - produced;  // codes of this round that are coefficients
        for (int c = lane; c < total_nz; c += 32) {
            const uint32_t ent = W.pos[c];
            const int p = ent & 0xffff, ci = ent >> 16;
            if (ci >= limit) continue;
            const int wi = p >> 5, sh = p & 31;
            const uint32_t a0 = W.words[wi], a1 = W.words[wi + 1];
            const uint32_t hi = __funnelshift_l(a1, a0, sh);
            int k;
            uint32_t mag = 0;
            int neg = 0;
            const uint32_t f = hi & 0xAAAAAAAAu;
            const int sv = slut[hi >> 24];
            if (sv != -128) {  // the common case: the whole code within 8 bits
                mag = (uint32_t)(sv < 0 ? -sv : sv);
                neg = sv < 0;
            } else if (f != 0) {  // at most 15 data bits: the whole code sits in 32 bits
                k = __clz(f) >> 1;
                uint32_t v = hi >> (32 - 2 * k);
                v = (v | (v >> 1)) & 0x33333333u;
                v = (v | (v >> 2)) & 0x0f0f0f0fu;
                v = (v | (v >> 4)) & 0x00ff00ffu;
                v = (v | (v >> 8)) & 0x0000ffffu;
                mag = ((1u << k) | v) - 1u;
                neg = (hi >> (30 - 2 * k)) & 1;
            } else {
                const uint32_t a2 = W.words[wi + 2];
                const uint32_t lo = __funnelshift_l(a2, a1, sh);
                const unsigned long long win = ((unsigned long long)hi << 32) | lo;
                const unsigned long long ff = win & 0xAAAAAAAAAAAAAAAAull;
                if (ff == 0) {  // 32 or more data bits: outside the int32 envelope
                    overflow = true;
                } else {
                    k = __clzll((long long)ff) >> 1;
                    unsigned long long v = win >> (64 - 2 * k);
                    v = (v | (v >> 1)) & 0x3333333333333333ull;
                    v = (v | (v >> 2)) & 0x0f0f0f0f0f0f0f0full;
                    v = (v | (v >> 4)) & 0x00ff00ff00ff00ffull;
                    v = (v | (v >> 8)) & 0x0000ffff0000ffffull;
                    v = (v | (v >> 16)) & 0x00000000ffffffffull;
                    mag = ((1u << k) | (uint32_t)v) - 1u;
                    neg = (int)((win >> (62 - 2 * k)) & 1);
                }
            }
            const int idx = produced + ci;
            const int jj = INTERLEAVED ? (idx >> 1) : idx;
            const uint32_t loc = cmap ? cmap[jj] : locate(T, nb, jj);
            const int b = loc >> 26;
            // inverse_quant (pseudocode/quantization.py:18): (m*qf + qo + 2) // 4
            const unsigned long long t = ((unsigned long long)mag * T.qf[b] + T.qo[b]) >> 2;
            if (t > 0x7fffffffull || T.qf[b] == 0xffffffffu) overflow = true;
            maxabs = max(maxabs, (int)t);
            int32_t *dst = (INTERLEAVED && (idx & 1)) ? out2 : out;
            dst[T.base[b] + (loc & 0x3ffffffu)] = neg ? -(int32_t)t : (int32_t)t;
        }
        produced += total;
    }
    overflow = __any_sync(0xffffffffu, overflow);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) maxabs = max(maxabs, __shfl_xor_sync(0xffffffffu, maxabs, d));
    return overflow ? -1 : maxabs;
}

// slice rectangle + quantiser table for component `comp` of slice (sx, sy)
// (slice_left/right/top/bottom pseudocode/slice_sizes.py:108-127, slice_quantizers
// decoder/transform_data_syntax.py:255), and the coefficient -> destination map (reused from
// the previous slice of this warp when the rectangles have the same dimensions).
// Returns the number of coefficients; *use_map tells whether W.map is valid for it.
constexpr int kQTab = 256;  // quantisers 0..255 (qindex is at most 8 bits)

// quant_factor / quant_offset + 2 of every quantiser (pseudocode/quantization.py:40-62), saturated
__device__ __forceinline__ void build_qtab(uint2 *qtab, int tid, int nthreads)
{
    for (int q = tid; q < kQTab; q += nthreads) {
        const uint64_t qf = quant_factor_u64(q);
        qtab[q] = qf > 0xfffffffeull ? make_uint2(0xffffffffu, 0u)
                                      : make_uint2((uint32_t)qf, (uint32_t)quant_offset_u64(q) + 2u);
    }
}

__device__ __forceinline__ int build_bands(const DevParams &P, int comp, int sx, int sy, int qindex,
                                           WarpBands &T, WarpScratch &W, const uint2 *__restrict__ qtab, int map_id,
                                           bool *use_map, int lane)
{
    const int nb = P.nb;
    int cnt = 0;
    bool same = true;
    __syncwarp();
    if (lane < nb) {
        const Band &B = P.band[comp][lane];
        int x1, x2, y1, y2;
        if ((unsigned)B.w < 32768u && (unsigned)B.h < 32768u && P.slices_x < 65536 && P.slices_y < 65536) {
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
        int q = max(qindex - B.qm, 0);
        T.w[lane] = max(x2 - x1, 1);
        T.h[lane] = y2 - y1;
        T.stride[lane] = B.w;
        T.base[lane] = B.off + y1 * B.w + x1;
        const uint2 qq = qtab[min(q, kQTab - 1)];
        T.qf[lane] = qq.x;
        T.qo[lane] = qq.y;
        cnt = (x2 - x1) * (y2 - y1);
        same = T.map_valid[map_id] && (T.map_w[map_id][lane] == x2 - x1) && (T.map_h[map_id][lane] == y2 - y1);
        T.map_w[map_id][lane] = x2 - x1;
        T.map_h[map_id][lane] = y2 - y1;
    }
    int pre = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, pre, d);
        if (lane >= d) pre += o;
    }
    if (lane < nb) T.cum[lane + 1] = pre;
    if (lane == 0) T.cum[0] = 0;
    const int ncoef = __shfl_sync(0xffffffffu, pre, 31);
    same = __all_sync(0xffffffffu, same);
    __syncwarp();
    *use_map = ncoef <= kMapMax;
    if (*use_map && !same) {
        // the map stores band << 26 | (y * stride + x): stride is a property of the band, so the map
        // only depends on the rectangle dimensions
        for (int j = lane; j < ncoef; j += 32) W.map[map_id][j] = locate(T, nb, j);
        if (lane == 0) T.map_valid[map_id] = 1;
    }
    if (!*use_map && lane == 0) T.map_valid[map_id] = 0;
    __syncwarp();
    return ncoef;
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
    __shared__ uint16_t lut[1024];
    __shared__ int8_t slut[256];
    __shared__ WarpBands tabs[kWarpsPerBlock];
    __shared__ WarpScratch scratch[kWarpsPerBlock];
    __shared__ uint2 qtab[kQTab];
    build_lut(lut, threadIdx.x, blockDim.x);
    build_short_lut(slut, threadIdx.x, blockDim.x);
    build_qtab(qtab, threadIdx.x, blockDim.x);
    if ((threadIdx.x & 31) == 0) tabs[threadIdx.x >> 5].map_valid[0] = tabs[threadIdx.x >> 5].map_valid[1] = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nslices = P.slices_x * P.slices_y;
    const long long total_slices = (long long)npictures * nslices;
    WarpBands &T = tabs[warp];
    WarpScratch &W = scratch[warp];
    // persistent warps: consecutive slices of a picture mostly share their rectangle dimensions,
    // so the destination map survives from one slice to the next
    const long long per_warp = (total_slices + (long long)gridDim.x * kWarpsPerBlock - 1) /
                               ((long long)gridDim.x * kWarpsPerBlock);
    const long long gw0 = ((long long)blockIdx.x * kWarpsPerBlock + warp) * per_warp;
    for (long long gw = gw0; gw < gw0 + per_warp && gw < total_slices; gw++) {
        const int pic = total_slices < 0x7fffffffLL ? (int)((unsigned)gw / (unsigned)nslices) : (int)(gw / nslices);
        const int n = (int)(gw - (long long)pic * nslices);
        const int sy = n / P.slices_x, sx = n - sy * P.slices_x;
        const uint32_t off = slice_offset[gw];
        if (off == 0xffffffffu) continue;  // at or after the slice where the stream ended
        const uint8_t *base = data + pic_offset[pic];
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
            bool use_map;
            // luma and chroma rectangles differ: the map is tagged with the component class
            const int ncoef = build_bands(P, c, sx, sy, qindex, T, W, qtab, c == 0 ? 0 : 1, &use_map, lane);
            int m = decode_block<false>(base, pos * 8, 8 * length, ncoef, T, W, use_map ? W.map[c == 0 ? 0 : 1] : nullptr,
                                        P.nb, pc + P.comp_off[c], nullptr, lut, slut, lane);
            if (m < 0) bad = true; else maxabs = max(maxabs, m);
            pos += length;
        }
        if (lane == 0) atomicMax(&status[4 * pic + 3], bad ? 0x7fffffff : maxabs);
    }
}

// One warp per slice, LD (decoder/transform_data_syntax.py:145 ld_slice).
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
unpack_ld_kernel(DevParams P, const uint8_t *__restrict__ data, const int64_t *__restrict__ pic_offset,
                 int npictures, int32_t *__restrict__ coeffs, int32_t *__restrict__ qindex_out,
                 int32_t *__restrict__ status)
{
    __shared__ uint16_t lut[1024];
    __shared__ int8_t slut[256];
    __shared__ WarpBands tabs[kWarpsPerBlock];
    __shared__ WarpScratch scratch[kWarpsPerBlock];
    __shared__ uint2 qtab[kQTab];
    build_lut(lut, threadIdx.x, blockDim.x);
    build_short_lut(slut, threadIdx.x, blockDim.x);
    build_qtab(qtab, threadIdx.x, blockDim.x);
    if ((threadIdx.x & 31) == 0) tabs[threadIdx.x >> 5].map_valid[0] = tabs[threadIdx.x >> 5].map_valid[1] = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nslices = P.slices_x * P.slices_y;
    const long long total_slices = (long long)npictures * nslices;
    WarpBands &T = tabs[warp];
    WarpScratch &W = scratch[warp];
    const long long per_warp = (total_slices + (long long)gridDim.x * kWarpsPerBlock - 1) /
                               ((long long)gridDim.x * kWarpsPerBlock);
    const long long gw0 = ((long long)blockIdx.x * kWarpsPerBlock + warp) * per_warp;
    for (long long gw = gw0; gw < gw0 + per_warp && gw < total_slices; gw++) {
        const int pic = total_slices < 0x7fffffffLL ? (int)((unsigned)gw / (unsigned)nslices) : (int)(gw / nslices);
        const int n = (int)(gw - (long long)pic * nslices);
        const int sy = n / P.slices_x, sx = n - sy * P.slices_x;
        const long long nbytes = pic_offset[pic + 1] - pic_offset[pic];
        const uint8_t *base = data + pic_offset[pic];
        const long long s0 = ld_slice_offset(P, n), s1 = ld_slice_offset(P, n + 1);
        if (s1 > nbytes) continue;  // stream ended inside this slice (reported by ld_init_status_kernel)
        const long long sbytes = s1 - s0;
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
            continue;
        }
        const long long ybit = s0 * 8 + 7 + length_bits;
        int maxabs = 0;
        bool bad = false;
        bool use_map;
        int ncoef = build_bands(P, 0, sx, sy, qindex, T, W, qtab, 0, &use_map, lane);
        int m = decode_block<false>(base, ybit, slice_y_length, ncoef, T, W, use_map ? W.map[0] : nullptr, P.nb,
                                    pc + P.comp_off[0], nullptr, lut, slut, lane);
        if (m < 0) bad = true; else maxabs = max(maxabs, m);
        ncoef = build_bands(P, 1, sx, sy, qindex, T, W, qtab, 1, &use_map, lane);
        m = decode_block<true>(base, ybit + slice_y_length, bits_left - slice_y_length, 2 * ncoef, T, W,
                               use_map ? W.map[1] : nullptr, P.nb, pc + P.comp_off[1], pc + P.comp_off[2], lut, slut, lane);
        if (m < 0) bad = true; else maxabs = max(maxabs, m);
        if (lane == 0) atomicMax(&status[4 * pic + 3], bad ? 0x7fffffff : maxabs);
    }
}

}  // namespace vc2b

#include "unpack_lanes.cuh"

namespace vc2b {

// ---------------------------------------------------------------------------
// DC prediction (decoder/transform_data_syntax.py:63 dc_prediction and its inverse
// encoder/pictures.py:202 apply_dc_prediction).  One CTA per (picture, component).
// mean(a,b,c) = (a+b+c+1)//3 with floor division (pseudocode/vc2_math.py:57).
// ---------------------------------------------------------------------------
constexpr int kDcThreads = 256;

__device__ __forceinline__ long long floordiv3(long long v)
{
    if (v == (long long)(int)v) {  // the usual case: a 32-bit division by a constant
        const int iv = (int)v;
        const int q = iv / 3;
        return (iv - q * 3 < 0) ? q - 1 : q;
    }
    long long q = v / 3;
    return (v - q * 3 < 0) ? q - 1 : q;
}

// |v| < 2^28: three such values plus the bias 3*2^30 + 1 stay below 2^32 (fast path of dc_group)
__device__ __forceinline__ bool dc_small(int v) { return (unsigned)v + (1u << 28) < (1u << 29); }

// Narrows to int32 and keeps the running max |value| in *maxabs (0x7fffffff: left the envelope).
__device__ __forceinline__ int narrow_or_flag(long long v, int *maxabs)
{
    if (v > 0x7fffffffLL || v < -0x7fffffffLL) {  // left the int32 envelope
        *maxabs = 0x7fffffff;
        return 0;
    }
    const int a = (int)(v < 0 ? -v : v);
    *maxabs = max(*maxabs, a);
    return (int)v;
}

// ---- decoder direction: warp-pipelined wavefront ---------------------------------------------------
// out[y][x] = in[y][x] + mean(out[y][x-1], out[y-1][x-1], out[y-1][x]) is a true 2-D recurrence; what is
// parallel is the anti-diagonal.  A warp owns 32 consecutive rows, lane l = row, and runs skewed: at
// step s lane l handles column group s - l (four columns), so the group of the row above was finished
// by lane l-1 one step earlier and arrives by ONE warp shuffle per value -- no shared memory, no block
// barrier on the chain.  The warps of a CTA are pipelined down the band: lane 0 of warp b reads the row
// above (the last row of warp b-1) back from global memory, eight groups ahead, once warp b-1 has
// published (fence + a shared progress word) that it is written; in steady state the producer is 31
// steps ahead and nobody spins.  Own rows are staged eight steps ahead by cp.async.
// Measured (64 1080i fields, 202 us): without the stores 124 us, without the dependent chain 87 us.  What
// binds the kernel is the L1TEX pipe: every lane of a warp touches a different row, so each 16-byte
// load / store instruction is 32 separate wavefronts (about 64 per warp and step, 16 warps per SM).  A
// skewed scratch layout S[x + y][y] would make the wavefront's accesses contiguous (next step).
constexpr int kDcWarps = 8;      // rows per pass = 32 * kDcWarps
constexpr int kDcAhead = 8;      // groups prefetched ahead (register ring; loop unrolled by it)

struct DcLane {
    int left, upleft;
    int maxabs;       // from the exact path (0x7fffffff: left the int32 envelope)
    int vmin, vmax;   // range of the values produced by the fast path
};

// four columns of one row: `cur` the residues, `upv` the finished row above; returns the finished values.
// first_group: x0 == 0.  EDGE: w is not a multiple of four (columns >= w are padding).
template <bool EDGE>
__device__ __forceinline__ int4 dc_group(DcLane &L, const int4 cur, const int4 upv, int x0, int y, int w)
{
    const int in[4] = {cur.x, cur.y, cur.z, cur.w};
    const int uv[4] = {upv.x, upv.y, upv.z, upv.w};
    int o[4];
    // The recurrence is a dependent chain through `left`, so it is kept short: with all operands within
    // +-2^28 (any real stream) the mean is one biased multiply-high in 32 bits, the edge cases are
    // selects off the chain, and the range checks are OR-ed into one word tested once per group.  If
    // the test fails the group is redone exactly in 64 bits.
    constexpr unsigned B = 1u << 28;
    const bool ypos = y > 0, xpos = x0 > 0;
    unsigned rng = ((unsigned)L.left + B) | ((unsigned)L.upleft + B);
    {
        int l = L.left, ul = L.upleft;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const unsigned t1 = (unsigned)ul + (unsigned)uv[i] + 0xC0000001u;  // + 1 + 3*2^30
            const unsigned inb = (unsigned)in[i] - (1u << 30);
            rng |= ((unsigned)uv[i] + B) | ((unsigned)in[i] + B);
            const int q = (int)((__umulhi((unsigned)l + t1, 0xAAAAAAABu) >> 1) + inb);
            // first row: left; first column: up (zero in the corner)
            const bool has_left = (i > 0) || xpos;
            const int e = (int)((unsigned)in[i] + (unsigned)(has_left ? l : uv[i]));
            const int v = (has_left && ypos) ? q : e;
            rng |= (unsigned)v + B;
            o[i] = v;
            l = v;
            ul = uv[i];
        }
    }
    if ((rng >> 29) == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (!EDGE || x0 + i < w) {
                L.vmin = min(L.vmin, o[i]);
                L.vmax = max(L.vmax, o[i]);
            }
        }
        L.left = o[3];
        L.upleft = uv[3];
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int x = x0 + i;
            long long pred;
            if (x > 0 && y > 0) pred = floordiv3((long long)L.left + L.upleft + uv[i] + 1);
            else if (x > 0) pred = L.left;
            else if (y > 0) pred = uv[i];
            else pred = 0;
            o[i] = (x < w) ? narrow_or_flag((long long)in[i] + pred, &L.maxabs) : 0;
            L.left = o[i];
            L.upleft = uv[i];
        }
    }
    return make_int4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(kDcWarps * 32)
dc_prediction_fwd_kernel(DevParams P, int32_t *__restrict__ coeffs, int32_t *__restrict__ status)
{
    __shared__ volatile int progress[kDcWarps];  // groups of the warp's last row that are final in global memory
    const int nwarps = blockDim.x >> 5;
    const int pic = blockIdx.x / 3, comp = blockIdx.x % 3;
    const Band &B = P.band[comp][0];
    const int w = B.w, h = B.h;
    int32_t *band = coeffs + (size_t)pic * P.pic_coeffs + P.comp_off[comp] + B.off;
    const int lane = threadIdx.x & 31, wb = threadIdx.x >> 5;
    const int ngroups = (w + 3) >> 2;
    const bool vec = (w & 3) == 0 && ((reinterpret_cast<uintptr_t>(band) & 15) == 0);
    DcLane L;
    L.maxabs = 0;
    L.vmin = 0;
    L.vmax = 0;

    // Rows are read 16 bytes at a time with a row stride between neighbouring lanes, which is a poor
    // DRAM pattern: pull the band into L2 first with one coalesced sweep of prefetches.
    {
        const char *b0 = reinterpret_cast<const char *>(band);
        const size_t nbytes = (size_t)w * h * sizeof(int32_t);
        for (size_t off = (size_t)threadIdx.x * 128; off < nbytes; off += (size_t)blockDim.x * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(b0 + off));
    }

    // Own rows (and, for the first warp of a later pass, the final row above) are staged eight steps
    // ahead by cp.async into lane-private shared-memory slots, one commit group per step.
    __shared__ int4 stage[kDcAhead][kDcWarps * 32];
    __shared__ int4 stage_up[kDcAhead][kDcWarps];
    auto async_group = [&](int4 *dst, const int32_t *row, int g) {
        if (g < 0 || g >= ngroups) return;
        if (vec) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(row + 4 * g) : "memory");
        } else {
            int4 v = make_int4(0, 0, 0, 0);
            const int x = 4 * g;
            v.x = __ldcg(row + x);
            if (x + 1 < w) v.y = __ldcg(row + x + 1);
            if (x + 2 < w) v.z = __ldcg(row + x + 2);
            if (x + 3 < w) v.w = __ldcg(row + x + 3);
            *dst = v;
        }
    };

    for (int y0 = 0; y0 < h; y0 += 32 * nwarps) {
        if (threadIdx.x < kDcWarps) progress[threadIdx.x] = 0;
        __syncthreads();
        const int yw = y0 + 32 * wb;          // first row of this warp
        const int y = yw + lane;
        if (yw < h) {
            const bool row_ok = y < h;
            int32_t *row = band + (size_t)(row_ok ? y : h - 1) * w;
            const int32_t *above = band + (size_t)(yw > 0 ? yw - 1 : 0) * w;   // lane 0's row above
            const bool has_above = yw > 0;
            const bool wait_above = wb > 0;    // written by warp wb-1 of this pass (else: final since the last pass)
            // lane 0 may read group g of the row above once progress[wb-1] > g
            auto wait_for = [&](int g) {
                if (!wait_above) return;
                const int need = min(g, ngroups - 1) + 1;
                while (progress[wb - 1] < need) __nanosleep(64);
                __threadfence_block();
            };
            L.left = 0;
            L.upleft = 0;
            wait_for(kDcAhead - 1);
#pragma unroll
            for (int u = 0; u < kDcAhead; u++) {
                if (row_ok) async_group(&stage[u][threadIdx.x], row, u - lane);
                if (lane == 0 && has_above) async_group(&stage_up[u][wb], above, u);
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            int4 prev = make_int4(0, 0, 0, 0);   // this lane's output of the previous step
            const int nsteps = ngroups + 31;
            for (int s0 = 0; s0 < nsteps; s0 += kDcAhead) {
                wait_for(s0 + 2 * kDcAhead - 1);  // the groups refilled during this block
#pragma unroll
                for (int u = 0; u < kDcAhead; u++) {
                    const int s = s0 + u;
                    const int g = s - lane;
                    asm volatile("cp.async.wait_group %0;" ::"n"(kDcAhead - 1) : "memory");
                    int4 upv;
                    upv.x = __shfl_up_sync(0xffffffffu, prev.x, 1);
                    upv.y = __shfl_up_sync(0xffffffffu, prev.y, 1);
                    upv.z = __shfl_up_sync(0xffffffffu, prev.z, 1);
                    upv.w = __shfl_up_sync(0xffffffffu, prev.w, 1);
                    if (lane == 0) upv = has_above ? stage_up[u][wb] : make_int4(0, 0, 0, 0);
                    const int4 cur = stage[u][threadIdx.x];
                    // refill the slots with the groups of step s + kDcAhead
                    if (row_ok) async_group(&stage[u][threadIdx.x], row, g + kDcAhead);
                    if (lane == 0 && has_above) async_group(&stage_up[u][wb], above, s + kDcAhead);
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    int4 out = make_int4(0, 0, 0, 0);
                    const bool active = row_ok && g >= 0 && g < ngroups;
                    if (active) {
                        const int x0 = 4 * g;
                        if (vec) {
                            out = dc_group<false>(L, cur, upv, x0, y, w);
                            *reinterpret_cast<int4 *>(row + x0) = out;
                        } else {
                            out = dc_group<true>(L, cur, upv, x0, y, w);
                            row[x0] = out.x;
                            if (x0 + 1 < w) row[x0 + 1] = out.y;
                            if (x0 + 2 < w) row[x0 + 2] = out.z;
                            if (x0 + 3 < w) row[x0 + 3] = out.w;
                        }
                        // the next warp's row above: published once per block of kDcAhead steps (a fence per
                        // step waits for the lane's stores on the critical path)
                        if (lane == 31 && (u == kDcAhead - 1 || g == ngroups - 1)) {
                            __threadfence_block();
                            progress[wb] = g + 1;
                        }
                    }
                    prev = out;
                }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __threadfence_block();
        __syncthreads();   // rows of this pass are final for the next one
    }
    if (status) {
        int maxabs = max(L.maxabs, max(-L.vmin, L.vmax));  // fast-path values are within +-2^28
        maxabs = __reduce_max_sync(0xffffffffu, maxabs);
        if (lane == 0 && maxabs) atomicMax(&status[4 * pic + 3], maxabs);
    }
}

__global__ void __launch_bounds__(kDcThreads)
dc_prediction_inverse_kernel(DevParams P, int32_t *__restrict__ coeffs, int32_t *__restrict__ status)
{
    const int pic = blockIdx.x / 3, comp = blockIdx.x % 3;
    const Band &B = P.band[comp][0];
    const int w = B.w, h = B.h;
    int32_t *band = coeffs + (size_t)pic * P.pic_coeffs + P.comp_off[comp] + B.off;
    const int t = threadIdx.x;
    int maxabs = 0;  // running max |coefficient| of this thread

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
                    band[(size_t)y * w + x] = narrow_or_flag((long long)tile[ty][tx] - pred, &maxabs);
                }
            }
            __syncthreads();
        }
    }
    if (status) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) maxabs = max(maxabs, __shfl_xor_sync(0xffffffffu, maxabs, d));
        if ((t & 31) == 0 && maxabs) atomicMax(&status[4 * pic + 3], maxabs);
    }
}

static int launch_dc_fwd(const DevParams &P, int32_t *d_coeffs, int npictures, int32_t *d_status, cudaStream_t s)
{
    dc_prediction_fwd_kernel<<<npictures * 3, kDcWarps * 32, 0, s>>>(P, d_coeffs, d_status);
    VC2B_CHECK_LAUNCH();
    return 0;
}

}  // namespace vc2b

namespace vc2b {
bool lanes_unpack_ok(const DevParams &P) { return lanes16_eligible(P); }
// vc2b_set_unpack_kernel: 0 = by geometry, 1 = warp per slice, 2 = lane per slice
static int g_unpack_kernel = 0;
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
    if (reinterpret_cast<uintptr_t>(d_data) & 15) return VC2B_E_BAD_PARAMS;  // aligned 16-byte loads
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
    if (inverse)
        dc_prediction_inverse_kernel<<<npictures * 3, kDcThreads, 0, (cudaStream_t)stream>>>(P, d_coeffs, nullptr);
    else
        return launch_dc_fwd(P, d_coeffs, npictures, nullptr, (cudaStream_t)stream);
    VC2B_CHECK_LAUNCH();
    return 0;
}

// vc2b_unpack and vc2b_unpack16 (d_coeffs16 != nullptr: the int16 coefficient plane)
static int unpack_impl(const vc2b_params *p, const uint8_t *d_data, const int64_t *d_pic_offset, int npictures,
                       const uint32_t *d_slice_offset, int32_t *d_coeffs, int16_t *d_coeffs16, int32_t *d_qindex,
                       int32_t *d_status, void *stream)
{
    DevParams P;
    int st = make_dev_params(p, &P);
    if (st) return st;
    if (npictures <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (!P.is_ld && !d_slice_offset) return VC2B_E_BAD_PARAMS;
    if (reinterpret_cast<uintptr_t>(d_data) & 15) return VC2B_E_BAD_PARAMS;  // the kernels stage aligned 16-byte pieces
    // Small slices (the production geometry): one LANE per slice.  Large slices: one WARP per slice.
    // vc2b_set_unpack_kernel overrides the choice (tests exercise both kernels on the same streams).
    bool lanes = lanes_eligible(P);
    const int forced = g_unpack_kernel;
    if (forced == 1) lanes = false;
    if (forced == 2 && !lanes) return VC2B_E_UNSUPPORTED;
    if (d_coeffs16 && !(lanes && lanes16_eligible(P))) return VC2B_E_UNSUPPORTED;
    if (P.is_ld) {
        ld_init_status_kernel<<<(npictures + 127) / 128, 128, 0, s>>>(P, d_pic_offset, npictures, d_status);
        VC2B_CHECK_LAUNCH();
    }
    if (lanes) {
        const long long tasks = (long long)npictures * ((P.slices_x * P.slices_y + 31) / 32);
        long long nblk = (tasks + kLaneWarps - 1) / kLaneWarps;
        // persistent: one CTA of 24 warps per SM (222 KB of shared memory: tables + a tile, a rectangle table and
        // a bitstream FIFO per warp)
        const long long max_blocks = 148LL;
        if (nblk > max_blocks) nblk = max_blocks;
        const void *fn = P.is_ld ? (d_coeffs16 ? (const void *)unpack_lanes_kernel<true, true>
                                               : (const void *)unpack_lanes_kernel<true, false>)
                                 : (d_coeffs16 ? (const void *)unpack_lanes_kernel<false, true>
                                               : (const void *)unpack_lanes_kernel<false, false>);
        // the attribute is per device: set at every launch (a process may drive several GPUs)
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLaneSmemBytes);
        if (e != cudaSuccess) return (int)e;
        const uint32_t *so = P.is_ld ? nullptr : d_slice_offset;
        void *args[] = {&P, &d_data, &d_pic_offset, &npictures, &so, &d_coeffs, &d_coeffs16, &d_qindex, &d_status};
        e = cudaLaunchKernel(fn, dim3((unsigned)nblk), dim3(kLaneWarps * 32), args, kLaneSmemBytes, s);
        if (e != cudaSuccess) return (int)e;
        VC2B_CHECK_LAUNCH();
    } else {
        const long long warps = (long long)npictures * P.slices_x * P.slices_y;
        // persistent warps, each walking a run of consecutive slices (about 16 CTAs per SM worth of
        // work items keeps the tail short)
        const long long max_blocks = 148LL * 64;
        long long nblk = (warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
        if (nblk > max_blocks) nblk = max_blocks;
        const unsigned blocks = (unsigned)nblk;
        if (P.is_ld)
            unpack_ld_kernel<<<blocks, kWarpsPerBlock * 32, 0, s>>>(P, d_data, d_pic_offset, npictures, d_coeffs,
                                                                    d_qindex, d_status);
        else
            unpack_hq_kernel<<<blocks, kWarpsPerBlock * 32, 0, s>>>(P, d_data, d_pic_offset, npictures,
                                                                    d_slice_offset, d_coeffs, d_qindex, d_status);
        VC2B_CHECK_LAUNCH();
    }
    if (P.is_ld) {
        // using_dc_prediction (pseudocode/parse_code_functions.py:70): LD pictures only
        st = launch_dc_fwd(P, d_coeffs, npictures, d_status, s);
        if (st) return st;
    }
    long long safe = idwt_max_safe_coeff(p);
    if (safe > 0x7ffffffeLL) safe = 0x7ffffffeLL;  // 0x7fffffff marks a value that did not fit int32
    finalize_status_kernel<<<(npictures + 127) / 128, 128, 0, s>>>(npictures, d_status, (int)safe);
    VC2B_CHECK_LAUNCH();
    return 0;
}

int vc2b_unpack(const vc2b_params *p, const uint8_t *d_data, const int64_t *d_pic_offset, int npictures,
                const uint32_t *d_slice_offset, int32_t *d_coeffs, int32_t *d_qindex, int32_t *d_status,
                void *stream)
{
    return unpack_impl(p, d_data, d_pic_offset, npictures, d_slice_offset, d_coeffs, nullptr, d_qindex, d_status,
                       stream);
}

int vc2b_unpack16(const vc2b_params *p, const uint8_t *d_data, const int64_t *d_pic_offset, int npictures,
                  const uint32_t *d_slice_offset, int32_t *d_coeffs, int16_t *d_coeffs16, int32_t *d_qindex,
                  int32_t *d_status, void *stream)
{
    if (!d_coeffs16) return VC2B_E_BAD_PARAMS;
    return unpack_impl(p, d_data, d_pic_offset, npictures, d_slice_offset, d_coeffs, d_coeffs16, d_qindex, d_status,
                       stream);
}

int vc2b_set_unpack_kernel(int kernel)
{
    if (kernel < 0 || kernel > 2) return VC2B_E_BAD_PARAMS;
    g_unpack_kernel = kernel;
    return 0;
}

}  // extern "C"
