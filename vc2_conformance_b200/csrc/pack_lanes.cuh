// pack_lanes.cuh -- lane-per-slice slice serialisation (included by quant.cu).
//
// Same job as pack_slices_kernel (hq_slice / ld_slice of bitstream/vc2.py:727-839 as filled in by
// make_hq_slice / make_ld_slice, encoder/pictures.py:456-525: forward_quant of every coefficient at the slice's
// qindex, write_sint bitstream/io.py:575, bounded blocks with their padding) for the production geometry of many
// small slices that all have the same rectangles -- the mirror image of unpack_lanes_kernel:
//
//   * a warp task is 32 consecutive slices of a picture;
//   * LOAD: lanes are code positions -- sixteen independent loads per lane, each instruction reading 16
//     consecutive coefficients of two slices (coalesced), into a padded tile [code][slice];
//   * PACK: lanes are slices -- every lane reads its own 16 values back, quantises them with its slice's quantiser
//     for the band (factor and reciprocal in registers, switched at band boundaries; exact), turns each into
//     its exp-Golomb code (a 64-entry table for |v| <= 31) and appends it to a 64-bit accumulator that is flushed
//     one big-endian word at a time.  A slice is one contiguous bit string (length fields, bounded blocks,
//     padding), so a lane never seeks; words inside the slice are plain stores, the first and last (shared with
//     the neighbouring slices) are OR-ed into the pre-zeroed output.
//
// The warp-per-slice kernel spends most of its time on per-slice set-up, cross-lane prefix sums and an atomicOr
// per code; here a code costs a few dozen instructions of ONE lane (cfg2 89 -> 41 us, cfg5 690 -> 333 us per
// picture).  A launch needs about 32 x 148 x 32 slices to fill the GPU (a lane per slice, 32 resident warps per
// SM): fewer slices than that and the launch lasts as long as one warp's 32 slices.
#pragma once

namespace vc2b {

constexpr int kPkWarps = 16;
constexpr int kPkBands = 16;
constexpr int kPkMaxCoef = 512;
constexpr int kPkChunk = 16;
constexpr int kPkTS = 34;
constexpr int kPkBaseStride = kPkBands + 1;
constexpr int kPkLut = 31;   // codes of |v| <= kPkLut come from the table (at most 12 bits)

struct __align__(16) PkBlock {
    uint32_t map[2][kPkMaxCoef];   // [luma, chroma]: coefficient -> band << 26 | y * stride + x
    int32_t ref_w[2][kPkBands], ref_h[2][kPkBands], ref_cnt[2][kPkBands];
    int32_t qm[kPkBands];          // quant_matrix value of every band (the same for the three components)
    int32_t ref_ncoef[2];
    int32_t pad_[2];
    uint2 qrec[256];               // quantiser index -> {quant_factor saturated to 2^32-1, floor((2^32-1) / it)}
    uint32_t clut[2 * kPkLut + 2]; // v + kPkLut -> code (right aligned) | length << 16
};

struct __align__(16) PkWarp {
    int32_t tile[kPkChunk * kPkTS];      // [code of the chunk][slice]: quantised values
    int32_t base[32 * kPkBaseStride];    // [slice][band]: offset of the rectangle in the component
};

struct PkWriter {
    uint32_t *wp;     // the aligned word the top 32 bits of acc belong to
    uint64_t acc;     // pending bits, first bit at the top
    int nbits;        // pending bits (< 32 between calls)
    bool shared;      // the word at wp also holds bytes of the previous slice
};

__device__ __forceinline__ void pk_init(PkWriter &W, uint8_t *out, long long bitpos)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(out) + (uintptr_t)(bitpos >> 3);
    W.wp = reinterpret_cast<uint32_t *>(a & ~(uintptr_t)3);
    W.nbits = (int)((a & 3) << 3) + (int)(bitpos & 7);
    W.acc = 0;
    W.shared = W.nbits != 0;
}

__device__ __forceinline__ void pk_flush_word(PkWriter &W)
{
    const uint32_t w = __byte_perm((uint32_t)(W.acc >> 32), 0, 0x0123);
    if (W.shared) {
        if (w) atomicOr(W.wp, w);
        W.shared = false;
    } else {
        *W.wp = w;
    }
    W.wp++;
    W.acc <<= 32;
    W.nbits -= 32;
}

// appends `len` (1..32) bits of `code` (right aligned)
__device__ __forceinline__ void pk_emit(PkWriter &W, uint32_t code, int len)
{
    W.acc |= (uint64_t)code << (64 - W.nbits - len);
    W.nbits += len;
    if (W.nbits >= 32) pk_flush_word(W);
}

// the last, partial word of a slice is shared with the next one
__device__ __forceinline__ void pk_finish(PkWriter &W)
{
    if (W.nbits > 0) {
        const uint32_t w = __byte_perm((uint32_t)(W.acc >> 32), 0, 0x0123);
        if (w) atomicOr(W.wp, w);
    }
}

// One code into a bounded block with `left` bits to go: bits past its end are dropped (they are the 1-bits of
// trailing zeros, bitstream/io.py:457-470).
__device__ __forceinline__ void pk_code(PkWriter &W, const PkBlock &S, int32_t v, int &left)
{
    if (left <= 0) return;
    if (v >= -kPkLut && v <= kPkLut) {
        const uint32_t e = S.clut[v + kPkLut];
        uint32_t code = e & 0xffffu;
        int len = (int)(e >> 16);
        if (len > left) {
            code >>= (len - left);
            len = left;
        }
        pk_emit(W, code, len);
        left -= len;
        return;
    }
    int len;
    unsigned long long code = sint_code(v, &len);   // up to 64 bits
    if (len > left) {
        code >>= (len - left);
        len = left;
    }
    left -= len;
    if (len > 32) {
        pk_emit(W, (uint32_t)(code >> 32), len - 32);
        len = 32;
    }
    pk_emit(W, (uint32_t)code, len);
}

__device__ __forceinline__ void pk_zeros(PkWriter &W, int n)
{
    while (n > 0) {
        const int k = n > 32 ? 32 : n;
        pk_emit(W, 0u, k);
        n -= k;
    }
}

// forward_quant magnitude (pseudocode/quantization.py:30): floor(4m / qf), exact, by reciprocal multiplication
// (rec = floor((2^32-1) / qf)); qf == 2^32-1 stands for any factor of 2^32 or more
__device__ __forceinline__ uint32_t pk_quant(uint32_t m, uint32_t qf, uint32_t rec)
{
    if (__builtin_expect(m >= 0x40000000u, 0))   // 4m overflows 32 bits
        return qf == 0xffffffffu ? 0u : (uint32_t)((4ull * m) / qf);
    const uint32_t n = 4u * m;
    uint32_t r = __umulhi(n, rec);
    uint32_t rem = n - r * qf;
    if (rem >= qf) { r++; rem -= qf; }
    if (rem >= qf) r++;   // at most two corrections
    return qf == 0xffffffffu ? 0u : r;
}

// Quantises and packs one bounded block per lane (component `comp` of the lane's slice; INTERLEAVED: C1 and C2
// alternate, color_diff_slice_band order).  `left`: bits of the block; on return what is left of it.
template <bool INTERLEAVED>
__device__ __forceinline__ void pk_block(const DevParams &P, const int comp, PkWarp &Wp, const PkBlock &S,
                                         const bool active, const int sx, const int sy, const int qindex,
                                         const int32_t *__restrict__ src, const int32_t *__restrict__ src2,
                                         PkWriter &W, int &left, const int lane)
{
    constexpr int I = INTERLEAVED ? 1 : 0;
    const int mc = comp ? 1 : 0;
    __syncwarp();
    for (int b = 0; b < P.nb; b++) {
        const Band &B = P.band[comp][b];
        Wp.base[lane * kPkBaseStride + b] = active ? B.off + (S.ref_h[mc][b] * sy) * B.w + S.ref_w[mc][b] * sx : 0;
    }
    const int ncodes = min(S.ref_ncoef[mc], kPkMaxCoef) << I;
    const uint32_t act_mask = __ballot_sync(0xffffffffu, active);
    __syncwarp();
    const int c = lane & 15, h = lane >> 4;
    int b = -1, next = 0;             // band of the code being packed (slice_quantizers: qindex - quant matrix)
    uint32_t qf = 4, rec = 0;
    for (int j0 = 0; j0 < ncodes; j0 += kPkChunk) {
        // ---- load: lane = code, two slices per instruction, sixteen loads in flight ------------------
        const int j = j0 + c;
        const bool have = j < ncodes;
        const uint32_t loc = have ? S.map[mc][j >> I] : 0u;
        const int mb = loc >> 26, moff = loc & 0x3ffffffu;
        const char *p = reinterpret_cast<const char *>(((INTERLEAVED && (j & 1)) ? src2 : src) + moff);
        const int32_t *bb = &Wp.base[h * kPkBaseStride + mb];
        int32_t v[16];
        if (have && act_mask == 0xffffffffu) {
#pragma unroll
            for (int k = 0; k < 16; k++)
                v[k] = __ldg(reinterpret_cast<const int32_t *>(p + 4ull * (uint32_t)bb[2 * k * kPkBaseStride]));
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                v[k] = 0;
                if (have && ((act_mask >> (2 * k + h)) & 1))
                    v[k] = __ldg(reinterpret_cast<const int32_t *>(p + 4ull * (uint32_t)bb[2 * k * kPkBaseStride]));
            }
        }
#pragma unroll
        for (int k = 0; k < 16; k++) Wp.tile[c * kPkTS + 2 * k + h] = v[k];
        __syncwarp();
        // ---- forward_quant + pack: lane = slice --------------------------------------------------------
        const int n = min(kPkChunk, ncodes - j0);
#pragma unroll 1
        for (int i = 0; i < n; i++) {
            if (j0 + i == next) {   // the same for every lane: all slices have the reference rectangles
                do {
                    b++;
                    next += S.ref_cnt[mc][b] << I;
                } while (j0 + i == next);
                const uint2 t = S.qrec[min(max(qindex - S.qm[b], 0), 255)];
                qf = t.x;
                rec = t.y;
            }
            const int32_t cv = Wp.tile[i * kPkTS + lane];
            const uint32_t m = (uint32_t)abs(cv);   // one IABS; |INT_MIN| stays 2^31 as an unsigned value
            const uint32_t r = pk_quant(m, qf, rec);
            if (active) pk_code(W, S, cv < 0 ? -(int32_t)r : (int32_t)r, left);
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void build_pk_block(const DevParams &P, PkBlock &S, int tid, int nthreads)
{
    for (int q = tid; q < 256; q += nthreads) {
        const uint64_t qf = quant_factor_u64(q);
        const uint32_t q32 = qf > 0xfffffffeull ? 0xffffffffu : (uint32_t)qf;
        S.qrec[q] = make_uint2(q32, 0xffffffffu / q32);
    }
    for (int e = tid; e < 2 * kPkLut + 2; e += nthreads) {
        int len = 0;
        const unsigned long long code = e <= 2 * kPkLut ? sint_code(e - kPkLut, &len) : 0ull;
        S.clut[e] = (uint32_t)code | ((uint32_t)len << 16);
    }
    for (int idx = tid; idx < 2 * kPkBands; idx += nthreads) {
        const int mc = idx / kPkBands, b = idx - mc * kPkBands;
        int rw = 0, rh = 0;
        if (b < P.nb) {
            rw = P.band[mc][b].w / P.slices_x;
            rh = P.band[mc][b].h / P.slices_y;
        }
        S.ref_w[mc][b] = rw;
        S.ref_h[mc][b] = rh;
        S.ref_cnt[mc][b] = rw * rh;
        if (mc == 0) S.qm[b] = b < P.nb ? P.band[0][b].qm : 0;
    }
    __syncthreads();
    if (tid < 2) {
        int n = 0;
        for (int b = 0; b < P.nb; b++) n += S.ref_cnt[tid][b];
        S.ref_ncoef[tid] = n;
    }
    __syncthreads();
    for (int idx = tid; idx < 2 * kPkMaxCoef; idx += nthreads) {
        const int mc = idx / kPkMaxCoef, j = idx - mc * kPkMaxCoef;
        uint32_t loc = 0;
        if (j < S.ref_ncoef[mc]) {
            int b = 0, cum = 0;
            while (j >= cum + S.ref_cnt[mc][b]) {
                cum += S.ref_cnt[mc][b];
                b++;
            }
            const int local = j - cum;
            const int y = local / S.ref_w[mc][b];
            const int x = local - y * S.ref_w[mc][b];
            loc = ((uint32_t)b << 26) | (uint32_t)(y * P.band[mc][b].w + x);
        }
        S.map[mc][j] = loc;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kPkWarps * 32)
pack_lanes_kernel(DevParams P, FitArgs F, const int32_t *__restrict__ coeffs, int npictures,
                  const int32_t *__restrict__ qindex_in, const int32_t *__restrict__ bits_in,
                  uint8_t *__restrict__ out, long long out_stride)
{
    extern __shared__ __align__(16) uint8_t pk_smem[];
    PkBlock &S = *reinterpret_cast<PkBlock *>(pk_smem);
    PkWarp *Ws = reinterpret_cast<PkWarp *>(pk_smem + sizeof(PkBlock));
    build_pk_block(P, S, threadIdx.x, blockDim.x);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PkWarp &Wp = Ws[warp];
    const int nslices = P.slices_x * P.slices_y;
    const int tasks_per_pic = (nslices + 31) >> 5;
    const long long total_tasks = (long long)npictures * tasks_per_pic;
    for (long long t = (long long)blockIdx.x * kPkWarps + warp; t < total_tasks; t += (long long)gridDim.x * kPkWarps) {
        const int pic = (int)(t / tasks_per_pic);
        const int n = (int)(t - (long long)pic * tasks_per_pic) * 32 + lane;
        const bool active = n < nslices;
        const int sy = n / P.slices_x, sx = n - sy * P.slices_x;
        const long long gw = (long long)pic * nslices + n;
        const int32_t *pc = coeffs + (size_t)pic * P.pic_coeffs;
        uint8_t *po = out + (size_t)pic * out_stride;
        const int qindex = active ? qindex_in[gw] : 0;
        PkWriter W;
        W.wp = nullptr; W.acc = 0; W.nbits = 0; W.shared = false;
        if (P.is_ld) {
            // ld_slice (bitstream/vc2.py:727): 7-bit qindex, slice_y_length, luma block, chroma block
            long long y_len = 0, c_len = 0;
            if (active) {
                const long long s0 = ld_slice_offset(P, n), s1 = ld_slice_offset(P, n + 1);
                const int length_bits = ilog2_ceil(8 * (s1 - s0) - 7);
                y_len = bits_in[3 * gw + 0];
                c_len = 8 * (s1 - s0) - 7 - length_bits - y_len;
                pk_init(W, po, 8 * s0);
                pk_emit(W, (uint32_t)qindex, 7);
                if (length_bits > 0) pk_emit(W, (uint32_t)y_len, length_bits);   // < 32 bits for any slice of this kernel
            }
            int left = (int)y_len;
            pk_block<false>(P, 0, Wp, S, active, sx, sy, qindex, pc + P.comp_off[0], nullptr, W, left, lane);
            if (active) pk_zeros(W, left);
            left = (int)c_len;
            pk_block<true>(P, 1, Wp, S, active, sx, sy, qindex, pc + P.comp_off[1], pc + P.comp_off[2], W, left, lane);
            if (active) pk_finish(W);   // the rest of the slice is padding: the output is zero already
        } else {
            // hq_slice (bitstream/vc2.py:798): prefix bytes (zero), qindex, then per component a length byte and
            // a bounded block of 8*scaler*length bits (make_hq_slice :456)
            const long long scaler = P.slice_size_scaler;
            const long long align = 8 * scaler;
            long long len[3] = {0, 0, 0};
            if (active) {
                len[0] = (bits_in[3 * gw + 0] + align - 1) / align;
                len[1] = (bits_in[3 * gw + 1] + align - 1) / align;
                long long off;
                if (F.lossless_offset) {  // every length field minimal (make_hq_slice total_length=None, :456)
                    off = F.lossless_offset[(long long)pic * (nslices + 1) + n];
                    len[2] = (bits_in[3 * gw + 2] + align - 1) / align;
                } else {
                    const long long den = (long long)nslices * scaler;
                    off = (long long)n * (P.slice_prefix_bytes + 4) + scaler * ((n * F.total_coeff_bytes) / den);
                    len[2] = hq_total_length(P, F, n) - len[0] - len[1];
                }
                pk_init(W, po, 8 * (off + P.slice_prefix_bytes));
                pk_emit(W, (uint32_t)qindex, 8);
            }
#pragma unroll 1
            for (int c = 0; c < 3; c++) {
                if (active) pk_emit(W, (uint32_t)len[c], 8);
                int left = (int)(align * len[c]);
                pk_block<false>(P, c, Wp, S, active, sx, sy, qindex, pc + P.comp_off[c], nullptr, W, left, lane);
                if (active && c < 2) pk_zeros(W, left);
            }
            if (active) pk_finish(W);
        }
    }
}

constexpr size_t kPkSmemBytes = sizeof(PkBlock) + kPkWarps * sizeof(PkWarp);

// Whether every slice has the same small rectangles (slices_have_same_dimensions, slice_sizes.py:131).
static bool pack_lanes_eligible(const DevParams &P)
{
    if (P.nb > kPkBands) return false;
    for (int c = 0; c < 3; c++) {
        long long tot = 0;
        for (int b = 0; b < P.nb; b++) {
            if (P.band[c][b].w % P.slices_x || P.band[c][b].h % P.slices_y) return false;
            tot += (long long)(P.band[c][b].w / P.slices_x) * (P.band[c][b].h / P.slices_y);
            if (P.band[c][b].qm != P.band[0][b].qm) return false;
        }
        if (tot > kPkMaxCoef || tot < 1) return false;
    }
    // the LD header (7 + intlog2(8 * slice_bytes - 7) bits) is written with two 32-bit emits
    return true;
}

}  // namespace vc2b
