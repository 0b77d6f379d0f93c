// unpack_lanes.cuh -- lane-per-slice coefficient unpacking (included by unpack.cu).
//
// Same job as the warp-per-slice kernels of unpack.cu (hq_slice / ld_slice / slice_band /
// color_diff_slice_band / read_sintb / inverse_quant, decoder/transform_data_syntax.py:145-335,
// decoder/io.py:246-313, pseudocode/quantization.py:18), for the production-style geometry of many
// SMALL slices (<= kLaneMaxCoef coefficients per slice component).
//
// Design (B200): exp-Golomb decoding is a dependent chain per bounded block, but the blocks of
// different slices are independent, so here every LANE owns a slice and walks its own bit string, the 32 lanes
// of a warp in lockstep on the code position:
//   * a 64-bit bit buffer in registers, refilled every fourth code from a lane-private shared-memory FIFO that
//     cp.async tops up once per chunk of codes (LaneReader);
//   * TWO codes per look-up: a 12-bit window table (build_pair_lut) yields both magnitudes, both signs and the
//     bits consumed; a group of four codes in which some lane's pair does not fit the window falls back, for
//     the whole warp, to the one-code step (8-bit window table, long codes out of line);
//   * the dequantised values go through a padded shared-memory tile (written [code][slice], read
//     [slice][code], both conflict-free) so that the stores to the subband rectangles are issued per slice
//     pair with consecutive lanes on consecutive coefficients, every coefficient written exactly once;
//   * lanes_block writes the int32 plane; lanes_block16 the int16 plane of vc2b_unpack16 (two codes per tile
//     word, 32-bit pair stores, level-0 band straight to the int32 plane, overflow flag).
// The coefficient -> (band, y, x) map is built once per CTA for the dimensions of slice (0, 0); slices of
// ragged geometries whose rectangles differ take a slower path (lanes_block only) that locates each
// coefficient from the slice's own band table.  0.92 warp instructions per coefficient at cfg2.
#pragma once

namespace vc2b {

constexpr int kLaneWarps = 24;           // warps per CTA: one persistent CTA per SM (the tables of LaneBlock once per SM)
constexpr int kPairBits = 12;            // window of the two-code look-up table
constexpr int kPairEntries = 1 << kPairBits;
constexpr uint32_t kPairMaxQf = (1u << 26) - 1;  // 30 * qf + quant_offset + 2 < 2^31
constexpr int kLaneBands = 16;           // subbands per component this kernel supports
constexpr int kLaneMaxCoef = 512;        // coefficients per slice component this kernel supports
constexpr int kLaneBaseStride = kLaneBands + 1;
constexpr int kLaneChunk = 16;           // codes per lane between two store phases
constexpr uint32_t kLaneFifoBytes = 128;  // per lane: eight 16-byte pieces of its bitstream
constexpr int kLaneTS = 34;              // tile row stride: conflict-free for both access patterns

struct __align__(16) LaneWarp {          // per-warp shared memory
    int32_t tile[kLaneChunk * kLaneTS];  // [code of the chunk][slice]
    int32_t base[32 * kLaneBaseStride];  // [slice][band]: offset of the rectangle in the component
    uint4 stage[32][kLaneFifoBytes / 16];   // [lane][piece]: the lane's bitstream FIFO (lane_top_up / lane_fetch)
};

struct __align__(16) LaneBlock {                       // per-CTA shared memory
    uint32_t map[2][kLaneMaxCoef];       // [luma, chroma]: coefficient -> band << 26 | y * stride + x
    int32_t ref_w[2][kLaneBands];        // rectangle of slice (0, 0): floor(band / slices)
    int32_t ref_h[2][kLaneBands];
    int32_t even_x[2][kLaneBands];       // ref_w if the band divides evenly between the slices, else -1
    int32_t even_y[2][kLaneBands];
    int32_t ref_cnt[2][kLaneBands];      // ref_w * ref_h
    int32_t even16[2][kLaneBands];       // 1: every rectangle of the band starts at an even element offset
    int32_t ref_ncoef[2];
    int32_t pad_[2];
    uint32_t slut[256];                  // 8-bit window -> 2^length | magnitude << 16 | sign << 30 (0: longer code)
    uint2 qtab[kQTab];
    uint32_t plut[kPairEntries];         // 12-bit window -> the next TWO codes (build_pair_lut)
};

// Codes that fit an 8-bit window, sign included: 0 -> "1"; +-1..2 -> 4 bits; +-3..6 -> 6 bits;
// +-7..14 -> 8 bits (decoder/io.py:285-313).  The entry is laid out for cheap field extraction on two
// different pipes (the INT ALU pipe is the one that saturates): bits 0-8 = 2^length, so consuming the
// code is a multiplication of the bit buffer; bits 16-23 = magnitude; bits 30-31 = sign as a two-bit
// signed number (+1, 0, -1), read with one signed multiply-high.  0 marks a longer code.
__device__ __forceinline__ void build_lane_lut(uint32_t *slut, int tid, int nthreads)
{
    for (int e = tid; e < 256; e += nthreads) {
        int value = 1, pos = 7;
        uint32_t entry = 0;
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
            if (value == 0) {
                entry = 2u;  // one bit, value 0, sign 0
            } else if (pos >= 0) {
                const uint32_t sgn = ((e >> pos) & 1) ? 3u : 1u;
                entry = (1u << (8 - pos)) | ((uint32_t)value << 16) | (sgn << 30);
            }
        }
        slut[e] = entry;
    }
}

// The next TWO codes of a 12-bit window in one look-up (decoder/io.py:285-313 applied twice).  At the bit
// rates of the BASELINE streams 59 % of the codes are the single bit "1" (zero) and 37 % have four bits
// (+-1, +-2): two consecutive codes fit twelve bits unless one of them has ten bits or more, so the
// lanes of a warp stay in lockstep at two codes per look-up and a step in which some lane's pair does not
// fit (entry 0) falls back to the one-code path for the whole warp.
//   byte 0 = bits consumed by both codes (1..12) | sign 2 << 6, byte 1 = magnitude 1, byte 2 = magnitude 2,
//   byte 3 = sign 1 << 6; signs are two-bit signed numbers (+1, 0, -1), magnitudes are <= 30.
// 32-bit entries: one shared-memory wavefront per look-up when the lanes hit different banks (the LSU data
// pipe is the busiest unit of this kernel).
__device__ __forceinline__ void build_pair_lut(uint32_t *plut, int tid, int nthreads)
{
    for (int e = tid; e < kPairEntries; e += nthreads) {
        int pos = kPairBits - 1;
        uint32_t mag[2] = {0, 0}, sgn[2] = {0, 0};
        bool complete = true;
        for (int c = 0; c < 2 && complete; c++) {
            uint32_t value = 1;
            bool stop = false;
            while (pos >= 0) {
                const int follow = (e >> pos) & 1;
                pos--;
                if (follow) { stop = true; break; }
                if (pos < 0) break;
                value = (value << 1) | ((e >> pos) & 1);
                pos--;
            }
            if (!stop) { complete = false; break; }
            value -= 1;
            if (value) {
                if (pos < 0) { complete = false; break; }
                sgn[c] = ((e >> pos) & 1) ? 3u : 1u;
                pos--;
            }
            mag[c] = value;
        }
        uint32_t entry = 0;
        if (complete) {
            const uint32_t used = (uint32_t)(kPairBits - 1 - pos);
            entry = used | (sgn[1] << 6) | (mag[0] << 8) | (mag[1] << 16) | (sgn[0] << 30);
        }
        plut[e] = entry;
    }
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// The bitstream of a lane reaches it through shared memory: a lane-private FIFO of eight 16-byte pieces
// filled by cp.async.  (Per-lane 4-byte global loads were the largest stall of this kernel -- every load
// touches 32 different L1 lines and its scoreboard had to be waited for at the top of every group of codes.)
struct LaneReader {
    const uint4 *cp;      // next 16-byte piece of the stream to stage
    uint64_t buf;         // bit buffer, next bit at the top; bits below `avail` are zero
    uint32_t ahead;       // the word after the buffer (raw), read from the FIFO one refill early
    uint32_t slot;        // shared-memory address of the lane's FIFO (kLaneFifoBytes consecutive bytes)
    uint32_t rd, wr;      // bytes read from / staged into the FIFO so far (rd in words, wr in pieces)
    int avail;            // valid bits in buf
    int remaining;        // bits of the bounded block at and after the next word to read (<= 0: past its end)
};

__device__ __forceinline__ void lane_stage(uint32_t dst, const uint4 *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// Stages pieces while the FIFO has a free slot (the piece being read is not free) and the block goes on.
// Called once per chunk of codes, after a wait for what the previous call staged: a piece is read at least
// a chunk after it was requested, so its latency is never waited for, and the refill itself does no staging
// (the 32 lanes of a warp refill out of phase: anything on that path runs in almost every group of codes).
__device__ __forceinline__ void lane_top_up(LaneReader &R)
{
    while ((int)(R.wr - (R.rd & ~15u)) <= (int)(kLaneFifoBytes - 16) &&
           R.remaining > max(8 * (int)(R.wr - R.rd), 0)) {
        lane_stage(R.slot + (R.wr & (kLaneFifoBytes - 1)), R.cp);
        R.cp++;
        R.wr += 16;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void lane_chunk_start(LaneReader &R)
{
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    lane_top_up(R);
}

// Takes the next word of the block into `ahead`.  Past the end of the block the word is stale shared memory:
// lane_ahead_bits turns it into ones.  No check that the word has arrived: lane_chunk_start leaves at least
// 98 - 36 = 62 complete bytes ahead of a lane that used 36 bytes in the previous chunk, and a chunk of 32 codes
// of at most 8 bits (everything but lane_read_long) uses at most 36.
__device__ __forceinline__ void lane_fetch(LaneReader &R)
{
    R.ahead = lds_u32(R.slot + (R.rd & (kLaneFifoBytes - 1)));
    R.rd += 4;
    R.remaining -= 32;
}

// The word fetched last as 32 bits of the bounded block: ones past its end (decoder/io.py:246-252).
__device__ __forceinline__ uint32_t lane_ahead_bits(const LaneReader &R)
{
    uint32_t w = __byte_perm(R.ahead, 0, 0x0123);
    if (__builtin_expect(R.remaining < 0, 0))   // fewer than 32 bits of the block in the fetched word
        w |= 0xffffffffu >> max(R.remaining + 32, 0);
    return w;
}

__device__ __forceinline__ void lane_refill(LaneReader &R)
{
    if (R.avail <= 32) {
        R.buf |= ((uint64_t)lane_ahead_bits(R) << 32) >> R.avail;
        R.avail += 32;
        lane_fetch(R);
    }
}

// The refill of the long-code path, which may use more of the stream than a chunk's budget: everything staged
// is waited for, and the FIFO is filled up again (and waited for) when less than a chunk's budget is left.
__device__ __forceinline__ void lane_refill_long(LaneReader &R)
{
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if ((int)(R.wr - R.rd) < 48 && R.remaining > max(8 * (int)(R.wr - R.rd), 0)) {
        lane_top_up(R);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    lane_refill(R);
}

// `slot`: shared-memory address of the lane's FIFO (LaneWarp::stage).  Reads whole 16-byte pieces: up to 15
// bytes before the block and 15 bytes past its end, inside the stream buffer and its slack.
__device__ __forceinline__ void lane_reader_init(LaneReader &R, uint32_t slot, const uint8_t *blk, int bit_off,
                                                 long long nbits, bool active)
{
    const uintptr_t addr = reinterpret_cast<uintptr_t>(blk);
    R.cp = reinterpret_cast<const uint4 *>(addr & ~(uintptr_t)15);
    R.slot = slot;
    const uint32_t wa = (uint32_t)(addr & 12);       // byte offset of the first word within its piece
    const int skip = (int)(addr & 3) * 8 + bit_off;  // <= 31
    long long rem = nbits + skip;
    if (rem > (1LL << 30)) rem = 1LL << 30;          // a slice of this kernel never reads that far
    R.remaining = active ? (int)rem : 0;
    R.rd = wa;
    R.wr = 0;
    // nothing of the previous block may still be landing in the FIFO; then fill it (most of a block of the
    // BASELINE streams): one load latency per block
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    lane_top_up(R);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    uint32_t wd0 = __byte_perm(lds_u32(slot + wa), 0, 0x0123);
    uint32_t wd1 = __byte_perm(lds_u32(slot + wa + 4), 0, 0x0123);
    R.ahead = lds_u32(slot + wa + 8);
    if (R.remaining < 32) wd0 |= 0xffffffffu >> max(R.remaining, 0);
    if (R.remaining < 64) wd1 |= 0xffffffffu >> max(R.remaining - 32, 0);
    R.rd = wa + 12;
    R.remaining -= 96;
    R.buf = (((uint64_t)wd0 << 32) | wd1) << skip;
    R.avail = 64 - skip;
}

// A lane whose block left the int32 envelope stops parsing: it reads zeros from here on.
__device__ __forceinline__ void lane_reader_kill(LaneReader &R)
{
    R.buf = ~0ull;
    R.avail = 64;
    R.remaining = -32;
}

// The tail of read_sintb (decoder/io.py:302) for a code longer than the 8-bit window.
// Restores avail >= 32.  Returns false when the value has 32 or more data bits.
__device__ __forceinline__ bool lane_read_long(LaneReader &R, uint32_t &mag, uint32_t &neg)
{
    if (R.avail < 32) lane_refill_long(R);   // the short-path schedule only guarantees 8 bits
    const uint32_t h2 = (uint32_t)(R.buf >> 32);
    const uint32_t f = h2 & 0xAAAAAAAAu;
    bool ok = true;
    if (f) {  // at most 15 data bits: the whole code sits in 32 bits
        const int k = __clz(f) >> 1;
        uint32_t v = h2 >> (32 - 2 * k);
        v = (v | (v >> 1)) & 0x33333333u;
        v = (v | (v >> 2)) & 0x0f0f0f0fu;
        v = (v | (v >> 4)) & 0x00ff00ffu;
        v = (v | (v >> 8)) & 0x0000ffffu;
        mag = ((1u << k) | v) - 1u;
        neg = (h2 >> (30 - 2 * k)) & 1;
        R.buf <<= (2 * k + 2);
        R.avail -= 2 * k + 2;
    } else {  // 16 or more data bits: bit-serial
        uint32_t value = 1;
        int nd = 0;
        while (true) {
            const uint32_t top2 = (uint32_t)(R.buf >> 62);
            if (top2 & 2) break;
            if (nd >= 31) {  // a 32nd data bit: outside the int32 envelope
                ok = false;
                break;
            }
            value = (value << 1) | (top2 & 1);
            nd++;
            R.buf <<= 2;
            R.avail -= 2;
            if (R.avail < 32) lane_refill_long(R);
        }
        mag = value - 1u;
        neg = (uint32_t)(R.buf >> 62) & 1;
        R.buf <<= 2;
        R.avail -= 2;
    }
    if (R.avail < 32) lane_refill_long(R);
    return ok;
}

// Out-of-line copy for the hot loop: long codes are rare, and inlining them four times per loop
// body spreads the hot instructions over several KB of code (instruction-fetch stalls).
struct LaneLong {
    LaneReader R;
    uint32_t mag, neg;
    int ok;
};

__device__ __noinline__ LaneLong lane_read_long_cold(LaneReader R)
{
    LaneLong o;
    o.ok = lane_read_long(R, o.mag, o.neg) ? 1 : 0;
    o.R = R;
    return o;
}

// Byte `B` of x, sign extended (prmt in its default mode replicates the sign bit of the selected byte when
// bit 3 of a selector nibble is set; __byte_perm masks that bit away).
template <int B>
__device__ __forceinline__ int prmt_sext(uint32_t x)
{
    int v;
    asm("prmt.b32 %0, %1, %1, %2;" : "=r"(v) : "r"(x), "n"(0x8880 + B + 0x1110 * B));
    return v;
}

// Width | height << 16 of the rectangle of slice (sx, sy) in a band of bw x bh (slice_left/right/
// top/bottom, pseudocode/slice_sizes.py:108-127).  Out of line: only slices of ragged geometries
// whose rectangles differ from slice (0, 0)'s come here.
__device__ __noinline__ uint32_t lane_rect_wh(int bw, int bh, int nx, int ny, int sx, int sy)
{
    const int x1 = (int)(((long long)bw * sx) / nx), x2 = (int)(((long long)bw * (sx + 1)) / nx);
    const int y1 = (int)(((long long)bh * sy) / ny), y2 = (int)(((long long)bh * (sy + 1)) / ny);
    return (uint32_t)(x2 - x1) | ((uint32_t)(y2 - y1) << 16);
}

// Decodes one bounded block per lane (component `comp` of the lane's slice) and scatters the
// dequantised coefficients.  INTERLEAVED: values alternate C1, C2 (color_diff_slice_band,
// decoder/transform_data_syntax.py:315); `out2` is then the C2 base.
template <bool INTERLEAVED>
__device__ __forceinline__ void lanes_block(const DevParams &P, const int comp, LaneWarp &W, const LaneBlock &S,
                                            const bool active, const int sx, const int sy, const int qindex,
                                            const uint8_t *blk, const int bit_off, const long long nbits,
                                            int32_t *__restrict__ out, int32_t *__restrict__ out2, const int lane,
                                            uint32_t &maxabs, bool &ovf)
{
    constexpr int I = INTERLEAVED ? 1 : 0;
    const int mc = comp ? 1 : 0;
    const int nb = P.nb;

    // ---- the slice's rectangles (slice_left/right/top/bottom, pseudocode/slice_sizes.py:108-127)
    int ncoef = 0;
    bool match = true;
    __syncwarp();
    for (int b = 0; b < nb; b++) {
        const Band &B = P.band[comp][b];
        // lanes without a slice run along on the reference rectangle and decode zeros
        int w = S.ref_w[mc][b], h = S.ref_h[mc][b], base = 0;
        if (active) {
            int x1, x2, y1, y2;
            const int ex = S.even_x[mc][b], ey = S.even_y[mc][b];
            if (ex >= 0) {
                x1 = ex * sx;
                x2 = x1 + ex;
            } else if ((unsigned)B.w < 32768u && P.slices_x < 65536) {
                x1 = (int)(((unsigned)B.w * (unsigned)sx) / (unsigned)P.slices_x);
                x2 = (int)(((unsigned)B.w * (unsigned)(sx + 1)) / (unsigned)P.slices_x);
            } else {
                x1 = (int)(((long long)B.w * sx) / P.slices_x);
                x2 = (int)(((long long)B.w * (sx + 1)) / P.slices_x);
            }
            if (ey >= 0) {
                y1 = ey * sy;
                y2 = y1 + ey;
            } else if ((unsigned)B.h < 32768u && P.slices_y < 65536) {
                y1 = (int)(((unsigned)B.h * (unsigned)sy) / (unsigned)P.slices_y);
                y2 = (int)(((unsigned)B.h * (unsigned)(sy + 1)) / (unsigned)P.slices_y);
            } else {
                y1 = (int)(((long long)B.h * sy) / P.slices_y);
                y2 = (int)(((long long)B.h * (sy + 1)) / P.slices_y);
            }
            w = x2 - x1;
            h = y2 - y1;
            base = B.off + y1 * B.w + x1;
            match = match && (w == S.ref_w[mc][b]) && (h == S.ref_h[mc][b]);
        }
        W.base[lane * kLaneBaseStride + b] = base;
        ncoef += w * h;
    }
    const int ncodes = ncoef << I;
    int jmax = ncodes;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) jmax = max(jmax, __shfl_xor_sync(0xffffffffu, jmax, d));
    const uint32_t act_mask = __ballot_sync(0xffffffffu, active);
    const uint32_t match_mask = __ballot_sync(0xffffffffu, active && match);
    __syncwarp();
    if (act_mask == 0) return;
    const bool uniform = match_mask == act_mask;  // every lane decodes the reference rectangle
    const bool my_match = match || !active;

    LaneReader R;
    lane_reader_init(R, (uint32_t)__cvta_generic_to_shared(&W.stage[lane][0]), blk, bit_off, nbits, active);

    int b = -1, next = 0;
    uint32_t qf = 0, qo = 0, maxmag = 0;
    const int refcodes = min(S.ref_ncoef[mc], kLaneMaxCoef) << I;
    uint32_t slut_addr = (uint32_t)__cvta_generic_to_shared(S.slut);
    asm volatile("" : "+r"(slut_addr));  // keep the address in a register (not rebuilt at every use)
    int32_t *tile = &W.tile[lane];

    // Largest magnitude read in the current band: inverse_quant is monotone in the magnitude, so the
    // envelope bookkeeping (largest |coefficient|, overflow of m*qf+qo, saturated quantiser) is done
    // once per band instead of once per coefficient.
#define VC2B_LANE_CLOSE_BAND()                                                                       \
    if (maxmag) {                                                                                    \
        const unsigned long long tm = (unsigned long long)maxmag * qf + qo;                          \
        maxabs = max(maxabs, (uint32_t)(tm >> 2));                                                   \
        hiacc |= (uint32_t)(tm >> 32);                                                               \
        if (qf == 0xffffffffu) bad |= 2u;                                                            \
        maxmag = 0;                                                                                  \
    }

    // `mark` is multiplied by 2^length with the buffer, so the bits consumed since the last settle
    // are log2(mark): avail is only needed at refills and on the long-code path.
#define VC2B_LANE_SETTLE()                                                                           \
    {                                                                                                \
        R.avail -= mark ? 31 - __clz(mark) : 32;                                                     \
        mark = 1;                                                                                    \
    }

    // the band of code j0 + i starts here (slice_quantizers, transform_data_syntax.py:255)
#define VC2B_LANE_NEXT_BAND(i)                                                                       \
    {                                                                                                \
        VC2B_LANE_CLOSE_BAND()                                                                       \
        do {                                                                                         \
            b++;                                                                                     \
            int cnt = S.ref_cnt[mc][b];                                                              \
            if (!my_match) {                                                                         \
                const uint32_t wh = lane_rect_wh(P.band[comp][b].w, P.band[comp][b].h, P.slices_x,   \
                                                 P.slices_y, sx, sy);                                \
                cnt = (int)(wh & 0xffffu) * (int)(wh >> 16);                                         \
            }                                                                                        \
            next += cnt << I;                                                                        \
        } while (j0 + (i) == next);                                                                  \
        const int q = max(qindex - P.band[comp][b].qm, 0);                                           \
        const uint2 qq = S.qtab[min(q, kQTab - 1)];                                                  \
        qf = qq.x;                                                                                   \
        qo = qq.y;                                                                                   \
        qfast = qf <= kPairMaxQf;                                                                    \
    }

    // one code: read_sintb + inverse_quant, value to tile row i
#define VC2B_LANE_STEP(i, CHK)                                                                       \
    {                                                                                                \
        if (CHK && j0 + (i) == next) VC2B_LANE_NEXT_BAND(i)                                          \
        const uint32_t hi = (uint32_t)(R.buf >> 32);                                                 \
        const uint32_t e = lds_u32(slut_addr + 4 * (hi >> 24));                                      \
        const uint32_t M = e & 0x1ffu;                                                               \
        uint32_t mag = __byte_perm(e, 0, 0x4442);                                                    \
        int sgn = __mulhi((int)e, 4);  /* e >> 30, arithmetic, on the multiplier pipe */             \
        if (__builtin_expect(M == 0, 0)) {                                                           \
            VC2B_LANE_SETTLE()                                                                       \
            const LaneLong o = lane_read_long_cold(R);                                               \
            R = o.R;                                                                                 \
            mag = o.mag;                                                                             \
            sgn = o.neg ? -1 : 1;                                                                    \
            if (!o.ok) {                                                                             \
                bad |= 1u;                                                                           \
                lane_reader_kill(R);                                                                 \
            }                                                                                        \
        } else {                                                                                     \
            R.buf *= M;                                                                              \
            mark *= M;                                                                               \
        }                                                                                            \
        /* inverse_quant (pseudocode/quantization.py:18): (m*qf + qo + 2) // 4; sign 0 keeps zero zero */ \
        const uint32_t t = (uint32_t)(((unsigned long long)mag * qf + qo) >> 2);                     \
        maxmag = max(maxmag, mag);                                                                   \
        tile[(i) * kLaneTS] = (int32_t)t * sgn;                                                      \
    }

    // Two codes per look-up (build_pair_lut).  The magnitudes are at most 30 and the lane's quantiser factor
    // is below 2^26 (qfast), so m*qf + qo stays below 2^31: 32-bit arithmetic, no envelope bookkeeping beyond
    // the largest magnitude of the band.
#define VC2B_LANE_PAIR_VALUE(i, m, sg)                                                               \
    tile[(i) * kLaneTS] = (int32_t)(((m) * qf + qo) >> 2) * (sg);

    // the int32 envelope was left if some m*qf+qo reached 2^33 (hiacc >= 2), a non-zero value met a
    // saturated quantiser or a value had 32 or more data bits (bad != 0)
    uint32_t hiacc = 0, bad = 0, mark = 1;
    bool qfast = false;
    uint32_t plut_addr = (uint32_t)__cvta_generic_to_shared(S.plut);
    asm volatile("" : "+r"(plut_addr));
    const int c = lane & 15, h = lane >> 4;   // store phase: half-warp h takes slice 2k+h, lane c its code c
    for (int j0 = 0; j0 < jmax; j0 += kLaneChunk) {
        // ---- decode: lane = slice, kLaneChunk codes each -------------------------------------------
        lane_chunk_start(R);
        if (uniform && j0 + kLaneChunk <= refcodes) {
#pragma unroll 1
            for (int g = 0; g < kLaneChunk / 4; g++) {
                if (j0 == next) VC2B_LANE_NEXT_BAND(0)  // a band starts with this group
                if (next - j0 >= 4) {  // no band starts within these four codes
                    // two look-ups of two codes each; mark == 1 here, avail >= 33
                    const uint32_t ea = lds_u32(plut_addr + (((uint32_t)(R.buf >> 32) >> (32 - kPairBits)) << 2));
                    const uint32_t la = ea & 15u;
                    const uint64_t bufa = R.buf << la;
                    const uint32_t eb = lds_u32(plut_addr + (((uint32_t)(bufa >> 32) >> (32 - kPairBits)) << 2));
                    const uint32_t lb = eb & 15u;
                    if (__all_sync(0xffffffffu, qfast && la * lb != 0)) {
                        R.buf = bufa << lb;
                        R.avail -= (int)(la + lb);
                        const uint32_t ma1 = __byte_perm(ea, 0, 0x4441), ma2 = __byte_perm(ea, 0, 0x4442);
                        const uint32_t mb1 = __byte_perm(eb, 0, 0x4441), mb2 = __byte_perm(eb, 0, 0x4442);
                        VC2B_LANE_PAIR_VALUE(0, ma1, __mulhi((int)ea, 4))
                        VC2B_LANE_PAIR_VALUE(1, ma2, __mulhi((int)(ea << 24), 4))
                        VC2B_LANE_PAIR_VALUE(2, mb1, __mulhi((int)eb, 4))
                        VC2B_LANE_PAIR_VALUE(3, mb2, __mulhi((int)(eb << 24), 4))
                        maxmag = max(max(maxmag, max(ma1, ma2)), max(mb1, mb2));
                    } else {
                        VC2B_LANE_STEP(0, false)
                        VC2B_LANE_STEP(1, false)
                        VC2B_LANE_STEP(2, false)
                        VC2B_LANE_STEP(3, false)
                        VC2B_LANE_SETTLE()
                    }
                } else {
                    VC2B_LANE_STEP(0, false)
                    VC2B_LANE_STEP(1, true)
                    VC2B_LANE_STEP(2, true)
                    VC2B_LANE_STEP(3, true)
                    VC2B_LANE_SETTLE()
                }
                lane_refill(R);
                j0 += 4;
                tile += 4 * kLaneTS;
            }
        } else {
#pragma unroll 1
            for (int i = 0; i < kLaneChunk; i++) {
                if (j0 < ncodes) VC2B_LANE_STEP(0, true)
                VC2B_LANE_SETTLE()
                if (R.avail < 32) lane_refill(R);
                j0 += 1;
                tile += kLaneTS;
            }
        }
        j0 -= kLaneChunk;
        tile -= kLaneChunk * kLaneTS;
        __syncwarp();

        // ---- store: two slices per instruction, 16 consecutive codes each ---------------------------
        const int j = j0 + c;
        const int jj = j >> I;
        const uint32_t loc = j < refcodes ? S.map[mc][jj] : 0u;
        const int mb = loc >> 26, moff = loc & 0x3ffffffu;
        int32_t *dst = (INTERLEAVED && (j & 1)) ? out2 : out;
        if (uniform) {
            if (j < refcodes) {
                // one 64-bit multiply-add per address: byte pointer + 4 * (unsigned) rectangle offset
                char *p = reinterpret_cast<char *>(dst + moff);
                const int32_t *tb = &W.tile[c * kLaneTS + h];
                const int32_t *bb = &W.base[h * kLaneBaseStride + mb];
                if (act_mask == 0xffffffffu) {
#pragma unroll
                    for (int k = 0; k < 16; k++)
                        *reinterpret_cast<int32_t *>(p + 4ull * (uint32_t)bb[2 * k * kLaneBaseStride]) = tb[2 * k];
                } else {
#pragma unroll 4
                    for (int k = 0; k < 16; k++)
                        if ((act_mask >> (2 * k + h)) & 1)
                            *reinterpret_cast<int32_t *>(p + 4ull * (uint32_t)bb[2 * k * kLaneBaseStride]) = tb[2 * k];
                }
            }
        } else {
            for (int k = 0; k < 16; k++) {
                const int s = 2 * k + h;
                const int ncodes_s = __shfl_sync(0xffffffffu, ncodes, s);
                const int sxs = __shfl_sync(0xffffffffu, sx, s), sys = __shfl_sync(0xffffffffu, sy, s);
                if (!((act_mask >> s) & 1)) continue;
                if ((match_mask >> s) & 1) {
                    if (j < refcodes) dst[W.base[s * kLaneBaseStride + mb] + moff] = W.tile[c * kLaneTS + s];
                } else if (j < ncodes_s) {
                    // locate code j in slice s's own rectangles
                    int bb = 0, cum = 0;
                    uint32_t wh = lane_rect_wh(P.band[comp][0].w, P.band[comp][0].h, P.slices_x, P.slices_y, sxs, sys);
                    while (jj >= cum + (int)(wh & 0xffffu) * (int)(wh >> 16)) {
                        cum += (int)(wh & 0xffffu) * (int)(wh >> 16);
                        bb++;
                        wh = lane_rect_wh(P.band[comp][bb].w, P.band[comp][bb].h, P.slices_x, P.slices_y, sxs, sys);
                    }
                    const int local = jj - cum;
                    const int w = (int)(wh & 0xffffu);
                    const int y = local / w;
                    const int x = local - y * w;
                    dst[W.base[s * kLaneBaseStride + bb] + y * P.band[comp][bb].w + x] = W.tile[c * kLaneTS + s];
                }
            }
        }
        __syncwarp();
    }
    VC2B_LANE_CLOSE_BAND()
#undef VC2B_LANE_STEP
#undef VC2B_LANE_NEXT_BAND
#undef VC2B_LANE_PAIR_VALUE
#undef VC2B_LANE_SETTLE
#undef VC2B_LANE_CLOSE_BAND
    if ((hiacc >> 1) != 0 || (bad & 3u) != 0) ovf = true;
}

// ---- the int16 coefficient plane (vc2b_unpack16) -----------------------------------------------------
// Same decoding as lanes_block for parameter sets whose slices all have the reference rectangles
// (lanes16_eligible).  Every band but the level-0 band is stored as int16 in `o16` / `o16b` at the element
// offsets of the int32 layout, which halves what the coefficients cost on their way to the IDWT (the largest
// DRAM stream of the decoder) and the shared-memory and store traffic of this kernel: the tile holds TWO codes
// per 32-bit word, a chunk is 32 codes, and a store instruction writes 32 consecutive coefficients of two
// slices (one 32-bit store per lane where a pair of codes is adjacent in its band, which is everywhere but in
// the smallest bands).  The level-0 band -- a couple of values per slice that DC prediction and the first
// synthesis level read as int32 -- goes straight from the decoding lane to the int32 plane.  A high-pass value
// outside int16 sets `ovf16`; the picture is then decoded again through the int32 plane.
template <bool INTERLEAVED>
__device__ __forceinline__ void lanes_block16(const DevParams &P, const int comp, LaneWarp &W, const LaneBlock &S,
                                              const bool active, const int sx, const int sy, const int qindex,
                                              const uint8_t *blk, const int bit_off, const long long nbits,
                                              int32_t *__restrict__ out, int32_t *__restrict__ out2,
                                              int16_t *__restrict__ o16, int16_t *__restrict__ o16b, const int lane,
                                              uint32_t &maxabs, bool &ovf, bool &ovf16)
{
    constexpr int I = INTERLEAVED ? 1 : 0;
    constexpr int CH = 32;             // codes per chunk
    constexpr int TS2 = 2 * kLaneTS;   // int16 elements per tile row
    const int mc = comp ? 1 : 0;
    const int nb = P.nb;

    __syncwarp();
    for (int b = 0; b < nb; b++) {
        const Band &B = P.band[comp][b];
        W.base[lane * kLaneBaseStride + b] = active ? B.off + (S.ref_h[mc][b] * sy) * B.w + S.ref_w[mc][b] * sx : 0;
    }
    const int ncodes = min(S.ref_ncoef[mc], kLaneMaxCoef) << I;
    const uint32_t act_mask = __ballot_sync(0xffffffffu, active);
    __syncwarp();
    if (act_mask == 0) return;

    LaneReader R;
    lane_reader_init(R, (uint32_t)__cvta_generic_to_shared(&W.stage[lane][0]), blk, bit_off, nbits, active);

    int b = -1, next = 0;
    uint32_t qf = 0, qo = 0, maxmag = 0;
    uint32_t slut_addr = (uint32_t)__cvta_generic_to_shared(S.slut);
    asm volatile("" : "+r"(slut_addr));
    uint32_t plut_addr = (uint32_t)__cvta_generic_to_shared(S.plut);
    asm volatile("" : "+r"(plut_addr));
    int16_t *t16 = reinterpret_cast<int16_t *>(&W.tile[lane]);
    const int base0 = W.base[lane * kLaneBaseStride];
    uint32_t hiacc = 0, bad = 0, mark = 1;
    bool qfast = false;

#define VC2B_LANE_CLOSE_BAND()                                                                       \
    if (maxmag) {                                                                                    \
        const unsigned long long tm = (unsigned long long)maxmag * qf + qo;                          \
        maxabs = max(maxabs, (uint32_t)(tm >> 2));                                                   \
        hiacc |= (uint32_t)(tm >> 32);                                                               \
        if (qf == 0xffffffffu) bad |= 2u;                                                            \
        if (b > 0 && (tm >> 2) > 32767ull) bad |= 4u;                                                \
        maxmag = 0;                                                                                  \
    }

#define VC2B_LANE_SETTLE()                                                                           \
    {                                                                                                \
        R.avail -= mark ? 31 - __clz(mark) : 32;                                                     \
        mark = 1;                                                                                    \
    }

    // the band of code j0 + i starts here; the level-0 band never takes the two-code path (its values are
    // stored by VC2B_LANE_STEP16 as int32)
#define VC2B_LANE_NEXT_BAND(i)                                                                       \
    {                                                                                                \
        VC2B_LANE_CLOSE_BAND()                                                                       \
        do {                                                                                         \
            b++;                                                                                     \
            next += S.ref_cnt[mc][b] << I;                                                           \
        } while (j0 + (i) == next);                                                                  \
        const int q = max(qindex - P.band[comp][b].qm, 0);                                           \
        const uint2 qq = S.qtab[min(q, kQTab - 1)];                                                  \
        qf = qq.x;                                                                                   \
        qo = qq.y;                                                                                   \
        qfast = qf <= kPairMaxQf && b > 0;                                                           \
    }

    // one code: read_sintb + inverse_quant; value to half (i & 1) of tile row i >> 1
#define VC2B_LANE_STEP16(i, CHK)                                                                     \
    {                                                                                                \
        if (CHK && j0 + (i) == next) VC2B_LANE_NEXT_BAND(i)                                          \
        const uint32_t hi = (uint32_t)(R.buf >> 32);                                                 \
        const uint32_t e = lds_u32(slut_addr + 4 * (hi >> 24));                                      \
        const uint32_t M = e & 0x1ffu;                                                               \
        uint32_t mag = __byte_perm(e, 0, 0x4442);                                                    \
        int sgn = __mulhi((int)e, 4);                                                                \
        if (__builtin_expect(M == 0, 0)) {                                                           \
            VC2B_LANE_SETTLE()                                                                       \
            const LaneLong o = lane_read_long_cold(R);                                               \
            R = o.R;                                                                                 \
            mag = o.mag;                                                                             \
            sgn = o.neg ? -1 : 1;                                                                    \
            if (!o.ok) {                                                                             \
                bad |= 1u;                                                                           \
                lane_reader_kill(R);                                                                 \
            }                                                                                        \
        } else {                                                                                     \
            R.buf *= M;                                                                              \
            mark *= M;                                                                               \
        }                                                                                            \
        const uint32_t t = (uint32_t)(((unsigned long long)mag * qf + qo) >> 2);                     \
        maxmag = max(maxmag, mag);                                                                   \
        const int32_t val = (int32_t)t * sgn;                                                        \
        if (b == 0) {                                                                                \
            if (active) {                                                                            \
                int32_t *d32 = (INTERLEAVED && ((j0 + (i)) & 1)) ? out2 : out;                       \
                d32[base0 + (int)(S.map[mc][(j0 + (i)) >> I] & 0x3ffffffu)] = val;                   \
            }                                                                                        \
        } else {                                                                                     \
            t16[((i) >> 1) * TS2 + ((i) & 1)] = (int16_t)val;                                        \
        }                                                                                            \
    }

#define VC2B_LANE_PAIR16(i, m1, s1, m2, s2)                                                          \
    {                                                                                                \
        const int32_t v1 = (int32_t)(((m1) * qf + qo) >> 2) * (s1);                                  \
        const int32_t v2 = (int32_t)(((m2) * qf + qo) >> 2) * (s2);                                  \
        reinterpret_cast<uint32_t *>(t16)[(i) * kLaneTS] = __byte_perm((uint32_t)v1, (uint32_t)v2, 0x5410); \
    }

    // four codes with no band switch among them: two look-ups of two codes each (mark == 1, avail >= 33), or --
    // when some lane's pair does not fit the window or its quantiser is too large -- four one-code steps
#define VC2B_LANE_QUAD16()                                                                           \
    {                                                                                                \
        const uint32_t ea = lds_u32(plut_addr + (((uint32_t)(R.buf >> 32) >> (32 - kPairBits)) << 2)); \
        const uint32_t la = ea & 15u;                                                                \
        const uint64_t bufa = R.buf << la;                                                           \
        const uint32_t eb = lds_u32(plut_addr + (((uint32_t)(bufa >> 32) >> (32 - kPairBits)) << 2)); \
        const uint32_t lb = eb & 15u;                                                                \
        if (__all_sync(0xffffffffu, qfast && la * lb != 0)) {                                        \
            R.buf = bufa << lb;                                                                      \
            R.avail -= (int)(la + lb);                                                               \
            const uint32_t ma1 = __byte_perm(ea, 0, 0x4441), ma2 = __byte_perm(ea, 0, 0x4442);       \
            const uint32_t mb1 = __byte_perm(eb, 0, 0x4441), mb2 = __byte_perm(eb, 0, 0x4442);       \
            VC2B_LANE_PAIR16(0, ma1, __mulhi((int)ea, 4), ma2, __mulhi((int)(ea << 24), 4))          \
            VC2B_LANE_PAIR16(1, mb1, __mulhi((int)eb, 4), mb2, __mulhi((int)(eb << 24), 4))          \
            maxmag = max(max(maxmag, max(ma1, ma2)), max(mb1, mb2));                                 \
        } else {                                                                                     \
            VC2B_LANE_STEP16(0, false)                                                               \
            VC2B_LANE_STEP16(1, false)                                                               \
            VC2B_LANE_STEP16(2, false)                                                               \
            VC2B_LANE_STEP16(3, false)                                                               \
            VC2B_LANE_SETTLE()                                                                       \
        }                                                                                            \
    }

    const int c = lane & 15, h = lane >> 4;   // store phase: half-warp h takes slice 2k+h, lane c its pair c
    for (int j0 = 0; j0 < ncodes; j0 += CH) {
        // ---- decode: lane = slice, CH codes each ---------------------------------------------------
        lane_chunk_start(R);
        if (j0 + CH <= ncodes) {
            if (j0 == next) VC2B_LANE_NEXT_BAND(0)
            if (next - j0 >= CH) {  // one band covers the chunk (most chunks): no band tests inside
#pragma unroll 1
                for (int g = 0; g < CH / 4; g++) {
                    VC2B_LANE_QUAD16()
                    lane_refill(R);
                    j0 += 4;
                    t16 += 2 * TS2;
                }
            } else {
#pragma unroll 1
                for (int g = 0; g < CH / 4; g++) {
                    if (j0 == next) VC2B_LANE_NEXT_BAND(0)
                    if (next - j0 >= 4) {
                        VC2B_LANE_QUAD16()
                    } else {
                        VC2B_LANE_STEP16(0, false)
                        VC2B_LANE_STEP16(1, true)
                        VC2B_LANE_STEP16(2, true)
                        VC2B_LANE_STEP16(3, true)
                        VC2B_LANE_SETTLE()
                    }
                    lane_refill(R);
                    j0 += 4;
                    t16 += 2 * TS2;
                }
            }
            j0 -= CH;
            t16 -= (CH / 2) * TS2;
        } else {
#pragma unroll 1
            for (int i = 0; i < CH; i++) {
                if (j0 < ncodes) VC2B_LANE_STEP16(0, true)
                VC2B_LANE_SETTLE()
                if (R.avail < 32) lane_refill(R);
                j0 += 1;
                t16 += (i & 1) ? TS2 - 1 : 1;
            }
            j0 -= CH;
            t16 -= (CH / 2) * TS2;
        }
        __syncwarp();

        // ---- store ---------------------------------------------------------------------------------------
        if constexpr (INTERLEAVED) {
            // C1 and C2 alternate: tile word p holds (C1[jj], C2[jj]).  A lane takes words 2q and 2q + 1 of a slice
            // -- two adjacent coefficients of each plane -- and writes one 32-bit pair per plane; an instruction
            // covers four slices (lane = q + 8 r, slice 4 k + r), conflict-free on the tile.
            const int q = lane & 7, r = lane >> 3;
            const int jj = (j0 >> 1) + 2 * q;           // first of the lane's two coefficients
            const int ncoef = ncodes >> 1;
            if (jj < ncoef) {
                const bool two = jj + 1 < ncoef;
                const uint32_t loc0 = S.map[mc][jj];
                const uint32_t loc1 = two ? S.map[mc][jj + 1] : loc0;
                const int mb0 = loc0 >> 26, off0 = loc0 & 0x3ffffffu;
                const int mb1 = loc1 >> 26, off1 = loc1 & 0x3ffffffu;
                const uint32_t *tb = reinterpret_cast<const uint32_t *>(&W.tile[(2 * q) * kLaneTS + r]);
                const bool wide = two && mb0 == mb1 && mb0 != 0 && off1 == off0 + 1 && !(off0 & 1) && S.even16[mc][mb0];
                if (wide) {
                    char *p1 = reinterpret_cast<char *>(o16 + off0), *p2 = reinterpret_cast<char *>(o16b + off0);
                    const int32_t *bb = &W.base[r * kLaneBaseStride + mb0];
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        if (act_mask == 0xffffffffu || ((act_mask >> (4 * k + r)) & 1)) {
                            const uint32_t w0 = tb[4 * k], w1 = tb[kLaneTS + 4 * k];
                            const unsigned long long at = 2ull * (uint32_t)bb[4 * k * kLaneBaseStride];
                            *reinterpret_cast<uint32_t *>(p1 + at) = __byte_perm(w0, w1, 0x5410);
                            *reinterpret_cast<uint32_t *>(p2 + at) = __byte_perm(w0, w1, 0x7632);
                        }
                    }
                } else {
#pragma unroll 2
                    for (int k = 0; k < 8; k++) {
                        const int s = 4 * k + r;
                        if (!((act_mask >> s) & 1)) continue;
                        const uint32_t w0 = tb[4 * k], w1 = tb[kLaneTS + 4 * k];
                        if (mb0 != 0) {
                            const int at = W.base[s * kLaneBaseStride + mb0] + off0;
                            o16[at] = (int16_t)(w0 & 0xffffu);
                            o16b[at] = (int16_t)(w0 >> 16);
                        }
                        if (two && mb1 != 0) {
                            const int at = W.base[s * kLaneBaseStride + mb1] + off1;
                            o16[at] = (int16_t)(w1 & 0xffffu);
                            o16b[at] = (int16_t)(w1 >> 16);
                        }
                    }
                }
            }
            __syncwarp();
            continue;
        }
        // two slices per instruction, 16 pairs of consecutive codes each
        const int j = j0 + 2 * c;
        if (j < ncodes) {
            const bool two = j + 1 < ncodes;
            const uint32_t loc0 = S.map[mc][j >> I];
            const uint32_t loc1 = two ? S.map[mc][(j + 1) >> I] : loc0;
            const int mb0 = loc0 >> 26, off0 = loc0 & 0x3ffffffu;
            const int mb1 = loc1 >> 26, off1 = loc1 & 0x3ffffffu;
            int16_t *d0 = o16, *d1 = INTERLEAVED ? o16b : o16;
            const uint32_t *tb = reinterpret_cast<const uint32_t *>(&W.tile[c * kLaneTS + h]);
            const bool wide = !INTERLEAVED && two && mb0 == mb1 && mb0 != 0 && off1 == off0 + 1 && !(off0 & 1) &&
                              S.even16[mc][mb0];
            if (wide) {
                char *p = reinterpret_cast<char *>(d0 + off0);
                const int32_t *bb = &W.base[h * kLaneBaseStride + mb0];
                if (act_mask == 0xffffffffu) {
#pragma unroll
                    for (int k = 0; k < 16; k++)
                        *reinterpret_cast<uint32_t *>(p + 2ull * (uint32_t)bb[2 * k * kLaneBaseStride]) = tb[2 * k];
                } else {
#pragma unroll 4
                    for (int k = 0; k < 16; k++)
                        if ((act_mask >> (2 * k + h)) & 1)
                            *reinterpret_cast<uint32_t *>(p + 2ull * (uint32_t)bb[2 * k * kLaneBaseStride]) = tb[2 * k];
                }
            } else {
#pragma unroll 2
                for (int k = 0; k < 16; k++) {
                    const int s = 2 * k + h;
                    if (!((act_mask >> s) & 1)) continue;
                    const uint32_t word = tb[2 * k];
                    if (mb0 != 0) d0[W.base[s * kLaneBaseStride + mb0] + off0] = (int16_t)(word & 0xffffu);
                    if (two && mb1 != 0) d1[W.base[s * kLaneBaseStride + mb1] + off1] = (int16_t)(word >> 16);
                }
            }
        }
        __syncwarp();
    }
    VC2B_LANE_CLOSE_BAND()
#undef VC2B_LANE_QUAD16
#undef VC2B_LANE_STEP16
#undef VC2B_LANE_PAIR16
#undef VC2B_LANE_NEXT_BAND
#undef VC2B_LANE_SETTLE
#undef VC2B_LANE_CLOSE_BAND
    if ((hiacc >> 1) != 0 || (bad & 3u) != 0) ovf = true;
    if (bad & 4u) ovf16 = true;
}

__device__ __forceinline__ void build_lane_block(const DevParams &P, LaneBlock &S, int tid, int nthreads)
{
    build_lane_lut(S.slut, tid, nthreads);
    build_pair_lut(S.plut, tid, nthreads);
    build_qtab(S.qtab, tid, nthreads);
    for (int idx = tid; idx < 2 * kLaneBands; idx += nthreads) {
        const int mc = idx / kLaneBands, b = idx - mc * kLaneBands;
        int rw = 0, rh = 0, ex = -1, ey = -1;
        if (b < P.nb) {
            const Band &B = P.band[mc][b];
            rw = B.w / P.slices_x;
            rh = B.h / P.slices_y;
            if (rw * P.slices_x == B.w) ex = rw;
            if (rh * P.slices_y == B.h) ey = rh;
        }
        S.ref_w[mc][b] = rw;
        S.ref_h[mc][b] = rh;
        S.ref_cnt[mc][b] = rw * rh;
        S.even16[mc][b] = (b < P.nb && !(P.band[mc][b].off & 1) && !(P.band[mc][b].w & 1) && !(rw & 1)) ? 1 : 0;
        S.even_x[mc][b] = ex;
        S.even_y[mc][b] = ey;
    }
    __syncthreads();
    if (tid < 2) {
        int n = 0;
        for (int b = 0; b < P.nb; b++) n += S.ref_w[tid][b] * S.ref_h[tid][b];
        S.ref_ncoef[tid] = n;
    }
    __syncthreads();
    for (int idx = tid; idx < 2 * kLaneMaxCoef; idx += nthreads) {
        const int mc = idx / kLaneMaxCoef, j = idx - mc * kLaneMaxCoef;
        uint32_t loc = 0;
        if (j < S.ref_ncoef[mc]) {
            int b = 0, cum = 0;
            while (j >= cum + S.ref_w[mc][b] * S.ref_h[mc][b]) {
                cum += S.ref_w[mc][b] * S.ref_h[mc][b];
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

// One LANE per slice; a warp task is 32 consecutive slices of one picture.
template <bool LD, bool C16>
__global__ void __launch_bounds__(kLaneWarps * 32, 1)
unpack_lanes_kernel(DevParams P, const uint8_t *__restrict__ data, const int64_t *__restrict__ pic_offset,
                    int npictures, const uint32_t *__restrict__ slice_offset, int32_t *__restrict__ coeffs,
                    int16_t *__restrict__ coeffs16, int32_t *__restrict__ qindex_out, int32_t *__restrict__ status)
{
    extern __shared__ __align__(16) uint8_t lane_smem[];
    LaneBlock &S = *reinterpret_cast<LaneBlock *>(lane_smem);
    LaneWarp *Ws = reinterpret_cast<LaneWarp *>(lane_smem + sizeof(LaneBlock));
    build_lane_block(P, S, threadIdx.x, blockDim.x);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    LaneWarp &W = Ws[warp];
    const int nslices = P.slices_x * P.slices_y;
    const int tasks_per_pic = (nslices + 31) >> 5;
    const long long total_tasks = (long long)npictures * tasks_per_pic;
    for (long long t = (long long)blockIdx.x * kLaneWarps + warp; t < total_tasks;
         t += (long long)gridDim.x * kLaneWarps) {
        const int pic = (int)(t / tasks_per_pic);
        const int n = (int)(t - (long long)pic * tasks_per_pic) * 32 + lane;
        bool active = n < nslices;
        const int sy = n / P.slices_x, sx = n - sy * P.slices_x;
        const long long gw = (long long)pic * nslices + n;
        const uint8_t *base = data + pic_offset[pic];
        int32_t *pc = coeffs + (size_t)pic * P.pic_coeffs;
        int16_t *pc16 = C16 ? coeffs16 + (size_t)pic * P.pic_coeffs : nullptr;
        uint32_t maxabs = 0;
        bool ovf = false, ovf16 = false;

        if (!LD) {
            // hq_slice (decoder/transform_data_syntax.py:212)
            const uint32_t off = active ? slice_offset[gw] : 0xffffffffu;
            active = active && off != 0xffffffffu;  // at or after the slice where the stream ended
            long long pos = (long long)off + P.slice_prefix_bytes;
            const int qindex = active ? (int)__ldg(base + pos) : 0;
            pos += 1;
            if (qindex_out && active) qindex_out[gw] = qindex;
#pragma unroll 1
            for (int c = 0; c < 3; c++) {
                const long long length = active ? (long long)P.slice_size_scaler * __ldg(base + pos) : 0;
                pos += 1;
                if constexpr (C16)
                    lanes_block16<false>(P, c, W, S, active, sx, sy, qindex, base + pos, 0, 8 * length,
                                         pc + P.comp_off[c], nullptr, pc16 + P.comp_off[c], nullptr, lane, maxabs, ovf,
                                         ovf16);
                else
                    lanes_block<false>(P, c, W, S, active, sx, sy, qindex, base + pos, 0, 8 * length,
                                       pc + P.comp_off[c], nullptr, lane, maxabs, ovf);
                pos += length;
            }
        } else {
            // ld_slice (decoder/transform_data_syntax.py:145)
            const long long nbytes = pic_offset[pic + 1] - pic_offset[pic];
            long long s0 = 0, s1 = 0;
            if (active) {
                s0 = ld_slice_offset(P, n);
                s1 = ld_slice_offset(P, n + 1);
                if (s1 > nbytes) active = false;  // stream ended inside this slice (ld_init_status_kernel)
            }
            int qindex = 0, length_bits = 0;
            long long slice_y_length = 0, bits_left = 0;
            if (active) {
                const long long sbytes = s1 - s0;
                // 7-bit qindex, then intlog2(8*slice_bytes-7)-bit slice_y_length (decoder/io.py:230)
                length_bits = ilog2_ceil(8 * sbytes - 7);
                long long hdr = 0;
                const int hbytes = (7 + length_bits + 7) >> 3;
                for (int i = 0; i < hbytes; i++) hdr = (hdr << 8) | __ldg(base + s0 + i);
                hdr >>= (8 * hbytes - 7 - length_bits);
                qindex = (int)(hdr >> length_bits);
                slice_y_length = hdr & ((1LL << length_bits) - 1);
                bits_left = 8 * sbytes - 7 - length_bits;
                if (qindex_out) qindex_out[gw] = qindex;
                if (slice_y_length > bits_left) {  // InvalidSliceYLength, transform_data_syntax.py:166
                    report(status, pic, n, VC2B_ST_INVALID_SLICE_Y_LENGTH);
                    active = false;
                }
            }
            const long long ybit = s0 * 8 + 7 + length_bits;
            if constexpr (C16)
                lanes_block16<false>(P, 0, W, S, active, sx, sy, qindex, base + (ybit >> 3), (int)(ybit & 7),
                                     slice_y_length, pc + P.comp_off[0], nullptr, pc16 + P.comp_off[0], nullptr, lane,
                                     maxabs, ovf, ovf16);
            else
                lanes_block<false>(P, 0, W, S, active, sx, sy, qindex, base + (ybit >> 3), (int)(ybit & 7),
                                   slice_y_length, pc + P.comp_off[0], nullptr, lane, maxabs, ovf);
            const long long cbit = ybit + slice_y_length;
            if constexpr (C16)
                lanes_block16<true>(P, 1, W, S, active, sx, sy, qindex, base + (cbit >> 3), (int)(cbit & 7),
                                    bits_left - slice_y_length, pc + P.comp_off[1], pc + P.comp_off[2],
                                    pc16 + P.comp_off[1], pc16 + P.comp_off[2], lane, maxabs, ovf, ovf16);
            else
                lanes_block<true>(P, 1, W, S, active, sx, sy, qindex, base + (cbit >> 3), (int)(cbit & 7),
                                  bits_left - slice_y_length, pc + P.comp_off[1], pc + P.comp_off[2], lane, maxabs,
                                  ovf);
        }
        ovf = __any_sync(0xffffffffu, ovf);
        if (C16) {  // status word 1 is free until finalize_status_kernel: high-pass value outside int16
            if (__any_sync(0xffffffffu, ovf16) && lane == 0) atomicOr(&status[4 * pic + 1], 1);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) maxabs = max(maxabs, __shfl_xor_sync(0xffffffffu, maxabs, d));
        if (lane == 0 && (maxabs || ovf))
            atomicMax(&status[4 * pic + 3], (ovf || maxabs > 0x7fffffffu) ? 0x7fffffff : (int)maxabs);
    }
}

constexpr size_t kLaneSmemBytes = sizeof(LaneBlock) + kLaneWarps * sizeof(LaneWarp);
static_assert(sizeof(LaneBlock) % 16 == 0, "LaneWarp array alignment");

// Whether every slice component fits the lane-per-slice kernel.
static bool lanes_eligible(const DevParams &P)
{
    if (P.nb > kLaneBands) return false;
    for (int c = 0; c < 3; c++) {
        long long tot = 0;
        for (int b = 0; b < P.nb; b++) {
            const long long cw = (P.band[c][b].w + P.slices_x - 1) / P.slices_x;
            const long long ch = (P.band[c][b].h + P.slices_y - 1) / P.slices_y;
            if (cw > 0xffff) return false;
            tot += cw * ch;
            if (tot > kLaneMaxCoef) return false;
        }
    }
    return true;
}

// vc2b_unpack16 additionally needs every band to divide evenly between the slices
// (slices_have_same_dimensions, pseudocode/slice_sizes.py:131) and even component offsets.
static bool lanes16_eligible(const DevParams &P)
{
    if (!lanes_eligible(P)) return false;
    if (P.pic_coeffs & 1) return false;
    for (int c = 0; c < 3; c++) {
        if (P.comp_off[c] & 1) return false;
        for (int b = 0; b < P.nb; b++)
            if (P.band[c][b].w % P.slices_x || P.band[c][b].h % P.slices_y) return false;
    }
    return true;
}

}  // namespace vc2b
