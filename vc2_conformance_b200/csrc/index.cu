// index.cu -- host-side index of a VC-2 byte stream: where every data unit starts,
// and for every picture / fragment the parameters and the byte offset the device kernels need.
//
// Restates, in C, the header syntax the reference reads bit by bit in Python: parse_sequence /
// parse_info (decoder/stream.py:194-308), sequence_header / parse_parameters / source_parameters
// (decoder/sequence_header.py:118-800), picture_parse / picture_header / transform_parameters /
// extended_transform_parameters / slice_parameters / quant_matrix (decoder/picture_syntax.py:67-348),
// fragment_parse / fragment_header (decoder/fragment_syntax.py:44-140), picture_dimensions /
// video_depth (pseudocode/video_parameters.py:177-206) and the bit reader they use (decoder/io.py:
// 187-243 read_bit / read_bool / read_nbits / read_uint / byte_align).
//
// This is an INDEXER, not a validator: it performs none of the reference's conformance checks.  The tables
// of the un-vendored vc2_data_tables package that the picture path depends on -- base video formats
// (frame size, colour difference format, signal range: SMPTE ST 2042-1 Annex C / Table 11.x) and the
// default quantisation matrices (Annex D) -- are restated from the standard below; a stream that needs an
// entry this file does not carry gets a flag instead of a guess.  Header fields are untrusted input: every
// value that sizes or positions a read is range-checked before it is used.
#include <string.h>

#include "common.cuh"

namespace {

struct BitReader {
    const uint8_t *data;
    int64_t len;
    int64_t pos;  // bit position
    bool eof;

    int bit()
    {
        if ((pos >> 3) >= len) { eof = true; return 1; }
        const int b = (data[pos >> 3] >> (7 - (pos & 7))) & 1;
        pos++;
        return b;
    }
    void byte_align() { pos = (pos + 7) & ~(int64_t)7; }
    uint64_t nbits(int n)
    {
        uint64_t v = 0;
        for (int i = 0; i < n; i++) v = (v << 1) | (uint64_t)bit();
        return v;
    }
    uint64_t uint_lit(int nbytes) { return nbits(8 * nbytes); }
    bool boolean() { return bit() != 0; }
    // read_uint (A.4.3): interleaved exp-Golomb; saturates instead of growing without bound
    uint64_t uint()
    {
        uint64_t value = 1;
        while (bit() == 0) {
            if (eof) break;
            value = (value << 1) | (uint64_t)bit();
            if (value >> 62) { eof = true; break; }
        }
        return value - 1;
    }
};

// intlog2 (pseudocode/vc2_math.py:32)
int intlog2_u64(uint64_t n) { return vc2b::ilog2_ceil((int64_t)n); }

struct SeqState {
    bool have;
    int32_t major_version, minor_version, profile, level, base_video_format, picture_coding_mode;
    int64_t frame_width, frame_height, color_diff_format_index, luma_excursion, color_diff_excursion;
    int32_t flags;  // VC2B_IDX_NEEDS_* accumulated while reading the header
};

// preset_signal_range (11.4.9): the four presets of SMPTE ST 2042-1 Table 11.5 as excursions
// (vc2_data_tables.PRESET_SIGNAL_RANGES; restated from the standard)
bool preset_excursions(uint64_t index, int64_t *luma, int64_t *chroma)
{
    switch (index) {
    case 1: *luma = 255; *chroma = 255; return true;      // 8-bit full range
    case 2: *luma = 219; *chroma = 224; return true;      // 8-bit video
    case 3: *luma = 876; *chroma = 896; return true;      // 10-bit video
    case 4: *luma = 3504; *chroma = 3584; return true;    // 12-bit video
    case 5: *luma = 1023; *chroma = 1023; return true;    // 10-bit full range
    case 6: *luma = 4095; *chroma = 4095; return true;    // 12-bit full range
    case 7: *luma = 56064; *chroma = 57344; return true;  // 16-bit video
    case 8: *luma = 65535; *chroma = 65535; return true;  // 16-bit full range
    default: return false;
    }
}

// Base video formats (vc2_data_tables.BASE_VIDEO_FORMAT_PARAMETERS; SMPTE ST 2042-1 Annex C): the three
// columns the picture path needs -- frame size, colour difference format index, signal range preset.
struct BaseFormat { int32_t width, height, cdf, signal_range; };
const BaseFormat kBaseFormats[23] = {
    {640, 480, 2, 1},    // 0  custom format
    {176, 120, 2, 1},    // 1  QSIF525
    {176, 144, 2, 1},    // 2  QCIF
    {352, 240, 2, 1},    // 3  SIF525
    {352, 288, 2, 1},    // 4  CIF
    {704, 480, 2, 1},    // 5  4SIF525
    {704, 576, 2, 1},    // 6  4CIF
    {720, 480, 1, 3},    // 7  SD480I-60
    {720, 576, 1, 3},    // 8  SD576I-50
    {1280, 720, 1, 3},   // 9  HD720P-60
    {1280, 720, 1, 3},   // 10 HD720P-50
    {1920, 1080, 1, 3},  // 11 HD1080I-60
    {1920, 1080, 1, 3},  // 12 HD1080I-50
    {1920, 1080, 1, 3},  // 13 HD1080P-60
    {1920, 1080, 1, 3},  // 14 HD1080P-50
    {2048, 1080, 0, 4},  // 15 DC2K
    {4096, 2160, 0, 4},  // 16 DC4K
    {3840, 2160, 1, 3},  // 17 UHDTV 4K-60
    {3840, 2160, 1, 3},  // 18 UHDTV 4K-50
    {7680, 4320, 1, 3},  // 19 UHDTV 8K-60
    {7680, 4320, 1, 3},  // 20 UHDTV 8K-50
    {1920, 1080, 1, 3},  // 21 HD1080P-24
    {720, 486, 1, 3},    // 22 SD Pro486
};

// Default quantisation matrices (vc2_data_tables.QUANTISATION_MATRICES; SMPTE ST 2042-1 Annex D) of the
// symmetric transforms (wavelet_index == wavelet_index_ho, dwt_depth_ho == 0, dwt_depth <= 4) and of
// Deslauriers-Dubuc (9,7) with horizontal-only levels (the tables quoted in the reference tree:
// util/parse_default_quantisation_matrix_table.py:10-55).  Returns false where no table is carried.
bool default_quant_matrix(vc2b_params *p)
{
    const int w = p->wavelet_index, d = p->dwt_depth, dho = p->dwt_depth_ho;
    if (w != p->wavelet_index_ho || w < 0 || w > 6 || d < 0 || dho < 0) return false;
    memset(p->quant_matrix, 0, sizeof(p->quant_matrix));
    if (dho == 0) {
        if (d > 4) return false;
        // rows: HL = LH, HH of 2-D level 1..4; LL of depth >= 1
        static const int ll[7] = {5, 4, 5, 0 /* Haar, no shift: 4 * (depth + 1) */, 8, 0, 3};
        static const int hl[7][4] = {{3, 4, 5, 6}, {2, 4, 5, 7}, {3, 4, 5, 6}, {0, 0, 0, 0}, {4, 4, 4, 4},
                                     {4, 8, 13, 17}, {1, 4, 6, 9}};
        static const int hh[7][4] = {{0, 1, 2, 3}, {0, 2, 3, 5}, {0, 1, 2, 3}, {0, 0, 0, 0}, {0, 0, 0, 0},
                                     {8, 12, 17, 21}, {0, 2, 5, 7}};
        if (d == 0) return true;  // a single LL band: 0
        if (w == 3) {             // Haar without shift: every level adds 4 from the top down
            p->quant_matrix[0][0] = 4 * (d + 1);
            for (int l = 1; l <= d; l++) {
                p->quant_matrix[l][1] = p->quant_matrix[l][2] = 4 * (d - l + 1);
                p->quant_matrix[l][3] = 4 * (d - l);
            }
            return true;
        }
        p->quant_matrix[0][0] = ll[w];
        for (int l = 1; l <= d; l++) {
            p->quant_matrix[l][1] = p->quant_matrix[l][2] = hl[w][l - 1];
            p->quant_matrix[l][3] = hh[w][l - 1];
        }
        return true;
    }
    if (w != 0 || dho > 4 || d > 5 - dho || (dho == 1 && d > 4)) return false;
    static const int h_levels[4] = {0, 3, 5, 8};
    static const int hl_ho[5][4] = {{0, 0, 0, 0}, {3, 4, 5, 6}, {5, 6, 7, 0}, {8, 9, 0, 0}, {10, 0, 0, 0}};
    static const int hh_ho[5][4] = {{0, 0, 0, 0}, {0, 1, 2, 3}, {3, 4, 5, 0}, {5, 6, 0, 0}, {8, 0, 0, 0}};
    p->quant_matrix[0][0] = 3;
    for (int l = 1; l <= dho; l++) p->quant_matrix[l][1] = h_levels[l - 1];
    for (int l = 1; l <= d; l++) {
        p->quant_matrix[dho + l][1] = p->quant_matrix[dho + l][2] = hl_ho[dho][l - 1];
        p->quant_matrix[dho + l][3] = hh_ho[dho][l - 1];
    }
    return true;
}

void read_sequence_header(BitReader &r, SeqState &s)
{
    s.have = true;
    s.flags = 0;
    s.major_version = (int32_t)r.uint();
    s.minor_version = (int32_t)r.uint();
    s.profile = (int32_t)r.uint();
    s.level = (int32_t)r.uint();
    s.base_video_format = (int32_t)r.uint();
    s.frame_width = s.frame_height = s.color_diff_format_index = -1;
    s.luma_excursion = s.color_diff_excursion = -1;
    if (r.boolean()) {  // frame_size
        s.frame_width = (int64_t)r.uint();
        s.frame_height = (int64_t)r.uint();
    }
    if (r.boolean()) s.color_diff_format_index = (int64_t)r.uint();  // color_diff_sampling_format
    if (r.boolean()) (void)r.uint();                                 // scan_format: source_sampling
    if (r.boolean()) {                                               // frame_rate
        if (r.uint() == 0) { (void)r.uint(); (void)r.uint(); }
    }
    if (r.boolean()) {  // pixel_aspect_ratio
        if (r.uint() == 0) { (void)r.uint(); (void)r.uint(); }
    }
    if (r.boolean()) { (void)r.uint(); (void)r.uint(); (void)r.uint(); (void)r.uint(); }  // clean_area
    if (r.boolean()) {                                                                    // signal_range
        const uint64_t index = r.uint();
        if (index == 0) {
            (void)r.uint();
            s.luma_excursion = (int64_t)r.uint();
            (void)r.uint();
            s.color_diff_excursion = (int64_t)r.uint();
        } else if (!preset_excursions(index, &s.luma_excursion, &s.color_diff_excursion)) {
            s.flags |= VC2B_IDX_NEEDS_BASE_FORMAT;
        }
    }
    if (r.boolean()) {  // color_spec
        if (r.uint() == 0) {
            if (r.boolean()) (void)r.uint();  // color_primaries
            if (r.boolean()) (void)r.uint();  // color_matrix
            if (r.boolean()) (void)r.uint();  // transfer_function
        }
    }
    s.picture_coding_mode = (int32_t)r.uint();
    // the base video format supplies what the stream does not override (set_source_defaults,
    // pseudocode/video_parameters.py:133)
    if (s.frame_width < 0 || s.color_diff_format_index < 0 || s.luma_excursion < 0) {
        if (s.base_video_format < 0 || s.base_video_format > 22) {
            s.flags |= VC2B_IDX_NEEDS_BASE_FORMAT;  // an index the table does not carry
        } else {
            const BaseFormat &b = kBaseFormats[s.base_video_format];
            if (s.frame_width < 0) { s.frame_width = b.width; s.frame_height = b.height; }
            if (s.color_diff_format_index < 0) s.color_diff_format_index = b.cdf;
            if (s.luma_excursion < 0) preset_excursions((uint64_t)b.signal_range, &s.luma_excursion, &s.color_diff_excursion);
        }
    }
    // untrusted sizes: anything the picture path cannot represent is flagged, not used
    if (s.frame_width < 1 || s.frame_height < 1 || s.frame_width > (1 << 20) || s.frame_height > (1 << 20) ||
        s.color_diff_format_index > 2 || s.luma_excursion < 1 || s.color_diff_excursion < 1 ||
        s.luma_excursion > 0xffffffffLL || s.color_diff_excursion > 0xffffffffLL)
        s.flags |= VC2B_IDX_UNSUPPORTED;
}

// picture_dimensions + video_depth (pseudocode/video_parameters.py:177-206)
void fill_dimensions(const SeqState &s, vc2b_params *p)
{
    if (!s.have || (s.flags & (VC2B_IDX_NEEDS_BASE_FORMAT | VC2B_IDX_UNSUPPORTED))) return;
    int64_t lw = s.frame_width, lh = s.frame_height, cw = lw, ch = lh;
    if (s.color_diff_format_index == 1) cw /= 2;                 // 4:2:2
    if (s.color_diff_format_index == 2) { cw /= 2; ch /= 2; }    // 4:2:0
    if (s.picture_coding_mode == 1) { lh /= 2; ch /= 2; }        // pictures are fields
    p->luma_width = (int32_t)lw;
    p->luma_height = (int32_t)lh;
    p->color_diff_width = (int32_t)cw;
    p->color_diff_height = (int32_t)ch;
    p->luma_depth = intlog2_u64((uint64_t)s.luma_excursion + 1);
    p->color_diff_depth = intlog2_u64((uint64_t)s.color_diff_excursion + 1);
}

// transform_parameters (picture_syntax.py:94) for a picture (or the first fragment) of parse code `code`
void read_transform_parameters(BitReader &r, const SeqState &s, int code, vc2b_params *p, int32_t *flags)
{
    const bool ld = (code == 0xC8 || code == 0xCC);
    p->is_ld = ld ? 1 : 0;
    p->wavelet_index = (int32_t)r.uint();
    p->dwt_depth = (int32_t)r.uint();
    p->wavelet_index_ho = p->wavelet_index;
    p->dwt_depth_ho = 0;
    if (s.major_version >= 3) {  // extended_transform_parameters (:125)
        if (r.boolean()) p->wavelet_index_ho = (int32_t)r.uint();
        if (r.boolean()) p->dwt_depth_ho = (int32_t)r.uint();
    }
    // slice_parameters (:175).  Untrusted: values that do not fit the fields, or that would make the slice
    // walk below overflow, mark the unit unsupported (the casts alone would let them go negative)
    const uint64_t kMaxField = 0x7fffffffull;
    uint64_t v[4];
    v[0] = r.uint();
    v[1] = r.uint();
    v[2] = r.uint();
    v[3] = r.uint();
    p->slice_size_scaler = 1;
    p->slice_bytes_denominator = 1;
    bool bad = v[0] < 1 || v[1] < 1 || v[0] > kMaxField || v[1] > kMaxField || v[0] * v[1] > kMaxField ||
               (uint64_t)p->wavelet_index > 6 || (uint64_t)p->wavelet_index_ho > 6;
    if (ld) bad = bad || v[3] < 1 || v[2] > (1ull << 40) || v[3] > (1ull << 40);  // slices * numerator < 2^71 / 2^8
    else bad = bad || v[2] > (1u << 24) || v[3] < 1 || v[3] > (1u << 24);
    if (bad) {
        *flags |= VC2B_IDX_UNSUPPORTED;
    } else {
        p->slices_x = (int32_t)v[0];
        p->slices_y = (int32_t)v[1];
        if (ld) {
            p->slice_bytes_numerator = (int64_t)v[2];
            p->slice_bytes_denominator = (int64_t)v[3];
        } else {
            p->slice_prefix_bytes = (int32_t)v[2];
            p->slice_size_scaler = (int32_t)v[3];
        }
    }
    if (r.boolean()) {  // quant_matrix (:267), custom
        const int dho = p->dwt_depth_ho, d = p->dwt_depth;
        if (dho < 0 || d < 0 || dho + d + 1 > VC2B_MAX_LEVELS) { *flags |= VC2B_IDX_UNSUPPORTED; return; }
        p->quant_matrix[0][0] = (int32_t)r.uint();  // LL, or L when there are horizontal-only levels
        for (int level = 1; level <= dho; level++) p->quant_matrix[level][1] = (int32_t)r.uint();  // H
        for (int level = dho + 1; level <= dho + d; level++) {
            p->quant_matrix[level][1] = (int32_t)r.uint();  // HL
            p->quant_matrix[level][2] = (int32_t)r.uint();  // LH
            p->quant_matrix[level][3] = (int32_t)r.uint();  // HH
        }
    } else if (p->dwt_depth < 0 || p->dwt_depth_ho < 0 || p->dwt_depth_ho + p->dwt_depth + 1 > VC2B_MAX_LEVELS) {
        *flags |= VC2B_IDX_UNSUPPORTED;
    } else if (!default_quant_matrix(p)) {
        *flags |= VC2B_IDX_NEEDS_DEFAULT_QUANT_MATRIX;  // a vc2_data_tables.QUANTISATION_MATRICES entry not carried here
    }
}

}  // namespace

extern "C" {

int64_t vc2b_index_stream(const uint8_t *data, int64_t len, vc2b_unit *units, int64_t max_units)
{
    if (!data || len < 0 || (max_units > 0 && !units)) return VC2B_E_BAD_PARAMS;
    SeqState seq;
    memset(&seq, 0, sizeof(seq));
    vc2b_params frag_params;  // transform parameters of the fragmented picture in progress
    memset(&frag_params, 0, sizeof(frag_params));
    int32_t frag_flags = 0;
    int64_t n = 0, off = 0;
    while (off + 13 <= len) {
        BitReader r = {data, len, off * 8, false};
        const uint64_t prefix = r.uint_lit(4);
        if (prefix != 0x42424344u) return n > 0 ? n : VC2B_E_BAD_PARAMS;  // BadParseInfoPrefix: stop indexing here
        const int code = (int)r.uint_lit(1);
        const int64_t next = (int64_t)r.uint_lit(4);
        (void)r.uint_lit(4);  // previous_parse_offset
        vc2b_unit u;
        memset(&u, 0, sizeof(u));
        u.offset = off;
        u.parse_code = code;
        u.next_offset = next ? off + next : 0;
        u.data_offset = -1;
        const bool picture = (code == 0xC8 || code == 0xE8), fragment = (code == 0xCC || code == 0xEC);
        if (code == 0x00) {
            read_sequence_header(r, seq);
            u.flags |= seq.flags;
        } else if (picture) {
            r.byte_align();
            u.picture_number = (uint32_t)r.uint_lit(4);  // picture_header (:76)
            r.byte_align();
            read_transform_parameters(r, seq, code, &u.params, &u.flags);
            fill_dimensions(seq, &u.params);
            r.byte_align();
            u.data_offset = r.pos >> 3;
            u.flags |= seq.flags;
        } else if (fragment) {
            r.byte_align();
            u.picture_number = (uint32_t)r.uint_lit(4);  // fragment_header (fragment_syntax.py:55)
            u.fragment_data_length = (int32_t)r.uint_lit(2);
            u.fragment_slice_count = (int32_t)r.uint_lit(2);
            if (u.fragment_slice_count == 0) {
                r.byte_align();
                memset(&frag_params, 0, sizeof(frag_params));
                frag_flags = 0;
                read_transform_parameters(r, seq, code, &frag_params, &frag_flags);
                fill_dimensions(seq, &frag_params);
            } else {
                u.fragment_x_offset = (int32_t)r.uint_lit(2);
                u.fragment_y_offset = (int32_t)r.uint_lit(2);
                r.byte_align();
                u.data_offset = r.pos >> 3;
            }
            u.params = frag_params;
            u.flags |= frag_flags | seq.flags;
        }
        if (seq.have) {
            u.major_version = seq.major_version;
            u.minor_version = seq.minor_version;
            u.profile = seq.profile;
            u.level = seq.level;
            u.base_video_format = seq.base_video_format;
            u.picture_coding_mode = seq.picture_coding_mode;
        } else if (picture || fragment) {
            u.flags |= VC2B_IDX_NO_SEQUENCE_HEADER;
        }
        if (r.eof) u.flags |= VC2B_IDX_TRUNCATED;
        if (n < max_units) units[n] = u;
        n++;
        if (code == 0x10) {  // end of sequence: another sequence may follow immediately (10.3)
            off += 13;
            memset(&seq, 0, sizeof(seq));
            continue;
        }
        if (next == 0) {
            // Without a next_parse_offset the end of a picture is only known by walking its slices.
            if (picture && !(u.flags & (VC2B_IDX_TRUNCATED | VC2B_IDX_UNSUPPORTED)) && u.params.slices_x > 0 &&
                u.params.slices_y > 0) {
                // slices_x * slices_y < 2^31, numerator <= 2^40, prefix and scaler <= 2^24 (checked when they
                // were read): every product below fits 128 / 64 bits and p only ever grows
                const int64_t ns = (int64_t)u.params.slices_x * u.params.slices_y;
                int64_t p = u.data_offset;
                if (u.params.is_ld) {
                    const unsigned __int128 total = ((unsigned __int128)ns * (unsigned __int128)u.params.slice_bytes_numerator) /
                                                    (unsigned __int128)u.params.slice_bytes_denominator;
                    p = total > (unsigned __int128)(len - p) ? len : p + (int64_t)total;
                } else {
                    for (int64_t i = 0; i < ns && p < len; i++) {  // hq_slice (transform_data_syntax.py:212)
                        p += (int64_t)u.params.slice_prefix_bytes + 1;
                        for (int c = 0; c < 3 && p < len; c++) p += 1 + (int64_t)data[p] * u.params.slice_size_scaler;
                    }
                    if (p > len) p = len;
                }
                if (p <= off || p < u.data_offset) break;  // no progress: nothing says where the next unit starts
                off = p;
                continue;
            }
            break;  // nothing says where the next data unit starts
        }
        if (next < 13 || next > len - off) break;  // a next_parse_offset that does not lead to another header
        off += next;
    }
    return n;
}

}  // extern "C"
