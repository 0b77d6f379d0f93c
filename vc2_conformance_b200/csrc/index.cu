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
// This is an INDEXER, not a validator: it performs none of the reference's conformance checks, and the
// tables of the un-vendored vc2_data_tables package (base video formats, default quantisation
// matrices) are not restated -- a stream that relies on them gets a flag instead of a guess.
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
    case 1: *luma = 255; *chroma = 255; return true;    // 8-bit full range
    case 2: *luma = 219; *chroma = 224; return true;    // 8-bit video
    case 3: *luma = 876; *chroma = 896; return true;    // 10-bit video
    case 4: *luma = 3504; *chroma = 3584; return true;  // 12-bit video
    default: return false;
    }
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
    if (s.frame_width < 0 || s.color_diff_format_index < 0 || s.luma_excursion < 0)
        s.flags |= VC2B_IDX_NEEDS_BASE_FORMAT;  // the base video format supplies what the stream does not override
}

// picture_dimensions + video_depth (pseudocode/video_parameters.py:177-206)
void fill_dimensions(const SeqState &s, vc2b_params *p)
{
    if (!s.have || (s.flags & VC2B_IDX_NEEDS_BASE_FORMAT)) return;
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
    p->slices_x = (int32_t)r.uint();  // slice_parameters (:175)
    p->slices_y = (int32_t)r.uint();
    p->slice_size_scaler = 1;
    p->slice_bytes_denominator = 1;
    if (ld) {
        p->slice_bytes_numerator = (int64_t)r.uint();
        p->slice_bytes_denominator = (int64_t)r.uint();
    } else {
        p->slice_prefix_bytes = (int32_t)r.uint();
        p->slice_size_scaler = (int32_t)r.uint();
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
    } else {
        *flags |= VC2B_IDX_NEEDS_DEFAULT_QUANT_MATRIX;  // vc2_data_tables.QUANTISATION_MATRICES
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
                const int64_t ns = (int64_t)u.params.slices_x * u.params.slices_y;
                int64_t p = u.data_offset;
                if (u.params.is_ld) {
                    p += (ns * u.params.slice_bytes_numerator) / (u.params.slice_bytes_denominator > 0
                                                                      ? u.params.slice_bytes_denominator : 1);
                } else {
                    for (int64_t i = 0; i < ns && p < len; i++) {  // hq_slice (transform_data_syntax.py:212)
                        p += u.params.slice_prefix_bytes + 1;
                        for (int c = 0; c < 3 && p < len; c++) p += 1 + (int64_t)data[p] * u.params.slice_size_scaler;
                    }
                }
                off = p;
                continue;
            }
            break;  // nothing says where the next data unit starts
        }
        off += next;
    }
    return n;
}

}  // extern "C"
