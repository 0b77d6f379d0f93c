/*
 * vc2_b200.h -- C ABI of the B200-native VC-2 picture path (libvc2b200.so).
 *
 * This is the drop-in boundary for the data-parallel picture path of
 * bbc/vc2_conformance.  The reference has no FFI of its own (it is pure
 * Python); each entry point below replaces the Python function(s) cited next
 * to it (paths relative to /root/reference/vc2_conformance/), and
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference adds.
 *
 * Conventions
 *  - every `d_*` pointer is a DEVICE pointer; everything else is host memory;
 *  - every call only ENQUEUES work on `stream` (a cudaStream_t passed as
 *    void*) and returns immediately; results are valid after the stream is
 *    synchronised;
 *  - return value: 0 on success, VC2B_E_* (<0) for argument errors, or a
 *    positive cudaError_t from the launch;
 *  - the library never allocates device memory: callers size buffers with the
 *    vc2b_*_count / vc2b_*_bytes helpers;
 *  - a "batch" is `npictures` pictures coded with identical parameters.
 *
 * Numeric envelope: the reference computes in unbounded Python integers.  This
 * library stores coefficients as int32 and reports VC2B_ST_INT32_ENVELOPE (per
 * picture) if a decoded or dequantised value does not fit, instead of wrapping.
 */
#ifndef VC2_B200_H
#define VC2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VC2B_MAX_LEVELS 16 /* dwt_depth + dwt_depth_ho + 1 <= 16 */
#define VC2B_MAX_BANDS 32  /* 1 + dwt_depth_ho + 3*dwt_depth <= 32 */

/* Picture-path parameters: the subset of the reference's `State`
 * (pseudocode/state.py:22-390) that the picture path reads.
 * quant_matrix[level][slot]: slot 0 = LL/L (level 0), 1 = HL or H, 2 = LH, 3 = HH
 * (State["quant_matrix"], pseudocode/state.py:120). */
typedef struct vc2b_params {
    int32_t luma_width, luma_height;             /* pseudocode/video_parameters.py:177 */
    int32_t color_diff_width, color_diff_height;
    int32_t luma_depth, color_diff_depth;        /* pseudocode/video_parameters.py:202 */
    int32_t wavelet_index, wavelet_index_ho;     /* decoder/picture_syntax.py:94-173 */
    int32_t dwt_depth, dwt_depth_ho;
    int32_t slices_x, slices_y;                  /* decoder/picture_syntax.py:175 */
    int32_t is_ld;                               /* parse code 0xC8 (1) or 0xE8 (0) */
    int32_t slice_prefix_bytes, slice_size_scaler; /* HQ */
    int64_t slice_bytes_numerator, slice_bytes_denominator; /* LD */
    int32_t quant_matrix[VC2B_MAX_LEVELS][4];
} vc2b_params;

/* argument errors */
#define VC2B_E_BAD_PARAMS (-1)
#define VC2B_E_UNSUPPORTED (-2)
#define VC2B_E_SCRATCH_TOO_SMALL (-3)

/* per-picture status words written by the unpack kernels (d_status[4*i+0]) */
#define VC2B_ST_OK 0
#define VC2B_ST_UNEXPECTED_END_OF_STREAM 1 /* decoder/io.py:197 UnexpectedEndOfStream */
#define VC2B_ST_INVALID_SLICE_Y_LENGTH 2   /* decoder/transform_data_syntax.py:166 */
#define VC2B_ST_INT32_ENVELOPE 4           /* value outside the int32 envelope */
#define VC2B_ST_COEFF16_RANGE 5            /* vc2b_unpack16 only: a high-pass coefficient does not fit int16;
                                              decode this picture with vc2b_unpack + vc2b_idwt instead */

/* ---- introspection --------------------------------------------------------------- */
const char *vc2b_version(void);
/* Number of CUDA kernels this library has launched in this process so far (every launch site
 * counts; bench.py reports the difference across its timed region as gpu_launches). */
int64_t vc2b_launch_count(void);
/* Checks a parameter set; 0 if the library supports it. */
int vc2b_check_params(const vc2b_params *p);
/* Subband geometry (pseudocode/slice_sizes.py:53-92).  Subbands are numbered in bitstream
 * order: level 0 (LL or L), H of each horizontal-only level, then HL,LH,HH per 2-D level. */
int vc2b_num_subbands(const vc2b_params *p);
int vc2b_subband_info(const vc2b_params *p, int comp, int band, int32_t *level, int32_t *slot,
                      int32_t *width, int32_t *height, int64_t *offset);
/* Coefficients (int32) of one component / one whole picture (Y,C1,C2 consecutive). */
int64_t vc2b_component_coeff_count(const vc2b_params *p, int comp);
int64_t vc2b_picture_coeff_count(const vc2b_params *p);
/* Padded (transform) dimensions and decoded sample count of a picture (Y,C1,C2 planes). */
int vc2b_padded_dims(const vc2b_params *p, int comp, int32_t *width, int32_t *height);
int64_t vc2b_picture_sample_count(const vc2b_params *p);
/* Largest |coefficient| for which the int32 inverse transform provably cannot overflow (the
 * bound propagates magnitudes through every lifting stage).  vc2b_unpack reports
 * VC2B_ST_INT32_ENVELOPE for a picture whose largest decoded coefficient exceeds it. */
int64_t vc2b_idwt_max_safe_coeff(const vc2b_params *p);
/* Bytes of scratch one picture needs in vc2b_idwt / vc2b_dwt (intermediate LL bands). */
int64_t vc2b_transform_scratch_bytes(const vc2b_params *p);

/* ---- decoder ------------------------------------------------------------------------ */

/* HQ slice offsets: the dependent chain hidden in the sequential reader of
 * hq_slice (decoder/transform_data_syntax.py:212-251; each slice's start is known only after
 * the three length bytes of the previous one).
 *   d_data        transform-data bytes of all pictures, back to back; 16-byte aligned (the kernels read
 *                 aligned 16-byte pieces) and followed by at least 16 readable padding bytes
 *   d_pic_offset  [npictures+1] int64: byte offset of each picture's transform data in d_data;
 *                 entry i+1 bounds picture i (reads past it are UnexpectedEndOfStream)
 *   d_slice_offset [npictures*slices] uint32 out: start of each slice relative to its picture
 *   d_status      [npictures*4] int32 out: {status, first bad slice, bytes consumed, max |coeff|}
 */
int vc2b_hq_slice_offsets(const vc2b_params *p, const uint8_t *d_data, const int64_t *d_pic_offset,
                          int npictures, uint32_t *d_slice_offset, int32_t *d_status, void *stream);

/* transform_data (decoder/transform_data_syntax.py:116): slice unpack (hq_slice :212 / ld_slice
 * :145, slice_band :286, color_diff_slice_band :315, read_sintb decoder/io.py:307), fused
 * inverse_quant (pseudocode/quantization.py:18) and, for LD, dc_prediction (:63).
 *   d_slice_offset from vc2b_hq_slice_offsets (ignored for LD: offsets are closed-form)
 *   d_coeffs      [npictures * vc2b_picture_coeff_count] int32 out
 *   d_qindex      [npictures*slices] int32 out (per-slice qindex, for the level-constraint
 *                 checks the host keeps: transform_data_syntax.py:154,222), may be NULL
 *   d_status      as above; d_status must have been initialised by vc2b_hq_slice_offsets for HQ;
 *                 for LD this call initialises it.
 */
int vc2b_unpack(const vc2b_params *p, const uint8_t *d_data, const int64_t *d_pic_offset,
                int npictures, const uint32_t *d_slice_offset, int32_t *d_coeffs,
                int32_t *d_qindex, int32_t *d_status, void *stream);

/* The same call with the INT16 COEFFICIENT PLANE: inverse-quantised coefficients of every subband but the
 * level-0 band are stored as int16 in d_coeffs16 (same element offsets as d_coeffs; only the level-0 band
 * -- which dc_prediction and the first synthesis level read as int32 -- is written to d_coeffs).  Halves the
 * bytes the coefficients cost on their way through HBM (the path is bandwidth-bound: DESIGN.md 5).  Values are
 * the same integers; a picture with a high-pass coefficient outside int16 gets VC2B_ST_COEFF16_RANGE and must
 * be decoded again through vc2b_unpack (device.BatchCodec does).  Needs vc2b_coeff16_supported(p). */
int vc2b_unpack16(const vc2b_params *p, const uint8_t *d_data, const int64_t *d_pic_offset,
                  int npictures, const uint32_t *d_slice_offset, int32_t *d_coeffs, int16_t *d_coeffs16,
                  int32_t *d_qindex, int32_t *d_status, void *stream);
/* 1 if the int16 coefficient plane is available for this parameter set: lane-per-slice unpack geometry and
 * every transform level on the wide synthesis kernel (2-D levels only, one filter, vector-aligned bands). */
int vc2b_coeff16_supported(const vc2b_params *p);
/* Test hook: which slice-unpack kernel vc2b_unpack uses.  0 = chosen by geometry (default), 1 = warp per
 * slice, 2 = lane per slice (VC2B_E_UNSUPPORTED from vc2b_unpack if the geometry does not allow it).
 * Process-wide; the GPU tests run both kernels on the same streams. */
int vc2b_set_unpack_kernel(int kernel);

/* dc_prediction (decoder/transform_data_syntax.py:63) of the level-0 band of every component,
 * in place.  vc2b_unpack already applies it for LD; exposed for the function-level API. */
int vc2b_dc_prediction(const vc2b_params *p, int32_t *d_coeffs, int npictures, int inverse,
                       void *stream);

/* picture_decode (pseudocode/picture_decoding.py:45): idwt :88 (h_synthesis :134, vh_synthesis
 * :159, lift1-4 :205-261, filter_bit_shift :198), idwt_pad_removal :282, clip_picture :327,
 * offset_picture :301.
 *   d_pictures  [npictures * vc2b_picture_sample_count] uint16 out: planes Y, C1, C2, each
 *               width x height row-major, values in [0, 2^depth)
 *   d_scratch   scratch_bytes >= vc2b_transform_scratch_bytes(p); more lets more pictures be
 *               in flight per launch */
int vc2b_idwt(const vc2b_params *p, const int32_t *d_coeffs, int npictures, uint16_t *d_pictures,
              void *d_scratch, int64_t scratch_bytes, void *stream);
/* picture_decode from the int16 coefficient plane of vc2b_unpack16 (d_coeffs supplies the level-0 bands).
 * The intermediate bands between the levels are int16 as well (half the scratch bytes of vc2b_idwt suffice);
 * a picture with an intermediate value outside int16 gets d_status[4*i] = VC2B_ST_COEFF16_RANGE (only over
 * VC2B_ST_OK) and must be decoded again with vc2b_unpack + vc2b_idwt.  d_status: the words vc2b_unpack16 wrote. */
int vc2b_idwt16(const vc2b_params *p, const int32_t *d_coeffs, const int16_t *d_coeffs16, int npictures,
                uint16_t *d_pictures, void *d_scratch, int64_t scratch_bytes, int32_t *d_status, void *stream);
/* idwt (:88) alone for ONE component: no crop, clip or offset; int32 out, padded dimensions. */
int vc2b_idwt_raw(const vc2b_params *p, int comp, const int32_t *d_coeffs, int32_t *d_out,
                  void *d_scratch, int64_t scratch_bytes, void *stream);

/* ---- encoder mirror ------------------------------------------------------------- */

/* picture_encode (pseudocode/picture_encoding.py:45): remove_offset_picture :264,
 * dwt_pad_addition :235, dwt :93 (h_analysis :141, vh_analysis :169).  Layouts as above. */
int vc2b_dwt(const vc2b_params *p, const uint16_t *d_pictures, int npictures, int32_t *d_coeffs,
             void *d_scratch, int64_t scratch_bytes, void *stream);
/* dwt (:93) alone for ONE component from an int32 padded picture (consumed). */
int vc2b_dwt_raw(const vc2b_params *p, int comp, const int32_t *d_in, int32_t *d_coeffs,
                 void *d_scratch, int64_t scratch_bytes, void *stream);

/* quantize_to_fit (encoder/pictures.py:295) for every slice of every picture, as driven by
 * make_transform_data_hq_lossy (:617, align 8*slice_size_scaler, target from picture_bytes) or
 * make_transform_data_ld_lossy (:724, align 1, target from slice_bytes): the smallest
 * qindex >= minimum_qindex whose quantised (forward_quant, quantization.py:30) slice fits,
 * sizes by calculate_coeffs_bits (:217).  For LD the caller applies dc prediction first
 * (vc2b_dc_prediction with inverse=1).
 *   picture_bytes  HQ only (LD uses slice_bytes_numerator/denominator)
 *   d_qindex       [npictures*slices] int32 out
 *   d_slice_bits   [npictures*slices*3] int32 out: bits of the Y, C1, C2 (LD: Y, C, 0) blocks at
 *                  the chosen qindex (calculate_coeffs_bits), may be NULL */
int vc2b_quantize_to_fit(const vc2b_params *p, const int32_t *d_coeffs, int npictures,
                         int64_t picture_bytes, int minimum_qindex, int32_t *d_qindex,
                         int32_t *d_slice_bits, void *stream);

/* calculate_coeffs_bits (encoder/pictures.py:217) of the Y, C1, C2 (LD: Y, C, 0) blocks of every slice,
 * quantised (quantize_coeffs :251) at `qindex`, or at d_qindex_in[slice] when that is not NULL -- no
 * search.  d_qindex_out [npictures*slices] receives the index used, d_slice_bits [npictures*slices*3]
 * the sizes.  With qindex 0 this is the size pass of make_transform_data_hq_lossless (:528). */
int vc2b_slice_bits(const vc2b_params *p, const int32_t *d_coeffs, int npictures, int qindex,
                    const int32_t *d_qindex_in, int32_t *d_qindex_out, int32_t *d_slice_bits, void *stream);

/* quantize_coeffs (encoder/pictures.py:251) for whole pictures: forward_quant (quantization.py:30) of
 * every coefficient at max(qindex of the slice that holds it - quant_matrix value, 0), the per-slice
 * index taken from d_qindex [npictures*slices] (or the constant `qindex` when d_qindex is NULL).
 * d_out has the coefficient layout of d_coeffs (may not alias it). */
int vc2b_quantize_slices(const vc2b_params *p, const int32_t *d_coeffs, int npictures, const int32_t *d_qindex,
                         int qindex, int32_t *d_out, void *stream);

/* quantize_to_fit (encoder/pictures.py:295) on explicit lists, the function-level form the reference's
 * encoder calls per slice: `nsets` (<= 8) coefficient sets stored back to back in d_values
 * (set s = [set_offsets[s], set_offsets[s+1]), set_offsets is HOST memory), one quantisation matrix value
 * per coefficient.  d_result[0] = the smallest qindex >= minimum_qindex whose blocks, each padded to
 * align_bits, total at most target_size bits; d_result[1+s] = calculate_coeffs_bits (:217) of set s at
 * that index; d_quantised = the quantised values (quantize_coeffs :251).  With target_size = INT64_MAX
 * and minimum_qindex = q it evaluates quantize_coeffs / calculate_coeffs_bits at q. */
int vc2b_quantize_to_fit_lists(const int32_t *d_values, const int32_t *d_quant_matrix_values,
                               const int32_t *set_offsets, int nsets, int64_t target_size, int64_t align_bits,
                               int minimum_qindex, int32_t *d_result, int32_t *d_quantised, void *stream);

/* Slice packing: the serialisation of hq_slice / ld_slice (bitstream/vc2.py:727-839) with
 * write_sint (bitstream/io.py:575) of the coefficients quantised at d_qindex, producing the
 * same bytes as make_transform_data_hq_lossy/ld_lossy + autofill_and_serialise_stream.
 *   d_out        [npictures * out_stride] bytes out (zero prefix bytes for HQ)
 *   out_stride   >= vc2b_packed_bytes(p, picture_bytes), a multiple of 4; d_out 4-byte aligned */
int64_t vc2b_packed_bytes(const vc2b_params *p, int64_t picture_bytes);
int vc2b_pack_slices(const vc2b_params *p, const int32_t *d_coeffs, int npictures,
                     int64_t picture_bytes, const int32_t *d_qindex, const int32_t *d_slice_bits,
                     uint8_t *d_out, int64_t out_stride, void *stream);

/* make_transform_data_hq_lossless (encoder/pictures.py:528): every length field is the smallest
 * multiple of slice_size_scaler bytes that holds its block, so slices have different sizes.
 * vc2b_hq_lossless_layout turns the block sizes (vc2b_slice_bits at qindex 0) into
 *   d_slice_offset [npictures*(slices+1)] uint32: byte offset of every slice in its picture's transform
 *                  data; entry [slices] is the total size
 *   d_max_length   [npictures] int32: the largest length field at p->slice_size_scaler (the reference
 *                  picks the scaler from this value at scaler 1: max(1, minimum, (max + 254) // 255)), may be NULL
 * and vc2b_pack_slices_lossless writes the slices (same bytes as the reference's serialiser). */
int vc2b_hq_lossless_layout(const vc2b_params *p, const int32_t *d_slice_bits, int npictures,
                            uint32_t *d_slice_offset, int32_t *d_max_length, void *stream);
int vc2b_pack_slices_lossless(const vc2b_params *p, const int32_t *d_coeffs, int npictures, const int32_t *d_qindex,
                              const int32_t *d_slice_bits, const uint32_t *d_slice_offset, uint8_t *d_out,
                              int64_t out_stride, void *stream);

/* ---- element-wise pseudocode functions ---------------------------------------------- */
/* inverse_quant / forward_quant (pseudocode/quantization.py:18,30) with per-element index. */
int vc2b_inverse_quant(const int32_t *d_in, const int32_t *d_qi, int32_t *d_out, int64_t n,
                       void *stream);
int vc2b_forward_quant(const int32_t *d_in, const int32_t *d_qi, int32_t *d_out, int64_t n,
                       void *stream);
/* clip_component + offset_component (pseudocode/picture_decoding.py:301-349) on int32 data:
 * mode bit0 = clip, bit1 = add offset, bit2 = subtract offset (remove_offset, :264). */
int vc2b_clip_offset(int32_t *d_data, int64_t n, int depth, int mode, void *stream);

/* ---- raw picture files and picture comparison --------------------- */
/* Bytes of one picture in the reference's raw planar file format: components Y, C1, C2, little
 * endian, ceil(depth / 8) rounded up to a power of two bytes per sample
 * (file_format.py:158-194, dimensions_and_depths.py:46-86). */
int64_t vc2b_raw_picture_bytes(const vc2b_params *p);
/* file_format.write_picture for npictures device pictures: d_raw[i * raw_stride ...] gets picture i. */
int vc2b_pictures_to_raw(const vc2b_params *p, const uint16_t *d_pictures, int npictures, uint8_t *d_raw,
                         int64_t raw_stride, void *stream);
/* file_format.read_picture (:263-314), including the masking of the bits above the sample depth. */
int vc2b_raw_to_pictures(const vc2b_params *p, const uint8_t *d_raw, int64_t raw_stride, int npictures,
                         uint16_t *d_pictures, void *stream);
/* scripts/vc2_picture_compare.py:370-420 measure_differences: d_result[(i * 3 + c) * 2 + {0, 1}] =
 * {samples of component c of picture i that differ, sum of their squared differences}. */
int vc2b_picture_compare(const vc2b_params *p, const uint16_t *d_a, const uint16_t *d_b, int npictures,
                         uint64_t *d_result, void *stream);

/* ---- stream index (host code, no device work) ------------------------ */
/* vc2b_unit.flags */
#define VC2B_IDX_NEEDS_BASE_FORMAT 1           /* dimensions / depths come from the base video format table */
#define VC2B_IDX_NEEDS_DEFAULT_QUANT_MATRIX 2  /* custom_quant_matrix is false: default matrix table */
#define VC2B_IDX_NO_SEQUENCE_HEADER 4          /* picture before any sequence header */
#define VC2B_IDX_TRUNCATED 8                   /* the stream ended inside this data unit's header */
#define VC2B_IDX_UNSUPPORTED 16                /* more transform levels than VC2B_MAX_LEVELS */

/* One data unit of a VC-2 stream (parse_info, decoder/stream.py:194): where it starts, and for
 * pictures (parse codes 0xC8 / 0xE8) and fragments (0xCC / 0xEC) the picture-path parameters
 * (picture_syntax.py:94-348, fragment_syntax.py:55-140, video_parameters.py:177-206) and the byte
 * offset of the transform data (pictures) or of the fragment's slices that vc2b_hq_slice_offsets /
 * vc2b_unpack take. */
typedef struct vc2b_unit {
    int64_t offset;        /* of the parse_info prefix 0x42424344 */
    int64_t next_offset;   /* offset + next_parse_offset, 0 if the stream gives none */
    int64_t data_offset;   /* transform data / slice data, -1 if the unit has none */
    int32_t parse_code;
    int32_t flags;         /* VC2B_IDX_* */
    uint32_t picture_number;
    int32_t fragment_data_length, fragment_slice_count, fragment_x_offset, fragment_y_offset;
    int32_t major_version, minor_version, profile, level, base_video_format, picture_coding_mode;
    vc2b_params params;    /* pictures and fragments; zero where a flag says the stream does not define it */
} vc2b_unit;

/* Walks the data units of `data` (HOST memory) from offset 0 (parse_sequence, decoder/stream.py:96-138)
 * following next_parse_offset (for pictures without one: the slice lengths).  Writes up to max_units
 * entries and returns the number of data units found (which may exceed max_units), or VC2B_E_BAD_PARAMS
 * if the buffer does not start with a parse_info header.  Performs no conformance checks. */
int64_t vc2b_index_stream(const uint8_t *data, int64_t len, vc2b_unit *units, int64_t max_units);

#ifdef __cplusplus
}
#endif
#endif /* VC2_B200_H */
