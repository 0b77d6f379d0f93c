/*
 * vc2_oracle.h -- CPU restatement of the VC-2 picture path of bbc/vc2_conformance.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is the checker for the CUDA path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product package
 * (vc2_conformance_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  Every function here is checked against the
 * unmodified Python reference (imported from /root/reference through
 * oracle/ref_import.py) by tools/make_golden.py; the resulting input/output
 * vectors are committed under tests/golden/ and re-checked by
 * tests/test_oracle_golden.py.  One caveat, stated in DESIGN.md: the lifting
 * tap table itself lives in the un-vendored `vc2_data_tables` package
 * (setup.py:48, >=0.1.1,<2.0) and is restated from SMPTE ST 2042-1 Table 15.x;
 * the reference's own tests do not pin tap values.
 *
 * Arithmetic: the reference computes in unbounded Python ints; this port
 * computes in int64_t, which is exact for every input whose intermediates fit
 * 63 bits (all streams the GPU path accepts; it stores int32).
 *
 * All `file:line` citations are relative to /root/reference/vc2_conformance/.
 */
#ifndef VC2_ORACLE_H
#define VC2_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VC2O_MAX_LEVELS 16 /* dwt_depth + dwt_depth_ho + 1 <= 16 */

/* Orientation slots inside quant_matrix[level][]: level 0 uses slot 0 (LL or
 * L); horizontal-only levels use slot 1 (H); 2-D levels use 1,2,3 = HL,LH,HH. */
typedef struct {
    int32_t luma_width, luma_height;             /* pseudocode/video_parameters.py:177 */
    int32_t color_diff_width, color_diff_height;
    int32_t luma_depth, color_diff_depth;        /* pseudocode/video_parameters.py:202 */
    int32_t wavelet_index, wavelet_index_ho;
    int32_t dwt_depth, dwt_depth_ho;
    int32_t slices_x, slices_y;
    int32_t is_ld;                               /* 1: LD profile (0xC8), 0: HQ (0xE8) */
    int32_t slice_prefix_bytes, slice_size_scaler; /* HQ */
    int64_t slice_bytes_numerator, slice_bytes_denominator; /* LD */
    int32_t quant_matrix[VC2O_MAX_LEVELS][4];
} vc2o_params;

/* status codes */
enum {
    VC2O_OK = 0,
    VC2O_UNEXPECTED_END_OF_STREAM = 1, /* decoder/io.py:197 */
    VC2O_INVALID_SLICE_Y_LENGTH = 2,   /* decoder/transform_data_syntax.py:166 */
    VC2O_BAD_PARAMS = 3,
    VC2O_INT64_ENVELOPE = 4
};

/* ---- maths (pseudocode/vc2_math.py:32) ---- */
int32_t vc2o_intlog2(int64_t n);

/* ---- geometry (pseudocode/slice_sizes.py:53-127) ---- */
int32_t vc2o_subband_width(const vc2o_params *p, int level, int comp);
int32_t vc2o_subband_height(const vc2o_params *p, int level, int comp);
int64_t vc2o_slice_bytes(const vc2o_params *p, int sx, int sy);
/* number of subbands of a component and, for subband index b in bitstream
 * order (level 0, then H levels, then HL,LH,HH per 2-D level), its level,
 * orientation slot, dimensions and offset (in coefficients) inside the flat
 * per-component coefficient array.  Returns total coefficients. */
int32_t vc2o_num_subbands(const vc2o_params *p);
int64_t vc2o_subband_layout(const vc2o_params *p, int comp, int32_t *level,
                            int32_t *orient, int32_t *w, int32_t *h,
                            int64_t *offset);

/* ---- quantisation (pseudocode/quantization.py:18-62) ---- */
int64_t vc2o_quant_factor(int index);
int64_t vc2o_quant_offset(int index);
int64_t vc2o_inverse_quant(int64_t q, int index);
int64_t vc2o_forward_quant(int64_t c, int index);

/* ---- bounded-block exp-Golomb reader (decoder/io.py:187-313) ---- */
/* Decodes `n` signed values from a bounded block of `bits_left` bits that
 * starts at bit offset `bit_offset` (MSB first) of `data`.  Returns status;
 * *bits_consumed receives the number of stream bits actually read. */
int vc2o_read_sintb_block(const uint8_t *data, size_t nbytes, uint64_t bit_offset,
                          int64_t bits_left, int64_t *out, size_t n,
                          uint64_t *bits_consumed);
/* exp-Golomb writer + length (bitstream/io.py:557-583, bitstream/exp_golomb.py:28) */
int32_t vc2o_signed_exp_golomb_length(int64_t v);

/* ---- transform data unpack (decoder/transform_data_syntax.py:116-335) ---- */
/* coeffs[c] : flat int64 arrays laid out by vc2o_subband_layout.
 * qindex_out / slice_len_out : per slice (raster), may be NULL.
 * *consumed : bytes of `data` consumed by transform_data.
 * *bad_slice : index of the offending slice for status 1/2. */
int vc2o_transform_data(const vc2o_params *p, const uint8_t *data, size_t nbytes,
                        int64_t *coeffs_y, int64_t *coeffs_c1, int64_t *coeffs_c2,
                        int32_t *qindex_out, int32_t *slice_len_out,
                        size_t *consumed, int32_t *bad_slice);
void vc2o_dc_prediction(int64_t *band, int32_t w, int32_t h);       /* :63 */
void vc2o_apply_dc_prediction(int64_t *band, int32_t w, int32_t h); /* encoder/pictures.py:202 */

/* ---- synthesis (pseudocode/picture_decoding.py:45-349) ---- */
int vc2o_filter_bit_shift(int filter); /* picture_decoding.py:198 */
void vc2o_oned_synthesis(int64_t *a, int64_t stride, int64_t len, int filter);
void vc2o_oned_analysis(int64_t *a, int64_t stride, int64_t len, int filter);
/* idwt of one component: coeffs (flat) -> out (padded_w x padded_h, row major) */
int vc2o_idwt(const vc2o_params *p, int comp, const int64_t *coeffs, int64_t *out);
/* picture_decode for one component: idwt + pad removal + clip + offset.
 * out: width x height (cropped) row major uint16 (depth <= 16). */
int vc2o_decode_component(const vc2o_params *p, int comp, const int64_t *coeffs,
                          uint16_t *out);

/* ---- analysis (pseudocode/picture_encoding.py:45-286) ---- */
int vc2o_dwt(const vc2o_params *p, int comp, int64_t *padded_picture, int64_t *coeffs);
/* picture_encode for one component: remove offset, edge pad, dwt. */
int vc2o_encode_component(const vc2o_params *p, int comp, const uint16_t *pic,
                          int64_t *coeffs);

/* ---- encoder slice maths (encoder/pictures.py:217-347, 456-792) ---- */
int64_t vc2o_calculate_coeffs_bits(const int64_t *coeffs, size_t n);
/* Gather one slice's coefficients of one component in bitstream order
 * (encoder/pictures.py:418-451); returns count.  qm_out gets the per-coefficient
 * quantisation matrix value. */
int64_t vc2o_gather_slice(const vc2o_params *p, int comp, const int64_t *coeffs,
                          int sx, int sy, int64_t *out, int32_t *qm_out);
/* HQ lossy transform data (make_transform_data_hq_lossy :617 + serialisation
 * bitstream/vc2.py:798): writes slices to out (capacity cap); returns bytes
 * written or -1.  p->slice_size_scaler is honoured as given.  DC prediction is
 * not applied (HQ).  qindex_out per slice may be NULL. */
int64_t vc2o_encode_hq_lossy(const vc2o_params *p, const int64_t *cy,
                             const int64_t *cc1, const int64_t *cc2,
                             int64_t picture_bytes, int minimum_qindex,
                             uint8_t *out, size_t cap, int32_t *qindex_out);
/* HQ with fixed qindex for every slice (lossless when qindex==0), minimal
 * lengths (make_transform_data_hq_lossless :528 with the given scaler). */
int64_t vc2o_encode_hq_fixed(const vc2o_params *p, const int64_t *cy,
                             const int64_t *cc1, const int64_t *cc2, int qindex,
                             uint8_t *out, size_t cap);
/* LD lossy transform data (make_transform_data_ld_lossy :724).  The caller
 * applies vc2o_apply_dc_prediction to the DC bands first.  Uses
 * p->slice_bytes_numerator/denominator. */
int64_t vc2o_encode_ld_lossy(const vc2o_params *p, const int64_t *cy,
                             const int64_t *cc1, const int64_t *cc2,
                             int minimum_qindex, uint8_t *out, size_t cap,
                             int32_t *qindex_out);
/* quantize_to_fit on explicit coefficient sets (encoder/pictures.py:295):
 * nsets sets, set i has n[i] values coeffs[i][] with matrix values qm[i][].
 * Returns the chosen qindex; quantised values are written to out[i][]. */
int vc2o_quantize_to_fit(int64_t target_size, int nsets, const int64_t *const *coeffs,
                         const int32_t *const *qm, const size_t *n, int64_t align_bits,
                         int minimum_qindex, int64_t *const *out);

#ifdef __cplusplus
}
#endif
#endif
