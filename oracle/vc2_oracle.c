/*
 * vc2_oracle.c -- CPU restatement of the VC-2 picture path of bbc/vc2_conformance.
 *
 * TEST INFRASTRUCTURE ONLY (see vc2_oracle.h).  Plain C99, int64 arithmetic,
 * written to follow the reference function by function; every function cites
 * the reference file:line it restates (paths relative to
 * /root/reference/vc2_conformance/).  Parity is pinned by tests/golden/ (.npz files),
 * produced by tools/make_golden.py from the unmodified Python reference.
 */
#include "vc2_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* small maths (pseudocode/vc2_math.py:32-60)                                */
/* ------------------------------------------------------------------------ */

/* floor division for a positive divisor (Python `//`) */
static int64_t floordiv(int64_t a, int64_t b)
{
    int64_t q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) q -= 1;
    return q;
}

/* pseudocode/vc2_math.py:32  intlog2(n) = (n-1).bit_length() */
int32_t vc2o_intlog2(int64_t n)
{
    int32_t bits = 0;
    uint64_t v;
    if (n <= 1) return 0;
    v = (uint64_t)(n - 1);
    while (v) { bits++; v >>= 1; }
    return bits;
}

/* pseudocode/vc2_math.py:57  mean(*S) = (sum + n//2) // n, here n == 3 */
static int64_t mean3(int64_t a, int64_t b, int64_t c)
{
    return floordiv(a + b + c + 1, 3);
}

/* ------------------------------------------------------------------------ */
/* geometry (pseudocode/slice_sizes.py:53-127)                               */
/* ------------------------------------------------------------------------ */

static int32_t comp_width(const vc2o_params *p, int comp)
{
    return comp == 0 ? p->luma_width : p->color_diff_width;
}
static int32_t comp_height(const vc2o_params *p, int comp)
{
    return comp == 0 ? p->luma_height : p->color_diff_height;
}
static int32_t comp_depth(const vc2o_params *p, int comp)
{
    return comp == 0 ? p->luma_depth : p->color_diff_depth;
}

/* slice_sizes.py:53 */
int32_t vc2o_subband_width(const vc2o_params *p, int level, int comp)
{
    int32_t w = comp_width(p, comp);
    int32_t scale_w = 1 << (p->dwt_depth_ho + p->dwt_depth);
    int32_t pw = scale_w * ((w + scale_w - 1) / scale_w);
    if (level == 0) return pw / (1 << (p->dwt_depth_ho + p->dwt_depth));
    return pw / (1 << (p->dwt_depth_ho + p->dwt_depth - level + 1));
}

/* slice_sizes.py:74 */
int32_t vc2o_subband_height(const vc2o_params *p, int level, int comp)
{
    int32_t h = comp_height(p, comp);
    int32_t scale_h = 1 << p->dwt_depth;
    int32_t ph = scale_h * ((h + scale_h - 1) / scale_h);
    if (level == 0 || level <= p->dwt_depth_ho) return ph / (1 << p->dwt_depth);
    return ph / (1 << (p->dwt_depth_ho + p->dwt_depth - level + 1));
}

/* slice_sizes.py:95 */
int64_t vc2o_slice_bytes(const vc2o_params *p, int sx, int sy)
{
    int64_t n = (int64_t)sy * p->slices_x + sx;
    int64_t b = ((n + 1) * p->slice_bytes_numerator) / p->slice_bytes_denominator;
    b -= (n * p->slice_bytes_numerator) / p->slice_bytes_denominator;
    return b;
}

/* slice_sizes.py:108-127 */
static int32_t slice_left(const vc2o_params *p, int sx, int comp, int level)
{
    return (int32_t)(((int64_t)vc2o_subband_width(p, level, comp) * sx) / p->slices_x);
}
static int32_t slice_right(const vc2o_params *p, int sx, int comp, int level)
{
    return (int32_t)(((int64_t)vc2o_subband_width(p, level, comp) * (sx + 1)) / p->slices_x);
}
static int32_t slice_top(const vc2o_params *p, int sy, int comp, int level)
{
    return (int32_t)(((int64_t)vc2o_subband_height(p, level, comp) * sy) / p->slices_y);
}
static int32_t slice_bottom(const vc2o_params *p, int sy, int comp, int level)
{
    return (int32_t)(((int64_t)vc2o_subband_height(p, level, comp) * (sy + 1)) / p->slices_y);
}

/* Subbands in bitstream order (decoder/transform_data_syntax.py:176-189,
 * 230-243): level 0 (LL or L), H of each horizontal-only level, then
 * HL, LH, HH of each 2-D level. */
int32_t vc2o_num_subbands(const vc2o_params *p)
{
    return 1 + p->dwt_depth_ho + 3 * p->dwt_depth;
}

int64_t vc2o_subband_layout(const vc2o_params *p, int comp, int32_t *level,
                            int32_t *orient, int32_t *w, int32_t *h, int64_t *offset)
{
    int b = 0, lv, o;
    int64_t off = 0;
    int top = p->dwt_depth_ho + p->dwt_depth;
    for (lv = 0; lv <= top; lv++) {
        int no = (lv == 0) ? 1 : (lv <= p->dwt_depth_ho ? 1 : 3);
        for (o = 0; o < no; o++) {
            int32_t bw = vc2o_subband_width(p, lv, comp);
            int32_t bh = vc2o_subband_height(p, lv, comp);
            if (level) level[b] = lv;
            if (orient) orient[b] = (lv == 0) ? 0 : o + 1;
            if (w) w[b] = bw;
            if (h) h[b] = bh;
            if (offset) offset[b] = off;
            off += (int64_t)bw * bh;
            b++;
        }
    }
    return off;
}

/* ------------------------------------------------------------------------ */
/* quantisation (pseudocode/quantization.py:18-62)                           */
/* ------------------------------------------------------------------------ */

/* quantization.py:40 ; saturates (INT64_MAX) once the exact value needs more
 * than 62 bits -- such a factor quantises every int64 coefficient to zero. */
int64_t vc2o_quant_factor(int index)
{
    int k = index / 4;
    int64_t base;
    if (index < 0) return 4;
    if (k > 40) return INT64_MAX;
    base = (int64_t)1 << k;
    switch (index % 4) {
    case 0: return 4 * base;
    case 1: return ((503829 * base) + 52958) / 105917;
    case 2: return ((665857 * base) + 58854) / 117708;
    default: return ((440253 * base) + 32722) / 65444;
    }
}

/* quantization.py:54 */
int64_t vc2o_quant_offset(int index)
{
    int64_t qf;
    if (index == 0) return 1;
    if (index == 1) return 2;
    qf = vc2o_quant_factor(index);
    if (qf == INT64_MAX) return INT64_MAX / 2;
    return (qf + 1) / 2;
}

/* quantization.py:18 */
int64_t vc2o_inverse_quant(int64_t q, int index)
{
    int64_t m = q < 0 ? -q : q;
    if (m != 0) {
        m *= vc2o_quant_factor(index);
        m += vc2o_quant_offset(index);
        m += 2;
        m /= 4;
    }
    return q < 0 ? -m : m;
}

/* quantization.py:30 */
int64_t vc2o_forward_quant(int64_t c, int index)
{
    int64_t m = c < 0 ? -c : c;
    int64_t qf = vc2o_quant_factor(index);
    int64_t r = (qf == INT64_MAX) ? 0 : (4 * m) / qf;
    return c >= 0 ? r : -r;
}

/* ------------------------------------------------------------------------ */
/* bit reader (decoder/io.py:152-313)                                        */
/* ------------------------------------------------------------------------ */

typedef struct {
    const uint8_t *data;
    uint64_t nbits_total; /* 8 * nbytes */
    uint64_t pos;         /* next bit to read, MSB first */
    int64_t bits_left;    /* bounded block (decoder/io.py:246) */
    int status;
} reader_t;

/* decoder/io.py:187 read_bit; raises UnexpectedEndOfStream at :197 */
static int read_bit(reader_t *r)
{
    int bit;
    if (r->pos >= r->nbits_total) {
        r->status = VC2O_UNEXPECTED_END_OF_STREAM;
        return 0;
    }
    bit = (r->data[r->pos >> 3] >> (7 - (r->pos & 7))) & 1;
    r->pos++;
    return bit;
}

/* decoder/io.py:230 */
static int64_t read_nbits(reader_t *r, int n)
{
    int64_t v = 0;
    int i;
    for (i = 0; i < n && !r->status; i++) v = (v << 1) + read_bit(r);
    return v;
}

/* decoder/io.py:246 */
static int read_bitb(reader_t *r)
{
    if (r->bits_left == 0) return 1;
    r->bits_left--;
    return read_bit(r);
}

/* decoder/io.py:265 */
static void flush_inputb(reader_t *r)
{
    while (r->bits_left > 0 && !r->status) {
        read_bit(r);
        r->bits_left--;
    }
}

/* decoder/io.py:285 read_uintb + :307 read_sintb */
static int64_t read_sintb(reader_t *r)
{
    int64_t value = 1;
    while (read_bitb(r) == 0 && !r->status) {
        if (value >= ((int64_t)1 << 61)) {
            r->status = VC2O_INT64_ENVELOPE;
            return 0;
        }
        value <<= 1;
        if (read_bitb(r)) value += 1;
    }
    value -= 1;
    if (value != 0) {
        if (read_bitb(r) == 1) value = -value;
    }
    return value;
}

int vc2o_read_sintb_block(const uint8_t *data, size_t nbytes, uint64_t bit_offset,
                          int64_t bits_left, int64_t *out, size_t n,
                          uint64_t *bits_consumed)
{
    reader_t r;
    size_t i;
    r.data = data;
    r.nbits_total = 8 * (uint64_t)nbytes;
    r.pos = bit_offset;
    r.bits_left = bits_left;
    r.status = VC2O_OK;
    for (i = 0; i < n && !r.status; i++) out[i] = read_sintb(&r);
    if (bits_consumed) *bits_consumed = r.pos - bit_offset;
    return r.status;
}

/* bitstream/exp_golomb.py:8-36 */
int32_t vc2o_signed_exp_golomb_length(int64_t v)
{
    uint64_t m = (uint64_t)(v < 0 ? -v : v) + 1;
    int32_t bl = 0;
    while (m) { bl++; m >>= 1; }
    return (bl - 1) * 2 + 1 + (v != 0 ? 1 : 0);
}

/* ------------------------------------------------------------------------ */
/* transform data unpack (decoder/transform_data_syntax.py:63-335)           */
/* ------------------------------------------------------------------------ */

typedef struct {
    int nb;
    int32_t level[3 * VC2O_MAX_LEVELS], orient[3 * VC2O_MAX_LEVELS];
    int32_t w[3][3 * VC2O_MAX_LEVELS], h[3][3 * VC2O_MAX_LEVELS];
    int64_t off[3][3 * VC2O_MAX_LEVELS];
    int64_t total[3];
} layout_t;

static void make_layout(const vc2o_params *p, layout_t *L)
{
    int c;
    L->nb = vc2o_num_subbands(p);
    for (c = 0; c < 3; c++)
        L->total[c] = vc2o_subband_layout(p, c, L->level, L->orient, L->w[c], L->h[c], L->off[c]);
}

/* transform_data_syntax.py:255 slice_quantizers */
static int slice_quantizer(const vc2o_params *p, int qindex, int level, int orient)
{
    int q = qindex - p->quant_matrix[level][orient];
    return q > 0 ? q : 0;
}

/* transform_data_syntax.py:286 slice_band */
static void slice_band(const vc2o_params *p, const layout_t *L, reader_t *r, int64_t *coeffs,
                       int comp, int b, int sx, int sy, int qindex)
{
    int lv = L->level[b], o = L->orient[b];
    int qi = slice_quantizer(p, qindex, lv, o);
    int32_t y1 = slice_top(p, sy, comp, lv), y2 = slice_bottom(p, sy, comp, lv);
    int32_t x1 = slice_left(p, sx, comp, lv), x2 = slice_right(p, sx, comp, lv);
    int32_t x, y;
    for (y = y1; y < y2; y++)
        for (x = x1; x < x2; x++) {
            int64_t val = read_sintb(r);
            if (r->status) return;
            coeffs[L->off[comp][b] + (int64_t)y * L->w[comp][b] + x] = vc2o_inverse_quant(val, qi);
        }
}

/* transform_data_syntax.py:315 color_diff_slice_band */
static void color_diff_slice_band(const vc2o_params *p, const layout_t *L, reader_t *r,
                                  int64_t *c1, int64_t *c2, int b, int sx, int sy, int qindex)
{
    int lv = L->level[b], o = L->orient[b];
    int qi = slice_quantizer(p, qindex, lv, o);
    int32_t y1 = slice_top(p, sy, 1, lv), y2 = slice_bottom(p, sy, 1, lv);
    int32_t x1 = slice_left(p, sx, 1, lv), x2 = slice_right(p, sx, 1, lv);
    int32_t x, y;
    for (y = y1; y < y2; y++)
        for (x = x1; x < x2; x++) {
            int64_t val = read_sintb(r);
            if (r->status) return;
            c1[L->off[1][b] + (int64_t)y * L->w[1][b] + x] = vc2o_inverse_quant(val, qi);
            val = read_sintb(r);
            if (r->status) return;
            c2[L->off[2][b] + (int64_t)y * L->w[2][b] + x] = vc2o_inverse_quant(val, qi);
        }
}

/* transform_data_syntax.py:145 ld_slice */
static void ld_slice(const vc2o_params *p, const layout_t *L, reader_t *r, int64_t *cy,
                     int64_t *c1, int64_t *c2, int sx, int sy, int32_t *qindex_out,
                     int32_t *len_out)
{
    int64_t sb = vc2o_slice_bytes(p, sx, sy);
    int64_t slice_bits_left = 8 * sb;
    int qindex, length_bits, b;
    int64_t slice_y_length;

    qindex = (int)read_nbits(r, 7);
    slice_bits_left -= 7;
    if (r->status) return;
    length_bits = vc2o_intlog2(8 * sb - 7);
    slice_y_length = read_nbits(r, length_bits);
    slice_bits_left -= length_bits;
    if (r->status) return;
    if (qindex_out) *qindex_out = qindex;
    if (len_out) *len_out = (int32_t)slice_y_length;
    if (slice_y_length > slice_bits_left) { /* :166 InvalidSliceYLength */
        r->status = VC2O_INVALID_SLICE_Y_LENGTH;
        return;
    }
    r->bits_left = slice_y_length;
    for (b = 0; b < L->nb && !r->status; b++) slice_band(p, L, r, cy, 0, b, sx, sy, qindex);
    flush_inputb(r);
    if (r->status) return;
    slice_bits_left -= slice_y_length;
    r->bits_left = slice_bits_left;
    for (b = 0; b < L->nb && !r->status; b++)
        color_diff_slice_band(p, L, r, c1, c2, b, sx, sy, qindex);
    flush_inputb(r);
}

/* transform_data_syntax.py:212 hq_slice */
static void hq_slice(const vc2o_params *p, const layout_t *L, reader_t *r, int64_t *cy,
                     int64_t *c1, int64_t *c2, int sx, int sy, int32_t *qindex_out,
                     int32_t *len_out)
{
    uint64_t start = r->pos;
    int qindex, c, b;
    int64_t *dst[3];
    dst[0] = cy; dst[1] = c1; dst[2] = c2;
    {   /* read_uint_lit(state, slice_prefix_bytes): value discarded */
        int i;
        for (i = 0; i < 8 * p->slice_prefix_bytes && !r->status; i++) read_bit(r);
    }
    qindex = (int)read_nbits(r, 8);
    if (r->status) return;
    if (qindex_out) *qindex_out = qindex;
    for (c = 0; c < 3; c++) {
        int64_t length = (int64_t)p->slice_size_scaler * read_nbits(r, 8);
        if (r->status) return;
        r->bits_left = 8 * length;
        for (b = 0; b < L->nb && !r->status; b++) slice_band(p, L, r, dst[c], c, b, sx, sy, qindex);
        flush_inputb(r);
        if (r->status) return;
    }
    if (len_out) *len_out = (int32_t)((r->pos - start) >> 3); /* total_slice_bytes :247 */
}

/* transform_data_syntax.py:63 dc_prediction */
void vc2o_dc_prediction(int64_t *band, int32_t w, int32_t h)
{
    int32_t x, y;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++) {
            int64_t prediction;
            if (x > 0 && y > 0)
                prediction = mean3(band[(int64_t)y * w + x - 1], band[(int64_t)(y - 1) * w + x - 1],
                                   band[(int64_t)(y - 1) * w + x]);
            else if (x > 0 && y == 0)
                prediction = band[x - 1];
            else if (x == 0 && y > 0)
                prediction = band[(int64_t)(y - 1) * w];
            else
                prediction = 0;
            band[(int64_t)y * w + x] += prediction;
        }
}

/* encoder/pictures.py:202 apply_dc_prediction */
void vc2o_apply_dc_prediction(int64_t *band, int32_t w, int32_t h)
{
    int32_t x, y;
    for (y = h - 1; y >= 0; y--)
        for (x = w - 1; x >= 0; x--) {
            int64_t prediction;
            if (x > 0 && y > 0)
                prediction = mean3(band[(int64_t)y * w + x - 1], band[(int64_t)(y - 1) * w + x - 1],
                                   band[(int64_t)(y - 1) * w + x]);
            else if (x > 0 && y == 0)
                prediction = band[x - 1];
            else if (x == 0 && y > 0)
                prediction = band[(int64_t)(y - 1) * w];
            else
                prediction = 0;
            band[(int64_t)y * w + x] -= prediction;
        }
}

/* transform_data_syntax.py:116 transform_data */
int vc2o_transform_data(const vc2o_params *p, const uint8_t *data, size_t nbytes,
                        int64_t *coeffs_y, int64_t *coeffs_c1, int64_t *coeffs_c2,
                        int32_t *qindex_out, int32_t *slice_len_out, size_t *consumed,
                        int32_t *bad_slice)
{
    layout_t L;
    reader_t r;
    int sx, sy;
    make_layout(p, &L);
    memset(coeffs_y, 0, sizeof(int64_t) * (size_t)L.total[0]);
    memset(coeffs_c1, 0, sizeof(int64_t) * (size_t)L.total[1]);
    memset(coeffs_c2, 0, sizeof(int64_t) * (size_t)L.total[2]);
    r.data = data;
    r.nbits_total = 8 * (uint64_t)nbytes;
    r.pos = 0;
    r.bits_left = 0;
    r.status = VC2O_OK;
    if (bad_slice) *bad_slice = -1;
    for (sy = 0; sy < p->slices_y; sy++)
        for (sx = 0; sx < p->slices_x; sx++) {
            int n = sy * p->slices_x + sx;
            if (p->is_ld)
                ld_slice(p, &L, &r, coeffs_y, coeffs_c1, coeffs_c2, sx, sy,
                         qindex_out ? qindex_out + n : NULL, slice_len_out ? slice_len_out + n : NULL);
            else
                hq_slice(p, &L, &r, coeffs_y, coeffs_c1, coeffs_c2, sx, sy,
                         qindex_out ? qindex_out + n : NULL, slice_len_out ? slice_len_out + n : NULL);
            if (r.status) {
                if (bad_slice) *bad_slice = n;
                if (consumed) *consumed = (size_t)((r.pos + 7) >> 3);
                return r.status;
            }
        }
    if (p->is_ld) { /* using_dc_prediction, pseudocode/parse_code_functions.py:70 */
        vc2o_dc_prediction(coeffs_y, L.w[0][0], L.h[0][0]);
        vc2o_dc_prediction(coeffs_c1, L.w[1][0], L.h[1][0]);
        vc2o_dc_prediction(coeffs_c2, L.w[2][0], L.h[2][0]);
    }
    if (consumed) *consumed = (size_t)((r.pos + 7) >> 3);
    return VC2O_OK;
}

/* ------------------------------------------------------------------------ */
/* lifting (pseudocode/picture_decoding.py:188-269)                          */
/* ------------------------------------------------------------------------ */

typedef struct { int lift_type, S, L, D; int taps[8]; } stage_t;
typedef struct { int nstages; stage_t stages[4]; int filter_bit_shift; } filter_t;

/* vc2_data_tables.LIFTING_FILTERS (un-vendored dependency; SMPTE ST 2042-1
 * Table 15.1-15.7).  lift_type: 1 even+=odd, 2 even-=odd, 3 odd+=even,
 * 4 odd-=even (pseudocode/picture_decoding.py:264-269). */
static const filter_t FILTERS[7] = {
    /* 0 Deslauriers-Dubuc (9,7) */
    {2, {{2, 2, 2, 0, {1, 1}}, {3, 4, 4, -1, {-1, 9, 9, -1}}}, 1},
    /* 1 LeGall (5,3) */
    {2, {{2, 2, 2, 0, {1, 1}}, {3, 1, 2, 0, {1, 1}}}, 1},
    /* 2 Deslauriers-Dubuc (13,7) */
    {2, {{2, 5, 4, -1, {-1, 9, 9, -1}}, {3, 4, 4, -1, {-1, 9, 9, -1}}}, 1},
    /* 3 Haar, no shift */
    {2, {{2, 1, 1, 1, {1}}, {3, 0, 1, 0, {1}}}, 0},
    /* 4 Haar, with shift */
    {2, {{2, 1, 1, 1, {1}}, {3, 0, 1, 0, {1}}}, 1},
    /* 5 Fidelity */
    {2, {{3, 8, 8, -3, {-2, 10, -25, 81, 81, -25, 10, -2}},
         {2, 8, 8, -3, {-8, 21, -46, 161, 161, -46, 21, -8}}}, 0},
    /* 6 Daubechies (9,7) */
    {4, {{2, 12, 2, 0, {1817, 1817}}, {4, 12, 2, 0, {3616, 3616}},
         {1, 12, 2, 0, {217, 217}}, {3, 12, 2, 0, {6497, 6497}}}, 1},
};

int vc2o_filter_bit_shift(int filter) { return FILTERS[filter].filter_bit_shift; }

/* picture_decoding.py:205-261 lift1..lift4 on a strided 1-D view */
static void lift(int64_t *A, int64_t stride, int64_t len, int type, int L, int D, const int *taps,
                 int S)
{
    int64_t n;
    int i;
    for (n = 0; n < len / 2; n++) {
        int64_t sum = 0;
        for (i = D; i < L + D; i++) {
            int64_t pos;
            if (type == 1 || type == 2) {
                pos = 2 * (n + i) - 1;
                if (pos > len - 1) pos = len - 1;
                if (pos < 1) pos = 1;
            } else {
                pos = 2 * (n + i);
                if (pos > len - 2) pos = len - 2;
                if (pos < 0) pos = 0;
            }
            sum += (int64_t)taps[i - D] * A[pos * stride];
        }
        if (S > 0) sum += (int64_t)1 << (S - 1);
        sum >>= S; /* arithmetic shift == Python floor shift */
        switch (type) {
        case 1: A[(2 * n) * stride] += sum; break;
        case 2: A[(2 * n) * stride] -= sum; break;
        case 3: A[(2 * n + 1) * stride] += sum; break;
        default: A[(2 * n + 1) * stride] -= sum; break;
        }
    }
}

/* picture_decoding.py:188 oned_synthesis */
void vc2o_oned_synthesis(int64_t *a, int64_t stride, int64_t len, int filter)
{
    const filter_t *f = &FILTERS[filter];
    int s;
    for (s = 0; s < f->nstages; s++) {
        const stage_t *st = &f->stages[s];
        lift(a, stride, len, st->lift_type, st->L, st->D, st->taps, st->S);
    }
}

/* picture_encoding.py:203 oned_analysis: reversed stages, add<->subtract (:214) */
void vc2o_oned_analysis(int64_t *a, int64_t stride, int64_t len, int filter)
{
    static const int swap[5] = {0, 2, 1, 4, 3};
    const filter_t *f = &FILTERS[filter];
    int s;
    for (s = f->nstages - 1; s >= 0; s--) {
        const stage_t *st = &f->stages[s];
        lift(a, stride, len, swap[st->lift_type], st->L, st->D, st->taps, st->S);
    }
}

/* ------------------------------------------------------------------------ */
/* synthesis (pseudocode/picture_decoding.py:45-184, 282-349)                */
/* ------------------------------------------------------------------------ */

static void bit_shift_down(int64_t *a, int64_t n, int shift)
{
    int64_t i;
    if (shift <= 0) return;
    for (i = 0; i < n; i++) a[i] = (a[i] + ((int64_t)1 << (shift - 1))) >> shift;
}

/* picture_decoding.py:88 idwt (+ :134 h_synthesis, :159 vh_synthesis) */
int vc2o_idwt(const vc2o_params *p, int comp, const int64_t *coeffs, int64_t *out)
{
    layout_t L;
    int b, n, top = p->dwt_depth_ho + p->dwt_depth;
    int32_t w, h;
    int64_t *cur, *next;
    int64_t maxn;
    int shift = FILTERS[p->wavelet_index_ho].filter_bit_shift; /* :198 filter_bit_shift */
    if (p->wavelet_index < 0 || p->wavelet_index > 6 || p->wavelet_index_ho < 0 ||
        p->wavelet_index_ho > 6 || top + 1 > VC2O_MAX_LEVELS)
        return VC2O_BAD_PARAMS;
    make_layout(p, &L);
    maxn = L.total[comp];
    cur = (int64_t *)malloc(sizeof(int64_t) * (size_t)maxn);
    next = (int64_t *)malloc(sizeof(int64_t) * (size_t)maxn);
    w = L.w[comp][0];
    h = L.h[comp][0];
    memcpy(cur, coeffs + L.off[comp][0], sizeof(int64_t) * (size_t)w * h);
    b = 1;
    for (n = 1; n <= p->dwt_depth_ho; n++, b++) { /* :134 h_synthesis */
        const int64_t *H = coeffs + L.off[comp][b];
        int32_t x, y, w2 = 2 * w;
        int64_t *tmp;
        for (y = 0; y < h; y++)
            for (x = 0; x < w; x++) {
                next[(int64_t)y * w2 + 2 * x] = cur[(int64_t)y * w + x];
                next[(int64_t)y * w2 + 2 * x + 1] = H[(int64_t)y * w + x];
            }
        for (y = 0; y < h; y++) vc2o_oned_synthesis(next + (int64_t)y * w2, 1, w2, p->wavelet_index_ho);
        bit_shift_down(next, (int64_t)w2 * h, shift);
        tmp = cur; cur = next; next = tmp;
        w = w2;
    }
    for (n = p->dwt_depth_ho + 1; n <= top; n++, b += 3) { /* :159 vh_synthesis */
        const int64_t *HL = coeffs + L.off[comp][b];
        const int64_t *LH = coeffs + L.off[comp][b + 1];
        const int64_t *HH = coeffs + L.off[comp][b + 2];
        int32_t x, y, w2 = 2 * w, h2 = 2 * h;
        int64_t *tmp;
        for (y = 0; y < h; y++)
            for (x = 0; x < w; x++) {
                next[(int64_t)(2 * y) * w2 + 2 * x] = cur[(int64_t)y * w + x];
                next[(int64_t)(2 * y) * w2 + 2 * x + 1] = HL[(int64_t)y * w + x];
                next[(int64_t)(2 * y + 1) * w2 + 2 * x] = LH[(int64_t)y * w + x];
                next[(int64_t)(2 * y + 1) * w2 + 2 * x + 1] = HH[(int64_t)y * w + x];
            }
        for (x = 0; x < w2; x++) vc2o_oned_synthesis(next + x, w2, h2, p->wavelet_index);
        for (y = 0; y < h2; y++) vc2o_oned_synthesis(next + (int64_t)y * w2, 1, w2, p->wavelet_index_ho);
        bit_shift_down(next, (int64_t)w2 * h2, shift);
        tmp = cur; cur = next; next = tmp;
        w = w2;
        h = h2;
    }
    memcpy(out, cur, sizeof(int64_t) * (size_t)w * h);
    free(cur);
    free(next);
    return VC2O_OK;
}

/* picture_decoding.py:45 picture_decode for one component: idwt, :282 pad
 * removal, :327 clip, :301 offset. */
int vc2o_decode_component(const vc2o_params *p, int comp, const int64_t *coeffs, uint16_t *out)
{
    int top = p->dwt_depth_ho + p->dwt_depth;
    int32_t pw = (top == 0) ? vc2o_subband_width(p, 0, comp) : 2 * vc2o_subband_width(p, top, comp);
    int32_t ph = (p->dwt_depth == 0) ? vc2o_subband_height(p, 0, comp)
                                     : 2 * vc2o_subband_height(p, top, comp);
    int32_t w = comp_width(p, comp), h = comp_height(p, comp), d = comp_depth(p, comp);
    int64_t lo = -((int64_t)1 << (d - 1)), hi = ((int64_t)1 << (d - 1)) - 1;
    int64_t *pic = (int64_t *)malloc(sizeof(int64_t) * (size_t)pw * ph);
    int32_t x, y;
    int st = vc2o_idwt(p, comp, coeffs, pic);
    if (st) { free(pic); return st; }
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++) {
            int64_t v = pic[(int64_t)y * pw + x];
            if (v < lo) v = lo;
            if (v > hi) v = hi;
            out[(int64_t)y * w + x] = (uint16_t)(v - lo);
        }
    free(pic);
    return VC2O_OK;
}

/* ------------------------------------------------------------------------ */
/* analysis (pseudocode/picture_encoding.py:45-286)                          */
/* ------------------------------------------------------------------------ */

/* picture_encoding.py:93 dwt (+ :141 h_analysis, :169 vh_analysis).
 * padded_picture is consumed (overwritten). */
int vc2o_dwt(const vc2o_params *p, int comp, int64_t *pic, int64_t *coeffs)
{
    layout_t L;
    int n, top = p->dwt_depth_ho + p->dwt_depth;
    int shift = FILTERS[p->wavelet_index_ho].filter_bit_shift;
    int32_t w, h;
    int64_t *cur = pic, *tmp;
    int b;
    make_layout(p, &L);
    w = (top == 0) ? L.w[comp][0] : 2 * L.w[comp][L.nb - 1];
    h = (p->dwt_depth == 0) ? L.h[comp][0] : 2 * L.h[comp][L.nb - 1];
    tmp = (int64_t *)malloc(sizeof(int64_t) * (size_t)w * h);
    b = L.nb;
    for (n = top; n >= p->dwt_depth_ho + 1; n--) { /* :169 vh_analysis */
        int32_t x, y, w2 = w / 2, h2 = h / 2;
        int64_t *HL, *LH, *HH;
        int64_t i;
        b -= 3;
        HL = coeffs + L.off[comp][b];
        LH = coeffs + L.off[comp][b + 1];
        HH = coeffs + L.off[comp][b + 2];
        if (shift > 0)
            for (i = 0; i < (int64_t)w * h; i++) cur[i] = cur[i] * ((int64_t)1 << shift);
        for (y = 0; y < h; y++) vc2o_oned_analysis(cur + (int64_t)y * w, 1, w, p->wavelet_index_ho);
        for (x = 0; x < w; x++) vc2o_oned_analysis(cur + x, w, h, p->wavelet_index);
        for (y = 0; y < h2; y++)
            for (x = 0; x < w2; x++) {
                tmp[(int64_t)y * w2 + x] = cur[(int64_t)(2 * y) * w + 2 * x];
                HL[(int64_t)y * w2 + x] = cur[(int64_t)(2 * y) * w + 2 * x + 1];
                LH[(int64_t)y * w2 + x] = cur[(int64_t)(2 * y + 1) * w + 2 * x];
                HH[(int64_t)y * w2 + x] = cur[(int64_t)(2 * y + 1) * w + 2 * x + 1];
            }
        memcpy(cur, tmp, sizeof(int64_t) * (size_t)w2 * h2);
        w = w2;
        h = h2;
    }
    for (n = p->dwt_depth_ho; n >= 1; n--) { /* :141 h_analysis */
        int32_t x, y, w2 = w / 2;
        int64_t *H;
        int64_t i;
        b -= 1;
        H = coeffs + L.off[comp][b];
        if (shift > 0)
            for (i = 0; i < (int64_t)w * h; i++) cur[i] = cur[i] * ((int64_t)1 << shift);
        for (y = 0; y < h; y++) vc2o_oned_analysis(cur + (int64_t)y * w, 1, w, p->wavelet_index_ho);
        for (y = 0; y < h; y++)
            for (x = 0; x < w2; x++) {
                tmp[(int64_t)y * w2 + x] = cur[(int64_t)y * w + 2 * x];
                H[(int64_t)y * w2 + x] = cur[(int64_t)y * w + 2 * x + 1];
            }
        memcpy(cur, tmp, sizeof(int64_t) * (size_t)w2 * h);
        w = w2;
    }
    memcpy(coeffs + L.off[comp][0], cur, sizeof(int64_t) * (size_t)w * h);
    free(tmp);
    return VC2O_OK;
}

/* picture_encoding.py:45 picture_encode for one component: :264 remove
 * offset, :235 dwt_pad_addition (edge replicate), :93 dwt. */
int vc2o_encode_component(const vc2o_params *p, int comp, const uint16_t *pic, int64_t *coeffs)
{
    int top = p->dwt_depth_ho + p->dwt_depth;
    int32_t pw = (top == 0) ? vc2o_subband_width(p, 0, comp) : 2 * vc2o_subband_width(p, top, comp);
    int32_t ph = (p->dwt_depth == 0) ? vc2o_subband_height(p, 0, comp)
                                     : 2 * vc2o_subband_height(p, top, comp);
    int32_t w = comp_width(p, comp), h = comp_height(p, comp), d = comp_depth(p, comp);
    int64_t off = (int64_t)1 << (d - 1);
    int64_t *buf = (int64_t *)malloc(sizeof(int64_t) * (size_t)pw * ph);
    int32_t x, y;
    int st;
    for (y = 0; y < ph; y++) {
        int32_t sy = y < h ? y : h - 1;
        for (x = 0; x < pw; x++) {
            int32_t sx = x < w ? x : w - 1;
            buf[(int64_t)y * pw + x] = (int64_t)pic[(int64_t)sy * w + sx] - off;
        }
    }
    st = vc2o_dwt(p, comp, buf, coeffs);
    free(buf);
    return st;
}

/* ------------------------------------------------------------------------ */
/* encoder slice maths (encoder/pictures.py:217-347, 418-792)                */
/* ------------------------------------------------------------------------ */

/* encoder/pictures.py:217 calculate_coeffs_bits */
int64_t vc2o_calculate_coeffs_bits(const int64_t *coeffs, size_t n)
{
    int64_t bits = 0;
    size_t i;
    while (n > 0 && coeffs[n - 1] == 0) n--;
    for (i = 0; i < n; i++) bits += vc2o_signed_exp_golomb_length(coeffs[i]);
    return bits;
}

/* encoder/pictures.py:418-451 gather in bitstream order */
int64_t vc2o_gather_slice(const vc2o_params *p, int comp, const int64_t *coeffs, int sx, int sy,
                          int64_t *out, int32_t *qm_out)
{
    layout_t L;
    int b;
    int64_t n = 0;
    make_layout(p, &L);
    for (b = 0; b < L.nb; b++) {
        int lv = L.level[b];
        int32_t y1 = slice_top(p, sy, comp, lv), y2 = slice_bottom(p, sy, comp, lv);
        int32_t x1 = slice_left(p, sx, comp, lv), x2 = slice_right(p, sx, comp, lv);
        int32_t x, y;
        for (y = y1; y < y2; y++)
            for (x = x1; x < x2; x++) {
                if (out) out[n] = coeffs[L.off[comp][b] + (int64_t)y * L.w[comp][b] + x];
                if (qm_out) qm_out[n] = p->quant_matrix[lv][L.orient[b]];
                n++;
            }
    }
    return n;
}

/* encoder/pictures.py:251 quantize_coeffs */
static void quantize_coeffs(int qindex, const int64_t *c, const int32_t *qm, size_t n, int64_t *out)
{
    size_t i;
    for (i = 0; i < n; i++) {
        int q = qindex - qm[i];
        out[i] = vc2o_forward_quant(c[i], q > 0 ? q : 0);
    }
}

/* encoder/pictures.py:295 quantize_to_fit */
int vc2o_quantize_to_fit(int64_t target_size, int nsets, const int64_t *const *coeffs,
                         const int32_t *const *qm, const size_t *n, int64_t align_bits,
                         int minimum_qindex, int64_t *const *out)
{
    int qindex;
    for (qindex = minimum_qindex;; qindex++) {
        int64_t total = 0;
        int s;
        for (s = 0; s < nsets; s++) {
            int64_t bits;
            quantize_coeffs(qindex, coeffs[s], qm[s], n[s], out[s]);
            bits = vc2o_calculate_coeffs_bits(out[s], n[s]);
            total += ((bits + align_bits - 1) / align_bits) * align_bits;
        }
        if (total <= target_size) return qindex;
    }
}

/* ---- bit writer (bitstream/io.py:457-583) ---- */
typedef struct {
    uint8_t *out;
    size_t cap;
    uint64_t pos;        /* bits written */
    int64_t bits_remaining; /* bounded block; <0 when not in a block (flag) */
    int bounded;
    int overflow;
} writer_t;

/* bitstream/io.py:457 write_bit: '1' past the end of a bounded block is
 * silently dropped (a '0' cannot occur: lengths are computed to fit). */
static void write_bit(writer_t *w, int bit)
{
    if (w->bounded) {
        w->bits_remaining--;
        if (w->bits_remaining <= -1) {
            if (!bit) w->overflow = 1;
            return;
        }
    }
    if ((w->pos >> 3) >= w->cap) { w->overflow = 1; return; }
    if (bit) w->out[w->pos >> 3] |= (uint8_t)(0x80 >> (w->pos & 7));
    else w->out[w->pos >> 3] &= (uint8_t)~(0x80 >> (w->pos & 7));
    w->pos++;
}
static void write_nbits(writer_t *w, int bits, uint64_t value)
{
    int i;
    for (i = bits - 1; i >= 0; i--) write_bit(w, (int)((value >> i) & 1));
}
/* bitstream/io.py:557 write_uint / :575 write_sint */
static void write_sint(writer_t *w, int64_t value)
{
    uint64_t m = (uint64_t)(value < 0 ? -value : value) + 1;
    int bl = 0, i;
    uint64_t t = m;
    while (t) { bl++; t >>= 1; }
    for (i = bl - 2; i >= 0; i--) {
        write_bit(w, 0);
        write_bit(w, (int)((m >> i) & 1));
    }
    write_bit(w, 1);
    if (value != 0) write_bit(w, value < 0);
}
static void block_begin(writer_t *w, int64_t length) { w->bounded = 1; w->bits_remaining = length; }
/* serdes.py:688 bounded_block_end: unused bits written as zero padding */
static void block_end(writer_t *w)
{
    int64_t unused = w->bits_remaining > 0 ? w->bits_remaining : 0;
    w->bounded = 0;
    while (unused-- > 0) write_bit(w, 0);
}

static int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

typedef struct {
    int64_t *c[3];
    int32_t *qm[3];
    int64_t *q[3];
    size_t n[3];
} slicebuf_t;

static void slicebuf_alloc(const vc2o_params *p, slicebuf_t *sb)
{
    /* worst-case slice size: whole component */
    layout_t L;
    int c;
    make_layout(p, &L);
    for (c = 0; c < 3; c++) {
        size_t cap = (size_t)L.total[c];
        if (c == 1) cap *= 2; /* LD interleaves C1/C2 into set 1 */
        sb->c[c] = (int64_t *)malloc(sizeof(int64_t) * (cap + 1));
        sb->q[c] = (int64_t *)malloc(sizeof(int64_t) * (cap + 1));
        sb->qm[c] = (int32_t *)malloc(sizeof(int32_t) * (cap + 1));
    }
}
static void slicebuf_free(slicebuf_t *sb)
{
    int c;
    for (c = 0; c < 3; c++) { free(sb->c[c]); free(sb->q[c]); free(sb->qm[c]); }
}

/* bitstream/vc2.py:798 hq_slice serialisation of encoder/pictures.py:456 make_hq_slice */
static void write_hq_slice(const vc2o_params *p, writer_t *w, slicebuf_t *sb, int qindex,
                           const int64_t len[3])
{
    int c;
    size_t i;
    for (i = 0; i < (size_t)p->slice_prefix_bytes; i++) write_nbits(w, 8, 0);
    write_nbits(w, 8, (uint64_t)qindex);
    for (c = 0; c < 3; c++) {
        write_nbits(w, 8, (uint64_t)len[c]);
        block_begin(w, 8 * (int64_t)p->slice_size_scaler * len[c]);
        for (i = 0; i < sb->n[c]; i++) write_sint(w, sb->q[c][i]);
        block_end(w);
    }
}

/* encoder/pictures.py:617 make_transform_data_hq_lossy */
int64_t vc2o_encode_hq_lossy(const vc2o_params *p, const int64_t *cy, const int64_t *cc1,
                             const int64_t *cc2, int64_t picture_bytes, int minimum_qindex,
                             uint8_t *out, size_t cap, int32_t *qindex_out)
{
    const int64_t *src[3];
    slicebuf_t sb;
    writer_t w;
    int64_t num_slices = (int64_t)p->slices_x * p->slices_y;
    int64_t scaler = p->slice_size_scaler;
    int64_t total_coeff_bytes = picture_bytes - num_slices * 4;
    int64_t den = num_slices * scaler;
    int sx, sy, c;
    src[0] = cy; src[1] = cc1; src[2] = cc2;
    if (total_coeff_bytes < 0) return -1; /* InsufficientHQPictureBytesError :663 */
    memset(&w, 0, sizeof(w));
    w.out = out;
    w.cap = cap;
    slicebuf_alloc(p, &sb);
    for (sy = 0; sy < p->slices_y; sy++)
        for (sx = 0; sx < p->slices_x; sx++) {
            int64_t n = (int64_t)sy * p->slices_x + sx;
            int64_t total_length = ((n + 1) * total_coeff_bytes) / den - (n * total_coeff_bytes) / den;
            int64_t target = 8 * scaler * total_length;
            int64_t len[3];
            int qindex;
            for (c = 0; c < 3; c++)
                sb.n[c] = (size_t)vc2o_gather_slice(p, c, src[c], sx, sy, sb.c[c], sb.qm[c]);
            qindex = vc2o_quantize_to_fit(target, 3, (const int64_t *const *)sb.c,
                                          (const int32_t *const *)sb.qm, sb.n, 8 * scaler,
                                          minimum_qindex, sb.q);
            /* :456 make_hq_slice */
            len[0] = ceil_div(vc2o_calculate_coeffs_bits(sb.q[0], sb.n[0]), 8 * scaler);
            len[1] = ceil_div(vc2o_calculate_coeffs_bits(sb.q[1], sb.n[1]), 8 * scaler);
            len[2] = total_length - len[0] - len[1];
            if (len[0] > 255 || len[1] > 255 || len[2] > 255 || qindex > 255) {
                slicebuf_free(&sb); /* bitstream/io.py:485 OutOfRangeError: scaler too small */
                return -2;
            }
            if (qindex_out) qindex_out[n] = qindex;
            write_hq_slice(p, &w, &sb, qindex, len);
        }
    slicebuf_free(&sb);
    if (w.overflow) return -1;
    return (int64_t)(w.pos >> 3);
}

/* encoder/pictures.py:528 make_transform_data_hq_lossless (fixed qindex) */
int64_t vc2o_encode_hq_fixed(const vc2o_params *p, const int64_t *cy, const int64_t *cc1,
                             const int64_t *cc2, int qindex, uint8_t *out, size_t cap)
{
    const int64_t *src[3];
    slicebuf_t sb;
    writer_t w;
    int64_t scaler = p->slice_size_scaler;
    int sx, sy, c;
    src[0] = cy; src[1] = cc1; src[2] = cc2;
    memset(&w, 0, sizeof(w));
    w.out = out;
    w.cap = cap;
    slicebuf_alloc(p, &sb);
    for (sy = 0; sy < p->slices_y; sy++)
        for (sx = 0; sx < p->slices_x; sx++) {
            int64_t len[3];
            for (c = 0; c < 3; c++) {
                sb.n[c] = (size_t)vc2o_gather_slice(p, c, src[c], sx, sy, sb.c[c], sb.qm[c]);
                quantize_coeffs(qindex, sb.c[c], sb.qm[c], sb.n[c], sb.q[c]);
                len[c] = ceil_div(vc2o_calculate_coeffs_bits(sb.q[c], sb.n[c]), 8 * scaler);
                if (len[c] > 255) { slicebuf_free(&sb); return -2; }
            }
            write_hq_slice(p, &w, &sb, qindex, len);
        }
    slicebuf_free(&sb);
    if (w.overflow) return -1;
    return (int64_t)(w.pos >> 3);
}

/* encoder/pictures.py:724 make_transform_data_ld_lossy + bitstream/vc2.py:727 ld_slice */
int64_t vc2o_encode_ld_lossy(const vc2o_params *p, const int64_t *cy, const int64_t *cc1,
                             const int64_t *cc2, int minimum_qindex, uint8_t *out, size_t cap,
                             int32_t *qindex_out)
{
    slicebuf_t sb;
    writer_t w;
    int sx, sy;
    memset(&w, 0, sizeof(w));
    w.out = out;
    w.cap = cap;
    slicebuf_alloc(p, &sb);
    for (sy = 0; sy < p->slices_y; sy++)
        for (sx = 0; sx < p->slices_x; sx++) {
            int64_t n = (int64_t)sy * p->slices_x + sx;
            int64_t sbytes = vc2o_slice_bytes(p, sx, sy);
            int64_t target = 8 * sbytes - 7;
            int64_t slice_bits_left = 8 * sbytes;
            int length_bits = vc2o_intlog2(8 * sbytes - 7);
            int64_t y_length;
            size_t i, nc;
            int qindex;
            const int64_t *sets_c[2];
            const int32_t *sets_qm[2];
            int64_t *sets_q[2];
            size_t sets_n[2];
            target -= vc2o_intlog2(target);
            if (target < 0) { slicebuf_free(&sb); return -1; } /* :763 */
            sb.n[0] = (size_t)vc2o_gather_slice(p, 0, cy, sx, sy, sb.c[0], sb.qm[0]);
            /* :712 interleave C1/C2 */
            nc = (size_t)vc2o_gather_slice(p, 1, cc1, sx, sy, sb.c[2], sb.qm[2]);
            for (i = 0; i < nc; i++) { sb.c[1][2 * i] = sb.c[2][i]; sb.qm[1][2 * i] = sb.qm[2][i]; }
            vc2o_gather_slice(p, 2, cc2, sx, sy, sb.c[2], sb.qm[2]);
            for (i = 0; i < nc; i++) { sb.c[1][2 * i + 1] = sb.c[2][i]; sb.qm[1][2 * i + 1] = sb.qm[2][i]; }
            sb.n[1] = 2 * nc;
            sets_c[0] = sb.c[0]; sets_c[1] = sb.c[1];
            sets_qm[0] = sb.qm[0]; sets_qm[1] = sb.qm[1];
            sets_q[0] = sb.q[0]; sets_q[1] = sb.q[1];
            sets_n[0] = sb.n[0]; sets_n[1] = sb.n[1];
            qindex = vc2o_quantize_to_fit(target, 2, sets_c, sets_qm, sets_n, 1, minimum_qindex, sets_q);
            if (qindex_out) qindex_out[n] = qindex;
            /* :507 make_ld_slice + bitstream/vc2.py:727 */
            y_length = vc2o_calculate_coeffs_bits(sb.q[0], sb.n[0]);
            write_nbits(&w, 7, (uint64_t)qindex);
            slice_bits_left -= 7;
            write_nbits(&w, length_bits, (uint64_t)y_length);
            slice_bits_left -= length_bits;
            block_begin(&w, y_length);
            for (i = 0; i < sb.n[0]; i++) write_sint(&w, sb.q[0][i]);
            block_end(&w);
            slice_bits_left -= y_length;
            block_begin(&w, slice_bits_left);
            for (i = 0; i < sb.n[1]; i++) write_sint(&w, sb.q[1][i]);
            block_end(&w);
        }
    slicebuf_free(&sb);
    if (w.overflow) return -1;
    return (int64_t)((w.pos + 7) >> 3);
}
