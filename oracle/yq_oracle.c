/*
 * oracle/yq_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or executed
 * from the product path (yolo_quantization_b200/).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load liboracle.so, and there only as the checker.
 *
 * A plain-C restatement (CPU, CHW layout like the reference) of the reference's QUANTIZATION=1
 * inference hot path.  Each function cites the reference file:line it follows.
 *
 * PARITY PINNING: the reference ships no tests / golden vectors, so this oracle is
 * pinned against outputs of the reference itself compiled here (oracle/_ref, `make -C oracle ref`):
 * tests/test_oracle_vs_reference.py runs both on the same seeded inputs, and the SHA-256 of the
 * reference's per-layer dumps are committed under tests/golden/ (generator: oracle/gen_golden.py).
 *
 * Exactness contract: yq_oracle_conv_acc is the EXACT-integer accumulator.  The
 * reference carries its "int32" accumulator through float32 (gemm.c:279-296) and is exact only
 * while every running partial sum stays <= 2^24; yq_oracle_conv_acc_reffloat restates that
 * float-carried behaviour so the deviation can be measured rather than assumed.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define YQ_LINEAR 3   /* ACTIVATION enum values, include/darknet.h:87-89 */
#define YQ_RELU 1
#define YQ_RELU6 8
#define YQ_LEAKY 9

/* im2col_get_pixel_uint8, src/im2col.c:5-14: out-of-bounds taps return the INPUT ZERO-POINT
   (yq_oracle_conv_acc below applies the same rule inline). */
static inline int px(const uint8_t *im, int h, int w, int row, int col, int ch, int pad, uint8_t zp_in)
{
    row -= pad;
    col -= pad;
    if (row < 0 || col < 0 || row >= h || col >= w) return zp_in;
    return im[col + w * (row + h * ch)];
}

/*
 * Exact-integer accumulator:
 *   acc[oc,oy,ox] = sum_{ci,ky,kx} (w[oc,ci,ky,kx] - zp_w[oc]) * A(ci, oy*stride+ky-pad, ox*stride+kx-pad)
 * Follows forward_convolutional_layer_quant_inputi_outputi, src/convolutional_layer.c:699-722
 * (im2col with pad value zp_in, then C = W*B, then C -= ZPW*B), im2col_cpu_uint8 src/im2col.c:26-50.
 * groups == 1 (the only case the shipped cfgs use).
 */
void yq_oracle_conv_acc(const uint8_t *in, int c, int h, int w, const uint8_t *weights, const uint8_t *zp_w,
                        int n, int size, int stride, int pad, int zp_in, int32_t *acc)
{
    int oh = (h + 2 * pad - size) / stride + 1;
    int ow = (w + 2 * pad - size) / stride + 1;
    /* loop order (oc | ci,ky,kx | oy,ox) keeps the inner loop unit-stride; int32 sums are exact
       because |acc| <= K*255*255 < 2^31 for K <= 33025. */
#pragma omp parallel for schedule(dynamic, 1)
    for (int oc = 0; oc < n; ++oc) {
        const uint8_t *wk = weights + (size_t)oc * c * size * size;
        int zw = zp_w[oc];
        int32_t *o = acc + (size_t)oc * oh * ow;
        memset(o, 0, sizeof(int32_t) * (size_t)oh * ow);
        for (int ci = 0; ci < c; ++ci)
            for (int ky = 0; ky < size; ++ky)
                for (int kx = 0; kx < size; ++kx) {
                    int wv = (int)wk[(ci * size + ky) * size + kx] - zw;
                    if (wv == 0) continue;
                    for (int oy = 0; oy < oh; ++oy) {
                        int32_t *orow = o + (size_t)oy * ow;
                        int iy = oy * stride + ky - pad;
                        if (iy < 0 || iy >= h) {
                            int t = wv * zp_in;
                            for (int ox = 0; ox < ow; ++ox) orow[ox] += t;
                            continue;
                        }
                        const uint8_t *irow = in + ((size_t)ci * h + iy) * w;
                        for (int ox = 0; ox < ow; ++ox) {
                            int ix = ox * stride + kx - pad;
                            int a = (ix < 0 || ix >= w) ? zp_in : irow[ix];
                            orow[ox] += wv * a;
                        }
                    }
                }
    }
}

/*
 * The reference's float-carried accumulator, restated: gemm_nn_uint8_int32_te (src/gemm.c:279-299)
 * executes  C[i][j] += ALPHA*A[i][k]*B[k][j]  with int32 C and float ALPHA, i.e. per step
 * C = (int32)((float)C + (float)product); first pass ALPHA=+1 over the weights, second pass
 * ALPHA=-1 over the zero-point matrix (convolutional_layer.c:718,721).  k runs in im2col row order
 * (ci, ky, kx).  Identical to yq_oracle_conv_acc while |partial| <= 2^24.
 */
void yq_oracle_conv_acc_reffloat(const uint8_t *in, int c, int h, int w, const uint8_t *weights,
                                 const uint8_t *zp_w, int n, int size, int stride, int pad, int zp_in,
                                 int32_t *acc)
{
    int oh = (h + 2 * pad - size) / stride + 1;
    int ow = (w + 2 * pad - size) / stride + 1;
#pragma omp parallel for schedule(static)
    for (int oc = 0; oc < n; ++oc) {
        const uint8_t *wk = weights + (size_t)oc * c * size * size;
        float zw = (float)zp_w[oc];
        for (int oy = 0; oy < oh; ++oy) {
            for (int ox = 0; ox < ow; ++ox) {
                int32_t C = 0;
                for (int pass = 0; pass < 2; ++pass)
                    for (int ci = 0; ci < c; ++ci)
                        for (int ky = 0; ky < size; ++ky)
                            for (int kx = 0; kx < size; ++kx) {
                                float a = (float)px(in, h, w, oy * stride + ky, ox * stride + kx, ci, pad,
                                                    (uint8_t)zp_in);
                                float wv = pass == 0 ? (float)wk[(ci * size + ky) * size + kx] : -zw;
                                volatile float t = (float)C + wv * a;   /* one fp32 rounding per step */
                                C = (int32_t)t;
                            }
                acc[((size_t)oc * oh + oy) * ow + ox] = C;
            }
        }
    }
}

/*
 * Requantize + activation + zero-point + uint8 store, faithful to
 * src/convolutional_layer.c:726-751:
 *   int64 t = (acc + bias_i32[oc]) * M_value[oc];          int32 add, ONE double multiply, trunc toward 0
 *   int32 q = t * M0_right_shift_value[oc];                int64->double, double multiply, trunc toward 0
 *   LEAKY : q<0 ? round(q*0.1)+zo : q+zo   (double)        RELU6: q<=0 ? zo : q+zo     LINEAR/RELU: q+zo
 *   store into uint8_t  => two's-complement WRAP mod 256; the clamp() on :749 is a no-op on a uint8_t.
 */
void yq_oracle_requant(const int32_t *acc, int n, int spatial, const int32_t *bias_i32, const double *M_value,
                       const double *rshift_value, int activation, int zp_out, uint8_t *out)
{
#pragma omp parallel for schedule(static)
    for (int oc = 0; oc < n; ++oc) {
        for (int j = 0; j < spatial; ++j) {
            size_t idx = (size_t)oc * spatial + j;
            int32_t x = (int32_t)((uint32_t)acc[idx] + (uint32_t)bias_i32[oc]);
            int64_t t = (int64_t)((double)x * M_value[oc]);
            int32_t q = (int32_t)((double)t * rshift_value[oc]);
            uint8_t r;
            switch (activation) {
            case YQ_LEAKY:
                if (q < 0) {
                    double v = round((double)q * 0.1) + (double)zp_out;
                    r = (uint8_t)(int32_t)v;
                } else {
                    r = (uint8_t)(q + zp_out);
                }
                break;
            case YQ_RELU6:
                r = (q <= 0) ? (uint8_t)zp_out : (uint8_t)(q + zp_out);
                break;
            default: /* LINEAR and RELU share one case, convolutional_layer.c:739-741 */
                r = (uint8_t)(q + zp_out);
                break;
            }
            out[idx] = r;
        }
    }
}

/* quant_stop dequantisation, src/convolutional_layer.c:752-760: f32 = (u8 - zp_out) * s_out */
void yq_oracle_dequant(const uint8_t *in, size_t n, int zp_out, float s_out, float *out)
{
    for (size_t i = 0; i < n; ++i) out[i] = (float)((int)in[i] - zp_out) * s_out;
}

/*
 * forward_maxpool_layer_quant, src/maxpool_layer.c:109-153.  `pad` is the layer's l.pad (= size-1 by
 * default, parser.c:415); window origin i*stride - pad/2; out-of-bounds taps contribute 0
 * ((uint8_t)(-FLT_MAX) == 0 and max starts at 0).
 */
void yq_oracle_maxpool(const uint8_t *in, int c, int h, int w, int size, int stride, int pad, uint8_t *out)
{
    int oh = (h + pad - size) / stride + 1;
    int ow = (w + pad - size) / stride + 1;
    int off = -pad / 2;
    for (int k = 0; k < c; ++k)
        for (int i = 0; i < oh; ++i)
            for (int j = 0; j < ow; ++j) {
                uint8_t m = 0;
                for (int a = 0; a < size; ++a)
                    for (int b = 0; b < size; ++b) {
                        int y = off + i * stride + a, x = off + j * stride + b;
                        if (y >= 0 && y < h && x >= 0 && x < w) {
                            uint8_t v = in[x + w * (y + h * k)];
                            if (v > m) m = v;
                        }
                    }
                out[j + ow * (i + oh * k)] = m;
            }
}

/* upsample_quant_cpu forward, src/blas.c:781-803: out[c, y, x] = in[c, y/stride, x/stride] */
void yq_oracle_upsample(const uint8_t *in, int c, int h, int w, int stride, uint8_t *out)
{
    for (int k = 0; k < c; ++k)
        for (int j = 0; j < h * stride; ++j)
            for (int i = 0; i < w * stride; ++i)
                out[(size_t)k * w * h * stride * stride + (size_t)j * w * stride + i] =
                    in[k * w * h + (j / stride) * w + i / stride];
}

/*
 * forward_yolo_layer inference part, src/yolo_layer.c:132-146 with entry_index :125-130 and
 * logistic_activate src/activations.h:32 (double exp, result narrowed to float):
 * per anchor a: channels a*(5+classes)+{0,1} and a*(5+classes)+{4..4+classes} <- 1/(1+exp(-x)).
 */
void yq_oracle_yolo(const float *in, int n_anchors, int classes, int h, int w, float *out)
{
    int per = 4 + classes + 1, hw = h * w;
    memcpy(out, in, sizeof(float) * (size_t)n_anchors * per * hw);
    for (int a = 0; a < n_anchors; ++a)
        for (int e = 0; e < per; ++e) {
            if (e == 2 || e == 3) continue;
            float *p = out + ((size_t)a * per + e) * hw;
            for (int i = 0; i < hw; ++i) p[i] = (float)(1. / (1. + exp(-(double)p[i])));
        }
}

/*
 * Layer-0 dynamic input quantiser, quant_weights_with_min_max_channel with size_channel=1
 * (src/blas.c:108-168, called at :279): mn=min(0,min x) mx=max(0,max x) s=(mx-mn)/255
 * zp=clamp(round(0 - mn/s)) ; u8 = clamp(round(x/s)+zp, 0, 255)   (this one saturates, :158).
 */
void yq_oracle_quantize_input(const float *x, size_t n, uint8_t *out, float *scale, uint8_t *zp)
{
    float mn = 0.f, mx = 0.f;
    for (size_t i = 0; i < n; ++i) {
        if (x[i] > mx) mx = x[i];
        if (x[i] < mn) mn = x[i];
    }
    float s = (mx - mn) / (255.f - 0.f);
    double izp = 0.f - mn / s;
    uint8_t z = izp < 0 ? 0 : izp > 255 ? 255 : (uint8_t)round(izp);
    for (size_t i = 0; i < n; ++i) {
        float t = roundf(x[i] / s) + z;
        int q = (int)t;
        out[i] = (uint8_t)(q < 0 ? 0 : q > 255 ? 255 : q);
    }
    *scale = s;
    *zp = z;
}

/* quant_multi_smaller_than_one_to_scale_and_shift, src/blas.c:387-418 */
static void mult_to_m0_shift(float m, int32_t *m0, int *shift)
{
    int s = 0;
    while (m < 0.5f) {
        m *= 2.0f;
        s++;
    }
    int64_t q = (int64_t)round((double)m * (double)(1ll << 31));
    if (q == (1ll << 31)) {
        q /= 2;
        s--;
    }
    *m0 = (int32_t)q;
    *shift = s;
}

/*
 * One conv layer's share of quantization_weights_and_activations (src/blas.c:282-334):
 * BN fold of the float bias (batch_normalize_bias, blas.c:594-601; the uint8 weights come from the
 * file and are NOT re-derived), weights_sum_int (:306-311, uint32 arithmetic), M = s_in*s_w/s_out in
 * float (:313), (M0, shift) (:314), M_value = 2^-31*M0, M0_right_shift_value = 2^-shift (:315-316),
 * biases_int32 = (int32)(bias/(s_in*s_w) + weights_sum_int) in float (:331-334).
 * The reference is built with -Ofast, so its float rounding may differ in the last ulp; tests
 * compare these outputs with the compiled reference's dump and report.
 */
void yq_oracle_prepare_conv(int n, int K, const uint8_t *weights, const uint8_t *zp_w, const float *s_w,
                            float s_in, int zp_in, float s_out, const float *bias, int bn, const float *bn_scales,
                            const float *bn_mean, const float *bn_var, int32_t *M0, int *shift, double *M_value,
                            double *rshift_value, int32_t *bias_i32)
{
    for (int oc = 0; oc < n; ++oc) {
        float b = bias[oc];
        /* batch_normalize_bias as ISO C evaluates it: the float product scales*mean, sqrt() in DOUBLE (its only prototype),
           .000001f promoted, double division and subtraction, one rounding back to float.  The reference's -Ofast build is
           free to narrow or reassociate this: measured here, its biases_int32 differs by +-1 from this form in ~1 of 3000
           channels with non-identity batch-norm statistics (and by as much from an all-float form, on other channels), so no
           single restatement reproduces that binary; tests against the compiled reference take ITS dumped parameters. */
        if (bn) b = (float)((double)b - (double)(bn_scales[oc] * bn_mean[oc]) / (sqrt((double)bn_var[oc]) + (double).000001f));
        uint32_t mult_zero_point = (uint32_t)(K * zp_in * (int)zp_w[oc]);
        int32_t wsum = 0;
        for (int k = 0; k < K; ++k) wsum += weights[(size_t)oc * K + k];
        int32_t weights_sum_int = (int32_t)(mult_zero_point - (uint32_t)(wsum * zp_in));
        float M = s_in * s_w[oc] / s_out;
        mult_to_m0_shift(M, &M0[oc], &shift[oc]);
        rshift_value[oc] = pow(2, -shift[oc]);
        M_value[oc] = pow(2, -31) * M0[oc];
        bias_i32[oc] = (int32_t)(b / (s_in * s_w[oc]) + (float)weights_sum_int);
    }
}

/* ------------------------------------------------------------------------------------------------
 * "next" row 8f-2: box decode + NMS
 * ---------------------------------------------------------------------------------------------- */
/*
 * get_yolo_detections + get_yolo_box + correct_yolo_boxes for ONE yolo layer (src/yolo_layer.c:83-91,247-273,
 * 316-343).  pred: the yolo layer's output, CHW [n_anchors*(5+classes)][lh][lw]; anchor_w/h: l.biases[2*mask[n]] and
 * [2*mask[n]+1].  Appends 5+classes floats per kept candidate (x, y, w, h, objectness, prob[classes]); returns the count.
 */
int yq_oracle_yolo_detections(const float *pred, int lw, int lh, int n_anchors, int classes, const float *anchor_w,
                              const float *anchor_h, int netw, int neth, int w, int h, int relative, float thresh,
                              float *dets)
{
    int cells = lw * lh, per = 5 + classes, count = 0;
    int new_w = 0, new_h = 0;
    if (((float)netw / w) < ((float)neth / h)) { new_w = netw; new_h = (h * netw) / w; }
    else { new_h = neth; new_w = (w * neth) / h; }
    for (int i = 0; i < cells; ++i) {
        int row = i / lw, col = i % lw;
        for (int n = 0; n < n_anchors; ++n) {
            const float *x = pred + (size_t)(n * per) * cells + i;
            float objectness = x[(size_t)4 * cells];
            if (objectness <= thresh) continue;
            float *d = dets + (size_t)count * per;
            float bx = (col + x[0]) / lw;
            float by = (row + x[(size_t)cells]) / lh;
            float bw = exp(x[(size_t)2 * cells]) * anchor_w[n] / netw;
            float bh = exp(x[(size_t)3 * cells]) * anchor_h[n] / neth;
            bx = (bx - (netw - new_w) / 2. / netw) / ((float)new_w / netw);
            by = (by - (neth - new_h) / 2. / neth) / ((float)new_h / neth);
            bw *= (float)netw / new_w;
            bh *= (float)neth / new_h;
            if (!relative) { bx *= w; bw *= w; by *= h; bh *= h; }
            d[0] = bx; d[1] = by; d[2] = bw; d[3] = bh; d[4] = objectness;
            for (int j = 0; j < classes; ++j) {
                float prob = objectness * x[(size_t)(5 + j) * cells];
                d[5 + j] = (prob > thresh) ? prob : 0;
            }
            ++count;
        }
    }
    return count;
}

static float ov1(float x1, float w1, float x2, float w2)          /* overlap, src/box.c:152-161 */
{
    float l1 = x1 - w1 / 2, l2 = x2 - w2 / 2;
    float left = l1 > l2 ? l1 : l2;
    float r1 = x1 + w1 / 2, r2 = x2 + w2 / 2;
    float right = r1 < r2 ? r1 : r2;
    return right - left;
}
static float iou(const float *a, const float *b)                  /* box_iou, src/box.c:163-182 */
{
    float w = ov1(a[0], a[2], b[0], b[2]), h = ov1(a[1], a[3], b[1], b[3]);
    float inter = (w < 0 || h < 0) ? 0 : w * h;
    float uni = a[2] * a[3] + b[2] * b[3] - inter;
    return inter / uni;
}

/*
 * do_nms_sort (src/box.c:58-89) on the flat detection array: per class, order by prob[k] descending, then every box
 * suppresses later boxes with IoU > thresh (prob[k] = 0).  Ties keep the order left by the previous class's sort
 * (stable), starting from detection order for class 0 -- see the comment in the loop.
 * The array itself is NOT reordered (only probabilities are zeroed), so it stays comparable element by element.
 */
void yq_oracle_nms_sort(float *dets, int total, int classes, float thresh)
{
    int per = 5 + classes;
    int *order = malloc(sizeof(int) * (size_t)(total > 0 ? total : 1));
    for (int i = 0; i < total; ++i) order[i] = i;
    for (int k = 0; k < classes; ++k) {
        /* the reference qsort()s the SAME array once per class, so each class starts from the previous class's order;
           glibc's qsort is a stable merge sort for arrays of this size, and uint8-quantised heads make exact probability
           ties common, so the carried-over order decides which of two tied boxes suppresses the other.  A stable
           insertion sort on the carried permutation reproduces it. */
        for (int i = 1; i < total; ++i) {
            int v = order[i], j = i - 1;
            float pv = dets[(size_t)v * per + 5 + k];
            while (j >= 0 && dets[(size_t)order[j] * per + 5 + k] < pv) { order[j + 1] = order[j]; --j; }
            order[j + 1] = v;
        }
        for (int i = 0; i < total; ++i) {
            float *a = dets + (size_t)order[i] * per;
            if (a[5 + k] == 0) continue;
            for (int j = i + 1; j < total; ++j) {
                float *b = dets + (size_t)order[j] * per;
                if (iou(a, b) > thresh) b[5 + k] = 0;
            }
        }
    }
    free(order);
}

/* ------------------------------------------------------------------------------------------------
 * "next" row 8f-3: quantized shortcut -- an EXTENSION, NOT in the reference.
 *
 * The reference has only the float shortcut (src/shortcut_layer.c:62-67: copy the input, add the `from` layer's output
 * through shortcut_cpu src/blas.c:456-477, then activate) and forward_network would hand the next quantized convolution a
 * stale input_uint8 (src/network.c:248-255): SURVEY 0.10, Appendix F.  So nothing in the reference can pin this function;
 * it IS the specification the CUDA path is held to ("parity unpinned by the reference" for this one layer type).
 *
 * Integer spec (dequant-add-requant with the layer's stored output scale, all integer):
 *   a = previous layer's uint8 output (s_a, zp_a);  b = the `from` layer's uint8 output (s_b, zp_b), same shape
 *   Ka = round(s_a / s_out * 2^16),  Kb = round(s_b / s_out * 2^16)          (double division of the float scales)
 *   t  = (a - zp_a) * Ka + (b - zp_b) * Kb                                   (exact int32: Ka, Kb < 2^22)
 *   q  = (t + 2^15) >> 16                                                    (arithmetic shift: round half up)
 *   out = clamp(q + zp_out, 0, 255)                                          (saturates; activation = linear only)
 * ---------------------------------------------------------------------------------------------- */
int yq_oracle_shortcut_mult(float s_x, float s_out, int32_t *K)
{
    if (!(s_x > 0.f) || !(s_out > 0.f)) return -1;
    double k = round((double)s_x / (double)s_out * 65536.0);
    if (!(k >= 1.0) || k >= 4194304.0) return -1;
    *K = (int32_t)k;
    return 0;
}

void yq_oracle_shortcut(const uint8_t *a, const uint8_t *b, size_t n, int zp_a, int zp_b, int32_t Ka, int32_t Kb,
                        int zp_out, uint8_t *out)
{
    for (size_t i = 0; i < n; ++i) {
        int32_t t = ((int32_t)a[i] - zp_a) * Ka + ((int32_t)b[i] - zp_b) * Kb;
        int32_t q = (t + 32768) >> 16;
        int32_t r = q + zp_out;
        out[i] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
    }
}

/*
 * letterbox_image, src/image.c:812-831 = resize_image (:1199-1245: horizontal pass into `part`, vertical pass with set_pixel +
 * add_pixel) into a w x h canvas filled with .5 (:695) at ((w - new_w)/2, (h - new_h)/2) (embed_image :428-439).  Plain float
 * arithmetic in the reference's order.  (The reference's -Ofast build may fuse or reorder the multiply-adds: compared with the
 * compiled reference the quantised bytes may differ by one LSB at rounding ties)
 */
void yq_oracle_letterbox(const float *im, int c, int ih, int iw, float *boxed, int h, int w)
{
    int new_w = iw, new_h = ih;
    if (((float)w / iw) < ((float)h / ih)) { new_w = w; new_h = (ih * w) / iw; }
    else { new_h = h; new_w = (iw * h) / ih; }
    float *part = malloc(sizeof(float) * (size_t)c * ih * new_w);
    float *res = malloc(sizeof(float) * (size_t)c * new_h * new_w);
    float w_scale = (float)(iw - 1) / (new_w - 1);
    float h_scale = (float)(ih - 1) / (new_h - 1);
    for (int k = 0; k < c; ++k)
        for (int r = 0; r < ih; ++r)
            for (int x = 0; x < new_w; ++x) {
                volatile float val;
                if (x == new_w - 1 || iw == 1) {
                    val = im[((size_t)k * ih + r) * iw + iw - 1];
                } else {
                    float sx = x * w_scale;
                    int ix = (int)sx;
                    float dx = sx - ix;
                    volatile float a = (1 - dx) * im[((size_t)k * ih + r) * iw + ix];
                    volatile float b = dx * im[((size_t)k * ih + r) * iw + ix + 1];
                    val = a + b;
                }
                part[((size_t)k * ih + r) * new_w + x] = val;
            }
    for (int k = 0; k < c; ++k)
        for (int r = 0; r < new_h; ++r) {
            float sy = r * h_scale;
            int iy = (int)sy;
            float dy = sy - iy;
            for (int x = 0; x < new_w; ++x) {
                volatile float val = (1 - dy) * part[((size_t)k * ih + iy) * new_w + x];
                if (!(r == new_h - 1 || ih == 1)) {
                    volatile float t = dy * part[((size_t)k * ih + iy + 1) * new_w + x];
                    val = val + t;
                }
                res[((size_t)k * new_h + r) * new_w + x] = val;
            }
        }
    for (size_t i = 0; i < (size_t)c * h * w; ++i) boxed[i] = .5f;
    int ox = (w - new_w) / 2, oy = (h - new_h) / 2;
    for (int k = 0; k < c; ++k)
        for (int y = 0; y < new_h; ++y)
            for (int x = 0; x < new_w; ++x) boxed[((size_t)k * h + oy + y) * w + ox + x] = res[((size_t)k * new_h + y) * new_w + x];
    free(part);
    free(res);
}
