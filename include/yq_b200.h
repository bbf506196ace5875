/*
 * yq_b200.h -- C ABI of libyq_b200.so: the B200-native (sm_100a) INT8 inference hot path of the
 * quantized darknet fork ArtyZe/yolo_quantization.
 *
 * This is the drop-in boundary.  Plain pointers, sizes and opaque handles only; no torch, no C++.
 * Every entry point names the reference interface it replaces (paths relative to the reference
 * repository).  INTEGRATION.md shows the stubs a maintainer adds to the reference
 * (make_*_layer / forward_*_gpu / forward_network_gpu) to bind these.
 *
 * Conventions
 *  - Device activations are uint8 NHWC with a channel stride `cs` >= c (yq_channel_stride(c));
 *    pad channels are ZERO.  The reference's host-visible layout is CHW (src/im2col.c:13); the
 *    yq_nchw_to_nhwc_* / yq_nhwc_to_nchw_* kernels convert at the two places data crosses the
 *    boundary (net->input_uint8 in, l.output / debug pulls out).
 *  - All functions return 0 on success, non-zero on failure; yq_last_error() returns the message.
 *    The reference has no return codes: check_error() prints, assert(0)s and exit(-1)s
 *    (src/cuda.c:27-49).  yq_set_abort_on_error(1) reproduces that convention.
 *  - `stream` arguments are cudaStream_t passed as void* (NULL = the legacy default stream, which is
 *    what the reference uses everywhere, src/cuda.c).
 *  - There is NO CPU fallback: without a CUDA device every compute entry point fails loudly.
 */
#ifndef YQ_B200_H
#define YQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YQ_API __attribute__((visibility("default")))

/* Same numeric values as the reference's ACTIVATION enum (include/darknet.h:87-89) so a binding can
 * pass l.activation straight through. */
typedef enum {
    YQ_LOGISTIC = 0,
    YQ_RELU = 1,
    YQ_LINEAR = 3,
    YQ_RELU6 = 8,
    YQ_LEAKY = 9
} yq_activation;

/* ------------------------------------------------------------------------------------------------
 * errors / device  (replaces check_error, cuda_set_device, cuda_get_device: src/cuda.c:11-49)
 * ---------------------------------------------------------------------------------------------- */
YQ_API const char *yq_last_error(void);
YQ_API void yq_set_abort_on_error(int enable);
YQ_API int yq_device_count(void);
YQ_API int yq_set_device(int device);
YQ_API const char *yq_version(void);

/* ------------------------------------------------------------------------------------------------
 * device memory  (replaces cuda_make_array / cuda_free / cuda_push_array_int8 /
 * cuda_pull_array_int8: src/cuda.c:90-104,145-179 -- without the sizeof(float) over-read of
 * cuda_make_array on uint8 buffers)
 * ---------------------------------------------------------------------------------------------- */
YQ_API void *yq_cuda_malloc(size_t bytes);
YQ_API int yq_cuda_free(void *dev);
YQ_API int yq_cuda_push(void *dev, const void *host, size_t bytes, void *stream);
YQ_API int yq_cuda_pull(void *host, const void *dev, size_t bytes, void *stream);
YQ_API int yq_cuda_memset(void *dev, int value, size_t bytes, void *stream);
YQ_API int yq_stream_synchronize(void *stream);

/* page-locked host memory for the host-buffer entry points (yq_network_submit_u8 / predict_*): cudaHostAlloc; write_combined = 1
 * asks for write-combined pages (host writes stream, host READS are very slow): for input buffers the CPU only fills. */
YQ_API void *yq_host_alloc(size_t bytes, int write_combined);
YQ_API int yq_host_free(void *host);

/* channel stride used for a c-channel NHWC activation tensor: 4 for c <= 4, else c rounded up to 16 */
YQ_API int yq_channel_stride(int c);

/* ------------------------------------------------------------------------------------------------
 * quantized convolution
 *   replaces forward_convolutional_layer_quant_inputi_outputi (src/convolutional_layer.c:694-761)
 *            = im2col_cpu_uint8 (src/im2col.c:26-50) + 2x gemm_nn_uint8_int32_te (src/gemm.c:279-299)
 *              + requantize/activation loop (:726-751) + quant_stop dequant (:752-760)
 *   binds as  l.forward_gpu of a CONVOLUTIONAL layer (slot assigned at src/convolutional_layer.c:320-340;
 *             prototype forward_convolutional_layer_quant_gpu, src/convolutional_layer.h:15)
 * ---------------------------------------------------------------------------------------------- */

/* Host-side description of one prepared layer; field names follow `struct layer`
 * (include/darknet.h:172-226) after quantization_weights_and_activations (src/blas.c:259-346). */
typedef struct yq_conv_desc {
    int h, w, c;                 /* l.h l.w l.c                                                          */
    int n, size, stride, pad;    /* l.n l.size l.stride l.pad (pad already = size/2 when cfg pad=1)      */
    int activation;              /* l.activation (yq_activation values)                                   */
    int quant_stop_flag;         /* l.quant_stop_flag: also emit float (u8 - zp_out) * s_out              */
    int zp_in;                   /* l.input_data_uint8_zero_point[0]  (pad value of im2col, im2col.c:5-14) */
    int zp_out;                  /* l.activ_data_uint8_zero_point[0]                                      */
    float s_out;                 /* l.activ_data_uint8_scales[0]                                          */
    const uint8_t *weights_uint8;          /* l.weights_uint8, OIHW [n][c][size][size]  (host)            */
    const uint8_t *weight_zero_point;      /* l.weight_data_uint8_zero_point [n]        (host)            */
    const int32_t *biases_int32;           /* l.biases_int32 [n]                        (host)            */
    const double *M_value;                 /* l.M_value [n]  = M0 * 2^-31               (host)            */
    const double *M0_right_shift_value;    /* l.M0_right_shift_value [n] = 2^-shift     (host)            */
    int saturate;                /* 0 = reference semantics: uint8 store WRAPS (convolutional_layer.c:737-749);
                                    1 = clamp to [0,255] (what the MKL variant does at :594)              */
} yq_conv_desc;

typedef struct yq_conv_layer yq_conv_layer;   /* opaque: packed device weights + per-channel params */

/* replaces the GPU block of make_convolutional_layer (src/convolutional_layer.c:320-340) and
 * push_convolutional_layer (src/convolutional_kernels.cu:339-350): packs weights into the kernel layout
 * once and uploads the per-channel parameters.  Returns NULL on failure. */
YQ_API yq_conv_layer *yq_make_convolutional_layer_quant(const yq_conv_desc *desc);
YQ_API void yq_free_convolutional_layer_quant(yq_conv_layer *l);
YQ_API int yq_conv_out_h(const yq_conv_layer *l);
YQ_API int yq_conv_out_w(const yq_conv_layer *l);
/* 0 = SIMT dp4a implicit GEMM, 1 = tcgen05 kind::i8 implicit GEMM; -1 (default) = choose per layer */
YQ_API int yq_conv_set_kernel(yq_conv_layer *l, int kind);
YQ_API int yq_conv_get_kernel(const yq_conv_layer *l);

/* in_u8   : device, NHWC [batch][h][w][yq_channel_stride(c)]
 * out_u8  : device, NHWC [batch][out_h][out_w][yq_channel_stride(n)]           (l.output_uint8_final)
 * out_f32 : device, NCHW [batch][n][out_h][out_w] or NULL; written when quant_stop_flag   (l.output)
 * out_acc : device, NHWC [batch][out_h][out_w][yq_channel_stride(n)] int32 or NULL -- the exact-integer
 *           accumulator (l.output_int32) for parity checks                                         */
YQ_API int yq_forward_convolutional_layer_quant_gpu(yq_conv_layer *l, const uint8_t *in_u8, uint8_t *out_u8,
                                                    float *out_f32, int32_t *out_acc, int batch, void *stream);

/* The same with the FOLLOWING max-pool layer (size 2, stride 2, default padding: maxpool_layer.c:109-153 with
 * l.pad = 1) fused into the epilogue: out_pool is NHWC [batch][(out_h+1)/2][(out_w+1)/2][yq_channel_stride(n)].
 * Either of out_u8 / out_pool may be NULL (not both).  Only flavours for which yq_conv_can_fuse_maxpool() is 1. */
YQ_API int yq_forward_convolutional_layer_quant_pool_gpu(yq_conv_layer *l, const uint8_t *in_u8, uint8_t *out_u8,
                                                         uint8_t *out_pool, float *out_f32, int32_t *out_acc, int batch,
                                                         void *stream);
YQ_API int yq_conv_can_fuse_maxpool(const yq_conv_layer *l);

/* (declared below) the same convolution between halo-padded tensors: yq_forward_convolutional_layer_quant_geom_gpu */

/* A halo-padded NHWC activation tensor: pixel (n, y, x) lives at
 *     base + (((size_t)n * rows_h + y + pad) * pitch_w + x + pad) * yq_channel_stride(c)
 * The halo (and any slack right / below the image) holds the CONSUMER's input zero point, which is what
 * im2col_cpu_uint8 pads with (src/im2col.c:5-14).  pad = 0, pitch_w = w, rows_h = h is the plain tensor. */
typedef struct yq_act_geom {
    int pad, pitch_w, rows_h;
} yq_act_geom;

/* forward_convolutional_layer_quant_inputi_outputi (src/convolutional_layer.c:694-761) between halo-padded tensors, any
 * stride the reference's im2col takes (src/im2col.c:26-50; stride 1 and 2 have a tcgen05 flavour, the rest run SIMT):
 * in_geom / out_geom describe the tensors (NULL = plain); only the interior of `out` is written.  in_halo_fill is the byte
 * the input's halo is known to hold, or -1: when it equals the layer's zp_in (and the halo is at least l.pad wide) the
 * kernel reads its padding from the halo (im2col.c:5-14 pads with zp_in), otherwise it treats everything outside the image
 * as out of bounds and restores zp_in * sum(w - zp_w) per border tap in the epilogue.  Flavours that only take plain
 * tensors (SIMT, c <= 32) fail on a padded geometry: yq_conv_geom_supported() says which kind the layer has. */
YQ_API int yq_conv_geom_supported(const yq_conv_layer *l);
/* 1: the layer also has the resident-bank kernel in patch mode (3x3, stride 2, n <= 255, filter bank resident in shared memory): handed a
 * halo-padded input whose halo holds its zp_in (in_halo_fill == zp_in) and no side outputs, yq_forward_convolutional_layer_quant_geom_gpu
 * runs it there (persistent, one tcgen05.commit per tile or filter row) instead of the one-tile-per-CTA per-tap / small-c flavours. */
YQ_API int yq_conv_patch_supported(const yq_conv_layer *l);
/* 1 when the layer's flavour can at least WRITE a halo-padded output tensor (plain input): the per-tap flavour, and the small-c
 * flavour (c <= 32), whose threads store their pixels themselves. */
YQ_API int yq_conv_out_geom_supported(const yq_conv_layer *l);
YQ_API int yq_forward_convolutional_layer_quant_geom_gpu(yq_conv_layer *l, const uint8_t *in_u8, const yq_act_geom *in_geom, int in_halo_fill,
                                                         uint8_t *out_u8, const yq_act_geom *out_geom, float *out_f32, int32_t *out_acc,
                                                         int batch, void *stream);

/* The "rows" flavour: 3x3 / stride 1 / pad 1 convolution + RELU6 + the following 2x2/2 max-pool in ONE launch
 * (convolutional_layer.c:694-751 + maxpool_layer.c:109-153 fused), for c <= 32 and c = 64 (n = 64 or 128).  It reads a halo-padded input
 * (geometry from yq_conv_rows_input_geom) with no im2col gather and writes only the pooled tensor, into any
 * geometry.  yq_conv_rows_supported() is nonzero when the layer has it: 2 when the filters went in as two signed blocks
 * h + l = w - zp_w (the accumulator needs no zero-point correction), 1 when some w - zp_w = 255 forced the all-ones-rows form. */
YQ_API int yq_conv_rows_supported(const yq_conv_layer *l);
YQ_API int yq_conv_rows_input_geom(const yq_conv_layer *l, yq_act_geom *geom);
YQ_API size_t yq_act_geom_bytes(const yq_act_geom *geom, int batch, int c);
YQ_API int yq_forward_convolutional_layer_quant_rows_pool_gpu(yq_conv_layer *l, const uint8_t *in_padded, uint8_t *out_pool,
                                                              const yq_act_geom *out_geom, int batch, void *stream);
/* The same launch reading the network input where the reference keeps it -- net.input_uint8 as [batch][3][h][w] planes
 * (network.c:248-250) -- with no layout transform before it: TMA fetches the three planes of a tile (zero outside the image,
 * hence zp_in = 0) and the kernel's producer warp interleaves them.  Needs c = 3, zp_in = 0, w % 16 = 0, w >= 64 and a 16-byte
 * aligned pointer; yq_conv_rows_nchw_supported() says whether the layer qualifies. */
YQ_API int yq_conv_rows_nchw_supported(const yq_conv_layer *l);
YQ_API int yq_forward_convolutional_layer_quant_rows_pool_nchw_gpu(yq_conv_layer *l, const uint8_t *in_nchw, uint8_t *out_pool,
                                                                   const yq_act_geom *out_geom, int batch, void *stream);

/* The "flat" flavour (stride 1, size 1 or 3, pad = size/2, c % 64 == 0): input and output are FLAT halo-padded
 * tensors of identical geometry {pad 1, pitch_w w+1, rows_h h+1} (yq_act_geom_flat): one shared halo pixel after
 * every image row, one shared halo row after every image, one trailing halo row (yq_act_geom_bytes counts it).
 * Every filter tap reads the same shared-memory patch (no per-tap activation loads, no border fix-ups: the halo
 * holds zp_in, src/im2col.c:5-14).  The kernel writes `halo_fill` (the CONSUMER's input zero point) at the halo
 * positions of its output, so a chain of flat convolutions keeps its halos by itself.
 * out_f32 / out_acc as in yq_forward_convolutional_layer_quant_gpu (dense, un-padded). */
YQ_API int yq_conv_flat_supported(const yq_conv_layer *l);
YQ_API int yq_act_geom_flat(int h, int w, yq_act_geom *geom);
YQ_API int yq_forward_convolutional_layer_quant_flat_gpu(yq_conv_layer *l, const uint8_t *in_flat, uint8_t *out_flat, int halo_fill,
                                                         float *out_f32, int32_t *out_acc, int batch, void *stream);

/* A flat convolution with the FOLLOWING quantized shortcut (the extension layer below) fused into its epilogue: out_flat receives
 * the SHORTCUT's output  clamp((((a - zp_a) * Ka + (b - zp_b) * Kb + 2^15) >> 16) + zp_out, 0, 255)  where a = this convolution's
 * uint8 result (never stored) and b = from_flat at the same position / channel (same flat geometry, same channel count);
 * halo_fill = the byte the shortcut's consumers pad with.  Saves the shortcut's launch: 2 bytes read + 1 written per element. */
YQ_API int yq_conv_flat_shortcut_supported(const yq_conv_layer *l);
YQ_API int yq_forward_convolutional_layer_quant_flat_shortcut_gpu(yq_conv_layer *l, const uint8_t *in_flat, const uint8_t *from_flat,
                                                                  uint8_t *out_flat, int halo_fill, int zp_from, int Ka, int Kb,
                                                                  int zp_out_shortcut, int batch, void *stream);

/* A flat 1x1 convolution with the FOLLOWING stride-2 upsample layer fused (forward_upsample_layer, src/upsample_layer.c:92-101 ->
 * upsample_cpu, src/blas.c:334-351: every pixel four times): out_up_flat = the flat tensor of (2h x 2w) pixels with the layer's output
 * channel stride.  Only its interior is written (the halo keeps the caller's fill); the layer's own tensor is not written.
 * Runs on the pointwise flavour (yq_conv_tc_pw.cu): 1x1, n <= 255, filter bank resident in shared memory. */
YQ_API int yq_conv_flat_up2_supported(const yq_conv_layer *l);
YQ_API int yq_forward_convolutional_layer_quant_flat_up2_gpu(yq_conv_layer *l, const uint8_t *in_flat, uint8_t *out_up_flat, int batch, void *stream);

/* A flat convolution behind a [route] that is never materialised: the input is the channel concatenation
 * [in_first (c_first channels) | in_second (c - c_first channels)] of two flat tensors of the layer's input geometry, as
 * forward_route_layer (src/route_layer.c:77-95) would have copied them; the convolution's patch loads pick the tensor per
 * channel chunk instead.  Both halos must hold the layer's input zero point.  yq_conv_flat_cat_supported: 1 when the layer's
 * kernel can do this for a split at c_first (the CTA-pair flavour, c >= 256, or the pointwise flavour of narrow 1x1 layers; c_first a
 * multiple of the kernel's channel chunk). */
YQ_API int yq_conv_flat_cat_supported(const yq_conv_layer *l, int c_first);
YQ_API int yq_forward_convolutional_layer_quant_flat_cat_gpu(yq_conv_layer *l, const uint8_t *in_first, int c_first, const uint8_t *in_second,
                                                             uint8_t *out_flat, int halo_fill, int batch, void *stream);

/* A quant_stop head with the FOLLOWING yolo layer fused (forward_yolo_layer's inference part, src/yolo_layer.c:132-146):
 * out_yolo [batch][n][h][w] receives the yolo layer's output.  The head's float values (u8 - zp_out) * s_out
 * (convolutional_layer.c:752-760) take only 256 values, so logistic_activate (src/activations.h:32) is a 256-entry
 * table built on the host in the reference's own double arithmetic.  out_f32 (the conv's own l.output) may be NULL. */
YQ_API int yq_forward_convolutional_layer_quant_flat_yolo_gpu(yq_conv_layer *l, const uint8_t *in_flat, uint8_t *out_flat,
                                                              int halo_fill, float *out_f32, float *out_yolo, int classes,
                                                              int32_t *out_acc, int batch, void *stream);

/* ------------------------------------------------------------------------------------------------
 * memory-bound layers (all device pointers, uint8 NHWC, channel stride yq_channel_stride(c))
 * ---------------------------------------------------------------------------------------------- */
/* replaces forward_maxpool_layer_quant (src/maxpool_layer.c:109-153); `pad` is l.pad (size-1 by default) */
YQ_API int yq_forward_maxpool_layer_quant_gpu(const uint8_t *in, uint8_t *out, int batch, int h, int w, int c,
                                              int size, int stride, int pad, void *stream);
/* replaces forward_upsample_layer_quant + upsample_quant_cpu (src/upsample_layer.c:96-113, src/blas.c:781-803) */
YQ_API int yq_forward_upsample_layer_quant_gpu(const uint8_t *in, uint8_t *out, int batch, int h, int w, int c,
                                               int stride, void *stream);
/* replaces forward_route_layer_quant + copy_cpu_uint8 (src/route_layer.c:107-130, src/blas.c:656-660):
 * channel concat of n_inputs tensors of identical h, w; in_c[i] real channels each. */
YQ_API int yq_forward_route_layer_quant_gpu(const uint8_t *const *inputs, const int *in_c, int n_inputs,
                                            uint8_t *out, int batch, int h, int w, void *stream);
/* the same three layers between halo-padded tensors (NULL geometry = plain); only interiors are read / written */
YQ_API int yq_forward_maxpool_layer_quant_geom_gpu(const uint8_t *in, const yq_act_geom *in_geom, uint8_t *out,
                                                   const yq_act_geom *out_geom, int batch, int h, int w, int c, int size,
                                                   int stride, int pad, void *stream);
YQ_API int yq_forward_upsample_layer_quant_geom_gpu(const uint8_t *in, const yq_act_geom *in_geom, uint8_t *out,
                                                    const yq_act_geom *out_geom, int batch, int h, int w, int c, int stride,
                                                    void *stream);
YQ_API int yq_forward_route_layer_quant_geom_gpu(const uint8_t *const *inputs, const yq_act_geom *in_geoms, const int *in_c,
                                                 int n_inputs, uint8_t *out, const yq_act_geom *out_geom, int batch, int h,
                                                 int w, void *stream);
/* forward_upsample_layer_quant (src/upsample_layer.c:96-113, src/blas.c:781-803) folded into the forward_route_layer_quant
 * (src/route_layer.c:107-130) that consumes it: input k holds (h / in_up[k]) x (w / in_up[k]) pixels and is read as
 * in[y / up][x / up]; in_up = NULL or all ones is the plain route.  in_geoms[k] describes the stored (small) tensor. */
YQ_API int yq_forward_route_layer_quant_up_gpu(const uint8_t *const *inputs, const yq_act_geom *in_geoms, const int *in_c,
                                               const int *in_up, int n_inputs, uint8_t *out, const yq_act_geom *out_geom,
                                               int batch, int h, int w, void *stream);
/* The same route with the inputs named by input_mask only (bit k = input k): the other inputs' channels of `out` are left
 * untouched, so the copies of one route_layer (src/route_layer.c:107-117 loops over l.n inputs) may be issued as separate
 * launches, each as soon as its input exists. */
YQ_API int yq_forward_route_layer_quant_part_gpu(const uint8_t *const *inputs, const yq_act_geom *in_geoms, const int *in_c,
                                                 const int *in_up, int n_inputs, unsigned input_mask, uint8_t *out,
                                                 const yq_act_geom *out_geom, int batch, int h, int w, void *stream);
/* Quantized shortcut -- an EXTENSION for the full yolov3 (BASELINE configs[4]): the reference's shortcut is float
 * only (src/shortcut_layer.c:62-67, shortcut_cpu src/blas.c:456-477) and cannot feed a quantized convolution
 * (src/network.c:248-255), so this layer's integer arithmetic is defined here and pinned by oracle/yq_oracle.c:yq_oracle_shortcut:
 *     Ka = round(s_a / s_out * 2^16)   Kb = round(s_b / s_out * 2^16)                 (yq_shortcut_multiplier, in [1, 2^22))
 *     out = clamp((((a - zp_a) * Ka + (b - zp_b) * Kb + 2^15) >> 16) + zp_out, 0, 255)       (activation = linear)
 * a = the previous layer's output (net.input_uint8), b = the `from` layer's output_uint8_final, same h, w, c. */
YQ_API int yq_shortcut_multiplier(float s_x, float s_out, int32_t *K);
YQ_API int yq_forward_shortcut_layer_quant_gpu(const uint8_t *a, const uint8_t *b, uint8_t *out, int batch, int h, int w, int c,
                                               int zp_a, int zp_b, int Ka, int Kb, int zp_out, void *stream);
YQ_API int yq_forward_shortcut_layer_quant_geom_gpu(const uint8_t *a, const yq_act_geom *a_geom, const uint8_t *b,
                                                    const yq_act_geom *b_geom, uint8_t *out, const yq_act_geom *out_geom, int batch,
                                                    int h, int w, int c, int zp_a, int zp_b, int Ka, int Kb, int zp_out, void *stream);
/* replaces forward_yolo_layer's inference part (src/yolo_layer.c:132-146): float NCHW in/out,
 * logistic on channels {0,1} and {4..4+classes} of each anchor. */
YQ_API int yq_forward_yolo_layer_gpu(const float *in, float *out, int batch, int n_anchors, int classes, int h,
                                     int w, void *stream);

/* replaces the layer-0 input quantiser quant_weights_with_min_max_channel(1, net->input, net->input_uint8, ...)
 * (src/blas.c:108-168, called at :279) for a batch of float CHW images already on the device: per image
 * s = (max(0,max x) - min(0,min x)) / 255, zp = clamp(round(-min/s)), u8 = clamp(round(x/s) + zp).
 * scales [batch] and zero_points [batch] receive (s, zp) (zp = -1 flags the all-zero image the reference asserts on);
 * scratch: 2*batch ints of device memory.  n = c*h*w elements per image. */
YQ_API int yq_quantize_input_gpu(const float *in_f32_chw, uint8_t *out_u8_chw, float *scales, int *zero_points, int *scratch,
                                 int batch, int n, void *stream);

/* letterbox_image (src/image.c:812-831: resize_image :1199-1245 into a .5-filled w x h canvas, embed_image :428-439) for a
 * batch of float CHW images of one size [batch][c][ih][iw] already on the device -> [batch][c][h][w].  Float operations in the
 * reference's order, each rounded on its own (the test oracle's plain-C restatement matches the compiled reference bit for bit). */
YQ_API int yq_letterbox_image_gpu(const float *in_chw, int batch, int c, int ih, int iw, float *out_chw, int h, int w, void *stream);

/* Layer 0 behind the dynamic input quantiser when the images of a batch quantise DIFFERENTLY: the reference derives
 * (s_in, zp_in) from its one image (src/blas.c:279) and M, biases_int32 (blas.c:301-334) and the im2col padding value
 * (src/im2col.c:5-14) of layer 0 follow; here they are per-image device tables: biases_int32 [batch][table_pitch],
 * multiplier = M_value * M0_right_shift_value [batch][table_pitch] (doubles), zp_in [batch].  Plain tensors, generic flavour. */
YQ_API int yq_forward_convolutional_layer_quant_per_image_gpu(yq_conv_layer *l, const uint8_t *in_u8, uint8_t *out_u8, int32_t *out_acc,
                                                              int batch, const int32_t *biases_int32_dev, const double *multiplier_dev,
                                                              const uint8_t *zp_in_dev, int table_pitch, void *stream);

/* layout conversion at the boundary (reference tensors are CHW per image) */
YQ_API int yq_nchw_to_nhwc_u8(const uint8_t *in_nchw, uint8_t *out_nhwc, int batch, int c, int h, int w,
                              void *stream);
YQ_API int yq_nhwc_to_nchw_u8(const uint8_t *in_nhwc, uint8_t *out_nchw, int batch, int c, int h, int w,
                              void *stream);
YQ_API int yq_nhwc_to_nchw_i32(const int32_t *in_nhwc, int32_t *out_nchw, int batch, int c, int h, int w,
                               void *stream);
/* the same conversions from / to a halo-padded tensor (only the h x w interior is read / written) */
YQ_API int yq_nchw_to_nhwc_u8_geom(const uint8_t *in_nchw, uint8_t *out_nhwc, int batch, int c, int h, int w,
                                   const yq_act_geom *geom, void *stream);
YQ_API int yq_nhwc_to_nchw_u8_geom(const uint8_t *in_nhwc, uint8_t *out_nchw, int batch, int c, int h, int w,
                                   const yq_act_geom *geom, void *stream);

/* ------------------------------------------------------------------------------------------------
 * network level -- host-side mirror of the reference's public API for this path
 *   load_network / parse_network_cfg / load_weights   (src/network.c:49-57, src/parser.c:682-815,1201-1305)
 *   quantization_weights_and_activations              (src/blas.c:259-346; done once inside yq_load_network)
 *   set_batch_network                                 (src/network.c:415-432; batch is fixed at load here
 *                                                      because buffers are sized at parse time, Appendix E.6)
 *   network_predict / forward_network                 (src/network.c:229-261,570-581)
 * ---------------------------------------------------------------------------------------------- */
typedef struct yq_network yq_network;

typedef struct yq_layer_info {
    int type;                 /* 0 conv, 1 maxpool, 2 route, 3 upsample, 4 yolo, 5 shortcut (extension) */
    int c, h, w;              /* input  */
    int out_c, out_h, out_w;  /* output */
    int n, size, stride, pad, activation, batch_normalize, quant_stop_flag;
    float s_in, s_out;
    int zp_in, zp_out;
    int kernel;               /* conv only: 0 SIMT, 1 tcgen05 (per-tap TMA / small-c), 2 tcgen05 flat strip, 3 tcgen05 rows */
    int classes, n_anchors;   /* yolo only */
    int fused;                /* conv: 0 own launch writing the conv tensor, 1 following maxpool fused into the epilogue,
                                 2 rows flavour (halo input, pooled tensor only), 3 following yolo layer fused,
                                 4 following quantized shortcut fused (the launch writes the shortcut's tensor only),
                                 6 following upsample fused (the launch writes the upsampled tensor only);
                                 maxpool / upsample / yolo / shortcut: 1 = produced by a neighbour's launch;
                                 route: 5 = never written, the conv behind it reads its inputs itself */
} yq_layer_info;

/* batch <= 0 keeps the cfg's [net] batch.  Returns NULL on failure (see yq_last_error). */
YQ_API yq_network *yq_load_network(const char *cfg, const char *weights, int batch, int device);
YQ_API void yq_free_network(yq_network *net);
YQ_API int yq_network_num_layers(const yq_network *net);
YQ_API int yq_network_batch(const yq_network *net);
YQ_API int yq_network_input_dims(const yq_network *net, int *c, int *h, int *w);
YQ_API int yq_network_layer_info(const yq_network *net, int i, yq_layer_info *info);
/* Layer-0 input quantisation (s_in, zp_in).  The reference derives it per image on the host
 * (quant_weights_with_min_max_channel, src/blas.c:108-168 via :279) and then overwrites the values
 * loaded from the file; here it is set explicitly (default: the file's values). Re-runs the layer-0 prep. */
YQ_API int yq_network_set_input_quant(yq_network *net, float s_in, int zp_in);
/* keep per-layer int32 accumulators / uint8 outputs for yq_network_pull_layer (parity checks) */
YQ_API int yq_network_set_debug(yq_network *net, int keep_acc);
/* fuse conv -> maxpool(2,2) pairs into one launch where the conv flavour supports it (default 1) */
YQ_API int yq_network_set_fusion(yq_network *net, int enable);
/* force one conv kernel flavour for every conv layer (-1 auto, 0 SIMT, 1 tcgen05) */
YQ_API int yq_network_set_conv_kernel(yq_network *net, int kind);

/* forward_network on device-resident input: in_u8_nchw is [batch][c][h][w] uint8 on the device
 * (the reference's net->input_uint8 layout).  Asynchronous on the network's stream. */
YQ_API int yq_forward_network_device(yq_network *net, const uint8_t *in_u8_nchw);
/* network_predict with HOST buffers: H2D of the uint8 input, forward, D2H of every yolo head into
 * out_f32 (heads concatenated in layer order, each [batch][out_c][out_h][out_w]); synchronous. */
YQ_API int yq_network_predict_u8(yq_network *net, const uint8_t *in_u8_nchw_host, float *out_f32_host);
/* network_predict with HOST FLOAT images (the reference's net->input): H2D, device input quantiser
 * (yq_quantize_input_gpu), layer-0 re-prep when (s_in, zp_in) changed -- exactly what test_detector does per image
 * (examples/detector.c:915-922) -- forward, D2H.  Layer 0's multipliers, biases_int32 and padding value depend on
 * (s_in, zp_in): when every image of the batch quantises to the same pair the planned (tensor-core) layer 0 runs, re-prepared
 * if the pair changed; when they differ, layer 0 runs yq_forward_convolutional_layer_quant_per_image_gpu with per-image tables
 * and the rest of the plan is unchanged. */
YQ_API int yq_network_predict_f32(yq_network *net, const float *in_f32_nchw_host, float *out_f32_host);
/* test_detector's whole input path (examples/detector.c:903-904, :915-922): host float CHW images of one size
 * [batch][c][ih][iw] (what load_image_color returns) -> H2D, letterbox_image on the device, dynamic input quantiser, forward. */
YQ_API int yq_network_predict_image_f32(yq_network *net, const float *images_host, int ih, int iw, float *out_f32_host);
/* the same, pipelined two deep so H2D / forward / D2H of neighbouring batches overlap (serving loop):
 * submit enqueues the H2D of one batch and its forward and returns a slot (>= 0; < 0 on error);
 * collect waits for that slot and copies the yolo heads to out_f32.  Use pinned host memory. */
YQ_API int yq_network_submit_u8(yq_network *net, const uint8_t *in_u8_nchw_host);
YQ_API int yq_network_collect(yq_network *net, int slot, float *out_f32_host);
YQ_API size_t yq_network_output_floats(const yq_network *net);
YQ_API int yq_network_synchronize(yq_network *net);
YQ_API void *yq_network_stream(yq_network *net);
/* capture the forward into a CUDA graph and replay it on later calls (0 disables) */
YQ_API int yq_network_use_graph(yq_network *net, int enable);
/* one un-graphed forward with a CUDA event after every layer on the network's stream; layer_ms has
 * num_layers+1 entries: [0] = input layout transform, [1+i] = layer i (0 for aliased routes). Synchronous. */
YQ_API int yq_network_profile_forward(yq_network *net, const uint8_t *in_u8_nchw, float *layer_ms);
/* number of kernel launches one forward issues (for bench.py's gpu_launches) */
YQ_API int yq_network_launches_per_forward(const yq_network *net);
/* ... and how many of them belong to layer i in the current plan (0: fused into a neighbour's launch, or an aliased route);
 * the input layout transform, when the plan has one, is the launch in front of layer 0's */
YQ_API int yq_network_layer_launches(const yq_network *net, int i);

/* ---- data-parallel inference over the GPUs of one box, C level ----------------------------------------------
 * The path shards by image: a full replica per GPU, no per-step collective.  The reference's multi-GPU code is one host thread
 * per device that calls cuda_set_device first (src/network.c:930-937, :1164-1194) and averages weights through the host; here
 * replica 0 parses, prepares and packs, the packed filter images travel to the other devices in ONE ncclBroadcast (libnccl is
 * bound at run time with dlopen; without it this entry fails with a message, everything else works), and one host thread per
 * remaining device builds its replica from device-to-device copies out of that blob.
 * yq_dp_network_predict_u8: in_host [n_devices][batch_per_device][c][h][w] uint8, out_host [n_devices][yq_network_output_floats];
 * image block i runs on devices[i]; all devices' copies and forwards are enqueued before the first result is awaited. */
typedef struct yq_dp_network yq_dp_network;
YQ_API yq_dp_network *yq_dp_load_network(const char *cfg, const char *weights, int batch_per_device, const int *devices, int n_devices);
YQ_API void yq_dp_free_network(yq_dp_network *dp);
YQ_API int yq_dp_num_devices(const yq_dp_network *dp);
YQ_API yq_network *yq_dp_replica(yq_dp_network *dp, int i);
YQ_API size_t yq_dp_arena_bytes(const yq_dp_network *dp);          /* bytes the one broadcast carried */
YQ_API int yq_dp_images_from_arena(const yq_dp_network *dp);      /* filter images replicas 1.. took from it (device-to-device) */
YQ_API int yq_dp_network_predict_u8(yq_dp_network *dp, const uint8_t *in_u8_nchw_host, float *out_f32_host);
/* the same, pipelined two deep like yq_network_submit_u8 / yq_network_collect (use pinned host memory: yq_host_alloc) */
YQ_API int yq_dp_network_submit_u8(yq_dp_network *dp, const uint8_t *in_u8_nchw_host);
YQ_API int yq_dp_network_collect(yq_dp_network *dp, int slot, float *out_f32_host);

/* ---- "next" row 8f-4: packed-weight arena --------------------------------------------------------------------
 * The reference repacks nothing because its GEMM reads `weights_uint8` as parsed (parser.c:1124-1159); the kernels here
 * read kernel-layout filter images (OHWI rows padded to the channel stride, Toeplitz / even-odd tiles, ...) that the layer
 * constructors build from the OIHW stream.  The arena keeps those images under a content key (layout version, layer shape,
 * zero points, the u8 weights themselves), so a later load of the same `.weights` uploads them without repacking:
 *   yq_pack_arena_load(path)  before yq_load_network / yq_make_convolutional_layer_quant: entries read (0 for a file of
 *                             another layout version), -1 on error;
 *   yq_pack_arena_save(path)  afterwards, when yq_pack_arena_stats reports dirty: entries written;
 *   yq_pack_arena_clear()     drops the in-memory arena and its counters (and disables collecting);
 *   yq_pack_arena_enable(1)   starts collecting freshly built images without a file to load (a successful load enables too).
 * A stale or foreign file misses (the weights are part of the key); every entry carries a checksum of its data and a damaged
 * entry is dropped at load (a miss).  It is a cache, not an authenticated container (see yq_pack.cu). */
YQ_API int yq_pack_arena_enable(int enable);
YQ_API int yq_pack_arena_load(const char *path);
YQ_API int yq_pack_arena_save(const char *path);
YQ_API int yq_pack_arena_clear(void);
YQ_API int yq_pack_arena_stats(int *entries, int *hits, int *misses, int *dirty);

/* ---- "next" row 8f-2: box decode + NMS on the device -------------------------------------------------------
 * get_network_boxes (src/network.c:635-640: get_yolo_detections + correct_yolo_boxes, src/yolo_layer.c:247-343) followed by
 * do_nms_sort (src/box.c:58-89) when nms_thresh > 0, for every image of the last forward.
 * counts_host[batch]; dets_host[batch][yq_network_box_capacity][5 + classes] = x, y, w, h, objectness, prob[classes],
 * candidates in the reference's pre-NMS order (yolo layers in network order, cell, anchor); NMS zeroes suppressed probs. */
YQ_API int yq_network_box_capacity(const yq_network *net);
YQ_API int yq_network_classes(const yq_network *net);
YQ_API int yq_network_get_boxes(yq_network *net, int w, int h, float thresh, float nms_thresh, int relative,
                                int *counts_host, float *dets_host);

/* pull one layer's output to the host in the REFERENCE's layout (CHW per image):
 * what = 0: output_uint8_final [batch][out_c][out_h][out_w] u8
 *        1: output_int32       [batch][out_c][out_h][out_w] i32  (conv, needs yq_network_set_debug(net,1))
 *        2: output (float)     [batch][out_c][out_h][out_w] f32  (quant_stop convs and yolo layers)   */
YQ_API int yq_network_pull_layer(yq_network *net, int layer, int what, void *host_out, size_t bytes);
/* device pointer of a yolo head's float output (NCHW), for device-resident consumers */
YQ_API const float *yq_network_layer_output_f32_device(const yq_network *net, int layer);
/* prepared per-channel parameters of conv layer `layer` (host copies, n entries each; any pointer may be NULL) */
YQ_API int yq_network_conv_params(const yq_network *net, int layer, int32_t *M0, int *M0_right_shift,
                                  double *M_value, double *M0_right_shift_value, int32_t *biases_int32);

#ifdef __cplusplus
}
#endif
#endif /* YQ_B200_H */
