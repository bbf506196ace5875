/*
 * oracle/ref_bind_harness.c -- TEST INFRASTRUCTURE ONLY: the drop-in proof.
 *
 * Compiled against the UNMODIFIED reference (its own include/darknet.h and src/ objects, see oracle/Makefile target `ref`)
 * and linked with libyq_b200.so.  Everything around the hot path is the reference's own code:
 *     load_network / parse_network_cfg / load_weights          (src/network.c:49-57, src/parser.c:682-815,1201-1305)
 *     set_batch_network, quantization_weights_and_activations   (src/network.c:415-432, src/blas.c:259-346)
 *     network_predict -> forward_network                         (src/network.c:570-581, :229-261)
 *     get_network_boxes, do_nms_sort                             (src/network.c:635-640, src/box.c:58-89)
 * Only the per-layer function-pointer slot `void (*forward)(struct layer, struct network)` (include/darknet.h:158-163;
 * the reference fills it in make_convolutional_layer src/convolutional_layer.c:258-273, make_maxpool_layer
 * src/maxpool_layer.c:56-65, make_upsample_layer src/upsample_layer.c:41-52, make_route_layer src/route_layer.c:38-49,
 * make_yolo_layer src/yolo_layer.c:48-52) is re-pointed at the stubs below, which are the INTEGRATION.md stubs made
 * to compile: each calls one yq_forward_*_gpu entry of include/yq_b200.h.  The reference's forward_network then drives
 * the B200 kernels layer by layer, through its own loop and its own uint8 hand-off (network.c:248-250).
 *
 * The CPU build of the reference has no device-pointer fields in `struct layer` (INTEGRATION.md section 1 appends them in a
 * GPU build), so the per-layer device state lives in a side table indexed by net.index, which forward_network sets before
 * every call (network.c:238).  After each launch the stub copies the layer's result back into the reference's own host
 * buffers (l.output_uint8_final, l.output_int32, l.output) -- that is what makes the reference's hand-off, its yolo decode
 * and the per-layer comparison with a pure-reference run work unchanged.
 *
 * Usage:  ref_bind_harness <cfg> <weights> <input.f32> <outdir>      (same dump format as ref_harness net)
 * Nothing in the product links or executes this file.
 */
#include "darknet.h"
#include "yq_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <sys/stat.h>

static void die(const char *msg) { fprintf(stderr, "ref_bind_harness: %s\n", msg); exit(2); }
#define YQ(call) do { if (call) { fprintf(stderr, "ref_bind_harness: %s failed: %s\n", #call, yq_last_error()); exit(3); } } while (0)

/* ---- per-layer device state (what INTEGRATION.md appends to struct layer in a GPU build) ---- */
typedef struct {
    yq_conv_layer *conv;       /* l.yq_conv */
    uint8_t *out_u8;           /* l.output_uint8_gpu : device, NHWC, yq_channel_stride(out_c) bytes per pixel */
    int32_t *out_acc;          /* device int32 accumulators (parity dump of l.output_int32) */
    float *out_f32;            /* l.output_gpu : device float CHW (quant_stop heads, yolo) */
} dev_layer;

static dev_layer *g_dev;       /* [net->n] */
static network *g_net;
static uint8_t *g_in_nchw, *g_in_nhwc, *g_scratch;   /* net.input_uint8_nchw_gpu, net.input_uint8_gpu (layer 0), pull staging */

static size_t nhwc_bytes(int b, int h, int w, int c) { return (size_t)b * h * w * yq_channel_stride(c); }

/* device NHWC -> the reference's host CHW buffer (l.output_uint8_final / l.output_int32) */
static void pull_u8(const uint8_t *dev, uint8_t *host, int b, int c, int h, int w)
{
    YQ(yq_nhwc_to_nchw_u8(dev, g_scratch, b, c, h, w, 0));
    YQ(yq_cuda_pull(host, g_scratch, (size_t)b * c * h * w, 0));
    YQ(yq_stream_synchronize(0));
}
static void pull_i32(const int32_t *dev, int32_t *host, int b, int c, int h, int w)
{
    YQ(yq_nhwc_to_nchw_i32(dev, (int32_t *)g_scratch, b, c, h, w, 0));
    YQ(yq_cuda_pull(host, g_scratch, (size_t)b * c * h * w * 4, 0));
    YQ(yq_stream_synchronize(0));
}

/* the device tensor the reference's net.input_uint8 stands for at layer i: the previous layer's output, or the network input */
static const uint8_t *input_dev(network net)
{
    int i = net.index;
    if (i == 0) {
        /* forward_network_gpu would push net->input_uint8 once (network.c:839 pushes the float input today) */
        size_t n = (size_t)net.batch * net.c * net.h * net.w;
        YQ(yq_cuda_push(g_in_nchw, net.input_uint8, n, 0));
        YQ(yq_nchw_to_nhwc_u8(g_in_nchw, g_in_nhwc, net.batch, net.c, net.h, net.w, 0));
        return g_in_nhwc;
    }
    return g_dev[i - 1].out_u8;
}

/* ---- the stubs (INTEGRATION.md sections 2 and 3) ---- */
static void forward_convolutional_layer_quant_yq(layer l, network net)
{
    dev_layer *d = &g_dev[net.index];
    YQ(yq_forward_convolutional_layer_quant_gpu(d->conv, input_dev(net), d->out_u8, l.quant_stop_flag ? d->out_f32 : 0, d->out_acc, l.batch, 0));
    pull_u8(d->out_u8, l.output_uint8_final, l.batch, l.out_c, l.out_h, l.out_w);
    pull_i32(d->out_acc, l.output_int32, l.batch, l.out_c, l.out_h, l.out_w);
    if (l.quant_stop_flag) {
        YQ(yq_cuda_pull(l.output, d->out_f32, (size_t)l.batch * l.outputs * sizeof(float), 0));
        YQ(yq_stream_synchronize(0));
    }
}
static void forward_maxpool_layer_quant_yq(layer l, network net)
{
    dev_layer *d = &g_dev[net.index];
    YQ(yq_forward_maxpool_layer_quant_gpu(input_dev(net), d->out_u8, l.batch, l.h, l.w, l.c, l.size, l.stride, l.pad, 0));
    pull_u8(d->out_u8, l.output_uint8_final, l.batch, l.out_c, l.out_h, l.out_w);
}
static void forward_upsample_layer_quant_yq(layer l, network net)
{
    dev_layer *d = &g_dev[net.index];
    YQ(yq_forward_upsample_layer_quant_gpu(input_dev(net), d->out_u8, l.batch, l.h, l.w, l.c, l.stride, 0));
    pull_u8(d->out_u8, l.output_uint8_final, l.batch, l.out_c, l.out_h, l.out_w);
}
static void forward_route_layer_quant_yq(layer l, network net)
{
    dev_layer *d = &g_dev[net.index];
    const uint8_t *ins[8];
    int cs[8];
    if (l.n > 8) die("route with more than 8 inputs");
    for (int k = 0; k < l.n; ++k) {
        ins[k] = g_dev[l.input_layers[k]].out_u8;
        cs[k] = net.layers[l.input_layers[k]].out_c;
    }
    YQ(yq_forward_route_layer_quant_gpu(ins, cs, l.n, d->out_u8, l.batch, l.out_h, l.out_w, 0));
    pull_u8(d->out_u8, l.output_uint8_final, l.batch, l.out_c, l.out_h, l.out_w);
}
static void forward_yolo_layer_yq(layer l, network net)
{
    dev_layer *d = &g_dev[net.index];
    /* net.input is the previous (quant_stop) convolution's l.output: its device copy is that layer's out_f32 */
    YQ(yq_forward_yolo_layer_gpu(g_dev[net.index - 1].out_f32, d->out_f32, l.batch, l.n, l.classes, l.h, l.w, 0));
    YQ(yq_cuda_pull(l.output, d->out_f32, (size_t)l.batch * l.outputs * sizeof(float), 0));
    YQ(yq_stream_synchronize(0));
}

/* push_convolutional_layer_quant of INTEGRATION.md: runs after quantization_weights_and_activations */
static void bind_layers(network *net)
{
    size_t scratch = 0;
    g_dev = calloc(net->n, sizeof(dev_layer));
    for (int i = 0; i < net->n; ++i) {
        layer *l = &net->layers[i];
        dev_layer *d = &g_dev[i];
        size_t outs = (size_t)l->batch * l->out_c * l->out_h * l->out_w;
        if (outs * 4 > scratch) scratch = outs * 4;
        if (l->type != YOLO) d->out_u8 = yq_cuda_malloc(nhwc_bytes(l->batch, l->out_h, l->out_w, l->out_c));
        switch (l->type) {
        case CONVOLUTIONAL: {
            if (!l->layer_quant_flag) die("convolution without quantized=1");
            yq_conv_desc c;
            memset(&c, 0, sizeof c);
            c.h = l->h; c.w = l->w; c.c = l->c;
            c.n = l->n; c.size = l->size; c.stride = l->stride; c.pad = l->pad;
            c.activation = l->activation;                       /* same enum values, darknet.h:87-89 */
            c.quant_stop_flag = l->quant_stop_flag;
            c.zp_in = l->input_data_uint8_zero_point[0];
            c.zp_out = l->activ_data_uint8_zero_point[0];
            c.s_out = l->activ_data_uint8_scales[0];
            c.weights_uint8 = l->weights_uint8;                 /* OIHW as loaded by parser.c:1143-1145 */
            c.weight_zero_point = l->weight_data_uint8_zero_point;
            c.biases_int32 = l->biases_int32;
            c.M_value = l->M_value;
            c.M0_right_shift_value = l->M0_right_shift_value;
            c.saturate = 0;
            d->conv = yq_make_convolutional_layer_quant(&c);
            if (!d->conv) { fprintf(stderr, "layer %d: %s\n", i, yq_last_error()); exit(3); }
            d->out_acc = yq_cuda_malloc(nhwc_bytes(l->batch, l->out_h, l->out_w, l->out_c) * 4);
            if (l->quant_stop_flag) d->out_f32 = yq_cuda_malloc(outs * sizeof(float));
            l->forward = forward_convolutional_layer_quant_yq;
            break;
        }
        case MAXPOOL: l->forward = forward_maxpool_layer_quant_yq; break;
        case UPSAMPLE: l->forward = forward_upsample_layer_quant_yq; break;
        case ROUTE: l->forward = forward_route_layer_quant_yq; break;
        case YOLO:
            d->out_f32 = yq_cuda_malloc(outs * sizeof(float));
            l->forward = forward_yolo_layer_yq;
            break;
        default: die("layer type outside the quantized inference path");
        }
    }
    size_t nin = (size_t)net->batch * net->c * net->h * net->w;
    g_in_nchw = yq_cuda_malloc(nin);
    g_in_nhwc = yq_cuda_malloc(nhwc_bytes(net->batch, net->h, net->w, net->c));
    if (nin * 4 > scratch) scratch = nin * 4;
    g_scratch = yq_cuda_malloc(scratch);
    if (!g_in_nchw || !g_in_nhwc || !g_scratch) die("device allocation failed");
}

/* ---- dump, same files as ref_harness.c so one reader serves both ---- */
static void dump(const char *dir, int li, const char *name, const void *p, size_t bytes)
{
    char path[1024];
    snprintf(path, sizeof path, "%s/L%02d_%s.bin", dir, li, name);
    FILE *fp = fopen(path, "wb");
    if (!fp) die("cannot open dump file");
    if (bytes && fwrite(p, 1, bytes, fp) != bytes) die("short write");
    fclose(fp);
}
static const char *tname(LAYER_TYPE t)
{
    return t == CONVOLUTIONAL ? "conv" : t == MAXPOOL ? "maxpool" : t == ROUTE ? "route" : t == UPSAMPLE ? "upsample" : t == YOLO ? "yolo" : "other";
}

int main(int argc, char **argv)
{
    if (argc < 5) die("usage: ref_bind_harness <cfg> <weights> <input.f32> <outdir>");
    if (yq_device_count() <= 0) die("no CUDA device");
    network *net = load_network(argv[1], argv[2], 0);                 /* the reference's own cfg / weights loader */
    set_batch_network(net, 1);
    size_t nin = (size_t)net->c * net->h * net->w;
    float *X = calloc(nin, sizeof(float));
    FILE *fp = fopen(argv[3], "rb");
    if (!fp || fread(X, sizeof(float), nin, fp) != nin) die("cannot read input");
    fclose(fp);
    net->input = X;
    quantization_weights_and_activations(net);                        /* the reference's own host prep, exactly once */
    g_net = net;
    bind_layers(net);                                                 /* the only change: l.forward slots */
    network_predict(net, X);                                          /* the reference's forward_network drives the stubs */

    const char *dir = argv[4];
    mkdir(dir, 0777);
    char path[1024];
    snprintf(path, sizeof path, "%s/manifest.txt", dir);
    FILE *man = fopen(path, "w");
    if (!man) die("cannot open manifest");
    fprintf(man, "net n %d c %d h %d w %d\n", net->n, net->c, net->h, net->w);
    int nboxes = 0;
    layer last = net->layers[net->n - 1];
    detection *dets = get_network_boxes(net, net->w, net->h, 0.5f, 0.5f, 0, 1, &nboxes);   /* host decode on the pulled heads */
    size_t per = 5 + last.classes;
    float *flat = calloc((size_t)nboxes * per + 1, sizeof(float));
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) do_nms_sort(dets, nboxes, last.classes, 0.45f);
        for (int i = 0; i < nboxes; ++i) {
            float *d = flat + (size_t)i * per;
            d[0] = dets[i].bbox.x; d[1] = dets[i].bbox.y; d[2] = dets[i].bbox.w; d[3] = dets[i].bbox.h;
            d[4] = dets[i].objectness;
            for (int j = 0; j < last.classes; ++j) d[5 + j] = dets[i].prob[j];
        }
        dump(dir, 99, pass ? "boxes_post_nms" : "boxes_pre_nms", flat, (size_t)nboxes * per * sizeof(float));
    }
    fprintf(man, "boxes n %d classes %d\n", nboxes, last.classes);
    dump(dir, 0, "input_uint8", net->input_uint8, nin);
    for (int i = 0; i < net->n; ++i) {
        layer *l = &net->layers[i];
        size_t outs = (size_t)l->out_c * l->out_h * l->out_w;
        fprintf(man, "layer %d type %s c %d h %d w %d n %d size %d stride %d pad %d out_c %d out_h %d out_w %d activation %d quant %d quant_stop %d bn %d",
                i, tname(l->type), l->c, l->h, l->w, l->n, l->size, l->stride, l->pad, l->out_c, l->out_h, l->out_w, (int)l->activation,
                l->layer_quant_flag, l->quant_stop_flag, l->batch_normalize);
        if (l->type == CONVOLUTIONAL) {
            dump(dir, i, "output_int32", l->output_int32, outs * sizeof(int32_t));
            dump(dir, i, "output_uint8", l->output_uint8_final, outs);
            if (l->quant_stop_flag) dump(dir, i, "output_f32", l->output, outs * sizeof(float));
        } else if (l->type == YOLO) {
            dump(dir, i, "output_f32", l->output, (size_t)l->outputs * sizeof(float));
        } else {
            dump(dir, i, "output_uint8", l->output_uint8_final, outs);
        }
        fprintf(man, "\n");
    }
    fclose(man);
    fprintf(stderr, "ref_bind_harness: %d layers through the reference's forward_network with yq_b200 stubs in l.forward\n", net->n);
    return 0;
}
