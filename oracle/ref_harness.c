/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A small driver of OUR OWN that is compiled against, and linked with, the UNMODIFIED
 * reference sources where they lie under /root/reference (see oracle/Makefile, target `ref`).
 * It replays what the reference's `test_detector` does around the hot path
 * (examples/detector.c:885-924: load_network, set_batch_network(net,1), net->input = X,
 * quantization_weights_and_activations(net) exactly once, network_predict) and then writes the
 * reference's own per-layer state to disk so the oracle restatement (oracle/yq_oracle.c) and the
 * CUDA path can be pinned against it.
 *
 * Nothing in the product links or executes this file.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may run the binary.
 *
 * Usage:
 *   ref_harness net   <cfg> <weights> <input.f32> <outdir>   whole net via network_predict, dump every layer
 *   ref_harness layer <cfg> <weights> <input.f32> <outdir>   1-conv cfg: call l.forward directly
 *                                                            (forward_network reads layers[i+1] out of bounds on a
 *                                                             net that does not end in [yolo], network.c:247)
 *   ref_harness time  <cfg> <weights> <input.f32> <iters>    time network_predict, print seconds per call on stderr
 *   ref_harness letterbox <in.f32> <c> <ih> <iw> <w> <h> <out.f32>   letterbox_image alone (src/image.c:812-831)
 *
 * <input.f32> is a raw little-endian float32 CHW image of net->c * net->h * net->w values.
 */
#include "darknet.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <sys/stat.h>

static void die(const char *msg) { fprintf(stderr, "ref_harness: %s\n", msg); exit(2); }

static float *read_f32(const char *path, size_t n)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) die("cannot open input");
    float *x = calloc(n, sizeof(float));
    if (fread(x, sizeof(float), n, fp) != n) die("short input file");
    fclose(fp);
    return x;
}

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
    switch (t) {
    case CONVOLUTIONAL: return "conv";
    case MAXPOOL: return "maxpool";
    case ROUTE: return "route";
    case UPSAMPLE: return "upsample";
    case YOLO: return "yolo";
    default: return "other";
    }
}

static void dump_layer(const char *dir, FILE *man, int i, layer *l)
{
    size_t outs = (size_t)l->out_c * l->out_h * l->out_w;
    fprintf(man, "layer %d type %s c %d h %d w %d n %d size %d stride %d pad %d out_c %d out_h %d out_w %d "
                 "activation %d quant %d quant_stop %d bn %d",
            i, tname(l->type), l->c, l->h, l->w, l->n, l->size, l->stride, l->pad, l->out_c, l->out_h, l->out_w,
            (int)l->activation, l->layer_quant_flag, l->quant_stop_flag, l->batch_normalize);
    if (l->type == CONVOLUTIONAL && l->layer_quant_flag) {
        fprintf(man, " s_in %.9g zp_in %d s_out %.9g zp_out %d",
                l->input_data_uint8_scales[0], (int)l->input_data_uint8_zero_point[0],
                l->activ_data_uint8_scales[0], (int)l->activ_data_uint8_zero_point[0]);
        dump(dir, i, "output_int32", l->output_int32, outs * sizeof(int32_t));
        dump(dir, i, "output_uint8", l->output_uint8_final, outs);
        dump(dir, i, "M0", l->M0, l->n * sizeof(int32_t));
        dump(dir, i, "M0_right_shift", l->M0_right_shift, l->n * sizeof(int));
        dump(dir, i, "M_value", l->M_value, l->n * sizeof(double));
        dump(dir, i, "M0_right_shift_value", l->M0_right_shift_value, l->n * sizeof(double));
        dump(dir, i, "biases_int32", l->biases_int32, l->n * sizeof(int32_t));
        dump(dir, i, "weight_zero_point", l->weight_data_uint8_zero_point, l->n);
        dump(dir, i, "weight_scales", l->weight_data_uint8_scales, l->n * sizeof(float));
        dump(dir, i, "biases_folded", l->biases, l->n * sizeof(float));
        if (l->quant_stop_flag) dump(dir, i, "output_f32", l->output, outs * sizeof(float));
    } else if (l->type == MAXPOOL || l->type == ROUTE || l->type == UPSAMPLE) {
        fprintf(man, " s_out %.9g zp_out %d", l->activ_data_uint8_scales[0], (int)l->activ_data_uint8_zero_point[0]);
        dump(dir, i, "output_uint8", l->output_uint8_final, outs);
    } else if (l->type == YOLO) {
        dump(dir, i, "output_f32", l->output, (size_t)l->outputs * sizeof(float));
    }
    fprintf(man, "\n");
}

int main(int argc, char **argv)
{
 /* ref_harness letterbox <in.f32> <c> <ih> <iw> <w> <h> <out.f32> : the reference's letterbox_image (src/image.c:812-831) alone */
    if (argc == 9 && !strcmp(argv[1], "letterbox")) {
        int c = atoi(argv[3]), ih = atoi(argv[4]), iw = atoi(argv[5]), w = atoi(argv[6]), h = atoi(argv[7]);
        image im = make_image(iw, ih, c);
        float *src = read_f32(argv[2], (size_t)c * ih * iw);
        memcpy(im.data, src, sizeof(float) * (size_t)c * ih * iw);
        image boxed = letterbox_image(im, w, h);
        FILE *fp = fopen(argv[8], "wb");
        if (!fp || fwrite(boxed.data, sizeof(float), (size_t)c * h * w, fp) != (size_t)c * h * w) die("cannot write the letterboxed image");
        fclose(fp);
        return 0;
    }
    if (argc < 6) die("usage: ref_harness net|layer|time <cfg> <weights> <input.f32> <outdir|iters>");
    const char *mode = argv[1];
    network *net = load_network(argv[2], argv[3], 0);
    set_batch_network(net, 1);
    size_t nin = (size_t)net->c * net->h * net->w;
    float *X = read_f32(argv[4], nin);
    net->input = X;
    /* one-time host preparation; NOT idempotent (blas.c:285-286,309) -> exactly once */
    quantization_weights_and_activations(net);

    if (!strcmp(mode, "time")) {
        int iters = atoi(argv[5]);
        for (int it = 0; it < iters; ++it) {
            double t0 = what_time_is_it_now();
            network_predict(net, X);
            double t1 = what_time_is_it_now();
            fprintf(stderr, "REF_TIME %d %.6f\n", it, t1 - t0);
        }
        return 0;
    }

    const char *dir = argv[5];
    mkdir(dir, 0777);
    char path[1024];
    snprintf(path, sizeof path, "%s/manifest.txt", dir);
    FILE *man = fopen(path, "w");
    if (!man) die("cannot open manifest");
    fprintf(man, "net n %d c %d h %d w %d\n", net->n, net->c, net->h, net->w);

    if (!strcmp(mode, "net")) {
        double t0 = what_time_is_it_now();
        network_predict(net, X);
        fprintf(stderr, "REF_TIME 0 %.6f\n", what_time_is_it_now() - t0);
    } else if (!strcmp(mode, "layer")) {
        net->train = 0;
        network nn = *net;
        nn.input = X;
        layer l = net->layers[0];
        l.forward(l, nn);
    } else {
        die("unknown mode");
    }
    if (!strcmp(mode, "net")) {
        /* get_network_boxes + do_nms_sort exactly as test_detector calls them (examples/detector.c:926-930, nms = .45);
           relative = 1, image size = network size (the synthetic image is already 416x416, no letterbox offset) */
        int nboxes = 0;
        layer last = net->layers[net->n - 1];
        detection *dets = get_network_boxes(net, net->w, net->h, 0.5f, 0.5f, 0, 1, &nboxes);
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
        free(flat);
        free_detections(dets, nboxes);
    }
    dump(dir, 0, "input_uint8", net->input_uint8, nin);
    int nl = !strcmp(mode, "layer") ? 1 : net->n;
    for (int i = 0; i < nl; ++i) dump_layer(dir, man, i, &net->layers[i]);
    fclose(man);
    return 0;
}
