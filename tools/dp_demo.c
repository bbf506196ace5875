/* tools/dp_demo.c -- the C-level data-parallel entry from plain C (no Python, no torch): images/s over all visible GPUs.
 *   gcc -O2 -Iinclude tools/dp_demo.c -Lyolo_quantization_b200 -lyq_b200 -Wl,-rpath,'$ORIGIN/../yolo_quantization_b200' -o tools/dp_demo
 *   tools/dp_demo <cfg> <weights> <batch_per_device> <steps>
 * Times `steps` batches through yq_dp_network_submit_u8 / yq_dp_network_collect (two in flight; pinned host buffers; H2D + forward + D2H
 * on every device) with the wall clock. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "yq_b200.h"

int main(int argc, char **argv)
{
    if (argc < 5) { fprintf(stderr, "usage: dp_demo <cfg> <weights> <batch_per_device> <steps>\n"); return 2; }
    int n = yq_device_count(), batch = atoi(argv[3]), steps = atoi(argv[4]), devs[64];
    if (n <= 0) { fprintf(stderr, "no CUDA device\n"); return 2; }
    if (n > 64) n = 64;
    for (int i = 0; i < n; ++i) devs[i] = i;
    yq_dp_network *dp = yq_dp_load_network(argv[1], argv[2], batch, devs, n);
    if (!dp) { fprintf(stderr, "%s\n", yq_last_error()); return 3; }
    int c, h, w;
    yq_network *r0 = yq_dp_replica(dp, 0);
    yq_network_input_dims(r0, &c, &h, &w);
    size_t in_bytes = (size_t)n * batch * c * h * w, out_floats = (size_t)n * yq_network_output_floats(r0);
    unsigned char *in = yq_host_alloc(2 * in_bytes, 0);
    float *out = yq_host_alloc(2 * out_floats * sizeof(float), 0);
    if (!in || !out) { fprintf(stderr, "%s\n", yq_last_error()); return 3; }
    for (size_t i = 0; i < 2 * in_bytes; ++i) in[i] = (unsigned char)(i * 2654435761u >> 24);
    for (int i = 0; i < 3; ++i) if (yq_dp_network_predict_u8(dp, in, out)) { fprintf(stderr, "%s\n", yq_last_error()); return 3; }
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int prev = -1;
    for (int i = 0; i < steps; ++i) {
        int slot = yq_dp_network_submit_u8(dp, in + (size_t)(i & 1) * in_bytes);
        if (slot < 0) { fprintf(stderr, "%s\n", yq_last_error()); return 3; }
        if (prev >= 0 && yq_dp_network_collect(dp, prev, out + (size_t)((i - 1) & 1) * out_floats)) { fprintf(stderr, "%s\n", yq_last_error()); return 3; }
        prev = slot;
    }
    if (prev >= 0 && yq_dp_network_collect(dp, prev, out)) { fprintf(stderr, "%s\n", yq_last_error()); return 3; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    double s = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    printf("{\"devices\": %d, \"batch_per_device\": %d, \"steps\": %d, \"images_per_s\": %.1f, \"arena_bytes_broadcast\": %zu, \"images_from_arena\": %d}\n", n, batch,
           steps, (double)n * batch * steps / s, yq_dp_arena_bytes(dp), yq_dp_images_from_arena(dp));
    yq_host_free(in);
    yq_host_free(out);
    yq_dp_free_network(dp);
    return 0;
}
