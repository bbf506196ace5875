// yq_detect.cu -- box decode + NMS on the device (the step right after the hot path).
//
// Restates, per image of the batch,
//   get_yolo_detections + get_yolo_box + correct_yolo_boxes   src/yolo_layer.c:83-91,247-273,316-343
//   fill_network_boxes (yolo layers in network order)        src/network.c:613-633
//   do_nms_sort / nms_comparator / box_iou                   src/box.c:6-19,58-89 (do_nms_sort), 152-182
// A detection is 5 + classes floats: x, y, w, h, objectness, prob[classes].
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include "yq_common.h"

namespace {

constexpr int DEC_THREADS = 256;
constexpr int MAX_HEADS = 4;
constexpr int MAX_ANCH = 8;

struct HeadDesc {
    const float *pred;          // yolo layer output, NCHW [B][n_anchors*(5+classes)][lh][lw]
    int lw, lh, n_anchors;
    float bias_w[MAX_ANCH], bias_h[MAX_ANCH];   // anchors selected by the layer's mask (l.biases[2*mask[n]], +1)
};

struct DecodeArgs {
    HeadDesc head[MAX_HEADS];
    int n_heads, classes, cap, netw, neth, w, h, relative;
    float thresh;
    float *dets;                // [B][cap][5+classes]
    int *counts;                // [B]
};

// one block per image; candidates are compacted in the reference's order (head, cell, anchor)
__global__ void __launch_bounds__(DEC_THREADS) yolo_decode_kernel(const DecodeArgs a)
{
    __shared__ int warp_sums[DEC_THREADS / 32];
    __shared__ int base;
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int stride_det = 5 + a.classes;
    float *dets = a.dets + (size_t)b * a.cap * stride_det;
    if (t == 0) base = 0;
    __syncthreads();
    // correct_yolo_boxes geometry (integer divisions as in the reference)
    int new_w, new_h;
    if (((float)a.netw / a.w) < ((float)a.neth / a.h)) { new_w = a.netw; new_h = (a.h * a.netw) / a.w; }
    else { new_h = a.neth; new_w = (a.w * a.neth) / a.h; }
    for (int hd = 0; hd < a.n_heads; ++hd) {
        const HeadDesc &H = a.head[hd];
        const int cells = H.lw * H.lh, total = cells * H.n_anchors, per = 5 + a.classes;
        const float *pred = H.pred + (size_t)b * H.n_anchors * per * cells;
        for (int p0 = 0; p0 < total; p0 += DEC_THREADS) {
            const int p = p0 + t;
            bool keep = false;
            int cell = 0, n = 0;
            float obj = 0.f;
            if (p < total) {
                cell = p / H.n_anchors;
                n = p - cell * H.n_anchors;
                obj = pred[(size_t)(n * per + 4) * cells + cell];      // entry_index(l, 0, n*w*h + i, 4)
                keep = obj > a.thresh;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) warp_sums[warp] = __popc(bal);
            __syncthreads();
            int off = base;
            for (int wq = 0; wq < warp; ++wq) off += warp_sums[wq];
            off += __popc(bal & ((1u << lane) - 1));
            if (keep) {
                const int row = cell / H.lw, col = cell - row * H.lw;
                const float *x = pred + (size_t)(n * per) * cells + cell;
                // get_yolo_box (yolo_layer.c:83-91)
                float bx = (col + x[0]) / H.lw;
                float by = (row + x[(size_t)cells]) / H.lh;
                float bw = (float)(exp((double)x[(size_t)2 * cells]) * H.bias_w[n] / a.netw);
                float bh = (float)(exp((double)x[(size_t)3 * cells]) * H.bias_h[n] / a.neth);
                // correct_yolo_boxes (yolo_layer.c:247-273)
                bx = (float)((bx - (a.netw - new_w) / 2. / a.netw) / ((float)new_w / a.netw));
                by = (float)((by - (a.neth - new_h) / 2. / a.neth) / ((float)new_h / a.neth));
                bw *= (float)a.netw / new_w;
                bh *= (float)a.neth / new_h;
                if (!a.relative) { bx *= a.w; bw *= a.w; by *= a.h; bh *= a.h; }
                float *d = dets + (size_t)off * stride_det;
                d[0] = bx; d[1] = by; d[2] = bw; d[3] = bh; d[4] = obj;
                for (int j = 0; j < a.classes; ++j) {
                    const float prob = obj * x[(size_t)(5 + j) * cells];
                    d[5 + j] = prob > a.thresh ? prob : 0.f;
                }
            }
            __syncthreads();
            if (t == 0) {
                int s = 0;
                for (int wq = 0; wq < DEC_THREADS / 32; ++wq) s += warp_sums[wq];
                base += s;
            }
            __syncthreads();
        }
    }
    if (t == 0) a.counts[b] = base;
}

__device__ __forceinline__ float overlap1d(float x1, float w1, float x2, float w2)
{
    const float l1 = x1 - w1 / 2, l2 = x2 - w2 / 2;
    const float left = l1 > l2 ? l1 : l2;
    const float r1 = x1 + w1 / 2, r2 = x2 + w2 / 2;
    const float right = r1 < r2 ? r1 : r2;
    return right - left;
}
__device__ __forceinline__ float box_iou(const float *a, const float *b)
{
    const float w = overlap1d(a[0], a[2], b[0], b[2]);
    const float h = overlap1d(a[1], a[3], b[1], b[3]);
    const float inter = (w < 0 || h < 0) ? 0.f : __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(__fmul_rn(a[2], a[3]), __fmul_rn(b[2], b[3])), inter);
    return __fdiv_rn(inter, uni);
}

// do_nms_sort for ONE image (one block), classes in sequence.  The reference qsort()s the same array once per class
// with glibc's stable merge sort, so every class starts from the order the previous class left, and -- uint8-quantised
// heads make exact probability ties common -- that carried order decides which of two tied boxes survives.  Here:
// key = prob bits << 32 | (0xffffffff - position in the current order); a descending bitonic sort of the keys is the
// same stable sort.  Then the greedy pass: every surviving box suppresses later boxes with IoU > thresh (prob[k] = 0).
// Dynamic smem: npow2 * 8 bytes of keys + 2 * npow2 * 4 bytes of order arrays.
__global__ void __launch_bounds__(256) yolo_nms_kernel(float *dets_all, const int *counts, int cap, int classes, float thresh, int npow2)
{
    extern __shared__ unsigned long long sort_buf[];
    int *order = (int *)(sort_buf + npow2);
    int *order2 = order + npow2;
    __shared__ int npos;
    const int b = blockIdx.x, t = threadIdx.x;
    const int stride_det = 5 + classes;
    float *dets = dets_all + (size_t)b * cap * stride_det;
    const int m = counts[b];
    for (int i = t; i < npow2; i += blockDim.x) order[i] = i;
    __syncthreads();
    for (int k = 0; k < classes; ++k) {
        for (int i = t; i < npow2; i += blockDim.x) {
            unsigned long long v = 0ull;
            if (i < m) {
                const float p = dets[(size_t)order[i] * stride_det + 5 + k];   // probabilities are >= 0: bit pattern orders like the value
                v = ((unsigned long long)__float_as_uint(p) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
            }
            sort_buf[i] = v;
        }
        if (t == 0) npos = 0;
        __syncthreads();
        for (int size = 2; size <= npow2; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int i = t; i < npow2; i += blockDim.x) {
                    const int j = i ^ stride;
                    if (j > i) {
                        const unsigned long long x = sort_buf[i], y = sort_buf[j];
                        const bool desc = (i & size) == 0;
                        if (desc ? (x < y) : (x > y)) { sort_buf[i] = y; sort_buf[j] = x; }
                    }
                }
                __syncthreads();
            }
        // new order (entries past m keep their padding slots) and the number of strictly positive probabilities
        for (int i = t; i < npow2; i += blockDim.x) {
            const unsigned long long v = sort_buf[i];
            order2[i] = i < m ? order[(int)(0xffffffffu - (unsigned)(v & 0xffffffffull))] : i;
            if (i < m && (v >> 32) != 0ull && (i + 1 == m || (sort_buf[i + 1] >> 32) == 0ull)) npos = i + 1;
        }
        __syncthreads();
        const int np = npos;
        for (int i = 0; i < np; ++i) {
            if ((sort_buf[i] >> 32) == 0ull) continue;         // uniform: every thread reads the same word
            const float *bi = dets + (size_t)order2[i] * stride_det;
            for (int j = i + 1 + t; j < np; j += blockDim.x) {
                const unsigned long long vj = sort_buf[j];
                if ((vj >> 32) == 0ull) continue;
                float *bj = dets + (size_t)order2[j] * stride_det;
                if (box_iou(bi, bj) > thresh) {
                    sort_buf[j] = vj & 0xffffffffull;
                    bj[5 + k] = 0.f;
                }
            }
            __syncthreads();
        }
        __syncthreads();
        int *tmp = order; order = order2; order2 = tmp;
    }
}

}  // namespace

int yq_detect_run(const float *const *head_pred, const int *lw, const int *lh, const int *n_anchors, const float *const *bias_w,
                  const float *const *bias_h, int n_heads, int classes, int batch, int cap, int netw, int neth, int w, int h, int relative,
                  float thresh, float nms_thresh, float *dets_dev, int *counts_dev, cudaStream_t stream)
{
    if (n_heads <= 0 || n_heads > MAX_HEADS) return yq::fail("detect: %d yolo heads (max %d)", n_heads, MAX_HEADS);
    DecodeArgs a;
    memset(&a, 0, sizeof a);
    for (int i = 0; i < n_heads; ++i) {
        if (n_anchors[i] > MAX_ANCH) return yq::fail("detect: %d anchors per head (max %d)", n_anchors[i], MAX_ANCH);
        a.head[i].pred = head_pred[i]; a.head[i].lw = lw[i]; a.head[i].lh = lh[i]; a.head[i].n_anchors = n_anchors[i];
        for (int n = 0; n < n_anchors[i]; ++n) { a.head[i].bias_w[n] = bias_w[i][n]; a.head[i].bias_h[n] = bias_h[i][n]; }
    }
    a.n_heads = n_heads; a.classes = classes; a.cap = cap; a.netw = netw; a.neth = neth; a.w = w; a.h = h; a.relative = relative;
    a.thresh = thresh; a.dets = dets_dev; a.counts = counts_dev;
    yolo_decode_kernel<<<batch, DEC_THREADS, 0, stream>>>(a);
    YQ_CHECK_LAUNCH();
    if (nms_thresh > 0.f) {
        int npow2 = 1;
        while (npow2 < cap) npow2 <<= 1;
        const size_t smem = (size_t)npow2 * (sizeof(unsigned long long) + 2 * sizeof(int));
        if (smem > 200 * 1024) return yq::fail("detect: %d candidate slots per image exceed the NMS sort capacity", cap);
        if (yq::ensure_dynamic_smem((const void *)yolo_nms_kernel, (int)smem)) return -1;
        yolo_nms_kernel<<<batch, 256, smem, stream>>>(dets_dev, counts_dev, cap, classes, nms_thresh, npow2);
        YQ_CHECK_LAUNCH();
    }
    return 0;
}
