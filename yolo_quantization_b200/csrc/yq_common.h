// yq_common.h -- shared host-side declarations of libyq_b200.so (error convention, layer object).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/yq_b200.h"

namespace yq {

// The reference's convention is check_error(): print + assert(0) + exit(-1) (src/cuda.c:27-49).
// Here: record the message, return non-zero; yq_set_abort_on_error(1) restores print-and-abort.
int fail(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
void clear_error();

#define YQ_CUDA(call)                                                                                       \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess)                                                                             \
            return yq::fail("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__,        \
                            cudaGetErrorString(e__));                                                       \
    } while (0)

#define YQ_CHECK_LAUNCH() YQ_CUDA(cudaPeekAtLastError())

// Per-DEVICE facts and settings.  Function attributes (cudaFuncAttributeMaxDynamicSharedMemorySize) belong to a device's
// context and the ABI takes a device argument (yq_load_network(..., device), yq_set_device), so nothing of this kind may be
// cached per process: these helpers key their caches by the CURRENT device and take a mutex (two host threads may drive two
// network instances).
int device_index();                 // cudaGetDevice, -1 on error
int device_sm_count();              // multiprocessors of the current device (0 on error)
int device_smem_optin();            // cudaDevAttrMaxSharedMemoryPerBlockOptin of the current device
int device_smem_per_sm();           // cudaDevAttrMaxSharedMemoryPerMultiprocessor of the current device
// make sure `kern` may be launched with `bytes` of dynamic shared memory on the current device (0 on success)
int ensure_dynamic_smem(const void *kern, int bytes);
// per-(kernel, device) integer memo for derived launch parameters (CTAs per SM, ...): returns false when not yet stored
bool memo_get(const void *kern, int *value);
void memo_put(const void *kern, int value);

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int channel_stride(int c) { return c <= 4 ? 4 : round_up(c, 16); }

}  // namespace yq

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): every kernel of the forward is launched with the programmatic-stream-serialization
// attribute, runs its own prologue (barrier init, TMEM allocation, constant filter tiles into shared memory) while the
// previous kernel drains, and only then executes griddepcontrol.wait -- which returns once the previous grid has completed
// and its writes are visible -- before it touches any activation tensor.  No global memory is written before the wait.
// YQ_PDL=0 launches plainly.
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void yq_pdl_wait_then_release()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // my own dependents may start their prologues
}
namespace yq {
inline bool pdl_enabled()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("YQ_PDL");
        v = e ? (atoi(e) != 0) : 1;
    }
    return v != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
}  // namespace yq
#endif

// One prepared quantized conv layer on the device.
struct yq_conv_layer {
    int h, w, c, cs_in;
    int n, cs_out, size, stride, pad, out_h, out_w;
    int activation, quant_stop_flag, zp_in, zp_out, saturate;
    float s_out;
    int fused_mult;          // 1: rshift values are exact powers of two -> q = trunc(RN(x * (M_value*rshift)))
    int kernel;              // resolved flavour: 0 SIMT, 1 tcgen05
    int kernel_req;          // requested: -1 auto
    // SIMT packing: [n_pad][k_pad] bytes, K order (ky, kx, ci < cs_in), zero padded
    int n_pad, k_pad;
    uint8_t *w_simt = nullptr;
    // per-channel parameters, n_pad entries (pad rows: zero weights, bias 0, multiplier 0)
    int32_t *bias = nullptr;
    int32_t *zw = nullptr;
    double *mcomb = nullptr;  // M_value * rshift
    double *mval = nullptr;
    double *rsh = nullptr;
    void *chanq = nullptr;    // int4 {bias, zw, 2*M0, shift} per channel: integer-form epilogue (tcgen05 flavours)
    int int_form = 0;         // 1 when every (M_value, rshift) pair is M0*2^-31, 2^-s with integer M0 < 2^31, 0 <= s <= 31
    // tcgen05 packing lives behind this pointer (yq_conv_tc.cu)
    void *tc = nullptr;
    void *tc_small = nullptr;   // small-c tcgen05 flavour state (yq_conv_tc_small.cu)
    void *tc_rows = nullptr;    // halo-input conv + pool flavour state (yq_conv_tc_rows.cu), or nullptr
    void *tc_flat = nullptr;    // flat-strip patch flavour state (yq_conv_tc_flat.cu), or nullptr
    void *tc_flat2 = nullptr;   // its persistent two-tiles-per-weight-stage form (yq_conv_tc_flat2.cu), or nullptr
    void *tc_flat2x = nullptr;  // the same on CTA pairs, tcgen05 cta_group::2 (yq_conv_tc_flat2x.cu), or nullptr
    void *tc_pw = nullptr;      // pointwise streaming form: narrow 1x1 layers and detection heads (yq_conv_tc_pw.cu), or nullptr
    void *tc_pwt = nullptr;     // the same kernel in patch mode: narrow 3x3 stride-2 layers between halo-padded tensors, or nullptr
    std::vector<uint8_t> host_w;  // OIHW copy kept for repacking
    uint64_t pack_key = 0;        // content key of this layer's filter images in the packed-weight arena (yq_pack.cu)
    std::vector<uint8_t> host_zw;
    std::vector<int32_t> host_chanq;   // 4 ints per channel {bias, zw, 2*M0, shift} (copy of chanq)
    std::vector<double> host_mcomb;
};

// implemented in yq_detect.cu: yolo box decode (+ NMS when nms_thresh > 0) for a whole batch
int yq_detect_run(const float *const *head_pred, const int *lw, const int *lh, const int *n_anchors, const float *const *bias_w,
                  const float *const *bias_h, int n_heads, int classes, int batch, int cap, int netw, int neth, int w, int h, int relative,
                  float thresh, float nms_thresh, float *dets_dev, int *counts_dev, cudaStream_t stream);

// implemented in yq_conv_tc_small.cu (threads build the im2col rows; c <= 32)
int yq_tc_small_supported(const yq_conv_layer *l);
// packed-weight arena (yq_pack.cu): fetch a filter image built earlier for the same weights, or keep a freshly built one
namespace yq {
uint64_t pack_layer_key(const yq_conv_layer *l);
bool pack_fetch(const yq_conv_layer *l, const char *tag, std::vector<uint8_t> &out);
void pack_put(const yq_conv_layer *l, const char *tag, const std::vector<uint8_t> &img);
// data-parallel replicas (yq_dp.cu): the arena as one blob + index; a replica thread that was handed the blob on ITS device
// (one ncclBroadcast) takes its filter images from there by device-to-device copy -- no host packing, no host-to-device upload
struct PackIndexEntry {
    uint64_t key;
    std::string tag;
    size_t offset, bytes;
};
void pack_serialize(std::vector<uint8_t> &blob, std::vector<PackIndexEntry> &index);
void pack_set_device_arena(const uint8_t *blob_dev, const std::vector<PackIndexEntry> *index);   // per host thread; nullptr switches it off
int pack_device_arena_hits();
bool pack_fetch_device(const yq_conv_layer *l, const char *tag, size_t bytes, void **dev_out);
}  // namespace yq
int yq_tc_small_prepare(yq_conv_layer *l, void **state);
void yq_tc_small_free(void *state);
// out_geom (null = plain): geometry of the conv output tensor; the input and the pooled output are always plain
int yq_tc_small_forward(yq_conv_layer *l, void *state, const uint8_t *in_u8, uint8_t *out_u8, uint8_t *out_pool, float *out_f32,
                        int32_t *out_acc, int batch, cudaStream_t stream, const yq_act_geom *out_geom = nullptr);

// implemented in yq_conv_tc_rows.cu (3x3/1/1 + RELU6 + 2x2 pool from a halo-padded input, no im2col; c <= 32)
int yq_tc_rows_supported(const yq_conv_layer *l);
void yq_tc_rows_input_geom(const yq_conv_layer *l, yq_act_geom *g);
int yq_tc_rows_prepare(yq_conv_layer *l, void **state);
void yq_tc_rows_free(void *state);
int yq_tc_rows_two_blocks(const void *state);
int yq_tc_rows_launches(const void *state);      // kernel launches of one forward (c = 64: one per 64 output channels)
int yq_tc_rows_forward(yq_conv_layer *l, void *state, const uint8_t *in, uint8_t *out_pool, const yq_act_geom *og, int batch, cudaStream_t stream,
                       int planar = 0);
int yq_tc_rows_planar_supported(const yq_conv_layer *l);

// implemented in yq_conv_tc_l0.cu (the network's first layer from its CHW planes: c = 3, n = 16, RELU6 + 2x2 pool, dense Toeplitz MMAs;
// owned by the rows state, which hands it the planar launches)
int yq_tc_l0_supported(const yq_conv_layer *l);
int yq_tc_l0_prepare(yq_conv_layer *l, void **state);      // *state stays null when the weights do not fit the two-signed-block form
void yq_tc_l0_free(void *state);
int yq_tc_l0_forward(yq_conv_layer *l, void *state, const uint8_t *in_planes, uint8_t *out_pool, const yq_act_geom *og, int batch, cudaStream_t stream);

// implemented in yq_conv_tc_flat.cu (flat halo-padded strip, one patch per channel chunk shared by all taps; c % 64 == 0)
int yq_tc_flat_supported(const yq_conv_layer *l);
int yq_tc_flat_eligible(const yq_conv_layer *l);   // shape conditions shared by flat / flat2 / flat2x (no row-width limit)
void yq_tc_flat_geom(int h, int w, yq_act_geom *g);
int yq_tc_flat_prepare(yq_conv_layer *l, void **state);
void yq_tc_flat_free(void *state);
int yq_tc_flat_forward(yq_conv_layer *l, void *state, const uint8_t *in_flat, uint8_t *out_flat, int halo_fill, float *out_f32, float *out_yolo,
                       int yolo_classes, int32_t *out_acc, int batch, cudaStream_t stream);

// implemented in yq_conv_tc_flat2.cu (same tensors as the flat flavour; persistent, 256 positions per weight stage; n % 128 == 0)
int yq_tc_flat2_supported(const yq_conv_layer *l);
int yq_tc_flat2_prepare(yq_conv_layer *l, void **state);
void yq_tc_flat2_free(void *state);
// the FOLLOWING quantized shortcut fused into a flat2 / flat2x launch: `resid` = the shortcut's `from` tensor in the convolution's own
// (flat) output geometry and channel stride; C0 = 2^15 + (zp_out << 16) - zp_a * Ka - zp_b * Kb.  The launch stores the shortcut's output.
struct yq_fused_shortcut {
    const uint8_t *resid;
    int Ka, Kb, C0;
};
// plain = 1 (1x1 layers only): in / out are plain [B][H][W][C] tensors -- a 1x1 convolution needs no halo
int yq_tc_flat2_forward(yq_conv_layer *l, void *state, const uint8_t *in_flat, uint8_t *out_flat, int halo_fill, int32_t *out_acc, int batch,
                        cudaStream_t stream, int plain = 0, const yq_fused_shortcut *sc = nullptr);

// implemented in yq_conv_tc_flat2x.cu (flat2 on CTA pairs: cta_group::2 MMAs, each SM holds half of every weight stage)
int yq_tc_flat2x_supported(const yq_conv_layer *l);
int yq_tc_flat2x_prepare(yq_conv_layer *l, void **state);
void yq_tc_flat2x_free(void *state);
// in_first != null: the input is the channel concatenation [in_first (c_first channels) | in_flat] of two flat tensors of the same geometry
// (a route that is never materialised); c_first a multiple of the layer's channel chunk
int yq_tc_flat2x_forward(yq_conv_layer *l, void *state, const uint8_t *in_flat, uint8_t *out_flat, int halo_fill, int32_t *out_acc, int batch,
                         cudaStream_t stream, const yq_fused_shortcut *sc = nullptr, const uint8_t *in_first = nullptr, int c_first = 0);
int yq_tc_flat2x_chunk(const void *state);     // channels per patch chunk (64 or 128)

// implemented in yq_conv_tc_pw.cu (1x1 layers with n <= 255 whose filter bank fits shared memory: persistent streaming GEMM; flat or
// plain strips; out_yolo != null: the layer is a detection head and the following yolo layer's tensor is written as well)
int yq_tc_pw_supported(const yq_conv_layer *l);
int yq_tc_pw_prepare(yq_conv_layer *l, void **state);
int yq_tc_pw_head_supported(const yq_conv_layer *l);
void yq_tc_pw_free(void *state);
// out_up2 != null (flat tensors only): the FOLLOWING stride-2 upsample fused -- the launch writes the (2H x 2W) flat tensor out_up2 (same
// channel stride, interior only) INSTEAD of the layer's own tensor (out_u8 only names the tensor-map cache entry then)
int yq_tc_pw_forward(yq_conv_layer *l, void *state, const uint8_t *in, uint8_t *out_u8, int halo_fill, float *out_yolo, int yolo_classes, int batch,
                     cudaStream_t stream, int plain, uint8_t *out_up2 = nullptr, const uint8_t *in_first = nullptr, int c_first = 0);
int yq_tc_pw_chunk(const void *state);         // channels per ring stage (64 or 128)
// patch mode of the same kernel: 3x3 stride-2 layers with n <= 255 whose filter bank [n + 1][9 c] fits shared memory; the input is
// halo-padded (pad >= 1) with the halo holding zp_in, the output tensor may have any geometry (its interior is written)
int yq_tc_pwt_supported(const yq_conv_layer *l);
int yq_tc_pwt_prepare(yq_conv_layer *l, void **state);
int yq_tc_pwt_forward(yq_conv_layer *l, void *state, const uint8_t *in, const yq_act_geom *in_geom, uint8_t *out_u8, const yq_act_geom *out_geom, int batch,
                      cudaStream_t stream);

// implemented in yq_conv_tc.cu (TMA-fed; c % 64 == 0)
int yq_tc_supported(const yq_conv_layer *l);
int yq_tc_prepare(yq_conv_layer *l);
void yq_tc_free(yq_conv_layer *l);
// in_geom / out_geom (null = plain) and in_halo_fill (-1 = unknown): only the per-tap TMA flavour takes halo-padded tensors
int yq_tc_forward(yq_conv_layer *l, const uint8_t *in_u8, uint8_t *out_u8, uint8_t *out_pool, float *out_f32, int32_t *out_acc,
                  int batch, cudaStream_t stream, const yq_act_geom *in_geom = nullptr, int in_halo_fill = -1,
                  const yq_act_geom *out_geom = nullptr);
int yq_conv_plain_1x1_fast(const yq_conv_layer *l);   // 1: the plain entry runs this 1x1 layer on conv_u8_tc_flat2_kernel (plain-strip mode)
int yq_tc_out_geom_supported(const yq_conv_layer *l);   // 1: plain input, but the output may have any geometry (the small-c flavour)
int yq_tc_geom_supported(const yq_conv_layer *l);   // 1: the layer's current flavour is the per-tap TMA one (any tensor geometry)
int yq_tc_cluster_enabled();                        // YQ_TC_CLUSTER: multicast clusters in the per-tap flavour (A/B switch, off)
// 1 when this layer's current flavour can also emit the 2x2/stride-2 max-pooled tensor from its epilogue
int yq_tc_can_fuse_pool(const yq_conv_layer *l);
