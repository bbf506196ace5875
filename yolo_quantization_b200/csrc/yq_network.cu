// yq_network.cu -- host-side network runtime behind the network-level C ABI (include/yq_b200.h).
//
// Mirrors, for the QUANTIZATION=1 inference path only, the reference's
//   parse_network_cfg            src/parser.c:682-815   (INI dialect: read_cfg :817-870)
//   load_weights_upto            src/parser.c:1201-1305 (per-layer readers :1124-1199)
//   quantization_weights_and_activations  src/blas.c:259-346  (one-time host prep)
//   forward_network              src/network.c:229-261  (uint8 hand-off between quantized layers)
// and drives the sm_100a kernels of yq_kernels.cu / yq_conv_tc.cu.  Written from scratch in C++;
// device tensors are uint8 NHWC, one buffer per layer, optional CUDA-graph replay.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <cctype>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "yq_common.h"

namespace {

enum LType { L_CONV = 0, L_MAXPOOL = 1, L_ROUTE = 2, L_UPSAMPLE = 3, L_YOLO = 4, L_SHORTCUT = 5 };
constexpr int MAX_ROUTE_INPUTS = 8;   // = ROUTE_MAX_INPUTS of the route kernel's by-value argument block (yq_kernels.cu)

struct Section {
    std::string type;
    std::vector<std::pair<std::string, std::string>> kv;
    const std::string *find(const char *key) const
    {
        for (auto &p : kv)
            if (p.first == key) return &p.second;
        return nullptr;
    }
    int geti(const char *key, int def) const
    {
        const std::string *v = find(key);
        return v ? atoi(v->c_str()) : def;
    }
    std::string gets(const char *key, const char *def) const
    {
        const std::string *v = find(key);
        return v ? *v : std::string(def);
    }
};

// read_cfg (parser.c:817-870): every line is stripped of ALL blanks (utils.c strip()), '[' opens a
// section, '#', ';' and empty lines are skipped, everything else is key=value.
bool read_cfg(const char *path, std::vector<Section> &out, std::string &err)
{
    std::ifstream f(path);
    if (!f) {
        err = std::string("cannot open cfg file ") + path;
        return false;
    }
    std::string line;
    int ln = 0;
    while (std::getline(f, line)) {
        ++ln;
        std::string s;
        for (char ch : line)
            if (ch != ' ' && ch != '\t' && ch != '\n' && ch != '\r') s.push_back(ch);
        if (s.empty() || s[0] == '#' || s[0] == ';') continue;
        if (s[0] == '[') {
            Section sec;
            sec.type = s;
            out.push_back(sec);
            continue;
        }
        size_t eq = s.find('=');
        if (eq == std::string::npos || out.empty()) {
            err = "config file error line " + std::to_string(ln) + ", could not parse: " + s;
            return false;
        }
        out.back().kv.emplace_back(s.substr(0, eq), s.substr(eq + 1));
    }
    return true;
}

int activation_from_string(const std::string &s)
{
    // get_activation (src/activations.c:44-64); only the four with a quantized form are accepted
    if (s == "linear") return YQ_LINEAR;
    if (s == "relu") return YQ_RELU;
    if (s == "relu6") return YQ_RELU6;
    if (s == "leaky") return YQ_LEAKY;
    if (s == "logistic") return YQ_LOGISTIC;
    return -1;
}

struct Layer {
    LType type;
    int c = 0, h = 0, w = 0, out_c = 0, out_h = 0, out_w = 0;
    int n = 0, size = 0, stride = 1, pad = 0, activation = YQ_LINEAR, bn = 0;
    int quantized = 0, quant_stop = 0, first_time = 0;
    std::vector<int> inputs;   // route: the concatenated layers; shortcut: {from}
    int Ka = 0, Kb = 0;        // shortcut (extension): round(s_prev / s_out * 2^16), round(s_from / s_out * 2^16)
    int zp_a = 0, zp_b = 0;    // shortcut: zero points of the previous layer's and the `from` layer's outputs
    bool use_geom = false;     // conv: per-tap TMA flavour between tensors of any geometry (yq_forward_convolutional_layer_quant_geom_gpu)
    bool use_outgeom = false;  // conv: plain input, halo-padded output (the small-c flavour through the same entry point)
    int classes = 0, n_anchors = 0;
    std::vector<float> anchor_w, anchor_h;   // yolo: anchors selected by mask (l.biases[2*mask[n]], [2*mask[n]+1])
    // quantisation state as in `struct layer`
    float s_in = 0, s_out = 0;
    int zp_in = 0, zp_out = 0;
    std::vector<float> biases, bn_scales, bn_mean, bn_var, s_w;
    std::vector<uint8_t> zp_w, w_u8;
    std::vector<int32_t> M0, biases_int32;
    std::vector<int> M0_right_shift;
    std::vector<double> M_value, rshift_value;
    // device
    yq_conv_layer *conv = nullptr;
    uint8_t *out_u8 = nullptr;     // NHWC; may alias another layer's buffer (single-input route)
    bool owns_u8 = false;
    float *out_f32 = nullptr;      // NCHW (quant_stop convs, yolo)
    int32_t *out_acc = nullptr;    // NHWC, debug
    size_t u8_bytes = 0, f32_count = 0;
    bool fuse_pool = false;        // conv: the next layer (maxpool 2/2) is produced by this layer's epilogue
    bool fused_away = false;       // maxpool: produced by the previous conv, no launch
    bool use_rows = false;         // conv: halo-input conv + pool flavour (yq_conv_tc_rows.cu); its input tensor is halo-padded
    bool use_flat = false;         // conv: flat-strip flavour (yq_conv_tc_flat.cu); input and output tensors are flat
    bool fuse_yolo = false;        // quant_stop conv: the following yolo layer is produced by this layer's epilogue (yolo: fused_away)
    bool fuse_shortcut = false;    // conv: the following quantized shortcut is produced by this layer's epilogue (shortcut: fused_away)
    bool fuse_up = false;          // route: inputs that are upsample layers are read through the upsample (upsample: fused_away)
    bool cat = false;              // route (two inputs): never materialised -- the flat conv behind it reads [input 0 | input 1] itself
                                   // (yq_forward_convolutional_layer_quant_flat_cat_gpu); out_u8 then holds input 0 alone when that
                                   // one has to be brought into shape first (the upsample folded into this route)
    bool cat_copy = false;         //   ... input 0 is copied (upsampled) into out_u8 by the route's own launch
    bool cat_buf = false;          //   ... input 0 lives in out_u8 (copied there, or written there by the conv in front of the upsample)
    int up2_route = -1;            // conv: the stride-2 upsample behind it is fused -- the launch writes the upsampled tensor into that route's buffer
    unsigned early_mask = 0;       // route: inputs copied on the side stream as soon as they exist (see plan_early_copies())
    cudaEvent_t ev_early = nullptr;
    std::vector<std::pair<int, int>> early_copies;   // (route layer, input index) to issue behind this layer
    bool side = false;             // runs on the side stream: a detection branch nothing later reads (see plan())
    int halo_fill = 0;             // byte kept in the halo of out_u8: the zero point its consumer convolutions pad with
    int src = -2;                  // layer whose out_u8 this layer reads (-1: the network input); routes use `inputs`
    yq_act_geom geom = {0, 0, 0};  // geometry of out_u8 (plain unless the only consumer is a rows-flavour conv)
};

}  // namespace

#ifndef YQ_DEFAULT_UPROUTE
#define YQ_DEFAULT_UPROUTE 1
#endif
#ifndef YQ_DEFAULT_EARLY_ROUTE
#define YQ_DEFAULT_EARLY_ROUTE 0
#endif
#ifndef YQ_DEFAULT_BRANCH_STREAM
#define YQ_DEFAULT_BRANCH_STREAM 1
#endif

struct yq_network {
    int device = 0;
    int batch = 1, c = 0, h = 0, w = 0;
    std::vector<Layer> layers;
    cudaStream_t stream = nullptr;
    uint8_t *in_stage_nchw = nullptr;   // staging for host-input predict
    uint8_t *in_nhwc = nullptr;
    float *in_f32 = nullptr;            // yq_network_predict_f32: float staging and per-image (scale, zero point, scratch)
    void *in_quant = nullptr;
    float *in_raw = nullptr;            // yq_network_predict_image_f32: the images before letterbox_image
    size_t in_raw_floats = 0;
    // layer 0 with PER-IMAGE input quantisation (images of one batch quantise differently): plain NHWC input, per-image tables
    uint8_t *in_plain = nullptr;
    int32_t *img_bias = nullptr;
    double *img_mcomb = nullptr;
    uint8_t *img_zp = nullptr;
    int img_pitch = 0;
    size_t in_nhwc_bytes = 0;
    yq_act_geom in_geom = {0, 0, 0};    // geometry of in_nhwc (halo-padded when layer 0 runs a halo-input flavour)
    int in_halo_fill = 0;
    uint8_t *scratch = nullptr;         // pull_layer conversions
    size_t scratch_bytes = 0;
    int keep_acc = 0;
    bool plain_l0_out = false;     // set by the first per-image-quantised forward whose plan had given layer 0 a halo-padded output (see yq_network_predict_f32)
    bool no_planar_input = getenv("YQ_NO_PLANAR") && atoi(getenv("YQ_NO_PLANAR"));   // A/B: keep the layout-transform launch in front of layer 0
    // YQ_UPROUTE: an upsample whose only reader is the route right behind it is folded into that route's launch;
    // YQ_BRANCH_STREAM: a detection branch that nothing later reads runs on a second stream beside the layers after it
    int fuse_uproute = getenv("YQ_UPROUTE") ? atoi(getenv("YQ_UPROUTE")) : YQ_DEFAULT_UPROUTE;
    int branch_stream = getenv("YQ_BRANCH_STREAM") ? atoi(getenv("YQ_BRANCH_STREAM")) : YQ_DEFAULT_BRANCH_STREAM;
    // YQ_EARLY_ROUTE (needs the side stream): an old tensor a route concatenates is copied into the route's output on the
    // side stream right behind its producer, under the convolutions in between, instead of in front of the route's consumer
    int early_route = getenv("YQ_EARLY_ROUTE") ? atoi(getenv("YQ_EARLY_ROUTE")) : YQ_DEFAULT_EARLY_ROUTE;
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int conv_kernel = -1;
    bool plan_error = false;            // a CUDA call inside plan() failed (halo fills): every later forward refuses to run
    int fusion = 1;
    int use_graph = 0;
    std::vector<std::pair<const uint8_t *, cudaGraphExec_t>> graphs;   // one captured forward per input pointer
    std::vector<cudaEvent_t> prof_events;                               // per-layer profiling (yq_network_profile_forward)
    int launches = 0;
    float *out_host_pinned = nullptr;
    size_t out_floats = 0;
    // 2-deep host pipeline (yq_network_submit_u8 / yq_network_collect): H2D, forward and D2H of
    // neighbouring batches overlap on three streams
    static const int PIPE = 2;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    uint8_t *pipe_in[PIPE] = {nullptr, nullptr};
    float *pipe_out[PIPE] = {nullptr, nullptr};
    cudaEvent_t ev_in_ready[PIPE] = {nullptr, nullptr}, ev_fwd_done[PIPE] = {nullptr, nullptr};
    int pipe_next = 0;
    bool pipe_busy[PIPE] = {false, false};
    // detections (yq_network_get_boxes)
    float *dets_dev = nullptr;
    int *counts_dev = nullptr;
    int det_cap = 0, det_classes = 0;
};

namespace {

// quant_multi_smaller_than_one_to_scale_and_shift (src/blas.c:387-418)
bool mult_to_m0_shift(float m, int32_t *m0, int *shift)
{
    if (!(m > 0.f) || !(m < 1.f)) return false;   // the reference assert()s here
    int s = 0;
    while (m < 0.5f) {
        m *= 2.0f;
        s++;
    }
    long long q = (long long)round((double)m * (double)(1ll << 31));
    if (q == (1ll << 31)) {
        q /= 2;
        s--;
    }
    if (s < 0) return false;
    *m0 = (int32_t)q;
    *shift = s;
    return true;
}

// One conv layer's share of quantization_weights_and_activations (src/blas.c:282-334).
int prepare_conv(Layer &l, int index)
{
    const int n = l.n, K = l.c * l.size * l.size;
    l.M0.assign(n, 0); l.M0_right_shift.assign(n, 0); l.M_value.assign(n, 0.0); l.rshift_value.assign(n, 0.0);
    l.biases_int32.assign(n, 0);
    if (l.s_out == 0.f) return yq::fail("layer %d: activation scale is 0 (blas.c:312 asserts)", index);
    for (int oc = 0; oc < n; ++oc) {
        float b = l.biases[oc];
        // batch_normalize_bias (blas.c:594-601): sqrt() is the double overload, .000001f promotes
        if (l.bn) b = (float)((double)b - (double)(l.bn_scales[oc] * l.bn_mean[oc]) / (sqrt((double)l.bn_var[oc]) + (double).000001f));
        if (l.s_w[oc] == 0.f) return yq::fail("layer %d: weight scale of channel %d is 0 (blas.c:293 asserts)", index, oc);
        uint32_t mult_zero_point = (uint32_t)(K * l.zp_in * (int)l.zp_w[oc]);                 // blas.c:307
        int32_t wsum = 0;
        for (int k = 0; k < K; ++k) wsum += l.w_u8[(size_t)oc * K + k];                       // :308-310
        int32_t weights_sum_int = (int32_t)(mult_zero_point - (uint32_t)(wsum * l.zp_in));    // :311
        float M = l.s_in * l.s_w[oc] / l.s_out;                                               // :313
        if (!mult_to_m0_shift(M, &l.M0[oc], &l.M0_right_shift[oc]))
            return yq::fail("layer %d channel %d: multiplier %g outside (0,1) (blas.c:391-392 asserts)", index, oc, (double)M);
        l.rshift_value[oc] = pow(2, -l.M0_right_shift[oc]);                                   // :315
        l.M_value[oc] = pow(2, -31) * l.M0[oc];                                               // :316
        l.biases_int32[oc] = (int32_t)(b / (l.s_in * l.s_w[oc]) + (float)weights_sum_int);    // :333
    }
    return 0;
}

// kind 0: SIMT everywhere; 1: tcgen05 wherever the layer shape has one (SIMT elsewhere); -1: per-layer default
void apply_kernel_choice(yq_conv_layer *cl, int kind)
{
    if (kind == 1 && yq_conv_set_kernel(cl, 1) != 0) yq_conv_set_kernel(cl, 0);
    else if (kind != 1) yq_conv_set_kernel(cl, kind);
    yq::clear_error();
}

int build_conv_device(yq_network *net, Layer &l)
{
    if (l.conv) {
        yq_free_convolutional_layer_quant(l.conv);
        l.conv = nullptr;
    }
    yq_conv_desc d;
    memset(&d, 0, sizeof d);
    d.h = l.h; d.w = l.w; d.c = l.c; d.n = l.n; d.size = l.size; d.stride = l.stride; d.pad = l.pad;
    d.activation = l.activation; d.quant_stop_flag = l.quant_stop; d.zp_in = l.zp_in; d.zp_out = l.zp_out; d.s_out = l.s_out;
    d.weights_uint8 = l.w_u8.data(); d.weight_zero_point = l.zp_w.data(); d.biases_int32 = l.biases_int32.data();
    d.M_value = l.M_value.data(); d.M0_right_shift_value = l.rshift_value.data(); d.saturate = 0;
    l.conv = yq_make_convolutional_layer_quant(&d);
    if (!l.conv) return -1;
    apply_kernel_choice(l.conv, net->conv_kernel);
    return 0;
}

template <typename T>
bool read_vec(FILE *fp, std::vector<T> &v, size_t n)
{
    v.resize(n);
    return fread(v.data(), sizeof(T), n, fp) == n;
}

void drop_graph(yq_network *net)
{
    for (auto &g : net->graphs) cudaGraphExecDestroy(g.second);
    net->graphs.clear();
}

// Decide which conv -> maxpool(2,2) pairs run as one launch, and count launches.  A pair fuses when the conv's
// flavour can pool in its epilogue and the pool is the 2x2 / stride-2 / default-padding kind.  The conv's own
// (unpooled) tensor is still written when another layer routes from it or when debug pulls are enabled.
bool conv_output_needed(const yq_network *net, int i);

bool routed_from(const yq_network *net, int i)
{
    for (const auto &l : net->layers)
        if (l.type == L_ROUTE || l.type == L_SHORTCUT)
            for (int idx : l.inputs)
                if (idx == i) return true;
    return false;
}

bool same_geom(const yq_act_geom &a, const yq_act_geom &b) { return a.pad == b.pad && a.pitch_w == b.pitch_w && a.rows_h == b.rows_h; }

// the layer whose buffer really holds layer i's output (single-input routes are aliases, route_layer.c:107-117 with n = 1)
int tensor_of(const yq_network *net, int i)
{
    while (i >= 0 && net->layers[i].type == L_ROUTE && net->layers[i].inputs.size() == 1) i = net->layers[i].inputs[0];
    return i;
}

// A route at layer i whose inputs all lie at or before layer m < i - 1 starts a new branch from old tensors: the layers
// m+1 .. i-1 in front of it (a detection head: convs + yolo) are read by nothing later, so they run on the side stream
// while the main stream goes on with layer i.  Every main-stream layer must read main-stream tensors only; if any
// does not, no layer is marked.
bool side_stream_safe(const yq_network *net);

// A route input that is at least three layers old (layer 8 under route 20 of yolov3-tiny) is copied behind its producer.
void plan_early_copies(yq_network *net)
{
    const int n = (int)net->layers.size();
    for (auto &l : net->layers) {
        l.early_mask = 0;
        l.early_copies.clear();
    }
    if (!net->branch_stream || !net->early_route || net->keep_acc || !side_stream_safe(net)) return;
    for (int i = 0; i < n; ++i) {
        Layer &r = net->layers[i];
        if (r.type != L_ROUTE || r.inputs.size() < 2 || r.side || r.cat) continue;
        for (size_t k = 0; k < r.inputs.size(); ++k) {
            const Layer &p = net->layers[r.inputs[k]];
            if (p.type == L_UPSAMPLE && p.fused_away) continue;
            const int t = tensor_of(net, r.inputs[k]);
            if (t < 0 || t >= i - 3 || net->layers[t].side) continue;
            if (!r.ev_early && cudaEventCreateWithFlags(&r.ev_early, cudaEventDisableTiming) != cudaSuccess) return;
            r.early_mask |= 1u << k;
            net->layers[t].early_copies.push_back({i, (int)k});
            if (r.early_mask != (1u << r.inputs.size()) - 1u) ++net->launches;   // (one launch more unless the route's own launch vanished)
        }
    }
}

bool side_stream_safe(const yq_network *net)
{
    // (only the multicast-cluster form of the per-tap flavour has CTAs that wait for each other; it is an A/B switch, off by default)
    if (!yq_tc_cluster_enabled()) return true;
    for (const auto &l : net->layers)
        if (l.type == L_CONV && l.conv && !l.use_rows && !l.use_flat && l.conv->kernel != 0) return false;
    return true;
}

void plan_side_branches(yq_network *net)
{
    const int n = (int)net->layers.size();
    if (!net->branch_stream) return;
    // Two launches may now be in flight at once.  The flat-strip CTA pairs are sized so that two of them never share an SM
    // (yq_conv_tc_flat2x.cu); the multicast clusters of the per-tap flavour (yq_conv_tc.cu) carry no such guarantee, so a
    // network that runs any convolution on that flavour keeps the single stream.
    if (!side_stream_safe(net)) return;
    for (int i = 1; i < n; ++i) {
        const Layer &r = net->layers[i];
        if (r.type != L_ROUTE) continue;
        int m = -1;
        for (int idx : r.inputs) m = idx > m ? idx : m;
        if (m < 0 || m >= i - 1) continue;
        bool ok = true;
        for (int j = m + 1; j < i && ok; ++j) ok = !net->layers[j].side && net->layers[j].src >= m;
        for (int j = i; j < n && ok; ++j)
            if (net->layers[j].type == L_ROUTE || net->layers[j].type == L_SHORTCUT)
                for (int idx : net->layers[j].inputs) ok = ok && !(idx > m && idx < i);
        if (ok)
            for (int j = m + 1; j < i; ++j) net->layers[j].side = true;
    }
    bool ok = true;
    for (int i = 0; i < n && ok; ++i) {
        const Layer &l = net->layers[i];
        if (l.side) continue;
        if (l.type == L_ROUTE || l.type == L_SHORTCUT)
            for (int idx : l.inputs) ok = ok && !net->layers[idx].side;
        if (l.type != L_ROUTE && l.src >= 0) ok = ok && !net->layers[l.src].side;
    }
    if (!ok)
        for (auto &l : net->layers) l.side = false;
}

// Decide the kernel flavour of every convolution, which conv -> maxpool(2,2) pairs run as one launch, and the geometry
// of every activation tensor.  Halo-input flavours constrain the tensors next to them:
//   rows conv  : input = its own padded geometry (any producer that can write one); writes the POOLED tensor, any geometry
//   flat conv  : input and output = the flat strip of its image size
//   other conv flavours (SIMT, small-c, per-tap TMA): plain tensors only
//   maxpool / upsample / route / the input transform: any geometry on either side
// A tensor touched by conflicting requirements makes the halo-input convs next to it fall back to a plain flavour.
void plan(yq_network *net)
{
    const int n = (int)net->layers.size();
    std::vector<char> rows_ok(n, 0), flat_ok(n, 0), plain1_ok(n, 0), patch_ok(n, 0);
    int prev = -1;
    for (int i = 0; i < n; ++i) {
        Layer &l = net->layers[i];
        l.src = prev;
        if (l.type != L_YOLO) prev = i;
        if (l.type != L_CONV || !l.conv || net->conv_kernel == 0) continue;
        const bool pool22 = i + 1 < n && net->layers[i + 1].type == L_MAXPOOL && net->layers[i + 1].size == 2 && net->layers[i + 1].stride == 2 &&
                            net->layers[i + 1].pad == 1;
        // YQ_NO_ROWS / YQ_NO_FLAT = 1 keep the older flavours for A/B measurements
        static const bool no_rows = getenv("YQ_NO_ROWS") && atoi(getenv("YQ_NO_ROWS")), no_flat = getenv("YQ_NO_FLAT") && atoi(getenv("YQ_NO_FLAT"));
        rows_ok[i] = !no_rows && net->fusion && pool22 && yq_conv_rows_supported(l.conv) && !conv_output_needed(net, i);
        flat_ok[i] = !no_flat && yq_conv_flat_supported(l.conv) != 0;
        // a 1x1 layer that cannot keep flat tensors still runs the persistent kernel between PLAIN ones (no halo needed at all)
        plain1_ok[i] = yq_conv_plain_1x1_fast(l.conv) != 0;
        // a narrow 3x3 stride-2 layer on the resident-bank kernel (patch mode): wants its input as a flat strip whose halo holds its zp_in
        patch_ok[i] = net->fusion && !net->keep_acc && !rows_ok[i] && !flat_ok[i] && yq_conv_patch_supported(l.conv) != 0;
        // (layer 0 behind the per-image input quantiser writes a plain tensor: once such a forward has run, its reader gives the flat input up)
        if (net->plain_l0_out && tensor_of(net, l.src) == 0) patch_ok[i] = 0;
    }
    // requirement on tensor t (index t + 1; t = -1 is the network input): 0 none, 1 plain, 2 a specific padded geometry
    struct Req { int kind = 0; yq_act_geom g = {0, 0, 0}; int fill = -1, wish = -1; bool conflict = false; };
    std::vector<Req> req;
    auto need = [&](int t, int kind, const yq_act_geom *g, int fill) {
        Req &r = req[tensor_of(net, t) + 1];
        if (kind == 2 && r.kind == 2 && !same_geom(r.g, *g)) r.conflict = true;
        if ((kind == 1 && r.kind == 2) || (kind == 2 && r.kind == 1)) r.conflict = true;
        if (kind == 2 && fill >= 0) {
            if (r.fill >= 0 && r.fill != fill) r.conflict = true;
            r.fill = fill;
        } else if (kind == 0 && fill >= 0) {
            r.wish = fill;      // a geometry-agnostic consumer's wish: used when somebody else makes the tensor halo-padded
        }
        if (kind > r.kind) r.kind = kind;
        if (kind == 2) r.g = *g;
    };
    // conv i (flat) + the quantized shortcut right behind it in one launch: nobody else may read conv i's tensor, the shortcut's
    // `from` tensor must be a flat strip of the same shape (it is whenever a flat conv of the same size reads it), and every
    // convolution that reads the shortcut's tensor must accept a flat strip (flat convs of this size, or geometry-agnostic ones)
    auto can_fuse_shortcut = [&](int i) -> bool {
        if (!net->fusion || net->keep_acc || i + 1 >= n) return false;
        const Layer &c = net->layers[i], &sc = net->layers[i + 1];
        if (sc.type != L_SHORTCUT || !flat_ok[i] || !yq_conv_flat_shortcut_supported(c.conv) || routed_from(net, i)) return false;
        const int from = tensor_of(net, sc.inputs[0]);
        if (from < 0 || net->layers[from].out_c != c.out_c) return false;
        bool from_flat = false;
        for (int j = 0; j < n; ++j) {
            const Layer &k = net->layers[j];
            if (k.type != L_CONV || !k.conv) continue;
            if (tensor_of(net, k.src) == from && flat_ok[j] && k.h == c.out_h && k.w == c.out_w) from_flat = true;
            if (tensor_of(net, k.src) == i + 1 && !(flat_ok[j] && k.h == c.out_h && k.w == c.out_w) && !yq_conv_geom_supported(k.conv)) return false;
        }
        return from_flat;
    };
    for (int pass = 0; pass < n + 2; ++pass) {
        req.assign(n + 1, Req());
        for (int i = 0; i < n; ++i) {
            Layer &l = net->layers[i];
            if (l.type != L_CONV) continue;
            yq_act_geom g;
            if (rows_ok[i]) {
                yq_conv_rows_input_geom(l.conv, &g);
                need(l.src, 2, &g, l.zp_in);            // (the pooled tensor it writes may have any geometry)
            } else if (flat_ok[i]) {
                yq_act_geom_flat(l.h, l.w, &g);
                need(l.src, 2, &g, l.zp_in);
                need(i, 2, &g, -1);
                if (can_fuse_shortcut(i)) need(i + 1, 2, &g, -1);      // the launch stores the shortcut's tensor, as a flat strip
            } else if (patch_ok[i]) {
                yq_act_geom_flat(l.h, l.w, &g);
                need(l.src, 2, &g, l.zp_in);            // (the tensor it writes may have any geometry)
            } else if (plain1_ok[i]) {
                need(l.src, 1, nullptr, -1);
                need(i, 1, nullptr, -1);
            } else if (yq_conv_geom_supported(l.conv)) {
                // the per-tap TMA flavour reads and writes through tensor maps built for whatever geometry the tensors have;
                // if its input turns out halo-padded it would like the halo to hold its zp_in (no border correction then)
                need(l.src, 0, nullptr, l.zp_in);
            } else {
                need(l.src, 1, nullptr, -1);
                const bool fuses = net->fusion && i + 1 < n && net->layers[i + 1].type == L_MAXPOOL && net->layers[i + 1].size == 2 &&
                                   net->layers[i + 1].stride == 2 && net->layers[i + 1].pad == 1 && yq_conv_can_fuse_maxpool(l.conv);
                // (the small-c flavour's threads store their pixels themselves: its conv output may take whatever geometry the
                // consumers ask for -- a stride-2 c = 32 layer in front of flat convolutions writes the flat strip directly)
                if (fuses || net->keep_acc || !yq_conv_out_geom_supported(l.conv)) need(i, 1, nullptr, -1);
                if (fuses) need(i + 1, 1, nullptr, -1);
            }
        }
        bool changed = false;
        // a plain-strip 1x1 layer next to a conflicting tensor gives way first (it falls to the per-tap flavour, which takes any geometry)
        for (int i = 0; i < n; ++i) {
            Layer &l = net->layers[i];
            if (l.type != L_CONV || rows_ok[i] || flat_ok[i] || !plain1_ok[i] || !yq_conv_geom_supported(l.conv)) continue;
            if (req[tensor_of(net, l.src) + 1].conflict || req[tensor_of(net, i) + 1].conflict) {
                plain1_ok[i] = 0;
                changed = true;
            }
        }
        for (int i = 0; i < n && !changed; ++i) {
            Layer &l = net->layers[i];
            if (l.type != L_CONV || !patch_ok[i]) continue;
            if (req[tensor_of(net, l.src) + 1].conflict) {
                patch_ok[i] = 0;
                changed = true;
            }
        }
        if (changed) continue;
        for (int i = 0; i < n; ++i) {
            Layer &l = net->layers[i];
            if (l.type != L_CONV || !(rows_ok[i] || flat_ok[i])) continue;
            const bool bad = req[tensor_of(net, l.src) + 1].conflict || (!rows_ok[i] && req[tensor_of(net, i) + 1].conflict);
            if (bad) {
                if (rows_ok[i]) rows_ok[i] = 0;     // (a rows conv that loses its input geometry may still run flat or plain)
                else flat_ok[i] = 0;
                changed = true;
            }
        }
        if (getenv("YQ_DEBUG_PLAN"))
            for (int i = 0; i < n && i < 12; ++i)
                if (net->layers[i].type == L_CONV)
                    fprintf(stderr, "yq plan pass %d layer %d: rows %d flat %d plain1 %d | src tensor %d kind %d conflict %d fill %d | own kind %d conflict %d\n", pass, i, rows_ok[i],
                            flat_ok[i], plain1_ok[i], tensor_of(net, net->layers[i].src), req[tensor_of(net, net->layers[i].src) + 1].kind,
                            (int)req[tensor_of(net, net->layers[i].src) + 1].conflict, req[tensor_of(net, net->layers[i].src) + 1].fill, req[tensor_of(net, i) + 1].kind,
                            (int)req[tensor_of(net, i) + 1].conflict);
        if (!changed) break;
    }
    // ---- apply
    for (int i = 0; i < n; ++i) {
        Layer &l = net->layers[i];
        l.fuse_pool = l.fused_away = l.use_rows = l.use_flat = l.use_geom = l.use_outgeom = l.fuse_yolo = l.fuse_shortcut = l.fuse_up = l.side = false;
        l.cat = l.cat_copy = l.cat_buf = false;
        l.up2_route = -1;
        l.geom = yq_act_geom{0, l.out_w, l.out_h};
        l.halo_fill = 0;
    }
    net->in_geom = yq_act_geom{0, net->w, net->h};
    net->in_halo_fill = 0;
    for (int t = -1; t < n; ++t) {
        const Req &r = req[t + 1];
        if (r.kind != 2 || (t >= 0 && tensor_of(net, t) != t)) continue;
        const int fill = r.fill >= 0 ? r.fill : (r.wish >= 0 ? r.wish : 0);
        uint8_t *buf = t < 0 ? net->in_nhwc : net->layers[t].out_u8;
        const int c = t < 0 ? net->c : net->layers[t].out_c;
        (t < 0 ? net->in_geom : net->layers[t].geom) = r.g;
        (t < 0 ? net->in_halo_fill : net->layers[t].halo_fill) = fill;
        // the halo (and any slack around the images) holds the consumers' input zero point (im2col.c:5-14); producers
        // only ever write the interior (flat convs also rewrite the halo of their output with the same value)
        if (cudaMemsetAsync(buf, fill, yq_act_geom_bytes(&r.g, net->batch, c), net->stream) != cudaSuccess) net->plan_error = true;
    }
    for (int i = 0; i < n; ++i)   // aliases share their tensor's geometry
        if (tensor_of(net, i) != i && tensor_of(net, i) >= 0) {
            net->layers[i].geom = net->layers[tensor_of(net, i)].geom;
            net->layers[i].halo_fill = net->layers[tensor_of(net, i)].halo_fill;
        }
    // upsample -> route: the route reads the small tensor itself when nothing else needs the upsampled one
    if (net->fusion && net->fuse_uproute && !net->keep_acc)
        for (int i = 1; i < n; ++i) {
            Layer &l = net->layers[i], &u = net->layers[i - 1];
            if (l.type != L_ROUTE || l.inputs.size() < 2 || u.type != L_UPSAMPLE || u.src < 0) continue;
            int reads = 0, others = 0;
            for (int idx : l.inputs) reads += idx == i - 1;
            for (int j = 0; j < n; ++j)
                if (j != i && (net->layers[j].type == L_ROUTE || net->layers[j].type == L_SHORTCUT))
                    for (int idx : net->layers[j].inputs) others += idx == i - 1;
            if (reads == 1 && others == 0 && l.out_h % u.stride == 0 && l.out_w % u.stride == 0) l.fuse_up = u.fused_away = true;
        }
    int launches = 1;
    for (int i = 0; i < n; ++i) {
        Layer &l = net->layers[i];
        if (l.type == L_CONV && l.conv) {
            if (rows_ok[i]) {
                l.use_rows = l.fuse_pool = net->layers[i + 1].fused_away = true;
                launches += yq_tc_rows_launches(l.conv->tc_rows) - 1;
            } else if (flat_ok[i]) {
                l.use_flat = true;
                if (net->fusion && l.quant_stop && i + 1 < n && net->layers[i + 1].type == L_YOLO && l.n % (net->layers[i + 1].classes + 5) == 0)
                    l.fuse_yolo = net->layers[i + 1].fused_away = true;
                if (can_fuse_shortcut(i)) {
                    const Layer &sc = net->layers[i + 1], &from = net->layers[tensor_of(net, sc.inputs[0])];
                    yq_act_geom g;
                    yq_act_geom_flat(l.out_h, l.out_w, &g);
                    if (same_geom(sc.geom, g) && same_geom(from.geom, g)) l.fuse_shortcut = net->layers[i + 1].fused_away = true;
                }
            } else if (patch_ok[i] || (yq_conv_geom_supported(l.conv) && !plain1_ok[i])) {
                l.use_geom = true;
            } else if (net->fusion && i + 1 < n && yq_conv_can_fuse_maxpool(l.conv)) {
                Layer &p = net->layers[i + 1];
                if (p.type == L_MAXPOOL && p.size == 2 && p.stride == 2 && p.pad == 1) l.fuse_pool = p.fused_away = true;
            }
            if (!l.use_geom && !l.fuse_pool && !plain1_ok[i] && !(l.geom.pad == 0 && l.geom.pitch_w == l.out_w && l.geom.rows_h == l.out_h)) l.use_outgeom = true;
        }
        if (l.fused_away || (l.type == L_ROUTE && l.inputs.size() == 1)) continue;
        ++launches;
    }
    // route -> flat conv: the convolution reads the route's two inputs itself when its kernel can (the CTA-pair flavour) and nobody
    // else wants the concatenated tensor; an input 0 that is an upsample folded into the route is still written, alone, into the
    // route's buffer (a quarter of the bytes the full route moved: route 20 of yolov3-tiny copied 22 MB of layer 8 per step)
    if (net->fusion && !net->keep_acc)
        for (int i = 0; i + 1 < n; ++i) {
            Layer &r = net->layers[i], &cv = net->layers[i + 1];
            if (r.type != L_ROUTE || r.inputs.size() != 2 || cv.type != L_CONV || !cv.conv || !cv.use_flat || cv.fuse_yolo || routed_from(net, i)) continue;
            const Layer &a0 = net->layers[r.inputs[0]], &a1 = net->layers[r.inputs[1]];
            const bool up0 = r.fuse_up && a0.type == L_UPSAMPLE && a0.fused_away;
            if (a1.type == L_UPSAMPLE && a1.fused_away) continue;
            if (!yq_conv_flat_cat_supported(cv.conv, a0.out_c) || a0.out_c + a1.out_c != cv.c) continue;
            yq_act_geom g;
            yq_act_geom_flat(cv.h, cv.w, &g);
            auto in_place = [&](const Layer &a) {
                const int t = tensor_of(net, (int)(&a - &net->layers[0]));
                // (a 1x1 convolution never looks at a halo position: whatever the tensor's other consumers pad with is fine)
                return t >= 0 && same_geom(net->layers[t].geom, g) && (cv.size == 1 || net->layers[t].halo_fill == r.halo_fill);
            };
            if (!same_geom(r.geom, g) || !in_place(a1) || (!up0 && !in_place(a0))) continue;
            r.cat = true;
            r.cat_copy = r.cat_buf = up0;
            if (!up0) --launches;        // nothing is copied at all
            // ... and the upsampled part is written by the 1x1 convolution in front of the upsample when that one runs the pointwise
            // flavour and nobody else reads its tensor (layer 18 -> upsample 19 -> route 20 -> layer 21 of yolov3-tiny: no launch in between)
            if (up0) {
                const int ui = r.inputs[0], pc = ui - 1;
                Layer &u = net->layers[ui];
                if (pc >= 0 && u.src == pc && u.stride == 2 && net->layers[pc].type == L_CONV && net->layers[pc].conv && net->layers[pc].use_flat &&
                    !net->layers[pc].fuse_yolo && !net->layers[pc].fuse_shortcut && !routed_from(net, pc) && !net->layers[pc].side && !r.side &&
                    yq_conv_flat_up2_supported(net->layers[pc].conv) && net->layers[pc].out_c == a0.out_c) {
                    net->layers[pc].up2_route = i;
                    r.cat_copy = false;
                    --launches;
                }
            }
        }
    // (layer 0 in the rows flavour reads the CHW planes itself when it can: no layout-transform launch, see forward_body)
    if (n > 0 && net->layers[0].type == L_CONV && net->layers[0].use_rows && !net->no_planar_input && yq_conv_rows_nchw_supported(net->layers[0].conv))
        --launches;
    net->launches = launches;
    plan_side_branches(net);
    plan_early_copies(net);
}

bool conv_output_needed(const yq_network *net, int i)
{
    if (net->keep_acc) return true;   // debug: every layer stays pullable
    return routed_from(net, i);
}

// the kernel sequence of one forward_network pass (network.c:229-261)
int forward_body(yq_network *net, const uint8_t *in_u8_nchw, int *launches, bool profile = false, bool per_image = false)
{
    cudaStream_t st = net->stream;
    int nl = 0;
    bool on_side = false, forked = false;   // (per-layer profiling keeps everything on the one stream)
    size_t first = 0;
    if (profile) cudaEventRecord(net->prof_events[0], st);
    if (per_image) {
        // Layer 0 behind the dynamic input quantiser with a different (s_in, zp_in) per image (src/blas.c:279 does it for the
        // reference's one image): the generic flavour with per-image (biases_int32, multiplier, padding value) tables writes
        // layer 0's own tensor; a max-pool the plan had fused into layer 0 runs as its own launch; the rest of the plan is unchanged.
        Layer &l0 = net->layers[0];
        if (l0.type != L_CONV) return yq::fail("per-image input quantisation: layer 0 is not a convolution");
        if (!l0.fuse_pool && !(l0.geom.pad == 0 && l0.geom.pitch_w == l0.out_w && l0.geom.rows_h == l0.out_h))
            return yq::fail("per-image input quantisation: layer 0's output tensor is halo-padded in this plan (yq_network_set_fusion(net, 0) keeps it plain)");
        if (yq_nchw_to_nhwc_u8(in_u8_nchw, net->in_plain, net->batch, net->c, net->h, net->w, st)) return -1;
        if (yq_forward_convolutional_layer_quant_per_image_gpu(l0.conv, net->in_plain, l0.out_u8, net->keep_acc ? l0.out_acc : nullptr, net->batch, net->img_bias,
                                                               net->img_mcomb, net->img_zp, net->img_pitch, st))
            return -1;
        nl += 2;
        first = 1;
        if (l0.fuse_pool) {
            Layer &p = net->layers[1];
            if (yq_forward_maxpool_layer_quant_geom_gpu(l0.out_u8, nullptr, p.out_u8, &p.geom, net->batch, p.h, p.w, p.c, p.size, p.stride, p.pad, st)) return -1;
            ++nl;
            first = 2;
        }
    }
    // layer 0 in the rows flavour reads the planes itself when it can (no layout-transform launch, no padded copy)
    const bool planar_in = !net->layers.empty() && net->layers[0].type == L_CONV && net->layers[0].use_rows && !net->no_planar_input &&
                           yq_conv_rows_nchw_supported(net->layers[0].conv) && ((uintptr_t)in_u8_nchw & 15) == 0;
    if (planar_in || per_image) {
        // nothing to do
    } else if (net->in_geom.pad) {
        if (yq_nchw_to_nhwc_u8_geom(in_u8_nchw, net->in_nhwc, net->batch, net->c, net->h, net->w, &net->in_geom, st)) return -1;
        ++nl;
    } else if (yq_nchw_to_nhwc_u8(in_u8_nchw, net->in_nhwc, net->batch, net->c, net->h, net->w, st)) {
        return -1;
    } else {
        ++nl;
    }
    if (profile) cudaEventRecord(net->prof_events[1], st);
    const uint8_t *cur = net->in_nhwc;
    const yq_act_geom *cur_geom = &net->in_geom;
    int cur_fill = net->in_halo_fill;       // the byte the current tensor's halo holds (meaningful when cur_geom->pad > 0)
    const float *cur_f32 = nullptr;
    if (first) {                            // (per-image layer 0 already ran: pick up behind it)
        const Layer &done = net->layers[first - 1];
        cur = done.out_u8;
        cur_geom = &done.geom;
        cur_fill = done.halo_fill;
    }
    const bool early_ok = !profile && net->side_stream;
    auto issue_route = [&](Layer &r, unsigned mask, cudaStream_t s, bool first_alone = false) -> int {
        const uint8_t *ins[8];
        yq_act_geom gs[8];
        int cs[8], ups[8];
        for (size_t k = 0; k < r.inputs.size(); ++k) {
            const Layer *p = &net->layers[r.inputs[k]];
            ups[k] = 1;
            if (r.fuse_up && p->type == L_UPSAMPLE && p->fused_away) {   // read the upsample's own input
                ups[k] = p->stride;
                p = &net->layers[p->src];
            }
            ins[k] = p->out_u8;
            gs[k] = p->geom;
            cs[k] = p->out_c;
        }
        // first_alone: the buffer receives input 0 only, as a tensor of that input's channel count (a route the next conv reads in two parts)
        return yq_forward_route_layer_quant_part_gpu(ins, gs, cs, ups, first_alone ? 1 : (int)r.inputs.size(), first_alone ? 1u : mask, r.out_u8, &r.geom, net->batch,
                                                     r.out_h, r.out_w, s);
    };
    for (size_t i = first; i < net->layers.size(); ++i) {
        Layer &l = net->layers[i];
        const bool side = l.side && !profile && net->side_stream;
        if (side && !on_side) {   // fork: the side stream starts behind everything issued so far
            YQ_CUDA(cudaEventRecord(net->ev_fork, net->stream));
            YQ_CUDA(cudaStreamWaitEvent(net->side_stream, net->ev_fork, 0));
            forked = true;
        }
        on_side = side;
        st = side ? net->side_stream : net->stream;
        switch (l.type) {
        case L_CONV:
            if (l.use_rows && i == 0 && planar_in) {
                if (yq_forward_convolutional_layer_quant_rows_pool_nchw_gpu(l.conv, in_u8_nchw, net->layers[1].out_u8, &net->layers[1].geom, net->batch, st))
                    return -1;
            } else if (l.use_rows) {
                if (yq_forward_convolutional_layer_quant_rows_pool_gpu(l.conv, cur, net->layers[i + 1].out_u8, &net->layers[i + 1].geom, net->batch, st))
                    return -1;
                nl += yq_tc_rows_launches(l.conv->tc_rows) - 1;
            } else if (l.use_flat && l.fuse_yolo) {
                // the head's own float tensor (l.output) is only materialised for debug pulls
                if (yq_forward_convolutional_layer_quant_flat_yolo_gpu(l.conv, cur, l.out_u8, l.halo_fill, net->keep_acc ? l.out_f32 : nullptr,
                                                                       net->layers[i + 1].out_f32, net->layers[i + 1].classes,
                                                                       net->keep_acc ? l.out_acc : nullptr, net->batch, st))
                    return -1;
            } else if (l.use_flat && l.up2_route >= 0) {
                // conv + the upsample behind it: the launch writes the upsampled tensor (the first part of what the conv behind the route reads)
                if (yq_forward_convolutional_layer_quant_flat_up2_gpu(l.conv, cur, net->layers[l.up2_route].out_u8, net->batch, st)) return -1;
            } else if (l.use_flat && i > 0 && net->layers[i - 1].cat) {
                // the route in front is not materialised: [its input 0 (brought into the route's buffer when it had to be upsampled) | its input 1]
                const Layer &r = net->layers[i - 1];
                const Layer &a0 = net->layers[r.inputs[0]], &a1 = net->layers[tensor_of(net, r.inputs[1])];
                const uint8_t *first_in = r.cat_buf ? r.out_u8 : net->layers[tensor_of(net, r.inputs[0])].out_u8;
                if (yq_forward_convolutional_layer_quant_flat_cat_gpu(l.conv, first_in, a0.out_c, a1.out_u8, l.out_u8, l.halo_fill, net->batch, st)) return -1;
            } else if (l.use_flat && l.fuse_shortcut) {
                // conv + the quantized shortcut behind it: the launch stores the shortcut's tensor (the conv's own is never written)
                Layer &sc = net->layers[i + 1];
                const Layer &from = net->layers[sc.inputs[0]];
                if (yq_forward_convolutional_layer_quant_flat_shortcut_gpu(l.conv, cur, from.out_u8, sc.out_u8, sc.halo_fill, sc.zp_b, sc.Ka, sc.Kb, sc.zp_out,
                                                                           net->batch, st))
                    return -1;
            } else if (l.use_flat) {
                if (yq_forward_convolutional_layer_quant_flat_gpu(l.conv, cur, l.out_u8, l.halo_fill, l.out_f32, net->keep_acc ? l.out_acc : nullptr,
                                                                  net->batch, st))
                    return -1;
            } else if (l.use_geom) {
                if (yq_forward_convolutional_layer_quant_geom_gpu(l.conv, cur, cur_geom, cur_geom->pad ? cur_fill : -1, l.out_u8, &l.geom, l.out_f32,
                                                                  net->keep_acc ? l.out_acc : nullptr, net->batch, st))
                    return -1;
            } else if (l.use_outgeom) {
                if (yq_forward_convolutional_layer_quant_geom_gpu(l.conv, cur, nullptr, -1, l.out_u8, &l.geom, l.out_f32, nullptr, net->batch, st)) return -1;
            } else if (l.fuse_pool) {
                uint8_t *conv_out = conv_output_needed(net, (int)i) ? l.out_u8 : nullptr;
                if (yq_forward_convolutional_layer_quant_pool_gpu(l.conv, cur, conv_out, net->layers[i + 1].out_u8, l.out_f32,
                                                                  net->keep_acc ? l.out_acc : nullptr, net->batch, st))
                    return -1;
            } else if (yq_forward_convolutional_layer_quant_gpu(l.conv, cur, l.out_u8, l.out_f32, net->keep_acc ? l.out_acc : nullptr,
                                                                net->batch, st)) {
                return -1;
            }
            ++nl;
            cur = l.out_u8;
            cur_geom = &l.geom;
            cur_fill = l.halo_fill;
            cur_f32 = l.out_f32;
            break;
        case L_SHORTCUT: {
            if (l.fused_away) {          // produced by the convolution in front of it
                cur = l.out_u8;
                cur_geom = &l.geom;
                cur_fill = l.halo_fill;
                break;
            }
            const Layer &f = net->layers[l.inputs[0]];
            if (yq_forward_shortcut_layer_quant_geom_gpu(cur, cur_geom, f.out_u8, &f.geom, l.out_u8, &l.geom, net->batch, l.h, l.w, l.c, l.zp_a, l.zp_b, l.Ka,
                                                         l.Kb, l.zp_out, st))
                return -1;
            ++nl;
            cur = l.out_u8;
            cur_geom = &l.geom;
            cur_fill = l.halo_fill;
            break;
        }
        case L_MAXPOOL:
            if (!l.fused_away) {
                if (yq_forward_maxpool_layer_quant_geom_gpu(cur, cur_geom, l.out_u8, &l.geom, net->batch, l.h, l.w, l.c, l.size, l.stride, l.pad, st))
                    return -1;
                ++nl;
            }
            cur = l.out_u8;
            cur_geom = &l.geom;
            cur_fill = l.halo_fill;
            break;
        case L_UPSAMPLE:
            if (l.fused_away) break;   // the route behind it reads `cur` through the upsample
            if (yq_forward_upsample_layer_quant_geom_gpu(cur, cur_geom, l.out_u8, &l.geom, net->batch, l.h, l.w, l.c, l.stride, st)) return -1;
            ++nl;
            cur = l.out_u8;
            cur_geom = &l.geom;
            cur_fill = l.halo_fill;
            break;
        case L_ROUTE:
            if (l.cat) {
                if (l.cat_copy) {
                    if (issue_route(l, 1u, st, true)) return -1;
                    ++nl;
                }
            } else if (l.inputs.size() > 1) {
                unsigned mask = (1u << l.inputs.size()) - 1u;
                if (early_ok && l.early_mask) {   // those inputs were copied on the side stream behind their producers
                    YQ_CUDA(cudaStreamWaitEvent(st, l.ev_early, 0));
                    mask &= ~l.early_mask;
                }
                if (mask) {
                    if (issue_route(l, mask, st)) return -1;
                    ++nl;
                }
            }   // a single-input route is an alias of its input (no copy)
            cur = l.out_u8;
            cur_geom = &l.geom;
            cur_fill = l.halo_fill;
            break;
        case L_YOLO:
            if (l.fused_away) break;
            if (!cur_f32) return yq::fail("layer %zu: yolo layer needs a float input (previous layer must be a quant_stop conv)", i);
            if (yq_forward_yolo_layer_gpu(cur_f32, l.out_f32, net->batch, l.n_anchors, l.classes, l.h, l.w, st)) return -1;
            ++nl;
            break;
        }
        if (early_ok && !side)
            for (const auto &ec : l.early_copies) {   // this layer's tensor into the routes that concatenate it much later
                Layer &r = net->layers[ec.first];
                YQ_CUDA(cudaEventRecord(net->ev_fork, net->stream));
                YQ_CUDA(cudaStreamWaitEvent(net->side_stream, net->ev_fork, 0));
                if (issue_route(r, 1u << ec.second, net->side_stream)) return -1;
                ++nl;
                YQ_CUDA(cudaEventRecord(r.ev_early, net->side_stream));
                forked = true;
            }
        if (profile) cudaEventRecord(net->prof_events[i + 2], st);
    }
    if (forked) {   // join: the forward is complete on net->stream
        YQ_CUDA(cudaEventRecord(net->ev_join, net->side_stream));
        YQ_CUDA(cudaStreamWaitEvent(net->stream, net->ev_join, 0));
    }
    if (launches) *launches = nl;
    return 0;
}

}  // namespace

extern "C" yq_network *yq_load_network(const char *cfg, const char *weights, int batch, int device)
{
    yq::clear_error();
    if (!cfg) {
        yq::fail("yq_load_network: cfg path is NULL");
        return nullptr;
    }
    if (yq_device_count() <= 0) {
        yq::fail("yq_load_network: no CUDA device visible (this library has no CPU fallback)");
        return nullptr;
    }
    std::vector<Section> secs;
    std::string err;
    if (!read_cfg(cfg, secs, err)) {
        yq::fail("%s", err.c_str());
        return nullptr;
    }
    if (secs.empty() || (secs[0].type != "[net]" && secs[0].type != "[network]")) {
        yq::fail("First section must be [net] or [network]");   // parser.c:692
        return nullptr;
    }
    std::unique_ptr<yq_network> net(new yq_network());
    net->device = device;
    if (cudaSetDevice(device) != cudaSuccess) {
        yq::fail("cudaSetDevice(%d) failed", device);
        return nullptr;
    }
    const Section &ns = secs[0];
    net->batch = batch > 0 ? batch : ns.geti("batch", 1);
    net->h = ns.geti("height", 0);
    net->w = ns.geti("width", 0);
    net->c = ns.geti("channels", 0);
    if (!net->h || !net->w || !net->c) {
        yq::fail("No input parameters supplied");   // parser.c:624
        return nullptr;
    }
    int c = net->c, h = net->h, w = net->w;
    for (size_t si = 1; si < secs.size(); ++si) {
        const Section &s = secs[si];
        const int index = (int)si - 1;
        Layer l;
        l.quantized = s.geti("quantized", 0);
        l.quant_stop = s.geti("quant_stop", 0);
        l.first_time = s.geti("first_time", 0);
        if (s.type == "[convolutional]" || s.type == "[conv]") {
            l.type = L_CONV;
            l.n = s.geti("filters", 1);
            l.size = s.geti("size", 1);
            l.stride = s.geti("stride", 1);
            int pad = s.geti("pad", 0), padding = s.geti("padding", 0);
            if (pad) padding = l.size / 2;                                  // parser.c:178
            l.pad = padding;
            if (s.geti("groups", 1) != 1) {
                yq::fail("layer %d: grouped convolution is outside the quantized path", index);
                return nullptr;
            }
            l.activation = activation_from_string(s.gets("activation", "logistic"));
            l.bn = s.geti("batch_normalize", 0);
            if (!l.quantized) {
                yq::fail("layer %d: convolutional layer without quantized=1 is outside the INT8 path", index);
                return nullptr;
            }
            if (l.activation != YQ_LINEAR && l.activation != YQ_RELU && l.activation != YQ_RELU6 && l.activation != YQ_LEAKY) {
                yq::fail("layer %d: activation '%s' has no quantized form (convolutional_layer.c:734-748)", index,
                         s.gets("activation", "logistic").c_str());
                return nullptr;
            }
            if (l.n <= 0 || l.size <= 0 || l.stride <= 0 || l.pad < 0) {
                yq::fail("layer %d: convolution needs filters, size, stride > 0 (got %d, %d, %d)", index, l.n, l.size, l.stride);
                return nullptr;
            }
            l.c = c; l.h = h; l.w = w;
            l.out_c = l.n;
            l.out_h = (h + 2 * l.pad - l.size) / l.stride + 1;
            l.out_w = (w + 2 * l.pad - l.size) / l.stride + 1;
        } else if (s.type == "[maxpool]" || s.type == "[max]") {
            l.type = L_MAXPOOL;
            l.stride = s.geti("stride", 1);
            l.size = s.geti("size", l.stride);
            l.pad = s.geti("padding", l.size - 1);                          // parser.c:415
            if (l.size <= 0 || l.stride <= 0) {
                yq::fail("layer %d: maxpool needs size, stride > 0 (got %d, %d)", index, l.size, l.stride);
                return nullptr;
            }
            l.c = c; l.h = h; l.w = w;
            l.out_c = c;
            l.out_h = (h + l.pad - l.size) / l.stride + 1;                  // maxpool_layer.c:31-32
            l.out_w = (w + l.pad - l.size) / l.stride + 1;
        } else if (s.type == "[upsample]") {
            l.type = L_UPSAMPLE;
            l.stride = s.geti("stride", 2);
            if (l.stride <= 0) {
                yq::fail("layer %d: upsample stride must be positive", index);
                return nullptr;
            }
            if (atof(s.gets("scale", "1").c_str()) != 1.0) {
                yq::fail("layer %d: upsample scale must be 1 on the quantized path (blas.c:785 asserts)", index);
                return nullptr;
            }
            l.c = c; l.h = h; l.w = w;
            l.out_c = c; l.out_h = h * l.stride; l.out_w = w * l.stride;
        } else if (s.type == "[route]") {
            l.type = L_ROUTE;
            std::string ls = s.gets("layers", "");
            if (ls.empty()) {
                yq::fail("Route Layer must specify input layers");          // parser.c:527
                return nullptr;
            }
            std::stringstream ss(ls);
            std::string tok;
            while (std::getline(ss, tok, ',')) {
                int idx = atoi(tok.c_str());
                if (idx < 0) idx = index + idx;                             // parser.c:538
                if (idx < 0 || idx >= index) {
                    yq::fail("layer %d: route input %d out of range", index, idx);
                    return nullptr;
                }
                l.inputs.push_back(idx);
            }
            if ((int)l.inputs.size() > MAX_ROUTE_INPUTS) {
                yq::fail("layer %d: route concatenates %zu layers, at most %d are supported", index, l.inputs.size(), MAX_ROUTE_INPUTS);
                return nullptr;
            }
            const Layer &first = net->layers[l.inputs[0]];
            l.out_h = first.out_h; l.out_w = first.out_w; l.out_c = 0;
            for (int idx : l.inputs) {
                const Layer &in = net->layers[idx];
                if (in.out_h != first.out_h || in.out_w != first.out_w) {
                    yq::fail("layer %d: route inputs differ in spatial size", index);
                    return nullptr;
                }
                if (in.type == L_YOLO) {
                    yq::fail("layer %d: route from a yolo layer has no uint8 tensor", index);
                    return nullptr;
                }
                l.out_c += in.out_c;
            }
            l.c = l.out_c; l.h = l.out_h; l.w = l.out_w;
        } else if (s.type == "[shortcut]") {
            // EXTENSION: the reference's shortcut (parse_shortcut parser.c:484-503, src/shortcut_layer.c) is float only; this is
            // the integer dequant-add-requant layer of include/yq_b200.h (yq_forward_shortcut_layer_quant_gpu)
            l.type = L_SHORTCUT;
            const std::string *from = s.find("from");
            int idx = from ? atoi(from->c_str()) : index;                      // parser.c:487-489
            if (idx < 0) idx = index + idx;
            if (!from || idx < 0 || idx >= index || index == 0) {
                yq::fail("layer %d: shortcut needs from= naming an earlier layer", index);
                return nullptr;
            }
            if (activation_from_string(s.gets("activation", "linear")) != YQ_LINEAR) {
                yq::fail("layer %d: the quantized shortcut has activation=linear only", index);
                return nullptr;
            }
            const Layer &f = net->layers[idx];
            if (f.type == L_YOLO || net->layers[index - 1].type == L_YOLO) {
                yq::fail("layer %d: shortcut from / after a yolo layer has no uint8 tensor", index);
                return nullptr;
            }
            if (f.out_c != c || f.out_h != h || f.out_w != w) {
                yq::fail("layer %d: shortcut inputs differ in shape (%dx%dx%d from layer %d, %dx%dx%d before)", index, f.out_c, f.out_h, f.out_w, idx, c, h, w);
                return nullptr;
            }
            l.inputs.push_back(idx);
            l.activation = YQ_LINEAR;
            l.c = c; l.h = h; l.w = w; l.out_c = c; l.out_h = h; l.out_w = w;
        } else if (s.type == "[yolo]") {
            l.type = L_YOLO;
            l.classes = s.geti("classes", 20);
            int total = s.geti("num", 1);
            std::string mask = s.gets("mask", "");
            l.n_anchors = mask.empty() ? total : (int)std::count(mask.begin(), mask.end(), ',') + 1;   // parse_yolo_mask
            {   // parse_yolo (parser.c:306-340): biases default .5, anchors= overrides, mask= selects
                std::vector<float> biases((size_t)2 * total, .5f);
                std::string an = s.gets("anchors", "");
                std::stringstream as(an);
                std::string tk;
                size_t bi = 0;
                while (std::getline(as, tk, ',') && bi < biases.size()) biases[bi++] = (float)atof(tk.c_str());
                std::vector<int> mk;
                if (mask.empty()) for (int q = 0; q < total; ++q) mk.push_back(q);
                else {
                    std::stringstream ms(mask);
                    while (std::getline(ms, tk, ',')) mk.push_back(atoi(tk.c_str()));
                }
                for (int q : mk) {
                    if (q < 0 || q >= total) {
                        yq::fail("layer %d: yolo mask index %d out of range (num=%d)", index, q, total);
                        return nullptr;
                    }
                    l.anchor_w.push_back(biases[2 * q]);
                    l.anchor_h.push_back(biases[2 * q + 1]);
                }
            }
            l.c = c; l.h = h; l.w = w; l.out_c = c; l.out_h = h; l.out_w = w;
            if (c != l.n_anchors * (l.classes + 5)) {
                yq::fail("layer %d: yolo expects %d channels, previous layer has %d", index, l.n_anchors * (l.classes + 5), c);
                return nullptr;
            }
            if (index == 0 || net->layers[index - 1].type != L_CONV || !net->layers[index - 1].quant_stop) {
                yq::fail("layer %d: yolo must follow a quant_stop=1 convolution", index);
                return nullptr;
            }
        } else {
            yq::fail("layer %d: type %s is outside the quantized inference path", index, s.type.c_str());
            return nullptr;
        }
        if (l.type != L_YOLO && l.type != L_CONV && !l.quantized) {
            yq::fail("layer %d: %s without quantized=1 is outside the INT8 path", index, s.type.c_str());
            return nullptr;
        }
        if (l.out_h <= 0 || l.out_w <= 0) {
            yq::fail("layer %d: empty output", index);
            return nullptr;
        }
        net->layers.push_back(l);
        c = l.out_c; h = l.out_h; w = l.out_w;
    }
    if (net->layers.empty()) {
        yq::fail("cfg has no layers");
        return nullptr;
    }

    // ---- weights: load_weights_upto (parser.c:1201-1305), QUANTIZATION field order
    if (!weights || !weights[0]) {
        yq::fail("yq_load_network: a quantized .weights file is required");
        return nullptr;
    }
    FILE *fp = fopen(weights, "rb");
    if (!fp) {
        yq::fail("Couldn't open file: %s", weights);   // file_error, utils.c
        return nullptr;
    }
    struct Closer { FILE *f; ~Closer() { fclose(f); } } closer{fp};
    int major = 0, minor = 0, revision = 0;
    bool ok = fread(&major, 4, 1, fp) == 1 && fread(&minor, 4, 1, fp) == 1 && fread(&revision, 4, 1, fp) == 1;
    if (ok) {
        if ((major * 10 + minor) >= 2 && major < 1000 && minor < 1000) {
            uint64_t seen;
            ok = fread(&seen, 8, 1, fp) == 1;
        } else {
            int iseen;
            ok = fread(&iseen, 4, 1, fp) == 1;
        }
    }
    for (size_t i = 0; ok && i < net->layers.size(); ++i) {
        Layer &l = net->layers[i];
        uint8_t z;
        if (l.type == L_CONV) {
            const size_t K = (size_t)l.c * l.size * l.size;
            ok = read_vec(fp, l.biases, l.n);
            if (ok && l.bn) ok = read_vec(fp, l.bn_scales, l.n) && read_vec(fp, l.bn_mean, l.n) && read_vec(fp, l.bn_var, l.n);
            ok = ok && fread(&l.s_in, 4, 1, fp) == 1 && fread(&z, 1, 1, fp) == 1;
            l.zp_in = z;
            ok = ok && fread(&l.s_out, 4, 1, fp) == 1 && fread(&z, 1, 1, fp) == 1;
            l.zp_out = z;
            ok = ok && read_vec(fp, l.s_w, l.n) && read_vec(fp, l.zp_w, l.n) && read_vec(fp, l.w_u8, K * l.n);
            if (ok) ok = fseek(fp, (long)(K * l.n * sizeof(float)), SEEK_CUR) == 0;   // float weights: unused at inference
        } else if (l.type == L_MAXPOOL || (l.type == L_UPSAMPLE && l.quantized && !l.first_time)) {
            ok = fread(&l.s_out, 4, 1, fp) == 1 && fread(&z, 1, 1, fp) == 1;
            l.zp_out = z;
        } else if (l.type == L_SHORTCUT) {
            // extension record: the layer's own output (scale, zero point), 5 bytes like load_maxpool_weights (parser.c:1161-1172)
            ok = fread(&l.s_out, 4, 1, fp) == 1 && fread(&z, 1, 1, fp) == 1;
            l.zp_out = z;
            if (ok) {
                const Layer &a = net->layers[i - 1], &b = net->layers[l.inputs[0]];
                l.s_in = a.s_out;
                l.zp_in = l.zp_a = a.zp_out;
                l.zp_b = b.zp_out;
                int32_t ka = 0, kb = 0;
                if (yq_shortcut_multiplier(a.s_out, l.s_out, &ka) || yq_shortcut_multiplier(b.s_out, l.s_out, &kb)) {
                    std::string why = yq_last_error();
                    yq::fail("layer %zu: %s", i, why.c_str());
                    return nullptr;
                }
                l.Ka = ka;
                l.Kb = kb;
            }
        } else if (l.type == L_ROUTE && l.quantized) {
            if (l.inputs.size() > 1 && !l.first_time) {                     // parser.c:1176-1182
                ok = fread(&l.s_out, 4, 1, fp) == 1 && fread(&z, 1, 1, fp) == 1;
                l.zp_out = z;
            } else {
                l.s_out = net->layers[l.inputs[0]].s_out;
                l.zp_out = net->layers[l.inputs[0]].zp_out;
            }
        }
    }
    if (!ok) {
        yq::fail("weights file %s is truncated for this cfg", weights);
        return nullptr;
    }

    // ---- one-time host prep (blas.c:259-346) and device objects
    if (cudaStreamCreateWithFlags(&net->stream, cudaStreamNonBlocking) != cudaSuccess) {
        yq::fail("cudaStreamCreate failed");
        return nullptr;
    }
    if (net->branch_stream && (cudaStreamCreateWithFlags(&net->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
                               cudaEventCreateWithFlags(&net->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                               cudaEventCreateWithFlags(&net->ev_join, cudaEventDisableTiming) != cudaSuccess)) {
        yq::fail("cudaStreamCreate (side stream) failed");
        return nullptr;
    }
    yq_network *raw = net.release();
    auto bail = [&]() -> yq_network * {
        std::string keep = yq_last_error();
        yq_free_network(raw);
        yq::fail("%s", keep.c_str());
        return nullptr;
    };
    for (size_t i = 0; i < raw->layers.size(); ++i) {
        Layer &l = raw->layers[i];
        if (l.type == L_CONV) {
            if (i > 0) {   // blas.c:301-305: input params come from the previous layer's activation params
                const Layer &p = raw->layers[i - 1];
                if (p.type == L_YOLO) {
                    yq::fail("layer %zu: convolution directly after a yolo layer has no quantized input", i);
                    return bail();
                }
                l.s_in = p.s_out;
                l.zp_in = p.zp_out;
            }
            if (prepare_conv(l, (int)i) || build_conv_device(raw, l)) return bail();
        }
        l.geom = yq_act_geom{0, l.out_w, l.out_h};
        if ((l.type == L_CONV && l.quant_stop) || l.type == L_YOLO) {
            l.f32_count = (size_t)raw->batch * l.out_c * l.out_h * l.out_w;
            if (cudaMalloc((void **)&l.out_f32, l.f32_count * sizeof(float)) != cudaSuccess) {
                yq::fail("cudaMalloc of layer %zu float output failed", i);
                return bail();
            }
            if (l.type == L_YOLO) raw->out_floats += l.f32_count;
        }
    }
    // activation buffers: plan() may switch a tensor between its plain form, the flat strip and the padded geometry a
    // rows-flavour consumer asks for -- size every buffer for the largest of them
    auto tensor_bytes = [&](int c, int h, int w) -> size_t {
        const int cs = yq::channel_stride(c);
        size_t need = (size_t)raw->batch * h * w * cs;
        yq_act_geom g;
        yq_act_geom_flat(h, w, &g);
        const size_t flat = yq_act_geom_bytes(&g, raw->batch, c);
        if (flat > need) need = flat;
        for (const Layer &k : raw->layers)   // any rows-flavour conv that could read a tensor of this shape
            if (k.type == L_CONV && k.conv && k.c == c && k.h == h && k.w == w && yq_conv_rows_supported(k.conv)) {
                yq_conv_rows_input_geom(k.conv, &g);
                const size_t rows = yq_act_geom_bytes(&g, raw->batch, c);
                if (rows > need) need = rows;
            }
        return need;
    };
    for (size_t i = 0; i < raw->layers.size(); ++i) {
        Layer &l = raw->layers[i];
        if (l.type == L_YOLO) continue;
        if (l.type == L_ROUTE && l.inputs.size() == 1) {
            l.out_u8 = raw->layers[l.inputs[0]].out_u8;   // alias (its input precedes it: already allocated)
            l.u8_bytes = raw->layers[l.inputs[0]].u8_bytes;
            continue;
        }
        l.u8_bytes = tensor_bytes(l.out_c, l.out_h, l.out_w);
        if (cudaMalloc((void **)&l.out_u8, l.u8_bytes) != cudaSuccess) {
            yq::fail("cudaMalloc of layer %zu output (%zu bytes) failed", i, l.u8_bytes);
            return bail();
        }
        l.owns_u8 = true;
        if (cudaMemset(l.out_u8, 0, l.u8_bytes) != cudaSuccess) {
            yq::fail("cudaMemset of layer %zu output failed", i);
            return bail();
        }
    }
    raw->in_nhwc_bytes = tensor_bytes(raw->c, raw->h, raw->w);
    const size_t in_bytes = (size_t)raw->batch * raw->c * raw->h * raw->w;
    if (cudaMalloc((void **)&raw->in_stage_nchw, in_bytes) != cudaSuccess || cudaMalloc((void **)&raw->in_nhwc, raw->in_nhwc_bytes) != cudaSuccess) {
        yq::fail("cudaMalloc of network input failed");
        return bail();
    }
    plan(raw);
    if (raw->plan_error) {
        yq::fail("yq_load_network: filling the tensor halos failed: %s", cudaGetErrorString(cudaGetLastError()));
        return bail();
    }
    return raw;
}

extern "C" void yq_free_network(yq_network *net)
{
    if (!net) return;
    cudaSetDevice(net->device);
    drop_graph(net);
    for (auto e : net->prof_events) cudaEventDestroy(e);
    for (auto &l : net->layers) {
        if (l.conv) yq_free_convolutional_layer_quant(l.conv);
        if (l.owns_u8) cudaFree(l.out_u8);
        cudaFree(l.out_f32);
        cudaFree(l.out_acc);
    }
    for (int k = 0; k < yq_network::PIPE; ++k) {
        cudaFree(net->pipe_in[k]);
        cudaFree(net->pipe_out[k]);
        if (net->ev_in_ready[k]) cudaEventDestroy(net->ev_in_ready[k]);
        if (net->ev_fwd_done[k]) cudaEventDestroy(net->ev_fwd_done[k]);
    }
    if (net->h2d_stream) cudaStreamDestroy(net->h2d_stream);
    if (net->d2h_stream) cudaStreamDestroy(net->d2h_stream);
    cudaFree(net->dets_dev);
    cudaFree(net->counts_dev);
    cudaFree(net->in_stage_nchw);
    cudaFree(net->in_nhwc);
    cudaFree(net->in_f32);
    cudaFree(net->in_quant);
    cudaFree(net->in_raw);
    cudaFree(net->in_plain);
    cudaFree(net->img_bias);
    cudaFree(net->img_mcomb);
    cudaFree(net->img_zp);
    cudaFree(net->scratch);
    if (net->out_host_pinned) cudaFreeHost(net->out_host_pinned);
    for (auto &l : net->layers)
        if (l.ev_early) cudaEventDestroy(l.ev_early);
    if (net->ev_fork) cudaEventDestroy(net->ev_fork);
    if (net->ev_join) cudaEventDestroy(net->ev_join);
    if (net->side_stream) cudaStreamDestroy(net->side_stream);
    if (net->stream) cudaStreamDestroy(net->stream);
    delete net;
}

extern "C" int yq_network_num_layers(const yq_network *net) { return (int)net->layers.size(); }
extern "C" int yq_network_batch(const yq_network *net) { return net->batch; }
extern "C" int yq_network_input_dims(const yq_network *net, int *c, int *h, int *w)
{
    if (c) *c = net->c;
    if (h) *h = net->h;
    if (w) *w = net->w;
    return 0;
}
extern "C" size_t yq_network_output_floats(const yq_network *net) { return net->out_floats; }
extern "C" void *yq_network_stream(yq_network *net) { return (void *)net->stream; }
extern "C" int yq_network_launches_per_forward(const yq_network *net) { return net->launches; }
extern "C" int yq_network_layer_launches(const yq_network *net, int i)
{
    if (!net || i < 0 || i >= (int)net->layers.size()) return -1;
    const Layer &l = net->layers[i];
    if (l.fused_away) return 0;
    if (l.type == L_CONV) return l.use_rows ? yq_tc_rows_launches(l.conv->tc_rows) : 1;
    if (l.type == L_ROUTE && l.cat) return l.cat_copy ? 1 : 0;      // (read in two parts by the conv behind it; only an upsampled part is written)
    if (l.type == L_ROUTE) return l.inputs.size() > 1 ? 1 : 0;      // (a single-input route is an alias; early copies aside)
    return 1;
}

extern "C" int yq_network_layer_info(const yq_network *net, int i, yq_layer_info *o)
{
    if (!net || !o || i < 0 || i >= (int)net->layers.size()) return yq::fail("yq_network_layer_info: bad index");
    const Layer &l = net->layers[i];
    memset(o, 0, sizeof *o);
    o->type = l.type; o->c = l.c; o->h = l.h; o->w = l.w; o->out_c = l.out_c; o->out_h = l.out_h; o->out_w = l.out_w;
    o->n = l.n; o->size = l.size; o->stride = l.stride; o->pad = l.pad; o->activation = l.activation;
    o->batch_normalize = l.bn; o->quant_stop_flag = l.quant_stop; o->s_in = l.s_in; o->s_out = l.s_out;
    o->zp_in = l.zp_in; o->zp_out = l.zp_out; o->kernel = l.conv ? (l.use_rows ? 3 : (l.use_flat ? 2 : yq_conv_get_kernel(l.conv))) : 0;
    o->classes = l.classes; o->n_anchors = l.n_anchors;
    o->fused = l.type == L_CONV ? (l.use_rows ? 2 : (l.fuse_pool ? 1 : (l.fuse_yolo ? 3 : (l.fuse_shortcut ? 4 : (l.up2_route >= 0 ? 6 : 0))))) : (l.fused_away ? 1 : (l.type == L_ROUTE && l.cat ? 5 : 0));
    return 0;
}

extern "C" int yq_network_set_input_quant(yq_network *net, float s_in, int zp_in)
{
    if (!net || net->layers.empty() || net->layers[0].type != L_CONV) return yq::fail("set_input_quant: layer 0 is not a convolution");
    YQ_CUDA(cudaSetDevice(net->device));
    YQ_CUDA(cudaStreamSynchronize(net->stream));
    drop_graph(net);
    Layer &l = net->layers[0];
    l.s_in = s_in;
    l.zp_in = zp_in & 0xff;
    if (prepare_conv(l, 0) || build_conv_device(net, l)) return -1;
    plan(net);
    return 0;
}

extern "C" int yq_network_set_debug(yq_network *net, int keep_acc)
{
    YQ_CUDA(cudaSetDevice(net->device));
    drop_graph(net);
    net->keep_acc = keep_acc;
    if (keep_acc)
        for (auto &l : net->layers)
            if (l.type == L_CONV && !l.out_acc) {
                size_t bytes = (size_t)net->batch * l.out_h * l.out_w * yq::channel_stride(l.out_c) * sizeof(int32_t);
                YQ_CUDA(cudaMalloc((void **)&l.out_acc, bytes));
            }
    plan(net);
    return 0;
}

extern "C" int yq_network_set_conv_kernel(yq_network *net, int kind)
{
    drop_graph(net);
    net->conv_kernel = kind;
    for (auto &l : net->layers)
        if (l.conv) apply_kernel_choice(l.conv, kind);
    plan(net);
    return 0;
}

extern "C" int yq_network_set_fusion(yq_network *net, int enable)
{
    drop_graph(net);
    net->fusion = enable;
    plan(net);
    return 0;
}

extern "C" int yq_network_use_graph(yq_network *net, int enable)
{
    if (!enable) drop_graph(net);
    net->use_graph = enable;
    return 0;
}

extern "C" int yq_forward_network_device(yq_network *net, const uint8_t *in_u8_nchw)
{
    if (!net || !in_u8_nchw) return yq::fail("yq_forward_network_device: null argument");
    if (net->plan_error) return yq::fail("yq_forward_network_device: the last re-plan failed to fill the tensor halos");
    YQ_CUDA(cudaSetDevice(net->device));
    if (!net->use_graph) return forward_body(net, in_u8_nchw, nullptr);
    cudaGraphExec_t exec = nullptr;
    for (auto &g : net->graphs)
        if (g.first == in_u8_nchw) exec = g.second;
    if (!exec) {
        if (net->graphs.size() >= 16) drop_graph(net);
        cudaGraph_t g = nullptr;
        YQ_CUDA(cudaStreamBeginCapture(net->stream, cudaStreamCaptureModeThreadLocal));
        int rc = forward_body(net, in_u8_nchw, nullptr);
        cudaError_t e = cudaStreamEndCapture(net->stream, &g);
        if (rc) {
            if (g) cudaGraphDestroy(g);
            return rc;
        }
        if (e != cudaSuccess) return yq::fail("cudaStreamEndCapture: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&exec, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return yq::fail("cudaGraphInstantiate: %s", cudaGetErrorString(e));
        net->graphs.emplace_back(in_u8_nchw, exec);
    }
    YQ_CUDA(cudaGraphLaunch(exec, net->stream));
    return 0;
}

extern "C" int yq_network_profile_forward(yq_network *net, const uint8_t *in_u8_nchw, float *layer_ms)
{
    if (!net || !in_u8_nchw || !layer_ms) return yq::fail("yq_network_profile_forward: null argument");
    YQ_CUDA(cudaSetDevice(net->device));
    const size_t ne = net->layers.size() + 2;
    while (net->prof_events.size() < ne) {
        cudaEvent_t e;
        YQ_CUDA(cudaEventCreate(&e));
        net->prof_events.push_back(e);
    }
    if (forward_body(net, in_u8_nchw, nullptr, true)) return -1;
    YQ_CUDA(cudaStreamSynchronize(net->stream));
    for (size_t i = 0; i + 1 < ne; ++i) YQ_CUDA(cudaEventElapsedTime(&layer_ms[i], net->prof_events[i], net->prof_events[i + 1]));
    return 0;
}

extern "C" int yq_network_synchronize(yq_network *net)
{
    YQ_CUDA(cudaStreamSynchronize(net->stream));
    return 0;
}

extern "C" int yq_network_predict_u8(yq_network *net, const uint8_t *in_host, float *out_host)
{
    if (!net || !in_host) return yq::fail("yq_network_predict_u8: null argument");
    YQ_CUDA(cudaSetDevice(net->device));
    const size_t in_bytes = (size_t)net->batch * net->c * net->h * net->w;
    YQ_CUDA(cudaMemcpyAsync(net->in_stage_nchw, in_host, in_bytes, cudaMemcpyHostToDevice, net->stream));
    if (yq_forward_network_device(net, net->in_stage_nchw)) return -1;
    if (out_host) {
        size_t off = 0;
        for (auto &l : net->layers)
            if (l.type == L_YOLO) {
                YQ_CUDA(cudaMemcpyAsync(out_host + off, l.out_f32, l.f32_count * sizeof(float), cudaMemcpyDeviceToHost, net->stream));
                off += l.f32_count;
            }
    }
    YQ_CUDA(cudaStreamSynchronize(net->stream));
    return 0;
}

static int predict_f32_device(yq_network *net, float *out_host);
extern "C" int yq_network_predict_f32(yq_network *net, const float *in_host, float *out_host)
{
    if (!net || !in_host || !out_host) return yq::fail("yq_network_predict_f32: null argument");
    if (net->layers.empty() || net->layers[0].type != L_CONV) return yq::fail("yq_network_predict_f32: layer 0 is not a convolution");
    YQ_CUDA(cudaSetDevice(net->device));
    const int n = net->c * net->h * net->w, B = net->batch;
    if (!net->in_f32) {
        YQ_CUDA(cudaMalloc((void **)&net->in_f32, sizeof(float) * (size_t)B * n));
        YQ_CUDA(cudaMalloc((void **)&net->in_quant, (sizeof(float) + 3 * sizeof(int)) * (size_t)B));
    }
    YQ_CUDA(cudaMemcpyAsync(net->in_f32, in_host, sizeof(float) * (size_t)B * n, cudaMemcpyHostToDevice, net->stream));
    return predict_f32_device(net, out_host);
}

// the float images are in net->in_f32: quantise per image, forward, copy the yolo heads out
static int predict_f32_device(yq_network *net, float *out_host)
{
    const int n = net->c * net->h * net->w, B = net->batch;
    float *scales = (float *)net->in_quant;
    int *zps = (int *)(scales + B), *scratch = zps + B;
    if (yq_quantize_input_gpu(net->in_f32, net->in_stage_nchw, scales, zps, scratch, B, n, net->stream)) return -1;
    std::vector<float> hs(B);
    std::vector<int> hz(B);
    YQ_CUDA(cudaMemcpyAsync(hs.data(), scales, sizeof(float) * B, cudaMemcpyDeviceToHost, net->stream));
    YQ_CUDA(cudaMemcpyAsync(hz.data(), zps, sizeof(int) * B, cudaMemcpyDeviceToHost, net->stream));
    YQ_CUDA(cudaStreamSynchronize(net->stream));
    bool uniform = true;
    for (int b = 0; b < B; ++b) {
        if (hz[b] < 0) return yq::fail("yq_network_predict_f32: image %d is all zero (blas.c:124-127 asserts)", b);
        uniform = uniform && hs[b] == hs[0] && hz[b] == hz[0];
    }
    if (uniform) {
        if (hs[0] != net->layers[0].s_in || hz[0] != net->layers[0].zp_in)
            if (yq_network_set_input_quant(net, hs[0], hz[0])) return -1;      // blas.c:279 overwrites layer 0's file values per image
        if (yq_forward_network_device(net, net->in_stage_nchw)) return -1;
    } else {
        // the images of this batch quantise differently: layer 0's multipliers, biases_int32 and padding value per image
        // (blas.c:301-334 with each image's own (s_in, zp_in)); everything behind layer 0 is the same for all of them
        const Layer &l0 = net->layers[0];
        const int pitch = yq::round_up(l0.n, 16);
        if (!net->img_bias) {
            net->img_pitch = pitch;
            YQ_CUDA(cudaMalloc((void **)&net->in_plain, (size_t)B * net->h * net->w * yq::channel_stride(net->c)));
            YQ_CUDA(cudaMalloc((void **)&net->img_bias, sizeof(int32_t) * (size_t)B * pitch));
            YQ_CUDA(cudaMalloc((void **)&net->img_mcomb, sizeof(double) * (size_t)B * pitch));
            YQ_CUDA(cudaMalloc((void **)&net->img_zp, (size_t)B));
        }
        std::vector<int32_t> hb((size_t)B * pitch, 0);
        std::vector<double> hm((size_t)B * pitch, 0.0);
        std::vector<uint8_t> hzp(B);
        Layer tmp = l0;               // host copy: prepare_conv per distinct (s_in, zp_in)
        tmp.conv = nullptr;
        for (int b = 0; b < B; ++b) {
            int same = -1;
            for (int k = 0; k < b && same < 0; ++k)
                if (hs[k] == hs[b] && hz[k] == hz[b]) same = k;
            if (same >= 0) {
                memcpy(&hb[(size_t)b * pitch], &hb[(size_t)same * pitch], sizeof(int32_t) * pitch);
                memcpy(&hm[(size_t)b * pitch], &hm[(size_t)same * pitch], sizeof(double) * pitch);
            } else {
                tmp.s_in = hs[b];
                tmp.zp_in = hz[b];
                if (prepare_conv(tmp, 0)) return -1;
                for (int oc = 0; oc < l0.n; ++oc) {
                    hb[(size_t)b * pitch + oc] = tmp.biases_int32[oc];
                    hm[(size_t)b * pitch + oc] = tmp.M_value[oc] * tmp.rshift_value[oc];   // exact: the shift is a power of two
                }
            }
            hzp[b] = (uint8_t)hz[b];
        }
        YQ_CUDA(cudaMemcpyAsync(net->img_bias, hb.data(), hb.size() * sizeof(int32_t), cudaMemcpyHostToDevice, net->stream));
        YQ_CUDA(cudaMemcpyAsync(net->img_mcomb, hm.data(), hm.size() * sizeof(double), cudaMemcpyHostToDevice, net->stream));
        YQ_CUDA(cudaMemcpyAsync(net->img_zp, hzp.data(), hzp.size(), cudaMemcpyHostToDevice, net->stream));
        {
            // the per-image layer 0 (generic flavour) writes a plain tensor: if this plan had made layer 0's output halo-padded for the
            // patch-mode convolution behind it (full yolov3: layer 1), re-plan once with that reader on its plain-input flavour
            const Layer &f0 = net->layers[0];
            if (!f0.fuse_pool && !(f0.geom.pad == 0 && f0.geom.pitch_w == f0.out_w && f0.geom.rows_h == f0.out_h) && !net->plain_l0_out) {
                YQ_CUDA(cudaStreamSynchronize(net->stream));
                drop_graph(net);
                net->plain_l0_out = true;
                plan(net);
                if (net->plan_error) return yq::fail("yq_network_predict_f32: re-plan for the per-image input quantiser failed");
            }
        }
        if (forward_body(net, net->in_stage_nchw, nullptr, false, true)) return -1;
        YQ_CUDA(cudaStreamSynchronize(net->stream));      // (the host tables above live on this stack frame)
    }
    size_t off = 0;
    for (auto &l : net->layers)
        if (l.type == L_YOLO) {
            YQ_CUDA(cudaMemcpyAsync(out_host + off, l.out_f32, l.f32_count * sizeof(float), cudaMemcpyDeviceToHost, net->stream));
            off += l.f32_count;
        }
    YQ_CUDA(cudaStreamSynchronize(net->stream));
    return 0;
}

// test_detector's input path (examples/detector.c:903-904: load_image_color, letterbox_image) for a batch of float CHW images
// of one size already decoded on the host: H2D, letterbox_image on the device, then yq_network_predict_f32's quantiser + forward.
extern "C" int yq_network_predict_image_f32(yq_network *net, const float *images_host, int ih, int iw, float *out_host)
{
    if (!net || !images_host || !out_host || ih <= 0 || iw <= 0) return yq::fail("yq_network_predict_image_f32: bad argument");
    YQ_CUDA(cudaSetDevice(net->device));
    const size_t raw = (size_t)net->batch * net->c * ih * iw, boxed = (size_t)net->batch * net->c * net->h * net->w;
    if (net->in_raw_floats < raw) {
        cudaFree(net->in_raw);
        net->in_raw = nullptr;
        net->in_raw_floats = 0;
        YQ_CUDA(cudaMalloc((void **)&net->in_raw, raw * sizeof(float)));
        net->in_raw_floats = raw;
    }
    if (!net->in_f32) {
        YQ_CUDA(cudaMalloc((void **)&net->in_f32, sizeof(float) * boxed));
        YQ_CUDA(cudaMalloc((void **)&net->in_quant, (sizeof(float) + 3 * sizeof(int)) * (size_t)net->batch));
    }
    YQ_CUDA(cudaMemcpyAsync(net->in_raw, images_host, raw * sizeof(float), cudaMemcpyHostToDevice, net->stream));
    if (yq_letterbox_image_gpu(net->in_raw, net->batch, net->c, ih, iw, net->in_f32, net->h, net->w, net->stream)) return -1;
    return predict_f32_device(net, out_host);
}

static int pipe_init(yq_network *net)
{
    if (net->h2d_stream) return 0;
    YQ_CUDA(cudaStreamCreateWithFlags(&net->h2d_stream, cudaStreamNonBlocking));
    YQ_CUDA(cudaStreamCreateWithFlags(&net->d2h_stream, cudaStreamNonBlocking));
    const size_t in_bytes = (size_t)net->batch * net->c * net->h * net->w;
    for (int k = 0; k < yq_network::PIPE; ++k) {
        YQ_CUDA(cudaMalloc((void **)&net->pipe_in[k], in_bytes));
        YQ_CUDA(cudaMalloc((void **)&net->pipe_out[k], net->out_floats * sizeof(float) + 16));
        YQ_CUDA(cudaEventCreateWithFlags(&net->ev_in_ready[k], cudaEventDisableTiming));
        YQ_CUDA(cudaEventCreateWithFlags(&net->ev_fwd_done[k], cudaEventDisableTiming));
    }
    return 0;
}

// network_predict, pipelined: enqueue H2D of one uint8 CHW batch (pinned host memory for true overlap) and its
// forward; returns the slot to pass to yq_network_collect.  At most PIPE batches may be in flight.
extern "C" int yq_network_submit_u8(yq_network *net, const uint8_t *in_host)
{
    if (!net || !in_host) return yq::fail("yq_network_submit_u8: null argument");
    YQ_CUDA(cudaSetDevice(net->device));
    if (pipe_init(net)) return -1;
    const int k = net->pipe_next;
    if (net->pipe_busy[k]) return yq::fail("yq_network_submit_u8: %d batches already in flight; collect one first", yq_network::PIPE);
    const size_t in_bytes = (size_t)net->batch * net->c * net->h * net->w;
    // pipe_in[k] was last read by the forward whose completion is ev_fwd_done[k] (already collected => complete)
    YQ_CUDA(cudaMemcpyAsync(net->pipe_in[k], in_host, in_bytes, cudaMemcpyHostToDevice, net->h2d_stream));
    YQ_CUDA(cudaEventRecord(net->ev_in_ready[k], net->h2d_stream));
    YQ_CUDA(cudaStreamWaitEvent(net->stream, net->ev_in_ready[k], 0));
    if (yq_forward_network_device(net, net->pipe_in[k])) return -1;
    size_t off = 0;
    for (auto &l : net->layers)
        if (l.type == L_YOLO) {   // heads -> per-slot device copy so the next forward may overwrite the layer buffers
            YQ_CUDA(cudaMemcpyAsync(net->pipe_out[k] + off, l.out_f32, l.f32_count * sizeof(float), cudaMemcpyDeviceToDevice, net->stream));
            off += l.f32_count;
        }
    YQ_CUDA(cudaEventRecord(net->ev_fwd_done[k], net->stream));
    net->pipe_busy[k] = true;
    net->pipe_next = (k + 1) % yq_network::PIPE;
    return k;
}

// wait for slot's forward, copy every yolo head (layer order, each [batch][out_c][out_h][out_w]) to out_host
extern "C" int yq_network_collect(yq_network *net, int slot, float *out_host)
{
    if (!net || slot < 0 || slot >= yq_network::PIPE || !net->pipe_busy[slot]) return yq::fail("yq_network_collect: slot %d is not in flight", slot);
    YQ_CUDA(cudaSetDevice(net->device));
    YQ_CUDA(cudaStreamWaitEvent(net->d2h_stream, net->ev_fwd_done[slot], 0));
    if (out_host) YQ_CUDA(cudaMemcpyAsync(out_host, net->pipe_out[slot], net->out_floats * sizeof(float), cudaMemcpyDeviceToHost, net->d2h_stream));
    YQ_CUDA(cudaStreamSynchronize(net->d2h_stream));
    net->pipe_busy[slot] = false;
    return 0;
}

extern "C" int yq_network_box_capacity(const yq_network *net)
{
    int cap = 0;
    for (auto &l : net->layers)
        if (l.type == L_YOLO) cap += l.out_h * l.out_w * l.n_anchors;
    return cap;
}

extern "C" int yq_network_classes(const yq_network *net)
{
    for (auto &l : net->layers)
        if (l.type == L_YOLO) return l.classes;
    return 0;
}

// get_network_boxes (src/network.c:635-640) + do_nms_sort (src/box.c:58-89) for every image of the last forward,
// on the device.  counts_host[batch]; dets_host[batch][capacity][5+classes] = x, y, w, h, objectness, prob[classes]
// in the reference's order before NMS (yolo layers in network order, cell, anchor); NMS zeroes suppressed probs.
extern "C" int yq_network_get_boxes(yq_network *net, int w, int h, float thresh, float nms_thresh, int relative, int *counts_host,
                                    float *dets_host)
{
    if (!net || !counts_host || !dets_host) return yq::fail("yq_network_get_boxes: null argument");
    YQ_CUDA(cudaSetDevice(net->device));
    const float *pred[8];
    int lw[8], lh[8], na[8];
    const float *bw[8], *bh[8];
    int nh = 0, classes = 0;
    for (auto &l : net->layers)
        if (l.type == L_YOLO) {
            if (nh == 8) return yq::fail("too many yolo layers");
            if (nh && l.classes != classes) return yq::fail("yolo layers disagree on the number of classes");
            classes = l.classes;
            pred[nh] = l.out_f32; lw[nh] = l.out_w; lh[nh] = l.out_h; na[nh] = l.n_anchors;
            bw[nh] = l.anchor_w.data(); bh[nh] = l.anchor_h.data();
            ++nh;
        }
    if (!nh) return yq::fail("network has no yolo layer");
    const int cap = yq_network_box_capacity(net);
    const size_t det_floats = (size_t)net->batch * cap * (5 + classes);
    if (!net->dets_dev || net->det_cap != cap || net->det_classes != classes) {
        cudaFree(net->dets_dev);
        cudaFree(net->counts_dev);
        net->dets_dev = nullptr;
        net->counts_dev = nullptr;
        YQ_CUDA(cudaMalloc((void **)&net->dets_dev, det_floats * sizeof(float)));
        YQ_CUDA(cudaMalloc((void **)&net->counts_dev, net->batch * sizeof(int)));
        net->det_cap = cap;
        net->det_classes = classes;
    }
    if (yq_detect_run(pred, lw, lh, na, bw, bh, nh, classes, net->batch, cap, net->w, net->h, w, h, relative, thresh, nms_thresh,
                      net->dets_dev, net->counts_dev, net->stream))
        return -1;
    YQ_CUDA(cudaMemcpyAsync(counts_host, net->counts_dev, net->batch * sizeof(int), cudaMemcpyDeviceToHost, net->stream));
    YQ_CUDA(cudaMemcpyAsync(dets_host, net->dets_dev, det_floats * sizeof(float), cudaMemcpyDeviceToHost, net->stream));
    YQ_CUDA(cudaStreamSynchronize(net->stream));
    return 0;
}

// a tensor the current plan never writes (its launch is fused into a neighbour's): pulls of it must fail, not return stale bytes
static const char *not_materialised(const yq_network *net, int layer, int what)
{
    const Layer &l = net->layers[layer];
    if (what == 0) {
        if (l.type == L_UPSAMPLE && l.fused_away) return "the route behind it reads its input through the upsample";
        if (l.type == L_CONV && l.fuse_pool && (l.use_rows || !conv_output_needed(net, layer))) return "its launch writes the max-pooled tensor of the next layer only";
        if (l.type == L_CONV && l.fuse_shortcut) return "its launch writes the shortcut layer's tensor only";
        if (l.type == L_ROUTE && l.cat) return "the convolution behind it reads the route's inputs itself";
        if (l.type == L_CONV && l.up2_route >= 0) return "its launch writes the upsampled tensor the convolution behind the route reads";
    } else if (what == 2) {
        if (l.type == L_CONV && l.fuse_yolo && !net->keep_acc) return "its launch writes the yolo layer's output only (pull that layer, or enable yq_network_set_debug)";
    }
    return nullptr;
}

extern "C" const float *yq_network_layer_output_f32_device(const yq_network *net, int layer)
{
    if (!net || layer < 0 || layer >= (int)net->layers.size()) return nullptr;
    if (not_materialised(net, layer, 2)) return nullptr;
    return net->layers[layer].out_f32;
}

extern "C" int yq_network_pull_layer(yq_network *net, int layer, int what, void *host_out, size_t bytes)
{
    if (!net || !host_out || layer < 0 || layer >= (int)net->layers.size()) return yq::fail("yq_network_pull_layer: bad argument");
    YQ_CUDA(cudaSetDevice(net->device));
    Layer &l = net->layers[layer];
    const size_t elems = (size_t)net->batch * l.out_c * l.out_h * l.out_w;
    const size_t esz = what == 0 ? 1 : 4;
    if (bytes != elems * esz) return yq::fail("yq_network_pull_layer: expected %zu bytes, got %zu", elems * esz, bytes);
    if (const char *why = not_materialised(net, layer, what))
        return yq::fail("layer %d is not materialised in this plan: %s (yq_network_set_fusion(net, 0) or yq_network_set_debug(net, 1) keep every tensor)", layer, why);
    if (what == 2) {
        if (!l.out_f32) return yq::fail("layer %d has no float output", layer);
        YQ_CUDA(cudaMemcpyAsync(host_out, l.out_f32, bytes, cudaMemcpyDeviceToHost, net->stream));
        YQ_CUDA(cudaStreamSynchronize(net->stream));
        return 0;
    }
    if (net->scratch_bytes < bytes) {
        cudaFree(net->scratch);
        net->scratch = nullptr;
        net->scratch_bytes = 0;
        YQ_CUDA(cudaMalloc((void **)&net->scratch, bytes));
        net->scratch_bytes = bytes;
    }
    if (what == 0) {
        if (!l.out_u8) return yq::fail("layer %d has no uint8 output", layer);
        if (yq_nhwc_to_nchw_u8_geom(l.out_u8, net->scratch, net->batch, l.out_c, l.out_h, l.out_w, &l.geom, net->stream)) return -1;
    } else if (what == 1) {
        if (!l.out_acc) return yq::fail("layer %d has no int32 accumulator (conv layers only, after yq_network_set_debug(net,1))", layer);
        if (yq_nhwc_to_nchw_i32(l.out_acc, (int32_t *)net->scratch, net->batch, l.out_c, l.out_h, l.out_w, net->stream)) return -1;
    } else {
        return yq::fail("yq_network_pull_layer: what=%d", what);
    }
    YQ_CUDA(cudaMemcpyAsync(host_out, net->scratch, bytes, cudaMemcpyDeviceToHost, net->stream));
    YQ_CUDA(cudaStreamSynchronize(net->stream));
    return 0;
}

extern "C" int yq_network_conv_params(const yq_network *net, int layer, int32_t *M0, int *M0_right_shift, double *M_value,
                                      double *M0_right_shift_value, int32_t *biases_int32)
{
    if (!net || layer < 0 || layer >= (int)net->layers.size() || net->layers[layer].type != L_CONV)
        return yq::fail("yq_network_conv_params: layer %d is not a convolution", layer);
    const Layer &l = net->layers[layer];
    if (M0) memcpy(M0, l.M0.data(), l.n * sizeof(int32_t));
    if (M0_right_shift) memcpy(M0_right_shift, l.M0_right_shift.data(), l.n * sizeof(int));
    if (M_value) memcpy(M_value, l.M_value.data(), l.n * sizeof(double));
    if (M0_right_shift_value) memcpy(M0_right_shift_value, l.rshift_value.data(), l.n * sizeof(double));
    if (biases_int32) memcpy(biases_int32, l.biases_int32.data(), l.n * sizeof(int32_t));
    return 0;
}
