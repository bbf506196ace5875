// yq_kernels.cu -- sm_100a kernels of the quantized inference hot path + the layer-level C ABI.
//
//   conv_u8_simt_kernel   implicit-GEMM uint8 x uint8 -> int32 convolution on the dp4a pipe with the
//                         requantize / bias / activation / zero-point / uint8-wrap epilogue fused in.
//                         It is the generic flavour (any c, n, size, stride, pad, zp_in); the tcgen05
//                         flavour for tensor-core-shaped layers lives in yq_conv_tc.cu.
//   maxpool / upsample / route / yolo / layout kernels: coalesced, vectorised, HBM-bound.
//
// Arithmetic contract (reference src/convolutional_layer.c:694-761):
//   acc = sum (w - zp_w[oc]) * A,  A = input or zp_in out of bounds   -- EXACT int32
//   x = acc + bias_i32[oc];  t = trunc((double)x * M_value[oc]);  q = trunc((double)t * rshift[oc])
//   activation; + zp_out; store to uint8 WRAPS (no saturation) unless desc.saturate.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <string.h>

#include <cmath>

#include <map>
#include <mutex>
#include <utility>

#include "yq_common.h"
#include "yq_epilogue.cuh"

// ------------------------------------------------------------------------------------------------
// error state
// ------------------------------------------------------------------------------------------------
namespace yq {
static thread_local char g_err[1024] = "";
static int g_abort = 0;

int fail(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    if (g_abort) {
        fprintf(stderr, "yq_b200: %s\n", g_err);
        abort();
    }
    return -1;
}
void clear_error() { g_err[0] = 0; }

namespace {
std::mutex g_dev_mu;
struct DevFacts { int n_sm = 0, smem_optin = 0, smem_sm = 0; };
std::map<int, DevFacts> g_dev_facts;
std::map<std::pair<const void *, int>, int> g_smem_optin, g_memo;
const DevFacts *dev_facts()
{
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    auto it = g_dev_facts.find(dev);
    if (it == g_dev_facts.end()) {
        DevFacts f;
        if (cudaDeviceGetAttribute(&f.n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&f.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&f.smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) != cudaSuccess)
            return nullptr;
        it = g_dev_facts.emplace(dev, f).first;
    }
    return &it->second;
}
}  // namespace
int device_index()
{
    int dev = -1;
    return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}
int device_sm_count()
{
    std::lock_guard<std::mutex> lk(g_dev_mu);
    const DevFacts *f = dev_facts();
    return f ? f->n_sm : 0;
}
int device_smem_optin()
{
    std::lock_guard<std::mutex> lk(g_dev_mu);
    const DevFacts *f = dev_facts();
    return f ? f->smem_optin : 0;
}
int device_smem_per_sm()
{
    std::lock_guard<std::mutex> lk(g_dev_mu);
    const DevFacts *f = dev_facts();
    return f ? f->smem_sm : 0;
}
int ensure_dynamic_smem(const void *kern, int bytes)
{
    std::lock_guard<std::mutex> lk(g_dev_mu);
    const int dev = device_index();
    if (dev < 0) return fail("cudaGetDevice failed");
    int &have = g_smem_optin[{kern, dev}];
    if (bytes > have) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return fail("cudaFuncSetAttribute(MaxDynamicSharedMemorySize = %d) on device %d: %s", bytes, dev, cudaGetErrorString(e));
        have = bytes;
    }
    return 0;
}
bool memo_get(const void *kern, int *value)
{
    std::lock_guard<std::mutex> lk(g_dev_mu);
    auto it = g_memo.find({kern, device_index()});
    if (it == g_memo.end()) return false;
    *value = it->second;
    return true;
}
void memo_put(const void *kern, int value)
{
    std::lock_guard<std::mutex> lk(g_dev_mu);
    g_memo[{kern, device_index()}] = value;
}
}  // namespace yq

extern "C" {
const char *yq_last_error(void) { return yq::g_err; }
void yq_set_abort_on_error(int enable) { yq::g_abort = enable; }
const char *yq_version(void) { return "yq_b200 0.1 (sm_100a)"; }
int yq_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
int yq_set_device(int device)
{
    YQ_CUDA(cudaSetDevice(device));
    return 0;
}
void *yq_cuda_malloc(size_t bytes)
{
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        yq::fail("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}
int yq_cuda_free(void *dev)
{
    YQ_CUDA(cudaFree(dev));
    return 0;
}
int yq_cuda_push(void *dev, const void *host, size_t bytes, void *stream)
{
    YQ_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return 0;
}
int yq_cuda_pull(void *host, const void *dev, size_t bytes, void *stream)
{
    YQ_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    YQ_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}
int yq_cuda_memset(void *dev, int value, size_t bytes, void *stream)
{
    YQ_CUDA(cudaMemsetAsync(dev, value, bytes, (cudaStream_t)stream));
    return 0;
}
void *yq_host_alloc(size_t bytes, int write_combined)
{
    void *p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
    if (e != cudaSuccess) {
        yq::fail("cudaHostAlloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}
int yq_host_free(void *host)
{
    YQ_CUDA(cudaFreeHost(host));
    return 0;
}
int yq_stream_synchronize(void *stream)
{
    YQ_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}
int yq_channel_stride(int c) { return yq::channel_stride(c); }
}

// ------------------------------------------------------------------------------------------------
// SIMT implicit-GEMM convolution (dp4a)
// ------------------------------------------------------------------------------------------------
struct ConvArgs {
    const uint8_t *in;
    uint8_t *out;
    float *out_f32;
    int32_t *out_acc;
    const uint8_t *wpk;
    yq::EpiParams ep;
    int B, H, W, C, CS, OH, OW, N, CSO, size, stride, pad, k_pad, k_bytes, zp_in, M_total;
    // per-IMAGE input quantisation (layer 0 behind the dynamic input quantiser, src/blas.c:279): device tables or nullptr
    const int32_t *img_bias;     // [B][img_pitch]  biases_int32 of image b
    const double *img_mcomb;     // [B][img_pitch]  M_value * M0_right_shift_value of image b
    const uint8_t *img_zp;       // [B]             input zero point of image b (the im2col pad value, src/im2col.c:5-14)
    int img_pitch;
};

constexpr int SIMT_BM = 64;        // output pixels per block
constexpr int SIMT_KC = 64;        // K bytes per stage
constexpr int SIMT_PITCH = 80;     // smem row pitch (64 + 16): conflict-free 128-bit reads
constexpr int SIMT_THREADS = 256;

// VEC = gather granularity in bytes (16 when the channel stride is a multiple of 16, else 4).
// OCT = output channels per thread; block covers BN = 16*OCT output channels.
template <int VEC, int OCT>
__global__ void __launch_bounds__(SIMT_THREADS) conv_u8_simt_kernel(const ConvArgs a)
{
    constexpr int BN = 16 * OCT;
    __shared__ __align__(16) uint8_t As[SIMT_BM * SIMT_PITCH];
    __shared__ __align__(16) uint8_t Bs[BN * SIMT_PITCH];
    __shared__ __align__(16) uint8_t Os[SIMT_BM * BN];

    const int t = threadIdx.x;
    const int m0 = blockIdx.x * SIMT_BM;
    const int oc0 = blockIdx.y * BN;

    // ---- loader role: pixel lp, 16-byte slot ls of the 64-byte stage row
    const int lp = t >> 2, ls = t & 3;
    const int lm = m0 + lp;
    const bool lvalid = lm < a.M_total;
    int ln = 0, loy = 0, lox = 0;
    if (lvalid) {
        ln = lm / (a.OH * a.OW);
        int r = lm - ln * a.OH * a.OW;
        loy = r / a.OW;
        lox = r - loy * a.OW;
    }
    const int units_per_tap = a.CS / VEC;
    const int total_units = a.size * a.size * units_per_tap;
    const int zp_in = a.img_zp && lvalid ? (int)__ldg(a.img_zp + ln) : a.zp_in;   // (per image when layer 0 follows the dynamic input quantiser)
    const uint32_t zp4 = (uint32_t)zp_in * 0x01010101u;

    auto fill_word = [&](int ch0) -> uint32_t {   // zp_in in real channels, 0 in pad channels
        if (ch0 + 4 <= a.C) return zp4;
        uint32_t v = 0;
        for (int b = 0; b < 4; ++b)
            if (ch0 + b < a.C) v |= (uint32_t)zp_in << (8 * b);
        return v;
    };

    auto load_a = [&](int stage) -> uint4 {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (!lvalid) return v;
        if (VEC == 16) {
            int gu = stage * 4 + ls;
            if (gu >= total_units) return v;
            int tap = gu / units_per_tap, cv = gu - tap * units_per_tap;
            int ky = tap / a.size, kx = tap - ky * a.size;
            int iy = loy * a.stride + ky - a.pad, ix = lox * a.stride + kx - a.pad;
            if (iy < 0 || iy >= a.H || ix < 0 || ix >= a.W) {
                v.x = fill_word(cv * 16);
                v.y = fill_word(cv * 16 + 4);
                v.z = fill_word(cv * 16 + 8);
                v.w = fill_word(cv * 16 + 12);
            } else {
                v = __ldg(reinterpret_cast<const uint4 *>(a.in + ((size_t)(ln * a.H + iy) * a.W + ix) * a.CS + cv * 16));
            }
        } else {
            uint32_t wv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int gu = stage * 16 + ls * 4 + j;
                uint32_t x = 0;
                if (gu < total_units) {
                    int tap = gu / units_per_tap, cv = gu - tap * units_per_tap;
                    int ky = tap / a.size, kx = tap - ky * a.size;
                    int iy = loy * a.stride + ky - a.pad, ix = lox * a.stride + kx - a.pad;
                    if (iy < 0 || iy >= a.H || ix < 0 || ix >= a.W)
                        x = fill_word(cv * 4);
                    else
                        x = __ldg(reinterpret_cast<const uint32_t *>(a.in + ((size_t)(ln * a.H + iy) * a.W + ix) * a.CS + cv * 4));
                }
                wv[j] = x;
            }
            v = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
        return v;
    };
    auto load_b = [&](int stage) -> uint4 {
        if (lp < BN)
            return __ldg(reinterpret_cast<const uint4 *>(a.wpk + (size_t)(oc0 + lp) * a.k_pad + stage * SIMT_KC + ls * 16));
        return make_uint4(0, 0, 0, 0);
    };

    // ---- compute role
    const int tx = t & 15, ty = t >> 4;   // oc = oc0 + tx + 16*jj ; px = ty*4 + j
    uint32_t acc[4][OCT];
    uint32_t sa[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sa[j] = 0;
#pragma unroll
        for (int jj = 0; jj < OCT; ++jj) acc[j][jj] = 0;
    }

    const int nstages = a.k_pad / SIMT_KC;
    uint4 ra = load_a(0), rb = load_b(0);
    for (int s = 0; s < nstages; ++s) {
        __syncthreads();
        *reinterpret_cast<uint4 *>(&As[lp * SIMT_PITCH + ls * 16]) = ra;
        if (lp < BN) *reinterpret_cast<uint4 *>(&Bs[lp * SIMT_PITCH + ls * 16]) = rb;
        __syncthreads();
        if (s + 1 < nstages) {
            ra = load_a(s + 1);
            rb = load_b(s + 1);
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            uint4 av[4], bv[OCT];
#pragma unroll
            for (int j = 0; j < 4; ++j) av[j] = *reinterpret_cast<const uint4 *>(&As[(ty * 4 + j) * SIMT_PITCH + kk * 16]);
#pragma unroll
            for (int jj = 0; jj < OCT; ++jj) bv[jj] = *reinterpret_cast<const uint4 *>(&Bs[(tx + 16 * jj) * SIMT_PITCH + kk * 16]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sa[j] = __dp4a(av[j].x, 0x01010101u, sa[j]);
                sa[j] = __dp4a(av[j].y, 0x01010101u, sa[j]);
                sa[j] = __dp4a(av[j].z, 0x01010101u, sa[j]);
                sa[j] = __dp4a(av[j].w, 0x01010101u, sa[j]);
#pragma unroll
                for (int jj = 0; jj < OCT; ++jj) {
                    acc[j][jj] = __dp4a(av[j].x, bv[jj].x, acc[j][jj]);
                    acc[j][jj] = __dp4a(av[j].y, bv[jj].y, acc[j][jj]);
                    acc[j][jj] = __dp4a(av[j].z, bv[jj].z, acc[j][jj]);
                    acc[j][jj] = __dp4a(av[j].w, bv[jj].w, acc[j][jj]);
                }
            }
        }
    }

    // ---- epilogue: zero-point correction, requantize, activation, wrap; stage uint8 tile in smem
#pragma unroll
    for (int jj = 0; jj < OCT; ++jj) {
        const int oc = oc0 + tx + 16 * jj;
        yq::ChanParams cp = yq::load_chan(a.ep, oc);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int p = ty * 4 + j;
            const int m = m0 + p;
            if (a.img_bias && m < a.M_total && oc < a.N) {      // this pixel's image has its own (bias, multiplier)
                const int nimg = m / (a.OH * a.OW);
                cp.bias = __ldg(a.img_bias + (size_t)nimg * a.img_pitch + oc);
                cp.m0 = __ldg(a.img_mcomb + (size_t)nimg * a.img_pitch + oc);
            }
            int accv = (int)acc[j][jj] - cp.zw * (int)sa[j];
            uint8_t r = 0;
            if (oc < a.N) r = yq::requant_u8(a.ep, cp, accv);
            Os[p * BN + tx + 16 * jj] = r;
            if (m < a.M_total && oc < a.N) {
                if (a.out_acc) a.out_acc[(size_t)m * a.CSO + oc] = accv;
                if (a.out_f32) {
                    int n = m / (a.OH * a.OW);
                    int rr = m - n * a.OH * a.OW;
                    a.out_f32[((size_t)n * a.N + oc) * a.OH * a.OW + rr] = yq::dequant_f32(a.ep, r);
                }
            }
        }
    }
    __syncthreads();
    // coalesced store: BN bytes per pixel, 16-byte vectors (BN and CSO are multiples of 16)
    constexpr int VPP = BN / 16;
    for (int i = t; i < SIMT_BM * VPP; i += SIMT_THREADS) {
        int p = i / VPP, v = i - p * VPP;
        int m = m0 + p;
        if (m < a.M_total && oc0 + v * 16 < a.CSO)
            *reinterpret_cast<uint4 *>(a.out + (size_t)m * a.CSO + oc0 + v * 16) = *reinterpret_cast<const uint4 *>(&Os[p * BN + v * 16]);
    }
}

static int launch_simt(yq_conv_layer *l, const uint8_t *in, uint8_t *out, float *out_f32, int32_t *out_acc, int batch,
                       cudaStream_t stream, const int32_t *img_bias = nullptr, const double *img_mcomb = nullptr, const uint8_t *img_zp = nullptr,
                       int img_pitch = 0)
{
    ConvArgs a;
    memset(&a, 0, sizeof a);
    a.img_bias = img_bias; a.img_mcomb = img_mcomb; a.img_zp = img_zp; a.img_pitch = img_pitch;
    a.in = in;
    a.out = out;
    a.out_f32 = l->quant_stop_flag ? out_f32 : nullptr;
    a.out_acc = out_acc;
    a.wpk = l->w_simt;
    a.ep = yq::make_epi(l);
    a.B = batch; a.H = l->h; a.W = l->w; a.C = l->c; a.CS = l->cs_in;
    a.OH = l->out_h; a.OW = l->out_w; a.N = l->n; a.CSO = l->cs_out;
    a.size = l->size; a.stride = l->stride; a.pad = l->pad;
    a.k_pad = l->k_pad; a.k_bytes = l->size * l->size * l->cs_in;
    a.zp_in = l->zp_in;
    a.M_total = batch * l->out_h * l->out_w;
    if (l->cs_out < 16) return yq::fail("conv: output channel stride %d < 16 unsupported", l->cs_out);
    const int oct = l->n > 32 ? 4 : (l->n > 16 ? 2 : 1);
    const int bn = 16 * oct;
    dim3 grid((a.M_total + SIMT_BM - 1) / SIMT_BM, l->n_pad / bn);
    const bool v16 = (l->cs_in % 16) == 0;
#define YQ_LAUNCH(V, O) conv_u8_simt_kernel<V, O><<<grid, SIMT_THREADS, 0, stream>>>(a)
    if (v16) {
        if (oct == 4) YQ_LAUNCH(16, 4); else if (oct == 2) YQ_LAUNCH(16, 2); else YQ_LAUNCH(16, 1);
    } else {
        if (oct == 4) YQ_LAUNCH(4, 4); else if (oct == 2) YQ_LAUNCH(4, 2); else YQ_LAUNCH(4, 1);
    }
#undef YQ_LAUNCH
    YQ_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// layer object
// ------------------------------------------------------------------------------------------------
static bool is_pow2_le1(double v)
{
    if (!(v > 0.0) || v > 1.0) return false;
    int e;
    return std::frexp(v, &e) == 0.5;
}

template <typename T>
static int upload(T **dst, const std::vector<T> &src)
{
    YQ_CUDA(cudaMalloc((void **)dst, src.size() * sizeof(T)));
    YQ_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" yq_conv_layer *yq_make_convolutional_layer_quant(const yq_conv_desc *d)
{
    yq::clear_error();
    if (!d || !d->weights_uint8 || !d->weight_zero_point || !d->biases_int32 || !d->M_value || !d->M0_right_shift_value) {
        yq::fail("yq_make_convolutional_layer_quant: null descriptor field");
        return nullptr;
    }
    if (d->c <= 0 || d->n <= 0 || d->size <= 0 || d->stride <= 0 || d->h <= 0 || d->w <= 0 || d->pad < 0) {
        yq::fail("yq_make_convolutional_layer_quant: bad geometry");
        return nullptr;
    }
    if (d->activation != YQ_LINEAR && d->activation != YQ_RELU && d->activation != YQ_RELU6 && d->activation != YQ_LEAKY) {
        // the reference's switch has no other case (convolutional_layer.c:734-748: default: break)
        yq::fail("yq_make_convolutional_layer_quant: activation %d has no quantized form in the reference", d->activation);
        return nullptr;
    }
    if ((long long)d->c * d->size * d->size > 33025) {
        yq::fail("yq_make_convolutional_layer_quant: K=%lld exceeds the exact-int32 bound 33025", (long long)d->c * d->size * d->size);
        return nullptr;
    }
    if (yq_device_count() <= 0) {
        yq::fail("yq_make_convolutional_layer_quant: no CUDA device (there is no CPU fallback)");
        return nullptr;
    }
    yq_conv_layer *l = new yq_conv_layer();
    l->h = d->h; l->w = d->w; l->c = d->c; l->cs_in = yq::channel_stride(d->c);
    l->n = d->n; l->cs_out = yq::channel_stride(d->n);
    l->size = d->size; l->stride = d->stride; l->pad = d->pad;
    l->out_h = (d->h + 2 * d->pad - d->size) / d->stride + 1;   // convolutional_out_height, convolutional_layer.c:55-63
    l->out_w = (d->w + 2 * d->pad - d->size) / d->stride + 1;
    l->activation = d->activation; l->quant_stop_flag = d->quant_stop_flag;
    l->zp_in = d->zp_in & 0xff; l->zp_out = d->zp_out & 0xff; l->saturate = d->saturate; l->s_out = d->s_out;
    l->kernel_req = -1;

    const int oct = l->n > 32 ? 4 : (l->n > 16 ? 2 : 1);
    l->n_pad = yq::round_up(l->n, 16 * oct);
    if (l->n_pad < l->cs_out) l->n_pad = l->cs_out;
    const int kb = l->size * l->size * l->cs_in;
    l->k_pad = yq::round_up(kb, SIMT_KC);

    const size_t K = (size_t)d->c * d->size * d->size;
    l->host_w.assign(d->weights_uint8, d->weights_uint8 + K * d->n);
    l->host_zw.assign(d->weight_zero_point, d->weight_zero_point + d->n);

    l->pack_key = yq::pack_layer_key(l);

    // pack OIHW -> [oc][ky][kx][ci (stride cs_in)], zero padded (push_convolutional_layer's role)
    std::vector<uint8_t> wp;
    const bool simt_on_dev = yq::pack_fetch_device(l, "simt", (size_t)l->n_pad * l->k_pad, (void **)&l->w_simt);
    if (!simt_on_dev && (!yq::pack_fetch(l, "simt", wp) || wp.size() != (size_t)l->n_pad * l->k_pad)) {
        wp.assign((size_t)l->n_pad * l->k_pad, 0);
        for (int oc = 0; oc < l->n; ++oc)
            for (int ci = 0; ci < l->c; ++ci)
                for (int ky = 0; ky < l->size; ++ky)
                    for (int kx = 0; kx < l->size; ++kx)
                        wp[(size_t)oc * l->k_pad + (size_t)(ky * l->size + kx) * l->cs_in + ci] =
                            d->weights_uint8[((size_t)oc * l->c + ci) * l->size * l->size + ky * l->size + kx];
        yq::pack_put(l, "simt", wp);
    }
    // per-channel parameter arrays are padded to a multiple of 128 so either flavour can read whole tiles
    const int p_pad = yq::round_up(l->n_pad, 128);
    std::vector<int32_t> bias(p_pad, 0), zw(p_pad, 0);
    std::vector<double> mcomb(p_pad, 0.0), mval(p_pad, 0.0), rsh(p_pad, 0.0);
    l->fused_mult = 1;
    l->int_form = 1;
    std::vector<int4> chanq(p_pad, make_int4(0, 0, 0, 0));
    for (int oc = 0; oc < l->n; ++oc) {
        bias[oc] = d->biases_int32[oc];
        zw[oc] = d->weight_zero_point[oc];
        mval[oc] = d->M_value[oc];
        rsh[oc] = d->M0_right_shift_value[oc];
        mcomb[oc] = mval[oc] * rsh[oc];   // exact: rsh is a power of two
        if (!is_pow2_le1(rsh[oc])) l->fused_mult = 0;
        // integer form: M_value = M0 * 2^-31 (blas.c:316), rshift = 2^-s (blas.c:315)
        const double m0d = std::ldexp(mval[oc], 31);
        int e = 0;
        const double mant = std::frexp(rsh[oc], &e);
        const int sh = 1 - e;
        if (!(m0d > 0.0 && m0d < 2147483648.0 && m0d == std::floor(m0d) && mant == 0.5 && sh >= 0 && sh <= 31)) l->int_form = 0;
        else chanq[oc] = make_int4(bias[oc], zw[oc], (int)(uint32_t)((uint64_t)m0d * 2ull), sh);
    }
    l->host_chanq.resize((size_t)p_pad * 4);
    memcpy(l->host_chanq.data(), chanq.data(), (size_t)p_pad * 16);
    l->host_mcomb = mcomb;
    if ((!simt_on_dev && upload(&l->w_simt, wp)) || upload(&l->bias, bias) || upload(&l->zw, zw) || upload(&l->mcomb, mcomb) ||
        upload(&l->mval, mval) || upload(&l->rsh, rsh) || upload((int4 **)&l->chanq, chanq)) {
        yq_free_convolutional_layer_quant(l);
        return nullptr;
    }
    l->kernel = 0;
    if (yq_tc_supported(l)) {
        if (yq_tc_prepare(l) == 0) l->kernel = 1;
    }
    if ((yq_tc_rows_supported(l) && yq_tc_rows_prepare(l, &l->tc_rows) != 0) || (yq_tc_flat_supported(l) && yq_tc_flat_prepare(l, &l->tc_flat) != 0) ||
        (yq_tc_flat2_supported(l) && yq_tc_flat2_prepare(l, &l->tc_flat2) != 0) ||
        (yq_tc_flat2x_supported(l) && yq_tc_flat2x_prepare(l, &l->tc_flat2x) != 0) ||
        (yq_tc_flat_eligible(l) && yq_tc_pw_supported(l) && yq_tc_pw_prepare(l, &l->tc_pw) != 0) ||
        (yq_tc_pwt_supported(l) && yq_tc_pwt_prepare(l, &l->tc_pwt) != 0)) {
        yq_free_convolutional_layer_quant(l);
        return nullptr;
    }
    return l;
}

extern "C" void yq_free_convolutional_layer_quant(yq_conv_layer *l)
{
    if (!l) return;
    yq_tc_free(l);
    yq_tc_rows_free(l->tc_rows);
    yq_tc_flat_free(l->tc_flat);
    yq_tc_flat2_free(l->tc_flat2);
    yq_tc_flat2x_free(l->tc_flat2x);
    yq_tc_pw_free(l->tc_pw);
    yq_tc_pw_free(l->tc_pwt);
    cudaFree(l->w_simt); cudaFree(l->bias); cudaFree(l->zw); cudaFree(l->mcomb); cudaFree(l->mval); cudaFree(l->rsh); cudaFree(l->chanq);
    delete l;
}

extern "C" int yq_conv_out_h(const yq_conv_layer *l) { return l->out_h; }
extern "C" int yq_conv_out_w(const yq_conv_layer *l) { return l->out_w; }
extern "C" int yq_conv_get_kernel(const yq_conv_layer *l) { return l->kernel; }
extern "C" int yq_conv_set_kernel(yq_conv_layer *l, int kind)
{
    if (kind == 1 && !(yq_tc_supported(l) && l->tc)) return yq::fail("tcgen05 flavour not available for this layer shape");
    l->kernel_req = kind;
    if (kind == 0) l->kernel = 0;
    else if (kind == 1) l->kernel = 1;
    else l->kernel = (yq_tc_supported(l) && l->tc) ? 1 : 0;
    return 0;
}

extern "C" int yq_forward_convolutional_layer_quant_pool_gpu(yq_conv_layer *l, const uint8_t *in_u8, uint8_t *out_u8, uint8_t *out_pool,
                                                             float *out_f32, int32_t *out_acc, int batch, void *stream)
{
    if (!l || !in_u8 || (!out_u8 && !out_pool) || batch <= 0) return yq::fail("yq_forward_convolutional_layer_quant_gpu: bad argument");
    if (l->quant_stop_flag && !out_f32) return yq::fail("quant_stop layer needs out_f32");
    if (out_pool && !yq_tc_can_fuse_pool(l)) return yq::fail("this layer's kernel flavour cannot fuse the max-pool (see yq_conv_can_fuse_maxpool)");
    // a 1x1 convolution between plain tensors is a flat strip without halo positions: the persistent two-tile kernel takes it
    if (yq_conv_plain_1x1_fast(l) && !out_pool && out_u8 && l->tc_pw && !out_acc && !(getenv("YQ_PW") && !atoi(getenv("YQ_PW"))))
        return yq_tc_pw_forward(l, l->tc_pw, in_u8, out_u8, 0, nullptr, 0, batch, (cudaStream_t)stream, 1);
    if (yq_conv_plain_1x1_fast(l) && !out_pool && out_u8)
        return yq_tc_flat2_forward(l, l->tc_flat2, in_u8, out_u8, 0, out_acc, batch, (cudaStream_t)stream, 1);
    if (l->kernel == 1) return yq_tc_forward(l, in_u8, out_u8, out_pool, out_f32, out_acc, batch, (cudaStream_t)stream);
    return launch_simt(l, in_u8, out_u8, out_f32, out_acc, batch, (cudaStream_t)stream);
}

extern "C" int yq_forward_convolutional_layer_quant_gpu(yq_conv_layer *l, const uint8_t *in_u8, uint8_t *out_u8,
                                                        float *out_f32, int32_t *out_acc, int batch, void *stream)
{
    if (!out_u8) return yq::fail("yq_forward_convolutional_layer_quant_gpu: bad argument");
    return yq_forward_convolutional_layer_quant_pool_gpu(l, in_u8, out_u8, nullptr, out_f32, out_acc, batch, stream);
}

extern "C" int yq_conv_can_fuse_maxpool(const yq_conv_layer *l) { return l ? yq_tc_can_fuse_pool(l) : 0; }

// Layer 0 behind the reference's DYNAMIC input quantiser: quantization_weights_and_activations derives (s_in, zp_in) from the
// image itself (quant_weights_with_min_max_channel, src/blas.c:108-168 via :279) and then M, biases_int32 and the padding value
// of layer 0 follow (blas.c:301-334, im2col.c:5-14).  The reference does that for its one image; a batch needs it PER IMAGE.
extern "C" int yq_forward_convolutional_layer_quant_per_image_gpu(yq_conv_layer *l, const uint8_t *in_u8, uint8_t *out_u8, int32_t *out_acc, int batch,
                                                                  const int32_t *biases_int32_dev, const double *multiplier_dev,
                                                                  const uint8_t *zp_in_dev, int table_pitch, void *stream)
{
    if (!l || !in_u8 || !out_u8 || batch <= 0 || !biases_int32_dev || !multiplier_dev || !zp_in_dev) return yq::fail("yq_forward_convolutional_layer_quant_per_image_gpu: bad argument");
    if (table_pitch < l->n) return yq::fail("per-image tables need a pitch of at least %d entries", l->n);
    if (!l->fused_mult) return yq::fail("per-image input quantisation needs power-of-two M0_right_shift values (blas.c:315)");
    if (l->quant_stop_flag) return yq::fail("per-image input quantisation: quant_stop layers are not supported");
    return launch_simt(l, in_u8, out_u8, nullptr, out_acc, batch, (cudaStream_t)stream, biases_int32_dev, multiplier_dev, zp_in_dev, table_pitch);
}

extern "C" int yq_conv_geom_supported(const yq_conv_layer *l) { return l && yq_tc_geom_supported(l) ? 1 : 0; }
// 1: the layer has the resident-bank patch-mode kernel (3x3 stride 2, narrow): it wants a halo-padded input whose halo holds its zp_in
// and then runs through yq_forward_convolutional_layer_quant_geom_gpu without side outputs
extern "C" int yq_conv_patch_supported(const yq_conv_layer *l) { return l && l->tc_pwt && !(getenv("YQ_PW") && !atoi(getenv("YQ_PW"))) ? 1 : 0; }
extern "C" int yq_conv_out_geom_supported(const yq_conv_layer *l) { return l && (yq_tc_geom_supported(l) || yq_tc_out_geom_supported(l)) ? 1 : 0; }
// 1: yq_forward_convolutional_layer_quant_gpu runs this (1x1) layer on conv_u8_tc_flat2_kernel in its plain-tensor mode
int yq_conv_plain_1x1_fast(const yq_conv_layer *l)
{
    static const bool off = getenv("YQ_NO_PLAIN_1X1") && atoi(getenv("YQ_NO_PLAIN_1X1"));   // A/B measurements
    return l && !off && l->kernel == 1 && l->size == 1 && l->tc_flat2 && !l->quant_stop_flag ? 1 : 0;
}
extern "C" int yq_forward_convolutional_layer_quant_geom_gpu(yq_conv_layer *l, const uint8_t *in_u8, const yq_act_geom *in_geom, int in_halo_fill,
                                                             uint8_t *out_u8, const yq_act_geom *out_geom, float *out_f32, int32_t *out_acc, int batch,
                                                             void *stream)
{
    if (!l || !in_u8 || !out_u8 || batch <= 0) return yq::fail("yq_forward_convolutional_layer_quant_geom_gpu: bad argument");
    if (l->quant_stop_flag && !out_f32) return yq::fail("quant_stop layer needs out_f32");
    const bool plain_in = !in_geom || (in_geom->pad == 0 && in_geom->pitch_w == l->w && in_geom->rows_h == l->h);
    const bool plain_out = !out_geom || (out_geom->pad == 0 && out_geom->pitch_w == l->out_w && out_geom->rows_h == l->out_h);
    if (plain_in && plain_out) return yq_forward_convolutional_layer_quant_pool_gpu(l, in_u8, out_u8, nullptr, out_f32, out_acc, batch, stream);
    if (plain_in && yq_tc_out_geom_supported(l) && !out_acc) {
        // the small-c flavour's threads store their pixels themselves: any output geometry
        if (out_geom->pad < 0 || out_geom->pitch_w < l->out_w + out_geom->pad || out_geom->rows_h < l->out_h + out_geom->pad)
            return yq::fail("output geometry does not hold a %dx%d tensor", l->out_h, l->out_w);
        return yq_tc_forward(l, in_u8, out_u8, nullptr, out_f32, out_acc, batch, (cudaStream_t)stream, nullptr, -1, out_geom);
    }
    // a narrow 3x3 stride-2 layer whose input halo holds its zp_in: the resident-bank kernel in patch mode (no side outputs)
    if (l->tc_pwt && !plain_in && !out_acc && !out_f32 && in_geom->pad >= l->pad && in_halo_fill == l->zp_in && in_geom->pitch_w >= l->w + in_geom->pad &&
        in_geom->rows_h >= l->h + in_geom->pad && !(getenv("YQ_PW") && !atoi(getenv("YQ_PW")))) {
        if (out_geom && (out_geom->pad < 0 || out_geom->pitch_w < l->out_w + out_geom->pad || out_geom->rows_h < l->out_h + out_geom->pad))
            return yq::fail("output geometry does not hold a %dx%d tensor", l->out_h, l->out_w);
        return yq_tc_pwt_forward(l, l->tc_pwt, in_u8, in_geom, out_u8, out_geom, batch, (cudaStream_t)stream);
    }
    if (!yq_tc_geom_supported(l)) return yq::fail("this layer's kernel flavour takes plain tensors only (see yq_conv_geom_supported)");
    if (in_geom && (in_geom->pad < 0 || in_geom->pitch_w < l->w + in_geom->pad || in_geom->rows_h < l->h + in_geom->pad))
        return yq::fail("input geometry does not hold a %dx%d tensor", l->h, l->w);
    if (out_geom && (out_geom->pad < 0 || out_geom->pitch_w < l->out_w + out_geom->pad || out_geom->rows_h < l->out_h + out_geom->pad))
        return yq::fail("output geometry does not hold a %dx%d tensor", l->out_h, l->out_w);
    return yq_tc_forward(l, in_u8, out_u8, nullptr, out_f32, out_acc, batch, (cudaStream_t)stream, in_geom, in_halo_fill, out_geom);
}

static int check_geom(const yq_act_geom *g, int h, int w);
extern "C" int yq_conv_flat_supported(const yq_conv_layer *l)
{
    if (!l) return 0;
    if (l->quant_stop_flag) return l->tc_flat ? 1 : 0;   // only the one-tile form writes the float side output
    return (l->tc_flat || l->tc_flat2 || l->tc_flat2x) ? 1 : 0;
}
extern "C" int yq_act_geom_flat(int h, int w, yq_act_geom *g)
{
    if (!g || h <= 0 || w <= 0) return yq::fail("yq_act_geom_flat: bad argument");
    yq_tc_flat_geom(h, w, g);
    return 0;
}
static int flat_dispatch(yq_conv_layer *l, const uint8_t *in_flat, uint8_t *out_flat, int halo_fill, float *out_f32, int32_t *out_acc, int batch, void *stream,
                         const yq_fused_shortcut *sc)
{
    if (!l || !in_flat || !out_flat || batch <= 0) return yq::fail("yq_forward_convolutional_layer_quant_flat_gpu: bad argument");
    if (!yq_conv_flat_supported(l)) return yq::fail("this layer has no flat flavour (see yq_conv_flat_supported)");
    if (l->quant_stop_flag && !out_f32) return yq::fail("quant_stop layer needs out_f32");
    static const bool no_flat2 = getenv("YQ_NO_FLAT2") && atoi(getenv("YQ_NO_FLAT2"));   // A/B measurements
    // CTA-pair form for the long-K layers (measured: layers 10/12/14/21 1.07-1.24x faster, the short-K layers 6/8 slower);
    // YQ_FLAT2X=0 disables it, =2 forces it wherever it exists (A/B measurements)
    static const int use_2x = getenv("YQ_FLAT2X") ? atoi(getenv("YQ_FLAT2X")) : 1;
    // 1x1 layers: the persistent two-tile form pays once there are a few waves of tile pairs (the 52 x 52 ... 208 x 208 maps of
    // the full yolov3); with one wave or less the one-tile form wins (layer 13 of yolov3-tiny: 0.0215 vs 0.0245 ms).
    // YQ_FLAT2_1X1 = 0 / 1 forces the choice (A/B measurements)
    const int one_env = getenv("YQ_FLAT2_1X1") ? atoi(getenv("YQ_FLAT2_1X1")) : -1;   // (read per call: the tests flip it)
    bool two = l->tc_flat2 != nullptr && !no_flat2 && !l->quant_stop_flag;
    if (two && l->size == 1 && l->tc_flat) {
        const long long pairs = ((long long)batch * (l->h + 1) * (l->w + 1) + 255) / 256 * ((l->n + 127) / 128);
        two = one_env >= 0 ? one_env != 0 : pairs >= 2 * 148;
    }
    if (sc && !l->tc_flat2 && !l->tc_flat2x) return yq::fail("the fused shortcut needs the persistent flat flavours (see yq_conv_flat_shortcut_supported)");
    if (sc) two = l->tc_flat2 != nullptr;          // (the one-tile form has no shortcut epilogue)
    // narrow 1x1 layers: the streaming form with the filter bank resident in shared memory (no side outputs; YQ_PW=0 switches it off per call)
    if (l->tc_pw && !sc && !out_acc && !l->quant_stop_flag && !(getenv("YQ_PW") && !atoi(getenv("YQ_PW"))))
        return yq_tc_pw_forward(l, l->tc_pw, in_flat, out_flat, halo_fill, nullptr, 0, batch, (cudaStream_t)stream, 0);
    if (l->tc_flat2x && !no_flat2 && (use_2x == 2 || (use_2x == 1 && l->c >= 256)))
        return yq_tc_flat2x_forward(l, l->tc_flat2x, in_flat, out_flat, halo_fill, out_acc, batch, (cudaStream_t)stream, sc);
    if (two) return yq_tc_flat2_forward(l, l->tc_flat2, in_flat, out_flat, halo_fill, out_acc, batch, (cudaStream_t)stream, 0, sc);
    if (l->tc_flat) return yq_tc_flat_forward(l, l->tc_flat, in_flat, out_flat, halo_fill, out_f32, nullptr, 0, out_acc, batch, (cudaStream_t)stream);
    if (l->tc_flat2) return yq_tc_flat2_forward(l, l->tc_flat2, in_flat, out_flat, halo_fill, out_acc, batch, (cudaStream_t)stream, 0, sc);
    return yq_tc_flat2x_forward(l, l->tc_flat2x, in_flat, out_flat, halo_fill, out_acc, batch, (cudaStream_t)stream, sc);
}
extern "C" int yq_forward_convolutional_layer_quant_flat_gpu(yq_conv_layer *l, const uint8_t *in_flat, uint8_t *out_flat, int halo_fill, float *out_f32,
                                                             int32_t *out_acc, int batch, void *stream)
{
    return flat_dispatch(l, in_flat, out_flat, halo_fill, out_f32, out_acc, batch, stream, nullptr);
}
extern "C" int yq_conv_flat_shortcut_supported(const yq_conv_layer *l)
{
    static const bool off = getenv("YQ_NO_FUSE_SHORTCUT") && atoi(getenv("YQ_NO_FUSE_SHORTCUT"));   // A/B measurements
    return l && !off && !l->quant_stop_flag && !l->saturate && (l->tc_flat2 || l->tc_flat2x) ? 1 : 0;
}
// A flat convolution with the FOLLOWING quantized shortcut (extension layer) fused into its epilogue: out_flat receives the
// SHORTCUT's output, the convolution's own tensor is never written.  from_flat = the shortcut's `from` tensor in the same flat
// geometry and channel count as the convolution's output; halo_fill = the byte the shortcut's consumers pad with.
extern "C" int yq_forward_convolutional_layer_quant_flat_shortcut_gpu(yq_conv_layer *l, const uint8_t *in_flat, const uint8_t *from_flat, uint8_t *out_flat,
                                                                      int halo_fill, int zp_from, int Ka, int Kb, int zp_out_shortcut, int batch, void *stream)
{
    if (!l || !in_flat || !from_flat || !out_flat || batch <= 0) return yq::fail("yq_forward_convolutional_layer_quant_flat_shortcut_gpu: bad argument");
    if (!yq_conv_flat_shortcut_supported(l)) return yq::fail("this layer cannot fuse the shortcut (see yq_conv_flat_shortcut_supported)");
    if (Ka < 1 || Kb < 1 || Ka >= (1 << 22) || Kb >= (1 << 22)) return yq::fail("shortcut: multipliers must lie in [1, 2^22) (see yq_shortcut_multiplier)");
    yq_fused_shortcut sc;
    sc.resid = from_flat;
    sc.Ka = Ka;
    sc.Kb = Kb;
    sc.C0 = 32768 + ((zp_out_shortcut & 0xff) << 16) - l->zp_out * Ka - (zp_from & 0xff) * Kb;
    return flat_dispatch(l, in_flat, out_flat, halo_fill, nullptr, nullptr, batch, stream, &sc);
}
// A flat 1x1 convolution with the FOLLOWING stride-2 upsample fused (upsample_layer.c:92-101, blas.c:334-351: every pixel four times):
// out_up_flat = the flat tensor of (2h x 2w) pixels with the layer's output channel stride; only its interior is written (its halo keeps
// the caller's fill).  The layer's own tensor is not written.
extern "C" int yq_conv_flat_up2_supported(const yq_conv_layer *l)
{
    const bool off = (getenv("YQ_NO_UP2") && atoi(getenv("YQ_NO_UP2"))) || (getenv("YQ_PW") && !atoi(getenv("YQ_PW")));   // A/B measurements
    return l && !off && l->tc_pw && !l->quant_stop_flag ? 1 : 0;
}
extern "C" int yq_forward_convolutional_layer_quant_flat_up2_gpu(yq_conv_layer *l, const uint8_t *in_flat, uint8_t *out_up_flat, int batch, void *stream)
{
    if (!l || !in_flat || !out_up_flat || batch <= 0) return yq::fail("yq_forward_convolutional_layer_quant_flat_up2_gpu: bad argument");
    if (!yq_conv_flat_up2_supported(l)) return yq::fail("this layer cannot fuse the upsample behind it (see yq_conv_flat_up2_supported)");
    return yq_tc_pw_forward(l, l->tc_pw, in_flat, out_up_flat, 0, nullptr, 0, batch, (cudaStream_t)stream, 0, out_up_flat);
}
// A flat convolution whose input is the channel concatenation [in_first (c_first channels) | in_second (the rest)] of two flat tensors of
// the layer's input geometry: a [route] in front of it that is never materialised (route_layer.c:77-95 only copies bytes; here the
// convolution's patch loads pick the tensor per channel chunk).  Both halos must hold the layer's input zero point.
extern "C" int yq_conv_flat_cat_supported(const yq_conv_layer *l, int c_first)
{
    const bool off = getenv("YQ_NO_CAT") && atoi(getenv("YQ_NO_CAT"));   // A/B measurements (read per call: the tests flip it)
    if (!l || off || l->quant_stop_flag) return 0;
    if (l->tc_pw && !(getenv("YQ_PW") && !atoi(getenv("YQ_PW")))) {     // a 1x1 layer on the pointwise flavour (its dispatch comes first)
        const int kc = yq_tc_pw_chunk(l->tc_pw);
        return c_first > 0 && c_first < l->cs_in && c_first % kc == 0 && (l->cs_in - c_first) % 16 == 0 ? 1 : 0;
    }
    if (!l->tc_flat2x || l->c < 256) return 0;
    if (getenv("YQ_NO_FLAT2") && atoi(getenv("YQ_NO_FLAT2"))) return 0;
    if (getenv("YQ_FLAT2X") && atoi(getenv("YQ_FLAT2X")) == 0) return 0;
    const int kc = yq_tc_flat2x_chunk(l->tc_flat2x);
    return c_first > 0 && c_first < l->cs_in && c_first % kc == 0 && (l->cs_in - c_first) % 16 == 0 ? 1 : 0;
}
extern "C" int yq_forward_convolutional_layer_quant_flat_cat_gpu(yq_conv_layer *l, const uint8_t *in_first, int c_first, const uint8_t *in_second,
                                                                 uint8_t *out_flat, int halo_fill, int batch, void *stream)
{
    if (!l || !in_first || !in_second || !out_flat || batch <= 0) return yq::fail("yq_forward_convolutional_layer_quant_flat_cat_gpu: bad argument");
    if (!yq_conv_flat_cat_supported(l, c_first)) return yq::fail("this layer cannot read a two-tensor input split at channel %d (see yq_conv_flat_cat_supported)", c_first);
    if (l->tc_pw && !(getenv("YQ_PW") && !atoi(getenv("YQ_PW"))))
        return yq_tc_pw_forward(l, l->tc_pw, in_second, out_flat, halo_fill, nullptr, 0, batch, (cudaStream_t)stream, 0, nullptr, in_first, c_first);
    return yq_tc_flat2x_forward(l, l->tc_flat2x, in_second, out_flat, halo_fill, nullptr, batch, (cudaStream_t)stream, nullptr, in_first, c_first);
}
extern "C" int yq_forward_convolutional_layer_quant_flat_yolo_gpu(yq_conv_layer *l, const uint8_t *in_flat, uint8_t *out_flat, int halo_fill,
                                                                  float *out_f32, float *out_yolo, int classes, int32_t *out_acc, int batch, void *stream)
{
    if (!l || !in_flat || !out_flat || !out_yolo || batch <= 0 || classes < 0) return yq::fail("yq_forward_convolutional_layer_quant_flat_yolo_gpu: bad argument");
    if (!l->tc_flat || !l->quant_stop_flag) return yq::fail("the fused yolo head needs a quant_stop layer with the flat flavour");
    if (yq_tc_pw_head_supported(l) && !out_f32 && !out_acc && !(getenv("YQ_PW") && !atoi(getenv("YQ_PW"))))      // the throughput path: no side outputs
        return yq_tc_pw_forward(l, l->tc_pw, in_flat, out_flat, halo_fill, out_yolo, classes, batch, (cudaStream_t)stream, 0);
    return yq_tc_flat_forward(l, l->tc_flat, in_flat, out_flat, halo_fill, out_f32, out_yolo, classes, out_acc, batch, (cudaStream_t)stream);
}

extern "C" int yq_conv_rows_supported(const yq_conv_layer *l) { return l && l->tc_rows ? 1 + yq_tc_rows_two_blocks(l->tc_rows) : 0; }
extern "C" int yq_conv_rows_input_geom(const yq_conv_layer *l, yq_act_geom *g)
{
    if (!l || !g || !l->tc_rows) return yq::fail("yq_conv_rows_input_geom: the layer has no rows flavour");
    yq_tc_rows_input_geom(l, g);
    return 0;
}
extern "C" size_t yq_act_geom_bytes(const yq_act_geom *g, int batch, int c)
{
    // + one trailing halo row and corner (the halo below the last image when halos are shared)
    return g ? ((size_t)batch * g->rows_h * g->pitch_w + g->pitch_w + 2) * yq::channel_stride(c) : 0;
}
extern "C" int yq_forward_convolutional_layer_quant_rows_pool_gpu(yq_conv_layer *l, const uint8_t *in_padded, uint8_t *out_pool,
                                                                  const yq_act_geom *out_geom, int batch, void *stream)
{
    if (!l || !in_padded || !out_pool || !out_geom || batch <= 0) return yq::fail("yq_forward_convolutional_layer_quant_rows_pool_gpu: bad argument");
    if (!l->tc_rows) return yq::fail("this layer has no rows flavour (see yq_conv_rows_supported)");
    if (check_geom(out_geom, l->out_h / 2, l->out_w / 2)) return -1;
    return yq_tc_rows_forward(l, l->tc_rows, in_padded, out_pool, out_geom, batch, (cudaStream_t)stream);
}
extern "C" int yq_conv_rows_nchw_supported(const yq_conv_layer *l) { return l && l->tc_rows && yq_tc_rows_planar_supported(l) ? 1 : 0; }
extern "C" int yq_forward_convolutional_layer_quant_rows_pool_nchw_gpu(yq_conv_layer *l, const uint8_t *in_nchw, uint8_t *out_pool,
                                                                       const yq_act_geom *out_geom, int batch, void *stream)
{
    if (!l || !in_nchw || !out_pool || !out_geom || batch <= 0) return yq::fail("yq_forward_convolutional_layer_quant_rows_pool_nchw_gpu: bad argument");
    if (!l->tc_rows || !yq_tc_rows_planar_supported(l)) return yq::fail("this layer cannot read CHW planes (see yq_conv_rows_nchw_supported)");
    if (check_geom(out_geom, l->out_h / 2, l->out_w / 2)) return -1;
    return yq_tc_rows_forward(l, l->tc_rows, in_nchw, out_pool, out_geom, batch, (cudaStream_t)stream, 1);
}

// ------------------------------------------------------------------------------------------------
// maxpool (src/maxpool_layer.c:109-153): out = max(0, in-bounds taps); window origin i*stride - pad/2
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 vmax16(uint4 a, uint4 b)
{
    return make_uint4(__vmaxu4(a.x, b.x), __vmaxu4(a.y, b.y), __vmaxu4(a.z, b.z), __vmaxu4(a.w, b.w));
}

// halo-padded tensor geometry on the device: pixel (n, y, x) -> pixel index
struct Geo {
    int pad, pitch, rows;
    __device__ __forceinline__ size_t pix(int n, int y, int x) const { return ((size_t)(n * rows + y + pad)) * pitch + x + pad; }
};
static inline Geo geo_of(const yq_act_geom *g, int h, int w)
{
    if (!g) return Geo{0, w, h};
    return Geo{g->pad, g->pitch_w, g->rows_h};
}
static int check_geom(const yq_act_geom *g, int h, int w)
{
    // (a halo may be shared between neighbouring rows / images: pitch_w = w + pad, rows_h = h + pad is the tightest form)
    if (g && (g->pad < 0 || g->pitch_w < w + g->pad || g->rows_h < h + g->pad)) return yq::fail("activation geometry does not hold a %dx%d tensor", h, w);
    return 0;
}

// One block row (blockIdx.x) = one output row (n, oy); threads walk (ox, vector) of that row: 32-bit index arithmetic only.
template <typename V>
__global__ void maxpool_u8_kernel(const V *__restrict__ in, V *__restrict__ out, int H, int W, int OH, int OW, int vpp /* vectors per pixel */,
                                  int size, int stride, int off, Geo gi, Geo go)
{
    yq_pdl_wait_then_release();
    const int n = blockIdx.x / OH, oy = blockIdx.x - n * OH;
    const int per_row = OW * vpp;
    for (int j = blockIdx.y * blockDim.x + threadIdx.x; j < per_row; j += gridDim.y * blockDim.x) {
        const int ox = j / vpp, v = j - ox * vpp;
        V m;
        memset(&m, 0, sizeof(V));
        for (int a = 0; a < size; ++a) {
            const int y = off + oy * stride + a;
            if (y < 0 || y >= H) continue;
            for (int b = 0; b < size; ++b) {
                const int x = off + ox * stride + b;
                if (x < 0 || x >= W) continue;
                V t = __ldg(in + gi.pix(n, y, x) * vpp + v);
                if constexpr (sizeof(V) == 16) m = vmax16(m, t);
                else m = __vmaxu4(m, t);
            }
        }
        out[go.pix(n, oy, ox) * vpp + v] = m;
    }
}

// The two pools of the reference's cfgs (size 2, stride 2 or 1) over 16-channel vectors with a power-of-two number of vectors per
// pixel: grid (x blocks, output row, image) -- no division anywhere, the four taps' row bases are block-uniform, every thread does
// four loads, twelve byte-wise maxima and one store.  (The generic kernel above spent ~380 instructions per thread on index
// arithmetic and its runtime-sized tap loops: with the tensors resident in L2 it was instruction-bound at 10 % of the L2's
// throughput -- ncu, warm caches: 10.3 / 12.3 us for layers 9 / 11 of yolov3-tiny.)
template <int STRIDE>
__global__ void __launch_bounds__(256) maxpool2_u8_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, int H, int W, int OW, int vshift, int off,
                                                          Geo gi, Geo go)
{
    yq_pdl_wait_then_release();
    const int n = blockIdx.z, oy = blockIdx.y;
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= (OW << vshift)) return;
    const int ox = j >> vshift, v = j & ((1 << vshift) - 1);
    const int y0 = off + oy * STRIDE, x0 = off + ox * STRIDE;
    const uint4 *r0 = in + ((((size_t)n * gi.rows + y0 + gi.pad) * gi.pitch + gi.pad) << vshift);
    const uint4 *r1 = r0 + ((size_t)gi.pitch << vshift);
    const bool ya = y0 >= 0 && y0 < H, yb = y0 + 1 >= 0 && y0 + 1 < H;
    const bool xa = x0 >= 0 && x0 < W, xb = x0 + 1 >= 0 && x0 + 1 < W;
    const int o0 = x0 * (1 << vshift) + v, o1 = o0 + (1 << vshift);
    const uint4 z = make_uint4(0, 0, 0, 0);
    const uint4 t00 = ya && xa ? __ldg(r0 + o0) : z, t01 = ya && xb ? __ldg(r0 + o1) : z;
    const uint4 t10 = yb && xa ? __ldg(r1 + o0) : z, t11 = yb && xb ? __ldg(r1 + o1) : z;
    out[((((size_t)n * go.rows + oy + go.pad) * go.pitch + ox + go.pad) << vshift) + v] = vmax16(vmax16(t00, t01), vmax16(t10, t11));
}

// launch shape for the row-per-block kernels: rows x ceil(per_row / threads) blocks
static inline void row_launch_shape(int rows, int per_row, dim3 *grid, int *threads)
{
    *threads = per_row >= 256 ? 256 : (per_row + 31) / 32 * 32;
    int by = (per_row + *threads - 1) / *threads;
    if (by > 64) by = 64;
    *grid = dim3((unsigned)rows, (unsigned)by);
}

static inline int grid_for(long long total, int threads)
{
    long long b = (total + threads - 1) / threads;
    long long cap = 148LL * 16;   // a few waves of the 148 SMs; kernels are grid-stride
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

extern "C" int yq_forward_maxpool_layer_quant_geom_gpu(const uint8_t *in, const yq_act_geom *in_geom, uint8_t *out, const yq_act_geom *out_geom,
                                                       int batch, int h, int w, int c, int size, int stride, int pad, void *stream)
{
    if (!in || !out || batch <= 0) return yq::fail("maxpool: bad argument");
    const int cs = yq::channel_stride(c);
    const int oh = (h + pad - size) / stride + 1, ow = (w + pad - size) / stride + 1;   // maxpool_layer.c:31-32
    if (check_geom(in_geom, h, w) || check_geom(out_geom, oh, ow)) return -1;
    const Geo gi = geo_of(in_geom, h, w), go = geo_of(out_geom, oh, ow);
    const int off = -pad / 2;
    dim3 grid;
    int threads;
    const int vpp16 = cs / 16;
    static const bool lean_off = getenv("YQ_POOL_LEAN") && !atoi(getenv("YQ_POOL_LEAN"));     // A/B measurements
    if (!lean_off && cs % 16 == 0 && (vpp16 & (vpp16 - 1)) == 0 && size == 2 && (stride == 1 || stride == 2) && oh <= 65535 && batch <= 65535) {
        int vshift = 0;
        while ((1 << vshift) < vpp16) ++vshift;
        const dim3 g2((unsigned)((ow * vpp16 + 255) / 256), (unsigned)oh, (unsigned)batch);
        if (stride == 2)
            YQ_CUDA(yq::launch_pdl(maxpool2_u8_kernel<2>, g2, dim3(256), 0, (cudaStream_t)stream, (const uint4 *)in, (uint4 *)out, h, w, ow, vshift, off, gi, go));
        else
            YQ_CUDA(yq::launch_pdl(maxpool2_u8_kernel<1>, g2, dim3(256), 0, (cudaStream_t)stream, (const uint4 *)in, (uint4 *)out, h, w, ow, vshift, off, gi, go));
    } else if (cs % 16 == 0) {
        row_launch_shape(batch * oh, ow * (cs / 16), &grid, &threads);
        YQ_CUDA(yq::launch_pdl(maxpool_u8_kernel<uint4>, grid, dim3(threads), 0, (cudaStream_t)stream, (const uint4 *)in, (uint4 *)out, h, w, oh, ow, cs / 16, size,
                               stride, off, gi, go));
    } else {
        row_launch_shape(batch * oh, ow * (cs / 4), &grid, &threads);
        YQ_CUDA(yq::launch_pdl(maxpool_u8_kernel<uint32_t>, grid, dim3(threads), 0, (cudaStream_t)stream, (const uint32_t *)in, (uint32_t *)out, h, w, oh, ow,
                               cs / 4, size, stride, off, gi, go));
    }
    YQ_CHECK_LAUNCH();
    return 0;
}
extern "C" int yq_forward_maxpool_layer_quant_gpu(const uint8_t *in, uint8_t *out, int batch, int h, int w, int c,
                                                  int size, int stride, int pad, void *stream)
{
    return yq_forward_maxpool_layer_quant_geom_gpu(in, nullptr, out, nullptr, batch, h, w, c, size, stride, pad, stream);
}

// ------------------------------------------------------------------------------------------------
// upsample (src/blas.c:781-803): out[y][x] = in[y/stride][x/stride]
// ------------------------------------------------------------------------------------------------
template <typename V>
__global__ void upsample_u8_kernel(const V *__restrict__ in, V *__restrict__ out, int H, int W, int vpp, int stride, Geo gi, Geo go)
{
    yq_pdl_wait_then_release();
    const int OW = W * stride, OH = H * stride;
    const int n = blockIdx.x / OH, oy = blockIdx.x - n * OH;
    const int per_row = OW * vpp;
    for (int j = blockIdx.y * blockDim.x + threadIdx.x; j < per_row; j += gridDim.y * blockDim.x) {
        const int ox = j / vpp, v = j - ox * vpp;
        out[go.pix(n, oy, ox) * vpp + v] = __ldg(in + gi.pix(n, oy / stride, ox / stride) * vpp + v);
    }
}

extern "C" int yq_forward_upsample_layer_quant_geom_gpu(const uint8_t *in, const yq_act_geom *in_geom, uint8_t *out, const yq_act_geom *out_geom,
                                                        int batch, int h, int w, int c, int stride, void *stream)
{
    if (!in || !out || batch <= 0 || stride <= 0) return yq::fail("upsample: bad argument");
    if (check_geom(in_geom, h, w) || check_geom(out_geom, h * stride, w * stride)) return -1;
    const Geo gi = geo_of(in_geom, h, w), go = geo_of(out_geom, h * stride, w * stride);
    const int cs = yq::channel_stride(c);
    dim3 grid;
    int threads;
    if (cs % 16 == 0) {
        row_launch_shape(batch * h * stride, w * stride * (cs / 16), &grid, &threads);
        YQ_CUDA(yq::launch_pdl(upsample_u8_kernel<uint4>, grid, dim3(threads), 0, (cudaStream_t)stream, (const uint4 *)in, (uint4 *)out, h, w, cs / 16, stride, gi, go));
    } else {
        row_launch_shape(batch * h * stride, w * stride * (cs / 4), &grid, &threads);
        YQ_CUDA(yq::launch_pdl(upsample_u8_kernel<uint32_t>, grid, dim3(threads), 0, (cudaStream_t)stream, (const uint32_t *)in, (uint32_t *)out, h, w, cs / 4, stride,
                               gi, go));
    }
    YQ_CHECK_LAUNCH();
    return 0;
}
extern "C" int yq_forward_upsample_layer_quant_gpu(const uint8_t *in, uint8_t *out, int batch, int h, int w, int c,
                                                   int stride, void *stream)
{
    return yq_forward_upsample_layer_quant_geom_gpu(in, nullptr, out, nullptr, batch, h, w, c, stride, stream);
}

// ------------------------------------------------------------------------------------------------
// route (src/route_layer.c:107-117): channel concat, no rescale.  An input may be stored at 1/up of the route's size
// and read through the nearest-neighbour upsample (src/blas.c:781-803: out[y][x] = in[y/up][x/up]), so that an
// upsample -> route pair needs no intermediate tensor.
// ------------------------------------------------------------------------------------------------
constexpr int ROUTE_MAX_INPUTS = 8;
struct RouteArgs {
    const uint8_t *in[ROUTE_MAX_INPUTS];
    int c[ROUTE_MAX_INPUTS];     // real channels
    int cs[ROUTE_MAX_INPUTS];    // channel strides
    int off[ROUTE_MAX_INPUTS];   // channel offset in the output
    Geo g[ROUTE_MAX_INPUTS];     // input geometries
    int up[ROUTE_MAX_INPUTS];    // >= 1: the input holds (H/up) x (W/up) pixels
    Geo go;
    int n, cs_out, c_out, H, W;
    int ch_lo, ch_hi;            // channels of `out` this launch covers
    unsigned mask;               // inputs this launch copies
    int sel[ROUTE_MAX_INPUTS];   // the same as a list (route_rows_u8_kernel: blockIdx.y picks one)
    long long pixels;
};

__global__ void route_u8_kernel(const RouteArgs a, uint8_t *__restrict__ out, int vec)
{
    yq_pdl_wait_then_release();
    const int n = blockIdx.x / a.H, y = blockIdx.x - n * a.H;
    if (vec) {
        const int vpp = (a.ch_hi - a.ch_lo) / 16, per_row = a.W * vpp;
        for (int j = blockIdx.y * blockDim.x + threadIdx.x; j < per_row; j += gridDim.y * blockDim.x) {
            const int x = j / vpp, ch = a.ch_lo + (j - x * vpp) * 16;
            uint4 val = make_uint4(0, 0, 0, 0);
            bool mine = ch >= a.c_out;   // (channel-stride padding behind the last input: zeros, written with the last input)
            for (int k = 0; k < a.n; ++k)
                if ((a.mask >> k & 1) && ch >= a.off[k] && ch < a.off[k] + a.c[k]) {
                    val = __ldg(reinterpret_cast<const uint4 *>(a.in[k] + a.g[k].pix(n, y / a.up[k], x / a.up[k]) * a.cs[k] + (ch - a.off[k])));
                    mine = true;
                }
            if (mine) *reinterpret_cast<uint4 *>(out + a.go.pix(n, y, x) * a.cs_out + ch) = val;
        }
    } else {
        const int cpp = a.ch_hi - a.ch_lo, per_row = a.W * cpp;
        for (int j = blockIdx.y * blockDim.x + threadIdx.x; j < per_row; j += gridDim.y * blockDim.x) {
            const int x = j / cpp, ch = a.ch_lo + (j - x * cpp);
            uint8_t val = 0;
            bool mine = ch >= a.c_out;
            for (int k = 0; k < a.n; ++k)
                if ((a.mask >> k & 1) && ch >= a.off[k] && ch < a.off[k] + a.c[k]) {
                    val = a.in[k][a.g[k].pix(n, y / a.up[k], x / a.up[k]) * a.cs[k] + (ch - a.off[k])];
                    mine = true;
                }
            if (mine) out[a.go.pix(n, y, x) * a.cs_out + ch] = val;
        }
    }
}

// The 16-byte-vector route (default since r2; YQ_ROUTE_ROWS=0 selects route_u8_kernel) with one block per
// (output row, input) -- the per-vector loop over the inputs, the row decode and the pixel address arithmetic of
// route_u8_kernel (239 warp instructions per vector, profiles/r1_route_full.csv) move out of the copy loop.
__global__ void __launch_bounds__(256) route_rows_u8_kernel(const RouteArgs a, uint8_t *__restrict__ out)
{
    yq_pdl_wait_then_release();
    const int k = a.sel[blockIdx.y];
    const int n = blockIdx.x / a.H, y = blockIdx.x - n * a.H;
    const int up = a.up[k], vk = a.c[k] >> 4;   // (vector path: c % 16 == 0, so the channel stride is c)
    const int sh = (vk & (vk - 1)) == 0 ? __ffs(vk) - 1 : -1;
    const uint4 *__restrict__ src = reinterpret_cast<const uint4 *>(a.in[k] + a.g[k].pix(n, y / up, 0) * a.cs[k]);
    uint4 *__restrict__ dst = reinterpret_cast<uint4 *>(out + a.go.pix(n, y, 0) * a.cs_out + a.off[k]);
    const int vo = a.cs_out >> 4, per_row = a.W * vk;
#pragma unroll 2
    for (int j = threadIdx.x; j < per_row; j += 256) {
        const int x = sh >= 0 ? j >> sh : j / vk, v = j - x * vk;
        const int xs = up == 1 ? x : (up == 2 ? x >> 1 : x / up);
        dst[x * vo + v] = __ldg(src + xs * vk + v);
    }
}

// input_mask: bit k set = input k is copied by this call (the other inputs' channels of `out` are left alone), so that the
// inputs of one route can be written by separate launches as they become ready.
extern "C" int yq_forward_route_layer_quant_part_gpu(const uint8_t *const *inputs, const yq_act_geom *in_geoms, const int *in_c, const int *in_up,
                                                     int n_inputs, unsigned input_mask, uint8_t *out, const yq_act_geom *out_geom, int batch, int h,
                                                     int w, void *stream)
{
    if (!inputs || !in_c || !out || n_inputs <= 0 || n_inputs > ROUTE_MAX_INPUTS || batch <= 0) return yq::fail("route: bad argument");
    input_mask &= (1u << n_inputs) - 1u;
    if (!input_mask) return yq::fail("route: empty input mask");
    RouteArgs a;
    memset(&a, 0, sizeof a);
    int off = 0, vec = 1, lo = -1, hi = 0;
    for (int i = 0; i < n_inputs; ++i) {
        const int up = in_up ? in_up[i] : 1;
        if (up < 1 || h % up || w % up) return yq::fail("route: input %d: upsample factor %d does not divide %dx%d", i, up, h, w);
        if (check_geom(in_geoms ? &in_geoms[i] : nullptr, h / up, w / up)) return -1;
        if ((input_mask >> i & 1) && !inputs[i]) return yq::fail("route: input %d is NULL", i);
        a.in[i] = inputs[i];
        a.up[i] = up;
        a.c[i] = in_c[i];
        a.cs[i] = yq::channel_stride(in_c[i]);
        a.off[i] = off;
        a.g[i] = geo_of(in_geoms ? &in_geoms[i] : nullptr, h / up, w / up);
        if (input_mask >> i & 1) {
            if (lo < 0) lo = off;
            hi = off + in_c[i];
        }
        off += in_c[i];
        if (in_c[i] % 16) vec = 0;
    }
    if (check_geom(out_geom, h, w)) return -1;
    a.go = geo_of(out_geom, h, w);
    a.n = n_inputs;
    a.mask = input_mask;
    a.c_out = off;
    a.cs_out = yq::channel_stride(off);
    a.ch_lo = lo;
    a.ch_hi = (input_mask >> (n_inputs - 1) & 1) ? a.cs_out : hi;
    a.H = h; a.W = w;
    a.pixels = (long long)batch * h * w;
    if (a.cs_out % 16) vec = 0;
    dim3 grid;
    int threads;
    // one block per (output row, input) with the index arithmetic hoisted out of the copy loop: measured on B200 (r2) 264.9 k vs
    // 261.2 k img/s for the whole yolov3-tiny step against route_u8_kernel; YQ_ROUTE_ROWS=0 keeps the old kernel for A/B runs
    const char *rows_env = getenv("YQ_ROUTE_ROWS");
    if (vec && !(rows_env && !atoi(rows_env))) {
        int nsel = 0;
        for (int i = 0; i < n_inputs; ++i)
            if (input_mask >> i & 1) a.sel[nsel++] = i;
        YQ_CUDA(yq::launch_pdl(route_rows_u8_kernel, dim3((unsigned)(batch * h), (unsigned)nsel), dim3(256), 0, (cudaStream_t)stream, a, out));
        YQ_CHECK_LAUNCH();
        return 0;
    }
    row_launch_shape(batch * h, vec ? w * ((a.ch_hi - a.ch_lo) / 16) : w * (a.ch_hi - a.ch_lo), &grid, &threads);
    YQ_CUDA(yq::launch_pdl(route_u8_kernel, grid, dim3(threads), 0, (cudaStream_t)stream, a, out, vec));
    YQ_CHECK_LAUNCH();
    return 0;
}
extern "C" int yq_forward_route_layer_quant_up_gpu(const uint8_t *const *inputs, const yq_act_geom *in_geoms, const int *in_c, const int *in_up,
                                                   int n_inputs, uint8_t *out, const yq_act_geom *out_geom, int batch, int h, int w, void *stream)
{
    return yq_forward_route_layer_quant_part_gpu(inputs, in_geoms, in_c, in_up, n_inputs, ~0u, out, out_geom, batch, h, w, stream);
}
extern "C" int yq_forward_route_layer_quant_geom_gpu(const uint8_t *const *inputs, const yq_act_geom *in_geoms, const int *in_c, int n_inputs,
                                                     uint8_t *out, const yq_act_geom *out_geom, int batch, int h, int w, void *stream)
{
    return yq_forward_route_layer_quant_up_gpu(inputs, in_geoms, in_c, nullptr, n_inputs, out, out_geom, batch, h, w, stream);
}
extern "C" int yq_forward_route_layer_quant_gpu(const uint8_t *const *inputs, const int *in_c, int n_inputs,
                                                uint8_t *out, int batch, int h, int w, void *stream)
{
    return yq_forward_route_layer_quant_geom_gpu(inputs, nullptr, in_c, n_inputs, out, nullptr, batch, h, w, stream);
}

// ------------------------------------------------------------------------------------------------
// quantized shortcut -- EXTENSION, not in the reference: its shortcut is float only (src/shortcut_layer.c:62-67 copies the
// input and adds the `from` layer's output through shortcut_cpu, src/blas.c:456-477) and forward_network would hand the next
// quantized convolution a stale input_uint8 (src/network.c:248-255).  Integer spec (include/yq_b200.h; the tests hold it to
// a plain-C restatement of the same lines):
//   Ka = round(s_a / s_out * 2^16), Kb = round(s_b / s_out * 2^16)
//   out = clamp((((a - zp_a) * Ka + (b - zp_b) * Kb + 2^15) >> 16) + zp_out, 0, 255)
// Device form: t = a * Ka + b * Kb + C0 with C0 = 2^15 + (zp_out << 16) - zp_a * Ka - zp_b * Kb, out = sat_u8(t >> 16): the same
// integer (zp_out << 16 passes through the arithmetic shift unchanged).  HBM-bound: 2 bytes read + 1 written per element.
// ------------------------------------------------------------------------------------------------
extern "C" int yq_shortcut_multiplier(float s_x, float s_out, int32_t *K)
{
    if (!K) return yq::fail("yq_shortcut_multiplier: null argument");
    if (!(s_x > 0.f) || !(s_out > 0.f)) return yq::fail("shortcut: scales must be positive (%g, %g)", (double)s_x, (double)s_out);
    const double k = round((double)s_x / (double)s_out * 65536.0);
    if (!(k >= 1.0) || k >= 4194304.0) return yq::fail("shortcut: scale ratio %g outside [2^-16, 64)", (double)s_x / (double)s_out);
    *K = (int32_t)k;
    return 0;
}

__device__ __forceinline__ uint32_t shortcut4(uint32_t a, uint32_t b, int Ka, int Kb, int C0)
{
    int t0 = (int)(a & 0xff) * Ka + (int)(b & 0xff) * Kb + C0;
    int t1 = (int)((a >> 8) & 0xff) * Ka + (int)((b >> 8) & 0xff) * Kb + C0;
    int t2 = (int)((a >> 16) & 0xff) * Ka + (int)((b >> 16) & 0xff) * Kb + C0;
    int t3 = (int)(a >> 24) * Ka + (int)(b >> 24) * Kb + C0;
    uint32_t lo, hi;
    // cvt.pack.sat.u8.s32: d = {sat_u8(x), sat_u8(y)} in the low half, the upper half from c
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(t3 >> 16), "r"(t2 >> 16));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(t1 >> 16), "r"(t0 >> 16), "r"(hi));
    return lo;
}

// One block row = one image row (n, y); threads walk its 16-byte vectors.  VEC = 0: one byte per thread (c % 16 != 0 never
// happens for channel strides > 4, kept for cs = 4).
__device__ __forceinline__ uint32_t keep_bytes(uint32_t v, int n)   // keep the n low bytes (n <= 0: none, n >= 4: all)
{
    return n >= 4 ? v : (n <= 0 ? 0u : v & ((1u << (8 * n)) - 1u));
}

__global__ void __launch_bounds__(256) shortcut_u8_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, uint8_t *__restrict__ out, int H, int W,
                                                          int c, int cs, int Ka, int Kb, int C0, Geo ga, Geo gb, Geo go)
{
    yq_pdl_wait_then_release();
    const int n = blockIdx.x / H, y = blockIdx.x - n * H;
    if (cs % 16 == 0) {
        const int per_row = W * (cs / 16);   // a row's interior is contiguous in every geometry
        const uint4 *pa = reinterpret_cast<const uint4 *>(a + ga.pix(n, y, 0) * cs);
        const uint4 *pb = reinterpret_cast<const uint4 *>(b + gb.pix(n, y, 0) * cs);
        uint4 *po = reinterpret_cast<uint4 *>(out + go.pix(n, y, 0) * cs);
        for (int j = blockIdx.y * blockDim.x + threadIdx.x; j < per_row; j += gridDim.y * blockDim.x) {
            const uint4 va = __ldg(pa + j), vb = __ldg(pb + j);
            uint4 r = make_uint4(shortcut4(va.x, vb.x, Ka, Kb, C0), shortcut4(va.y, vb.y, Ka, Kb, C0), shortcut4(va.z, vb.z, Ka, Kb, C0),
                                 shortcut4(va.w, vb.w, Ka, Kb, C0));
            if (c != cs) {   // pad channels stay zero
                const int left = c - (j % (cs / 16)) * 16;
                r = make_uint4(keep_bytes(r.x, left), keep_bytes(r.y, left - 4), keep_bytes(r.z, left - 8), keep_bytes(r.w, left - 12));
            }
            po[j] = r;
        }
    } else {
        const int per_row = W * (cs / 4);
        const uint32_t *pa = reinterpret_cast<const uint32_t *>(a + ga.pix(n, y, 0) * cs);
        const uint32_t *pb = reinterpret_cast<const uint32_t *>(b + gb.pix(n, y, 0) * cs);
        uint32_t *po = reinterpret_cast<uint32_t *>(out + go.pix(n, y, 0) * cs);
        for (int j = blockIdx.y * blockDim.x + threadIdx.x; j < per_row; j += gridDim.y * blockDim.x)
            po[j] = keep_bytes(shortcut4(__ldg(pa + j), __ldg(pb + j), Ka, Kb, C0), c - (j % (cs / 4)) * 4);
    }
}

extern "C" int yq_forward_shortcut_layer_quant_geom_gpu(const uint8_t *a, const yq_act_geom *a_geom, const uint8_t *b, const yq_act_geom *b_geom, uint8_t *out,
                                                        const yq_act_geom *out_geom, int batch, int h, int w, int c, int zp_a, int zp_b, int Ka, int Kb,
                                                        int zp_out, void *stream)
{
    if (!a || !b || !out || batch <= 0 || h <= 0 || w <= 0 || c <= 0) return yq::fail("shortcut: bad argument");
    if (Ka < 1 || Kb < 1 || Ka >= (1 << 22) || Kb >= (1 << 22)) return yq::fail("shortcut: multipliers must lie in [1, 2^22) (see yq_shortcut_multiplier)");
    if (check_geom(a_geom, h, w) || check_geom(b_geom, h, w) || check_geom(out_geom, h, w)) return -1;
    const int cs = yq::channel_stride(c);
    const int C0 = 32768 + ((zp_out & 0xff) << 16) - (zp_a & 0xff) * Ka - (zp_b & 0xff) * Kb;
    dim3 grid;
    int threads;
    row_launch_shape(batch * h, w * (cs % 16 == 0 ? cs / 16 : cs / 4), &grid, &threads);
    YQ_CUDA(yq::launch_pdl(shortcut_u8_kernel, grid, dim3(threads), 0, (cudaStream_t)stream, a, b, out, h, w, c, cs, Ka, Kb, C0, geo_of(a_geom, h, w),
                           geo_of(b_geom, h, w), geo_of(out_geom, h, w)));
    YQ_CHECK_LAUNCH();
    return 0;
}
extern "C" int yq_forward_shortcut_layer_quant_gpu(const uint8_t *a, const uint8_t *b, uint8_t *out, int batch, int h, int w, int c, int zp_a, int zp_b,
                                                   int Ka, int Kb, int zp_out, void *stream)
{
    return yq_forward_shortcut_layer_quant_geom_gpu(a, nullptr, b, nullptr, out, nullptr, batch, h, w, c, zp_a, zp_b, Ka, Kb, zp_out, stream);
}

// ------------------------------------------------------------------------------------------------
// yolo head (src/yolo_layer.c:125-146): logistic on x,y and obj+classes, double exp like
// logistic_activate (src/activations.h:32)
// ------------------------------------------------------------------------------------------------
__global__ void yolo_kernel(const float *__restrict__ in, float *__restrict__ out, int per, int hw, long long total)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int e = (int)((i / hw) % per);
        float x = in[i];
        if (e != 2 && e != 3) x = (float)(1. / (1. + exp(-(double)x)));
        out[i] = x;
    }
}

extern "C" int yq_forward_yolo_layer_gpu(const float *in, float *out, int batch, int n_anchors, int classes, int h,
                                         int w, void *stream)
{
    if (!in || !out || batch <= 0) return yq::fail("yolo: bad argument");
    const int per = 4 + classes + 1;
    long long total = (long long)batch * n_anchors * per * h * w;
    yolo_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, per, h * w, total);
    YQ_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// layout conversion at the boundary
// ------------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_u8_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int C, int HW, int CS,
                                       long long total /* B*HW*(CS/4) */)
{
    const int wpp = CS / 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        // pixel-fastest thread order: reads of each channel plane are coalesced across the warp
        long long g = i / HW;            // (n, word)
        int p = (int)(i - g * HW);
        int wd = (int)(g % wpp);
        long long n = g / wpp;
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int ch = wd * 4 + b;
            if (ch < C) v |= (uint32_t)in[((size_t)n * C + ch) * HW + p] << (8 * b);
        }
        *reinterpret_cast<uint32_t *>(out + ((size_t)n * HW + p) * CS + wd * 4) = v;
    }
}

// c <= 4, HW % 4 == 0: one thread = 4 consecutive pixels: one 4-byte load per channel plane, one 16-byte store
__global__ void nchw_to_nhwc4_u8_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int C, int HW, long long total /* B*HW/4 */)
{
    const int q = HW / 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / q;
        const int p4 = (int)(i - n * q);
        uint32_t pl[4] = {0, 0, 0, 0};
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
            if (ch < C) pl[ch] = __ldg(reinterpret_cast<const uint32_t *>(in + ((size_t)n * C + ch) * HW) + p4);
        uint4 o;   // pixel k = byte k of every plane word
        o.x = __byte_perm(__byte_perm(pl[0], pl[1], 0x0040), __byte_perm(pl[2], pl[3], 0x0040), 0x5410);
        o.y = __byte_perm(__byte_perm(pl[0], pl[1], 0x0051), __byte_perm(pl[2], pl[3], 0x0051), 0x5410);
        o.z = __byte_perm(__byte_perm(pl[0], pl[1], 0x0062), __byte_perm(pl[2], pl[3], 0x0062), 0x5410);
        o.w = __byte_perm(__byte_perm(pl[0], pl[1], 0x0073), __byte_perm(pl[2], pl[3], 0x0073), 0x5410);
        reinterpret_cast<uint4 *>(out)[i] = o;
    }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T *__restrict__ in, T *__restrict__ out, int C, int HW, int CS, long long total)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int p = (int)(i % HW);
        long long g = i / HW;
        int ch = (int)(g % C);
        long long n = g / C;
        out[i] = in[((size_t)n * HW + p) * CS + ch];
    }
}

// ---- halo-padded variants: pixel (n, y, x) at ((n*rows_h + y + pad)*pitch_w + x + pad)*CS
__global__ void nchw_to_nhwc_u8_geom_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int C, int H, int W, int CS, int pad,
                                            int pitch_w, int rows_h, long long total /* B*HW*(CS/4) */)
{
    const int wpp = CS / 4, HW = H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long g = i / HW;
        int p = (int)(i - g * HW);
        int wd = (int)(g % wpp);
        long long n = g / wpp;
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int ch = wd * 4 + b;
            if (ch < C) v |= (uint32_t)in[((size_t)n * C + ch) * HW + p] << (8 * b);
        }
        const int y = p / W, x = p - y * W;
        *reinterpret_cast<uint32_t *>(out + (((size_t)n * rows_h + y + pad) * pitch_w + x + pad) * CS + wd * 4) = v;
    }
}

// c <= 4, W % 4 == 0, pad == 1, pitch_w % 4 == 0: one thread = the 16-byte-aligned output chunk holding pixels 4j-1 .. 4j+2
// of one row (two aligned 4-byte loads per channel plane, funnel-shifted by one pixel; one 16-byte store).  The two chunks
// at the row ends store only their in-image pixels, so the halo is never touched.
template <int RY>
__global__ void nchw_to_nhwc4_u8_geom_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int C, int H, int W, int pitch_w,
                                             int rows_h, int total_rows /* B*H */)
{
    // one block = RY consecutive image rows, one thread = one 16-byte output chunk of each of them; every load of the RY
    // rows is issued before the first use (the kernel is a pure latency / bandwidth problem)
    yq_pdl_wait_then_release();
    const int HW = H * W, wq = W / 4;
    const int row0 = blockIdx.x * RY;
    for (int j = blockIdx.y * blockDim.x + threadIdx.x; j <= wq; j += gridDim.y * blockDim.x) {
        uint32_t lo[RY][3], hi[RY][3];
#pragma unroll
        for (int r = 0; r < RY; ++r) {
            const int row = row0 + r, n = row / H, y = row - n * H;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                lo[r][ch] = hi[r][ch] = 0u;
                if (ch < C && row < total_rows) {
                    const uint32_t *src = reinterpret_cast<const uint32_t *>(in + ((size_t)n * C + ch) * HW + (size_t)y * W);
                    if (j > 0) lo[r][ch] = __ldg(src + j - 1);
                    if (j < wq) hi[r][ch] = __ldg(src + j);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RY; ++r) {
            const int row = row0 + r, n = row / H, y = row - n * H;
            if (row >= total_rows) break;
            uint32_t pl[4] = {0, 0, 0, 0};
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) pl[ch] = __byte_perm(lo[r][ch], hi[r][ch], 0x6543);   // pixels 4j-1, 4j, 4j+1, 4j+2 of this plane
            if (C == 4) {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(in + ((size_t)n * C + 3) * HW + (size_t)y * W);
                pl[3] = __byte_perm(j > 0 ? __ldg(src + j - 1) : 0u, j < wq ? __ldg(src + j) : 0u, 0x6543);
            }
            uint4 o;   // pixel k = byte k of every plane word
            o.x = __byte_perm(__byte_perm(pl[0], pl[1], 0x0040), __byte_perm(pl[2], pl[3], 0x0040), 0x5410);
            o.y = __byte_perm(__byte_perm(pl[0], pl[1], 0x0051), __byte_perm(pl[2], pl[3], 0x0051), 0x5410);
            o.z = __byte_perm(__byte_perm(pl[0], pl[1], 0x0062), __byte_perm(pl[2], pl[3], 0x0062), 0x5410);
            o.w = __byte_perm(__byte_perm(pl[0], pl[1], 0x0073), __byte_perm(pl[2], pl[3], 0x0073), 0x5410);
            uint32_t *dst = reinterpret_cast<uint32_t *>(out) + ((size_t)n * rows_h + y + 1) * pitch_w + 4 * j;   // pixel 4j-1 sits at column 4j
            if (j == 0) {
                dst[1] = o.y;
                *reinterpret_cast<uint2 *>(dst + 2) = make_uint2(o.z, o.w);
            } else if (j == wq) {
                dst[0] = o.x;
            } else {
                *reinterpret_cast<uint4 *>(dst) = o;
            }
        }
    }
}

__global__ void nhwc_to_nchw_u8_geom_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int C, int H, int W, int CS, int pad,
                                            int pitch_w, int rows_h, long long total)
{
    const int HW = H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int p = (int)(i % HW);
        long long g = i / HW;
        int ch = (int)(g % C);
        long long n = g / C;
        const int y = p / W, x = p - y * W;
        out[i] = in[(((size_t)n * rows_h + y + pad) * pitch_w + x + pad) * CS + ch];
    }
}


extern "C" int yq_nchw_to_nhwc_u8_geom(const uint8_t *in, uint8_t *out, int batch, int c, int h, int w, const yq_act_geom *g, void *stream)
{
    if (!in || !out) return yq::fail("nchw_to_nhwc: null pointer");
    if (!g || check_geom(g, h, w)) return -1;
    const int cs = yq::channel_stride(c);
    if (cs == 4 && w % 4 == 0 && ((uintptr_t)in % 4) == 0 && g->pad == 1 && g->pitch_w % 4 == 0 && ((uintptr_t)out % 16) == 0) {
        constexpr int RY = 4;
        const int cpr = w / 4 + 1, threads = cpr >= 128 ? 128 : (cpr + 31) / 32 * 32;
        const int by = (cpr + threads - 1) / threads;
        dim3 grid((unsigned)((batch * h + RY - 1) / RY), (unsigned)(by < 8 ? by : 8));
        YQ_CUDA(yq::launch_pdl(nchw_to_nhwc4_u8_geom_kernel<RY>, grid, dim3(threads), 0, (cudaStream_t)stream, in, out, c, h, w, g->pitch_w, g->rows_h, batch * h));
        YQ_CHECK_LAUNCH();
        return 0;
    }
    long long total = (long long)batch * h * w * (cs / 4);
    nchw_to_nhwc_u8_geom_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, c, h, w, cs, g->pad, g->pitch_w, g->rows_h, total);
    YQ_CHECK_LAUNCH();
    return 0;
}
extern "C" int yq_nhwc_to_nchw_u8_geom(const uint8_t *in, uint8_t *out, int batch, int c, int h, int w, const yq_act_geom *g, void *stream)
{
    if (!in || !out) return yq::fail("nhwc_to_nchw: null pointer");
    if (!g || check_geom(g, h, w)) return -1;
    long long total = (long long)batch * c * h * w;
    nhwc_to_nchw_u8_geom_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, c, h, w, yq::channel_stride(c), g->pad, g->pitch_w,
                                                                                        g->rows_h, total);
    YQ_CHECK_LAUNCH();
    return 0;
}

extern "C" int yq_nchw_to_nhwc_u8(const uint8_t *in, uint8_t *out, int batch, int c, int h, int w, void *stream)
{
    if (!in || !out) return yq::fail("nchw_to_nhwc: null pointer");
    const int cs = yq::channel_stride(c);
    if (cs == 4 && (h * w) % 4 == 0 && ((uintptr_t)in % 4) == 0) {
        long long total4 = (long long)batch * h * w / 4;
        nchw_to_nhwc4_u8_kernel<<<grid_for(total4, 256), 256, 0, (cudaStream_t)stream>>>(in, out, c, h * w, total4);
        YQ_CHECK_LAUNCH();
        return 0;
    }
    long long total = (long long)batch * h * w * (cs / 4);
    nchw_to_nhwc_u8_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, c, h * w, cs, total);
    YQ_CHECK_LAUNCH();
    return 0;
}
extern "C" int yq_nhwc_to_nchw_u8(const uint8_t *in, uint8_t *out, int batch, int c, int h, int w, void *stream)
{
    if (!in || !out) return yq::fail("nhwc_to_nchw: null pointer");
    long long total = (long long)batch * c * h * w;
    nhwc_to_nchw_kernel<uint8_t><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, c, h * w, yq::channel_stride(c), total);
    YQ_CHECK_LAUNCH();
    return 0;
}
extern "C" int yq_nhwc_to_nchw_i32(const int32_t *in, int32_t *out, int batch, int c, int h, int w, void *stream)
{
    if (!in || !out) return yq::fail("nhwc_to_nchw: null pointer");
    long long total = (long long)batch * c * h * w;
    nhwc_to_nchw_kernel<int32_t><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, c, h * w, yq::channel_stride(c), total);
    YQ_CHECK_LAUNCH();
    return 0;
}
