// yq_input.cu -- the layer-0 dynamic input quantiser on the device.
//
// Restates quant_weights_with_min_max_channel with size_channel = 1 (src/blas.c:108-168), which
// quantization_weights_and_activations runs on the float image before every forward (src/blas.c:279):
//     mn = min(0, min x)   mx = max(0, max x)   s = (mx - mn) / 255
//     zp = clamp(round(0 - mn / s), 0, 255)     u8 = clamp(round(x / s) + zp, 0, 255)      (this one saturates, :158)
// per IMAGE (the reference is batch-1: net->c*net->w*net->h elements).  Float operations are IEEE single precision in
// the reference's order (division, not reciprocal), round() is half-away-from-zero.
#include <cuda_runtime.h>
#include <math.h>

#include "yq_common.h"

namespace {

// pass 1: per-image max(x) and max(-x) (both >= 0, so their float bit patterns order like ints)
__global__ void input_minmax_kernel(const float *__restrict__ x, int n, int *__restrict__ mm /* [batch][2], zero-initialised */)
{
    const float *img = x + (size_t)blockIdx.y * n;
    float mx = 0.f, mneg = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float v = __ldg(img + i);
        mx = fmaxf(mx, v);
        mneg = fmaxf(mneg, -v);
    }
    for (int o = 16; o; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mneg = fmaxf(mneg, __shfl_xor_sync(0xffffffffu, mneg, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&mm[2 * blockIdx.y], __float_as_int(mx));
        atomicMax(&mm[2 * blockIdx.y + 1], __float_as_int(mneg));
    }
}

// pass 2: scale / zero point per image (recomputed by every block from the two floats), then the bytes
__global__ void input_quant_kernel(const float *__restrict__ x, uint8_t *__restrict__ out, int n, const int *__restrict__ mm, float *__restrict__ scales,
                                   int *__restrict__ zps)
{
    const int b = blockIdx.y;
    const float mx = __int_as_float(mm[2 * b]), mn = -__int_as_float(mm[2 * b + 1]);
    const float s = __fdiv_rn(__fsub_rn(mx, mn), 255.f);                    // blas.c:137
    const double izp = (double)__fsub_rn(0.f, __fdiv_rn(mn, s));            // :139 (float arithmetic, then widened)
    const int z = izp < 0. ? 0 : (izp > 255. ? 255 : (int)round(izp));     // :144-150
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        scales[b] = s;
        zps[b] = (mx == 0.f && mn == 0.f) ? -1 : z;                         // the reference assert()s on an all-zero image (:124-127)
    }
    const float *img = x + (size_t)b * n;
    uint8_t *o = out + (size_t)b * n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float t = roundf(__fdiv_rn(__ldg(img + i), s)) + (float)z;   // :154
        const int q = (int)t;
        o[i] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));                  // clamp(), :158
    }
}

// letterbox_image (src/image.c:812-831) = resize_image (:1199-1245, two separable passes of linear interpolation) into a
// w x h canvas filled with .5 (fill_image :695), embedded at ((w - new_w) / 2, (h - new_h) / 2) (embed_image :428-439).
// One thread per canvas pixel; the float operations are the reference's, in its order, each rounded separately
// (no fused multiply-add: the reference's horizontal pass stores `part` before the vertical pass reads it):
//     part(c, r)  = c == new_w - 1 || iw == 1 ? im(iw - 1, r) : (1 - dx) * im(ix, r) + dx * im(ix + 1, r),  sx = c * w_scale
//     res(c, r)   = (1 - dy) * part(c, iy)  [+ dy * part(c, iy + 1) unless r == new_h - 1 || ih == 1],        sy = r * h_scale
__device__ __forceinline__ float lb_part(const float *__restrict__ plane, int iw, int new_w, float w_scale, int c, int r)
{
    if (c == new_w - 1 || iw == 1) return __ldg(plane + (size_t)r * iw + iw - 1);
    const float sx = __fmul_rn((float)c, w_scale);
    const int ix = (int)sx;
    const float dx = __fsub_rn(sx, (float)ix);
    return __fadd_rn(__fmul_rn(__fsub_rn(1.f, dx), __ldg(plane + (size_t)r * iw + ix)), __fmul_rn(dx, __ldg(plane + (size_t)r * iw + ix + 1)));
}

__global__ void letterbox_kernel(const float *__restrict__ in, float *__restrict__ out, int ch, int ih, int iw, int h, int w, int new_h, int new_w, int off_y,
                                 int off_x, float w_scale, float h_scale)
{
    const size_t total = (size_t)ch * h * w;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % w), y = (int)((i / w) % h), k = (int)(i / ((size_t)w * h));
        const float *src = in + (size_t)blockIdx.y * ch * ih * iw + (size_t)k * ih * iw;
        const int c = x - off_x, r = y - off_y;
        float val = .5f;
        if (c >= 0 && c < new_w && r >= 0 && r < new_h) {
            const float sy = __fmul_rn((float)r, h_scale);
            const int iy = (int)sy;
            const float dy = __fsub_rn(sy, (float)iy);
            val = __fmul_rn(__fsub_rn(1.f, dy), lb_part(src, iw, new_w, w_scale, c, iy));
            if (!(r == new_h - 1 || ih == 1)) val = __fadd_rn(val, __fmul_rn(dy, lb_part(src, iw, new_w, w_scale, c, iy + 1)));
        }
        out[(size_t)blockIdx.y * total + i] = val;
    }
}

}  // namespace

// letterbox_image (src/image.c:812-831) for a batch of float CHW images of one size, on the device
extern "C" int yq_letterbox_image_gpu(const float *in_chw, int batch, int c, int ih, int iw, float *out_chw, int h, int w, void *stream)
{
    if (!in_chw || !out_chw || batch <= 0 || c <= 0 || ih <= 0 || iw <= 0 || h <= 0 || w <= 0) return yq::fail("yq_letterbox_image_gpu: bad argument");
    int new_w = iw, new_h = ih;
    if (((float)w / iw) < ((float)h / ih)) {          // image.c:816-822
        new_w = w;
        new_h = (ih * w) / iw;
    } else {
        new_h = h;
        new_w = (iw * h) / ih;
    }
    if (new_w < 1 || new_h < 1) return yq::fail("yq_letterbox_image_gpu: %dx%d does not fit a %dx%d canvas", iw, ih, w, h);
    const float w_scale = (float)(iw - 1) / (new_w - 1), h_scale = (float)(ih - 1) / (new_h - 1);   // image.c:1204-1205 (inf / nan when new == 1: unused then)
    const size_t total = (size_t)c * h * w;
    unsigned bx = (unsigned)((total + 255) / 256);
    if (bx > 148 * 8) bx = 148 * 8;
    letterbox_kernel<<<dim3(bx, (unsigned)batch), 256, 0, (cudaStream_t)stream>>>(in_chw, out_chw, c, ih, iw, h, w, new_h, new_w, (h - new_h) / 2, (w - new_w) / 2, w_scale,
                                                                                   h_scale);
    YQ_CHECK_LAUNCH();
    return 0;
}

extern "C" int yq_quantize_input_gpu(const float *in_f32, uint8_t *out_u8, float *scales, int *zero_points, int *scratch, int batch, int n, void *stream)
{
    if (!in_f32 || !out_u8 || !scales || !zero_points || !scratch || batch <= 0 || n <= 0) return yq::fail("yq_quantize_input_gpu: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    YQ_CUDA(cudaMemsetAsync(scratch, 0, sizeof(int) * 2 * (size_t)batch, st));
    int bx = (n + 256 * 8 - 1) / (256 * 8);
    if (bx > 64) bx = 64;
    dim3 grid((unsigned)bx, (unsigned)batch);
    input_minmax_kernel<<<grid, 256, 0, st>>>(in_f32, n, scratch);
    YQ_CHECK_LAUNCH();
    input_quant_kernel<<<grid, 256, 0, st>>>(in_f32, out_u8, n, scratch, scales, zero_points);
    YQ_CHECK_LAUNCH();
    return 0;
}
