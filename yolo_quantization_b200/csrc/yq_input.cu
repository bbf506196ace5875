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

}  // namespace

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
