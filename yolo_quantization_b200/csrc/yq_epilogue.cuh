// yq_epilogue.cuh -- the fused requantize / bias / activation / zero-point / uint8 epilogue shared by
// both convolution flavours.  Restates src/convolutional_layer.c:726-760 of the reference:
//
//   x = acc + biases_int32[oc]                                  (int32 add)
//   t = (int64) trunc( (double)x * M_value[oc] )                (ONE IEEE double multiply, RN, then trunc)
//   q = (int32) trunc( (double)t * M0_right_shift_value[oc] )
//   LEAKY : q < 0 ? (int)(round(q*0.1) + zp_out) : q + zp_out   (double round-half-away)
//   RELU6 : q <= 0 ? zp_out : q + zp_out                        (no upper clip)
//   LINEAR, RELU : q + zp_out
//   uint8 store WRAPS mod 256 (the clamp on :749 acts on a uint8_t and is a no-op)
//   quant_stop : f32 = (float)((int)u8 - zp_out) * s_out
//
// Fused multiplier: M0_right_shift_value is 2^-s, so trunc(trunc(y) * 2^-s) == trunc(y * 2^-s) and
// RN(x*Mv) * 2^-s == RN(x * (Mv*2^-s)) exactly; when the host verified that every rshift is a power of
// two (always true for the reference's prep, blas.c:315) the epilogue does one multiply + one convert.
#pragma once
#include <stdint.h>

#include "yq_common.h"

namespace yq {

struct EpiParams {
    const int32_t *bias;
    const int32_t *zw;
    const double *mcomb;
    const double *mval;
    const double *rsh;
    int fused, act, zp_out, saturate;
    float s_out;
};

struct ChanParams {
    int bias, zw;
    double m0, m1;
};

static inline EpiParams make_epi(const yq_conv_layer *l)
{
    EpiParams e;
    e.bias = l->bias; e.zw = l->zw; e.mcomb = l->mcomb; e.mval = l->mval; e.rsh = l->rsh;
    e.fused = l->fused_mult; e.act = l->activation; e.zp_out = l->zp_out; e.saturate = l->saturate; e.s_out = l->s_out;
    return e;
}

#ifdef __CUDACC__
__device__ __forceinline__ ChanParams load_chan(const EpiParams &e, int oc)
{
    ChanParams c;
    c.bias = __ldg(e.bias + oc);
    c.zw = __ldg(e.zw + oc);
    if (e.fused) {
        c.m0 = __ldg(e.mcomb + oc);
        c.m1 = 1.0;
    } else {
        c.m0 = __ldg(e.mval + oc);
        c.m1 = __ldg(e.rsh + oc);
    }
    return c;
}

__device__ __forceinline__ int requant_q(const EpiParams &e, const ChanParams &c, int acc)
{
    const int x = acc + c.bias;
    if (e.fused) return __double2int_rz(__dmul_rn((double)x, c.m0));
    const long long t = __double2ll_rz(__dmul_rn((double)x, c.m0));
    return __double2int_rz(__dmul_rn((double)t, c.m1));
}

__device__ __forceinline__ uint8_t requant_u8(const EpiParams &e, const ChanParams &c, int acc)
{
    const int q = requant_q(e, c, acc);
    int r;
    if (e.act == YQ_RELU6) {
        r = q <= 0 ? e.zp_out : q + e.zp_out;
    } else if (e.act == YQ_LEAKY) {
        r = q < 0 ? __double2int_rz(round((double)q * 0.1) + (double)e.zp_out) : q + e.zp_out;
    } else {
        r = q + e.zp_out;
    }
    if (e.saturate) r = r < 0 ? 0 : (r > 255 ? 255 : r);
    return (uint8_t)r;
}

__device__ __forceinline__ float dequant_f32(const EpiParams &e, uint8_t u8)
{
    return __fmul_rn((float)((int)u8 - e.zp_out), e.s_out);
}
#endif

}  // namespace yq
