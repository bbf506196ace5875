// yq_epilogue.cuh -- the fused requantize / bias / activation / zero-point / uint8 epilogue shared by the
// convolution flavours.  Restates src/convolutional_layer.c:726-760 of the reference:
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
// Two exact forms are used on the device:
//
//  (A) FP64 form.  M0_right_shift_value is 2^-s, so trunc(trunc(y) * 2^-s) == trunc(y * 2^-s) and
//      RN(x*Mv) * 2^-s == RN(x * (Mv*2^-s)); with every rshift a power of two (always true for the reference's
//      prep, blas.c:315) one DMUL + one truncating convert reproduces the reference bit for bit.  The SIMT
//      flavour uses it for every output (and keeps the literal two-step form when a binding hands in
//      something that is not a power of two).
//
//  (B) Integer form (tcgen05 flavours' hot path).  M_value = M0 * 2^-31 with integer M0 in [2^30, 2^31).
//      While |x| < 2^22 the product x*M0 is below 2^53, the double multiply is exact, and
//          q = sign(x) * ( umulhi(|x|, 2*M0) >> s )       ( = trunc(x * M0 / 2^(31+s)) )
//      is the same number with two integer instructions and no conversion-unit traffic.  Each thread tracks
//      max|x| over a chunk of outputs and re-does the chunk with form (A) in the (rare) case the bound is
//      exceeded, so the result is bit-identical for every input.
//      LEAKY's round(q*0.1) for q < 0 equals -((|q|+5)/10) exactly (the double product is q/10 + O(2^-22)
//      and ties round away from zero either way), i.e. umulhi(|q|+5, 0xCCCCCCCD) >> 3 -- or, while |q| + 5 < 81920,
//      (|q| * 52429 + 262145) >> 19: one multiply-add on the light multiplier pipe instead of an add and a 64-bit-product multiply.
//      make_epi() lowers xlim for LEAKY layers so that |x| < xlim implies that range; beyond it the chunk is redone in form (A).
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
    const int4 *chanq;       // {bias, zw, 2*M0, shift} per channel (integer form), or nullptr
    int fused, act, zp_out, saturate;
    float s_out;
    uint32_t xlim;           // the integer form is the reference's double arithmetic for every |x| < xlim (see requant_exact_limit)
};

struct ChanParams {
    int bias, zw;
    double m0, m1;
};

// Largest power of two P such that for every channel the reference's double product x * M_value (convolutional_layer.c:732) is
// EXACT for all |x| < P, i.e. |x| * M0 has at most 53 significant bits.  M0 = round(M * 2^31) of a FLOAT M (blas.c:313-316,
// :387-418) carries at most 24 significant bits -- at least 7 trailing zeros -- so for the reference's own host prep P = 2^29 and
// the FP64 re-do below never runs (|acc| <= K * 255 * 255 < 2^29 for K <= 8256); a binding that hands in a full 31-bit M0 gets the
// 2^22 of the worst case.  A power of two so that kernels may test an OR of magnitudes instead of their maximum.
static inline uint32_t requant_exact_limit(const yq_conv_layer *l)
{
    uint32_t lim = 1u << 31;
    for (int oc = 0; oc < l->n; ++oc) {
        uint32_t m = (uint32_t)l->host_chanq[(size_t)oc * 4 + 2] >> 1;      // chanq.z = 2 * M0
        if (!m) continue;
        while (!(m & 1u)) m >>= 1;                                           // significant bits of M0
        while (lim > 1 && (unsigned long long)(lim - 1) * m >= (1ull << 53)) lim >>= 1;
    }
    return lim;
}

// LEAKY: the epilogue divides h = trunc(|x| * M) by ten as (h * 52429 + 262145) >> 19, which is floor((h + 5) / 10) while h + 5 < 81920.
// Largest power of two P <= lim such that |x| < P keeps every channel's h = (|x| * 2 M0) >> (32 + s) at or below 81914.
static inline uint32_t leaky_div10_limit(const yq_conv_layer *l, uint32_t lim)
{
    for (int oc = 0; oc < l->n; ++oc) {
        const unsigned __int128 z = (uint32_t)l->host_chanq[(size_t)oc * 4 + 2];
        const int w = l->host_chanq[(size_t)oc * 4 + 3];
        if (!z) continue;
        const unsigned __int128 top = ((unsigned __int128)81915 << (32 + w)) - 1;      // (|x| * z) >> (32 + w) <= 81914  <=>  |x| * z <= top
        while (lim > 1 && (unsigned __int128)(lim - 1) * z > top) lim >>= 1;
    }
    return lim;
}

static inline EpiParams make_epi(const yq_conv_layer *l)
{
    EpiParams e;
    e.xlim = l->int_form ? requant_exact_limit(l) : (1u << 22);
    if (l->int_form && l->activation == YQ_LEAKY) e.xlim = leaky_div10_limit(l, e.xlim);
    e.bias = l->bias; e.zw = l->zw; e.mcomb = l->mcomb; e.mval = l->mval; e.rsh = l->rsh; e.chanq = (const int4 *)l->chanq;
    e.fused = l->fused_mult; e.act = l->activation; e.zp_out = l->zp_out; e.saturate = l->saturate; e.s_out = l->s_out;
    return e;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// form (A): per-output FP64 (SIMT flavour, slow path of the tcgen05 flavours)
// ---------------------------------------------------------------------------------------------
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

// ---------------------------------------------------------------------------------------------
// form (B): chunked integer epilogue for the tcgen05 flavours
// ---------------------------------------------------------------------------------------------
// One chunk of NV consecutive output channels of ONE pixel.
//   v[j]   raw tensor-core accumulator  sum_k w*a   (uint8 x uint8, zero-filled padding)
//   nsa    minus the pixel's activation sum          (so  acc = v + zw * nsa)
//   extra  per-output additive correction (border taps) or nullptr
//   cq     shared-memory {bias, zw, 2*M0, shift} of the chunk's first channel; mc = matching M_value*2^-s doubles
// Writes NV/4 packed little-endian words.  Returns nothing; bit-exact for all inputs (see header).
// four low bytes -> one little-endian word (PRMT takes the LOW byte of each value: the uint8 wrap comes for free)
__device__ __forceinline__ uint32_t pack_low_bytes(int r0, int r1, int r2, int r3)
{
    return __byte_perm(__byte_perm((uint32_t)r0, (uint32_t)r1, 0x0040), __byte_perm((uint32_t)r2, (uint32_t)r3, 0x0040), 0x5410);
}

// ACTM: 0 = RELU6, 1 = LINEAR / RELU, 2 = LEAKY.  SAT: clamp instead of wrap.  Value before the uint8 store.
template <int ACTM, bool SAT>
__device__ __forceinline__ int act_value(int q, int zo)
{
    int r;
    if (ACTM == 2 && q < 0) r = zo - (int)(__umulhi((uint32_t)(-q) + 5u, 0xCCCCCCCDu) >> 3);
    else r = q + zo;                                   // (RELU6 callers pass q >= 0 with q == 0 for every x <= 0)
    if (SAT) r = max(0, min(255, r));
    return r;
}

// the chunk's values BEFORE the uint8 store (the caller packs their low bytes, or feeds them to a fused quantized shortcut)
template <int ACTM, bool SAT, int NV, bool HAS_EXTRA>
__device__ __forceinline__ void requant_chunk_vals(const uint32_t (&v)[NV], int nsa, const int (&extra)[NV], const int4 *cq, const double *mc,
                                                   int zo, int (&r)[NV], uint32_t xlim)
{
    // xlim is a power of two: the OR of the magnitudes reaches it exactly when one of them does (two per LOP3 instead of one VIMNMX each)
    uint32_t mx = 0;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int4 c = cq[j];
        int x = c.y * nsa + (int)v[j];
        if (HAS_EXTRA) x += extra[j];
        x += c.x;
        int q;
        if (ACTM == 0) {
            const uint32_t xp = (uint32_t)max(x, 0);
            mx |= xp;
            q = (int)(__umulhi(xp, (uint32_t)c.z) >> c.w);
        } else {
            const uint32_t ax = (uint32_t)abs(x);
            mx |= ax;
            const int h = (int)(__umulhi(ax, (uint32_t)c.z) >> c.w);
            if (ACTM == 2) {
                // LEAKY from the magnitude: x < 0 -> zo - round(h / 10) (half away from zero), else zo + h; one select, no second sign test
                const int t = (int)(((uint32_t)h * 52429u + 262145u) >> 19);      // floor((h + 5) / 10): h <= 81914 whenever |x| < xlim (make_epi)
                int rr = x < 0 ? zo - t : zo + h;
                if (SAT) rr = max(0, min(255, rr));
                r[j] = rr;
                continue;
            }
            q = x < 0 ? -h : h;
        }
        r[j] = act_value<ACTM, SAT>(q, zo);
    }
    if (mx >= xlim) {
        // |x*M0| may need more than 53 bits: the reference's double multiply rounds -> redo this chunk in FP64 form (A)
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int4 c = cq[j];
            int x = c.y * nsa + (int)v[j];
            if (HAS_EXTRA) x += extra[j];
            x += c.x;
            int q = __double2int_rz(__dmul_rn((double)x, mc[j]));
            if (ACTM == 0) q = max(q, 0);
            r[j] = act_value<ACTM, SAT>(q, zo);
        }
    }
}

template <int ACTM, bool SAT, int NV, bool HAS_EXTRA>
__device__ __forceinline__ void requant_chunk(const uint32_t (&v)[NV], int nsa, const int (&extra)[NV], const int4 *cq, const double *mc,
                                              int zo, uint32_t (&packed)[NV / 4], uint32_t xlim)
{
    int r[NV];
    requant_chunk_vals<ACTM, SAT, NV, HAS_EXTRA>(v, nsa, extra, cq, mc, zo, r, xlim);
#pragma unroll
    for (int k = 0; k < NV / 4; ++k) packed[k] = pack_low_bytes(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
}

// The quantized shortcut (extension, include/yq_b200.h) fused behind a convolution's requantisation: a = the convolution's uint8
// result (the low byte of r: the reference's store wraps), b = the byte of the `from` tensor at the same position / channel:
//     out = sat_u8((a * Ka + b * Kb + C0) >> 16),   C0 = 2^15 + (zp_out << 16) - zp_a * Ka - zp_b * Kb
// 16 channels: r[16] and the 16 residual bytes (one 16-byte load) -> 4 packed words.
struct ShortcutParams {
    int Ka, Kb, C0;
};
__device__ __forceinline__ uint32_t shortcut_word(int r0, int r1, int r2, int r3, uint32_t b, const ShortcutParams &sp)
{
    const int t0 = (r0 & 0xff) * sp.Ka + (int)(b & 0xff) * sp.Kb + sp.C0;
    const int t1 = (r1 & 0xff) * sp.Ka + (int)((b >> 8) & 0xff) * sp.Kb + sp.C0;
    const int t2 = (r2 & 0xff) * sp.Ka + (int)((b >> 16) & 0xff) * sp.Kb + sp.C0;
    const int t3 = (r3 & 0xff) * sp.Ka + (int)(b >> 24) * sp.Kb + sp.C0;
    uint32_t lo, hi;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(t3 >> 16), "r"(t2 >> 16));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(t1 >> 16), "r"(t0 >> 16), "r"(hi));
    return lo;
}
__device__ __forceinline__ void shortcut_pack16(const int (&r)[16], const uint4 &b, const ShortcutParams &sp, uint32_t (&packed)[4])
{
    packed[0] = shortcut_word(r[0], r[1], r[2], r[3], b.x, sp);
    packed[1] = shortcut_word(r[4], r[5], r[6], r[7], b.y, sp);
    packed[2] = shortcut_word(r[8], r[9], r[10], r[11], b.z, sp);
    packed[3] = shortcut_word(r[12], r[13], r[14], r[15], b.w, sp);
}


// zero the bytes of pad channels (channel index >= n_real within this chunk): pad lanes of an activation
// tensor must be 0 because consumers sum every byte of a pixel
template <int NV>
__device__ __forceinline__ void mask_pad_channels(uint32_t (&packed)[NV / 4], int n_real)
{
    if (n_real >= NV) return;
#pragma unroll
    for (int k = 0; k < NV / 4; ++k) {
        const int left = n_real - 4 * k;
        const uint32_t m = left >= 4 ? 0xffffffffu : (left <= 0 ? 0u : (0xffffffffu >> (8 * (4 - left))));
        packed[k] &= m;
    }
}

__host__ __device__ constexpr int act_mode(int act) { return act == YQ_RELU6 ? 0 : (act == YQ_LEAKY ? 2 : 1); }
#endif

}  // namespace yq
