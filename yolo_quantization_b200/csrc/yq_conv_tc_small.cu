// yq_conv_tc_small.cu -- tcgen05 (kind::i8) convolution for SMALL input-channel counts (c <= 32: layers 0, 2, 4
// of yolov3-tiny: c = 3, 16, 32), 3x3 filters.
//
// These layers have tiny K (27 / 144 / 288) and huge pixel counts: the tensor-core time is negligible and a
// TMA box per tap would move 4..32-byte rows.  Instead each of the 128 threads of a CTA builds ONE im2col row
// (one output pixel: 9 taps x CS bytes, zero-padded to a multiple of 32) straight into shared memory in the
// canonical K-major SWIZZLE_32B layout the tensor core reads; out-of-image taps are written as zp_in
// (im2col.c:5-14), so no border correction is needed.  The whole filter bank (<= 18 KB, pre-swizzled on the host)
// stays resident in shared memory; CTAs are persistent and stride over 128-pixel tiles.
//
//   per tile:  build A rows (LDG -> STS) | fence.proxy.async + bar | one thread issues K/32 x {main MMA, ones MMA}
//              | tcgen05.commit -> mbarrier | every thread drains its own TMEM lane (= its pixel):
//              acc = sum w*a - zp_w * sum a -> FP64 requant -> activation -> +zp_out -> uint8 wrap -> 16-byte stores
//   128 consecutive pixels x N channels are contiguous in the NHWC output, so stores are plain coalesced STG.128.
//   Overlap comes from several co-resident CTAs per SM (each owns BN+16 TMEM columns).
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include <vector>

#include "yq_common.h"
#include "yq_epilogue.cuh"
#include "yq_tc_ptx.cuh"

using namespace yqtc;

namespace {

constexpr int SM_THREADS = 128;
constexpr int TAPS = 9;

struct SmallArgs {
    const uint8_t *in;
    uint8_t *out;
    float *out_f32;
    int32_t *out_acc;
    const uint8_t *wimg;   // pre-swizzled shared-memory image of the filter bank
    yq::EpiParams ep;
    int B, H, W, C, OH, OW, N, CSO, stride, pad, zp_in, M_total, num_tiles;
};

template <int CS>
struct SmallGeom {
    static constexpr int K = TAPS * CS;
    static constexpr int KPAD = (K + 31) / 32 * 32;
    static constexpr int P = KPAD / 32;          // 32-byte K panels = MMA K steps
    static constexpr int A_BYTES = P * 128 * 32;
};

template <int CS, int BN>
struct SmallSmem {
    using G = SmallGeom<CS>;
    static constexpr int A_OFF = 0;
    static constexpr int B_OFF = G::A_BYTES;
    static constexpr int B_BYTES = G::P * BN * 32;
    static constexpr int ONES_OFF = B_OFF + B_BYTES;
    static constexpr int PARAM_OFF = ONES_OFF + 512;
    static constexpr int BAR_OFF = PARAM_OFF + BN * 24;
    static constexpr int TOTAL = BAR_OFF + 64;
};

template <int BN>
__host__ __device__ constexpr int small_tmem_cols() { return BN + 16 <= 32 ? 32 : (BN + 16 <= 64 ? 64 : 128); }

template <int NV>
__device__ __forceinline__ void epi_chunk_small(int actm, int sat, const uint32_t (&v)[NV], int nsa, const int4 *cq, const double *mc, int zo,
                                                uint32_t (&packed)[NV / 4])
{
    int extra[NV];   // unused (HAS_EXTRA = false): padded taps already carry zp_in
    if (sat) {
        if (actm == 0) yq::requant_chunk<0, true, NV, false>(v, nsa, extra, cq, mc, zo, packed);
        else if (actm == 1) yq::requant_chunk<1, true, NV, false>(v, nsa, extra, cq, mc, zo, packed);
        else yq::requant_chunk<2, true, NV, false>(v, nsa, extra, cq, mc, zo, packed);
    } else {
        if (actm == 0) yq::requant_chunk<0, false, NV, false>(v, nsa, extra, cq, mc, zo, packed);
        else if (actm == 1) yq::requant_chunk<1, false, NV, false>(v, nsa, extra, cq, mc, zo, packed);
        else yq::requant_chunk<2, false, NV, false>(v, nsa, extra, cq, mc, zo, packed);
    }
}

template <int CS, int BN>
__global__ void __launch_bounds__(SM_THREADS, 4) conv_u8_tc_small_kernel(const SmallArgs a)
{
    using G = SmallGeom<CS>;
    using L = SmallSmem<CS, BN>;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    uint8_t *sA = smem + L::A_OFF;
    uint8_t *sB = smem + L::B_OFF;
    int4 *s_q = (int4 *)(smem + L::PARAM_OFF);          // {bias, zw, 2*M0, shift}
    double *s_mc = (double *)(s_q + BN);                // M_value * 2^-s (FP64 slow path)
    uint64_t *mma_done = (uint64_t *)(smem + L::BAR_OFF);
    uint32_t *tmem_slot = (uint32_t *)(mma_done + 1);

    const int r = threadIdx.x;            // A-tile row == TMEM lane == pixel within the tile
    const int warp = r >> 5;

    // ---- one-time setup: resident filter bank, ones tile, per-channel parameters, barrier, TMEM
    for (int i = r; i < L::B_BYTES / 16; i += SM_THREADS)
        reinterpret_cast<uint4 *>(sB)[i] = __ldg(reinterpret_cast<const uint4 *>(a.wimg) + i);
    for (int i = r; i < 512 / 4; i += SM_THREADS) reinterpret_cast<uint32_t *>(smem + L::ONES_OFF)[i] = 0x01010101u;
    for (int i = r; i < BN; i += SM_THREADS) {
        s_q[i] = __ldg(a.ep.chanq + i);
        s_mc[i] = __ldg(a.ep.mcomb + i);
    }
    if (r == 0) {
        mbar_init(mma_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc<small_tmem_cols<BN>()>(tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);

    // zp_in in the real channels, 0 in the pad channels (pad channels must stay 0: their weights are 0 but the
    // activation sum counts every byte)
    auto fill_word = [&](int ch0) -> uint32_t {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (ch0 + b < a.C) v |= (uint32_t)a.zp_in << (8 * b);
        return v;
    };
    const int actm = yq::act_mode(a.ep.act);
    const int swz = (r >> 2) & 1;          // SWIZZLE_32B: 16-byte chunk index ^= address bit 7 = (row >> 2) & 1
    uint32_t phase = 0;

    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int m = tile * 128 + r;
        const bool valid = m < a.M_total;
        int n = 0, oy = 0, ox = 0;
        if (valid) {
            n = m / (a.OH * a.OW);
            const int rem = m - n * a.OH * a.OW;
            oy = rem / a.OW;
            ox = rem - oy * a.OW;
        }
        const int iy0 = oy * a.stride - a.pad, ix0 = ox * a.stride - a.pad;
        const uint8_t *img = a.in + (size_t)n * a.H * a.W * CS;

        // ---- build this pixel's im2col row: chunk g (16 bytes) of the K axis -> panel g/2, half g%2
#pragma unroll
        for (int g = 0; g < 2 * G::P; ++g) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (valid) {
                if (CS == 4) {
                    uint32_t wv[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int tap = 4 * g + j;
                        if (tap < TAPS) {
                            const int iy = iy0 + tap / 3, ix = ix0 + tap % 3;
                            wv[j] = (iy < 0 || iy >= a.H || ix < 0 || ix >= a.W)
                                        ? fill_word(0)
                                        : __ldg(reinterpret_cast<const uint32_t *>(img + ((size_t)iy * a.W + ix) * 4));
                        }
                    }
                    v = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                } else {
                    constexpr int CPT = CS / 16;   // 16-byte chunks per tap
                    const int tap = g / CPT, sub = g % CPT;
                    if (tap < TAPS) {
                        const int iy = iy0 + tap / 3, ix = ix0 + tap % 3;
                        if (iy < 0 || iy >= a.H || ix < 0 || ix >= a.W)
                            v = make_uint4(fill_word(sub * 16), fill_word(sub * 16 + 4), fill_word(sub * 16 + 8), fill_word(sub * 16 + 12));
                        else
                            v = __ldg(reinterpret_cast<const uint4 *>(img + ((size_t)iy * a.W + ix) * CS + sub * 16));
                    }
                }
            }
            *reinterpret_cast<uint4 *>(sA + (g >> 1) * 4096 + r * 32 + (((g & 1) ^ swz) << 4)) = v;
        }
        fence_proxy_async();     // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncthreads();

        // ---- one thread issues the MMAs for the whole tile
        if (r == 0) {
            tc_fence_after();
            constexpr uint32_t idesc_main = make_idesc(BN);
            constexpr uint32_t idesc_ones = make_idesc(16);
            const uint64_t d_ones = make_desc<32>(smem_u32(smem + L::ONES_OFF));
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
            for (int p = 0; p < G::P; ++p) {
                const uint64_t da = make_desc<32>(a0 + p * 4096);
                umma_i8(tmem_base, da, make_desc<32>(b0 + p * BN * 32), idesc_main, p ? 1u : 0u);
                umma_i8(tmem_base + BN, da, d_ones, idesc_ones, p ? 1u : 0u);
            }
            umma_commit(mma_done);
        }
        mbar_wait(mma_done, phase);
        phase ^= 1;
        tc_fence_after();

        // ---- epilogue: this thread owns TMEM lane r = its pixel
        const int nsa = -(int)tmem_ld1(trow + BN);   // minus the pixel's activation sum (ones-tile columns)
        uint8_t *orow = a.out + (size_t)m * a.CSO;
        constexpr int STEP = 16;
        const bool side = (a.out_acc != nullptr) || (a.out_f32 != nullptr);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += STEP) {
            uint32_t v[STEP];
            tmem_ld16(trow + c0, v);
            uint32_t packed[STEP / 4];
            epi_chunk_small<STEP>(actm, a.ep.saturate, v, nsa, s_q + c0, s_mc + c0, a.ep.zp_out, packed);
            yq::mask_pad_channels<STEP>(packed, a.N - c0);
            if (side && valid) {   // parity / quant_stop side outputs (not on the throughput path)
#pragma unroll
                for (int j = 0; j < STEP; ++j) {
                    const int oc = c0 + j;
                    if (oc < a.N) {
                        if (a.out_acc) a.out_acc[(size_t)m * a.CSO + oc] = (int)v[j] + s_q[oc].y * nsa;
                        if (a.out_f32) {
                            const uint8_t u = (uint8_t)(packed[j / 4] >> (8 * (j % 4)));
                            a.out_f32[((size_t)n * a.N + oc) * a.OH * a.OW + (size_t)oy * a.OW + ox] = yq::dequant_f32(a.ep, u);
                        }
                    }
                }
            }
            if (valid) {
#pragma unroll
                for (int h = 0; h < STEP / 16; ++h)
                    *reinterpret_cast<uint4 *>(orow + c0 + h * 16) = make_uint4(packed[4 * h], packed[4 * h + 1], packed[4 * h + 2], packed[4 * h + 3]);
            }
        }
        tc_fence_before();
        __syncthreads();         // every lane drained before the next tile's MMAs overwrite TMEM / rows are rebuilt
    }
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<small_tmem_cols<BN>()>(tmem_base);
    }
}

struct SmallState {
    int CS, BN, ctas_per_sm;
    uint8_t *wimg = nullptr;
};

template <int CS, int BN>
int launch_small(SmallState *st, const SmallArgs &a, cudaStream_t stream)
{
    using L = SmallSmem<CS, BN>;
    static int ctas_per_sm = 0, n_sm = 0;
    const int smem = L::TOTAL + 1024;
    if (!ctas_per_sm) {
        YQ_CUDA(cudaFuncSetAttribute(conv_u8_tc_small_kernel<CS, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        // cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for kernels that allocate tensor memory (measured on
        // B200), although the hardware co-schedules as many CTAs as smem / registers / TMEM columns allow: count by hand.
        int dev = 0, smem_sm = 0;
        cudaFuncAttributes fa;
        YQ_CUDA(cudaFuncGetAttributes(&fa, conv_u8_tc_small_kernel<CS, BN>));
        YQ_CUDA(cudaGetDevice(&dev));
        YQ_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        YQ_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
        const int by_smem = smem_sm / (smem + 1024 + (int)fa.sharedSizeBytes);
        const int regs_per_cta = ((fa.numRegs + 7) / 8 * 8) * SM_THREADS;
        const int by_regs = 65536 / (regs_per_cta > 0 ? regs_per_cta : 1);
        const int by_threads = 2048 / SM_THREADS;
        int occ = by_smem < by_regs ? by_smem : by_regs;
        if (by_threads < occ) occ = by_threads;
        const int tmem_limit = 512 / small_tmem_cols<BN>();   // every resident CTA must own its TMEM columns
        ctas_per_sm = occ < tmem_limit ? occ : tmem_limit;
        if (getenv("YQ_DEBUG")) fprintf(stderr, "yq: small<%d,%d> occ=%d tmem_limit=%d n_sm=%d smem=%d\n", CS, BN, occ, tmem_limit, n_sm, smem);
        if (ctas_per_sm < 1) return yq::fail("conv_u8_tc_small_kernel<%d,%d> does not fit on an SM", CS, BN);
    }
    int grid = n_sm * ctas_per_sm;
    if (grid > a.num_tiles) grid = a.num_tiles;
    conv_u8_tc_small_kernel<CS, BN><<<grid, SM_THREADS, smem, stream>>>(a);
    YQ_CHECK_LAUNCH();
    return 0;
}

}  // namespace

int yq_tc_small_supported(const yq_conv_layer *l)
{
    if (!l->int_form || !l->fused_mult) return 0;   // the integer-form epilogue needs M0 * 2^-31 / 2^-s parameters
    if (l->size != 3 || l->pad != 1 || (l->stride != 1 && l->stride != 2)) return 0;
    if (!(l->cs_in == 4 || l->cs_in == 16 || l->cs_in == 32)) return 0;
    if (l->cs_out < 16 || l->cs_out > 64 || (l->cs_out != 16 && l->cs_out != 32 && l->cs_out != 64)) return 0;
    return 1;
}

int yq_tc_small_prepare(yq_conv_layer *l, void **state)
{
    SmallState *st = new SmallState();
    st->CS = l->cs_in;
    st->BN = l->cs_out;
    const int K = TAPS * st->CS, KPAD = (K + 31) / 32 * 32, P = KPAD / 32;
    // shared-memory image: panel p = K bytes [32p, 32p+32) of every row; SWIZZLE_32B inside each 8-row atom
    std::vector<uint8_t> img((size_t)P * st->BN * 32, 0);
    for (int oc = 0; oc < l->n; ++oc)
        for (int t = 0; t < TAPS; ++t)
            for (int ci = 0; ci < l->c; ++ci) {
                const int k = t * st->CS + ci;
                const int p = k / 32, c = (k % 32) / 16, b = k % 16;
                img[(size_t)p * st->BN * 32 + oc * 32 + ((c ^ ((oc >> 2) & 1)) << 4) + b] = l->host_w[((size_t)oc * l->c + ci) * TAPS + t];
            }
    if (cudaMalloc((void **)&st->wimg, img.size()) != cudaSuccess || cudaMemcpy(st->wimg, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(st->wimg);
        delete st;
        return yq::fail("tcgen05 small-c flavour: weight upload failed");
    }
    *state = st;
    return 0;
}

void yq_tc_small_free(void *state)
{
    SmallState *st = (SmallState *)state;
    if (!st) return;
    cudaFree(st->wimg);
    delete st;
}

int yq_tc_small_forward(yq_conv_layer *l, void *state, const uint8_t *in_u8, uint8_t *out_u8, float *out_f32, int32_t *out_acc, int batch,
                        cudaStream_t stream)
{
    SmallState *st = (SmallState *)state;
    SmallArgs a;
    memset(&a, 0, sizeof a);
    a.in = in_u8; a.out = out_u8; a.out_f32 = l->quant_stop_flag ? out_f32 : nullptr; a.out_acc = out_acc; a.wimg = st->wimg;
    a.ep = yq::make_epi(l);
    a.B = batch; a.H = l->h; a.W = l->w; a.C = l->c; a.OH = l->out_h; a.OW = l->out_w; a.N = l->n; a.CSO = l->cs_out;
    a.stride = l->stride; a.pad = l->pad; a.zp_in = l->zp_in;
    a.M_total = batch * l->out_h * l->out_w;
    a.num_tiles = (a.M_total + 127) / 128;
#define YQ_SM(CS_, BN_) if (st->CS == CS_ && st->BN == BN_) return launch_small<CS_, BN_>(st, a, stream)
    YQ_SM(4, 16); YQ_SM(4, 32); YQ_SM(4, 64);
    YQ_SM(16, 16); YQ_SM(16, 32); YQ_SM(16, 64);
    YQ_SM(32, 16); YQ_SM(32, 32); YQ_SM(32, 64);
#undef YQ_SM
    return yq::fail("tcgen05 small-c flavour: no instantiation for cs_in=%d cs_out=%d", st->CS, st->BN);
}
