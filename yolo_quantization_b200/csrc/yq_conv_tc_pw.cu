// yq_conv_tc_pw.cu -- tcgen05 (kind::i8) POINTWISE (1x1, stride 1) convolution as a persistent streaming GEMM with the
// whole filter bank RESIDENT in shared memory: the flavour of the narrow 1x1 layers and of the detection heads.
//
// Why (measured on B200, profiles/r2_forward_full.csv): in conv_u8_tc_flat_kernel a 1x1 layer with few output channels
// is one short dependent chain per CTA -- barrier init, TMEM allocation, a TMA round trip, a handful of MMAs, the
// epilogue, the store -- and the launch is a few waves of such chains: the heads of yolov3-tiny ran at 0.15 - 0.23 of
// the HBM roofline with every pipe below 25 %.  These layers have no arithmetic to speak of (K <= 512, N <= 255); they
// are a stream of activation rows through a small matrix.  So:
//   * B  the filter bank [NC = round_up(n + 1, 16)][K] is loaded ONCE per CTA (before griddepcontrol.wait: it overlaps the
//        previous kernel's tail) and stays in shared memory; row n is all ones, so TMEM column n of every accumulator is the
//        position's activation sum (uint8 weights with a uint8 zero point: acc = sum w*a - zp_w * sum a,
//        convolutional_layer.c:718-721) -- no separate ones rows per stage, no sum warps;
//   * A  tiles of 128 consecutive positions x KC channels stream through a deep TMA ring (up to 12 stages) that runs
//        across tile boundaries: the loads never drain between tiles;
//   * D  2 or 4 TMEM accumulators; 16 epilogue warps in 4 groups, group g works on accumulator g % nbuf (and, with two
//        accumulators, on one half of its channels): the epilogue of tile i overlaps the loads and MMAs of tiles i+1 ...;
//   * one tcgen05.commit per tile when K is a single chunk (the commit both frees the ring stage and publishes the
//        accumulator: commits cannot follow each other faster than every ~358 clocks per SM, see yq_conv_tc_flat2.cu).
// Tensors, position arithmetic, requantisation and halo handling are those of yq_conv_tc_flat.cu (flat halo-padded strips;
// halo positions of the output are written with the consumer's zero point), or plain [B][H][W][C] strips (plain = 1).
// Detection heads (quant_stop followed by [yolo]): the float values are a 256-entry table per channel class
// (yq_conv_tc_flat.cu), written straight to the yolo layer's NCHW tensor, one coalesced 128-byte store per channel and warp.
// Restates convolutional_layer.c:694-761 for size 1, stride 1, pad 0, c % 64 == 0, n <= 255.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <map>
#include <type_traits>
#include <vector>

#include "yq_common.h"
#include "yq_epilogue.cuh"
#include "yq_tc_ptx.cuh"

using namespace yqtc;

namespace {

constexpr int PW_GROUPS = 4;                       // epilogue groups of four warps (one warp per TMEM lane quarter)
constexpr int PW_EPI_WARPS = 4 * PW_GROUPS;
constexpr int PW_THREADS = 64 + 32 * PW_EPI_WARPS;
constexpr int PW_MAX_ASTAGES = 12;
constexpr int PW_MAX_NC = 256;
constexpr int PW_STAGE_SLICE = 2048;               // one epilogue warp's staging slice: 32 positions x up to 64 channels

struct PwArgs {
    yq::EpiParams ep;
    float *out_yolo;       // the following yolo layer's NCHW tensor (detection heads) or nullptr
    const float *lut;      // [0,256): (u8 - zp_out) * s_out   [256,512): its logistic
    int yolo_per;          // 4 + classes + 1 channels per anchor
    int N, CSO, NC;        // real channels, channel stride of the output, MMA N (row N of the bank = ones)
    int B, H, W, NP;       // NP = positions that belong to images (flat: their halo included)
    int plain;             // plain [B][H][W][C] strips: every p < NP is a pixel
    int cpt;               // KC-chunks per tile (K / KC)
    int a_stages, nbuf;    // ring stages, TMEM accumulators (2 or 4)
    int tmem_cols;         // power of two >= nbuf * NC
    int wc;                // channels one epilogue warp requantises per tile: CSO / (4 / nbuf)
    int rowb;              // bytes per staging row = inner box of the store: min(wc, 64)
    int store_u8;          // the uint8 tensor is stored (always, except heads on request and layers that only write the upsampled tensor)
    uint8_t *up_out;       // the FOLLOWING upsample layer (stride 2, upsample_layer.c:92-101 / blas.c:334-351) fused: the flat tensor of
                           // (2H x 2W) pixels, same channel stride, that receives every output pixel four times; or nullptr
    int num_tiles;
    int split;             // chunks [0, split) of the input channels come from the second tensor map (see yq_conv_tc_flat2x.cu)
    int group;             // chunks per ring stage: loaded under one barrier, released by one commit (1: pointwise; 3 or 9: patch mode)
    // patch mode (size x size filters, stride 1 or 2, bank [NC][taps][cs_in] resident): a tile is a TW x TH patch of output pixels of one
    // image (TW * TH = 128, TW a power of two), chunk c = (tap, channel chunk) is ONE 4-D tiled TMA box whose start is shifted by the tap
    // and which steps through the input with elementStrides = stride (yq_conv_tc.cu); the input's halo holds zp_in, the output map
    // covers the interior of the output tensor (partial tiles are clipped by the TMA unit)
    int patch;             // 1: nine (tap) boxes per channel chunk; 2: PAIR form for c = 32, stride 2 -- the input seen as pixel pairs (64 contiguous bytes),
                           //    a filter row is two boxes of consecutive pairs: [px 2x | px 2x+1] and [px 2x+2 | (zero weights)], K = 6 x 64 per tile
    int tw_shift, tiles_x, tiles_y, OH, OW;
    int size, stride, cptap;   // filter size, stride, KC-chunks per tap
    int trace;             // YQ_PW_TRACE
    uint32_t halo_word;
    uint32_t magic_w, magic_h;
};

// YQ_PW_TRACE=1: every CTA records globaltimer (ns) at a few events into yq_pw_trace[blockIdx.x * 64 + event] (tools/probes/pw_trace.py)
__device__ unsigned long long yq_pw_trace[256 * 64];
__device__ __forceinline__ void pw_mark(int on, int ev)
{
    if (on && blockIdx.x < 256 && ev < 64) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        yq_pw_trace[blockIdx.x * 64 + ev] = t;
    }
}

__device__ __forceinline__ void pw_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void pw_tmem_alloc(uint32_t *slot, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void pw_tmem_dealloc(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// ONE = K is a single chunk: one commit per tile on done[stage], which the producer (stage free) and the epilogue (accumulator
// complete) both wait on; a_stages % nbuf == 0 makes the stage name the accumulator.  Otherwise a commit per chunk frees its
// stage and one more per tile publishes the accumulator.
// PATCH: 0 = pointwise strips, 1 = patch mode with a box per tap, 2 = its PAIR form (a compile-time choice: the strip kernels carry none of
// the patch arithmetic -- with it in one kernel the narrow 1x1 layers of the full yolov3 ran 5 - 28 % slower, twice the code per tile)
template <int KC, bool ONE, bool YOLO, int ACTM, int PATCH>
__global__ void __launch_bounds__(PW_THREADS, 1) conv_u8_tc_pw_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                  const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmA2,
                                                                  const PwArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int A_CHUNK = 128 * KC;
    const int A_STAGE = a.group * A_CHUNK;
    const int b_chunk = a.NC * KC;                                   // whole swizzle atoms (NC % 16 == 0)
    uint8_t *sB = smem;                                              // [cpt][NC rows][KC] resident filter bank
    uint8_t *sA = sB + ((a.cpt * b_chunk + 1023) & ~1023);           // a.a_stages ring stages of a.group x (128 positions x KC)
    uint8_t *sOut = sA + a.a_stages * A_STAGE;                       // one staging slice per epilogue warp
    int4 *s_q = (int4 *)(sOut + PW_EPI_WARPS * PW_STAGE_SLICE);      // {bias, zw, 2*M0, shift}
    double *s_mc = (double *)(s_q + PW_MAX_NC);
    float *s_lut = (float *)(s_mc + PW_MAX_NC);
    int *s_sel = (int *)(s_lut + 512);
    uint64_t *a_full = (uint64_t *)(s_sel + PW_MAX_NC);
    uint64_t *a_empty = a_full + PW_MAX_ASTAGES;                     // ONE: "done" (MMAs of the tile in this stage have completed)
    uint64_t *acc_full = a_empty + PW_MAX_ASTAGES;
    uint64_t *acc_empty = acc_full + 4;
    uint64_t *b_full = acc_empty + 4;
    uint32_t *tmem_slot = (uint32_t *)(b_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nst = a.a_stages, nbuf = a.nbuf, chunks = a.cpt, group = PATCH ? a.group : 1, groups = PATCH ? a.cpt / a.group : a.cpt;
    if (threadIdx.x == 0) pw_mark(a.trace, 0);                       // CTA start

    if (threadIdx.x == 0) {
        for (int s = 0; s < PW_MAX_ASTAGES; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], (uint32_t)(PW_EPI_WARPS / nbuf));
        }
        mbar_init(b_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) pw_tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
        if (a.split) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
    }
    if (warp >= 2) {
        const int t = threadIdx.x - 64;
        for (int i = t; i < a.CSO; i += 32 * PW_EPI_WARPS) {   // (CSO <= the parameter arrays' padded length; NC may exceed it)
            s_q[i] = __ldg(a.ep.chanq + i);
            s_mc[i] = __ldg(a.ep.mcomb + i);
        }
        if (YOLO) {   // yolo_layer.c:137-146: channels 2, 3 (w, h) of every anchor stay linear, the rest go through the logistic
            for (int i = t; i < 512; i += 32 * PW_EPI_WARPS) s_lut[i] = __ldg(a.lut + i);
            for (int i = t; i < a.CSO; i += 32 * PW_EPI_WARPS) {
                const int e = i % a.yolo_per;
                s_sel[i] = (e == 2 || e == 3) ? 0 : 256;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) pw_mark(a.trace, 1);                       // setup done
    // the filter bank is a constant of the network: its load may overlap the previous kernel's tail
    if (warp == 0 && elect_one()) {
        mbar_expect_tx(b_full, (uint32_t)(chunks * b_chunk));
        for (int c = 0; c < chunks; ++c) tma_load_2d(sB + c * b_chunk, &tmB, b_full, c * KC, 0);
    }
    yq_pdl_wait_then_release();                             // no activation tensor was touched so far
    if (threadIdx.x == 0) pw_mark(a.trace, 2);                       // previous grid complete

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform control flow, one elected lane issues) =====================
        int s = 0, tr = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++tr) {
            const int p0 = tile * 128;
            int tx = 0, ty = 0, n0 = 0;
            if (PATCH) {
                const int t2 = tile / a.tiles_x;
                tx = tile - t2 * a.tiles_x;
                n0 = t2 / a.tiles_y;
                ty = t2 - n0 * a.tiles_y;
            }
            const int cx0 = (tx << a.tw_shift) * a.stride, cy0 = (ty << (7 - a.tw_shift)) * a.stride;            // tap (0, 0) of the patch: the map starts in the halo
            for (int g = 0; g < groups; ++g) {
                mbar_wait(&a_empty[s], ph ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&a_full[s], (uint32_t)A_STAGE);
                    for (int j = 0; j < group; ++j) {
                        const int c = g * group + j;
                        uint8_t *dst = sA + s * A_STAGE + j * A_CHUNK;
                        if (PATCH == 2) {
                            // chunk c = (filter row c / 2, pair x + c % 2): rows of 64 contiguous bytes, consecutive pairs, every second input row
                            tma_load_4d(dst, &tmA, &a_full[s], 0, (tx << a.tw_shift) + (c & 1), cy0 + (c >> 1), n0);
                        } else if (PATCH == 1) {
                            const int tap = c / a.cptap, chunk = c - tap * a.cptap;
                            const int ky = tap / a.size, kx = tap - ky * a.size;
                            tma_load_4d(dst, &tmA, &a_full[s], chunk * KC, cx0 + kx, cy0 + ky, n0);
                        } else {
                            // a.split > 0: the input is the channel concatenation [tmA2 | tmA] of two tensors (a route that is never materialised)
                            tma_load_2d(dst, c < a.split ? &tmA2 : &tmA, &a_full[s], (c < a.split ? c : c - a.split) * KC, p0);
                        }
                    }
                }
                if (++s == nst) { s = 0; ph ^= 1; }
            }
            if (lane == 0 && tr < 8) pw_mark(a.trace, 4 + 4 * tr + 3);                     // tile's loads issued
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc(a.NC);
        int s = 0, b = 0;
        uint32_t ph = 0, phb = 0;
        mbar_wait(b_full, 0);
        if (lane == 0) pw_mark(a.trace, 3);                  // filter bank landed
        int tr = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++tr) {
            mbar_wait(&acc_empty[b], phb ^ 1);               // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)(b * a.NC);
            for (int g = 0; g < groups; ++g) {
                mbar_wait(&a_full[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    for (int j = 0; j < group; ++j) {
                        const int c = g * group + j;
                        const uint64_t da = make_desc<KC>(smem_u32(sA + s * A_STAGE + j * A_CHUNK));
                        const uint64_t db = make_desc<KC>(smem_u32(sB + c * b_chunk));
#pragma unroll
                        for (int k = 0; k < KC / 32; ++k) umma_i8(acc, da + 2 * k, db + 2 * k, idesc, (c | k) ? 1u : 0u);
                    }
                    umma_commit(&a_empty[s]);
                }
                if (++s == nst) { s = 0; ph ^= 1; }
            }
            if (!ONE && elect_one()) umma_commit(&acc_full[b]);
            if (lane == 0 && tr < 8) pw_mark(a.trace, 4 + 4 * tr);                         // tile's operands landed, MMAs issued
            if (++b == nbuf) { b = 0; phb ^= 1; }
        }
    } else {
        // ===================== epilogue: 16 warps, each on its own =====================
        const int ew = warp - 2;
        const int q = warp & 3;                 // TMEM lane quarter the hardware lets this warp read
        const int g = ew >> 2;                  // group
        const int b = g % nbuf;                 // my accumulator
        const int part = g / nbuf;              // my share of its channels
        const int cbeg = part * a.wc;
        const int rowb = a.rowb;                // 32 or 64
        const int cpp = rowb / 16;              // chunks per store pass
        const int npass = a.wc / rowb;
        uint8_t *stage = sOut + ew * PW_STAGE_SLICE;
        const int pitch = a.W + (a.plain ? 0 : 1), rows_h = a.H + (a.plain ? 0 : 1);
        const int hw = a.H * a.W;
        for (int it = b; blockIdx.x + (long long)it * gridDim.x < a.num_tiles; it += nbuf) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int p0 = tile * 128;
            int n, yy, xx, px0 = 0, py0 = 0;
            bool valid;
            if (PATCH) {
                // tile row r = (hi, wi) of the TW x TH patch at (px0, py0) of image n; rows outside the image are clipped by the store
                const int tx = tile % a.tiles_x, t2 = tile / a.tiles_x, ty = t2 % a.tiles_y;
                const int r = q * 32 + lane;
                n = t2 / a.tiles_y;
                px0 = tx << a.tw_shift;
                py0 = ty << (7 - a.tw_shift);
                xx = px0 + (r & ((1 << a.tw_shift) - 1));
                yy = py0 + (r >> a.tw_shift);
                valid = xx < a.OW && yy < a.OH;
            } else {
                const int p = p0 + q * 32 + lane;
                const int row = (int)__umulhi((uint32_t)p, a.magic_w);
                const int col = p - row * pitch;
                n = (int)__umulhi((uint32_t)row, a.magic_h);
                const int y1 = row - n * rows_h;
                valid = p < a.NP && (a.plain || (col >= 1 && y1 >= 1));
                yy = a.plain ? y1 : y1 - 1;
                xx = a.plain ? col : col - 1;
            }
            if (ONE) {
                const int s = it % nst;
                mbar_wait(&a_empty[s], (uint32_t)((it / nst) & 1));
            } else {
                mbar_wait(&acc_full[b], (uint32_t)((it / nbuf) & 1));
            }
            tc_fence_after();
            if ((ew & 3) == 0 && lane == 0 && it < 8) pw_mark(a.trace, 4 + 4 * it + 1);    // tile's accumulator complete
            const uint32_t trow = tmem_base + (uint32_t)(b * a.NC) + ((uint32_t)(q * 32) << 16);
            const int nsa = -(int)tmem_ld1(trow + a.N);      // minus the position's activation sum (the ones row)
            // the head's plane of this pixel in the yolo tensor: channel c of it is ybase[c * hw]
            float *ybase = YOLO ? a.out_yolo + (size_t)n * a.N * hw + (size_t)yy * a.W + xx : nullptr;
            // One 16-channel chunk per trip of a ROLLED loop, its TMEM load not overlapped with the previous chunk's arithmetic by this
            // warp: the other fifteen epilogue warps hide it, and the loop body stays a few hundred instructions.  (The first version
            // unrolled pairs of chunks with the load of one in flight under the other: 1 300 instructions per tile and warp that each
            // warp walks a handful of times per launch -- half of the epilogue's stall samples were instruction-cache misses.)
#pragma unroll 1
            for (int pass = 0; pass < npass; ++pass) {
                if (a.store_u8) {
                    // my staging slice is free once the previous store has read it
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
                }
#pragma unroll 1
                for (int chl = 0; chl < cpp; ++chl) {
                    const int c0 = cbeg + 16 * (pass * cpp + chl);
                    uint32_t v[16];
                    tmem_ld16(trow + c0, v);
                    int r[16];
                    int extra[16];
                    yq::requant_chunk_vals<ACTM, false, 16, false>(v, nsa, extra, s_q + c0, s_mc + c0, a.ep.zp_out, r, a.ep.xlim);
                    if (YOLO && valid) {
                        // the detection heads: one table lookup and one 4-byte store per output (lanes = consecutive pixels)
                        float *dst = ybase + (size_t)c0 * hw;
                        const int4 *sel = reinterpret_cast<const int4 *>(s_sel + c0);
                        const int nreal = a.N - c0;
                        if (nreal >= 16) {
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4) {
                                const int4 sl = sel[j4];
                                dst[(4 * j4 + 0) * hw] = s_lut[sl.x + (r[4 * j4 + 0] & 255)];
                                dst[(4 * j4 + 1) * hw] = s_lut[sl.y + (r[4 * j4 + 1] & 255)];
                                dst[(4 * j4 + 2) * hw] = s_lut[sl.z + (r[4 * j4 + 2] & 255)];
                                dst[(4 * j4 + 3) * hw] = s_lut[sl.w + (r[4 * j4 + 3] & 255)];
                            }
                        } else {
#pragma unroll 1
                            for (int j = 0; j < nreal; ++j) {
                                int rj = r[0];
#pragma unroll
                                for (int k = 1; k < 16; ++k) rj = j == k ? r[k] : rj;
                                dst[j * hw] = s_lut[s_sel[c0 + j] + (rj & 255)];
                            }
                        }
                    }
                    if (!YOLO && !PATCH && a.up_out) {
                        // conv -> upsample(2) in one launch: the pixel's 16 bytes go to its four copies in the (2H x 2W) flat strip
                        // (that strip's halo is never written: it keeps the fill the plan gave it)
                        if (valid) {
                            uint32_t packed[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) packed[k] = yq::pack_low_bytes(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
                            yq::mask_pad_channels<16>(packed, a.N - c0);
                            const uint4 val = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                            const int W2 = 2 * a.W + 1;
                            const size_t P = ((size_t)n * (2 * a.H + 1) + 2 * yy + 1) * W2 + 2 * xx + 1;
                            uint8_t *d = a.up_out + P * a.CSO + c0;
                            *reinterpret_cast<uint4 *>(d) = val;
                            *reinterpret_cast<uint4 *>(d + a.CSO) = val;
                            *reinterpret_cast<uint4 *>(d + (size_t)W2 * a.CSO) = val;
                            *reinterpret_cast<uint4 *>(d + (size_t)(W2 + 1) * a.CSO) = val;
                        }
                    } else if (a.store_u8) {
                        uint32_t packed[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) packed[k] = yq::pack_low_bytes(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
                        if (!valid && !PATCH) packed[0] = packed[1] = packed[2] = packed[3] = a.halo_word;
                        yq::mask_pad_channels<16>(packed, a.N - c0);
                        // 64-byte rows: SWIZZLE_64B (16-byte chunk ^ row bits 1-2); 32-byte rows: SWIZZLE_32B (chunk ^ row bit 2)
                        const int sw = rowb == 64 ? (chl ^ ((lane >> 1) & 3)) : (chl ^ ((lane >> 2) & 1));
                        *reinterpret_cast<uint4 *>(stage + lane * rowb + sw * 16) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    }
                }
                if (pass + 1 == npass) {
                    // this warp's TMEM reads of the tile are done: hand the accumulator back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) pw_arrive(&acc_empty[b]);
                }
                if (a.store_u8) {
                    fence_proxy_async();          // my staging writes -> visible to the TMA unit
                    __syncwarp();
                    if (lane == 0) {
                        if (PATCH) tma_store_4d(&tmO, stage, cbeg + pass * rowb, px0, py0 + q * (32 >> a.tw_shift), n);   // my 32 / TW rows of the patch
                        else tma_store_2d(&tmO, stage, cbeg + pass * rowb, p0 + q * 32);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            }
            if ((ew & 3) == 0 && lane == 0 && it < 8) pw_mark(a.trace, 4 + 4 * it + 2);    // tile's epilogue done (stores issued)
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // global writes complete before the CTA retires
        if (ew == 0 && lane == 0) pw_mark(a.trace, 40);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        pw_tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
    }
    if (threadIdx.x == 0) pw_mark(a.trace, 41);                      // CTA end
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn pw_get_encode()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int pw_encode_2d(CUtensorMap *m, const void *ptr, uint64_t rows, int row_bytes, int box_c, int box_rows, CUtensorMapL2promotion prom)
{
    EncodeTiledFn enc = pw_get_encode();
    if (!enc) return yq::fail("cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     box_c >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (box_c == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B), prom,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return yq::fail("cuTensorMapEncodeTiled(%llu x %d, box %d x %d) failed: %d", (unsigned long long)rows, row_bytes, box_c, box_rows, (int)r);
    return 0;
}

struct PwState {
    int KC, NC;
    bool pair = false;          // patch mode, PAIR form (PwArgs::patch == 2)
    uint8_t *w = nullptr;       // [NC][cs_in]: rows < n the filters, row n all ones, the rest zero
    float *lut = nullptr;       // quant_stop layers: dequantized value and its logistic for each of the 256 output bytes
    CUtensorMap tmB;
    struct Key {
        const void *in, *in2;
        void *out;
        int batch;
        bool operator<(const Key &o) const { return in != o.in ? in < o.in : in2 != o.in2 ? in2 < o.in2 : out != o.out ? out < o.out : batch < o.batch; }
    };
    struct Maps {
        CUtensorMap a, o, a2;
    };
    std::map<Key, Maps> maps;
};

// shared memory of a launch with `stages` ring stages of `group` chunks each
int pw_smem_bytes(int NC, int K, int KC, int stages, int group = 1)
{
    return 1024 + ((NC * K + 1023) & ~1023) + stages * group * 128 * KC + PW_EPI_WARPS * PW_STAGE_SLICE + PW_MAX_NC * (16 + 8 + 4) + 2048 + 512;
}

template <int KC, bool ONE, bool YOLO, int ACTM, int PATCH>
int pw_launch_v(PwState *st, const CUtensorMap &tmA, const CUtensorMap &tmO, const CUtensorMap &tmA2, const PwArgs &a, int smem, int grid, cudaStream_t stream)
{
    auto kern = conv_u8_tc_pw_kernel<KC, ONE, YOLO, ACTM, PATCH>;
    if (yq::ensure_dynamic_smem((const void *)kern, smem)) return -1;
    YQ_CUDA(yq::launch_pdl(kern, dim3(grid), dim3(PW_THREADS), smem, stream, tmA, st->tmB, tmO, tmA2, a));
    return 0;
}

template <int KC, int PATCH>
int pw_launch(PwState *st, const CUtensorMap &tmA, const CUtensorMap &tmO, const CUtensorMap &tmA2, const PwArgs &a, int smem, int grid, cudaStream_t stream)
{
    // the activation is a launch constant: one epilogue form per kernel keeps the code each warp walks short (heads are LINEAR: checked by the caller)
    const int actm = yq::act_mode(a.ep.act);
    if (a.cpt == a.group) {      // one ring stage per tile: a single commit frees the stage and publishes the accumulator
        if (!PATCH && a.out_yolo) return pw_launch_v<KC, true, !PATCH, 1, PATCH>(st, tmA, tmO, tmA2, a, smem, grid, stream);
        if (actm == 0) return pw_launch_v<KC, true, false, 0, PATCH>(st, tmA, tmO, tmA2, a, smem, grid, stream);
        if (actm == 1) return pw_launch_v<KC, true, false, 1, PATCH>(st, tmA, tmO, tmA2, a, smem, grid, stream);
        return pw_launch_v<KC, true, false, 2, PATCH>(st, tmA, tmO, tmA2, a, smem, grid, stream);
    }
    if (!PATCH && a.out_yolo) return pw_launch_v<KC, false, !PATCH, 1, PATCH>(st, tmA, tmO, tmA2, a, smem, grid, stream);
    if (actm == 0) return pw_launch_v<KC, false, false, 0, PATCH>(st, tmA, tmO, tmA2, a, smem, grid, stream);
    if (actm == 1) return pw_launch_v<KC, false, false, 1, PATCH>(st, tmA, tmO, tmA2, a, smem, grid, stream);
    return pw_launch_v<KC, false, false, 2, PATCH>(st, tmA, tmO, tmA2, a, smem, grid, stream);
}

}  // namespace

int yq_tc_pw_supported(const yq_conv_layer *l)
{
    static const bool off = getenv("YQ_NO_PW") && atoi(getenv("YQ_NO_PW"));      // A/B measurements
    if (off || !l->int_form || !l->fused_mult || l->saturate) return 0;
    if (l->size != 1 || l->stride != 1 || l->pad != 0) return 0;
    if (l->c != l->cs_in || l->cs_in % 64) return 0;                             // no pad lanes: they would count in sum(a)
    const int NC = yq::round_up(l->n + 1, 16);
    if (NC > PW_MAX_NC) return 0;
    if (l->cs_out > NC) return 0;
    const int wc = l->cs_out / (PW_GROUPS / (4 * NC <= 512 ? 4 : 2));            // channels per epilogue warp: 32, 64 or a multiple of 64
    if (wc % 32 || (wc > 64 && wc % 64)) return 0;
    const int KC = (l->cs_in % 128) ? 64 : 128;
    if (pw_smem_bytes(NC, l->cs_in, KC, l->cs_in == KC ? 4 : 3) > 227 * 1024) return 0;   // the bank and a minimal ring
    return pw_get_encode() != nullptr;
}

// a detection head (quant_stop + fused yolo) runs here when its activation is LINEAR (every yolo cfg of the reference); others keep the flat kernel
int yq_tc_pw_chunk(const void *state) { return state ? ((const PwState *)state)->KC : 0; }
int yq_tc_pw_head_supported(const yq_conv_layer *l) { return l->tc_pw && l->quant_stop_flag && yq::act_mode(l->activation) == 1 ? 1 : 0; }

int yq_tc_pw_prepare(yq_conv_layer *l, void **state)
{
    PwState *st = new PwState();
    st->KC = (l->cs_in % 128) ? 64 : 128;
    st->NC = yq::round_up(l->n + 1, 16);
    const size_t K = (size_t)l->cs_in;
    std::vector<uint8_t> wp;
    char tag[24];
    snprintf(tag, sizeof tag, "pw1.%d", st->NC);
    // (a data-parallel replica takes the image from the arena blob broadcast to its device: no host packing, no upload)
    const bool on_dev = yq::pack_fetch_device(l, tag, (size_t)st->NC * K, (void **)&st->w);
    if (!on_dev && (!yq::pack_fetch(l, tag, wp) || wp.size() != (size_t)st->NC * K)) {
        wp.assign((size_t)st->NC * K, 0);
        for (int oc = 0; oc < l->n; ++oc)
            for (int ci = 0; ci < l->c; ++ci) wp[(size_t)oc * K + ci] = l->host_w[(size_t)oc * l->c + ci];
        for (size_t ci = 0; ci < K; ++ci) wp[(size_t)l->n * K + ci] = 1;         // the ones row: TMEM column n = sum of the position's activations
        yq::pack_put(l, tag, wp);
    }
    auto cleanup = [&]() {
        cudaFree(st->w);
        cudaFree(st->lut);
        delete st;
        return -1;
    };
    if (!on_dev) {
        if (cudaMalloc((void **)&st->w, wp.size()) != cudaSuccess) return cleanup();
        if (cudaMemcpy(st->w, wp.data(), wp.size(), cudaMemcpyHostToDevice) != cudaSuccess) return cleanup();
    }
    if (pw_encode_2d(&st->tmB, st->w, (uint64_t)st->NC, (int)K, st->KC, st->NC, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return cleanup();
    if (l->quant_stop_flag) {
        // the head's float values are a table of the 256 output bytes, tabulated with the host's libm (yq_conv_tc_flat.cu)
        float lut[512];
        for (int u = 0; u < 256; ++u) {
            const float x = (float)(u - l->zp_out) * l->s_out;
            lut[u] = x;
            lut[256 + u] = (float)(1. / (1. + exp(-(double)x)));
        }
        if (cudaMalloc((void **)&st->lut, sizeof lut) != cudaSuccess) return cleanup();
        if (cudaMemcpy(st->lut, lut, sizeof lut, cudaMemcpyHostToDevice) != cudaSuccess) return cleanup();
    }
    *state = st;
    return 0;
}

void yq_tc_pw_free(void *state)
{
    PwState *st = (PwState *)state;
    if (!st) return;
    cudaFree(st->w);
    cudaFree(st->lut);
    delete st;
}

int yq_tc_pw_forward(yq_conv_layer *l, void *state, const uint8_t *in, uint8_t *out_u8, int halo_fill, float *out_yolo, int yolo_classes, int batch,
                     cudaStream_t stream, int plain, uint8_t *out_up2, const uint8_t *in_first, int c_first)
{
    PwState *st = (PwState *)state;
    if (!st || !in || !out_u8) return yq::fail("tcgen05 pointwise flavour: bad argument");
    // in_first != null: the input is the concatenation [in_first (c_first channels) | in (the rest)] of two tensors of the same geometry
    if (in_first && (c_first <= 0 || c_first >= l->cs_in || c_first % st->KC || (l->cs_in - c_first) % 16))
        return yq::fail("tcgen05 pointwise flavour: a two-tensor input splits at a multiple of %d channels", st->KC);
    const int c_second = in_first ? l->cs_in - c_first : l->cs_in;
    if (out_up2 && (plain || l->quant_stop_flag)) return yq::fail("tcgen05 pointwise flavour: the fused upsample needs a flat, quantized layer");
    if (l->quant_stop_flag && !out_yolo) return yq::fail("tcgen05 pointwise flavour: a quant_stop layer runs here only as a fused yolo head");
    if (out_yolo && yq::act_mode(l->activation) != 1) return yq::fail("tcgen05 pointwise flavour: detection heads are LINEAR (see yq_tc_pw_head_supported)");
    const int W1 = l->w + (plain ? 0 : 1), H1 = l->h + (plain ? 0 : 1);
    const long long NP = (long long)batch * H1 * W1;
    const long long rows_alloc = plain ? NP : NP + W1 + 2;     // + the trailing halo row (yq_act_geom_bytes)
    if (rows_alloc * W1 >= 0x100000000ll || rows_alloc + 128 >= 0x7fffffffll) return yq::fail("tcgen05 pointwise flavour: tensor too large for 32-bit position arithmetic");
    const int n_sm = yq::device_sm_count(), smem_max = yq::device_smem_optin();   // (of the CURRENT device: nothing cached per process)
    if (n_sm <= 0 || smem_max <= 0) return yq::fail("cannot query the device's multiprocessor count / shared memory size");
    PwArgs a;
    memset(&a, 0, sizeof a);
    a.group = 1;
    a.tiles_x = a.tiles_y = 1;
    a.ep = yq::make_epi(l);
    a.out_yolo = out_yolo;
    a.lut = st->lut;
    a.yolo_per = 4 + yolo_classes + 1;
    if (out_yolo && (!st->lut || l->n % a.yolo_per)) return yq::fail("tcgen05 pointwise flavour: layer is not a yolo head for %d classes", yolo_classes);
    a.N = l->n; a.CSO = l->cs_out; a.NC = st->NC;
    a.B = batch; a.H = l->h; a.W = l->w; a.NP = (int)NP;
    a.plain = plain ? 1 : 0;
    a.cpt = l->cs_in / st->KC;
    a.nbuf = 4 * st->NC <= 512 ? 4 : 2;
    a.tmem_cols = 32;
    while (a.tmem_cols < a.nbuf * st->NC) a.tmem_cols *= 2;
    // the epilogue walks the output's channel stride (pad channels are stored as zeros)
    a.wc = l->cs_out / (PW_GROUPS / a.nbuf);
    a.rowb = a.wc < 64 ? a.wc : 64;
    if (a.wc % 32 || (a.wc > 64 && a.wc % 64)) return yq::fail("tcgen05 pointwise flavour: %d channels per epilogue warp", a.wc);
    {
        static const int head_u8 = getenv("YQ_PW_HEAD_U8") ? atoi(getenv("YQ_PW_HEAD_U8")) : 1;     // 0: heads skip their uint8 tensor (A/B measurements)
        a.store_u8 = l->quant_stop_flag ? head_u8 : (out_up2 ? 0 : 1);
        a.up_out = out_up2;
    }
    // ring depth: what fits, a multiple of the accumulator count when a commit serves stage and accumulator alike
    int stages = PW_MAX_ASTAGES;
    while (stages > 3 && pw_smem_bytes(st->NC, l->cs_in, st->KC, stages) > smem_max) --stages;
    if (a.cpt == 1) stages = stages / 4 * 4;
    if (pw_smem_bytes(st->NC, l->cs_in, st->KC, stages) > smem_max) return yq::fail("tcgen05 pointwise flavour: the filter bank does not fit shared memory");
    a.a_stages = stages;
    const int smem = pw_smem_bytes(st->NC, l->cs_in, st->KC, stages);
    a.num_tiles = (int)((rows_alloc + 127) / 128);
    a.trace = getenv("YQ_PW_TRACE") && atoi(getenv("YQ_PW_TRACE")) ? 1 : 0;
    a.halo_word = 0x01010101u * (uint32_t)(halo_fill & 0xff);
    a.magic_w = (uint32_t)((0x100000000ull + W1 - 1) / W1);
    a.magic_h = (uint32_t)((0x100000000ull + H1 - 1) / H1);
    PwState::Key key{in, in_first, out_u8, batch * 2 + (plain ? 1 : 0)};
    auto it = st->maps.find(key);
    if (it == st->maps.end()) {
        if (st->maps.size() > 64) st->maps.clear();
        PwState::Maps m;
        if (pw_encode_2d(&m.a, in, (uint64_t)rows_alloc, c_second, st->KC, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return -1;
        // the store box is one epilogue warp's pass: 32 positions x rowb channels
        if (pw_encode_2d(&m.o, out_u8, (uint64_t)rows_alloc, l->cs_out, a.rowb, 32, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return -1;
        m.a2 = m.a;
        if (in_first && pw_encode_2d(&m.a2, in_first, (uint64_t)rows_alloc, c_first, st->KC, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return -1;
        it = st->maps.emplace(key, m).first;
    }
    a.split = in_first ? c_first / st->KC : 0;
    const int grid = a.num_tiles < n_sm ? a.num_tiles : n_sm;
    const CUtensorMap &tmA = it->second.a, &tmO = it->second.o, &tmA2 = it->second.a2;
    if (st->KC == 128) return pw_launch<128, 0>(st, tmA, tmO, tmA2, a, smem, grid, stream);
    return pw_launch<64, 0>(st, tmA, tmO, tmA2, a, smem, grid, stream);
}

// ---------------------------------------------------------------------------------------------
// patch mode: size x size filters (3x3), stride 1 or 2, between halo-padded tensors; bank [NC][taps][cs_in] resident
// ---------------------------------------------------------------------------------------------
namespace {
// 4-D map over an NHWC tensor stored in geometry g (yq_conv_tc.cu: encode_nhwc).  halo > 0: coordinate (0, 0) is the halo pixel
// (-halo, -halo); halo = 0: the h x w interior, everything else out of bounds (zero fill on loads, clipped on stores).
int pw_encode_nhwc(CUtensorMap *m, const void *ptr, const yq_act_geom *g, int halo, int B, int H, int W, int CS, int box_c, int TW, int TH, int estride)
{
    EncodeTiledFn enc = pw_get_encode();
    if (!enc) return yq::fail("cuTensorMapEncodeTiled is not available from this driver");
    const uint8_t *base = (const uint8_t *)ptr + ((size_t)(g->pad - halo) * g->pitch_w + (g->pad - halo)) * CS;
    cuuint64_t dims[4] = {(cuuint64_t)CS, (cuuint64_t)(W + 2 * halo), (cuuint64_t)(H + 2 * halo), (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)CS, (cuuint64_t)g->pitch_w * CS, (cuuint64_t)g->rows_h * g->pitch_w * CS};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)(TW * estride), (cuuint32_t)(TH * estride), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<uint8_t *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     box_c >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (box_c == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return yq::fail("cuTensorMapEncodeTiled(NHWC %dx%dx%dx%d pitch %d rows %d box %d,%d,%d stride %d) failed: %d", B, H, W, CS, g->pitch_w, g->rows_h, box_c, TW, TH,
                        estride, (int)r);
    return 0;
}
int pwt_kc(const yq_conv_layer *l) { return (l->cs_in % 128) ? ((l->cs_in % 64) ? 32 : 64) : 128; }
// chunks per ring stage: the whole filter (one commit per tile) when four such stages fit beside the bank, else one filter row
int pwt_group(const yq_conv_layer *l, int NC, int KC, int smem_max)
{
    const int taps = l->size * l->size, cpt = taps * (l->cs_in / KC);
    if (pw_smem_bytes(NC, taps * l->cs_in, KC, 4, cpt) <= smem_max) return cpt;
    return l->size * (l->cs_in / KC);
}
}  // namespace

// PAIR form (c = 32, stride 2): see PwArgs::patch
static bool pwt_pair(const yq_conv_layer *l)
{
    const bool off = getenv("YQ_NO_PWT_PAIR") && atoi(getenv("YQ_NO_PWT_PAIR"));      // A/B measurements (read per call: the tests flip it)
    return !off && l->cs_in == 32 && l->stride == 2 && l->size == 3;
}

int yq_tc_pwt_supported(const yq_conv_layer *l)
{
    static const bool off = (getenv("YQ_NO_PW") && atoi(getenv("YQ_NO_PW"))) || (getenv("YQ_NO_PWT") && atoi(getenv("YQ_NO_PWT")));      // A/B measurements
    if (off || !l->int_form || !l->fused_mult || l->saturate || l->quant_stop_flag) return 0;
    if (l->size != 3 || l->pad != 1 || l->stride != 2) return 0;               // (stride-1 3x3 layers run the flat family: shared patches, fused shortcut)
    if (l->c != l->cs_in || l->cs_in % 32) return 0;                             // no pad lanes: they would count in sum(a)
    // c = 32 (32-byte rows: 128 strided 32-byte requests per tap box) is correct but slow: layer 1 of the full yolov3 0.339 ms against the
    // 0.266 ms of the small-c flavour, whose taps share an L1-cached patch -- nine boxes per tile re-read the input through L2.  YQ_PWT_C32=1 keeps it.
    {
        const bool c32 = getenv("YQ_PWT_C32") && atoi(getenv("YQ_PWT_C32"));      // (read per call: the tests flip it)
        if (l->cs_in % 64 && !c32 && !pwt_pair(l)) return 0;
    }
    if (l->out_w < 8 || l->out_h < 4) return 0;
    const int NC = yq::round_up(l->n + 1, 16);
    if (NC > PW_MAX_NC || l->cs_out > NC) return 0;
    const int wc = l->cs_out / (PW_GROUPS / (4 * NC <= 512 ? 4 : 2));
    if (wc % 32 || (wc > 64 && wc % 64)) return 0;
    const int KC = pwt_kc(l);
    if (pwt_pair(l)) return pw_smem_bytes(NC, 6 * 64, 64, 2, 6) <= 227 * 1024 && 2 * NC <= 512 && pw_get_encode() != nullptr;   // two stages of a whole tile
    if (pw_smem_bytes(NC, 9 * l->cs_in, KC, 3, 3 * (l->cs_in / KC)) > 227 * 1024) return 0;   // the bank and three stages of one filter row each
    return pw_get_encode() != nullptr;
}

int yq_tc_pwt_prepare(yq_conv_layer *l, void **state)
{
    PwState *st = new PwState();
    st->KC = pwt_kc(l);
    st->NC = yq::round_up(l->n + 1, 16);
    const int taps = l->size * l->size;
    if (pwt_pair(l)) {
        // [NC][filter row ky][j = 0: kx 0, kx 1 | j = 1: kx 2, zeros][32 channels]; the ones row counts every real input byte once
        st->KC = 64;
        st->pair = true;
        const size_t K = 6 * 64;
        std::vector<uint8_t> wp((size_t)st->NC * K, 0);
        for (int oc = 0; oc <= l->n; ++oc)
            for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx)
                    for (int ci = 0; ci < l->c; ++ci)
                        wp[(size_t)oc * K + (size_t)(ky * 2 + kx / 2) * 64 + (kx & 1) * 32 + ci] = oc == l->n ? 1 : l->host_w[((size_t)oc * l->c + ci) * taps + ky * 3 + kx];
        if (cudaMalloc((void **)&st->w, wp.size()) != cudaSuccess || cudaMemcpy(st->w, wp.data(), wp.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
            pw_encode_2d(&st->tmB, st->w, (uint64_t)st->NC, (int)K, 64, st->NC, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) {
            cudaFree(st->w);
            delete st;
            return -1;
        }
        *state = st;
        return 0;
    }
    const size_t K = (size_t)taps * l->cs_in;
    std::vector<uint8_t> wp;
    char tag[24];
    snprintf(tag, sizeof tag, "pwt.%d", st->NC);
    const bool on_dev = yq::pack_fetch_device(l, tag, (size_t)st->NC * K, (void **)&st->w);
    if (!on_dev && (!yq::pack_fetch(l, tag, wp) || wp.size() != (size_t)st->NC * K)) {
        wp.assign((size_t)st->NC * K, 0);
        for (int oc = 0; oc < l->n; ++oc)
            for (int t = 0; t < taps; ++t)
                for (int ci = 0; ci < l->c; ++ci) wp[(size_t)oc * K + (size_t)t * l->cs_in + ci] = l->host_w[((size_t)oc * l->c + ci) * taps + t];
        for (size_t k = 0; k < K; ++k) wp[(size_t)l->n * K + k] = 1;             // the ones row: TMEM column n = sum of the activations under the window
        yq::pack_put(l, tag, wp);
    }
    auto cleanup = [&]() {
        cudaFree(st->w);
        delete st;
        return -1;
    };
    if (!on_dev) {
        if (cudaMalloc((void **)&st->w, wp.size()) != cudaSuccess) return cleanup();
        if (cudaMemcpy(st->w, wp.data(), wp.size(), cudaMemcpyHostToDevice) != cudaSuccess) return cleanup();
    }
    if (pw_encode_2d(&st->tmB, st->w, (uint64_t)st->NC, (int)K, st->KC, st->NC, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return cleanup();
    *state = st;
    return 0;
}

// in_geom: halo-padded (pad >= 1) with the halo holding the layer's zp_in; out_geom: any geometry (only the interior is written)
int yq_tc_pwt_forward(yq_conv_layer *l, void *state, const uint8_t *in, const yq_act_geom *in_geom, uint8_t *out_u8, const yq_act_geom *out_geom, int batch,
                      cudaStream_t stream)
{
    PwState *st = (PwState *)state;
    if (!st || !in || !out_u8 || !in_geom || in_geom->pad < l->pad) return yq::fail("tcgen05 pointwise flavour (patch mode): bad argument");
    const int n_sm = yq::device_sm_count(), smem_max = yq::device_smem_optin();
    if (n_sm <= 0 || smem_max <= 0) return yq::fail("cannot query the device's multiprocessor count / shared memory size");
    const yq_act_geom go = out_geom ? *out_geom : yq_act_geom{0, l->out_w, l->out_h};
    PwArgs a;
    memset(&a, 0, sizeof a);
    a.ep = yq::make_epi(l);
    a.N = l->n; a.CSO = l->cs_out; a.NC = st->NC;
    a.B = batch; a.H = l->h; a.W = l->w; a.OH = l->out_h; a.OW = l->out_w;
    a.patch = st->pair ? 2 : 1;
    a.size = l->size; a.stride = l->stride; a.cptap = st->pair ? 1 : l->cs_in / st->KC;
    a.cpt = st->pair ? 6 : l->size * l->size * a.cptap;
    a.group = st->pair ? 6 : pwt_group(l, st->NC, st->KC, smem_max);
    a.tw_shift = l->out_w >= 16 ? 4 : 3;                        // TW = 16 (TH = 8) or 8 (TH = 16)
    const int TW = 1 << a.tw_shift, TH = 128 >> a.tw_shift;
    a.tiles_x = (l->out_w + TW - 1) / TW;
    a.tiles_y = (l->out_h + TH - 1) / TH;
    a.num_tiles = a.tiles_x * a.tiles_y * batch;
    const int K = st->pair ? 6 * 64 : l->size * l->size * l->cs_in;
    int stages = PW_MAX_ASTAGES;
    while (stages > 2 && pw_smem_bytes(st->NC, K, st->KC, stages, a.group) > smem_max) --stages;
    if (a.cpt == a.group) stages = stages >= 4 ? stages / 4 * 4 : 2;     // one commit per tile: the stage names the accumulator (stages % nbuf == 0)
    if ((a.cpt != a.group && stages < 3) || pw_smem_bytes(st->NC, K, st->KC, stages, a.group) > smem_max)
        return yq::fail("tcgen05 pointwise flavour (patch mode): the filter bank does not fit shared memory");
    a.nbuf = (4 * st->NC <= 512 && !(a.cpt == a.group && stages == 2)) ? 4 : 2;
    a.tmem_cols = 32;
    while (a.tmem_cols < a.nbuf * st->NC) a.tmem_cols *= 2;
    a.wc = l->cs_out / (PW_GROUPS / a.nbuf);
    a.rowb = a.wc < 64 ? a.wc : 64;
    if (a.wc % 32 || (a.wc > 64 && a.wc % 64)) return yq::fail("tcgen05 pointwise flavour (patch mode): %d channels per epilogue warp", a.wc);
    a.store_u8 = 1;
    a.a_stages = stages;
    const int smem = pw_smem_bytes(st->NC, K, st->KC, stages, a.group);
    a.trace = getenv("YQ_PW_TRACE") && atoi(getenv("YQ_PW_TRACE")) ? 1 : 0;
    PwState::Key key{in, (const void *)(uintptr_t)(in_geom->pitch_w * 65536 + go.pitch_w), out_u8, batch * 4 + 2 + (go.pad ? 1 : 0)};
    auto it = st->maps.find(key);
    if (it == st->maps.end()) {
        if (st->maps.size() > 64) st->maps.clear();
        PwState::Maps m;
        if (st->pair) {
            // pixel pairs from the halo pixel (-1, -1) on: 64 contiguous bytes each, consecutive in x, every second row in y
            EncodeTiledFn enc = pw_get_encode();
            const uint8_t *base = in + ((size_t)(in_geom->pad - 1) * in_geom->pitch_w + (in_geom->pad - 1)) * 32;
            cuuint64_t dims[4] = {64, (cuuint64_t)((l->w + 3) / 2), (cuuint64_t)(l->h + 2), (cuuint64_t)batch};
            cuuint64_t strides[3] = {64, (cuuint64_t)in_geom->pitch_w * 32, (cuuint64_t)in_geom->rows_h * in_geom->pitch_w * 32};
            cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)(TH * 2), 1};
            cuuint32_t es[4] = {1, 1, 2, 1};
            CUresult r = enc(&m.a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<uint8_t *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return yq::fail("cuTensorMapEncodeTiled(pixel pairs %dx%d pitch %d) failed: %d", l->h, l->w, in_geom->pitch_w, (int)r);
        } else if (pw_encode_nhwc(&m.a, in, in_geom, l->pad, batch, l->h, l->w, l->cs_in, st->KC, TW, TH, l->stride)) return -1;
        // the store box is one epilogue warp's pass: its 32 / TW rows of the patch x rowb channels
        if (pw_encode_nhwc(&m.o, out_u8, &go, 0, batch, l->out_h, l->out_w, l->cs_out, a.rowb, TW, 32 / TW, 1)) return -1;
        m.a2 = m.a;
        it = st->maps.emplace(key, m).first;
    }
    const int grid = a.num_tiles < n_sm ? a.num_tiles : n_sm;
    const CUtensorMap &tmA = it->second.a, &tmO = it->second.o, &tmA2 = it->second.a2;
    if (st->pair) return pw_launch<64, 2>(st, tmA, tmO, tmA2, a, smem, grid, stream);
    if (st->KC == 128) return pw_launch<128, 1>(st, tmA, tmO, tmA2, a, smem, grid, stream);
    if (st->KC == 64) return pw_launch<64, 1>(st, tmA, tmO, tmA2, a, smem, grid, stream);
    return pw_launch<32, 1>(st, tmA, tmO, tmA2, a, smem, grid, stream);
}

// YQ_PW_TRACE=1: the event times of the last traced launch (64 events x up to 256 CTAs, ns; tools/probes/pw_trace.py)
extern "C" __attribute__((visibility("default"))) int yq_debug_pw_trace(void *host, size_t bytes)
{
    return cudaMemcpyFromSymbol(host, yq_pw_trace, bytes < sizeof(yq_pw_trace) ? bytes : sizeof(yq_pw_trace)) == cudaSuccess ? 0 : -1;
}
