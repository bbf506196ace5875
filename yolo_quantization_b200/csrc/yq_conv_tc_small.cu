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

#include <type_traits>
#include <vector>

#include "yq_common.h"
#include "yq_epilogue.cuh"
#include "yq_tc_ptx.cuh"

using namespace yqtc;

namespace {

constexpr int SM_THREADS = 128;            // thread = im2col row = TMEM lane = output pixel of the tile
constexpr int TAPS = 9;
constexpr int TILE_W = 16, TILE_H = 8;     // 128 output pixels per tile; power-of-two so row -> (x,y) is shifts/masks

struct SmallArgs {
    const uint8_t *in;
    uint8_t *out;          // NHWC conv output, or nullptr when only the pooled output is wanted
    uint8_t *out_pool;     // NHWC 2x2/stride-2 max-pooled output (fused maxpool_layer.c:109-153), or nullptr
    float *out_f32;
    int32_t *out_acc;
    const uint8_t *wimg;   // pre-swizzled shared-memory image of the filter bank
    yq::EpiParams ep;
    int B, H, W, C, OH, OW, N, CSO, stride, pad, zp_in;
    int tiles_x, tiles_y, num_tiles, PH, PW;
    int OHP, OWP, opad;    // geometry of the conv output tensor (rows per image, pixels per row, halo width): OH, OW, 0 when plain
    // per-channel parameters live in kernel-parameter (constant-bank) space: with the channel loop fully unrolled
    // they become immediate constant operands of the epilogue's IMAD / SHF instructions (no loads, no registers)
    int4 cq[64];       // {bias, zw, 2*M0, shift}
    double mc[64];     // M_value * 2^-s (FP64 slow path)
};

template <int CS>
struct SmallGeom {
    static constexpr int K = TAPS * CS;
    static constexpr int KPAD = (K + 31) / 32 * 32;
    static constexpr int P = KPAD / 32;          // 32-byte K panels = MMA K steps
    static constexpr int A_BYTES = P * 128 * 32;
    static constexpr int NCHUNK = 2 * P;         // 16-byte chunks per im2col row
    static constexpr int NREAL = CS == 4 ? 3 : TAPS * (CS / 16);   // chunks that carry data (the rest are zero)
};

template <int CS, int BN>
struct SmallSmem {
    using G = SmallGeom<CS>;
    static constexpr int A_OFF = 0;                       // two A buffers
    static constexpr int B_OFF = 2 * G::A_BYTES;
    static constexpr int B_BYTES = G::P * BN * 32;
    static constexpr int ONES_OFF = B_OFF + B_BYTES;
    static constexpr int BAR_OFF = ONES_OFF + 512;
    static constexpr int TOTAL = BAR_OFF + 64;
};

// two accumulators of BN + 16 columns each
template <int BN>
__host__ __device__ constexpr int small_tmem_cols() { return 2 * (BN + 16) <= 64 ? 64 : (2 * (BN + 16) <= 128 ? 128 : 256); }

template <int ACTM, int NV>
__device__ __forceinline__ void epi_chunk_small(int sat, const uint32_t (&v)[NV], int nsa, const int4 *cq, const double *mc, int zo,
                                                uint32_t (&packed)[NV / 4], uint32_t xlim)
{
    int extra[NV];   // unused (HAS_EXTRA = false): padded taps already carry zp_in
    if (sat) yq::requant_chunk<ACTM, true, NV, false>(v, nsa, extra, cq, mc, zo, packed, xlim);
    else yq::requant_chunk<ACTM, false, NV, false>(v, nsa, extra, cq, mc, zo, packed, xlim);
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct TilePos { int n, ty, tx; };

// Tiles are visited with stride gridDim.x; (n, ty, tx) advance incrementally (no divisions in the loops).
struct TileWalker {
    int tiles_x, tiles_y, step_n, step_y, step_x;
    __device__ __forceinline__ void init(const SmallArgs &a, int step, int first, TilePos &t)
    {
        const int per_img = a.tiles_x * a.tiles_y;
        tiles_x = a.tiles_x; tiles_y = a.tiles_y;
        step_n = step / per_img; step_y = (step % per_img) / a.tiles_x; step_x = step % a.tiles_x;
        t.n = first / per_img; t.ty = (first % per_img) / a.tiles_x; t.tx = first % a.tiles_x;
    }
    __device__ __forceinline__ void advance(TilePos &t) const
    {
        t.tx += step_x;
        if (t.tx >= tiles_x) { t.tx -= tiles_x; ++t.ty; }
        t.ty += step_y;
        if (t.ty >= tiles_y) { t.ty -= tiles_y; ++t.n; }
        t.n += step_n;
    }
};

// SLOW = the variant that also serves the int32 / float side outputs and the saturate switch (parity checks,
// quant_stop heads); the production variant (SLOW = false) carries none of that code.
//
// PQ = "pool-first quad" production variant (RELU6, pooled output only).  Tile rows are ordered so that the four pixels
// of a 2x2 pooling window sit at TMEM lanes {i, i+8, i+16, i+24} of one warp's lane quarter; the 16x256b tcgen05.ld
// shape then hands ONE thread all four pixels of a window for BN/4 channels (no shuffles).  The thread max-reduces the
// raw accumulators first and requantizes the winner only: requantize + RELU6 is monotone in the accumulator, so
//     max_p u8(f(x_p)) == u8(f(max_p x_p))   as long as f(max) <= 255 (no uint8 wrap) and max < 2^22 (integer form exact);
// otherwise (rare) the window is redone pixel by pixel in FP64 form with the reference's wrap.  Filter rows are
// permuted on the host so that thread q = lane & 3 owns channels [q*BN/4, (q+1)*BN/4): its packed bytes are one
// contiguous BN/4-byte store and a warp writes 8 pooled pixels = 8*BN contiguous bytes.
template <int CS, int BN, int ACTM, bool SLOW, bool PQ>
__global__ void __launch_bounds__(SM_THREADS, (CS == 32 ? 2 : (CS == 16 ? 4 : 6))) conv_u8_tc_small_kernel(const __grid_constant__ SmallArgs a)
{
    static_assert(!PQ || (ACTM == 0 && !SLOW), "the pool-first variant exists for RELU6 production launches only");
    using G = SmallGeom<CS>;
    using L = SmallSmem<CS, BN>;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    uint8_t *sB = smem + L::B_OFF;
    uint64_t *mma_done = (uint64_t *)(smem + L::BAR_OFF);   // [2]
    uint32_t *tmem_slot = (uint32_t *)(mma_done + 2);

    const int r = threadIdx.x;            // A-tile row == TMEM lane == pixel within the tile
    const int warp = r >> 5, lane = r & 31;
    // pixel of the 16x8 tile this thread's im2col row / TMEM lane stands for
    const int wi = PQ ? 2 * (lane & 7) + ((lane >> 3) & 1) : (r & (TILE_W - 1));
    const int hi = PQ ? 2 * warp + (lane >> 4) : (r >> 4);

    // ---- one-time setup: resident filter bank, ones tile, barriers, TMEM
    for (int i = r; i < L::B_BYTES / 16; i += SM_THREADS)
        reinterpret_cast<uint4 *>(sB)[i] = __ldg(reinterpret_cast<const uint4 *>(a.wimg) + i);
    for (int i = r; i < 512 / 4; i += SM_THREADS) reinterpret_cast<uint32_t *>(smem + L::ONES_OFF)[i] = 0x01010101u;
    if (r == 0) {
        mbar_init(&mma_done[0], 1);
        mbar_init(&mma_done[1], 1);
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
    const int swz = (r >> 2) & 1;          // SWIZZLE_32B: 16-byte chunk index ^= address bit 7 = (row >> 2) & 1
    // index arithmetic = per-thread constant part (computed once) + per-tile uniform part
    const int thr_in_off = (hi * a.stride * a.W + wi * a.stride) * CS;
    const int row_pitch = a.W * CS;
    const int thr_out_off = (hi * a.OWP + wi) * a.CSO;
    const int thr_pool_off = ((hi >> 1) * a.PW + (wi >> 1)) * a.CSO;
    uint8_t *const sA_thr = smem + L::A_OFF + r * 32;

    const int first = blockIdx.x, step = gridDim.x;
    const int cnt = first < a.num_tiles ? (a.num_tiles - first + step - 1) / step : 0;
    TileWalker walk;
    TilePos lp;                          // position of the next tile to LOAD
    walk.init(a, step, first, lp);
    auto pack_pos = [](const TilePos &t) -> uint32_t { return ((uint32_t)t.n << 16) | ((uint32_t)t.ty << 8) | (uint32_t)t.tx; };

    // ---- gather this thread's im2col row for the tile at `t` straight into the swizzled A buffer with cp.async
    //      (no staging registers; the copies stay in flight while the previous tile's MMA and epilogue run).
    //      chunk g (16 bytes) of the K axis -> panel g/2, half g%2, SWIZZLE_32B.  Zero K-padding was written once.
    auto chunk_addr = [&](int buf, int g) -> uint32_t {
        return smem_u32(sA_thr + buf * G::A_BYTES + (g >> 1) * 4096 + (((g & 1) ^ swz) << 4));
    };
    auto issue_row = [&](const TilePos &t, int buf) {
        const int y0 = t.ty * TILE_H * a.stride - a.pad, x0 = t.tx * TILE_W * a.stride - a.pad;   // first tap of the tile
        const uint8_t *row0 = a.in + ((long long)(t.n * a.H + y0) * a.W + x0) * CS + thr_in_off;
        const uint8_t *rows[3] = {row0, row0 + row_pitch, row0 + 2 * row_pitch};
        // interior tile: every tap of every pixel is inside the image and every pixel inside the output
        const bool interior = y0 >= 0 && x0 >= 0 && y0 + (TILE_H - 1) * a.stride + 2 < a.H && x0 + (TILE_W - 1) * a.stride + 2 < a.W &&
                              t.ty * TILE_H + TILE_H <= a.OH && t.tx * TILE_W + TILE_W <= a.OW;
        if (interior) {
            if (CS == 4) {
#pragma unroll
                for (int tap = 0; tap < TAPS; ++tap) cp_async4(chunk_addr(buf, tap / 4) + (tap % 4) * 4, rows[tap / 3] + (tap % 3) * 4);
            } else {
                constexpr int CPT = CS / 16;
#pragma unroll
                for (int g = 0; g < G::NREAL; ++g) {
                    const int tap = g / CPT;
                    cp_async16(chunk_addr(buf, g), rows[tap / 3] + (tap % 3) * CS + (g % CPT) * 16);
                }
            }
            return;
        }
        const int oy = t.ty * TILE_H + hi, ox = t.tx * TILE_W + wi;
        const bool valid = ox < a.OW && oy < a.OH;
        const int iy0 = y0 + hi * a.stride, ix0 = x0 + wi * a.stride;
        bool vy[3], vx[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            vy[k] = valid && (unsigned)(iy0 + k) < (unsigned)a.H;
            vx[k] = (unsigned)(ix0 + k) < (unsigned)a.W;
        }
        if (CS == 4) {
            const uint32_t fw = valid ? fill_word(0) : 0u;
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
                const uint32_t dst = chunk_addr(buf, tap / 4) + (tap % 4) * 4;
                if (vy[tap / 3] && vx[tap % 3]) cp_async4(dst, rows[tap / 3] + (tap % 3) * 4);
                else st_shared32(dst, fw);
            }
        } else {
            constexpr int CPT = CS / 16;   // 16-byte chunks per tap
#pragma unroll
            for (int g = 0; g < G::NREAL; ++g) {
                const int tap = g / CPT, sub = g % CPT;
                const uint32_t dst = chunk_addr(buf, g);
                if (vy[tap / 3] && vx[tap % 3]) cp_async16(dst, rows[tap / 3] + (tap % 3) * CS + sub * 16);
                else if (valid) st_shared128(dst, make_uint4(fill_word(sub * 16), fill_word(sub * 16 + 4), fill_word(sub * 16 + 8), fill_word(sub * 16 + 12)));
                else st_shared128(dst, make_uint4(0, 0, 0, 0));
            }
        }
    };
    auto issue_mma = [&](int buf) {   // one thread issues the MMAs of a whole tile
        constexpr uint32_t idesc_main = make_idesc(BN);
        constexpr uint32_t idesc_ones = make_idesc(16);
        const uint64_t d_ones = make_desc<32>(smem_u32(smem + L::ONES_OFF));
        const uint32_t a0 = smem_u32(smem + L::A_OFF + buf * G::A_BYTES), b0 = smem_u32(sB);
        const uint32_t acc = tmem_base + buf * (BN + 16);
#pragma unroll
        for (int p = 0; p < G::P; ++p) {
            const uint64_t da = make_desc<32>(a0 + p * 4096);
            umma_i8(acc, da, make_desc<32>(b0 + p * BN * 32), idesc_main, p ? 1u : 0u);
            umma_i8(acc + BN, da, d_ones, idesc_ones, p ? 1u : 0u);   // sum of activations
        }
        umma_commit(&mma_done[buf]);
    };

    // zero the K padding of both A buffers once (cp.async only ever writes the data chunks)
    for (int buf = 0; buf < 2; ++buf) {
        if (CS == 4) {
#pragma unroll
            for (int g = 2; g < G::NCHUNK; ++g) st_shared128(chunk_addr(buf, g), make_uint4(0, 0, 0, 0));   // taps 8..15: tap 8 is rewritten per tile
        } else {
#pragma unroll
            for (int g = G::NREAL; g < G::NCHUNK; ++g) st_shared128(chunk_addr(buf, g), make_uint4(0, 0, 0, 0));
        }
    }

    uint32_t q0 = 0, q1 = 0;             // packed positions of the tiles whose epilogues are pending (q0 = oldest)
    uint32_t phase_bits = 0;
    if (cnt > 0) {
        issue_row(lp, 0);
        q0 = pack_pos(lp);
        walk.advance(lp);
        cp_async_wait_all();
        fence_proxy_async();     // generic-proxy smem writes -> visible to the tensor core (async proxy)
        tc_fence_before();
        __syncthreads();
        if (r == 0) {
            tc_fence_after();
            issue_mma(0);
        }
        if (cnt > 1) {
            issue_row(lp, 1);
            q1 = pack_pos(lp);
            walk.advance(lp);
        }
    }
    const bool side = SLOW && ((a.out_acc != nullptr) || (a.out_f32 != nullptr));
    // PQ: this thread's BN/4 channels (slot 4j+cc <-> channel qq*BN/4 + 4j + cc), held in registers for the whole launch
    constexpr int NPQ = PQ ? BN / 4 : 1;
    const int qi = lane >> 2, qq = lane & 3;
    int pq_bias[NPQ], pq_zw[NPQ], pq_sh[NPQ];
    uint32_t pq_m2[NPQ];
    if constexpr (PQ) {
#pragma unroll
        for (int k = 0; k < NPQ; ++k) {
            const int4 c = a.cq[qq * NPQ + k];
            pq_bias[k] = c.x; pq_zw[k] = c.y; pq_m2[k] = (uint32_t)c.z; pq_sh[k] = c.w;
        }
    }
    for (int i = 0; i < cnt; ++i) {
        const int b = i & 1;
        const uint32_t qe = q0;          // tile whose epilogue runs in this iteration
        q0 = q1;
        if (i + 1 < cnt) {
            // tile i+1's row has been landing in A[b^1] since the previous iteration; acc[b^1] was drained by every
            // thread before it reaches this barrier
            cp_async_wait_all();
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (r == 0) {
                tc_fence_after();
                issue_mma(b ^ 1);
            }
        }
        mbar_wait(&mma_done[b], (phase_bits >> b) & 1u);   // MMA(i) complete: acc[b] ready, A[b] free
        phase_bits ^= 1u << b;
        tc_fence_after();
        if (i + 2 < cnt) {
            issue_row(lp, b);
            q1 = pack_pos(lp);
            walk.advance(lp);
        }

        const int tn = (int)(qe >> 16), tty = (int)((qe >> 8) & 255u), ttx = (int)(qe & 255u);
        if constexpr (PQ) {
            // ---- pool-first epilogue of tile i: this thread owns the 2x2 window (qi, warp) of the tile for channels qq*BN/4 ..
            const int ppx = ttx * (TILE_W / 2) + qi, ppy = tty * (TILE_H / 2) + warp;
            const bool is_edge = ttx * TILE_W + TILE_W > a.OW || tty * TILE_H + TILE_H > a.OH;   // uniform per tile
            const uint32_t tq = trow + b * (BN + 16);
            // EDGE tiles hang over the right / bottom border: pixels outside the conv output take no part in the max
            auto pq_epilogue = [&](auto edge_tag) {
                constexpr bool edge = decltype(edge_tag)::value;
                bool pv[4] = {true, true, true, true};      // window pixel p = 2*dy + dx inside the conv output?
                if (edge) {
#pragma unroll
                    for (int p = 0; p < 4; ++p) pv[p] = 2 * ppx + (p & 1) < a.OW && 2 * ppy + (p >> 1) < a.OH;
                }
                int nsa[4];
                uint32_t words[BN / 16];
#pragma unroll
                for (int j = 0; j < BN / 16; ++j) {
                    uint32_t v0[8], v1[8];                  // lanes {qi, qi+8} and {qi+16, qi+24}
                    if (j == 0) {
                        uint32_t s0[4], s1[4];
                        tmem_ldq_first(tq + BN, tq, s0, s1, v0, v1);
                        nsa[0] = -(int)s0[0]; nsa[1] = -(int)s0[2]; nsa[2] = -(int)s1[0]; nsa[3] = -(int)s1[2];
                    } else {
                        tmem_ldq(tq + 16 * j, v0, v1);
                    }
                    int rr[4];
                    uint32_t orx = 0, orr = 0;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int k = 4 * j + cc, ri = 4 * (cc >> 1) + (cc & 1);   // register of (dx = 0, this channel)
                        int x0 = pq_zw[k] * nsa[0] + (int)v0[ri], x1 = pq_zw[k] * nsa[1] + (int)v0[ri + 2];
                        int x2 = pq_zw[k] * nsa[2] + (int)v1[ri], x3 = pq_zw[k] * nsa[3] + (int)v1[ri + 2];
                        if (edge) {
                            x0 = pv[0] ? x0 : INT_MIN; x1 = pv[1] ? x1 : INT_MIN; x2 = pv[2] ? x2 : INT_MIN; x3 = pv[3] ? x3 : INT_MIN;
                        }
                        const int m = max(max(x0, x1), max(x2, x3));
                        const uint32_t xq = (uint32_t)max(m + pq_bias[k], 0);   // (a window with no valid pixel is never stored)
                        rr[cc] = (int)(__umulhi(xq, pq_m2[k]) >> pq_sh[k]) + a.ep.zp_out;
                        orx |= xq; orr |= (uint32_t)rr[cc];
                    }
                    if (orx >= (1u << 22) || orr > 255u) {
                        // rare: the integer form may round differently from the reference's double multiply, or a byte wraps:
                        // redo the window pixel by pixel in FP64 form (convolutional_layer.c:732-749), then pool the bytes
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            const int k = 4 * j + cc, ri = 4 * (cc >> 1) + (cc & 1);
                            const int4 c = a.cq[qq * NPQ + k];
                            const double mcd = a.mc[qq * NPQ + k];
                            const int xs[4] = {(int)v0[ri], (int)v0[ri + 2], (int)v1[ri], (int)v1[ri + 2]};
                            int best = 0;
#pragma unroll
                            for (int p = 0; p < 4; ++p) {
                                const int x = c.y * nsa[p] + xs[p] + c.x;
                                const int q = max(__double2int_rz(__dmul_rn((double)x, mcd)), 0);
                                const int u8 = (q + a.ep.zp_out) & 255;
                                if (pv[p]) best = max(best, u8);
                            }
                            rr[cc] = best;
                        }
                    }
                    words[j] = __byte_perm(__byte_perm((uint32_t)rr[0], (uint32_t)rr[1], 0x0040), __byte_perm((uint32_t)rr[2], (uint32_t)rr[3], 0x0040), 0x5410);
                }
                if (!edge || (ppx < a.PW && ppy < a.PH)) {
                    uint8_t *dst = a.out_pool + ((size_t)(tn * a.PH + ppy) * a.PW + ppx) * a.CSO + qq * (BN / 4);
                    if constexpr (BN == 16) *reinterpret_cast<uint32_t *>(dst) = words[0];
                    else if constexpr (BN == 32) *reinterpret_cast<uint2 *>(dst) = make_uint2(words[0], words[1]);
                    else *reinterpret_cast<uint4 *>(dst) = make_uint4(words[0], words[1], words[2], words[3]);
                }
            };
            if (is_edge) pq_epilogue(std::true_type{});
            else pq_epilogue(std::false_type{});
            tc_fence_before();
            continue;
        }
        // ---- epilogue of tile i: this thread owns TMEM lane r = its pixel
        const int ox = ttx * TILE_W + wi, oy = tty * TILE_H + hi;
        const bool valid = ox < a.OW && oy < a.OH;
        const uint32_t tacc = trow + b * (BN + 16);
        const int nsa = -(int)tmem_ld1(tacc + BN);   // minus the pixel's activation sum (ones-tile columns)
        uint8_t *const orow = a.out + ((size_t)(tn * a.OHP + tty * TILE_H + a.opad) * a.OWP + ttx * TILE_W + a.opad) * a.CSO + thr_out_off;
        // fused 2x2 stride-2 max-pool: partners are lane^1 (x neighbour) and lane^16 (next row); out-of-image pixels count as 0
        uint8_t *const prow = a.out_pool + ((size_t)(tn * a.PH + tty * (TILE_H / 2)) * a.PW + ttx * (TILE_W / 2)) * a.CSO + thr_pool_off;
        const bool pool_writer = a.out_pool && ((lane & 17) == 0) && (ox >> 1) < a.PW && (oy >> 1) < a.PH;
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tacc + c0, v);
            uint32_t packed[4];
            epi_chunk_small<ACTM, 16>(SLOW ? a.ep.saturate : 0, v, nsa, a.cq + c0, a.mc + c0, a.ep.zp_out, packed, a.ep.xlim);
            yq::mask_pad_channels<16>(packed, a.N - c0);
            if (SLOW && side && valid) {   // parity / quant_stop side outputs (not on the throughput path)
                const size_t pix = ((size_t)tn * a.OH + oy) * a.OW + ox;
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                    const int oc = c0 + jj;
                    if (oc < a.N) {
                        if (a.out_acc) a.out_acc[pix * a.CSO + oc] = (int)v[jj] + a.cq[oc].y * nsa;
                        if (a.out_f32) {
                            const uint8_t u = (uint8_t)(packed[jj / 4] >> (8 * (jj % 4)));
                            a.out_f32[((size_t)tn * a.N + oc) * a.OH * a.OW + (size_t)oy * a.OW + ox] = yq::dequant_f32(a.ep, u);
                        }
                    }
                }
            }
            if (a.out && valid) *reinterpret_cast<uint4 *>(orow + c0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            if (a.out_pool) {
                uint32_t pm[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t m = valid ? packed[k] : 0u;
                    m = __vmaxu4(m, __shfl_xor_sync(0xffffffffu, m, 1));
                    m = __vmaxu4(m, __shfl_xor_sync(0xffffffffu, m, 16));
                    pm[k] = m;
                }
                if (pool_writer) *reinterpret_cast<uint4 *>(prow + c0) = make_uint4(pm[0], pm[1], pm[2], pm[3]);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<small_tmem_cols<BN>()>(tmem_base);
    }
}

struct SmallState {
    int CS, BN, ctas_per_sm;
    uint8_t *wimg = nullptr;
    uint8_t *wimg_pq = nullptr;   // filter rows in the pool-first variant's TMEM-column order, or nullptr
};

// pool-first variant: TMEM column c = 16j + 8k + 2q + e  holds channel  q*BN/4 + 4j + 2k + e
int pq_channel_of_column(int c, int BN) { return ((c % 8) / 2) * (BN / 4) + 4 * (c / 16) + 2 * ((c % 16) / 8) + (c % 2); }

template <int CS, int BN, int ACTM, bool SLOW, bool PQ = false>
int launch_small(SmallState *st, const SmallArgs &a, cudaStream_t stream)
{
    using L = SmallSmem<CS, BN>;
    const int smem = L::TOTAL + 1024;
    auto kern = conv_u8_tc_small_kernel<CS, BN, ACTM, SLOW, PQ>;
    // per device: the shared-memory opt-in and the CTAs-per-SM count derived from this device's limits
    if (yq::ensure_dynamic_smem((const void *)kern, smem)) return -1;
    const int n_sm = yq::device_sm_count();
    int ctas_per_sm = 0;
    if (!yq::memo_get((const void *)kern, &ctas_per_sm)) {
        // cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for kernels that allocate tensor memory (measured on
        // B200), although the hardware co-schedules as many CTAs as smem / registers / TMEM columns allow: count by hand.
        const int smem_sm = yq::device_smem_per_sm();
        cudaFuncAttributes fa;
        YQ_CUDA(cudaFuncGetAttributes(&fa, kern));
        const int by_smem = smem_sm / (smem + 1024 + (int)fa.sharedSizeBytes);
        const int regs_per_cta = ((fa.numRegs + 7) / 8 * 8) * SM_THREADS;
        const int by_regs = 65536 / (regs_per_cta > 0 ? regs_per_cta : 1);
        const int by_threads = 2048 / SM_THREADS;
        int occ = by_smem < by_regs ? by_smem : by_regs;
        if (by_threads < occ) occ = by_threads;
        const int tmem_limit = 512 / small_tmem_cols<BN>();   // every resident CTA must own its TMEM columns
        ctas_per_sm = occ < tmem_limit ? occ : tmem_limit;
        if (getenv("YQ_DEBUG")) fprintf(stderr, "yq: small<%d,%d,%d,%d> regs=%d occ=%d tmem_limit=%d n_sm=%d smem=%d\n", CS, BN, ACTM, (int)SLOW, fa.numRegs, occ, tmem_limit, n_sm, smem);
        if (ctas_per_sm < 1 || n_sm <= 0) return yq::fail("conv_u8_tc_small_kernel<%d,%d> does not fit on an SM", CS, BN);
        yq::memo_put((const void *)kern, ctas_per_sm);
    }
    int grid = n_sm * ctas_per_sm;
    if (grid > a.num_tiles) grid = a.num_tiles;
    conv_u8_tc_small_kernel<CS, BN, ACTM, SLOW, PQ><<<grid, SM_THREADS, smem, stream>>>(a);
    YQ_CHECK_LAUNCH();
    return 0;
}

template <int CS, int BN>
int launch_small_act(SmallState *st, const SmallArgs &a, int actm, cudaStream_t stream)
{
    const bool slow = a.out_acc || a.out_f32 || a.ep.saturate;
    if (!slow && actm == 0 && a.out_pool && !a.out && st->wimg_pq && a.N == BN && a.CSO == BN) {
        SmallArgs b = a;
        b.wimg = st->wimg_pq;
        return launch_small<CS, BN, 0, false, true>(st, b, stream);
    }
    if (slow) {
        if (actm == 0) return launch_small<CS, BN, 0, true>(st, a, stream);
        if (actm == 1) return launch_small<CS, BN, 1, true>(st, a, stream);
        return launch_small<CS, BN, 2, true>(st, a, stream);
    }
    if (actm == 0) return launch_small<CS, BN, 0, false>(st, a, stream);
    if (actm == 1) return launch_small<CS, BN, 1, false>(st, a, stream);
    return launch_small<CS, BN, 2, false>(st, a, stream);
}

}  // namespace

int yq_tc_small_supported(const yq_conv_layer *l)
{
    if (!l->int_form || !l->fused_mult) return 0;   // the integer-form epilogue needs M0 * 2^-31 / 2^-s parameters
    if (l->size != 3 || l->pad != 1 || (l->stride != 1 && l->stride != 2)) return 0;
    if (!(l->cs_in == 4 || l->cs_in == 16 || l->cs_in == 32)) return 0;
    // instantiated (cs_in, cs_out) pairs -- see yq_tc_small_forward
    const int ci = l->cs_in, co = l->cs_out;
    return (ci == 4 && (co == 16 || co == 32 || co == 64)) || (ci == 16 && (co == 16 || co == 32)) || (ci == 32 && co == 64);
}

int yq_tc_small_prepare(yq_conv_layer *l, void **state)
{
    SmallState *st = new SmallState();
    st->CS = l->cs_in;
    st->BN = l->cs_out;
    const int K = TAPS * st->CS, KPAD = (K + 31) / 32 * 32, P = KPAD / 32;
    // shared-memory image: panel p = K bytes [32p, 32p+32) of every row; SWIZZLE_32B inside each 8-row atom
    std::vector<uint8_t> img;
    if (!yq::pack_fetch(l, "small", img) || img.size() != (size_t)P * st->BN * 32) {
        img.assign((size_t)P * st->BN * 32, 0);
        for (int oc = 0; oc < l->n; ++oc)
            for (int t = 0; t < TAPS; ++t)
                for (int ci = 0; ci < l->c; ++ci) {
                    const int k = t * st->CS + ci;
                    const int p = k / 32, c = (k % 32) / 16, b = k % 16;
                    img[(size_t)p * st->BN * 32 + oc * 32 + ((c ^ ((oc >> 2) & 1)) << 4) + b] = l->host_w[((size_t)oc * l->c + ci) * TAPS + t];
                }
        yq::pack_put(l, "small", img);
    }
    if (cudaMalloc((void **)&st->wimg, img.size()) != cudaSuccess || cudaMemcpy(st->wimg, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(st->wimg);
        delete st;
        return yq::fail("tcgen05 small-c flavour: weight upload failed");
    }
    if (l->n == st->BN) {   // the pool-first variant needs every TMEM column to be a real channel
        std::vector<uint8_t> pq(img.size(), 0);
        for (int p = 0; p < P; ++p)
            for (int col = 0; col < st->BN; ++col) {
                const int oc = pq_channel_of_column(col, st->BN);
                // row `col` of the panel keeps its own SWIZZLE_32B phase: move the two 16-byte chunks accordingly
                for (int c = 0; c < 2; ++c)
                    memcpy(&pq[(size_t)p * st->BN * 32 + col * 32 + ((c ^ ((col >> 2) & 1)) << 4)],
                           &img[(size_t)p * st->BN * 32 + oc * 32 + ((c ^ ((oc >> 2) & 1)) << 4)], 16);
            }
        if (cudaMalloc((void **)&st->wimg_pq, pq.size()) != cudaSuccess ||
            cudaMemcpy(st->wimg_pq, pq.data(), pq.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaFree(st->wimg);
            cudaFree(st->wimg_pq);
            delete st;
            return yq::fail("tcgen05 small-c flavour: weight upload failed");
        }
    }
    *state = st;
    return 0;
}

void yq_tc_small_free(void *state)
{
    SmallState *st = (SmallState *)state;
    if (!st) return;
    cudaFree(st->wimg);
    cudaFree(st->wimg_pq);
    delete st;
}

int yq_tc_small_forward(yq_conv_layer *l, void *state, const uint8_t *in_u8, uint8_t *out_u8, uint8_t *out_pool, float *out_f32,
                        int32_t *out_acc, int batch, cudaStream_t stream, const yq_act_geom *out_geom)
{
    SmallState *st = (SmallState *)state;
    SmallArgs a;
    memset(&a, 0, sizeof a);
    a.in = in_u8; a.out = out_u8; a.out_pool = out_pool; a.out_f32 = l->quant_stop_flag ? out_f32 : nullptr; a.out_acc = out_acc; a.wimg = st->wimg;
    a.ep = yq::make_epi(l);
    a.B = batch; a.H = l->h; a.W = l->w; a.C = l->c; a.OH = l->out_h; a.OW = l->out_w; a.N = l->n; a.CSO = l->cs_out;
    a.stride = l->stride; a.pad = l->pad; a.zp_in = l->zp_in;
    // the conv output may be a halo-padded tensor (only its interior is written); the pooled output stays plain
    a.OHP = out_geom ? out_geom->rows_h : l->out_h; a.OWP = out_geom ? out_geom->pitch_w : l->out_w; a.opad = out_geom ? out_geom->pad : 0;
    a.tiles_x = (l->out_w + TILE_W - 1) / TILE_W;
    a.tiles_y = (l->out_h + TILE_H - 1) / TILE_H;
    a.num_tiles = a.tiles_x * a.tiles_y * batch;
    a.PH = (l->out_h + 1) / 2;   // maxpool 2/2 with the default padding size-1: (h + 1 - 2)/2 + 1 (maxpool_layer.c:31-32)
    a.PW = (l->out_w + 1) / 2;
    if (!out_u8 && !out_pool) return yq::fail("conv: neither the conv output nor the pooled output was requested");
    const int p = l->n < 64 ? l->n : 64;
    memcpy(a.cq, l->host_chanq.data(), (size_t)p * 16);
    memcpy(a.mc, l->host_mcomb.data(), (size_t)p * 8);
    const int actm = yq::act_mode(l->activation);
#define YQ_SM(CS_, BN_) if (st->CS == CS_ && st->BN == BN_) return launch_small_act<CS_, BN_>(st, a, actm, stream)
    YQ_SM(4, 16); YQ_SM(4, 32); YQ_SM(4, 64);   // (4, 32): layer 0 of the full yolov3
    YQ_SM(16, 16); YQ_SM(16, 32);
    YQ_SM(32, 64);
#undef YQ_SM
    return yq::fail("tcgen05 small-c flavour: no instantiation for cs_in=%d cs_out=%d", st->CS, st->BN);
}
