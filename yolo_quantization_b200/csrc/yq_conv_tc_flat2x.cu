// yq_conv_tc_flat2x.cu -- the persistent flat-strip convolution of yq_conv_tc_flat2.cu on CTA PAIRS (tcgen05 cta_group::2).
//
// Measured on B200: with SS-mode M = 128 MMAs the tensor pipe of layer 12 runs at ~107 clocks per N = 128 MMA against a
// floor of 64, because every MMA reads 8 KB of operands from shared memory while TMA writes 2.5 KB more -- the SM's
// 128 B/clk of shared-memory bandwidth is the bound.  A cta_group::2 MMA (M = 256: 128 rows from each CTA of a cluster
// pair) takes its B operand HALF from each CTA's shared memory: every SM stores, fills and reads only 64 of the 128
// filter rows of a stage.
//
// Cluster of 2 CTAs (one per SM of a TPC); each CTA owns 256 positions (two 128-row tiles) of a 512-position cluster tile.
//   * producer warp (both CTAs): own activation patch -> local a_full; own 64 filter rows of every stage -> the LEADER's
//     b_full (cp.async.bulk.tensor ... cta_group::2 with the leader's mbarrier);
//   * MMA warp: the leader issues tcgen05.mma.cta_group::2 (M = 256, N = 128) for tile j = 0 and 1 of every stage and
//     releases stages / publishes accumulators with multicast commits to both CTAs; the peer's MMA warp only relays
//     "my patch has landed" to the leader;
//   * sum warps and epilogue warps: as in flat2, per CTA, on the CTA's own TMEM lanes; the peer's epilogue also arrives on the
//     leader's barrier so that the leader knows both halves of an accumulator pair are drained.
// Same tensors, geometry, arithmetic as yq_conv_tc_flat.cu / flat2.  Restates convolutional_layer.c:694-751.
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include <map>
#include <type_traits>
#include <vector>

#include "yq_common.h"
#include "yq_epilogue.cuh"
#include "yq_tc_ptx.cuh"

using namespace yqtc;

namespace {

constexpr int F2X_BN = 128;            // channels per tile of the two-tile form; the WIDE form has one tile of 256 channels
constexpr int F2X_EPI_WARPS = 16;
constexpr int F2X_SUM_WARPS = 2;
constexpr int F2X_THREADS = 32 * (2 + F2X_SUM_WARPS + F2X_EPI_WARPS);
constexpr int F2X_ASTAGES = 6;              // barrier slots; a launch uses a.a_stages of them: 2 for 3x3 layers (nine taps of MMAs per patch chunk cover the next
                                            // chunk's load), up to 6 for 1x1 layers (one tap per chunk: the loads run several chunks ahead of the tensor pipe)
constexpr int F2X_MAX_BSTAGES = 8;
constexpr int F2X_MAX_ROWS = 384;           // patch rows: 256 + 2W + 4 (W <= 61)

struct Flat2xArgs {
    yq::EpiParams ep;
    int32_t *out_acc;      // dense NHWC [B][H][W][CSO] (parity checks) or nullptr
    int N, CSO;
    int B, H, W, NP;       // NP = B*(H+1)*(W+1)
    int size, taps, cpt /* KC-chunks per tap */, CS;
    int q_off;             // first patch position relative to the pair's first position: -(pad*(W+1) + pad)
    int patch_rows, box_rows, a_stage_bytes, a_stages, b_stages;
    int m_pairs, num_tiles;   // tile t -> (n-tile t / m_pairs, position pair t % m_pairs)
    uint32_t halo_word;
    uint32_t magic_w, magic_h, magic_m;
    // the FOLLOWING quantized shortcut fused into the epilogue (extension layer, include/yq_b200.h): the `from` tensor in this
    // layer's own flat geometry and channel stride; the launch then stores the SHORTCUT's output
    int split;             // chunks [0, split) of the input channels come from the second tensor map (virtual route), the rest from the first
    const uint8_t *resid;
    yq::ShortcutParams sc;
    int debug;             // YQ_FLAT2_DEBUG experiments (results are garbage): 1 = skip weight loads of taps > 0, 2 = skip the epilogue math
};

// WIDE = false: two 128-position tiles x 128 channels per CTA and accumulator pair;  WIDE = true: one tile x 256 channels
// (N = 256 MMAs: the patch is read once per 256 channels -> ~105 instead of ~126 B/clk of shared-memory traffic per SM)
template <int KC, bool WIDE>
struct Flat2xSmem {
    static constexpr int BNT = WIDE ? 256 : 128;                  // channels per tile = N of the pair MMA
    static constexpr int TPC = WIDE ? 1 : 2;                      // tiles per CTA
    static constexpr int PPC = 128 * TPC;                         // positions per CTA; a cluster tile has 2 * PPC
    static constexpr int B_STAGE = (BNT / 2) * KC;                // each CTA of the pair holds half of the stage's filter rows
    static constexpr int OUT_BYTES = 128 * 128;                   // one staging tile [128 positions][128 channels]; there are two
    static constexpr int PARAM_BYTES = BNT * 24;
    static constexpr int SUM_BYTES = 2 * F2X_MAX_ROWS * 4;         // S[pair buffer][patch row]
};

__device__ __forceinline__ void f2x_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }

// RESID: the following quantized shortcut fused into the epilogue (launch-uniform: a template parameter, see yq_conv_tc_flat2.cu)
template <int KC, bool SLOW, bool WIDE, bool RESID>
__global__ void __launch_bounds__(F2X_THREADS, 1) conv_u8_tc_flat2x_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                     const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmA2,
                                                                     const Flat2xArgs a)
{
    using L = Flat2xSmem<KC, WIDE>;
    constexpr int BNT = L::BNT, TPC = L::TPC, PPC = L::PPC;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sA = smem;                                              // F2X_ASTAGES patch stages
    uint8_t *sB = sA + a.a_stages * a.a_stage_bytes;                  // a.b_stages weight stages
    uint8_t *sOut = sB + a.b_stages * L::B_STAGE;                    // two output staging tiles
    int4 *s_q = (int4 *)(sOut + 2 * L::OUT_BYTES);                   // {bias, zw, 2*M0, shift} of the current n-tile
    double *s_mc = (double *)(s_q + BNT);
    int *s_sum = (int *)(s_mc + BNT);                              // [2][F2X_MAX_ROWS]
    uint64_t *a_full = (uint64_t *)(s_sum + 2 * F2X_MAX_ROWS);
    uint64_t *a_empty = a_full + F2X_ASTAGES;
    uint64_t *b_full = a_empty + F2X_ASTAGES;
    uint64_t *b_empty = b_full + F2X_MAX_BSTAGES;
    uint64_t *acc_full = b_empty + F2X_MAX_BSTAGES;
    uint64_t *acc_empty = acc_full + 2;
    uint64_t *sum_full = acc_empty + 2;
    uint64_t *a_peer_full = sum_full + 2;                          // leader: the peer's patch stage has landed
    uint64_t *acc_peer_empty = a_peer_full + F2X_ASTAGES;         // leader: the peer's epilogue has drained the accumulator pair
    uint32_t *tmem_slot = (uint32_t *)(acc_peer_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = a.cpt, nbs = a.b_stages;
    const uint32_t rank = cluster_ctarank();                       // 0 = leader (issues the pair's MMAs)
    const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;         // cluster index / count

    if (threadIdx.x == 0) {
        for (int s = 0; s < F2X_ASTAGES; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1 + F2X_SUM_WARPS);     // the (multicast) MMA commit and each local sum warp release a patch stage
            mbar_init(&a_peer_full[s], 1);
        }
        for (int s = 0; s < F2X_MAX_BSTAGES; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], F2X_EPI_WARPS);
            mbar_init(&sum_full[s], F2X_SUM_WARPS);
            mbar_init(&acc_peer_empty[s], F2X_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc_2cta<512>(tmem_slot);
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
        if (a.split) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
    }
    tc_fence_before();
    cluster_sync_all();            // both CTAs' barriers are initialised before anyone signals across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    yq_pdl_wait_then_release();                             // no activation tensor was touched so far

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform control flow, one elected lane issues) =====================
        int sa = 0, s = 0;
        uint32_t pha = 0, phb = 0;
        for (int tile = cid; tile < a.num_tiles; tile += ncl) {
            const int nt = a.m_pairs == 1 ? tile : (int)__umulhi((uint32_t)tile, a.magic_m), mp = tile - nt * a.m_pairs;
            const int p0 = mp * 2 * PPC + (int)rank * PPC, oc0 = nt * BNT + (int)rank * (BNT / 2);
            for (int c = 0; c < chunks; ++c) {
                mbar_wait(&a_empty[sa], pha ^ 1);
                if (elect_one()) {
                    // the input is the channel concatenation [tmA2 | tmA] of two tensors of the same geometry (a route that is never
                    // materialised): the first a.split chunks come from tmA2; a.split = 0: one tensor
                    const CUtensorMap *src = c < a.split ? &tmA2 : &tmA;
                    const int cc = (c < a.split ? c : c - a.split) * KC;
                    mbar_expect_tx(&a_full[sa], (uint32_t)(2 * a.box_rows * KC));
                    tma_load_2d(sA + sa * a.a_stage_bytes, src, &a_full[sa], cc, p0 + a.q_off);
                    tma_load_2d(sA + sa * a.a_stage_bytes + a.box_rows * KC, src, &a_full[sa], cc, p0 + a.q_off + a.box_rows);
                }
                if (++sa == a.a_stages) { sa = 0; pha ^= 1; }
                for (int tap = 0; tap < a.taps; ++tap) {
                    mbar_wait(&b_empty[s], phb ^ 1);           // (released by the leader's multicast commit)
                    if (elect_one()) {
                        // my 64 filter rows of the stage; both halves complete on the LEADER's barrier
                        if (rank == 0) mbar_expect_tx(&b_full[s], (uint32_t)(BNT * KC));
                        tma_load_2d_2cta(sB + s * L::B_STAGE, &tmB, map_to_cta(&b_full[s], 0), tap * a.CS + c * KC, oc0);
                    }
                    if (++s == nbs) { s = 0; phb ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // ===================== MMA issuer (leader CTA): M = 256 across the pair =====================
            constexpr uint32_t idesc = make_idesc_m(256, BNT);
            const uint32_t row_step = (uint32_t)((a.W + 1 - a.size) * KC);
            int sa = 0, s = 0;
            uint32_t pha = 0, phb = 0, it = 0;
            for (int tile = cid; tile < a.num_tiles; tile += ncl, ++it) {
                const int pb = it & 1;
                mbar_wait(&acc_empty[pb], ((it >> 1) & 1) ^ 1);          // my epilogue has drained this pair of accumulators
                mbar_wait(&acc_peer_empty[pb], ((it >> 1) & 1) ^ 1);     // ... and so has the peer's
                tc_fence_after();
                const uint32_t acc0 = tmem_base + pb * 256, acc1 = acc0 + 128;     // (WIDE: one accumulator of 256 columns)
                uint32_t accumulate = 0;
                for (int c = 0; c < chunks; ++c) {
                    mbar_wait(&a_full[sa], pha);                         // my patch
                    mbar_wait(&a_peer_full[sa], pha);                    // the peer's patch (relayed by its warp 1)
                    uint32_t tap_addr = smem_u32(sA + sa * a.a_stage_bytes);    // row-shifted descriptor start of the current tap
                    int kx = 0;
                    for (int tap = 0; tap < a.taps; ++tap) {
                        mbar_wait(&b_full[s], phb);                      // both halves of the weight stage
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t da0 = make_desc<KC>(tap_addr), da1 = make_desc<KC>(tap_addr + 128 * KC);
                            const uint64_t db = make_desc<KC>(smem_u32(sB + s * L::B_STAGE));
#pragma unroll
                            for (int k = 0; k < KC / 32; ++k) {
                                umma_i8_2cta(acc0, da0 + 2 * k, db + 2 * k, idesc, (k == 0) ? accumulate : 1u);
                                if (!WIDE) umma_i8_2cta(acc1, da1 + 2 * k, db + 2 * k, idesc, (k == 0) ? accumulate : 1u);
                            }
                            umma_commit_2cta(&b_empty[s], 3);
                        }
                        accumulate = 1;
                        tap_addr += KC;
                        if (++kx == a.size) { kx = 0; tap_addr += row_step; }
                        if (++s == nbs) { s = 0; phb ^= 1; }
                    }
                    if (elect_one()) umma_commit_2cta(&a_empty[sa], 3);
                    if (++sa == a.a_stages) { sa = 0; pha ^= 1; }
                }
                if (elect_one()) umma_commit_2cta(&acc_full[pb], 3);
            }
        } else {
            // ===================== peer CTA: tell the leader when my patch stages have landed =====================
            const uint32_t remote0 = map_to_cta(&a_peer_full[0], 0);       // (the barriers are consecutive 8-byte words)
            int sa = 0;
            uint32_t pha = 0;
            for (int tile = cid; tile < a.num_tiles; tile += ncl) {
                for (int c = 0; c < chunks; ++c) {
                    mbar_wait(&a_full[sa], pha);
                    if (elect_one()) mbar_arrive_cluster(remote0 + 8u * (uint32_t)sa);
                    if (++sa == a.a_stages) { sa = 0; pha ^= 1; }
                }
            }
        }
    } else if (warp < 2 + F2X_SUM_WARPS) {
        // ===================== activation sums: S[row] = sum of the patch row's bytes over all channel chunks =====================
        const int st = threadIdx.x - 64;                              // 0 .. 63
        constexpr int RPT = (F2X_MAX_ROWS + 32 * F2X_SUM_WARPS - 1) / (32 * F2X_SUM_WARPS);   // rows per thread
        int sa = 0;
        uint32_t pha = 0, it = 0;
        for (int tile = cid; tile < a.num_tiles; tile += ncl, ++it) {
            const int pb = it & 1;
            int acc[RPT];
#pragma unroll
            for (int k = 0; k < RPT; ++k) acc[k] = 0;
            for (int c = 0; c < chunks; ++c) {
                mbar_wait(&a_full[sa], pha);
                const uint8_t *patch = sA + sa * a.a_stage_bytes;
#pragma unroll
                for (int k = 0; k < RPT; ++k) {
                    const int row = st + k * 32 * F2X_SUM_WARPS;
                    if (row < a.patch_rows) {
                        const uint4 *rp = reinterpret_cast<const uint4 *>(patch + (size_t)row * KC);
                        unsigned sum = 0;
#pragma unroll
                        for (int j = 0; j < KC / 16; ++j) {
                            const uint4 v = rp[(j + st) % (KC / 16)];     // rotate the start chunk: fewer bank conflicts; the sum does not care
                            sum = __dp4a(v.x, 0x01010101u, sum);
                            sum = __dp4a(v.y, 0x01010101u, sum);
                            sum = __dp4a(v.z, 0x01010101u, sum);
                            sum = __dp4a(v.w, 0x01010101u, sum);
                        }
                        acc[k] += (int)sum;
                    }
                }
                __syncwarp();
                if (lane == 0) f2x_arrive(&a_empty[sa]);
                if (++sa == a.a_stages) { sa = 0; pha ^= 1; }
            }
            mbar_wait(&acc_empty[pb], ((it >> 1) & 1) ^ 1);           // S[pb] was consumed by the epilogue of pair it-2
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const int row = st + k * 32 * F2X_SUM_WARPS;
                if (row < a.patch_rows) s_sum[pb * F2X_MAX_ROWS + row] = acc[k];
            }
            __syncwarp();
            if (lane == 0) f2x_arrive(&sum_full[pb]);
        }
    } else {
        // ===================== epilogue: 16 warps, each on its own (see yq_conv_tc_flat2.cu) =====================
        // Warp (q, g): q = warp & 3 = its TMEM lane quarter (32 positions); g = 0..3 picks a 32 x 64 block of the CTA's part of the
        // accumulator pair -- two-tile form: tile g >> 1, channel half g & 1; WIDE: channels [64 g, 64 g + 64) of the one tile.
        // Private 2 KB staging slice, own TMA store, no barrier between epilogue warps on the tile path.
        const int ew = warp - (2 + F2X_SUM_WARPS);
        const int q = warp & 3;
        const int g = ew >> 2;
        const int j = WIDE ? 0 : (g >> 1);
        const int cbeg = WIDE ? 64 * g : 64 * (g & 1);
        const int r = q * 32 + lane;            // tile row = TMEM lane
        const int et = threadIdx.x - 32 * (2 + F2X_SUM_WARPS);
        constexpr int EPI_THREADS = 32 * F2X_EPI_WARPS;
        const bool side = SLOW && a.out_acc != nullptr;
        const int actm = yq::act_mode(a.ep.act);
        const int pitch = a.W + 1;
        uint8_t *stage = sOut + ew * 2048;      // [32 rows][64 bytes], SWIZZLE_64B
        int cur_nt = -1;
        uint32_t it = 0;
        for (int tile = cid; tile < a.num_tiles; tile += ncl, ++it) {
            const int nt = a.m_pairs == 1 ? tile : (int)__umulhi((uint32_t)tile, a.magic_m), mp = tile - nt * a.m_pairs;
            const int oc0 = nt * BNT;
            const int pb = it & 1;
            if (nt != cur_nt) {   // a new n-tile: everyone is done with the old parameters before they are overwritten
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                for (int i = et; i < BNT; i += EPI_THREADS) {
                    s_q[i] = __ldg(a.ep.chanq + oc0 + i);
                    s_mc[i] = __ldg(a.ep.mcomb + oc0 + i);
                }
                cur_nt = nt;
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            }
            const int p0 = mp * 2 * PPC + (int)rank * PPC + j * 128;
            const int p = p0 + r;
            const int row = (int)__umulhi((uint32_t)p, a.magic_w);
            const int col = p - row * pitch;
            const int n = (int)__umulhi((uint32_t)row, a.magic_h);
            const int y1 = row - n * (a.H + 1);
            const bool valid = p < a.NP && col >= 1 && y1 >= 1;
            const size_t pix = ((size_t)n * a.H + (y1 - 1)) * a.W + (col - 1);
            mbar_wait(&sum_full[pb], (it >> 1) & 1);
            int sa_sum;
            {
                const int *sp = s_sum + pb * F2X_MAX_ROWS + j * 128 + r;
                if (a.size == 3) {
                    const int *s1 = sp + pitch, *s2 = s1 + pitch;
                    sa_sum = (sp[0] + sp[1] + sp[2]) + (s1[0] + s1[1] + s1[2]) + (s2[0] + s2[1] + s2[2]);
                } else {
                    sa_sum = sp[0];
                }
            }
            const int nsa = -sa_sum;
            mbar_wait(&acc_full[pb], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t trow = tmem_base + pb * 256 + j * 128 + ((uint32_t)(q * 32) << 16) + cbeg;
            uint32_t vbuf[2][16];
            if (a.debug != 2) {
                tmem_ld16_issue(trow, vbuf[0]);
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // my staging slice is free again
                tmem_ld_wait16(vbuf[0]);
                __syncwarp();
                auto run = [&](auto actm_tag, auto sat_tag) {
                    constexpr int ACTM = decltype(actm_tag)::value;
                    constexpr bool SAT = decltype(sat_tag)::value;
#pragma unroll 1
                    for (int cp = 0; cp < 2; ++cp) {
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const int ch = 2 * cp + h2;
                        const int c0 = cbeg + 16 * ch;
                        uint32_t(&v)[16] = vbuf[h2];
                        if (ch + 1 < 4) tmem_ld16_issue(trow + 16 * (ch + 1), vbuf[h2 ^ 1]);   // in flight while this chunk is requantized
                        uint32_t packed[4];
                        int extra[16];
                        if (RESID) {
                            // conv -> shortcut in one launch: the residual bytes of this position / chunk travel while the chunk is requantized
                            uint4 rb = make_uint4(0, 0, 0, 0);
                            if (valid) rb = __ldg(reinterpret_cast<const uint4 *>(a.resid + (size_t)p * a.CSO + oc0 + c0));
                            int rv[16];
                            yq::requant_chunk_vals<ACTM, SAT, 16, false>(v, nsa, extra, s_q + c0, s_mc + c0, a.ep.zp_out, rv, a.ep.xlim);
                            yq::shortcut_pack16(rv, rb, a.sc, packed);
                        } else {
                            yq::requant_chunk<ACTM, SAT, 16, false>(v, nsa, extra, s_q + c0, s_mc + c0, a.ep.zp_out, packed, a.ep.xlim);
                        }
                        if (!valid) packed[0] = packed[1] = packed[2] = packed[3] = a.halo_word;
                        yq::mask_pad_channels<16>(packed, a.N - (oc0 + c0));
                        if (SLOW && side && valid) {
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) {
                                const int oc = oc0 + c0 + jj;
                                if (oc < a.N) a.out_acc[pix * a.CSO + oc] = (int)v[jj] + s_q[c0 + jj].y * nsa;
                            }
                        }
                        *reinterpret_cast<uint4 *>(stage + lane * 64 + ((ch ^ ((lane >> 1) & 3)) * 16)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                        if (ch + 1 < 4) tmem_ld_wait16(vbuf[h2 ^ 1]);
                    }
                    }
                };
                if (SLOW && a.ep.saturate) {
                    if (actm == 0) run(std::integral_constant<int, 0>{}, std::true_type{});
                    else if (actm == 1) run(std::integral_constant<int, 1>{}, std::true_type{});
                    else run(std::integral_constant<int, 2>{}, std::true_type{});
                } else {
                    if (actm == 0) run(std::integral_constant<int, 0>{}, std::false_type{});
                    else if (actm == 1) run(std::integral_constant<int, 1>{}, std::false_type{});
                    else run(std::integral_constant<int, 2>{}, std::false_type{});
                }
            }
            // this warp's TMEM and S reads of the pair are done: hand the accumulators back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                f2x_arrive(&acc_empty[pb]);                                              // local: my sum warps / (leader) MMA warp
                if (rank == 1) mbar_arrive_cluster(map_to_cta(&acc_peer_empty[pb], 0));      // the leader's MMA warp
            }
            if (a.debug != 2) {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmO, stage, oc0 + cbeg, p0 + q * 32);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // global writes complete before the CTA retires
    }
    tc_fence_before();
    cluster_sync_all();            // nobody leaves while the peer may still signal its barriers or use the pair's TMEM
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2cta<512>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn f2x_get_encode()
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

int f2x_encode_2d(CUtensorMap *m, const void *ptr, uint64_t rows, int row_bytes, int box_c, int box_rows, CUtensorMapL2promotion prom)
{
    EncodeTiledFn enc = f2x_get_encode();
    if (!enc) return yq::fail("cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     box_c >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return yq::fail("cuTensorMapEncodeTiled(%llu x %d, box %d x %d) failed: %d", (unsigned long long)rows, row_bytes, box_c, box_rows, (int)r);
    return 0;
}

struct Flat2xState {
    int KC, n_pad;
    uint8_t *w = nullptr;       // [n_pad][size*size*cs_in]
    CUtensorMap tmB, tmBw;      // weight boxes of 64 rows (two-tile form) / 128 rows (WIDE form)
    bool wide = false;
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

template <int KC, bool SLOW, bool WIDE, bool RESID>
int f2x_launch_v(Flat2xState *st, const CUtensorMap &tmA, const CUtensorMap &tmO, const CUtensorMap &tmA2, Flat2xArgs a, cudaStream_t stream)
{
    using L = Flat2xSmem<KC, WIDE>;
    const int n_sm = yq::device_sm_count(), smem_max = yq::device_smem_optin();   // (of the CURRENT device: nothing cached per process)
    if (n_sm <= 0 || smem_max <= 0) return yq::fail("cannot query the device's multiprocessor count / shared memory size");
    // 1x1 layers: as many patch stages as fit beside the staging tiles and four weight stages (at most F2X_ASTAGES)
    a.a_stages = 2;
    if (a.size == 1)
        while (a.a_stages < F2X_ASTAGES &&
               (a.a_stages + 1) * a.a_stage_bytes + 2 * L::OUT_BYTES + L::PARAM_BYTES + L::SUM_BYTES + 512 + 1024 + 4 * L::B_STAGE <= smem_max)
            ++a.a_stages;
    const int fixed = a.a_stages * a.a_stage_bytes + 2 * L::OUT_BYTES + L::PARAM_BYTES + L::SUM_BYTES + 512 + 1024;
    int nbs = (smem_max - fixed) / L::B_STAGE;
    if (nbs > F2X_MAX_BSTAGES) nbs = F2X_MAX_BSTAGES;
    if (nbs < 2) return yq::fail("conv_u8_tc_flat2x_kernel<%d>: shared memory does not hold two weight stages", KC);
    a.b_stages = nbs;
    int smem = fixed + nbs * L::B_STAGE;
    // Never two of these CTAs on one SM: each takes all 512 TMEM columns, and two CTA pairs of concurrent launches (the
    // side stream of yq_network.cu) that each hold one SM's columns while waiting for the other's would not finish.
    if (smem <= smem_max / 2) smem = smem_max / 2 + 1024;
    auto kern = conv_u8_tc_flat2x_kernel<KC, SLOW, WIDE, RESID>;
    if (yq::ensure_dynamic_smem((const void *)kern, smem)) return -1;
    int grid = 2 * a.num_tiles < n_sm ? 2 * a.num_tiles : n_sm / 2 * 2;      // whole CTA pairs
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(F2X_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = yq::pdl_enabled() ? 2 : 1;
    YQ_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, WIDE ? st->tmBw : st->tmB, tmO, tmA2, a));
    return 0;
}

template <int KC>
int f2x_launch(Flat2xState *st, const CUtensorMap &tmA, const CUtensorMap &tmO, const CUtensorMap &tmA2, const Flat2xArgs &a, cudaStream_t stream)
{
    if (st->wide) {
        if (a.resid) return f2x_launch_v<KC, false, true, true>(st, tmA, tmO, tmA2, a, stream);
        if (a.out_acc || a.ep.saturate) return f2x_launch_v<KC, true, true, false>(st, tmA, tmO, tmA2, a, stream);
        return f2x_launch_v<KC, false, true, false>(st, tmA, tmO, tmA2, a, stream);
    }
    if (a.resid) return f2x_launch_v<KC, false, false, true>(st, tmA, tmO, tmA2, a, stream);
    if (a.out_acc || a.ep.saturate) return f2x_launch_v<KC, true, false, false>(st, tmA, tmO, tmA2, a, stream);
    return f2x_launch_v<KC, false, false, false>(st, tmA, tmO, tmA2, a, stream);
}

}  // namespace

#ifndef F2X_1X1_DEFAULT
#define F2X_1X1_DEFAULT 0
#endif
int yq_tc_flat2x_supported(const yq_conv_layer *l)
{
    if (!yq_tc_flat_supported(l)) return 0;
    if (l->quant_stop_flag || l->cs_out % F2X_BN) return 0;
    // 1x1 (one weight stage per patch chunk): YQ_FLAT2X_1X1 = 1 / 0 switches it on / off for A/B measurements
    static const int one_env = getenv("YQ_FLAT2X_1X1") ? atoi(getenv("YQ_FLAT2X_1X1")) : -1;
    if (l->size != 3 && !(l->size == 1 && (one_env < 0 ? F2X_1X1_DEFAULT : one_env))) return 0;
    if (256 + (l->size - 1) * (l->w + 2) > F2X_MAX_ROWS) return 0;
    return 1;
}

int yq_tc_flat2x_prepare(yq_conv_layer *l, void **state)
{
    Flat2xState *st = new Flat2xState();
    st->KC = (l->cs_in % 128) ? 64 : 128;
    {
        const char *e = getenv("YQ_FLAT2X_WIDE");      // 0 keeps the two-tile N = 128 form everywhere (A/B measurements)
        st->wide = l->cs_out % 256 == 0 && !(e && atoi(e) == 0);
    }
    st->n_pad = yq::round_up(l->n, st->wide ? 256 : F2X_BN);
    const int taps = l->size * l->size;
    const size_t ktot = (size_t)taps * l->cs_in;
    // [n_pad][taps][cs_in] rows (shared by the flat kernels: same tag, same image)
    std::vector<uint8_t> wp;
    char tag[24];
    snprintf(tag, sizeof tag, "ohwi.%d", st->n_pad);
    // (a data-parallel replica takes the image from the arena blob broadcast to its device: no host packing, no upload)
    const bool on_dev = yq::pack_fetch_device(l, tag, (size_t)st->n_pad * ktot, (void **)&st->w);
    if (!on_dev && (!yq::pack_fetch(l, tag, wp) || wp.size() != (size_t)st->n_pad * ktot)) {
        wp.assign((size_t)st->n_pad * ktot, 0);
        for (int oc = 0; oc < l->n; ++oc)
            for (int t = 0; t < taps; ++t)
                for (int ci = 0; ci < l->c; ++ci) wp[(size_t)oc * ktot + (size_t)t * l->cs_in + ci] = l->host_w[((size_t)oc * l->c + ci) * taps + t];
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
    if (f2x_encode_2d(&st->tmB, st->w, (uint64_t)st->n_pad, (int)ktot, st->KC, 64, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return cleanup();
    if (f2x_encode_2d(&st->tmBw, st->w, (uint64_t)st->n_pad, (int)ktot, st->KC, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return cleanup();
    *state = st;
    return 0;
}

int yq_tc_flat2x_chunk(const void *state) { return state ? ((const Flat2xState *)state)->KC : 0; }

void yq_tc_flat2x_free(void *state)
{
    Flat2xState *st = (Flat2xState *)state;
    if (!st) return;
    cudaFree(st->w);
    delete st;
}

int yq_tc_flat2x_forward(yq_conv_layer *l, void *state, const uint8_t *in_flat, uint8_t *out_flat, int halo_fill, int32_t *out_acc, int batch,
                        cudaStream_t stream, const yq_fused_shortcut *sc, const uint8_t *in_first, int c_first)
{
    Flat2xState *st = (Flat2xState *)state;
    if (!st || !in_flat || !out_flat) return yq::fail("tcgen05 flat2 flavour: bad argument");
    // in_first != null: the input is the concatenation [in_first (c_first channels) | in_flat (the rest)] of two flat tensors
    if (in_first && (c_first <= 0 || c_first >= l->cs_in || c_first % st->KC || (l->cs_in - c_first) % 16))
        return yq::fail("tcgen05 flat2 flavour: a two-tensor input splits at a multiple of %d channels", st->KC);
    const int c_second = in_first ? l->cs_in - c_first : l->cs_in;
    const int W1 = l->w + 1, H1 = l->h + 1;
    const long long NP = (long long)batch * H1 * W1;
    const long long rows_alloc = NP + W1 + 2;     // + the trailing halo row (yq_act_geom_bytes)
    if (rows_alloc * W1 >= 0x100000000ll) return yq::fail("tcgen05 flat2 flavour: tensor too large for 32-bit position arithmetic");
    Flat2xArgs a;
    memset(&a, 0, sizeof a);
    const int pad = l->size / 2;
    const int ppc = st->wide ? 128 : 256, bnt = st->wide ? 256 : F2X_BN;      // positions per CTA, channels per tile
    a.patch_rows = ppc + (l->size - 1) * (W1 + 1);
    a.box_rows = yq::round_up((a.patch_rows + 1) / 2, 8);
    a.a_stage_bytes = yq::round_up(2 * a.box_rows * st->KC, 1024);
    Flat2xState::Key key{in_flat, in_first, out_flat, batch};
    auto it = st->maps.find(key);
    if (it == st->maps.end()) {
        if (st->maps.size() > 64) st->maps.clear();
        Flat2xState::Maps m;
        if (f2x_encode_2d(&m.a, in_flat, (uint64_t)rows_alloc, c_second, st->KC, a.box_rows, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return -1;
        // the store box is one epilogue warp's block: 32 positions x 64 channels
        if (f2x_encode_2d(&m.o, out_flat, (uint64_t)rows_alloc, l->cs_out, 64, 32, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return -1;
        m.a2 = m.a;
        if (in_first && f2x_encode_2d(&m.a2, in_first, (uint64_t)rows_alloc, c_first, st->KC, a.box_rows, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return -1;
        it = st->maps.emplace(key, m).first;
    }
    a.split = in_first ? c_first / st->KC : 0;
    a.ep = yq::make_epi(l);
    a.out_acc = out_acc;
    a.N = l->n; a.CSO = l->cs_out;
    if (sc && sc->resid) {
        a.resid = sc->resid;
        a.sc = yq::ShortcutParams{sc->Ka, sc->Kb, sc->C0};
    }
    a.B = batch; a.H = l->h; a.W = l->w; a.NP = (int)NP;
    a.size = l->size; a.taps = l->size * l->size; a.cpt = l->cs_in / st->KC; a.CS = l->cs_in;
    a.q_off = -(pad * W1 + pad);
    a.halo_word = 0x01010101u * (uint32_t)(halo_fill & 0xff);
    {
        static int dbg = -1;
        if (dbg < 0) dbg = getenv("YQ_FLAT2X_DEBUG") ? atoi(getenv("YQ_FLAT2X_DEBUG")) : 0;
        a.debug = dbg;
    }
    a.magic_w = (uint32_t)((0x100000000ull + W1 - 1) / W1);
    a.magic_h = (uint32_t)((0x100000000ull + H1 - 1) / H1);
    a.m_pairs = (int)((rows_alloc + 2 * ppc - 1) / (2 * ppc));      // cluster tiles of 2 * ppc positions
    a.num_tiles = a.m_pairs * (st->n_pad / bnt);
    a.magic_m = a.m_pairs == 1 ? 0u : (uint32_t)((0x100000000ull + a.m_pairs - 1) / a.m_pairs);
    if ((long long)a.num_tiles * a.m_pairs >= 0x100000000ll) return yq::fail("tcgen05 flat2 flavour: too many tiles for 32-bit tile arithmetic");
    const CUtensorMap &tmA = it->second.a, &tmO = it->second.o, &tmA2 = it->second.a2;
    if (st->KC == 128) return f2x_launch<128>(st, tmA, tmO, tmA2, a, stream);
    return f2x_launch<64>(st, tmA, tmO, tmA2, a, stream);
}
