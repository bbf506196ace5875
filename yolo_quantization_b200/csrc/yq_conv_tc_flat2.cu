// yq_conv_tc_flat2.cu -- the flat-strip tcgen05 (kind::i8) convolution, second form: PERSISTENT, two M-tiles (256
// positions) per weight stage, four TMEM accumulators (two tile pairs in flight), no all-ones filter rows.
//
// Why (measured on B200, see yq_conv_tc_flat.cu and tools/probes/):
//   * a weight stage (128 channels x KC bytes) is the dominant operand; one tile per stage needs 57 B/clk of TMA ingest per
//     SM and one tcgen05.commit per 4 MMAs (commits cannot be closer than ~358 clocks per SM).  Feeding TWO tiles from every
//     stage halves both: 8 MMAs (N = 128, 64 clocks each) per stage and per commit;
//   * that needs 2 x 2 accumulators to overlap the epilogue of pair i with the main loop of pair i+1: 4 x 128 TMEM columns
//     = all 512, so the 16 all-ones filter rows (sum of activations per position, for the weight zero point) have no room.
//     Instead two "sum warps" read every activation patch once from shared memory while the tensor core works on it,
//     S[row] = sum over the row's channels (dp4a), accumulated over the channel chunks; the epilogue then takes
//         sum_a(position p) = sum over taps S[p + (ky*(W+1) + kx)]          (9 shared-memory loads per position)
//     which is the same integer the ones rows produced (convolutional_layer.c:718-721: acc = sum w*a - zp_w * sum a).
//
// Same tensors, geometry, arithmetic and epilogue as yq_conv_tc_flat.cu (flat halo-padded strips, row-shifted swizzled
// descriptors per tap, integer-form requantize, halo positions rewritten with the consumer's zero point).
// Warp roles (20 warps): 0 = TMA producer, 1 = MMA issuer, 2-3 = activation sums, 4-19 = epilogue (4 per TMEM lane quarter).
// Restates convolutional_layer.c:694-751 for stride 1, pad = size/2, size in {1, 3}, c % 64 == 0 or c = 32 (KC = 32: 32-byte patch
// rows, SWIZZLE_32B, one K = 32 MMA per tile and tap), n % 128 == 0 or n in {32, 64}.
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

constexpr int F2_BN = 128;
constexpr int F2_EPI_WARPS = 16;
constexpr int F2_SUM_WARPS = 2;
constexpr int F2_THREADS = 32 * (2 + F2_SUM_WARPS + F2_EPI_WARPS);
constexpr int F2_ASTAGES = 4;              // barrier slots; a launch uses a.a_stages of them: 2 for 3x3 layers (nine taps of MMAs per
                                           // patch chunk cover the next chunk's load), up to 4 for 1x1 layers (one tap per chunk: the
                                           // loads have to run several chunks ahead of the tensor pipe)
constexpr int F2_MAX_BSTAGES = 8;
constexpr int F2_MAX_ROWS = 704;           // patch rows: 256 + 2W + 4 (W <= 222: the 208 x 208 and 104 x 104 maps of the full yolov3)
constexpr int F2_MAX_BOXES = 3;            // a patch arrives as up to three TMA boxes (a box holds at most 256 rows)

struct Flat2Args {
    yq::EpiParams ep;
    int32_t *out_acc;      // dense NHWC [B][H][W][CSO] (parity checks) or nullptr
    int N, CSO;
    int B, H, W, NP;       // NP = B*(H+1)*(W+1)
    int size, taps, cpt /* KC-chunks per tap */, CS;
    int tps;               // taps per weight stage (Flat2Smem)
    int q_off;             // first patch position relative to the pair's first position: -(pad*(W+1) + pad)
    int patch_rows, box_rows, n_boxes, a_stage_bytes, a_stages, b_stages;
    int plain;             // 1x1 only: the tensors are plain [B][H][W][C] strips (no halo positions: every p < NP is a pixel)
    int out_cols;          // bytes per position the launch stores: min(128, channel stride of the output); 32 / 64: narrow layers
    int warp_cols;         // channels of a tile one epilogue warp requantises and stores: 64, or out_cols / 2 in narrow layers (all 16 warps work)
    int m_pairs, num_tiles;   // tile t -> (n-tile t / m_pairs, position pair t % m_pairs)
    uint32_t halo_word;
    uint32_t magic_w, magic_h, magic_m;
    // the FOLLOWING quantized shortcut fused into the epilogue (extension layer, include/yq_b200.h): the `from` tensor in this
    // layer's own flat geometry and channel stride; the launch then stores the SHORTCUT's output
    const uint8_t *resid;
    yq::ShortcutParams sc;
    int debug;             // YQ_FLAT2_DEBUG experiments (results are garbage): 1 = skip weight loads of taps > 0, 2 = skip the epilogue math
};

template <int KC>
struct Flat2Smem {
    // a weight stage holds `tps` taps: 1, or -- KC = 32, 3x3 -- one filter row of three (a tap is then a single K = 32 MMA per tile: with
    // one tap per stage the tensor pipe would wait on the commits, which cannot follow each other faster than every ~358 clocks)
    static constexpr int TAP_BYTES = F2_BN * KC;
    static constexpr int B_STAGE = TAP_BYTES * (KC == 32 ? 3 : 1);
    static constexpr int OUT_BYTES = 128 * F2_BN;                 // one staging tile; there are two
    static constexpr int PARAM_BYTES = F2_BN * 24;
    static constexpr int SUM_BYTES = 2 * F2_MAX_ROWS * 4;         // S[pair buffer][patch row]
};

__device__ __forceinline__ void f2_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }

// RESID: the following quantized shortcut is fused into the epilogue (a launch-uniform choice: a template parameter keeps the
// other form's code out of the instruction cache -- the 16 epilogue warps run apart from each other and `no instruction` stalls
// were 15 % of the samples with both forms and four unrolled chunks in one loop body, r2 profile)
template <int KC, bool SLOW, bool RESID>
__global__ void __launch_bounds__(F2_THREADS, 1) conv_u8_tc_flat2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                     const __grid_constant__ CUtensorMap tmO, const Flat2Args a)
{
    using L = Flat2Smem<KC>;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sA = smem;                                              // F2_ASTAGES patch stages
    uint8_t *sB = sA + a.a_stages * a.a_stage_bytes;                 // a.b_stages weight stages
    uint8_t *sOut = sB + a.b_stages * L::B_STAGE;                    // two output staging tiles
    int4 *s_q = (int4 *)(sOut + 2 * L::OUT_BYTES);                   // {bias, zw, 2*M0, shift} of the current n-tile
    double *s_mc = (double *)(s_q + F2_BN);
    int *s_sum = (int *)(s_mc + F2_BN);                              // [2][F2_MAX_ROWS]
    uint64_t *a_full = (uint64_t *)(s_sum + 2 * F2_MAX_ROWS);
    uint64_t *a_empty = a_full + F2_ASTAGES;
    uint64_t *b_full = a_empty + F2_ASTAGES;
    uint64_t *b_empty = b_full + F2_MAX_BSTAGES;
    uint64_t *acc_full = b_empty + F2_MAX_BSTAGES;
    uint64_t *acc_empty = acc_full + 2;
    uint64_t *sum_full = acc_empty + 2;
    uint32_t *tmem_slot = (uint32_t *)(sum_full + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = a.cpt, nbs = a.b_stages;
    const int tps = KC == 32 ? a.tps : 1;      // (a compile-time 1 for the wide layers: their tap loops stay as tight as they were)

    if (threadIdx.x == 0) {
        for (int s = 0; s < F2_ASTAGES; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1 + F2_SUM_WARPS);     // the MMA commit and each sum warp release a patch stage
        }
        for (int s = 0; s < F2_MAX_BSTAGES; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], F2_EPI_WARPS);
            mbar_init(&sum_full[s], F2_SUM_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    yq_pdl_wait_then_release();                             // no activation tensor was touched so far

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform control flow, one elected lane issues) =====================
        int sa = 0, s = 0;
        uint32_t pha = 0, phb = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            const int nt = a.m_pairs == 1 ? tile : (int)__umulhi((uint32_t)tile, a.magic_m), mp = tile - nt * a.m_pairs;
            const int p0 = mp * 256, oc0 = nt * F2_BN;
            for (int c = 0; c < chunks; ++c) {
                mbar_wait(&a_empty[sa], pha ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&a_full[sa], (uint32_t)(a.n_boxes * a.box_rows * KC));
                    for (int b = 0; b < a.n_boxes; ++b)
                        tma_load_2d(sA + sa * a.a_stage_bytes + b * a.box_rows * KC, &tmA, &a_full[sa], c * KC, p0 + a.q_off + b * a.box_rows);
                }
                if (++sa == a.a_stages) { sa = 0; pha ^= 1; }
                for (int tap = 0; tap < a.taps; tap += tps) {
                    mbar_wait(&b_empty[s], phb ^ 1);
                    if (elect_one()) {
                        if (a.debug == 1 && tap > 0) {
                            f2_arrive(&b_full[s]);
                        } else {
                            mbar_expect_tx(&b_full[s], (uint32_t)(tps * L::TAP_BYTES));
                            for (int t = 0; t < tps; ++t)
                                tma_load_2d(sB + s * L::B_STAGE + t * L::TAP_BYTES, &tmB, &b_full[s], (tap + t) * a.CS + c * KC, oc0);
                        }
                    }
                    if (++s == nbs) { s = 0; phb ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc(F2_BN);
        const uint32_t row_step = (uint32_t)((a.W + 1 - a.size) * KC);
        int sa = 0, s = 0;
        uint32_t pha = 0, phb = 0, it = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
            const int pb = it & 1;
            mbar_wait(&acc_empty[pb], ((it >> 1) & 1) ^ 1);      // the epilogue has drained this pair of accumulators
            tc_fence_after();
            const uint32_t acc0 = tmem_base + pb * 2 * F2_BN, acc1 = acc0 + F2_BN;
            uint32_t accumulate = 0;
            for (int c = 0; c < chunks; ++c) {
                mbar_wait(&a_full[sa], pha);
                uint32_t tap_addr = smem_u32(sA + sa * a.a_stage_bytes);    // row-shifted descriptor start of the current tap
                int kx = 0;
                for (int tap = 0; tap < a.taps; tap += tps) {
                    mbar_wait(&b_full[s], phb);
                    tc_fence_after();
                    for (int t = 0; t < tps; ++t) {
                        if (elect_one()) {
                            const uint64_t da0 = make_desc<KC>(tap_addr), da1 = make_desc<KC>(tap_addr + 128 * KC);
                            const uint64_t db = make_desc<KC>(smem_u32(sB + s * L::B_STAGE + t * L::TAP_BYTES));
#pragma unroll
                            for (int k = 0; k < KC / 32; ++k) {
                                umma_i8(acc0, da0 + 2 * k, db + 2 * k, idesc, (k == 0) ? accumulate : 1u);
                                umma_i8(acc1, da1 + 2 * k, db + 2 * k, idesc, (k == 0) ? accumulate : 1u);
                            }
                            if (t + 1 == tps) umma_commit(&b_empty[s]);
                        }
                        accumulate = 1;
                        tap_addr += KC;
                        if (++kx == a.size) { kx = 0; tap_addr += row_step; }
                    }
                    if (++s == nbs) { s = 0; phb ^= 1; }
                }
                if (elect_one()) umma_commit(&a_empty[sa]);
                if (++sa == a.a_stages) { sa = 0; pha ^= 1; }
            }
            if (elect_one()) umma_commit(&acc_full[pb]);
        }
    } else if (warp < 2 + F2_SUM_WARPS) {
        // ===================== activation sums: S[row] = sum of the patch row's bytes over all channel chunks =====================
        const int st = threadIdx.x - 64;                              // 0 .. 63
        constexpr int RPT = (F2_MAX_ROWS + 32 * F2_SUM_WARPS - 1) / (32 * F2_SUM_WARPS);   // rows per thread
        int sa = 0;
        uint32_t pha = 0, it = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
            const int pb = it & 1;
            int acc[RPT];
#pragma unroll
            for (int k = 0; k < RPT; ++k) acc[k] = 0;
            for (int c = 0; c < chunks; ++c) {
                mbar_wait(&a_full[sa], pha);
                const uint8_t *patch = sA + sa * a.a_stage_bytes;
#pragma unroll
                for (int k = 0; k < RPT; ++k) {
                    const int row = st + k * 32 * F2_SUM_WARPS;
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
                if (lane == 0) f2_arrive(&a_empty[sa]);
                if (++sa == a.a_stages) { sa = 0; pha ^= 1; }
            }
            mbar_wait(&acc_empty[pb], ((it >> 1) & 1) ^ 1);           // S[pb] was consumed by the epilogue of pair it-2
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const int row = st + k * 32 * F2_SUM_WARPS;
                if (row < a.patch_rows) s_sum[pb * F2_MAX_ROWS + row] = acc[k];
            }
            __syncwarp();
            if (lane == 0) f2_arrive(&sum_full[pb]);
        }
    } else {
        // ===================== epilogue: 16 warps, each on its own =====================
        // Warp (q, g): q = warp & 3 is the TMEM lane quarter the hardware lets it read (32 positions), g = 0..3 picks the tile of the
        // pair (g >> 1) and the 64-channel half of it (g & 1).  A warp requantizes its 32 x 64 block into a PRIVATE 2 KB staging
        // slice and issues its own TMA store: no barrier between epilogue warps on the tile path, so the warps drift apart and one
        // warp's TMEM-load and store latencies hide behind the others' arithmetic (the CTA-wide barrier per tile of the first
        // version kept all 16 in lockstep: issue slots 55 % used, profiles/r2_flat2_before.txt).
        const int ew = warp - (2 + F2_SUM_WARPS);
        const int q = warp & 3;
        const int g = ew >> 2;
        const int j = g >> 1, half = g & 1;
        const int r = q * 32 + lane;            // tile row = TMEM lane
        const int et = threadIdx.x - 32 * (2 + F2_SUM_WARPS);
        constexpr int EPI_THREADS = 32 * F2_EPI_WARPS;
        const bool side = SLOW && a.out_acc != nullptr;
        const int actm = yq::act_mode(a.ep.act);
        const int pitch = a.W + 1;
        const int cbeg = half * a.warp_cols;
        const int nch = a.warp_cols / 16;                                // 16-channel chunks of mine
        const int rowb = a.warp_cols;                                    // bytes per staging row = inner box of the store: 64, 32 or 16
        uint8_t *stage = sOut + ew * 2048;                               // [32 rows][rowb] in the store map's swizzle
        const int lr = lane;                                             // row within my staging slice
        int cur_nt = -1;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
            const int nt = a.m_pairs == 1 ? tile : (int)__umulhi((uint32_t)tile, a.magic_m), mp = tile - nt * a.m_pairs;
            const int oc0 = nt * F2_BN;
            const int pb = it & 1;
            if (nt != cur_nt) {
                // a new n-tile: every epilogue warp is done with the old parameters before they are overwritten (the only barrier
                // among the epilogue warps; once per n-tile, i.e. once or twice per CTA on the big layers)
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                for (int i = et; i < F2_BN; i += EPI_THREADS) {
                    s_q[i] = __ldg(a.ep.chanq + oc0 + i);
                    s_mc[i] = __ldg(a.ep.mcomb + oc0 + i);
                }
                cur_nt = nt;
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            }
            const int p0 = mp * 256 + j * 128;
            const int p = p0 + r;
            const int row = (int)__umulhi((uint32_t)p, a.magic_w);
            const int col = p - row * pitch;
            const int n = (int)__umulhi((uint32_t)row, a.magic_h);
            const int y1 = row - n * (a.H + 1);
            const bool valid = p < a.NP && (a.plain || (col >= 1 && y1 >= 1));
            const size_t pix = a.plain ? (size_t)p : ((size_t)n * a.H + (y1 - 1)) * a.W + (col - 1);
            mbar_wait(&sum_full[pb], (it >> 1) & 1);
            // sum of activations under this position's window, from the per-row sums of the patch
            int sa_sum;
            {
                const int *sp = s_sum + pb * F2_MAX_ROWS + j * 128 + r;
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
            const uint32_t trow = tmem_base + (pb * 2 + j) * F2_BN + ((uint32_t)(q * 32) << 16) + cbeg;
            uint32_t vbuf[2][16];
            if (nch > 0 && a.debug != 2) {
                tmem_ld16_issue(trow, vbuf[0]);
                // my staging slice is free once the store of the previous pair has read it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
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
                        if (ch < nch) {
                            const int c0 = cbeg + 16 * ch;
                            uint32_t(&v)[16] = vbuf[h2];
                            if (ch + 1 < nch) tmem_ld16_issue(trow + 16 * (ch + 1), vbuf[h2 ^ 1]);   // in flight while this chunk is requantized
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
                            // 64-byte rows: SWIZZLE_64B (16-byte chunk ^ row bits 1-2); 32-byte rows: SWIZZLE_32B (chunk ^ row bit 2); 16-byte rows: plain
                            const int sw = rowb == 64 ? (ch ^ ((lr >> 1) & 3)) : (rowb == 32 ? (ch ^ ((lr >> 2) & 1)) : 0);     // (16-byte rows: no swizzle)
                            *reinterpret_cast<uint4 *>(stage + lr * rowb + sw * 16) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                            if (ch + 1 < nch) tmem_ld_wait16(vbuf[h2 ^ 1]);
                        }
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
            if (lane == 0) f2_arrive(&acc_empty[pb]);
            if (nch > 0 && a.debug != 2) {
                fence_proxy_async();          // my staging writes -> visible to the TMA unit
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
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn f2_get_encode()
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

int f2_encode_2d(CUtensorMap *m, const void *ptr, uint64_t rows, int row_bytes, int box_c, int box_rows, CUtensorMapL2promotion prom)
{
    EncodeTiledFn enc = f2_get_encode();
    if (!enc) return yq::fail("cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     box_c >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : (box_c == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : (box_c == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE)),
                     prom,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return yq::fail("cuTensorMapEncodeTiled(%llu x %d, box %d x %d) failed: %d", (unsigned long long)rows, row_bytes, box_c, box_rows, (int)r);
    return 0;
}

struct Flat2State {
    int KC, n_pad;
    uint8_t *w = nullptr;       // [n_pad][size*size*cs_in]
    CUtensorMap tmB;
    struct Key {
        const void *in;
        void *out;
        int batch;
        bool operator<(const Key &o) const { return in != o.in ? in < o.in : out != o.out ? out < o.out : batch < o.batch; }
    };
    std::map<Key, std::pair<CUtensorMap, CUtensorMap>> maps;
};

template <int KC, bool SLOW, bool RESID>
int f2_launch_v(Flat2State *st, const CUtensorMap &tmA, const CUtensorMap &tmO, Flat2Args a, cudaStream_t stream)
{
    using L = Flat2Smem<KC>;
    const int n_sm = yq::device_sm_count(), smem_max = yq::device_smem_optin();   // (of the CURRENT device: nothing cached per process)
    if (n_sm <= 0 || smem_max <= 0) return yq::fail("cannot query the device's multiprocessor count / shared memory size");
    // 1x1 layers: as many patch stages as fit beside the staging tiles and three weight stages (at most F2_ASTAGES)
    a.a_stages = 2;
    if (a.size == 1)
        while (a.a_stages < F2_ASTAGES &&
               (a.a_stages + 1) * a.a_stage_bytes + 2 * L::OUT_BYTES + L::PARAM_BYTES + L::SUM_BYTES + 512 + 1024 + 3 * L::B_STAGE <= smem_max)
            ++a.a_stages;
    const int fixed = a.a_stages * a.a_stage_bytes + 2 * L::OUT_BYTES + L::PARAM_BYTES + L::SUM_BYTES + 512 + 1024;
    int nbs = (smem_max - fixed) / L::B_STAGE;
    if (nbs > F2_MAX_BSTAGES) nbs = F2_MAX_BSTAGES;
    if (nbs < 2) return yq::fail("conv_u8_tc_flat2_kernel<%d>: shared memory does not hold two weight stages", KC);
    a.b_stages = nbs;
    const int smem = fixed + nbs * L::B_STAGE;
    auto kern = conv_u8_tc_flat2_kernel<KC, SLOW, RESID>;
    if (yq::ensure_dynamic_smem((const void *)kern, smem)) return -1;
    const int grid = a.num_tiles < n_sm ? a.num_tiles : n_sm;
    YQ_CUDA(yq::launch_pdl(kern, dim3(grid), dim3(F2_THREADS), smem, stream, tmA, st->tmB, tmO, a));
    return 0;
}

template <int KC>
int f2_launch(Flat2State *st, const CUtensorMap &tmA, const CUtensorMap &tmO, const Flat2Args &a, cudaStream_t stream)
{
    if (a.resid) return f2_launch_v<KC, false, true>(st, tmA, tmO, a, stream);      // (the fused shortcut has no side outputs: production plan only)
    if (a.out_acc || a.ep.saturate) return f2_launch_v<KC, true, false>(st, tmA, tmO, a, stream);
    return f2_launch_v<KC, false, false>(st, tmA, tmO, a, stream);
}

}  // namespace

int yq_tc_flat2_supported(const yq_conv_layer *l)
{
    if (!yq_tc_flat_eligible(l)) return 0;
    if (l->quant_stop_flag) return 0;
    if (l->cs_out % F2_BN && l->cs_out != 64 && l->cs_out != 32) return 0;   // narrow layers store 64 / 32 bytes per position of one 128-channel tile
    // (1x1: with one wave of tiles or less the one-tile form is faster -- layer 13 of yolov3-tiny 0.0215 vs 0.0245 ms; the
    // dispatch in yq_forward_convolutional_layer_quant_flat_gpu picks per launch from the tile count)
    if (256 + (l->size - 1) * (l->w + 2) > F2_MAX_ROWS) return 0;
    return 1;
}

int yq_tc_flat2_prepare(yq_conv_layer *l, void **state)
{
    Flat2State *st = new Flat2State();
    st->KC = (l->cs_in % 128) ? ((l->cs_in % 64) ? 32 : 64) : 128;
    st->n_pad = yq::round_up(l->n, F2_BN);
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
    if (f2_encode_2d(&st->tmB, st->w, (uint64_t)st->n_pad, (int)ktot, st->KC, F2_BN, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return cleanup();
    *state = st;
    return 0;
}

void yq_tc_flat2_free(void *state)
{
    Flat2State *st = (Flat2State *)state;
    if (!st) return;
    cudaFree(st->w);
    delete st;
}

int yq_tc_flat2_forward(yq_conv_layer *l, void *state, const uint8_t *in_flat, uint8_t *out_flat, int halo_fill, int32_t *out_acc, int batch,
                        cudaStream_t stream, int plain, const yq_fused_shortcut *sc)
{
    Flat2State *st = (Flat2State *)state;
    if (!st || !in_flat || !out_flat) return yq::fail("tcgen05 flat2 flavour: bad argument");
    if (plain && l->size != 1) return yq::fail("tcgen05 flat2 flavour: only 1x1 layers run on plain tensors");
    // plain: a 1x1 convolution needs no halo, so a plain [B][H][W][C] tensor is a strip of B*H*W positions, all of them pixels
    const int W1 = l->w + (plain ? 0 : 1), H1 = l->h + (plain ? 0 : 1);
    const long long NP = (long long)batch * H1 * W1;
    const long long rows_alloc = plain ? NP : NP + W1 + 2;     // + the trailing halo row (yq_act_geom_bytes)
    if (rows_alloc * W1 >= 0x100000000ll) return yq::fail("tcgen05 flat2 flavour: tensor too large for 32-bit position arithmetic");
    Flat2Args a;
    memset(&a, 0, sizeof a);
    const int pad = l->size / 2;
    a.patch_rows = 256 + (l->size - 1) * (W1 + 1);
    a.n_boxes = a.patch_rows <= 256 ? 1 : (a.patch_rows <= 512 ? 2 : F2_MAX_BOXES);
    a.box_rows = yq::round_up((a.patch_rows + a.n_boxes - 1) / a.n_boxes, 8);
    a.a_stage_bytes = yq::round_up(a.n_boxes * a.box_rows * st->KC, 1024);
    a.out_cols = l->cs_out < F2_BN ? l->cs_out : F2_BN;
    // narrow layers (n = 64 / 32): the two warps of a lane quarter and tile split the channels that exist, instead of one of them idling
    static const bool split_narrow = !(getenv("YQ_FLAT2_NARROW_SPLIT") && !atoi(getenv("YQ_FLAT2_NARROW_SPLIT")));     // A/B measurements
    a.warp_cols = a.out_cols >= F2_BN ? 64 : (split_narrow ? a.out_cols / 2 : (a.out_cols >= 64 ? 64 : 32));
    a.plain = plain ? 1 : 0;
    if (sc && sc->resid) {
        if (plain) return yq::fail("tcgen05 flat2 flavour: the fused shortcut needs flat tensors");
        a.resid = sc->resid;
        a.sc = yq::ShortcutParams{sc->Ka, sc->Kb, sc->C0};
    }
    Flat2State::Key key{in_flat, out_flat, batch * 2 + (plain ? 1 : 0)};
    auto it = st->maps.find(key);
    if (it == st->maps.end()) {
        if (st->maps.size() > 64) st->maps.clear();
        CUtensorMap tmA, tmO;
        if (f2_encode_2d(&tmA, in_flat, (uint64_t)rows_alloc, l->cs_in, st->KC, a.box_rows, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return -1;
        // the store box is one epilogue warp's block: 32 positions x 64 channels (narrow layers: half of their channels, 32 or 16)
        if (f2_encode_2d(&tmO, out_flat, (uint64_t)rows_alloc, l->cs_out, a.warp_cols, 32, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return -1;
        it = st->maps.emplace(key, std::make_pair(tmA, tmO)).first;
    }
    a.ep = yq::make_epi(l);
    a.out_acc = out_acc;
    a.N = l->n; a.CSO = l->cs_out;
    a.B = batch; a.H = l->h; a.W = l->w; a.NP = (int)NP;
    a.size = l->size; a.taps = l->size * l->size; a.cpt = l->cs_in / st->KC; a.CS = l->cs_in;
    a.tps = (st->KC == 32 && l->size == 3) ? 3 : 1;
    a.q_off = -(pad * W1 + pad);
    a.halo_word = 0x01010101u * (uint32_t)(halo_fill & 0xff);
    {
        static int dbg = -1;
        if (dbg < 0) dbg = getenv("YQ_FLAT2_DEBUG") ? atoi(getenv("YQ_FLAT2_DEBUG")) : 0;
        a.debug = dbg;
    }
    a.magic_w = (uint32_t)((0x100000000ull + W1 - 1) / W1);
    a.magic_h = (uint32_t)((0x100000000ull + H1 - 1) / H1);
    a.m_pairs = (int)((rows_alloc + 255) / 256);
    a.num_tiles = a.m_pairs * (st->n_pad / F2_BN);
    a.magic_m = a.m_pairs == 1 ? 0u : (uint32_t)((0x100000000ull + a.m_pairs - 1) / a.m_pairs);
    if ((long long)a.num_tiles * a.m_pairs >= 0x100000000ll) return yq::fail("tcgen05 flat2 flavour: too many tiles for 32-bit tile arithmetic");
    const CUtensorMap &tmA = it->second.first, &tmO = it->second.second;
    if (st->KC == 128) return f2_launch<128>(st, tmA, tmO, a, stream);
    if (st->KC == 32) return f2_launch<32>(st, tmA, tmO, a, stream);
    return f2_launch<64>(st, tmA, tmO, a, stream);
}
