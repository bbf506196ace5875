// yq_conv_tc_flat.cu -- tcgen05 (kind::i8) implicit-GEMM convolution over a FLAT halo-padded activation tensor:
// every filter tap reads the SAME shared-memory patch through a row-shifted descriptor (no per-tap activation loads).
//
// Flat geometry (yq_act_geom{pad 1, pitch_w W+1, rows_h H+1}): the batch is one long strip of "positions"
//     P(n, y, x) = (n*(H+1) + y + 1)*(W+1) + x + 1
// in which every image row is followed by ONE halo pixel (shared by the row's right edge and the next row's left edge)
// and every image by ONE halo row (shared likewise), plus a trailing halo row.  Halo positions hold the consumer's
// input zero point -- exactly what im2col_cpu_uint8 pads with (src/im2col.c:5-14) -- so tap (ky,kx) of output position
// p is simply position p + (ky-pad)*(W+1) + (kx-pad): a stride-1 convolution is a 1-D correlation over the strip.
//
// GEMM view:  D[M = 128 consecutive positions][N = BN output channels] += A[M][K] * B[N][K]^T,  K = (ky, kx, ci)
//   A  per KC-byte channel chunk ONE 2-D TMA load brings the patch [p0 - pad*(W+2), p0 + 128 + pad*(W+2)) x KC into
//      shared memory (SWIZZLE_128B / 64B, one position per row).  The MMA of tap (ky,kx) uses the descriptor start
//      patch + (ky*(W+1) + kx) * KC: the tensor core applies the swizzle to ABSOLUTE shared-memory address bits, so a
//      start shifted by whole rows reads the shifted rows correctly with base_offset = 0 (measured on B200:
//      tools/probes/probe_shift.cu).  Activation ingest per tile drops from 9 x 128 rows to 128 + 2W + 4 rows.
//   B  weights [oc][ky][kx][ci], one 2-D TMA load of BN x KC per (tap, chunk) through its own ring; 16 all-ones rows
//      appended to every stage give sum(a) per position in TMEM columns [BN, BN+16) (uint8 weights with a uint8 zero
//      point: acc = sum w*a - zp_w * sum a, convolutional_layer.c:718-721).
//   epilogue  thread = TMEM lane = position: integer-form requantize (yq_epilogue.cuh) -> swizzled smem tile -> TMA
//      store.  Halo positions of the OUTPUT strip are written with the consumer's zero point, so the kernel keeps its
//      output's halo intact by itself and garbage computed there never escapes.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-9 = epilogue (two per
// TMEM lane quarter, half of the channels each: the epilogue of a short-K layer is its critical path).
// One tile per CTA, two CTAs per SM (one CTA's epilogue overlaps the other's main loop).
// Restates convolutional_layer.c:694-761 for stride 1, pad = size/2, size in {1, 3}, c % 64 == 0.
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

// -DYQ_TIMELINE: every CTA records globaltimer (ns) at a few events into yq_flat_timeline[blockIdx.x * 8 + event]
#ifdef YQ_TIMELINE
__device__ unsigned long long yq_flat_timeline[8 * 16384];
__device__ __forceinline__ void tl_mark(int ev)
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const int b = blockIdx.y * gridDim.x + blockIdx.x;
    if (b < 16384) yq_flat_timeline[b * 8 + ev] = t;
}
#define TL_MARK(ev) tl_mark(ev)
#else
#define TL_MARK(ev)
#endif

__device__ unsigned int yq_flat_arrivals[256];

constexpr int FL_EPI_WARPS = 8;          // two warps per TMEM lane quarter, each takes half of the tile's channels
constexpr int FL_THREADS = 64 + 32 * FL_EPI_WARPS;
constexpr int FL_BSTAGES = 3;
constexpr int FL_ONES = 16;

struct FlatArgs {
    yq::EpiParams ep;
    float *out_f32;        // dense NCHW [B][N][H][W] (quant_stop) or nullptr
    float *out_yolo;       // same shape: the following yolo layer's output (logistic on x, y, objectness, classes), or nullptr
    const float *lut;      // [0,256): (u8 - zp_out) * s_out   [256,512): its logistic -- the head's float values are a 256-entry table
    int yolo_per;          // 4 + classes + 1 channels per anchor
    int32_t *out_acc;      // dense NHWC [B][H][W][CSO] (parity checks) or nullptr
    int N, CSO, n_pad;
    int B, H, W, NP;       // NP = B*(H+1)*(W+1): positions that belong to images (their halo included)
    int size, taps, cpt /* KC-chunks per tap */, CS;
    int q_off;             // first patch position relative to the tile's first position: -(pad*(W+1) + pad)
    int patch_rows, a_stage_bytes;
    uint32_t halo_word;    // byte the consumer pads with, replicated
    uint32_t magic_w, magic_h;   // ceil(2^32 / (W+1)), ceil(2^32 / (H+1))
    int n_sm, stagger_ns;        // first-wave stagger (see the kernel)
};

template <int BN>
__host__ __device__ constexpr int fl_tmem_cols()
{
    return BN + FL_ONES <= 32 ? 32 : BN + FL_ONES <= 64 ? 64 : BN + FL_ONES <= 128 ? 128 : 256;
}

template <int BN, int KC>
struct FlatSmem {
    static constexpr int B_BYTES = BN * KC;
    static constexpr int B_STAGE = (BN + FL_ONES) * KC;          // multiple of 1024 for every instantiated (BN, KC)
    static constexpr int PARAM_BYTES = BN * 24 + 2048 + 4 * BN;     // per-channel parameters, the head's 512-entry float table, its per-channel table offsets
    static_assert(B_STAGE % 1024 == 0, "stage buffers must keep the 1024-byte swizzle alignment");
    static_assert(128 * BN <= FL_BSTAGES * B_STAGE, "output staging aliases the weight ring");
};

// SLOW = the variant that also serves the int32 / float side outputs and the saturate switch.
template <int BN, int KC, bool SLOW>
__global__ void __launch_bounds__(FL_THREADS, 2) conv_u8_tc_flat_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                    const __grid_constant__ CUtensorMap tmO, const FlatArgs a)
{
    using L = FlatSmem<BN, KC>;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sA = smem;                                     // 2 patch stages of a.a_stage_bytes
    uint8_t *sB = smem + 2 * a.a_stage_bytes;               // FL_BSTAGES weight stages (+ constant ones rows)
    int4 *s_q = (int4 *)(sB + FL_BSTAGES * L::B_STAGE);     // {bias, zw, 2*M0, shift}
    double *s_mc = (double *)(s_q + BN);
    float *s_lut = (float *)(s_mc + BN);                    // yolo heads: (u8 - zp_out) * s_out, then its logistic
    int *s_sel = (int *)(s_lut + 512);                      // 0 / 256: which half of the table channel oc0 + i reads
    uint64_t *a_full = (uint64_t *)(s_sel + BN);
    uint64_t *a_empty = a_full + 2;
    uint64_t *b_full = a_empty + 2;
    uint64_t *b_empty = b_full + FL_BSTAGES;
    uint64_t *accum_full = b_empty + FL_BSTAGES;
    uint32_t *tmem_slot = (uint32_t *)(accum_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int oc0 = blockIdx.y * BN;
    const int p0 = blockIdx.x * 128;
    const int chunks = a.cpt;
    if (threadIdx.x == 0) TL_MARK(0);                       // CTA start
    // Two CTAs share an SM.  Launched together they run in lockstep -- both in the main loop (sharing the tensor pipe), then
    // both in the epilogue (pipe idle): measured 55 % pipe utilisation on layer 12.  The CTAs that take the SECOND slot of
    // every SM in the first wave therefore start half a tile period late; later waves inherit the offset because a new CTA
    // starts when an old one retires.
    if (threadIdx.x == 0 && a.stagger_ns > 0 && blockIdx.y * gridDim.x + blockIdx.x < 2 * a.n_sm) {
        // which of the SM's two first-wave CTAs am I?  (a never-reset arrival counter per SM: two CTAs that arrive together get
        // one even and one odd ticket whatever the count was)
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (atomicAdd(&yq_flat_arrivals[smid & 255], 1u) & 1u) {
            unsigned long long t0, t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            do {
                __nanosleep(200);
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            } while (t1 - t0 < (unsigned long long)a.stagger_ns);
        }
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < FL_BSTAGES; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        mbar_init(accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc<fl_tmem_cols<BN>()>(tmem_slot);
    if (warp >= 2) {
        const int t = threadIdx.x - 64;
        for (int i = t; i < BN; i += 32 * FL_EPI_WARPS) {
            s_q[i] = __ldg(a.ep.chanq + oc0 + i);
            s_mc[i] = __ldg(a.ep.mcomb + oc0 + i);
        }
        if (SLOW && a.out_yolo) {   // yolo_layer.c:137-146: channels 2, 3 (w, h) of every anchor stay linear, the rest go through the logistic
            for (int i = t; i < 512; i += 32 * FL_EPI_WARPS) s_lut[i] = __ldg(a.lut + i);
            for (int i = t; i < BN; i += 32 * FL_EPI_WARPS) {
                const int e = (oc0 + i) % a.yolo_per;
                s_sel[i] = (e == 2 || e == 3) ? 0 : 256;
            }
        }
        for (int s = 0; s < FL_BSTAGES; ++s) {   // the 16 all-ones filter rows behind the TMA-written BN rows of every stage
            uint32_t *ones = (uint32_t *)(sB + s * L::B_STAGE + L::B_BYTES);
            for (int i = t; i < FL_ONES * KC / 4; i += 32 * FL_EPI_WARPS) ones[i] = 0x01010101u;
        }
        fence_proxy_async();
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TL_MARK(1);                       // setup done
    yq_pdl_wait_then_release();                             // no activation tensor was touched so far

    if (warp == 0) {
        // ===================== TMA producer (whole warp in uniform control flow, one elected lane issues) =====================
        int it = 0;
        for (int c = 0; c < chunks; ++c) {
            const int sa = c & 1;
            mbar_wait(&a_empty[sa], ((c >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&a_full[sa], (uint32_t)(a.patch_rows * KC));
                tma_load_2d(sA + sa * a.a_stage_bytes, &tmA, &a_full[sa], c * KC, p0 + a.q_off);
            }
            for (int tap = 0; tap < a.taps; ++tap, ++it) {
                const int s = it % FL_BSTAGES;
                mbar_wait(&b_empty[s], ((it / FL_BSTAGES) & 1) ^ 1);
                if (elect_one()) {
#ifdef YQ_TIMELINE
                    if (a.stagger_ns == -1 && tap > 0) {     // experiment: skip the weight loads of taps 1.. (results are garbage)
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&b_full[s])) : "memory");
                        continue;
                    }
#endif
                    mbar_expect_tx(&b_full[s], (uint32_t)(BN * KC));
                    tma_load_2d(sB + s * L::B_STAGE, &tmB, &b_full[s], tap * a.CS + c * KC, oc0);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (same form) =====================
        constexpr uint32_t idesc = make_idesc(BN + FL_ONES);
        const uint32_t row_step = (uint32_t)((a.W + 1 - a.size) * KC);   // from the last tap of a filter row to the first of the next
        int it = 0;
        for (int c = 0; c < chunks; ++c) {
            const int sa = c & 1;
            mbar_wait(&a_full[sa], (c >> 1) & 1);
            uint32_t tap_addr = smem_u32(sA + sa * a.a_stage_bytes);    // row-shifted descriptor start of the current tap, base_offset 0
            int kx = 0;
            for (int tap = 0; tap < a.taps; ++tap, ++it) {
                const int s = it % FL_BSTAGES;
                mbar_wait(&b_full[s], (it / FL_BSTAGES) & 1);
                tc_fence_after();
                if (it == 0 && lane == 0) TL_MARK(2);       // first operands landed
                if (elect_one()) {
                    const uint64_t da = make_desc<KC>(tap_addr);
                    const uint64_t db = make_desc<KC>(smem_u32(sB + s * L::B_STAGE));
#pragma unroll
                    for (int k = 0; k < KC / 32; ++k) umma_i8(tmem_base, da + 2 * k, db + 2 * k, idesc, (it | k) ? 1u : 0u);
                    umma_commit(&b_empty[s]);
                }
                tap_addr += KC;
                if (++kx == a.size) { kx = 0; tap_addr += row_step; }
            }
            if (elect_one()) umma_commit(&a_empty[sa]);
        }
        if (elect_one()) umma_commit(accum_full);
        if (lane == 0) TL_MARK(3);                          // last MMA issued
    } else {
        // ===================== epilogue =====================
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;       // which half of the tile's channels this warp requantizes
        const int r = q * 32 + lane;            // tile row = TMEM lane = position p0 + r
        const int p = p0 + r;
        // position -> (image, y, x); halo positions (and positions past the last image) are not pixels
        const int row = (int)__umulhi((uint32_t)p, a.magic_w);
        const int col = p - row * (a.W + 1);
        const int n = (int)__umulhi((uint32_t)row, a.magic_h);
        const int y1 = row - n * (a.H + 1);
        const bool valid = p < a.NP && col >= 1 && y1 >= 1;
        uint8_t *stage_out = sB;                // aliases the weight ring (all MMAs have completed)
        const bool side = SLOW && ((a.out_acc != nullptr) || (a.out_f32 != nullptr) || (a.out_yolo != nullptr));
        const size_t pix = ((size_t)n * a.H + (y1 - 1)) * a.W + (col - 1);
        constexpr int HALF = BN / 2, NCHUNK = HALF / 16;
        const int cbeg = half * HALF;
        mbar_wait(accum_full, 0);
        tc_fence_after();
        if (threadIdx.x == 64) TL_MARK(4);                  // accumulator complete
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t vbuf[2][16];
        tmem_ld16_issue(trow + cbeg, vbuf[0]);
        const int nsa = -(int)tmem_ld1(trow + BN);   // minus the position's activation sum (ones-row columns); also completes the load above
        auto run = [&](auto actm_tag, auto sat_tag) {
            constexpr int ACTM = decltype(actm_tag)::value;
            constexpr bool SAT = decltype(sat_tag)::value;
#pragma unroll
            for (int ch = 0; ch < NCHUNK; ++ch) {
                const int c0 = cbeg + 16 * ch;
                uint32_t(&v)[16] = vbuf[ch & 1];
                if (ch + 1 < NCHUNK) tmem_ld16_issue(trow + c0 + 16, vbuf[(ch + 1) & 1]);   // in flight while this chunk is requantized
                uint32_t packed[4];
                int extra[16];
                yq::requant_chunk<ACTM, SAT, 16, false>(v, nsa, extra, s_q + c0, s_mc + c0, a.ep.zp_out, packed, a.ep.xlim);
                if (!valid) packed[0] = packed[1] = packed[2] = packed[3] = a.halo_word;
                yq::mask_pad_channels<16>(packed, a.N - (oc0 + c0));
                if (SLOW && side && valid) {
                    const int hw = a.H * a.W;
                    const size_t f0 = ((size_t)n * a.N + oc0 + c0) * hw + (size_t)(y1 - 1) * a.W + (col - 1);
                    const int nreal = a.N - (oc0 + c0);
                    if (a.out_yolo && !a.out_acc && !a.out_f32) {
                        // the detection heads on the throughput path: one table lookup and one 4-byte store per output
                        float *dst = a.out_yolo + f0;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (j < nreal) dst[j * hw] = s_lut[s_sel[c0 + j] + ((packed[j / 4] >> (8 * (j % 4))) & 255u)];
                    } else {   // parity / quant_stop side outputs (not on the throughput path)
#pragma unroll 1
                        for (int j = 0; j < 16; ++j) {
                            if (j < nreal) {
                                const int oc = oc0 + c0 + j;
                                if (a.out_acc) a.out_acc[pix * a.CSO + oc] = (int)v[j] + s_q[c0 + j].y * nsa;
                                const uint8_t u = (uint8_t)(packed[j / 4] >> (8 * (j % 4)));
                                const size_t fidx = f0 + (size_t)j * hw;
                                if (a.out_f32) a.out_f32[fidx] = yq::dequant_f32(a.ep, u);
                                if (a.out_yolo) a.out_yolo[fidx] = s_lut[s_sel[c0 + j] + u];
                            }
                        }
                    }
                }
                {   // swizzled staging (matches the TMA-store tensor map): 16-byte chunk index XOR row bits
                    const int chunk = c0 / 16;
                    int sw;
                    if (BN >= 128) sw = chunk ^ (r & 7);
                    else if (BN == 64) sw = chunk ^ ((r >> 1) & 3);
                    else sw = chunk ^ ((r >> 2) & 1);
                    *reinterpret_cast<uint4 *>(stage_out + (size_t)r * BN + sw * 16) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                }
                if (ch + 1 < NCHUNK) tmem_ld_wait16(vbuf[(ch + 1) & 1]);
            }
        };
        const int actm = yq::act_mode(a.ep.act);
        if (SLOW && a.ep.saturate) {
            if (actm == 0) run(std::integral_constant<int, 0>{}, std::true_type{});
            else if (actm == 1) run(std::integral_constant<int, 1>{}, std::true_type{});
            else run(std::integral_constant<int, 2>{}, std::true_type{});
        } else {
            if (actm == 0) run(std::integral_constant<int, 0>{}, std::false_type{});
            else if (actm == 1) run(std::integral_constant<int, 1>{}, std::false_type{});
            else run(std::integral_constant<int, 2>{}, std::false_type{});
        }
        tc_fence_before();
        fence_proxy_async();
        asm volatile("bar.sync 1, %0;" ::"n"(32 * FL_EPI_WARPS) : "memory");   // the epilogue warps only
        if (threadIdx.x == 64) {
            TL_MARK(5);                                     // epilogue math done
            tma_store_2d(&tmO, stage_out, oc0, p0);
            tma_store_commit_wait();
            TL_MARK(6);                                     // store drained
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<fl_tmem_cols<BN>()>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn fl_get_encode()
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

CUtensorMapSwizzle fl_swizzle_for(int inner_bytes)
{
    return inner_bytes >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : inner_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

int fl_encode_2d(CUtensorMap *m, const void *ptr, uint64_t rows, int row_bytes, int box_c, int box_rows, CUtensorMapL2promotion prom)
{
    EncodeTiledFn enc = fl_get_encode();
    if (!enc) return yq::fail("cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     fl_swizzle_for(box_c), prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return yq::fail("cuTensorMapEncodeTiled(%llu x %d, box %d x %d) failed: %d", (unsigned long long)rows, row_bytes, box_c, box_rows, (int)r);
    return 0;
}

struct FlatState {
    int BN, KC, n_pad;
    uint8_t *w = nullptr;       // [n_pad][size*size*cs_in]
    float *lut = nullptr;       // quant_stop layers: dequantized value and its logistic for each of the 256 output bytes
    CUtensorMap tmB;
    struct Key {
        const void *in;
        void *out;
        int batch;
        bool operator<(const Key &o) const { return in != o.in ? in < o.in : out != o.out ? out < o.out : batch < o.batch; }
    };
    std::map<Key, std::pair<CUtensorMap, CUtensorMap>> maps;
};

template <int BN, int KC, bool SLOW>
int fl_launch_v(FlatState *st, const CUtensorMap &tmA, const CUtensorMap &tmO, const FlatArgs &a, dim3 grid, cudaStream_t stream)
{
    using L = FlatSmem<BN, KC>;
    const int smem = 2 * a.a_stage_bytes + FL_BSTAGES * L::B_STAGE + L::PARAM_BYTES + 128 + 1024;
    auto kern = conv_u8_tc_flat_kernel<BN, KC, SLOW>;
    if (yq::ensure_dynamic_smem((const void *)kern, smem)) return -1;
    YQ_CUDA(yq::launch_pdl(kern, grid, dim3(FL_THREADS), smem, stream, tmA, st->tmB, tmO, a));
    return 0;
}

template <int BN, int KC>
int fl_launch(FlatState *st, const CUtensorMap &tmA, const CUtensorMap &tmO, const FlatArgs &a, dim3 grid, cudaStream_t stream)
{
    if (a.out_acc || a.out_f32 || a.out_yolo || a.ep.saturate) return fl_launch_v<BN, KC, true>(st, tmA, tmO, a, grid, stream);
    return fl_launch_v<BN, KC, false>(st, tmA, tmO, a, grid, stream);
}

}  // namespace

// what every flat-strip flavour needs of a layer (each adds its own limit on the row width)
int yq_tc_flat_eligible(const yq_conv_layer *l)
{
    if (!l->int_form || !l->fused_mult) return 0;
    if (l->stride != 1 || !(l->size == 1 || l->size == 3) || l->pad != l->size / 2) return 0;
    // no pad lanes: the halo fill would count in sum(a).  c = 32 (32-byte rows, SWIZZLE_32B, one K step per tap): flat2 only
    if (l->c != l->cs_in || (l->cs_in % 64 && l->cs_in != 32)) return 0;
    if (l->cs_out < 32) return 0;
    return fl_get_encode() != nullptr;
}

int yq_tc_flat_supported(const yq_conv_layer *l)
{
    if (!yq_tc_flat_eligible(l) || l->cs_in % 64) return 0;
    if (128 + (l->size - 1) * (l->w + 2) > 256) return 0;       // the patch is one TMA box (<= 256 rows)
    return 1;
}

void yq_tc_flat_geom(int h, int w, yq_act_geom *g)
{
    g->pad = 1;
    g->pitch_w = w + 1;
    g->rows_h = h + 1;
}

int yq_tc_flat_prepare(yq_conv_layer *l, void **state)
{
    FlatState *st = new FlatState();
    st->BN = l->cs_out >= 128 ? 128 : (l->cs_out >= 64 ? 64 : 32);
    st->KC = (l->cs_in % 128) ? 64 : 128;
    st->n_pad = yq::round_up(l->n, st->BN);
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
        cudaFree(st->lut);
        delete st;
        return -1;
    };
    if (!on_dev) {
        if (cudaMalloc((void **)&st->w, wp.size()) != cudaSuccess) return cleanup();
        if (cudaMemcpy(st->w, wp.data(), wp.size(), cudaMemcpyHostToDevice) != cudaSuccess) return cleanup();
    }
    if (fl_encode_2d(&st->tmB, st->w, (uint64_t)st->n_pad, (int)ktot, st->KC, st->BN, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return cleanup();
    if (l->quant_stop_flag) {
        // l.output = (u8 - zp_out) * s_out (convolutional_layer.c:752-760) takes 256 values; logistic_activate
        // (activations.h:32: 1./(1. + exp(-x)) in double, stored as float) of each is tabulated with the host's libm,
        // i.e. with the very arithmetic the reference runs
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

void yq_tc_flat_free(void *state)
{
    FlatState *st = (FlatState *)state;
    if (!st) return;
    cudaFree(st->w);
    cudaFree(st->lut);
    delete st;
}

int yq_tc_flat_forward(yq_conv_layer *l, void *state, const uint8_t *in_flat, uint8_t *out_flat, int halo_fill, float *out_f32, float *out_yolo,
                       int yolo_classes, int32_t *out_acc, int batch, cudaStream_t stream)
{
    FlatState *st = (FlatState *)state;
    if (!st || !in_flat || !out_flat) return yq::fail("tcgen05 flat flavour: bad argument");
    const int W1 = l->w + 1, H1 = l->h + 1;
    const long long NP = (long long)batch * H1 * W1;
    const long long rows_alloc = NP + W1 + 2;     // + the trailing halo row (yq_act_geom_bytes)
    if (rows_alloc * W1 >= 0x100000000ll) return yq::fail("tcgen05 flat flavour: tensor too large for 32-bit position arithmetic");
    FlatState::Key key{in_flat, out_flat, batch};
    FlatArgs a;
    memset(&a, 0, sizeof a);
    const int pad = l->size / 2;
    a.patch_rows = 128 + (l->size - 1) * (W1 + 1);
    a.a_stage_bytes = yq::round_up(a.patch_rows * st->KC, 1024);
    auto it = st->maps.find(key);
    if (it == st->maps.end()) {
        if (st->maps.size() > 64) st->maps.clear();
        CUtensorMap tmA, tmO;
        if (fl_encode_2d(&tmA, in_flat, (uint64_t)rows_alloc, l->cs_in, st->KC, a.patch_rows, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return -1;
        if (fl_encode_2d(&tmO, out_flat, (uint64_t)rows_alloc, l->cs_out, st->BN, 128, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return -1;
        it = st->maps.emplace(key, std::make_pair(tmA, tmO)).first;
    }
    a.ep = yq::make_epi(l);
    a.out_f32 = l->quant_stop_flag ? out_f32 : nullptr;
    a.out_yolo = l->quant_stop_flag ? out_yolo : nullptr;
    a.lut = st->lut;
    a.yolo_per = 4 + yolo_classes + 1;
    if (a.out_yolo && (!st->lut || l->n % a.yolo_per)) return yq::fail("tcgen05 flat flavour: layer is not a yolo head for %d classes", yolo_classes);
    a.out_acc = out_acc;
    a.N = l->n; a.CSO = l->cs_out; a.n_pad = st->n_pad;
    a.B = batch; a.H = l->h; a.W = l->w; a.NP = (int)NP;
    a.size = l->size; a.taps = l->size * l->size; a.cpt = l->cs_in / st->KC; a.CS = l->cs_in;
    a.q_off = -(pad * W1 + pad);
    a.halo_word = 0x01010101u * (uint32_t)(halo_fill & 0xff);
    a.magic_w = (uint32_t)((0x100000000ull + W1 - 1) / W1);
    a.magic_h = (uint32_t)((0x100000000ull + H1 - 1) / H1);
    dim3 grid((unsigned)((rows_alloc + 127) / 128), st->n_pad / st->BN);
    {
        const int n_sm = yq::device_sm_count();
        if (n_sm <= 0) return yq::fail("cannot query the device's multiprocessor count");
        static const int want = getenv("YQ_FLAT_STAGGER") ? atoi(getenv("YQ_FLAT_STAGGER")) : 1;   // 0 disables the first-wave stagger (A/B measurements)
        a.n_sm = n_sm;
        // half of one CTA's period: main loop alone on the pipe (N/2 clocks per MMA at ~1.9 GHz) + epilogue + prologue
        const double main_ns = (double)a.taps * a.cpt * (st->KC / 32) * ((st->BN + FL_ONES) / 2) / 1.9;
        a.stagger_ns = want && (long long)grid.x * grid.y > 2LL * n_sm ? (int)((main_ns + 3500.0) / 2.0) : 0;
        if (want < 0) a.stagger_ns = -1;
    }
    const CUtensorMap &tmA = it->second.first, &tmO = it->second.second;
#define YQ_FL(BN_, KC_) return fl_launch<BN_, KC_>(st, tmA, tmO, a, grid, stream)
    if (st->KC == 128) {
        if (st->BN == 128) YQ_FL(128, 128);
        if (st->BN == 64) YQ_FL(64, 128);
        YQ_FL(32, 128);
    } else {
        if (st->BN == 128) YQ_FL(128, 64);
        if (st->BN == 64) YQ_FL(64, 64);
        YQ_FL(32, 64);
    }
#undef YQ_FL
}

#ifdef YQ_TIMELINE
extern "C" __attribute__((visibility("default"))) int yq_debug_flat_timeline(void *host, size_t bytes)
{
    return cudaMemcpyFromSymbol(host, yq_flat_timeline, bytes < sizeof(yq_flat_timeline) ? bytes : sizeof(yq_flat_timeline)) == cudaSuccess ? 0 : -1;
}
#endif
