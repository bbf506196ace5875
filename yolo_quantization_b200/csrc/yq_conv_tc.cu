// yq_conv_tc.cu -- tcgen05 (kind::i8, u8 x u8 -> s32 in TMEM) implicit-GEMM convolution for sm_100a.
//
// GEMM view:  D[M = 128 output pixels][N = BN output channels] += A[M][K] * B[N][K]^T,  K = (ky, kx, ci)
//
//   A  activations, uint8 NHWC.  One pipeline stage = one filter tap (ky,kx) x one KC-byte channel chunk of a
//      TW x TH x TN patch of output pixels (TW*TH*TN <= 128 rows): a single 4-D *tiled* TMA load whose start
//      coordinate is shifted by (kx-pad, ky-pad).  Out-of-image taps are ZERO-filled by the TMA unit; the
//      reference pads with zp_in (im2col.c:5-14), which the epilogue restores exactly with a per-tap
//      correction  zp_in * sum_ci (w - zp_w)  on border pixels.
//   B  weights packed once at load as [oc][ky][kx][ci] (K-major), 2-D TMA load of a BN x KC tile.
//   D  int32 accumulator in TMEM, columns [0,BN).  Columns [BN,BN+16) accumulate A x ONES (16 constant all-ones
//      rows appended to every stage's B tile, so one N = BN+16 MMA per K step reads A from shared memory once)
//      = sum of activations per pixel, which the epilogue needs
//      because weights are uint8 with a per-channel uint8 zero point (9-bit w - zp_w does not fit s8):
//          acc = sum w*a - zp_w[oc] * sum a                      (convolutional_layer.c:718-721 restated)
//   epilogue (4 warps, one TMEM lane = one pixel per thread): tcgen05.ld -> zero-point correction ->
//      FP64 requantize -> activation -> +zp_out -> uint8 wrap -> swizzled smem tile -> TMA store (clips
//      partial tiles).  Optional int32 / float (quant_stop) side outputs go straight to global memory.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-9 = epilogue (two per TMEM lane
// quarter, each on half of the tile's channels: the epilogue is the longest phase of a tile, r2 profile).
// One output tile per CTA, up to two CTAs per SM so one CTA's epilogue overlaps the other's main loop.
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include <map>
#include <vector>

#include "yq_common.h"
#include "yq_epilogue.cuh"
#include "yq_tc_ptx.cuh"

using namespace yqtc;

namespace {

constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 32 * (2 + TC_EPI_WARPS);
constexpr int TC_STAGES = 3;
constexpr int ONES_ROWS = 16;

struct TcArgs {
    yq::EpiParams ep;
    float *out_f32;
    int32_t *out_acc;
    const int32_t *corr;      // [size*size][n_pad] border correction, or nullptr when zp_in == 0
    int B, OH, OW, N, CSO, n_pad;
    int TW, TH, TN, tiles_x, tiles_y;
    int size, pad, cpt /* KC-chunks per tap */, CS;
    int H, W;
    int stride;        // 1 or 2: the activation box steps through the input with elementStrides = stride
    int cpad;          // coordinate offset of pixel (0, 0): l.pad when the map starts in the input's halo (which then holds zp_in), else 0
};

template <int BN>
__host__ __device__ constexpr int tmem_cols()
{
    return BN + ONES_ROWS <= 32 ? 32 : BN + ONES_ROWS <= 64 ? 64 : BN + ONES_ROWS <= 128 ? 128 : BN + ONES_ROWS <= 256 ? 256 : 512;
}

template <int BN, int KC>
struct SmemLayout {
    static constexpr int A_BYTES = 128 * KC;
    static constexpr int B_BYTES = BN * KC;                  // written by TMA every stage
    static constexpr int ONES_BYTES = ONES_ROWS * KC;        // rows BN..BN+15 of the B tile: constant 0x01, written once
    static constexpr int STAGE = A_BYTES + B_BYTES + ONES_BYTES;   // multiple of 1024
    static constexpr int PARAM_OFF = TC_STAGES * STAGE;      // int4 {bias, zw, 2*M0, shift}[BN], double mcomb[BN]
    static constexpr int PARAM_BYTES = BN * 24;
    static_assert(STAGE % 1024 == 0, "stage buffers must keep the 1024-byte swizzle alignment");
    static constexpr int BAR_OFF = PARAM_OFF + PARAM_BYTES;
    static constexpr int TOTAL = BAR_OFF + 128;
    static_assert(128 * BN <= STAGE, "output staging aliases stage 0");
};

// activation / saturation are uniform per launch: dispatch once per 32-output chunk, not per output
template <bool HAS_EXTRA>
__device__ __forceinline__ void epi_chunk(int actm, int sat, const uint32_t (&v)[16], int nsa, const int (&extra)[16], const int4 *cq,
                                          const double *mc, int zo, uint32_t (&packed)[4], uint32_t xlim)
{
    if (sat) {
        if (actm == 0) yq::requant_chunk<0, true, 16, HAS_EXTRA>(v, nsa, extra, cq, mc, zo, packed, xlim);
        else if (actm == 1) yq::requant_chunk<1, true, 16, HAS_EXTRA>(v, nsa, extra, cq, mc, zo, packed, xlim);
        else yq::requant_chunk<2, true, 16, HAS_EXTRA>(v, nsa, extra, cq, mc, zo, packed, xlim);
    } else {
        if (actm == 0) yq::requant_chunk<0, false, 16, HAS_EXTRA>(v, nsa, extra, cq, mc, zo, packed, xlim);
        else if (actm == 1) yq::requant_chunk<1, false, 16, HAS_EXTRA>(v, nsa, extra, cq, mc, zo, packed, xlim);
        else yq::requant_chunk<2, false, 16, HAS_EXTRA>(v, nsa, extra, cq, mc, zo, packed, xlim);
    }
}

// SLOW = the variant that also serves the int32 / float side outputs and the saturate switch.
// CM x CN = thread-block cluster shape over (m-tile, n-tile).  CTAs of a cluster that share the m-tile take turns
// loading the activation stage and MULTICAST it to each other (same for the weight stage across CTAs sharing the
// n-tile), which divides the L2 -> shared-memory operand traffic (the measured bound of this kernel) by up to 2.
template <int BN, int KC, bool SLOW, int CM, int CN>
__global__ void __launch_bounds__(TC_THREADS, 2) conv_u8_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                               const __grid_constant__ CUtensorMap tmO, const TcArgs a)
{
    using L = SmemLayout<BN, KC>;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    uint64_t *full = (uint64_t *)(smem + L::BAR_OFF);
    uint64_t *empty = full + TC_STAGES;
    uint64_t *accum_full = empty + TC_STAGES;
    uint32_t *tmem_slot = (uint32_t *)(accum_full + 1);
    int4 *s_q = (int4 *)(smem + L::PARAM_OFF);          // {bias, zw, 2*M0, shift}
    double *s_mc = (double *)(s_q + BN);                // M_value * 2^-s (FP64 slow path)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool CLUSTER = CM * CN > 1;
    const int cx = CLUSTER ? (int)blockIdx.x % CM : 0, cy = CLUSTER ? (int)blockIdx.y % CN : 0;   // position in the cluster
    uint16_t mask_a = 0, mask_b = 0;      // cluster ranks (x fastest) sharing my activation tile / my weight tile
    for (int j = 0; j < CN; ++j) mask_a |= (uint16_t)(1u << (cx + CM * j));
    for (int i = 0; i < CM; ++i) mask_b |= (uint16_t)(1u << (i + CM * cy));
    const int oc0 = blockIdx.y * BN;
    const int tx = blockIdx.x % a.tiles_x;
    const int ty = (blockIdx.x / a.tiles_x) % a.tiles_y;
    const int tb = blockIdx.x / (a.tiles_x * a.tiles_y);
    const int x0 = tx * a.TW, y0 = ty * a.TH, n0 = tb * a.TN;
    const int rows = a.TW * a.TH * a.TN;
    const int kiters = a.size * a.size * a.cpt;

    // ---- one-time setup
    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CM + CN - 1);   // every CTA that my multicasts write into must have released the slot
        }
        mbar_init(accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc<tmem_cols<BN>()>(tmem_slot);
    if (warp >= 2) {
        const int t = threadIdx.x - 64;
        for (int i = t; i < BN; i += 32 * TC_EPI_WARPS) {
            const bool in = oc0 + i < a.n_pad;       // cluster padding may add CTAs past the last n-tile
            s_q[i] = in ? __ldg(a.ep.chanq + oc0 + i) : make_int4(0, 0, 0, 0);
            s_mc[i] = in ? __ldg(a.ep.mcomb + oc0 + i) : 0.0;
        }
        // the 16 all-ones filter rows that follow the TMA-written BN rows of every stage's B tile (never overwritten)
        for (int s = 0; s < TC_STAGES; ++s) {
            uint32_t *ones = (uint32_t *)(smem + s * L::STAGE + L::A_BYTES + L::B_BYTES);
            for (int i = t; i < L::ONES_BYTES / 4; i += 32 * TC_EPI_WARPS) ones[i] = 0x01010101u;
        }
        fence_proxy_async();   // generic-proxy writes of the ones rows must be visible to the tensor core (async proxy)
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
    }
    tc_fence_before();
    if (CLUSTER) cluster_sync_all();      // barrier inits of every CTA visible before any multicast / remote arrive
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)(rows * KC + BN * KC);
            for (int it = 0; it < kiters; ++it) {
                const int s = it % TC_STAGES;
                const uint32_t ph = (it / TC_STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], bytes);
                const int tap = it / a.cpt, chunk = it - tap * a.cpt;
                const int ky = tap / a.size, kx = tap - ky * a.size;
                uint8_t *sa = smem + s * L::STAGE;
                const int cx0 = x0 * a.stride + kx - a.pad + a.cpad, cy0 = y0 * a.stride + ky - a.pad + a.cpad;
                if (!CLUSTER) {
                    tma_load_4d(sa, &tmA, &full[s], chunk * KC, cx0, cy0, n0);
                    tma_load_2d(sa + L::A_BYTES, &tmB, &full[s], tap * a.CS + chunk * KC, oc0);
                } else {
                    // take turns: the CTA whose turn it is loads the stage for everyone sharing that operand
                    if (CN == 1) tma_load_4d(sa, &tmA, &full[s], chunk * KC, cx0, cy0, n0);
                    else if (it % CN == cy) tma_load_4d_mc(sa, &tmA, &full[s], chunk * KC, cx0, cy0, n0, mask_a);
                    if (CM == 1) tma_load_2d(sa + L::A_BYTES, &tmB, &full[s], tap * a.CS + chunk * KC, oc0);
                    else if (it % CM == cx) tma_load_2d_mc(sa + L::A_BYTES, &tmB, &full[s], tap * a.CS + chunk * KC, oc0, mask_b);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // ONE instruction per K step: N = BN filter rows + 16 all-ones rows, so A is read from smem once and
            // TMEM columns [BN, BN+16) receive the per-pixel activation sum
            constexpr uint32_t idesc_main = make_idesc(BN + ONES_ROWS);
            for (int it = 0; it < kiters; ++it) {
                const int s = it % TC_STAGES;
                const uint32_t ph = (it / TC_STAGES) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * L::STAGE);
                const uint64_t da = make_desc<KC>(sa), db = make_desc<KC>(sa + L::A_BYTES);
#pragma unroll
                for (int k = 0; k < KC / 32; ++k) {
                    const uint32_t accum = (it | k) ? 1u : 0u;
                    umma_i8(tmem_base, da + 2 * k, db + 2 * k, idesc_main, accum);       // +32 bytes = +2 in 16-byte units
                }
                if (CLUSTER) umma_commit_mc(&empty[s], (uint16_t)(mask_a | mask_b));   // release the slot towards every CTA that fills it
                else umma_commit(&empty[s]);
            }
            umma_commit(accum_full);
        }
    } else {
        // ===================== epilogue =====================
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;       // which half of the tile's channels (BN >= 32: at least one 16-channel chunk each)
        const int r = q * 32 + lane;            // tile row = TMEM lane = output pixel of the patch
        const int wi = r % a.TW, hi = (r / a.TW) % a.TH, ni = r / (a.TW * a.TH);
        const int ox = x0 + wi, oy = y0 + hi, n = n0 + ni;
        const bool valid = r < rows && ox < a.OW && oy < a.OH && n < a.B;
        // taps that fall outside the image for this pixel (zero-filled by TMA, zp_in in the reference)
        uint32_t oob = 0;
        if (a.corr && valid) {
            for (int ky = 0; ky < a.size; ++ky)
                for (int kx = 0; kx < a.size; ++kx) {
                    const int iy = oy * a.stride + ky - a.pad, ix = ox * a.stride + kx - a.pad;
                    if (iy < 0 || iy >= a.H || ix < 0 || ix >= a.W) oob |= 1u << (ky * a.size + kx);
                }
        }
        mbar_wait(accum_full, 0);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        const int nsa = -(int)tmem_ld1(trow + BN);   // minus the pixel's activation sum (ones-tile columns)
        uint8_t *stage_out = smem;              // aliases pipeline stage 0 (all MMAs have completed)
        const size_t pix = ((size_t)n * a.OH + oy) * a.OW + ox;
        const int actm = yq::act_mode(a.ep.act);
        const bool side = SLOW && ((a.out_acc != nullptr) || (a.out_f32 != nullptr));
        const int sat = SLOW ? a.ep.saturate : 0;
#pragma unroll 1
        for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 16) {
            uint32_t v[16];
            tmem_ld16(trow + c0, v);
            uint32_t packed[4];
            int extra[16];
            if (oob) {
#pragma unroll
                for (int j = 0; j < 16; ++j) extra[j] = 0;
                for (uint32_t m = oob; m; m &= m - 1) {
                    const int32_t *cr = a.corr + (size_t)(__ffs(m) - 1) * a.n_pad + oc0 + c0;
#pragma unroll
                    for (int j = 0; j < 16; ++j) extra[j] += __ldg(cr + j);
                }
                epi_chunk<true>(actm, sat, v, nsa, extra, s_q + c0, s_mc + c0, a.ep.zp_out, packed, a.ep.xlim);
            } else {
                epi_chunk<false>(actm, sat, v, nsa, extra, s_q + c0, s_mc + c0, a.ep.zp_out, packed, a.ep.xlim);
            }
            yq::mask_pad_channels<16>(packed, a.N - (oc0 + c0));
            if (SLOW && side && valid) {   // parity / quant_stop side outputs (not on the throughput path)
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int oc = oc0 + c0 + j;
                    if (oc < a.N) {
                        if (a.out_acc) a.out_acc[pix * a.CSO + oc] = (int)v[j] + s_q[c0 + j].y * nsa + (oob ? extra[j] : 0);
                        if (a.out_f32) {
                            const uint8_t u = (uint8_t)(packed[j / 4] >> (8 * (j % 4)));
                            a.out_f32[((size_t)n * a.N + oc) * a.OH * a.OW + (size_t)oy * a.OW + ox] = yq::dequant_f32(a.ep, u);
                        }
                    }
                }
            }
            // swizzled staging (matches the TMA-store tensor map): 16-byte chunk index XOR row bits
            {
                const int chunk = c0 / 16;
                int sw;
                if (BN >= 128) sw = chunk ^ (r & 7);
                else if (BN == 64) sw = chunk ^ ((r >> 1) & 3);
                else sw = chunk ^ ((r >> 2) & 1);
                *reinterpret_cast<uint4 *>(stage_out + (size_t)r * BN + sw * 16) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            }
        }
        tc_fence_before();
        fence_proxy_async();
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");   // the epilogue warps only
        if (threadIdx.x == 64) {
            tma_store_4d(&tmO, stage_out, oc0, x0, y0, n0);
            tma_store_commit_wait();
        }
    }
    tc_fence_before();
    if (CLUSTER) cluster_sync_all();      // nobody exits while a peer may still multicast into it or arrive on its barriers
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<tmem_cols<BN>()>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode()
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

CUtensorMapSwizzle swizzle_for(int inner_bytes)
{
    return inner_bytes >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : inner_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

struct TcState {
    int BN, KC, TW, TH, TN, n_pad;
    uint8_t *w = nullptr;       // [n_pad][size*size*cs_in]
    int32_t *corr = nullptr;    // [size*size][n_pad]
    CUtensorMap tmB;
    // tensor maps for the activations depend on the pointers handed to forward(); cache the last few
    struct Key {
        const void *in;
        void *out;
        int batch;
        yq_act_geom gi, go;
        int halo;
        bool operator<(const Key &o) const
        {
            if (in != o.in) return in < o.in;
            if (out != o.out) return out < o.out;
            if (batch != o.batch) return batch < o.batch;
            if (halo != o.halo) return halo < o.halo;
            const int a[6] = {gi.pad, gi.pitch_w, gi.rows_h, go.pad, go.pitch_w, go.rows_h};
            const int b[6] = {o.gi.pad, o.gi.pitch_w, o.gi.rows_h, o.go.pad, o.go.pitch_w, o.go.rows_h};
            for (int i = 0; i < 6; ++i)
                if (a[i] != b[i]) return a[i] < b[i];
            return false;
        }
    };
    std::map<Key, std::pair<CUtensorMap, CUtensorMap>> maps;
};

// 4-D map over an NHWC tensor stored in geometry g (plain when g is null).  halo > 0: coordinate (0, 0) is the halo pixel
// (-halo, -halo) and the extents grow by 2 * halo (rows / images overlap their neighbours' shared halo: fine for a tensor map,
// only the strides must be multiples of 16 bytes); halo = 0: the h x w interior, everything else out of bounds (zero fill on
// loads, clipped on stores).  estride: the box picks every estride-th pixel in x and y (TW x TH pixels either way).
int encode_nhwc(CUtensorMap *m, const void *ptr, const yq_act_geom *g, int halo, int B, int H, int W, int CS, int box_c, int TW, int TH, int TN, int estride)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) return yq::fail("cuTensorMapEncodeTiled is not available from this driver");
    const int gpad = g ? g->pad : 0, pitch = g ? g->pitch_w : W, rows = g ? g->rows_h : H;
    const uint8_t *base = (const uint8_t *)ptr + ((size_t)(gpad - halo) * pitch + (gpad - halo)) * CS;
    cuuint64_t dims[4] = {(cuuint64_t)CS, (cuuint64_t)(W + 2 * halo), (cuuint64_t)(H + 2 * halo), (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)CS, (cuuint64_t)pitch * CS, (cuuint64_t)rows * pitch * CS};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)(TW * estride), (cuuint32_t)(TH * estride), (cuuint32_t)TN};   // N pixels = box N * stride
    cuuint32_t es[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<uint8_t *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle_for(box_c), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return yq::fail("cuTensorMapEncodeTiled(NHWC %dx%dx%dx%d pitch %d rows %d box %d,%d,%d,%d stride %d) failed: %d", B, H, W, CS, pitch, rows, box_c, TW, TH,
                        TN, estride, (int)r);
    return 0;
}

// choose the output-pixel patch TW x TH x TN (<= 128 rows) that wastes the fewest MMA rows
void choose_tile(int B, int OH, int OW, int *tw, int *th, int *tn)
{
    double best = -1;
    for (int w = 1; w <= OW && w <= 128; ++w) {   // (box extent w * stride <= 256 holds for stride <= 2)
        const int txs = (OW + w - 1) / w;
        for (int h = 1; h <= OH && w * h <= 128; ++h) {
            const int tys = (OH + h - 1) / h;
            int n = 128 / (w * h);
            if (n > B) n = B;
            if (n < 1) n = 1;
            const int tbs = (B + n - 1) / n;
            const double eff = (double)B * OH * OW / ((double)txs * tys * tbs * 128.0);
            // prefer fewer, wider boxes on ties (longer contiguous TMA rows)
            const double score = eff + 1e-6 * w;
            if (score > best) {
                best = score;
                *tw = w; *th = h; *tn = n;
            }
        }
    }
}

template <int BN, int KC, bool SLOW, int CM, int CN>
int launch_v(TcState *st, const CUtensorMap &tmA, const CUtensorMap &tmO, const TcArgs &a, dim3 grid, cudaStream_t stream)
{
    using L = SmemLayout<BN, KC>;
    const int smem = L::TOTAL + 1024;
    auto kern = conv_u8_tc_kernel<BN, KC, SLOW, CM, CN>;
    if (yq::ensure_dynamic_smem((const void *)kern, smem)) return -1;
    if (CM * CN == 1) {
        kern<<<grid, TC_THREADS, smem, stream>>>(tmA, st->tmB, tmO, a);
    } else {
        grid.x = (grid.x + CM - 1) / CM * CM;   // pad to whole clusters: the extra CTAs see only out-of-range tiles
        grid.y = (grid.y + CN - 1) / CN * CN;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = grid;
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CM;
        attr[0].val.clusterDim.y = CN;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        YQ_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, st->tmB, tmO, a));
    }
    YQ_CHECK_LAUNCH();
    return 0;
}

template <int BN, int KC, int CM, int CN>
int launch_c(TcState *st, const CUtensorMap &tmA, const CUtensorMap &tmO, const TcArgs &a, dim3 grid, cudaStream_t stream)
{
    if (a.out_acc || a.out_f32 || a.ep.saturate) return launch_v<BN, KC, true, CM, CN>(st, tmA, tmO, a, grid, stream);
    return launch_v<BN, KC, false, CM, CN>(st, tmA, tmO, a, grid, stream);
}

template <int BN, int KC>
int launch(yq_conv_layer *l, TcState *st, const CUtensorMap &tmA, const CUtensorMap &tmO, const TcArgs &a, dim3 grid, cudaStream_t stream)
{
    // Cluster multicast is OFF by default: measured on B200 (profiles/README.md) it is 10-15 % SLOWER than independent
    // CTAs -- the bound is the per-SM shared-memory ingest rate, which multicast does not reduce (every SM still
    // receives every byte); it only saves L2 reads.  YQ_TC_CLUSTER=1 enables it for A/B measurements.
    static int want = -1;
    if (want < 0) {
        const char *e = getenv("YQ_TC_CLUSTER");
        want = e ? atoi(e) : 0;
    }
    if constexpr (BN == 128) {
        if (want && grid.x >= 2) {
            if (grid.y % 2 == 0) return launch_c<BN, KC, 2, 2>(st, tmA, tmO, a, grid, stream);
            return launch_c<BN, KC, 2, 1>(st, tmA, tmO, a, grid, stream);
        }
    }
    return launch_c<BN, KC, 1, 1>(st, tmA, tmO, a, grid, stream);
}

}  // namespace

static int tc_big_supported(const yq_conv_layer *l);

int yq_tc_supported(const yq_conv_layer *l) { return tc_big_supported(l) || yq_tc_small_supported(l); }

static int tc_big_supported(const yq_conv_layer *l)
{
    if (!l->int_form || !l->fused_mult) return 0;   // the integer-form epilogue needs M0 * 2^-31 / 2^-s parameters
    if (l->stride != 1 && l->stride != 2) return 0;   // stride 2: the activation box walks the input with elementStrides = 2
    if (!(l->size == 1 || l->size == 3) || l->pad != l->size / 2) return 0;
    if (l->cs_in % 64) return 0;
    if (l->cs_out < 32) return 0;
    return get_encode() != nullptr;
}

int yq_tc_prepare(yq_conv_layer *l)
{
    if (!tc_big_supported(l)) {
        if (yq_tc_small_prepare(l, &l->tc_small)) return -1;
        l->tc = l->tc_small;   // non-null marks "a tcgen05 flavour is ready"
        return 0;
    }
    TcState *st = new TcState();
    st->BN = l->cs_out >= 128 ? 128 : (l->cs_out >= 64 ? 64 : 32);   // TMA-store box inner extent == BN <= cs_out
    st->KC = (l->cs_in % 128) ? 64 : 128;
    st->n_pad = yq::round_up(l->n, st->BN);
    const int taps = l->size * l->size;
    const size_t ktot = (size_t)taps * l->cs_in;
    // [n_pad][taps][cs_in] rows: the same image (and arena tag) as the flat kernels'
    std::vector<uint8_t> wp;
    std::vector<int32_t> corr((size_t)taps * st->n_pad, 0);
    char tag[24];
    snprintf(tag, sizeof tag, "ohwi.%d", st->n_pad);
    const bool on_dev = yq::pack_fetch_device(l, tag, (size_t)st->n_pad * ktot, (void **)&st->w);
    const bool cached = on_dev || (yq::pack_fetch(l, tag, wp) && wp.size() == (size_t)st->n_pad * ktot);
    if (!cached) wp.assign((size_t)st->n_pad * ktot, 0);
    for (int oc = 0; oc < l->n; ++oc)
        for (int t = 0; t < taps; ++t) {
            int tsum = 0;
            for (int ci = 0; ci < l->c; ++ci) {
                const uint8_t w = l->host_w[((size_t)oc * l->c + ci) * taps + t];
                if (!cached) wp[(size_t)oc * ktot + (size_t)t * l->cs_in + ci] = w;
                tsum += w;
            }
            corr[(size_t)t * st->n_pad + oc] = l->zp_in * (tsum - (int)l->host_zw[oc] * l->c);
        }
    if (!cached) yq::pack_put(l, tag, wp);
    auto cleanup = [&]() {
        cudaFree(st->w);
        cudaFree(st->corr);
        delete st;
        return -1;
    };
    if (!on_dev) {
        if (cudaMalloc((void **)&st->w, wp.size()) != cudaSuccess) return cleanup();
        if (cudaMemcpy(st->w, wp.data(), wp.size(), cudaMemcpyHostToDevice) != cudaSuccess) return cleanup();
    }
    if (l->zp_in != 0 && l->size > 1) {
        if (cudaMalloc((void **)&st->corr, corr.size() * 4) != cudaSuccess) return cleanup();
        if (cudaMemcpy(st->corr, corr.data(), corr.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return cleanup();
    }
    // weights: 2-D [n_pad][ktot], box KC x BN
    EncodeTiledFn enc = get_encode();
    cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)st->n_pad};
    cuuint64_t strides[1] = {(cuuint64_t)ktot};
    cuuint32_t box[2] = {(cuuint32_t)st->KC, (cuuint32_t)st->BN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&st->tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, st->w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle_for(st->KC), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        yq::fail("cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
        return cleanup();
    }
    l->tc = st;
    return 0;
}

void yq_tc_free(yq_conv_layer *l)
{
    if (l->tc_small) {
        yq_tc_small_free(l->tc_small);
        l->tc_small = nullptr;
        l->tc = nullptr;
        return;
    }
    TcState *st = (TcState *)l->tc;
    if (!st) return;
    cudaFree(st->w);
    cudaFree(st->corr);
    delete st;
    l->tc = nullptr;
}

int yq_tc_can_fuse_pool(const yq_conv_layer *l) { return l->kernel == 1 && l->tc_small != nullptr; }

int yq_tc_geom_supported(const yq_conv_layer *l) { return l->kernel == 1 && l->tc && !l->tc_small; }
int yq_tc_out_geom_supported(const yq_conv_layer *l) { return l->kernel == 1 && l->tc_small != nullptr && !l->quant_stop_flag; }

int yq_tc_cluster_enabled()
{
    const char *e = getenv("YQ_TC_CLUSTER");
    return e ? atoi(e) != 0 : 0;
}

int yq_tc_forward(yq_conv_layer *l, const uint8_t *in_u8, uint8_t *out_u8, uint8_t *out_pool, float *out_f32, int32_t *out_acc, int batch,
                  cudaStream_t stream, const yq_act_geom *in_geom, int in_halo_fill, const yq_act_geom *out_geom)
{
    const bool plain_in = !in_geom || (in_geom->pad == 0 && in_geom->pitch_w == l->w && in_geom->rows_h == l->h);
    const bool plain_out = !out_geom || (out_geom->pad == 0 && out_geom->pitch_w == l->out_w && out_geom->rows_h == l->out_h);
    if (l->tc_small) {
        if (!plain_in) return yq::fail("the small-c tcgen05 flavour reads plain tensors only");
        if (!plain_out && (out_pool || out_acc || l->quant_stop_flag)) return yq::fail("the small-c tcgen05 flavour: a halo-padded output comes without side outputs");
        return yq_tc_small_forward(l, l->tc_small, in_u8, out_u8, out_pool, out_f32, out_acc, batch, stream, plain_out ? nullptr : out_geom);
    }
    if (out_pool || !out_u8) return yq::fail("the TMA tcgen05 flavour has no fused max-pool output");
    TcState *st = (TcState *)l->tc;
    if (!st) return yq::fail("tcgen05 flavour was not prepared for this layer");
    int TW = 0, TH = 0, TN = 0;
    choose_tile(batch, l->out_h, l->out_w, &TW, &TH, &TN);
    // padding comes out of the input's halo when that is wide enough and known to hold zp_in (im2col.c:5-14 pads with zp_in);
    // otherwise the TMA unit zero-fills outside the image and the epilogue restores zp_in * sum(w - zp_w) per border tap
    const int halo = (in_geom && in_geom->pad >= l->pad && in_halo_fill == l->zp_in) ? l->pad : 0;
    const yq_act_geom gi = in_geom ? *in_geom : yq_act_geom{0, l->w, l->h}, go = out_geom ? *out_geom : yq_act_geom{0, l->out_w, l->out_h};
    TcState::Key key{in_u8, out_u8, batch, gi, go, halo};
    auto it = st->maps.find(key);
    if (it == st->maps.end()) {
        if (st->maps.size() > 64) st->maps.clear();
        CUtensorMap tmA, tmO;
        if (encode_nhwc(&tmA, in_u8, &gi, halo, batch, l->h, l->w, l->cs_in, st->KC, TW, TH, TN, l->stride)) return -1;
        if (encode_nhwc(&tmO, out_u8, &go, 0, batch, l->out_h, l->out_w, l->cs_out, st->BN, TW, TH, TN, 1)) return -1;
        it = st->maps.emplace(key, std::make_pair(tmA, tmO)).first;
    }
    TcArgs a;
    memset(&a, 0, sizeof a);
    a.ep = yq::make_epi(l);
    a.out_f32 = l->quant_stop_flag ? out_f32 : nullptr;
    a.out_acc = out_acc;
    a.corr = halo ? nullptr : st->corr;
    a.stride = l->stride;
    a.cpad = halo;
    a.B = batch; a.OH = l->out_h; a.OW = l->out_w; a.N = l->n; a.CSO = l->cs_out; a.n_pad = st->n_pad;
    a.TW = TW; a.TH = TH; a.TN = TN;
    a.tiles_x = (l->out_w + TW - 1) / TW;
    a.tiles_y = (l->out_h + TH - 1) / TH;
    const int tiles_b = (batch + TN - 1) / TN;
    a.size = l->size; a.pad = l->pad; a.cpt = l->cs_in / st->KC; a.CS = l->cs_in;
    a.H = l->h; a.W = l->w;
    dim3 grid(a.tiles_x * a.tiles_y * tiles_b, st->n_pad / st->BN);
    const CUtensorMap &tmA = it->second.first, &tmO = it->second.second;
#define YQ_TC(BN_, KC_) return launch<BN_, KC_>(l, st, tmA, tmO, a, grid, stream)
    if (st->KC == 128) {
        if (st->BN == 128) YQ_TC(128, 128);
        if (st->BN == 64) YQ_TC(64, 128);
        YQ_TC(32, 128);
    } else {
        if (st->BN == 128) YQ_TC(128, 64);
        if (st->BN == 64) YQ_TC(64, 64);
        YQ_TC(32, 64);
    }
#undef YQ_TC
}
