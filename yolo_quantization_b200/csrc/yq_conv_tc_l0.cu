// yq_conv_tc_l0.cu -- the network's FIRST convolution (c = 3, 3x3 / stride 1 / pad 1, RELU6, fused 2x2/2 max-pool) read
// straight from the CHW planes the reference keeps its input in (net.input_uint8, network.c:248-250), as dense tcgen05
// kind::i8 MMAs.  Restates convolutional_layer.c:694-751 + maxpool_layer.c:109-153 for this one layer class; the generic
// rows flavour (yq_conv_tc_rows.cu) stays behind it for every shape / weight set this file does not take.
//
// Why a second kernel for one layer: the rows flavour's layer-0 form spends 3 (x 2 signed blocks) MMAs of N = 64 on a
// 16 x 32 pixel tile -- K = 32 bytes per MMA row of which 12 carry weights -- and 8192 accumulators per tile leave every
// epilogue thread with 8 pooled outputs per ~100 instructions of per-tile bookkeeping (tile split, barrier wait, addresses):
// profiles/r1_forward_full.csv shows it issue-bound (56 M warp instructions, IPC 2.2), not HBM-bound.  Here
//
//   * one MMA row = EIGHT output pixels of one image row.  Its K = 32 bytes are the three colour runs of the 10 input
//     pixels under them, [R x-1..x+8][G x-1..x+8][B x-1..x+8][0 0] -- 30 of 32 bytes carry weights -- and the filter bank
//     is the Toeplitz matrix [8 pixels x n channels][3 colours x 10 pixels] of one filter row ky.  A tile is 16 rows x 64
//     pixels = 128 MMA rows (row m = 8 * image row + segment), N = 8 n = 128 columns, 3 MMAs (one per ky, the operand start
//     shifted by one image row), x 2 for the two signed weight blocks h + l = w - zp_w: 384 tensor clocks per 16384
//     accumulators (43 per clock and SM; the rows form: 28);
//   * 16384 accumulators per tile: an epilogue thread requantises 32 (SPLIT = 2: 16) pooled outputs per tile, so the
//     per-tile bookkeeping weighs a quarter (half) of what it did;
//   * the zero point of the output rides in the multiply (IMAD.HI's addend: (umulhi(x, 2 M0) + (zp_out << s)) >> s) and the
//     |x| < xlim test is dropped when the layer's weights cannot reach the limit (host-side bound).
//
// Pipeline (as the rows flavour's planar form): two producer warps -- one per TMEM accumulator, tiles alternate between them --
// each: TMA box {96 bytes, 18 rows, 3 planes} at (x0 - 16, y0 - 1) into a staging slot (zero fill outside the image: zp_in = 0),
// all 32 lanes rearrange it into the operand tile, one lane issues the 6 MMAs + commit.  4 * SPLIT epilogue warps: TMEM lane
// quarter = 4 conv rows = 2 pooled rows; the 16x256b load shape hands thread (qi, qq) the complete 2x2 windows of segment qi,
// channels 4 qq .. 4 qq + 3: max in the accumulator domain, requantise the winner (exact: the map is monotone), per-pixel FP64
// redo when a byte would wrap (the reference's uint8 store wraps BEFORE its pool) or |x| >= xlim.
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include <map>
#include <utility>
#include <vector>

#include "yq_common.h"
#include "yq_epilogue.cuh"
#include "yq_tc_ptx.cuh"

using namespace yqtc;

namespace {

constexpr int L0_ROWS = 16;             // conv output rows per tile
constexpr int L0_SEGS = 8;              // 8-pixel segments per tile row
constexpr int L0_TWPX = 64;             // conv output pixels per tile row
constexpr int L0_AROWS = L0_ROWS + 2;   // input rows of a tile
constexpr int L0_SCOLS = 96;            // staged bytes per plane row: x0 - 16 .. x0 + 79 (the tile reads x0 - 1 .. x0 + 64 = bytes 15 .. 80)
constexpr int L0_AROWB = 256;           // operand bytes per input row: [chunk 0: 8 segments x 16 B][chunk 1: 8 segments x 16 B]

// Forms of the kernel (GROUPS): 1 = one epilogue group (4 warps), 2 = two groups, group g owns accumulator g, 4 = two warp sets splitting
// every tile by pooled row ("window halves"); all three: two producer warps, two accumulators (256 TMEM columns), two CTAs per SM.
// 3 = THREE groups over FOUR accumulators with four producer warps, one CTA per SM: tile `it` goes to accumulator it % 4 and group
// it % 3, so a group that finishes a tile finds its next accumulator already filled (with as many groups as accumulators -- form 2 --
// group and producer wait for each other a quarter of their time)
template <int GROUPS>
struct L0Mode {
    static constexpr int NGRP = GROUPS == 3 ? 3 : (GROUPS == 1 ? 1 : 2);      // epilogue warp sets of four warps
    static constexpr int NPROD = GROUPS == 3 ? 4 : 2;                          // producer warps = accumulators
    static constexpr int NACC = NPROD;
    static constexpr int MINB = GROUPS == 3 ? 1 : 2;                           // CTAs per SM
    static constexpr int NT = 128 * NGRP;                                      // epilogue threads; the producer warps come after them
    static constexpr int THREADS = NT + 32 * NPROD;
};

template <int NCH, int GROUPS>
struct L0Cfg {
    using M = L0Mode<GROUPS>;
    static constexpr int N = 8 * NCH;                         // MMA N = TMEM columns of one accumulator
    static_assert(N % 16 == 0 && N <= 256, "kind::i8 N");
    static constexpr int TMEM_COLS = M::NACC * N;
    static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "power of two");
    static constexpr int NMMA = 6;                            // 2 signed blocks x 3 filter rows
    static constexpr int BSUB = N * 32;
    static constexpr int B_BYTES = NMMA * BSUB;
    static constexpr int A_STRIDE = L0_AROWS * L0_AROWB;      // 4608
    static constexpr int NA = 2 * M::NPROD;                   // operand tiles (two per producer warp)
    static constexpr int NBUF = 2 * M::NPROD;                 // staging slots (two per producer warp)
    static constexpr int S_BYTES = 3 * L0_AROWS * L0_SCOLS;   // 5184: what TMA delivers per tile
    static constexpr int S_STRIDE = (S_BYTES + 127) / 128 * 128;
    static constexpr int A_OFF = 0;
    static constexpr int S_OFF = NA * A_STRIDE;
    static constexpr int B_OFF = S_OFF + NBUF * S_STRIDE;
    static constexpr int ST_OFF = B_OFF + B_BYTES;            // BULK: epilogue warps x 2 buffers x [2 pooled rows][32 pixels][16 channels]
    static constexpr int BAR_OFF = ST_OFF + 4 * M::NGRP * 2 * 1024;
    static constexpr int TOTAL = BAR_OFF + 256;
    static_assert(S_OFF % 128 == 0 && B_OFF % 128 == 0 && BAR_OFF % 8 == 0, "alignment");
};

struct L0Args {
    uint8_t *out_pool;      // pooled output: pixel (n, y, x) at out + ((n*OHP + y + opad)*OWP + x + opad)*out_cs
    const uint8_t *wimg;    // shared-memory image of the six filter tiles
    int PH, PW, OHP, OWP, opad, out_cs;
    int tiles_x, tiles_y, num_tiles, zp_out;
    uint32_t magic_x, magic_y;   // ceil(2^32 / tiles_x), ceil(2^32 / tiles_y)
    uint32_t xlim;
    int tstore;             // BULK: 1 = the staged rows leave through a 4-D tensor store (see the kernel), 0 = linear bulk copies
    int knobs;              // -DYQ_L0_KNOBS experiments (results are garbage): 1 no MMAs, 2 no epilogue arithmetic, 4 no rearrangement, 8 no stores
    int4 cq[32];            // {bias, zw, 2*M0, shift} per channel
    double mc[32];          // M_value * 2^-s (FP64 redo)
};

// -DYQ_L0_TRACE: lane 0 of producer warp 0 and of epilogue warp 0 add up the clocks of their phases per CTA (yq_l0_trace[cta * 16 + ..]):
// producer 0 wait full, 1 rearrange, 2 refill + wait acc_empty, 3 MMA issue, 4 tiles, 5 loop clocks, 6 loop ns;
// epilogue 8 wait acc_full, 9 tile body, 10 tiles, 11 loop clocks; 12 ns kernel entry -> loop start, 13 ns loop start (absolute), 14 ns at exit
#ifdef YQ_L0_TRACE
__device__ unsigned long long yq_l0_trace[1024 * 16];
__device__ __forceinline__ unsigned long long l0_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define L0TR_MARK(ev) do { const long long now_ = clock64(); tr_acc[ev] += (unsigned long long)(now_ - tr_prev); tr_prev = now_; } while (0)
#else
#define L0TR_MARK(ev)
#endif

#ifdef YQ_L0_KNOBS
#define L0KNOB(bit) (a.knobs & (bit))
#else
#define L0KNOB(bit) 0
#endif

__device__ __forceinline__ uint64_t l0_desc_ns(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

// both 16-lane halves of this warp's TMEM lane quarter, 32 columns each: thread (qi, qq) gets lanes {qi, qi + 8} in v0 and
// {qi + 16, qi + 24} in v1; register 4 g + {0, 1} = columns 8 g + 2 qq + {0, 1} of the first lane, + {2, 3} of the second
__device__ __forceinline__ void l0_ld_issue(uint32_t taddr, uint32_t (&v0)[16], uint32_t (&v1)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%32];\n\t"
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%33];"
        : "=r"(v0[0]), "=r"(v0[1]), "=r"(v0[2]), "=r"(v0[3]), "=r"(v0[4]), "=r"(v0[5]), "=r"(v0[6]), "=r"(v0[7]), "=r"(v0[8]), "=r"(v0[9]),
          "=r"(v0[10]), "=r"(v0[11]), "=r"(v0[12]), "=r"(v0[13]), "=r"(v0[14]), "=r"(v0[15]), "=r"(v1[0]), "=r"(v1[1]), "=r"(v1[2]),
          "=r"(v1[3]), "=r"(v1[4]), "=r"(v1[5]), "=r"(v1[6]), "=r"(v1[7]), "=r"(v1[8]), "=r"(v1[9]), "=r"(v1[10]), "=r"(v1[11]),
          "=r"(v1[12]), "=r"(v1[13]), "=r"(v1[14]), "=r"(v1[15])
        : "r"(taddr), "r"(taddr + (16u << 16)));
}
__device__ __forceinline__ void l0_ld_wait(uint32_t (&v0)[16], uint32_t (&v1)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v0[0]), "+r"(v0[1]), "+r"(v0[2]), "+r"(v0[3]), "+r"(v0[4]), "+r"(v0[5]), "+r"(v0[6]), "+r"(v0[7]), "+r"(v0[8]), "+r"(v0[9]),
                   "+r"(v0[10]), "+r"(v0[11]), "+r"(v0[12]), "+r"(v0[13]), "+r"(v0[14]), "+r"(v0[15]), "+r"(v1[0]), "+r"(v1[1]), "+r"(v1[2]),
                   "+r"(v1[3]), "+r"(v1[4]), "+r"(v1[5]), "+r"(v1[6]), "+r"(v1[7]), "+r"(v1[8]), "+r"(v1[9]), "+r"(v1[10]), "+r"(v1[11]),
                   "+r"(v1[12]), "+r"(v1[13]), "+r"(v1[14]), "+r"(v1[15])::"memory");
}

// one 16-lane half only (GROUPS = 4: a warp requantises ONE pooled row of its lane quarter)
__device__ __forceinline__ void l0_ld1_issue(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void l0_ld1_wait(uint32_t (&v)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]),
                   "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])::"memory");
}

// four bytes -> one little-endian word on the multiplier pipe (the alu pipe is this kernel's bottleneck: PRMT lives there)
__device__ __forceinline__ uint32_t l0_pack(const int (&r)[4])
{
    uint32_t lo, hi, w;
    asm("mad.lo.u32 %0, %1, 256, %2;" : "=r"(lo) : "r"(r[1]), "r"(r[0]));
    asm("mad.lo.u32 %0, %1, 256, %2;" : "=r"(hi) : "r"(r[3]), "r"(r[2]));
    asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(w) : "r"(hi), "r"(lo));
    return w;
}

// the reference's per-pixel arithmetic for one window (FP64 multiply, uint8 wrap, then the pool); the accumulators are the
// zero-point-corrected ones (two signed weight blocks)
__device__ __forceinline__ int l0_window_slow(int bias, double mcd, int zo, int v0, int v1, int v2, int v3)
{
    const int xs[4] = {v0 + bias, v1 + bias, v2 + bias, v3 + bias};
    int best = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int q = max(__double2int_rz(__dmul_rn((double)xs[p], mcd)), 0);
        best = max(best, (q + zo) & 255);
    }
    return best;
}

// GROUPS = 2: two epilogue groups of four warps, group g owns accumulator g and every second tile (all four pixel pairs of its
// threads' segments: 32 pooled outputs per thread and tile, like GROUPS = 1, with twice the warps to hide latencies behind)
// BULK: a warp stages its two pooled rows (2 x 512 B) in shared memory and one lane writes them out.  In pixel order (two linear bulk
// copies per warp) the eight segments' words of one pixel pair sit 64 bytes apart -- a 4-way bank conflict on every staging store, and
// this kernel's shared-memory data pipe is its bottleneck (ncu r2: 41 % LSU + 40 % tensor-core operand wavefronts).  `tstore` stages
// [row][pair][segment][16 channels] instead -- 32 lanes, 32 banks -- and a 4-D tensor map over the pooled tensor viewed as
// [rows][4 pixels of a segment][segments][16 channels] (strides W*16, 16, 64 bytes) lets ONE TMA store per row put the pixels back in order
// (cp.async.bulk shared -> global) -- whole 512-byte runs instead of 16 predicated 4-byte stores per thread
template <int NCH, int GROUPS, bool CHECKX, bool BULK, bool BIASF>
__global__ void __launch_bounds__(L0Mode<GROUPS>::THREADS, L0Mode<GROUPS>::MINB) conv_u8_tc_l0_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmO, const __grid_constant__ L0Args a)
{
    using L = L0Cfg<NCH, GROUPS>;
    using MD = L0Mode<GROUPS>;
    static_assert(NCH == 16, "one 32-column chunk per pixel pair");
    constexpr int NT = MD::NT, NPROD = MD::NPROD, NACC = MD::NACC, NGRP = MD::NGRP;
    constexpr int NBUF = L::NBUF;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint64_t *full = (uint64_t *)(smem + L::BAR_OFF);   // [NBUF] staged planes landed
    uint64_t *acc_full = full + NBUF;                   // [NACC] accumulator complete
    uint64_t *acc_empty = acc_full + NACC;              // [NACC] accumulator read out
    uint64_t *b_full = acc_empty + NACC;                // the resident filter tiles have landed
    uint32_t *tmem_slot = (uint32_t *)(b_full + 1);
    volatile int *epi_seq = (volatile int *)(tmem_slot + 1);     // GROUPS = 3: how many tiles have entered their epilogue (see below)

#ifdef YQ_L0_TRACE
    const unsigned long long tr_entry = l0_ns();
#endif
    const int t = threadIdx.x, warp = (t >> 5) & 3, grp = t >> 7, lane = t & 31;      // (grp: meaningful for epilogue threads)
    const bool producer = t >= NT;
    const int qi = lane >> 2, qq = lane & 3;

    if (t == 0) {
        for (int b = 0; b < NBUF; ++b) mbar_init(&full[b], 1);
        for (int b = 0; b < NACC; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], GROUPS == 4 ? 8 : 4);      // the warps that read an accumulator
        }
        mbar_init(b_full, 1);
        *epi_seq = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(b_full, (uint32_t)L::B_BYTES);
        bulk_load(smem + L::B_OFF, a.wimg, (uint32_t)L::B_BYTES, b_full);      // constants: before the wait on the previous kernel
    }
    if (t < 32) tmem_alloc<L::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    struct TileXY { int tx, ty, n; };
    auto split_tile = [&](int tile) -> TileXY {
        const int r1 = a.tiles_x == 1 ? tile : (int)__umulhi((uint32_t)tile, a.magic_x);
        const int n = a.tiles_y == 1 ? r1 : (int)__umulhi((uint32_t)r1, a.magic_y);
        return TileXY{tile - r1 * a.tiles_x, r1 - n * a.tiles_y, n};
    };
    const int first = blockIdx.x, step = gridDim.x;

    yq_pdl_wait_then_release();      // everything above touched only constants and on-chip state

    if (producer) {
        const uint32_t sA = smem_u32(smem + L::A_OFF), b0 = smem_u32(smem + L::B_OFF);
        const int pw = (t - NT) >> 5;                   // producer warp pw: tiles it = pw, pw + NPROD, ... into accumulator pw
        constexpr int SPP = NBUF / NPROD, APP = L::NA / NPROD;  // staging slots / operand tiles per producer warp
        auto load_planes = [&](int tile, int sbuf) {
            const TileXY p = split_tile(tile);
            mbar_expect_tx(&full[sbuf], (uint32_t)L::S_BYTES);
            tma_load_3d(smem + L::S_OFF + sbuf * L::S_STRIDE, &tmA, &full[sbuf], p.tx * L0_TWPX - 16, p.ty * L0_ROWS - 1, p.n * 3);
        };
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
#pragma unroll 1
            for (int d = 0; d < SPP; ++d)
                if (first + (pw + NPROD * d) * step < a.num_tiles) load_planes(first + (pw + NPROD * d) * step, pw * SPP + d);
        }
        __syncwarp();
        constexpr int PLW = L0_AROWS * L0_SCOLS / 4;    // words of one staged plane
        constexpr int NUNITS = L0_AROWS * L0_SEGS, NG = (NUNITS + 31) / 32;     // (input row, segment) units: 144 -> 5 rounds
        int src_w[NG], dst_off[NG];
#pragma unroll
        for (int i = 0; i < NG; ++i) {
            const int g = lane + 32 * i, r = g >> 3, s = g & 7;
            src_w[i] = (r * L0_SCOLS + 12 + 8 * s) / 4;         // word holding byte 12 + 8 s of the row: the unit's pixels are bytes 15 + 8 s .. 24 + 8 s
            dst_off[i] = r * L0_AROWB + 16 * s;
        }
#ifdef YQ_L0_TRACE
        unsigned long long tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long tr_prev = clock64();
        const long long tr_c0 = tr_prev;
        const unsigned long long tr_n0 = l0_ns();
#endif
        int k = 0;
#pragma unroll 1
        for (int tile = first + pw * step; tile < a.num_tiles; tile += NPROD * step, ++k) {
            const int sbuf = pw * SPP + k % SPP, abuf = pw * APP + k % APP;
            mbar_wait(&full[sbuf], (uint32_t)((k / SPP) & 1));
            L0TR_MARK(0);
            const uint32_t *S = reinterpret_cast<const uint32_t *>(smem + L::S_OFF + sbuf * L::S_STRIDE);
            uint8_t *A = smem + L::A_OFF + abuf * L::A_STRIDE;
            if (!L0KNOB(4)) {
#pragma unroll
                for (int i = 0; i < NG; ++i) {
                    if (i + 1 < NG || lane + 32 * i < NUNITS) {
                        const uint32_t *w = S + src_w[i];
                        uint32_t x[3][4];
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (int j = 0; j < 4; ++j) x[c][j] = w[c * PLW + j];
                        // colour run of 10 pixels = bytes 3 .. 12 of the four words
                        const uint32_t r03 = __funnelshift_r(x[0][0], x[0][1], 24), r47 = __funnelshift_r(x[0][1], x[0][2], 24), r8 = __funnelshift_r(x[0][2], x[0][3], 24);
                        const uint32_t g03 = __funnelshift_r(x[1][0], x[1][1], 24), g47 = __funnelshift_r(x[1][1], x[1][2], 24), g8 = __funnelshift_r(x[1][2], x[1][3], 24);
                        const uint32_t b03 = __funnelshift_r(x[2][0], x[2][1], 24), b47 = __funnelshift_r(x[2][1], x[2][2], 24), b8 = __funnelshift_r(x[2][2], x[2][3], 24);
                        // K bytes 0..15: R0-9 G0-5; 16..31: G6-9 B0-9 p q  (p, q = 0, 0 -- or 255, 1 when the bias rides in the filter bank, see prepare)
                        *reinterpret_cast<uint4 *>(A + dst_off[i]) = make_uint4(r03, r47, __byte_perm(r8, g03, 0x5410), __funnelshift_r(g03, g47, 16));
                        *reinterpret_cast<uint4 *>(A + dst_off[i] + 128) = make_uint4(__funnelshift_r(g47, g8, 16), b03, b47, (b8 & 0xffffu) | (BIASF ? 0x01ff0000u : 0u));
                    }
                }
            }
            fence_proxy_async();      // the operand tile (generic-proxy stores) -> visible to the tensor core
            __syncwarp();
            L0TR_MARK(1);
            if (lane == 0) {
                if (k == 0) mbar_wait(b_full, 0);
                if (tile + NBUF * step < a.num_tiles) load_planes(tile + NBUF * step, sbuf);      // every lane has read the slot (SPP * NPROD = NBUF tiles on)
                if (k >= 1) mbar_wait(&acc_empty[pw], ((uint32_t)k & 1u) ^ 1u);                    // the accumulator's previous tile has been read out
                tc_fence_after();
                L0TR_MARK(2);
                if (L0KNOB(1)) {
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_full[pw])) : "memory");
                } else {
                    constexpr uint32_t idesc = make_idesc(L::N) | (1u << 10);       // B (the filters) is SINT8
                    const uint32_t a0 = sA + abuf * L::A_STRIDE;
                    const uint32_t tacc = tmem_base + (uint32_t)(pw * L::N);
#pragma unroll
                    for (int part = 0; part < 2; ++part)
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
                            umma_i8(tacc, l0_desc_ns(a0 + ky * L0_AROWB, 128, L0_AROWB), l0_desc_ns(b0 + (part * 3 + ky) * L::BSUB, 128, 256), idesc,
                                    (part | ky) ? 1u : 0u);
                    umma_commit(&acc_full[pw]);
                }
                L0TR_MARK(3);
            }
            __syncwarp();             // (the operand tile rewritten next had its MMAs two of this warp's tiles back: complete, see the acc_empty wait)
        }
#ifdef YQ_L0_TRACE
        if (pw == 0 && lane == 0 && blockIdx.x < 1024) {
            unsigned long long *q = yq_l0_trace + blockIdx.x * 16;
            for (int e = 0; e < 4; ++e) q[e] = tr_acc[e];
            q[4] = (unsigned long long)k; q[5] = (unsigned long long)(clock64() - tr_c0); q[6] = l0_ns() - tr_n0; q[12] = tr_n0 - tr_entry; q[13] = tr_n0;
        }
#endif
    } else {
        // ---- epilogue: warp = TMEM lane quarter = conv rows 4w .. 4w+3 = pooled rows 2w, 2w+1; thread (qi, qq) = segment qi, channels 4qq .. 4qq+3
        int bias[4], sh[4];
        uint32_t m2[4], zsh[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int4 c = a.cq[4 * qq + g];
            bias[g] = BIASF ? 0 : c.x; m2[g] = (uint32_t)c.z; sh[g] = c.w; zsh[g] = (uint32_t)a.zp_out << c.w;
        }
        const int zo = a.zp_out;
        if constexpr (GROUPS == 4) {
            // ---- window halves: warp (w, hv) requantises pooled row 2w + hv of EVERY tile (lanes 16 hv .. 16 hv + 15 of quarter w): both
            // accumulators alternate, each drained by all eight warps in half the time, so one fills while the other drains
            static_assert(BULK, "window halves store by bulk copy");
            const int hv = grp;
            const uint32_t tq = tmem_base + ((uint32_t)(warp * 32 + hv * 16) << 16);
            uint8_t *const stage = smem + L::ST_OFF + (t >> 5) * 1024;      // two buffers of [32 pixels][16 channels]
            const uint32_t st_thr = (uint32_t)(4 * qi * 16 + 4 * qq);
            const uint32_t zo4 = (uint32_t)zo * 0x01010101u, qlim = 255u - (uint32_t)zo;
            const uint32_t out_row = (uint32_t)(a.OWP * a.out_cs);
#ifdef YQ_L0_TRACE
            unsigned long long tr_acc[2] = {0, 0};
            long long tr_prev = clock64();
            const long long tr_c0 = tr_prev;
#endif
            int j = 0;
#pragma unroll 1
            for (int tile = first; tile < a.num_tiles; tile += step, ++j) {
                const int acc = j & 1;
                mbar_wait(&acc_full[acc], (uint32_t)((j >> 1) & 1));
                L0TR_MARK(0);
                tc_fence_after();
                const uint32_t tqa = tq + (uint32_t)(acc * L::N);
                uint8_t *const st = stage + (j & 1) * 512;
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");      // the copy that read this buffer two tiles ago
                __syncwarp();
                uint32_t V[2][16];
                l0_ld1_issue(tqa, V[0]);
                l0_ld1_wait(V[0]);
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    if (pp + 1 < 4) l0_ld1_issue(tqa + 32 * (pp + 1), V[(pp + 1) & 1]);
                    uint32_t(&v)[16] = V[pp & 1];
                    int r[4];
                    uint32_t orx = 0, orr = 0;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint32_t x;
                        if (BIASF) x = (uint32_t)max(max(max((int)v[4 * g], (int)v[4 * g + 1]), (int)v[4 * g + 2]), max((int)v[4 * g + 3], 0));
                        else x = (uint32_t)max(max(max((int)v[4 * g], (int)v[4 * g + 1]), max((int)v[4 * g + 2], (int)v[4 * g + 3])) + bias[g], 0);
                        r[g] = (int)(__umulhi(x, m2[g]) >> sh[g]);
                        if (CHECKX) orx |= x;
                        orr |= (uint32_t)r[g];
                    }
                    uint32_t w;
                    // (the OR of the values bounds each of them: a conservative test, exact for zp_out = 0)
                    if (orr > qlim || (CHECKX && orx >= a.xlim)) {
#pragma unroll
                        for (int g = 0; g < 4; ++g)
                            r[g] = l0_window_slow(bias[g], a.mc[4 * qq + g], zo, (int)v[4 * g], (int)v[4 * g + 1], (int)v[4 * g + 2], (int)v[4 * g + 3]);
                        w = l0_pack(r);
                    } else {
                        w = l0_pack(r) + zo4;      // no byte carries: every value is <= 255 - zp_out
                    }
                    *reinterpret_cast<uint32_t *>(st + st_thr + pp * 16) = w;
                    if (pp + 1 < 4) {
                        l0_ld1_wait(V[(pp + 1) & 1]);
                        if (pp + 2 == 4) {       // this warp's TMEM reads of the tile are complete
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[acc])) : "memory");
                        }
                    }
                }
                fence_proxy_async();      // the staged row (generic-proxy stores) -> visible to the bulk copy engine
                __syncwarp();
                if (lane == 0) {
                    const TileXY cur = split_tile(tile);
                    const int py = cur.ty * (L0_ROWS / 2) + 2 * warp + hv;
                    const int npx = min(L0_TWPX / 2, a.PW - cur.tx * (L0_TWPX / 2));      // pooled pixels of this tile inside the image
                    if (py < a.PH && !L0KNOB(8)) {
                        uint8_t *g0 = a.out_pool + (size_t)((uint32_t)((cur.n * a.OHP + py + a.opad) * a.OWP + cur.tx * (L0_TWPX / 2) + a.opad) * (uint32_t)a.out_cs);
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g0), "r"(smem_u32(st)), "r"((uint32_t)(npx * 16)) : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                L0TR_MARK(1);
            }
            (void)out_row;
#ifdef YQ_L0_TRACE
            if (t == 0 && blockIdx.x < 1024) {
                unsigned long long *q = yq_l0_trace + blockIdx.x * 16;
                q[8] = tr_acc[0]; q[9] = tr_acc[1]; q[10] = (unsigned long long)j; q[11] = (unsigned long long)(clock64() - tr_c0);
            }
#endif
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        } else {
        const uint32_t tq = tmem_base + ((uint32_t)(warp * 32) << 16);
        const uint32_t out_thr = (uint32_t)(((2 * warp + a.opad) * a.OWP + 4 * qi + a.opad) * a.out_cs + 4 * qq);
        const uint32_t out_row = (uint32_t)(a.OWP * a.out_cs);
        // BULK: this warp's staging, two buffers of [2 pooled rows][32 pixels][16 channels]
        uint8_t *const stage = smem + L::ST_OFF + ((t >> 5) * 2) * 1024;
        const uint32_t st_thr = (uint32_t)((a.tstore ? qi * 16 : 4 * qi * 16) + 4 * qq), st_pp = a.tstore ? 128u : 16u;
#ifdef YQ_L0_TRACE
        unsigned long long tr_acc[2] = {0, 0};
        long long tr_prev = clock64();
        const long long tr_c0 = tr_prev;
#endif
        int j = 0;
#pragma unroll 1
        for (int tile = first + grp * step; tile < a.num_tiles; tile += NGRP * step, ++j) {
            const TileXY cur = split_tile(tile);
            const int it = grp + NGRP * j;               // this CTA's it-th tile: accumulator it % NACC, its (it / NACC)-th use
            const int acc = it % NACC;
            if constexpr (NGRP != NACC && NGRP > 1) {
                // A parity wait tells "this use" from "the use before" only.  With fewer groups than accumulators a group may come to wait
                // for use u of an accumulator before its use u - 1 (drained by ANOTHER group) has even been filled -- and the phase of equal
                // parity two uses back would let it through.  So tiles enter their epilogues in order: tile it waits until tile it - 1
                // has seen its accumulator complete, which (by induction) puts every earlier use of this accumulator behind us.
                const long long t0 = clock64();
                while (*epi_seq < it) {
                    __nanosleep(20);
                    if (clock64() - t0 > 4000000000LL) __trap();
                }
            }
            mbar_wait(&acc_full[acc], (uint32_t)((it / NACC) & 1));
            if constexpr (NGRP != NACC && NGRP > 1) {
                if (warp == 0 && lane == 0) *epi_seq = it + 1;
            }
            L0TR_MARK(0);
            tc_fence_after();
            const uint32_t tqa = tq + (uint32_t)(acc * L::N);
            auto release_acc = [&]() {       // this warp's TMEM reads of the tile are complete
                tc_fence_before();
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[acc])) : "memory");
            };
            if (L0KNOB(2)) {
                release_acc();
                continue;
            }
            const int py0 = cur.ty * (L0_ROWS / 2) + 2 * warp;
            const int px0 = cur.tx * (L0_TWPX / 2) + 4 * qi;
            const uint32_t tile_off = (uint32_t)((cur.n * a.OHP + cur.ty * (L0_ROWS / 2)) * a.OWP + cur.tx * (L0_TWPX / 2)) * (uint32_t)a.out_cs;
            uint8_t *const out_tile = a.out_pool + (size_t)(tile_off + out_thr);
            const bool row0 = py0 < a.PH, row1 = py0 + 1 < a.PH;
            uint8_t *const st = stage + (j & 1) * 1024 + st_thr;
            if (BULK) {
                // the bulk copies that read this staging buffer two tiles ago must have finished reading
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
            }
            // GROUPS = 1 (registers for it): the whole accumulator first, then the release -- the next tile's MMAs run under ALL of this
            // tile's arithmetic; GROUPS = 2: chunk by chunk; GROUPS = 3 (128 registers): chunk by chunk, the next chunk's load in flight
            constexpr int NV = GROUPS == 1 ? 4 : (GROUPS == 3 ? 2 : 1);
            uint32_t V[NV][2][16];
            if (GROUPS == 1) {
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) l0_ld_issue(tqa + 32 * pp, V[pp][0], V[pp][1]);
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) l0_ld_wait(V[pp][0], V[pp][1]);
                release_acc();
            } else if (GROUPS == 3) {
                l0_ld_issue(tqa, V[0][0], V[0][1]);
                l0_ld_wait(V[0][0], V[0][1]);
            }
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
                uint32_t(&v0)[16] = V[GROUPS == 1 ? pp : (GROUPS == 3 ? pp & 1 : 0)][0];
                uint32_t(&v1)[16] = V[GROUPS == 1 ? pp : (GROUPS == 3 ? pp & 1 : 0)][1];
                if (GROUPS == 2) {
                    l0_ld_issue(tqa + 32 * pp, v0, v1);
                    l0_ld_wait(v0, v1);
                    if (pp == 3) release_acc();
                }
                if (GROUPS == 3 && pp + 1 < 4) l0_ld_issue(tqa + 32 * (pp + 1), V[NV == 2 ? (pp + 1) & 1 : 0][0], V[NV == 2 ? (pp + 1) & 1 : 0][1]);
                int r0[4], r1[4];
                uint32_t orx = 0, orr = 0;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t x0, x1;
                    if (BIASF) {      // the accumulator already holds the bias: RELU6's max with 0 joins the second three-input max
                        x0 = (uint32_t)max(max(max((int)v0[4 * g], (int)v0[4 * g + 1]), (int)v0[4 * g + 2]), max((int)v0[4 * g + 3], 0));
                        x1 = (uint32_t)max(max(max((int)v1[4 * g], (int)v1[4 * g + 1]), (int)v1[4 * g + 2]), max((int)v1[4 * g + 3], 0));
                    } else {
                        const int m0 = max(max((int)v0[4 * g], (int)v0[4 * g + 1]), max((int)v0[4 * g + 2], (int)v0[4 * g + 3]));
                        const int m1 = max(max((int)v1[4 * g], (int)v1[4 * g + 1]), max((int)v1[4 * g + 2], (int)v1[4 * g + 3]));
                        x0 = (uint32_t)max(m0 + bias[g], 0);
                        x1 = (uint32_t)max(m1 + bias[g], 0);
                    }
                    r0[g] = (int)((__umulhi(x0, m2[g]) + zsh[g]) >> sh[g]);
                    r1[g] = (int)((__umulhi(x1, m2[g]) + zsh[g]) >> sh[g]);
                    if (CHECKX) orx |= x0 | x1;
                    orr |= (uint32_t)r0[g] | (uint32_t)r1[g];
                }
                if (orr > 255u || (CHECKX && orx >= a.xlim)) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const double mcd = a.mc[4 * qq + g];
                        r0[g] = l0_window_slow(bias[g], mcd, zo, (int)v0[4 * g], (int)v0[4 * g + 1], (int)v0[4 * g + 2], (int)v0[4 * g + 3]);
                        r1[g] = l0_window_slow(bias[g], mcd, zo, (int)v1[4 * g], (int)v1[4 * g + 1], (int)v1[4 * g + 2], (int)v1[4 * g + 3]);
                    }
                }
                const uint32_t w0 = l0_pack(r0), w1 = l0_pack(r1);      // every value is a byte here (fast path: checked; redo: masked)
                if (BULK) {
                    *reinterpret_cast<uint32_t *>(st + pp * st_pp) = w0;
                    *reinterpret_cast<uint32_t *>(st + pp * st_pp + 512) = w1;
                } else if (px0 + pp < a.PW && !L0KNOB(8)) {
                    uint8_t *dst = out_tile + pp * a.out_cs;
                    if (row0) *reinterpret_cast<uint32_t *>(dst) = w0;
                    if (row1) *reinterpret_cast<uint32_t *>(dst + out_row) = w1;
                }
                if (GROUPS == 3 && pp + 1 < 4) {
                    l0_ld_wait(V[NV == 2 ? (pp + 1) & 1 : 0][0], V[NV == 2 ? (pp + 1) & 1 : 0][1]);
                    if (pp + 2 == 4) release_acc();
                }
            }
            if (BULK) {
                fence_proxy_async();      // the staged rows (generic-proxy stores) -> visible to the bulk copy engine
                __syncwarp();
                if (lane == 0) {
                    const int npx = min(L0_TWPX / 2, a.PW - cur.tx * (L0_TWPX / 2));      // pooled pixels of this tile inside the image
                    const uint32_t nbytes = (uint32_t)(npx * 16);
                    uint8_t *g0 = a.out_pool + (size_t)(tile_off + (uint32_t)(((2 * warp + a.opad) * a.OWP + a.opad) * a.out_cs));
                    const uint32_t s0 = smem_u32(stage + (j & 1) * 1024);
                    if (a.tstore) {
                        // box {16 channels, 8 segments, 4 pixels, 1 row} at (0, first segment, 0, row); segments beyond the image are clipped
                        const int row = cur.n * a.OHP + py0;
                        if (row0 && !L0KNOB(8)) tma_store_4d(&tmO, stage + (j & 1) * 1024, 0, cur.tx * L0_SEGS, 0, row);
                        if (row1 && !L0KNOB(8)) tma_store_4d(&tmO, stage + (j & 1) * 1024 + 512, 0, cur.tx * L0_SEGS, 0, row + 1);
                    } else {
                        if (row0 && !L0KNOB(8)) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g0), "r"(s0), "r"(nbytes) : "memory");
                        if (row1 && !L0KNOB(8)) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g0 + out_row), "r"(s0 + 512), "r"(nbytes) : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            L0TR_MARK(1);
        }
#ifdef YQ_L0_TRACE
        if (t == 0 && blockIdx.x < 1024) {
            unsigned long long *q = yq_l0_trace + blockIdx.x * 16;
            q[8] = tr_acc[0]; q[9] = tr_acc[1]; q[10] = (unsigned long long)j; q[11] = (unsigned long long)(clock64() - tr_c0);
        }
#endif
        if (BULK && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // the rows are in global memory before the grid may count as complete
        }
    }
    tc_fence_before();
    __syncthreads();
#ifdef YQ_L0_TRACE
    if (t == 0 && blockIdx.x < 1024) yq_l0_trace[blockIdx.x * 16 + 14] = l0_ns();
#endif
    if (t < 32) {
        tc_fence_after();
        tmem_dealloc<L::TMEM_COLS>(tmem_base);
    }
}

struct L0State {
    int NCH = 0;
    bool checkx = true;
    bool biasf = false;     // the bias rides in the filter bank (operand pad bytes 255, 1): no bias add in the epilogue
    uint32_t xlim = 1u << 22;
    uint8_t *wimg = nullptr;
    std::map<std::pair<const void *, int>, CUtensorMap> maps;      // input tensor map per (input pointer, batch)
    std::map<std::pair<const void *, int>, CUtensorMap> omaps;     // pooled-output tensor map per (output pointer, batch)
};

typedef CUresult (*L0EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                               const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// the network input as [planes][h][w bytes]; box = 3 planes x 18 rows x 96 bytes at (x0 - 16, y0 - 1, 3 n), zero outside the image
int l0_encode(CUtensorMap *m, const void *in, int w, int h, int planes)
{
    static L0EncodeFn enc = nullptr;
    if (!enc) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) enc = (L0EncodeFn)p;
    }
    if (!enc) return yq::fail("cuTensorMapEncodeTiled is not available from this driver");
    const cuuint32_t es[3] = {1, 1, 1};
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)w, (cuuint64_t)w * h};
    const cuuint32_t box[3] = {(cuuint32_t)L0_SCOLS, (cuuint32_t)L0_AROWS, 3};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return yq::fail("layer-0 flavour: cuTensorMapEncodeTiled(planar %d x %d x %d) failed: %d", w, h, planes, (int)r);
    return 0;
}

// the pooled output [rows][pitch][16 ch] viewed as [rows][4 pixels of a segment][segments][16 ch]: box {16, 8 segments, 4 pixels, 1 row}
// = one staged row in the kernel's bank-conflict-free order; `base` = pixel (0, 0) of image 0 (behind the halo)
int l0_encode_out(CUtensorMap *m, void *base, int segs, int pitch_px, long long rows)
{
    static L0EncodeFn enc = nullptr;
    if (!enc) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) enc = (L0EncodeFn)p;
    }
    if (!enc) return yq::fail("cuTensorMapEncodeTiled is not available from this driver");
    const cuuint32_t es[4] = {1, 1, 1, 1};
    const cuuint64_t dims[4] = {16, (cuuint64_t)segs, 4, (cuuint64_t)rows};
    const cuuint64_t strides[3] = {64, 16, (cuuint64_t)pitch_px * 16};
    const cuuint32_t box[4] = {16, (cuuint32_t)L0_SEGS, 4, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return yq::fail("layer-0 flavour: cuTensorMapEncodeTiled(pooled output, %d segments x %lld rows) failed: %d", segs, rows, (int)r);
    return 0;
}

// byte position of element (row n, k) inside one [N][32] filter tile: 8-row x 16-byte core matrices, the two K chunks of a
// group side by side (LBO = 128), groups 256 bytes apart (SBO = 256)
inline size_t l0_bpos(int n, int k) { return (size_t)(n / 8) * 256 + (size_t)(k / 16) * 128 + (size_t)(n % 8) * 16 + (size_t)(k % 16); }

template <int NCH, int SPLIT, bool CHECKX, bool BULK, bool BIASF>
int l0_launch(const CUtensorMap &tmA, const CUtensorMap &tmO, const L0Args &a, cudaStream_t stream)
{
    using L = L0Cfg<NCH, SPLIT>;
    using MD = L0Mode<SPLIT>;
    const int smem = L::TOTAL + 128;
    auto kern = conv_u8_tc_l0_kernel<NCH, SPLIT, CHECKX, BULK, BIASF>;
    if (yq::ensure_dynamic_smem((const void *)kern, smem)) return -1;
    const int n_sm = yq::device_sm_count();
    int ctas_per_sm = 0;
    if (!yq::memo_get((const void *)kern, &ctas_per_sm)) {
        cudaFuncAttributes fa;
        YQ_CUDA(cudaFuncGetAttributes(&fa, kern));
        // counted by hand: the occupancy API answers 1 for kernels that allocate tensor memory
        const int by_smem = yq::device_smem_per_sm() / (smem + 1024 + (int)fa.sharedSizeBytes);
        const int by_regs = 65536 / (((fa.numRegs + 7) / 8 * 8) * MD::THREADS);
        const int by_tmem = 512 / L::TMEM_COLS;
        int occ = by_smem < by_regs ? by_smem : by_regs;
        if (by_tmem < occ) occ = by_tmem;
        if (MD::MINB < occ) occ = MD::MINB;
        if (getenv("YQ_DEBUG")) fprintf(stderr, "yq: l0<%d,%d,%d,%d,%d> regs=%d by_smem=%d by_regs=%d by_tmem=%d smem=%d\n", NCH, SPLIT, (int)CHECKX, (int)BULK, (int)BIASF, fa.numRegs, by_smem, by_regs, by_tmem, smem);
        if (occ < 1 || n_sm <= 0) return yq::fail("conv_u8_tc_l0_kernel<%d> does not fit on an SM", NCH);
        ctas_per_sm = occ;
        yq::memo_put((const void *)kern, occ);
    }
    int grid = n_sm * ctas_per_sm;
    if (grid > a.num_tiles) grid = a.num_tiles;
    YQ_CUDA(yq::launch_pdl(kern, dim3(grid), dim3(MD::THREADS), smem, stream, tmA, tmO, a));
    return 0;
}

}  // namespace

// c = 3 planes read in place (zp_in = 0: TMA's zero fill is the padding), n = 16, RELU6 + fused 2x2 pool, integer-form requantisation
int yq_tc_l0_supported(const yq_conv_layer *l)
{
    if (getenv("YQ_NO_L0") && atoi(getenv("YQ_NO_L0"))) return 0;
    if (!l->int_form || !l->fused_mult || l->saturate || l->quant_stop_flag) return 0;
    if (l->size != 3 || l->pad != 1 || l->stride != 1 || yq::act_mode(l->activation) != 0) return 0;
    if ((l->h & 1) || (l->w & 1) || l->n != l->cs_out) return 0;
    return l->c == 3 && l->cs_in == 4 && l->cs_out == 16 && l->zp_in == 0 && l->w % 16 == 0 && l->w >= 64;
}

int yq_tc_l0_prepare(yq_conv_layer *l, void **state)
{
    *state = nullptr;
    auto zpw = [&](int oc) -> int { return l->host_chanq[(size_t)oc * 4 + 1]; };
    // two signed blocks h + l = w - zp_w need every difference <= 254; (h + zp_out << s) must not leave 32 bits
    for (int oc = 0; oc < l->n; ++oc) {
        if (l->host_chanq[(size_t)oc * 4 + 3] > 22 || l->zp_out < 0 || l->zp_out > 255) return 0;
        for (int i = 0; i < 27; ++i)
            if ((int)l->host_w[(size_t)oc * 27 + i] - zpw(oc) == 255) return 0;
    }
    L0State *st = new L0State();
    st->NCH = l->cs_out;
    const int NCH = st->NCH, N = 8 * NCH;
    st->xlim = yq::requant_exact_limit(l);
    // can max(x, 0) reach xlim?  x <= 255 * sum of the positive differences + bias
    st->checkx = false;
    for (int oc = 0; oc < l->n; ++oc) {
        long long pos = 0;
        for (int i = 0; i < 27; ++i) {
            const int d = (int)l->host_w[(size_t)oc * 27 + i] - zpw(oc);
            if (d > 0) pos += d;
        }
        if (255 * pos + (long long)l->host_chanq[(size_t)oc * 4] >= (long long)st->xlim) st->checkx = true;
    }
    // bias folding: the two pad bytes of every operand row hold (255, 1), so bias = 255 q + r with r in [-127, 127] in one filter byte
    // and q = q_h + q_l in the two signed blocks' byte 30 of the ky = 0 tile -- possible while |bias| <= 254 * 255 + 127
    st->biasf = !(getenv("YQ_L0_BIASF") && !atoi(getenv("YQ_L0_BIASF")));
    for (int oc = 0; oc < l->n; ++oc) {
        const long long b = l->host_chanq[(size_t)oc * 4];
        if (b > 64897 || b < -64897) st->biasf = false;
    }
    const char *tag = st->biasf ? "l0.2b" : "l0.2";
    const size_t img_bytes = (size_t)6 * N * 32;
    std::vector<uint8_t> img;
    const bool cached = yq::pack_fetch(l, tag, img) && img.size() == img_bytes;
    if (!cached) {
        img.assign(img_bytes, 0);
        // TMEM column col = 32 pair + 8 g + 2 q + e: channel 4 q + g, output pixel 2 pair + e of the 8-pixel segment;
        // K byte k = 10 ci + xi: colour ci, input pixel xi (0..9; the segment's window starts one pixel to the left)
        for (int part = 0; part < 2; ++part)
            for (int ky = 0; ky < 3; ++ky) {
                uint8_t *tile = img.data() + (size_t)(part * 3 + ky) * N * 32;
                for (int col = 0; col < N; ++col) {
                    const int pair = col / 32, g = (col % 32) / 8, q = (col % 8) / 2, e = col % 2;
                    const int oc = 4 * q + g, px = 2 * pair + e;
                    if (oc >= l->n) continue;
                    for (int ci = 0; ci < 3; ++ci)
                        for (int kx = 0; kx < 3; ++kx) {
                            const int d = (int)l->host_w[(((size_t)oc * 3 + ci) * 3 + ky) * 3 + kx] - zpw(oc);
                            const int h = d < -128 ? -128 : d > 127 ? 127 : d;
                            tile[l0_bpos(col, 10 * ci + px + kx)] = (uint8_t)(int8_t)(part == 0 ? h : d - h);
                        }
                    if (st->biasf && ky == 0) {
                        const int b = l->host_chanq[(size_t)oc * 4];
                        int qq = (b >= 0 ? b + 127 : b - 127) / 255;      // round to nearest: |b - 255 qq| <= 127
                        const int r = b - 255 * qq;
                        const int qh = qq < -128 ? -128 : qq > 127 ? 127 : qq;
                        tile[l0_bpos(col, 30)] = (uint8_t)(int8_t)(part == 0 ? qh : qq - qh);
                        tile[l0_bpos(col, 31)] = (uint8_t)(int8_t)(part == 0 ? r : 0);
                    }
                }
            }
        yq::pack_put(l, tag, img);
    }
    if (yq::pack_fetch_device(l, tag, img.size(), (void **)&st->wimg)) {
        // (data-parallel replica: the image came from the arena blob on this device)
    } else if (cudaMalloc((void **)&st->wimg, img.size()) != cudaSuccess || cudaMemcpy(st->wimg, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(st->wimg);
        delete st;
        return yq::fail("layer-0 flavour: weight upload failed");
    }
    *state = st;
    return 0;
}

void yq_tc_l0_free(void *state)
{
    L0State *st = (L0State *)state;
    if (!st) return;
    cudaFree(st->wimg);
    delete st;
}

int yq_tc_l0_forward(yq_conv_layer *l, void *state, const uint8_t *in_planes, uint8_t *out_pool, const yq_act_geom *og, int batch, cudaStream_t stream)
{
    L0State *st = (L0State *)state;
    if (!st || !in_planes || !out_pool || !og) return yq::fail("layer-0 flavour: bad argument");
    if ((uintptr_t)in_planes & 15) return yq::fail("layer-0 flavour: the input must be 16-byte aligned");
    L0Args a;
    memset(&a, 0, sizeof a);
    a.out_pool = out_pool; a.wimg = st->wimg;
    a.PH = l->out_h / 2; a.PW = l->out_w / 2; a.OHP = og->rows_h; a.OWP = og->pitch_w; a.opad = og->pad; a.out_cs = l->cs_out;
    a.tiles_x = (l->out_w + L0_TWPX - 1) / L0_TWPX;
    a.tiles_y = (l->out_h + L0_ROWS - 1) / L0_ROWS;
    a.num_tiles = a.tiles_x * a.tiles_y * batch;
    a.zp_out = l->zp_out;
    a.xlim = st->xlim;
    a.knobs = getenv("YQ_L0_KNOBS") ? atoi(getenv("YQ_L0_KNOBS")) : 0;
    a.magic_x = (uint32_t)((0x100000000ull + a.tiles_x - 1) / a.tiles_x);
    a.magic_y = (uint32_t)((0x100000000ull + a.tiles_y - 1) / a.tiles_y);
    if ((unsigned long long)a.num_tiles * (a.tiles_x > a.tiles_y ? a.tiles_x : a.tiles_y) >= 0x100000000ull || yq_act_geom_bytes(og, batch, l->n) >= 0x100000000ull)
        return yq::fail("layer-0 flavour: tensor too large for 32-bit tile arithmetic");
    memcpy(a.cq, l->host_chanq.data(), (size_t)st->NCH * 16);
    memcpy(a.mc, l->host_mcomb.data(), (size_t)st->NCH * 8);
    const auto key = std::make_pair((const void *)in_planes, batch);
    auto it = st->maps.find(key);
    if (it == st->maps.end()) {
        CUtensorMap m;
        if (l0_encode(&m, in_planes, l->w, l->h, 3 * batch)) return -1;
        if (st->maps.size() > 64) st->maps.clear();
        it = st->maps.emplace(key, m).first;
    }
    // form of the kernel: defaults from measurements on B200 (profiles/README.md); YQ_L0_GROUPS = 1 / 2 and YQ_L0_BULK = 0 / 1 force
    static const int groups_env = getenv("YQ_L0_GROUPS") ? atoi(getenv("YQ_L0_GROUPS")) : -1;
    static const int bulk_env = getenv("YQ_L0_BULK") ? atoi(getenv("YQ_L0_BULK")) : -1;
    const int groups = groups_env < 0 ? 2 : (groups_env >= 1 && groups_env <= 4 ? groups_env : 2);
    const bool bulk = bulk_env < 0 ? true : bulk_env != 0;
    const CUtensorMap &tm = it->second;
    // YQ_L0_TSTORE = 1: pooled rows by tensor store (conflict-free staging) when the row is a whole number of segments.  Off by default:
    // measured equal on B200 (0.0835 vs 0.0831 ms) -- the 16-byte boxes cost the TMA unit what the bank conflicts cost the LSU
    static const int tstore_env = getenv("YQ_L0_TSTORE") ? atoi(getenv("YQ_L0_TSTORE")) : -1;
    a.tstore = (tstore_env < 0 ? 0 : tstore_env != 0) && bulk && groups != 4 && a.PW % 4 == 0 && l->cs_out == 16;
    CUtensorMap tmo;
    memset(&tmo, 0, sizeof tmo);
    if (a.tstore) {
        const auto okey = std::make_pair((const void *)out_pool, batch);
        auto ot = st->omaps.find(okey);
        if (ot == st->omaps.end()) {
            CUtensorMap m;
            if (l0_encode_out(&m, out_pool + ((size_t)og->pad * og->pitch_w + og->pad) * 16, a.PW / 4, og->pitch_w, (long long)batch * og->rows_h - og->pad)) return -1;
            if (st->omaps.size() > 64) st->omaps.clear();
            ot = st->omaps.emplace(okey, m).first;
        }
        tmo = ot->second;
    }
#define YQ_L0(G_, B_)                                                                                                                   \
    (st->biasf ? (st->checkx ? l0_launch<16, G_, true, B_, true>(tm, tmo, a, stream) : l0_launch<16, G_, false, B_, true>(tm, tmo, a, stream))    \
               : (st->checkx ? l0_launch<16, G_, true, B_, false>(tm, tmo, a, stream) : l0_launch<16, G_, false, B_, false>(tm, tmo, a, stream)))
    if (groups == 4) return YQ_L0(4, true);
    if (groups == 3) return bulk ? YQ_L0(3, true) : YQ_L0(3, false);
    if (groups == 2) return bulk ? YQ_L0(2, true) : YQ_L0(2, false);
    return bulk ? YQ_L0(1, true) : YQ_L0(1, false);
#undef YQ_L0
}

#ifdef YQ_L0_TRACE
extern "C" __attribute__((visibility("default"))) int yq_debug_l0_trace(void *host, size_t bytes)
{
    return cudaMemcpyFromSymbol(host, yq_l0_trace, bytes < sizeof(yq_l0_trace) ? bytes : sizeof(yq_l0_trace)) == cudaSuccess ? 0 : -1;
}
#endif
