// yq_conv_tc_rows.cu -- tcgen05 (kind::i8) 3x3 / stride 1 / pad 1 convolution + RELU6 + fused 2x2/2 max-pool for
// SMALL input-channel counts (c <= 32 and c = 64: layers 0, 2, 4, 6 of yolov3-tiny) with NO im2col gather at all.
//
// These layers are HBM-bound on paper (45..384 op/B) but were issue-bound in practice: every thread built one im2col
// row per pixel.  Here a producer warp streams each activation tile ONCE into shared memory with TMA (from a halo-padded
// NHWC tensor, or -- layer 0 -- straight from the network's CHW planes, interleaved on chip) and the tensor core reads
// overlapping windows of it through a NO-SWIZZLE K-major descriptor whose leading-byte-offset makes consecutive K
// chunks overlap:
//
//   c = 4  (layer 0): one MMA row = 4 output pixels.  Row i's K = 32 bytes = input pixels 4i-1 .. 4i+6 of one image
//          row = two 16-byte chunks at 16i and 16i+16 (LBO = 16: chunk 1 of row i IS chunk 0 of row i+1).  The filter
//          bank is a Toeplitz matrix [4 pixels x n channels][8 input pixels x 4 channels]; the tensor core multiplies
//          by its zeros for free.  3 MMAs per tile (one per filter row ky, A start shifted by one tile row).
//   c = 16, 32: TMA boxes over the input viewed as [rows][pitch / 2][parity][c] land the EVEN and the ODD pixel columns
//          in separate planes, and one MMA computes the even and the odd output pixel of a row into different TMEM
//          column ranges from K chunks (E[i + s], O[i + s]) (6 MMAs per tile and 16-channel block).
//
// Either way one TMEM lane then holds BOTH x-neighbours of a pooling window in its columns, tile rows are ordered so
// that lanes {i, i+8} / {i+16, i+24} are y-neighbours, and the 16x256b tcgen05.ld shape hands one thread complete 2x2
// windows: max in the accumulator domain, requantize the winner only (exact, see yq_conv_tc_small.cu), FP64 per-pixel
// fallback when a byte would wrap or |x| >= 2^22.  Restates convolutional_layer.c:694-751 + maxpool_layer.c:109-153.
//
// Zero point of the weights: either 16 extra all-ones filter rows per MMA group give sum(a) per output pixel
// (acc = sum w*a - zp_w * sum a, convolutional_layer.c:718-721), or the filters enter as two signed blocks
// h + l = w - zp_w (RowsCfg, TWO).  Padding: the input tensor carries a halo filled with zp_in (im2col.c:5-14), written
// once by the host runtime -- or TMA's zero fill outside the image for the planar layer-0 form (zp_in = 0); pad
// channels never count (their weights are 0).  Pipeline, template forms and their measured trade-offs: DESIGN.md 4.1.
#include <cuda.h>
#include <cuda_runtime.h>
#include <limits.h>
#include <string.h>

#include <map>
#include <type_traits>
#include <utility>
#include <vector>

#include "yq_common.h"
#include "yq_epilogue.cuh"
#include "yq_tc_ptx.cuh"

using namespace yqtc;

namespace {

constexpr int RW_THREADS = 128;
constexpr int TILE_ROWS = 16;           // conv output rows per tile
constexpr int PLANE = 144;              // bytes of one shared-memory plane row: 9 x 16
#ifndef YQ_ROWS_NBUF
#define YQ_ROWS_NBUF 3
#endif

// -DYQ_ROWS_TRACE: the producer lane of every CTA adds up the clocks of its phases (yq_rows_trace[cta * 8 + phase], [.. + 7] = tiles):
// 0 wait acc_empty, 1 wait full, 2 MMA issue + commit, 3 wait acc_full of the previous tile, 4 refill (tile split + TMA issue),
// 5 / 6 = clocks / nanoseconds of the whole loop (their ratio is the SM clock the kernel really ran at)
#ifdef YQ_ROWS_TRACE
__device__ unsigned long long yq_rows_trace[5 * 8 * 1024];      // slot (0: c = 4, 1: c = 16, 2: c = 32, 3 / 4: c = 64 slices) x CTA x phase
__device__ __forceinline__ unsigned long long tr_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ unsigned long long yq_rows_trace2[5 * 4 * 1024];     // ns: kernel entry, after the PDL wait, loop start, kernel exit -- minus loop start
#define TR_ENTRY const unsigned long long tr_entry = tr_ns(); unsigned long long tr_pdl = 0, tr_loop = 0; bool tr_me = false
#define TR_PDL() tr_pdl = tr_ns()
#define TR_EXIT() do { if (tr_me && blockIdx.x < 1024) { unsigned long long *q = yq_rows_trace2 + (a.trace_slot * 1024 + blockIdx.x) * 4; q[0] = tr_loop - tr_entry; q[1] = tr_loop - tr_pdl; q[2] = tr_loop; q[3] = tr_ns() - tr_loop; } } while (0)
#define TR_DECL tr_loop = tr_ns(); tr_me = true; long long tr_prev = clock64(); const long long tr_c0 = tr_prev; const unsigned long long tr_n0 = tr_ns(); unsigned long long tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define TR_MARK(ev) do { const long long tr_now = clock64(); tr_acc[ev] += (unsigned long long)(tr_now - tr_prev); tr_prev = tr_now; } while (0)
#define TR_FLUSH() do { tr_acc[5] = (unsigned long long)(clock64() - tr_c0); tr_acc[6] = tr_ns() - tr_n0; if (blockIdx.x < 1024) for (int e = 0; e < 8; ++e) yq_rows_trace[(a.trace_slot * 1024 + blockIdx.x) * 8 + e] = tr_acc[e]; } while (0)
#else
#define TR_ENTRY
#define TR_PDL()
#define TR_EXIT()
#define TR_DECL
#define TR_MARK(ev)
#define TR_FLUSH()
#endif

#ifdef YQ_ROWS_KNOBS
#define KNOB(bit) (a.knobs & (bit))
#else
#define KNOB(bit) 0
#endif

template <int CS>
struct RowsGeom {
    static constexpr int TWPX = CS == 4 ? 32 : 16;                  // conv output pixels per tile row
    static constexpr int NBLK = CS == 4 ? 1 : CS / 16;              // 16-channel blocks (c = 16, 32, 64)
    static constexpr int ROWP = CS == 4 ? PLANE : 2 * PLANE * NBLK; // shared-memory bytes per tile row
    static constexpr int A_ROWS = TILE_ROWS + 2;
    static constexpr int NPLANES = CS == 4 ? 1 : 2 * NBLK;          // [18 rows][9 x 16 B] arrays of a tile: one per 16-channel block and pixel parity
    static constexpr int TX_BYTES = NPLANES * A_ROWS * PLANE;       // what TMA delivers per tile
    static constexpr int PLANE_BYTES = (A_ROWS * PLANE + 127) / 128 * 128;   // plane stride: TMA destinations are 128-byte aligned (2592 -> 2688)
    static constexpr int A_BYTES = NPLANES * PLANE_BYTES;
    static constexpr int NMMA = CS == 4 ? 3 : 6 * NBLK;
};

// TWO: the weights enter as two SIGNED blocks h + l = w - zp_w (each in [-128, 127]) multiplied with the same activation
// tile, so the accumulator is the zero-point-corrected one and neither the all-ones rows nor the per-output correction
// exist; twice the MMAs, 4 multiply-adds and one TMEM load less per pooled output (prepare falls back when some w - zp_w = 255)
// PLANAR (c = 3 only): the kernel reads the network input itself -- CHW planes, no halo (TMA zero-fills outside the image,
// so zp_in must be 0) -- and its producer warp interleaves each tile into the pixel-major form the Toeplitz MMAs read,
// instead of a separate layout-transform launch writing a padded NHWC4 copy.
template <int CS, int NCH, bool TWO, bool DBL, bool PLANAR = false>
struct RowsCfg {
    static_assert(!PLANAR || CS == 4, "planar input: layer-0 class only");
    using G = RowsGeom<CS>;
    // all-ones filter rows per MMA group: c = 4 needs one column pair per pixel pair of the segment (16); c >= 16 reads one
    // 8-column TMEM atom of identical sums per group
    static constexpr int NSUM = TWO ? 0 : (CS == 4 ? 16 : 8);
    static constexpr int NB = CS == 4 ? 4 * NCH + NSUM : NCH + NSUM;    // filter rows (TMEM columns) per MMA group
    static constexpr int NACC = CS == 4 ? NB : 2 * NB;              // TMEM columns of one accumulator
    // DB: two accumulators, the next tile's MMAs run under this tile's epilogue (at the price of CTAs per SM once two of
    // them need more than 128 columns)
    static constexpr bool DB = DBL;
    static constexpr int NTM = NACC * (DB ? 2 : 1);
    static constexpr int TMEM_COLS = NTM <= 32 ? 32 : NTM <= 64 ? 64 : NTM <= 128 ? 128 : NTM <= 256 ? 256 : 512;
    static constexpr int MAX_CTAS = 512 / TMEM_COLS;
    // input-tile ring: the copy of tile i + NBUF - 1 is in flight while tile i is computed (DB consumes a tile one step earlier)
    static constexpr int NBUF = DB ? YQ_ROWS_NBUF + 1 : YQ_ROWS_NBUF;
    static_assert(NBUF <= 4, "barrier block");
    static constexpr int NMMA_N = CS == 4 ? NB : 2 * NB;            // N of one MMA: for c >= 16 the even and the odd group share it
    static constexpr int BSUB = NMMA_N * 32;                        // one MMA's filter tile
    static constexpr int NMMA = G::NMMA * (TWO ? 2 : 1);
    static constexpr int B_BYTES = NMMA * BSUB;
    static constexpr int A_STRIDE = (G::A_BYTES + 32 + 127) / 128 * 128;   // +32: the last window's (zero-weight) overhang
    static constexpr int A_OFF = 0;
    // PLANAR with two accumulators: two producer warps, one per accumulator (tiles alternate between them), each with its own
    // staging slots and operand tiles -- one warp's interleave + MMA issue per tile is otherwise the slowest stage
    static constexpr int NPROD = PLANAR && DB ? 2 : 1;
    static constexpr int NA = PLANAR ? (NPROD == 2 ? 4 : 3) : NBUF; // MMA operand tiles (PLANAR: written by the producer warps)
    // PLANAR staging: [3 planes][18 rows][64 bytes] per tile.  A TMA box must start on a 16-byte boundary of a row, so it starts 16
    // columns left of the tile (x0 - 16); the 36 columns the tile reads (x0 - 1 ...) sit at bytes 15 .. 50
    static constexpr int S_COLS = 64;
    static constexpr int S_BYTES = 3 * G::A_ROWS * S_COLS;
    static constexpr int S_STRIDE = (S_BYTES + 127) / 128 * 128;
    static constexpr int S_OFF = NA * A_STRIDE;
    static constexpr int B_OFF = S_OFF + (PLANAR ? NBUF * S_STRIDE : 0);
    static constexpr int BAR_OFF = B_OFF + B_BYTES;
    static constexpr int TOTAL = BAR_OFF + 128;
    static_assert(NMMA_N % 16 == 0 && NMMA_N <= 256, "kind::i8 N");
};

struct RowsArgs {
    const uint8_t *in;      // halo-padded NHWC input: pixel (n, y, x) at in + ((n*HP + y + 1)*WP + x + 1)*CS
    uint8_t *out_pool;      // pooled output:          pixel (n, y, x) at out + ((n*OHP + y + opad)*OWP + x + opad)*NCH
    const uint8_t *wimg;    // shared-memory image of the filter tiles, one per MMA in issue order
    int HP, WP, OH, OW, PH, PW, OHP, OWP, opad;
    int tiles_x, tiles_y, num_tiles, zp_out;
    int trace_slot;         // -DYQ_ROWS_TRACE: which block of yq_rows_trace this launch fills
    int out_cs, ch_off;     // channel stride of the output tensor and first channel this launch writes (c = 64: n is done in slices of 64)
    int knobs;              // -DYQ_ROWS_KNOBS experiments (results are garbage): 1 no MMAs, 2 no epilogue arithmetic, 4 no tile copies, 8 no stores
    uint32_t magic_x, magic_y;   // ceil(2^32 / tiles_x), ceil(2^32 / tiles_y): exact quotients by __umulhi for tile < 2^32 / tiles
    int4 cq[64];            // {bias, zw, 2*M0, shift} per channel
    double mc[64];          // M_value * 2^-s (FP64 fallback)
};

// K-major, no swizzle: 8 rows x 16 bytes contiguous per core matrix; LBO = byte step between the two K chunks of one
// MMA, SBO = byte step between 8-row groups (cute::UMMA canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units)
__device__ __forceinline__ uint64_t make_desc_ns(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void tmem_ldq4(uint32_t taddr, uint32_t (&v0)[16], uint32_t (&v1)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%32];\n\t"
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%33];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(v0[0]), "=r"(v0[1]), "=r"(v0[2]), "=r"(v0[3]), "=r"(v0[4]), "=r"(v0[5]), "=r"(v0[6]), "=r"(v0[7]), "=r"(v0[8]), "=r"(v0[9]),
          "=r"(v0[10]), "=r"(v0[11]), "=r"(v0[12]), "=r"(v0[13]), "=r"(v0[14]), "=r"(v0[15]), "=r"(v1[0]), "=r"(v1[1]), "=r"(v1[2]),
          "=r"(v1[3]), "=r"(v1[4]), "=r"(v1[5]), "=r"(v1[6]), "=r"(v1[7]), "=r"(v1[8]), "=r"(v1[9]), "=r"(v1[10]), "=r"(v1[11]),
          "=r"(v1[12]), "=r"(v1[13]), "=r"(v1[14]), "=r"(v1[15])
        : "r"(taddr), "r"(taddr + (16u << 16)));
}
// even / odd column groups, both 16-lane halves: 4 x (.x2)
__device__ __forceinline__ void tmem_ldq_eo(uint32_t te, uint32_t to, uint32_t (&e0)[8], uint32_t (&e1)[8], uint32_t (&o0)[8], uint32_t (&o1)[8])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%32];\n\t"
        "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%8, %9, %10, %11, %12, %13, %14, %15}, [%33];\n\t"
        "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%16, %17, %18, %19, %20, %21, %22, %23}, [%34];\n\t"
        "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%24, %25, %26, %27, %28, %29, %30, %31}, [%35];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(e0[0]), "=r"(e0[1]), "=r"(e0[2]), "=r"(e0[3]), "=r"(e0[4]), "=r"(e0[5]), "=r"(e0[6]), "=r"(e0[7]), "=r"(e1[0]), "=r"(e1[1]),
          "=r"(e1[2]), "=r"(e1[3]), "=r"(e1[4]), "=r"(e1[5]), "=r"(e1[6]), "=r"(e1[7]), "=r"(o0[0]), "=r"(o0[1]), "=r"(o0[2]), "=r"(o0[3]),
          "=r"(o0[4]), "=r"(o0[5]), "=r"(o0[6]), "=r"(o0[7]), "=r"(o1[0]), "=r"(o1[1]), "=r"(o1[2]), "=r"(o1[3]), "=r"(o1[4]), "=r"(o1[5]),
          "=r"(o1[6]), "=r"(o1[7])
        : "r"(te), "r"(te + (16u << 16)), "r"(to), "r"(to + (16u << 16)));
}
// activation-sum columns of the even / odd groups, both halves: 4 x (.x1)
__device__ __forceinline__ void tmem_ldq_sums(uint32_t te, uint32_t to, uint32_t (&e0)[4], uint32_t (&e1)[4], uint32_t (&o0)[4], uint32_t (&o1)[4])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%16];\n\t"
        "tcgen05.ld.sync.aligned.16x256b.x1.b32 {%4, %5, %6, %7}, [%17];\n\t"
        "tcgen05.ld.sync.aligned.16x256b.x1.b32 {%8, %9, %10, %11}, [%18];\n\t"
        "tcgen05.ld.sync.aligned.16x256b.x1.b32 {%12, %13, %14, %15}, [%19];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(e0[0]), "=r"(e0[1]), "=r"(e0[2]), "=r"(e0[3]), "=r"(e1[0]), "=r"(e1[1]), "=r"(e1[2]), "=r"(e1[3]), "=r"(o0[0]), "=r"(o0[1]),
          "=r"(o0[2]), "=r"(o0[3]), "=r"(o1[0]), "=r"(o1[1]), "=r"(o1[2]), "=r"(o1[3])
        : "r"(te), "r"(te + (16u << 16)), "r"(to), "r"(to + (16u << 16)));
}

// The same loads split into issue and wait, so that a chunk's TMEM round trip runs under the previous chunk's arithmetic
// (the wait names the registers as in/out operands: their consumers are ordered after it)
__device__ __forceinline__ void tmem_ldq4_issue(uint32_t taddr, uint32_t (&v0)[16], uint32_t (&v1)[16])
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
__device__ __forceinline__ void tmem_wait32(uint32_t (&v0)[16], uint32_t (&v1)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v0[0]), "+r"(v0[1]), "+r"(v0[2]), "+r"(v0[3]), "+r"(v0[4]), "+r"(v0[5]), "+r"(v0[6]), "+r"(v0[7]), "+r"(v0[8]), "+r"(v0[9]),
                   "+r"(v0[10]), "+r"(v0[11]), "+r"(v0[12]), "+r"(v0[13]), "+r"(v0[14]), "+r"(v0[15]), "+r"(v1[0]), "+r"(v1[1]), "+r"(v1[2]),
                   "+r"(v1[3]), "+r"(v1[4]), "+r"(v1[5]), "+r"(v1[6]), "+r"(v1[7]), "+r"(v1[8]), "+r"(v1[9]), "+r"(v1[10]), "+r"(v1[11]),
                   "+r"(v1[12]), "+r"(v1[13]), "+r"(v1[14]), "+r"(v1[15])::"memory");
}
__device__ __forceinline__ void tmem_ldq_eo_issue(uint32_t te, uint32_t to, uint32_t (&e0)[8], uint32_t (&e1)[8], uint32_t (&o0)[8], uint32_t (&o1)[8])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%32];\n\t"
        "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%8, %9, %10, %11, %12, %13, %14, %15}, [%33];\n\t"
        "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%16, %17, %18, %19, %20, %21, %22, %23}, [%34];\n\t"
        "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%24, %25, %26, %27, %28, %29, %30, %31}, [%35];"
        : "=r"(e0[0]), "=r"(e0[1]), "=r"(e0[2]), "=r"(e0[3]), "=r"(e0[4]), "=r"(e0[5]), "=r"(e0[6]), "=r"(e0[7]), "=r"(e1[0]), "=r"(e1[1]),
          "=r"(e1[2]), "=r"(e1[3]), "=r"(e1[4]), "=r"(e1[5]), "=r"(e1[6]), "=r"(e1[7]), "=r"(o0[0]), "=r"(o0[1]), "=r"(o0[2]), "=r"(o0[3]),
          "=r"(o0[4]), "=r"(o0[5]), "=r"(o0[6]), "=r"(o0[7]), "=r"(o1[0]), "=r"(o1[1]), "=r"(o1[2]), "=r"(o1[3]), "=r"(o1[4]), "=r"(o1[5]),
          "=r"(o1[6]), "=r"(o1[7])
        : "r"(te), "r"(te + (16u << 16)), "r"(to), "r"(to + (16u << 16)));
}
__device__ __forceinline__ void tmem_wait_eo(uint32_t (&e0)[8], uint32_t (&e1)[8], uint32_t (&o0)[8], uint32_t (&o1)[8])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(e0[0]), "+r"(e0[1]), "+r"(e0[2]), "+r"(e0[3]), "+r"(e0[4]), "+r"(e0[5]), "+r"(e0[6]), "+r"(e0[7]), "+r"(e1[0]), "+r"(e1[1]),
                   "+r"(e1[2]), "+r"(e1[3]), "+r"(e1[4]), "+r"(e1[5]), "+r"(e1[6]), "+r"(e1[7]), "+r"(o0[0]), "+r"(o0[1]), "+r"(o0[2]), "+r"(o0[3]),
                   "+r"(o0[4]), "+r"(o0[5]), "+r"(o0[6]), "+r"(o0[7]), "+r"(o1[0]), "+r"(o1[1]), "+r"(o1[2]), "+r"(o1[3]), "+r"(o1[4]), "+r"(o1[5]),
                   "+r"(o1[6]), "+r"(o1[7])::"memory");
}

// Per-launch channel parameters of ONE thread (its NCH/4 channels), kept in registers.
template <int NPQ>
struct ThreadChan {
    int bias[NPQ], zw[NPQ], sh[NPQ];
    uint32_t m2[NPQ];
};

// One pooled output: the window's four raw accumulators v0..v3 with their activation sums -> exact zero-point
// correction, max, bias, RELU6 requantize of the winner.  Returns q + zp_out; *xq is the requantizer's input.
// nzw = -zp_w, n0..n3 = the four pixels' activation sums (as read from TMEM)
__device__ __forceinline__ int pool_requant(int nzw, int bias, uint32_t m2, int sh, int zo, int v0, int v1, int v2, int v3, int n0, int n1, int n2,
                                            int n3, uint32_t *xq)
{
    const int m = max(max(nzw * n0 + v0, nzw * n1 + v1), max(nzw * n2 + v2, nzw * n3 + v3));
    *xq = (uint32_t)max(m + bias, 0);
    return (int)(__umulhi(*xq, m2) >> sh) + zo;
}
// the reference's per-pixel arithmetic for one window (FP64 multiply, uint8 wrap, then the pool)
__device__ __forceinline__ int pool_requant_slow(int nzw, int bias, double mcd, int zo, int v0, int v1, int v2, int v3, int n0, int n1, int n2, int n3)
{
    const int xs[4] = {nzw * n0 + v0 + bias, nzw * n1 + v1 + bias, nzw * n2 + v2 + bias, nzw * n3 + v3 + bias};
    int best = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int q = max(__double2int_rz(__dmul_rn((double)xs[p], mcd)), 0);
        best = max(best, (q + zo) & 255);
    }
    return best;
}
__device__ __forceinline__ uint32_t pack4(const int (&r)[4])
{
    return __byte_perm(__byte_perm((uint32_t)r[0], (uint32_t)r[1], 0x0040), __byte_perm((uint32_t)r[2], (uint32_t)r[3], 0x0040), 0x5410);
}

// Warp-specialised, persistent: one producer warp (a single elected lane) streams input tiles with TMA into a ring of
// NBUF buffers and issues each tile's MMAs; 4 * SPLIT epilogue warps (warp % 4 = TMEM lane quarter) requantize + pool +
// store.  No CTA-wide barrier and no proxy fence per tile: full[] (TMA bytes landed) -> MMAs -> acc_full[] (tcgen05.commit)
// -> epilogue -> acc_empty[] (one arrival per epilogue warp, right after its last TMEM load).
// SPLIT = 2: two warps per lane quarter share a tile's epilogue (c = 4: one pixel pair each; c >= 16: half of the thread's
// channels each).  DBL: two accumulators, so the next tile's MMAs run under this tile's epilogue.
constexpr int rows_min_ctas(int CS, int SPLIT, int by_tmem)
{
    const int want = CS == 64 ? 1 : CS == 32 ? 2 : SPLIT == 2 ? 3 : 4;
    return want < by_tmem ? want : by_tmem;
}

template <int CS, int NCH, int SPLIT, bool TWO, bool DBL, bool PLANAR>
__global__ void __launch_bounds__(RW_THREADS * SPLIT + 32 * RowsCfg<CS, NCH, TWO, DBL, PLANAR>::NPROD,
                                  rows_min_ctas(CS, SPLIT, RowsCfg<CS, NCH, TWO, DBL, PLANAR>::MAX_CTAS))
conv_u8_tc_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ RowsArgs a)
{
    using G = RowsGeom<CS>;
    using L = RowsCfg<CS, NCH, TWO, DBL, PLANAR>;
    constexpr int NPQ = NCH / 4;
    constexpr int NT = RW_THREADS * SPLIT;                 // epilogue threads; the producer warp comes after them
    constexpr int NPT = CS == 4 ? NPQ : NPQ / SPLIT;      // channels whose parameters this thread keeps
    // double-buffered TMEM loads (a chunk's round trip under the previous chunk's arithmetic) where the registers are there:
    // c = 32 runs 2 CTAs per SM; measured slower for the 4-CTA kernels (layers 0, 2), faster for layer 4
    constexpr bool PIPE = CS >= 32;
    constexpr int NACCS = L::DB ? 2 : 1;
    constexpr int NBUF = L::NBUF;
    static_assert(SPLIT == 1 || CS == 4 || NCH >= 32, "a thread needs at least one 4-channel chunk");
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    uint64_t *full = (uint64_t *)(smem + L::BAR_OFF);          // [NBUF] input tile landed
    uint64_t *acc_full = full + NBUF;                          // [2] accumulator complete
    uint64_t *acc_empty = acc_full + 2;                        // [2] accumulator read out
    uint64_t *b_full = acc_empty + 2;                          // the resident filter tiles have landed
    uint32_t *tmem_slot = (uint32_t *)(b_full + 1);

    TR_ENTRY;
    const int t = threadIdx.x, warp = (t >> 5) & 3, half = SPLIT == 2 ? (t >> 7) & 1 : 0, lane = t & 31;
    const bool producer = t >= NT;
    const int qi = lane >> 2, qq = lane & 3;

    // ---- one-time setup: barriers, TMEM, and the resident filter tiles as bulk copies (up to 123 KB per CTA: fetched by
    // dependent 16-byte loads of 160 threads this took ~ 13 us of a 30 us layer-6 launch; they are constants, so they start
    // before the wait on the previous kernel)
    if (t == 0) {
        for (int b = 0; b < NBUF; ++b) mbar_init(&full[b], 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], NT / 32);
        }
        mbar_init(b_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(b_full, (uint32_t)L::B_BYTES);
        constexpr int CHUNK = 32768;
#pragma unroll 1
        for (int off = 0; off < L::B_BYTES; off += CHUNK)
            bulk_load(smem + L::B_OFF + off, a.wimg + off, (uint32_t)(L::B_BYTES - off < CHUNK ? L::B_BYTES - off : CHUNK), b_full);
    }
    if (t < 32) tmem_alloc<L::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    struct TileXY { int tx, ty, n; };
    auto split_tile = [&](int tile) -> TileXY {
        // (a divisor of 1 has no 32-bit magic: 2^32)
        const int r1 = a.tiles_x == 1 ? tile : (int)__umulhi((uint32_t)tile, a.magic_x);
        const int n = a.tiles_y == 1 ? r1 : (int)__umulhi((uint32_t)r1, a.magic_y);
        return TileXY{tile - r1 * a.tiles_x, r1 - n * a.tiles_y, n};
    };
    const int first = blockIdx.x, step = gridDim.x;

    yq_pdl_wait_then_release();      // everything above touched only constants and on-chip state
    TR_PDL();

    if (producer) {
        const uint32_t sA = smem_u32(smem + L::A_OFF), b0 = smem_u32(smem + L::B_OFF);
        if constexpr (PLANAR) {
            // the whole warp: wait for a tile's three planes, interleave them into an operand tile, then one lane refills the
            // staging slot and issues the MMAs
            auto load_planes = [&](int tile, int sbuf) {
                const TileXY p = split_tile(tile);
                mbar_expect_tx(&full[sbuf], (uint32_t)L::S_BYTES);
                tma_load_3d(smem + L::S_OFF + sbuf * L::S_STRIDE, &tmA, &full[sbuf], p.tx * G::TWPX - 16, p.ty * TILE_ROWS - 1, p.n * 3);
            };
            auto issue_mma_planar = [&](int abuf, int acc) {
                constexpr uint32_t idesc = make_idesc(L::NMMA_N) | (TWO ? 1u << 10 : 0u);
                const uint32_t a0 = sA + abuf * L::A_STRIDE;
                const uint32_t tacc = tmem_base + (uint32_t)(acc * L::NACC);
#pragma unroll
                for (int part = 0; part < (TWO ? 2 : 1); ++part)
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
                        umma_i8(tacc, make_desc_ns(a0 + ky * PLANE, 16, PLANE), make_desc_ns(b0 + (part * 3 + ky) * L::BSUB, 128, 256), idesc,
                                (part | ky) ? 1u : 0u);
                umma_commit(&acc_full[acc]);
            };
            constexpr int NPROD = L::NPROD, SPP = NBUF / NPROD, APP = L::NA / NPROD;      // staging slots / operand tiles per producer warp
            static_assert(NBUF % NPROD == 0 && L::NA % NPROD == 0 && (NPROD == 1 || NACCS == 2), "producer warps split the rings evenly");
            const int pw = (t - NT) >> 5;                  // this producer warp handles tiles it = pw, pw + NPROD, ...
            if (lane == 0) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
#pragma unroll 1
                for (int d = 0; d < SPP; ++d)
                    if (first + (pw + d * NPROD) * step < a.num_tiles) load_planes(first + (pw + d * NPROD) * step, pw * SPP + d);
            }
            __syncwarp();
            constexpr int PL = G::A_ROWS * L::S_COLS;      // bytes of one staged plane
            constexpr int NGROUPS = G::A_ROWS * 9, NG = (NGROUPS + 31) / 32;
            int src_off[NG], dst_off[NG];                  // this lane's groups: (row r, pixels 4 q4 .. 4 q4 + 3), fixed for the launch
#pragma unroll
            for (int i = 0; i < NG; ++i) {
                const int g = lane + 32 * i, r = g / 9, q4 = g - r * 9;
                src_off[i] = r * L::S_COLS + 12 + 4 * q4;
                dst_off[i] = r * PLANE + 16 * q4;
            }
            int k = 0;
#pragma unroll 1
            for (int tile = first + pw * step; tile < a.num_tiles; tile += NPROD * step, ++k) {
                const int it = pw + k * NPROD;
                const int sbuf = pw * SPP + k % SPP, abuf = pw * APP + k % APP;
                mbar_wait(&full[sbuf], (uint32_t)((k / SPP) & 1));
                const uint8_t *S = smem + L::S_OFF + sbuf * L::S_STRIDE;
                uint8_t *A = smem + L::A_OFF + abuf * L::A_STRIDE;
                // 18 rows x 9 groups of 4 pixels: (R, G, B) words of 4 pixels -> 4 pixel-major words (byte 3 of a pixel meets zero
                // weights).  All loads of the lane's 6 groups first, so their latencies overlap.
                uint32_t wr[NG][6];
#pragma unroll
                for (int i = 0; i < NG; ++i) {
                    const uint32_t *w = reinterpret_cast<const uint32_t *>(S + src_off[i]);       // bytes 12 + 4 q4 .. 19 + 4 q4 of the row
                    if (i + 1 < NG || lane + 32 * i < NGROUPS) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            wr[i][2 * c] = w[c * (PL / 4)];
                            wr[i][2 * c + 1] = w[c * (PL / 4) + 1];
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < NG; ++i) {
                    if (i + 1 < NG || lane + 32 * i < NGROUPS) {
                        const uint32_t R4 = __funnelshift_r(wr[i][0], wr[i][1], 24);              // bytes 15 + 4 q4 .. 18 + 4 q4
                        const uint32_t G4 = __funnelshift_r(wr[i][2], wr[i][3], 24);
                        const uint32_t B4 = __funnelshift_r(wr[i][4], wr[i][5], 24);
                        const uint32_t t01 = __byte_perm(R4, G4, 0x5140), t23 = __byte_perm(R4, G4, 0x7362);
                        *reinterpret_cast<uint4 *>(A + dst_off[i]) = make_uint4(__byte_perm(t01, B4, 0x4410), __byte_perm(t01, B4, 0x5532),
                                                                                __byte_perm(t23, B4, 0x6610), __byte_perm(t23, B4, 0x7732));
                    }
                }
                fence_proxy_async();      // the operand tile (generic-proxy stores) -> visible to the tensor core
                __syncwarp();
                if (lane == 0) {
                    if (k == 0) mbar_wait(b_full, 0);
                    if (tile + NBUF * step < a.num_tiles) load_planes(tile + NBUF * step, sbuf);      // every lane has read the slot (SPP * NPROD = NBUF tiles on)
                    const int acc = NACCS == 2 ? it & 1 : 0;
                    if (it >= NACCS) mbar_wait(&acc_empty[acc], ((uint32_t)(it / NACCS) & 1u) ^ 1u);
                    tc_fence_after();
                    issue_mma_planar(abuf, acc);
                }
                __syncwarp();             // (the operand tile rewritten next had its MMAs >= 2 of this warp's tiles back: known complete by now)
            }
        } else if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            auto load_tile = [&](int tile, int buf) {     // 18 halo rows of the tile: one box (c = 4) or one box per 16-channel block and pixel parity
                const TileXY p = split_tile(tile);
                uint8_t *dst = smem + L::A_OFF + buf * L::A_STRIDE;
                const int row = p.n * a.HP + p.ty * TILE_ROWS;
                if (KNOB(4)) {
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[buf])) : "memory");
                    return;
                }
                mbar_expect_tx(&full[buf], (uint32_t)G::TX_BYTES);
                if (CS == 4) {
                    tma_load_2d(dst, &tmA, &full[buf], p.tx * G::TWPX * 4, row);
                } else {
#pragma unroll
                    for (int blk = 0; blk < G::NBLK; ++blk)
#pragma unroll
                        for (int par = 0; par < 2; ++par)      // plane E = odd padded columns (parity 1), plane O = even ones
                            tma_load_4d(dst + (blk * 2 + (par ^ 1)) * G::PLANE_BYTES, &tmA, &full[buf], blk * 16, par, p.tx * (G::TWPX / 2), row);
                }
            };
            auto issue_mma = [&](int buf, int acc) {
                constexpr uint32_t idesc = make_idesc(L::NMMA_N) | (TWO ? 1u << 10 : 0u);      // TWO: B (the filters) is SINT8
                const uint32_t a0 = sA + buf * L::A_STRIDE;
                const uint32_t tacc = tmem_base + (uint32_t)(acc * L::NACC);
                if (CS == 4) {
#pragma unroll
                    for (int part = 0; part < (TWO ? 2 : 1); ++part)
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
                            umma_i8(tacc, make_desc_ns(a0 + ky * PLANE, 16, PLANE), make_desc_ns(b0 + (part * 3 + ky) * L::BSUB, 128, 256), idesc,
                                    (part | ky) ? 1u : 0u);
                } else {
                    // one MMA = K chunks (E[i + step], O[i + step]) of one image row and 16-channel block, N = even group | odd group:
                    //   even pixel 2i   : kx 1 = E[i], kx 0 = O[i]   (step 0);                 kx 2 = O[i+1] (step 1)
                    //   odd  pixel 2i+1 : kx 0 = E[i]                (step 0);  kx 2 = E[i+1], kx 1 = O[i+1] (step 1)       [O[] is the shifted plane]
                    // planes are separate [18 rows][9 x 16 B] arrays: LBO = E -> O plane, SBO = next image row
                    int m = 0;
#pragma unroll
                    for (int part = 0; part < (TWO ? 2 : 1); ++part)
#pragma unroll
                        for (int blk = 0; blk < G::NBLK; ++blk)
#pragma unroll
                            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                                for (int st = 0; st < 2; ++st, ++m)
                                    umma_i8(tacc, make_desc_ns(a0 + blk * 2 * G::PLANE_BYTES + ky * PLANE + st * 16, G::PLANE_BYTES, PLANE),
                                            make_desc_ns(b0 + m * L::BSUB, 128, 256), idesc, m ? 1u : 0u);
                }
                umma_commit(&acc_full[acc]);
            };
#pragma unroll 1
            for (int d = 0; d < NBUF; ++d)
                if (first + d * step < a.num_tiles) load_tile(first + d * step, d);
            mbar_wait(b_full, 0);
            int buf = 0, it = 0;
            TR_DECL;
#pragma unroll 1
            for (int tile = first; tile < a.num_tiles; tile += step, buf = buf + 1 == NBUF ? 0 : buf + 1, ++it) {
                const int acc = NACCS == 2 ? it & 1 : 0;
                const uint32_t use = (uint32_t)(it / NACCS);            // how many times this accumulator has been used before
                if (it >= NACCS) mbar_wait(&acc_empty[acc], (use & 1u) ^ 1u);      // its previous tile has been read out
                TR_MARK(0);
                mbar_wait(&full[buf], (uint32_t)((it / NBUF) & 1));
                TR_MARK(1);
                tc_fence_after();
                if (KNOB(1)) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_full[acc])) : "memory");
                else issue_mma(buf, acc);
                TR_MARK(2);
                // refill the previous tile's buffer once its MMAs have finished reading it
                // (one accumulator: the acc_empty wait above already implies it -- and by now acc_full may be a phase further)
                if (it >= 1 && tile + (NBUF - 1) * step < a.num_tiles) {
                    if (NACCS == 2) mbar_wait(&acc_full[(it - 1) & 1], (uint32_t)(((it - 1) >> 1) & 1));
                    TR_MARK(3);
                    load_tile(tile + (NBUF - 1) * step, buf == 0 ? NBUF - 1 : buf - 1);
                    TR_MARK(4);
                }
#ifdef YQ_ROWS_TRACE
                tr_acc[7] += 1;
#endif
            }
            TR_FLUSH();
        }
        __syncwarp();
    } else {
        const int ch0 = qq * NPQ + (CS == 4 ? 0 : half * NPT);       // first channel of this thread
        ThreadChan<NPT> ch;
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            const int4 c = a.cq[ch0 + k];
            ch.bias[k] = c.x; ch.zw[k] = TWO ? 0 : -c.y; ch.m2[k] = (uint32_t)c.z; ch.sh[k] = c.w;   // zw holds MINUS the zero point
        }
        const int zo = a.zp_out;
        const uint32_t tq = tmem_base + ((uint32_t)(warp * 32) << 16);
        // per-thread part of the pooled-output address (window 0 of this thread; window 1 is one pooled row further)
        const uint32_t out_thr = (uint32_t)(((2 * warp + a.opad) * a.OWP + (CS == 4 ? 2 * qi : qi) + a.opad) * a.out_cs + a.ch_off + ch0);
        const uint32_t out_row = (uint32_t)(a.OWP * a.out_cs);
        int it = 0;
#pragma unroll 1
        for (int tile = first; tile < a.num_tiles; tile += step, ++it) {
            const TileXY cur = split_tile(tile);
            const int acc = NACCS == 2 ? it & 1 : 0;
            mbar_wait(&acc_full[acc], (uint32_t)((it / NACCS) & 1));
            tc_fence_after();
            const uint32_t tqa = tq + (uint32_t)(acc * L::NACC);      // this tile's accumulator, this warp's lane quarter
            auto release_acc = [&]() {       // this warp's TMEM reads of the tile are complete
                tc_fence_before();
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[acc])) : "memory");
            };
            if (KNOB(2)) {
                release_acc();
                continue;
            }

        // ---- epilogue: warp = TMEM lane quarter = 4 conv rows = 2 pooled rows; thread (qi, qq) = pooled column(s) qi, channels qq*NPQ ..
            const int tx = cur.tx, ty = cur.ty, n = cur.n;
            const int py0 = ty * (TILE_ROWS / 2) + 2 * warp;
            uint8_t *const out_tile = a.out_pool + (size_t)((uint32_t)((n * a.OHP + ty * (TILE_ROWS / 2)) * a.OWP + tx * (G::TWPX / 2)) * (uint32_t)a.out_cs + out_thr);
            if (CS == 4) {
                constexpr int NPAIR = 2 / SPLIT, NJ = NCH / 16, NCK = NPAIR * NJ;    // NCK chunks of 2 x 16 accumulators, double-buffered
                uint32_t V[2][2][16];
                auto chunk_addr = [&](int ck) { return tqa + ((SPLIT == 2 ? half : ck / NJ) * NPQ + 4 * (ck % NJ)) * 8; };
                if (PIPE) tmem_ldq4_issue(chunk_addr(0), V[0][0], V[0][1]);
                uint32_t s0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                if (!TWO) tmem_ldq(tqa + 4 * NCH, s0, s1);        // its wait covers chunk 0 as well
                if (PIPE) {
                    tmem_wait32(V[0][0], V[0][1]);
                    if (NCK == 1) release_acc();
                }
                uint32_t w0[NPAIR][NCH / 16], w1[NPAIR][NCH / 16];
#pragma unroll
                for (int pp = 0; pp < NPAIR; ++pp) {
                    const int pair = SPLIT == 2 ? half : pp;
                    int nsa0[4], nsa1[4];       // the pair's four activation sums
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        nsa0[k] = (int)(SPLIT == 2 ? (half ? s0[4 + k] : s0[k]) : s0[4 * pp + k]);
                        nsa1[k] = (int)(SPLIT == 2 ? (half ? s1[4 + k] : s1[k]) : s1[4 * pp + k]);
                    }
#pragma unroll
                    for (int jj = 0; jj < NCH / 16; ++jj) {
                        const int ck = pp * NJ + jj;
                        if (PIPE) {
                            if (ck + 1 < NCK) tmem_ldq4_issue(chunk_addr(ck + 1), V[(ck + 1) & 1][0], V[(ck + 1) & 1][1]);
                        } else {
                            tmem_ldq4(chunk_addr(ck), V[0][0], V[0][1]);
                            if (ck + 1 == NCK) release_acc();
                        }
                        uint32_t(&v0)[16] = V[PIPE ? ck & 1 : 0][0];
                        uint32_t(&v1)[16] = V[PIPE ? ck & 1 : 0][1];
                        int r0[4], r1w[4];
                        uint32_t orx = 0, orr = 0, xq;
#pragma unroll
                        for (int gg = 0; gg < 4; ++gg) {
                            const int k = 4 * jj + gg;
                            r0[gg] = pool_requant(ch.zw[k], ch.bias[k], ch.m2[k], ch.sh[k], zo, (int)v0[4 * gg], (int)v0[4 * gg + 1], (int)v0[4 * gg + 2],
                                                  (int)v0[4 * gg + 3], nsa0[0], nsa0[1], nsa0[2], nsa0[3], &xq);
                            orx |= xq; orr |= (uint32_t)r0[gg];
                            r1w[gg] = pool_requant(ch.zw[k], ch.bias[k], ch.m2[k], ch.sh[k], zo, (int)v1[4 * gg], (int)v1[4 * gg + 1], (int)v1[4 * gg + 2],
                                                   (int)v1[4 * gg + 3], nsa1[0], nsa1[1], nsa1[2], nsa1[3], &xq);
                            orx |= xq; orr |= (uint32_t)r1w[gg];
                        }
                        if (orx >= (1u << 22) || orr > 255u) {
#pragma unroll
                            for (int gg = 0; gg < 4; ++gg) {
                                const int k = 4 * jj + gg;
                                const double mcd = a.mc[ch0 + k];
                                r0[gg] = pool_requant_slow(ch.zw[k], ch.bias[k], mcd, zo, (int)v0[4 * gg], (int)v0[4 * gg + 1], (int)v0[4 * gg + 2],
                                                           (int)v0[4 * gg + 3], nsa0[0], nsa0[1], nsa0[2], nsa0[3]);
                                r1w[gg] = pool_requant_slow(ch.zw[k], ch.bias[k], mcd, zo, (int)v1[4 * gg], (int)v1[4 * gg + 1], (int)v1[4 * gg + 2],
                                                            (int)v1[4 * gg + 3], nsa1[0], nsa1[1], nsa1[2], nsa1[3]);
                            }
                        }
                        w0[pp][jj] = pack4(r0);
                        w1[pp][jj] = pack4(r1w);
                        if (PIPE && ck + 1 < NCK) {
                            tmem_wait32(V[(ck + 1) & 1][0], V[(ck + 1) & 1][1]);
                            if (ck + 2 == NCK) release_acc();
                        }
                    }
                }
#pragma unroll
                for (int pp = 0; pp < NPAIR; ++pp) {
                    const int pair = SPLIT == 2 ? half : pp;
                    const int px = tx * (G::TWPX / 2) + 2 * qi + pair;
                    if (px < a.PW && !KNOB(8)) {
                        uint8_t *dst = out_tile + pair * a.out_cs;
                        if (py0 < a.PH) {
                            if constexpr (NCH == 16) *reinterpret_cast<uint32_t *>(dst) = w0[pp][0];
                            else *reinterpret_cast<uint2 *>(dst) = make_uint2(w0[pp][0], w0[pp][1]);
                        }
                        if (py0 + 1 < a.PH) {
                            if constexpr (NCH == 16) *reinterpret_cast<uint32_t *>(dst + out_row) = w1[pp][0];
                            else *reinterpret_cast<uint2 *>(dst + out_row) = make_uint2(w1[pp][0], w1[pp][1]);
                        }
                    }
                }
            } else {
                constexpr int JN = NCH / 16 / SPLIT > 0 ? NCH / 16 / SPLIT : 1;      // 4-channel chunks of this thread, double-buffered
                uint32_t E[2][4][8];
                auto chunk_addr = [&](int j) { return tqa + 16 * (half * JN + j); };
                if (PIPE) tmem_ldq_eo_issue(chunk_addr(0), chunk_addr(0) + L::NB, E[0][0], E[0][1], E[0][2], E[0][3]);
                uint32_t se0[4] = {0, 0, 0, 0}, se1[4] = {0, 0, 0, 0}, so0[4] = {0, 0, 0, 0}, so1[4] = {0, 0, 0, 0};
                if (!TWO) tmem_ldq_sums(tqa + NCH, tqa + L::NB + NCH, se0, se1, so0, so1);      // its wait covers chunk 0 as well
                if (PIPE) {
                    tmem_wait_eo(E[0][0], E[0][1], E[0][2], E[0][3]);
                    if (JN == 1) release_acc();
                }
                // window 0: conv rows (4w, 4w+1) = lanes (qi, qi+8) of half 0; window 1: rows (4w+2, 4w+3) = half 1
                const int n0[4] = {(int)se0[0], (int)se0[2], (int)so0[0], (int)so0[2]};
                const int n1[4] = {(int)se1[0], (int)se1[2], (int)so1[0], (int)so1[2]};
                uint32_t w0[JN], w1[JN];
#pragma unroll
                for (int j = 0; j < JN; ++j) {
                    constexpr int NOW = 0;
                    const int set = PIPE ? j & 1 : NOW;
                    if (PIPE) {
                        if (j + 1 < JN)
                            tmem_ldq_eo_issue(chunk_addr(j + 1), chunk_addr(j + 1) + L::NB, E[(j + 1) & 1][0], E[(j + 1) & 1][1], E[(j + 1) & 1][2], E[(j + 1) & 1][3]);
                    } else {
                        tmem_ldq_eo(chunk_addr(j), chunk_addr(j) + L::NB, E[NOW][0], E[NOW][1], E[NOW][2], E[NOW][3]);
                        if (j + 1 == JN) release_acc();
                    }
                    uint32_t(&e0)[8] = E[set][0];
                    uint32_t(&e1)[8] = E[set][1];
                    uint32_t(&o0)[8] = E[set][2];
                    uint32_t(&o1)[8] = E[set][3];
                    int r0[4], r1w[4];
                    uint32_t orx = 0, orr = 0, xq;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int k = 4 * j + cc, ri = 4 * (cc >> 1) + (cc & 1);
                        r0[cc] = pool_requant(ch.zw[k], ch.bias[k], ch.m2[k], ch.sh[k], zo, (int)e0[ri], (int)e0[ri + 2], (int)o0[ri], (int)o0[ri + 2], n0[0],
                                              n0[1], n0[2], n0[3], &xq);
                        orx |= xq; orr |= (uint32_t)r0[cc];
                        r1w[cc] = pool_requant(ch.zw[k], ch.bias[k], ch.m2[k], ch.sh[k], zo, (int)e1[ri], (int)e1[ri + 2], (int)o1[ri], (int)o1[ri + 2], n1[0],
                                               n1[1], n1[2], n1[3], &xq);
                        orx |= xq; orr |= (uint32_t)r1w[cc];
                    }
                    if (orx >= (1u << 22) || orr > 255u) {
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            const int k = 4 * j + cc, ri = 4 * (cc >> 1) + (cc & 1);
                            const double mcd = a.mc[ch0 + k];
                            r0[cc] = pool_requant_slow(ch.zw[k], ch.bias[k], mcd, zo, (int)e0[ri], (int)e0[ri + 2], (int)o0[ri], (int)o0[ri + 2], n0[0], n0[1],
                                                       n0[2], n0[3]);
                            r1w[cc] = pool_requant_slow(ch.zw[k], ch.bias[k], mcd, zo, (int)e1[ri], (int)e1[ri + 2], (int)o1[ri], (int)o1[ri + 2], n1[0], n1[1],
                                                        n1[2], n1[3]);
                        }
                    }
                    w0[j] = pack4(r0);
                    w1[j] = pack4(r1w);
                    if (PIPE && j + 1 < JN) {
                        tmem_wait_eo(E[(j + 1) & 1][0], E[(j + 1) & 1][1], E[(j + 1) & 1][2], E[(j + 1) & 1][3]);
                        if (j + 2 == JN) release_acc();
                    }
                }
                const int px = tx * (G::TWPX / 2) + qi;
                if (px < a.PW && !KNOB(8)) {
                    uint8_t *dst = out_tile;
                    const uint32_t rowb = out_row;
                    if (py0 < a.PH) {
                        if constexpr (JN == 1) *reinterpret_cast<uint32_t *>(dst) = w0[0];
                        else if constexpr (JN == 2) *reinterpret_cast<uint2 *>(dst) = make_uint2(w0[0], w0[1]);
                        else *reinterpret_cast<uint4 *>(dst) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
                    }
                    if (py0 + 1 < a.PH) {
                        if constexpr (JN == 1) *reinterpret_cast<uint32_t *>(dst + rowb) = w1[0];
                        else if constexpr (JN == 2) *reinterpret_cast<uint2 *>(dst + rowb) = make_uint2(w1[0], w1[1]);
                        else *reinterpret_cast<uint4 *>(dst + rowb) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    TR_EXIT();
    if (t < 32) {
        tc_fence_after();
        tmem_dealloc<L::TMEM_COLS>(tmem_base);
    }
}

struct RowsState {
    int CS, NCH;
    int slices = 1;         // launches per forward, NCH output channels each (c = 64 only)
    size_t slice_bytes = 0; // filter image bytes of one slice
    bool two = false;       // two signed weight blocks (see RowsCfg)
    uint8_t *wimg = nullptr;
    void *l0 = nullptr;     // the dense layer-0 flavour (yq_conv_tc_l0.cu) when this layer qualifies: takes the planar launches
    std::map<std::pair<const void *, int>, CUtensorMap> maps;      // input tensor map per (input pointer, batch)
};

typedef CUresult (*RowsEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// The input tile as TMA boxes.  c = 4: the padded tensor as [rows][pitch * 4 bytes], box = 18 rows x 144 bytes.
// c >= 16: [rows][pitch / 2][parity][c bytes], box = 18 rows x 9 pixels of ONE parity x 16 channels, so that the even and the
// odd pixel columns land in separate planes (what the even/odd MMA groups read).
RowsEncodeFn rows_encoder()
{
    static RowsEncodeFn enc = nullptr;
    if (!enc) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) enc = (RowsEncodeFn)p;
    }
    return enc;
}

// PLANAR: the network input as [planes][h][w bytes]; box = 3 planes x 18 rows x 64 bytes at (x0 - 16, y0 - 1, 3 n), zero outside the image
int rows_encode_planar(CUtensorMap *m, const void *in, int w, int h, int planes)
{
    RowsEncodeFn enc = rows_encoder();
    if (!enc) return yq::fail("cuTensorMapEncodeTiled is not available from this driver");
    const cuuint32_t es[3] = {1, 1, 1};
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)w, (cuuint64_t)w * h};
    const cuuint32_t box[3] = {64, (cuuint32_t)(TILE_ROWS + 2), 3};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return yq::fail("tcgen05 rows flavour: cuTensorMapEncodeTiled(planar %d x %d x %d) failed: %d", w, h, planes, (int)r);
    return 0;
}

int rows_encode(CUtensorMap *m, const void *in, int CS, int rows, int pitch)
{
    RowsEncodeFn enc = rows_encoder();
    if (!enc) return yq::fail("cuTensorMapEncodeTiled is not available from this driver");
    CUresult r;
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (CS == 4) {
        const cuuint64_t dims[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)rows};
        const cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
        const cuuint32_t box[2] = {(cuuint32_t)PLANE, (cuuint32_t)(TILE_ROWS + 2)};
        r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t dims[4] = {(cuuint64_t)CS, 2, (cuuint64_t)pitch / 2, (cuuint64_t)rows};
        const cuuint64_t strides[3] = {(cuuint64_t)CS, (cuuint64_t)2 * CS, (cuuint64_t)pitch * CS};
        const cuuint32_t box[4] = {16, 1, PLANE / 16, (cuuint32_t)(TILE_ROWS + 2)};
        r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void *>(in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return yq::fail("tcgen05 rows flavour: cuTensorMapEncodeTiled(c %d, %d rows, pitch %d) failed: %d", CS, rows, pitch, (int)r);
    return 0;
}

// byte position of element (row n, k) inside one [NB][32] filter tile: 8-row x 16-byte core matrices, the two K chunks of
// a group side by side (LBO = 128), groups 256 bytes apart (SBO = 256)
inline size_t bpos(int n, int k) { return (size_t)(n / 8) * 256 + (size_t)(k / 16) * 128 + (size_t)(n % 8) * 16 + (size_t)(k % 16); }

template <int CS, int NCH, int SPLIT, bool TWO, bool DBL, bool PLANAR = false>
int launch_rows(const CUtensorMap &tmA, const RowsArgs &a, cudaStream_t stream)
{
    using L = RowsCfg<CS, NCH, TWO, DBL, PLANAR>;
    constexpr int NT = RW_THREADS * SPLIT;
    const int smem = L::TOTAL + 128;
    auto kern = conv_u8_tc_rows_kernel<CS, NCH, SPLIT, TWO, DBL, PLANAR>;
    // per device: the shared-memory opt-in and the CTAs-per-SM count derived from this device's limits
    if (yq::ensure_dynamic_smem((const void *)kern, smem)) return -1;
    const int n_sm = yq::device_sm_count();
    int ctas_per_sm = 0;
    if (!yq::memo_get((const void *)kern, &ctas_per_sm)) {
        const int smem_sm = yq::device_smem_per_sm();
        cudaFuncAttributes fa;
        YQ_CUDA(cudaFuncGetAttributes(&fa, kern));
        // counted by hand: the occupancy API answers 1 for kernels that allocate tensor memory (see yq_conv_tc_small.cu)
        const int by_smem = smem_sm / (smem + 1024 + (int)fa.sharedSizeBytes);
        const int by_regs = 65536 / (((fa.numRegs + 7) / 8 * 8) * (NT + 32 * L::NPROD));
        const int by_tmem = 512 / L::TMEM_COLS;
        int occ = by_smem < by_regs ? by_smem : by_regs;
        if (by_tmem < occ) occ = by_tmem;
        if (getenv("YQ_DEBUG")) fprintf(stderr, "yq: rows<%d,%d,%d,%d,%d> regs=%d by_smem=%d by_regs=%d by_tmem=%d smem=%d\n", CS, NCH, SPLIT, (int)TWO, (int)DBL, fa.numRegs, by_smem, by_regs, by_tmem, smem);
        if (occ < 1 || n_sm <= 0) return yq::fail("conv_u8_tc_rows_kernel<%d,%d> does not fit on an SM", CS, NCH);
        ctas_per_sm = occ;
        yq::memo_put((const void *)kern, occ);
    }
    int grid = n_sm * ctas_per_sm;
    if (grid > a.num_tiles) grid = a.num_tiles;
    YQ_CUDA(yq::launch_pdl(conv_u8_tc_rows_kernel<CS, NCH, SPLIT, TWO, DBL, PLANAR>, dim3(grid), dim3(NT + 32 * L::NPROD), smem, stream, tmA, a));
    return 0;
}

// two signed blocks always come with two accumulators; SPLIT needs a 4-channel chunk per thread
template <int CS, int NCH>
int run_rows(bool two, bool split, bool dbl, bool planar, const CUtensorMap &tmA, const RowsArgs &a, cudaStream_t stream)
{
    constexpr bool CAN_SPLIT = CS == 4 || NCH >= 32;
    if constexpr (CS == 4) {
        if (planar) {
            if (split) {
                if (two) return launch_rows<CS, NCH, 2, true, true, true>(tmA, a, stream);
                return dbl ? launch_rows<CS, NCH, 2, false, true, true>(tmA, a, stream) : launch_rows<CS, NCH, 2, false, false, true>(tmA, a, stream);
            }
            if (two) return launch_rows<CS, NCH, 1, true, true, true>(tmA, a, stream);
            return dbl ? launch_rows<CS, NCH, 1, false, true, true>(tmA, a, stream) : launch_rows<CS, NCH, 1, false, false, true>(tmA, a, stream);
        }
    }
    if constexpr (CAN_SPLIT) {
        if (split) {
            if (two) return launch_rows<CS, NCH, 2, true, true>(tmA, a, stream);
            return dbl ? launch_rows<CS, NCH, 2, false, true>(tmA, a, stream) : launch_rows<CS, NCH, 2, false, false>(tmA, a, stream);
        }
    }
    if (two) return launch_rows<CS, NCH, 1, true, true>(tmA, a, stream);
    return dbl ? launch_rows<CS, NCH, 1, false, true>(tmA, a, stream) : launch_rows<CS, NCH, 1, false, false>(tmA, a, stream);
}

}  // namespace

int yq_tc_rows_supported(const yq_conv_layer *l)
{
    if (!l->int_form || !l->fused_mult || l->saturate || l->quant_stop_flag) return 0;
    if (l->size != 3 || l->pad != 1 || l->stride != 1) return 0;
    if (yq::act_mode(l->activation) != 0) return 0;                 // RELU6 only (pool-first needs a monotone, non-negative map)
    if ((l->h & 1) || (l->w & 1)) return 0;                         // whole 2x2 windows only
    if (l->n != l->cs_out) return 0;
    const int ci = l->cs_in, co = l->cs_out;
    // (c = 64: the filter tiles of 64 output channels fill shared memory -- n = 128 runs as two launches of 64 channels each)
    return (ci == 4 && (co == 16 || co == 32)) || (ci == 16 && (co == 16 || co == 32 || co == 64)) || (ci == 32 && (co == 32 || co == 64)) ||
           (ci == 64 && l->c == 64 && (co == 64 || co == 128));
}

void yq_tc_rows_input_geom(const yq_conv_layer *l, yq_act_geom *g)
{
    const int twpx = l->cs_in == 4 ? 32 : 16;
    g->pad = 1;
    g->pitch_w = yq::round_up(yq::round_up(l->w, twpx) + 4, 4);   // tile overhang + halo; rows start 16-byte aligned for c = 4
    g->rows_h = yq::round_up(l->h, TILE_ROWS) + 2;
}

int yq_tc_rows_prepare(yq_conv_layer *l, void **state)
{
    RowsState *st = new RowsState();
    st->CS = l->cs_in;
    st->slices = l->cs_in == 64 ? l->cs_out / 64 : 1;
    st->NCH = l->cs_out / st->slices;       // output channels per launch
    const int CS = st->CS, NCH = st->NCH, NPQ = NCH / 4;
    // two signed blocks h + l = w - zp_w unless some difference is 255 (h = 127 would leave l = 128).  By default only where the
    // kernel can then double-buffer its accumulator (RowsCfg::DB): elsewhere twice the MMAs cost more than the epilogue saves
    // (measured: layer 4 51 -> 76 us).  YQ_ROWS_TWO = 0 / 1 forces it off / on.
    const int two_env = getenv("YQ_ROWS_TWO") ? atoi(getenv("YQ_ROWS_TWO")) : -1;
    bool two = two_env >= 0 ? two_env != 0 : ((CS == 4 && NCH == 16) || (CS == 16 && NCH <= 32));
    if (CS == 64) two = false;              // (twice the filter tiles would not fit in shared memory)
    auto Wraw = [&](int oc, int ci, int ky, int kx) -> int { return l->host_w[(((size_t)oc * l->c + ci) * 3 + ky) * 3 + kx]; };
    auto zpw = [&](int oc) -> int { return l->host_chanq[(size_t)oc * 4 + 1]; };
    for (int oc = 0; oc < l->n && two; ++oc)
        for (int i = 0; i < l->c * 9; ++i)
            if ((int)l->host_w[(size_t)oc * l->c * 9 + i] - zpw(oc) == 255) { two = false; break; }
    st->two = two;
    const int nsum = two ? 0 : (CS == 4 ? 16 : 8);
    const int NB = CS == 4 ? 4 * NCH + nsum : NCH + nsum;
    const int nblk = CS == 4 ? 1 : CS / 16;
    const int nmma = CS == 4 ? 3 : 6 * nblk;
    const int parts = two ? 2 : 1;
    const size_t slice_bytes = (size_t)parts * nmma * (CS == 4 ? NB : 2 * NB) * 32;
    const size_t img_bytes = slice_bytes * st->slices;
    st->slice_bytes = slice_bytes;
    std::vector<uint8_t> img;
    char tag[24];
    snprintf(tag, sizeof tag, "rows.%d", two ? 2 : 1);
    const bool cached = yq::pack_fetch(l, tag, img) && img.size() == img_bytes;
    if (!cached) img.assign(img_bytes, 0);
    int part = 0, slice = 0;
    // the filter byte of the block being laid out (channel oc of the current slice): the u8 weight, or (two) the signed block
    // `part` of w - zp_w
    auto W = [&](int oc, int ci, int ky, int kx) -> uint8_t {
        if (ci >= l->c) return 0;
        oc += slice * NCH;
        const int w = Wraw(oc, ci, ky, kx);
        if (!two) return (uint8_t)w;
        const int d = w - zpw(oc), h = d < -128 ? -128 : d > 127 ? 127 : d;
        return (uint8_t)(int8_t)(part == 0 ? h : d - h);
    };
    for (slice = 0; slice < (cached ? 0 : st->slices); ++slice)
    for (part = 0; part < parts; ++part)
    if (CS == 4) {
        // TMEM column c = 8g + 2q + e: channel q*NPQ + g % NPQ, output pixel 2*(g / NPQ) + e of the 4-pixel segment;
        // K byte k = 4*ip + ci: input pixel ip (0..7, the segment's window starts one pixel to the left), channel ci
        for (int ky = 0; ky < 3; ++ky) {
            uint8_t *tile = img.data() + slice * slice_bytes + (size_t)(part * 3 + ky) * NB * 32;
            for (int c = 0; c < 4 * NCH; ++c) {
                const int g = c / 8, q = (c % 8) / 2, e = c % 2;
                const int oc = q * NPQ + g % NPQ, px = 2 * (g / NPQ) + e;
                for (int kx = 0; kx < 3; ++kx)
                    for (int ci = 0; ci < 4; ++ci) tile[bpos(c, 4 * (px + kx) + ci)] = W(oc, ci, ky, kx);
            }
            for (int c = 0; c < nsum; ++c) {   // activation sums: column group c/8 = pixel pair, c % 2 = pixel of the pair
                const int px = 2 * (c / 8) + c % 2;
                for (int kx = 0; kx < 3; ++kx)
                    for (int ci = 0; ci < l->c; ++ci) tile[bpos(4 * NCH + c, 4 * (px + kx) + ci)] = 1;
            }
        }
    } else {
        // TMEM column c = 8g + 2q + e of a group: channel q*NPQ + 2g + e.  Filter rows [0, NB) = even-pixel group, [NB, 2NB) =
        // odd-pixel group; K chunk 0 = plane E, chunk 1 = plane O of the MMA's step (see issue_mma)
        int m = part * nmma;
        for (int blk = 0; blk < nblk; ++blk)
            for (int ky = 0; ky < 3; ++ky)
                for (int step = 0; step < 2; ++step, ++m) {
                    uint8_t *tile = img.data() + slice * slice_bytes + (size_t)m * 2 * NB * 32;
                    // filter column kx read by [group][chunk] of this MMA (-1: nothing)
                    const int kxmap[2][2][2] = {{{1, 0}, {-1, 2}}, {{0, -1}, {2, 1}}};   // [group even/odd][step][chunk]
                    for (int grp = 0; grp < 2; ++grp)
                        for (int chunk = 0; chunk < 2; ++chunk) {
                            const int kx = kxmap[grp][step][chunk];
                            if (kx < 0) continue;
                            for (int b = 0; b < 16; ++b) {
                                const int ci = blk * 16 + b;
                                for (int c = 0; c < NCH; ++c) {
                                    const int g = c / 8, q = (c % 8) / 2, e = c % 2;
                                    tile[bpos(grp * NB + c, chunk * 16 + b)] = W(q * NPQ + 2 * g + e, ci, ky, kx);
                                }
                                if (ci < l->c)
                                    for (int c = 0; c < nsum; ++c) tile[bpos(grp * NB + NCH + c, chunk * 16 + b)] = 1;
                            }
                        }
                }
    }
    if (!cached) yq::pack_put(l, tag, img);
    if (yq::pack_fetch_device(l, tag, img.size(), (void **)&st->wimg)) {
        // (data-parallel replica: the image came from the arena blob on this device)
    } else if (cudaMalloc((void **)&st->wimg, img.size()) != cudaSuccess || cudaMemcpy(st->wimg, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(st->wimg);
        delete st;
        return yq::fail("tcgen05 rows flavour: weight upload failed");
    }
    if (yq_tc_l0_supported(l) && yq_tc_l0_prepare(l, &st->l0) != 0) {
        cudaFree(st->wimg);
        delete st;
        return -1;
    }
    *state = st;
    return 0;
}

int yq_tc_rows_two_blocks(const void *state) { return state && ((const RowsState *)state)->two ? 1 : 0; }
int yq_tc_rows_launches(const void *state) { return state ? ((const RowsState *)state)->slices : 0; }

void yq_tc_rows_free(void *state)
{
    RowsState *st = (RowsState *)state;
    if (!st) return;
    yq_tc_l0_free(st->l0);
    cudaFree(st->wimg);
    delete st;
}

int yq_tc_rows_planar_supported(const yq_conv_layer *l)
{
    // CHW planes read in place: three real channels, zero halo (TMA's out-of-bounds fill), rows a multiple of 16 bytes
    return l->tc_rows && l->c == 3 && l->cs_in == 4 && l->zp_in == 0 && l->w % 16 == 0 && l->w >= 64;
}

// planar = 0: in = the halo-padded NHWC tensor of yq_tc_rows_input_geom; planar = 1: in = the plain [batch][3][h][w] planes
int yq_tc_rows_forward(yq_conv_layer *l, void *state, const uint8_t *in_padded, uint8_t *out_pool, const yq_act_geom *og, int batch, cudaStream_t stream,
                       int planar)
{
    RowsState *st = (RowsState *)state;
    if (!st || !in_padded || !out_pool || !og) return yq::fail("tcgen05 rows flavour: bad argument");
    if (planar && !yq_tc_rows_planar_supported(l)) return yq::fail("tcgen05 rows flavour: this layer cannot read CHW planes (c = 3, zp_in = 0, w %% 16 = 0)");
    if (planar && st->l0) return yq_tc_l0_forward(l, st->l0, in_padded, out_pool, og, batch, stream);      // the dense layer-0 form
    RowsArgs a;
    memset(&a, 0, sizeof a);
    yq_act_geom ig;
    yq_tc_rows_input_geom(l, &ig);
    a.in = in_padded; a.out_pool = out_pool; a.wimg = st->wimg;
    a.HP = ig.rows_h; a.WP = ig.pitch_w; a.OH = l->out_h; a.OW = l->out_w; a.PH = l->out_h / 2; a.PW = l->out_w / 2;
    a.OHP = og->rows_h; a.OWP = og->pitch_w; a.opad = og->pad;
    const int twpx = st->CS == 4 ? 32 : 16;
    a.tiles_x = (l->out_w + twpx - 1) / twpx;
    a.tiles_y = (l->out_h + TILE_ROWS - 1) / TILE_ROWS;
    a.num_tiles = a.tiles_x * a.tiles_y * batch;
    a.zp_out = l->zp_out;
    a.knobs = getenv("YQ_ROWS_KNOBS") ? atoi(getenv("YQ_ROWS_KNOBS")) : 0;
    a.magic_x = (uint32_t)((0x100000000ull + a.tiles_x - 1) / a.tiles_x);
    a.magic_y = (uint32_t)((0x100000000ull + a.tiles_y - 1) / a.tiles_y);
    if ((unsigned long long)a.num_tiles * (a.tiles_x > a.tiles_y ? a.tiles_x : a.tiles_y) >= 0x100000000ull || yq_act_geom_bytes(og, batch, l->n) >= 0x100000000ull)
        return yq::fail("tcgen05 rows flavour: tensor too large for 32-bit tile arithmetic");
    if (((uintptr_t)in_padded & 15) || (ig.pitch_w & 3)) return yq::fail("tcgen05 rows flavour: the input must be 16-byte aligned with a pitch that is a multiple of 4");
    const auto key = std::make_pair((const void *)in_padded, planar ? -batch : batch);
    auto it = st->maps.find(key);
    if (it == st->maps.end()) {
        CUtensorMap m;
        if (planar ? rows_encode_planar(&m, in_padded, l->w, l->h, 3 * batch) : rows_encode(&m, in_padded, st->CS, ig.rows_h * batch, ig.pitch_w)) return -1;
        if (st->maps.size() > 64) st->maps.clear();
        it = st->maps.emplace(key, m).first;
    }
    const CUtensorMap &tmA = it->second;
    a.out_cs = l->cs_out;
    // form of the kernel: defaults from measurements on B200 (profiles/README.md), YQ_ROWS_SPLIT / YQ_ROWS_DB = 0 / 1 force
    static const int split_env = getenv("YQ_ROWS_SPLIT") ? atoi(getenv("YQ_ROWS_SPLIT")) : -1;
    static const int db_env = getenv("YQ_ROWS_DB") ? atoi(getenv("YQ_ROWS_DB")) : -1;
    for (int slice = 0; slice < st->slices; ++slice) {
        a.ch_off = slice * st->NCH;
        a.trace_slot = st->CS == 4 ? 0 : st->CS == 16 ? 1 : st->CS == 32 ? 2 : 3 + (slice & 1);
        a.wimg = st->wimg + slice * st->slice_bytes;
        memcpy(a.cq, l->host_chanq.data() + (size_t)a.ch_off * 4, (size_t)st->NCH * 16);
        memcpy(a.mc, l->host_mcomb.data() + a.ch_off, (size_t)st->NCH * 8);
        int rc = -2;
#define YQ_RW(CS_, N_, SPLIT_DEF_, DB_DEF_)                                                                       \
        if (st->CS == CS_ && st->NCH == N_) {                                                                    \
            const bool split = (CS_ == 4 || N_ >= 32) && (split_env < 0 ? SPLIT_DEF_ : split_env);               \
            const bool dbl = st->two || (db_env < 0 ? DB_DEF_ : db_env);                                         \
            rc = run_rows<CS_, N_>(st->two, split, dbl, planar != 0, tmA, a, stream);                            \
        }
        YQ_RW(4, 16, 0, 0); YQ_RW(4, 32, 0, 0);
        YQ_RW(16, 16, 0, 0); YQ_RW(16, 32, 0, 0); YQ_RW(16, 64, 0, 0);
        YQ_RW(32, 32, 0, 0); YQ_RW(32, 64, 0, 1);
        YQ_RW(64, 64, 0, 1);
#undef YQ_RW
        if (rc == -2) break;
        if (rc) return rc;
        if (slice + 1 == st->slices) return 0;
    }
    return yq::fail("tcgen05 rows flavour: no instantiation for cs_in=%d cs_out=%d", st->CS, st->NCH);
}


#ifdef YQ_ROWS_TRACE
extern "C" __attribute__((visibility("default"))) int yq_debug_rows_trace2(void *host, size_t bytes)
{
    return cudaMemcpyFromSymbol(host, yq_rows_trace2, bytes < sizeof(yq_rows_trace2) ? bytes : sizeof(yq_rows_trace2)) == cudaSuccess ? 0 : -1;
}
extern "C" __attribute__((visibility("default"))) int yq_debug_rows_trace(void *host, size_t bytes)
{
    return cudaMemcpyFromSymbol(host, yq_rows_trace, bytes < sizeof(yq_rows_trace) ? bytes : sizeof(yq_rows_trace)) == cudaSuccess ? 0 : -1;
}
#endif
