// yq_tc_ptx.cuh -- raw PTX wrappers for the sm_100a building blocks both tcgen05 convolution flavours use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05.alloc / mma kind::i8 / commit / ld, proxy fences.
// Instruction forms follow the PTX emitted by the CuTe headers (cute/arch/mma_sm100_umma.hpp:968-1003,
// copy_sm90_tma.hpp, tmem_allocator_sm100.hpp, copy_sm100.hpp) -- written out by hand here.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace yqtc {

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"      // (no suspend-time hint: measured on B200, a 10 ms hint as the CuTe
        "selp.u32 %0, 1, 0, p;\n\t}"                                         // pipelines pass makes the waiters wake late -- layer 0 0.087 -> 0.110 ms)
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    for (;;) {
        // four polls per look at the clock: the retry loop's compare / add / branch instructions compete with the working warps
        // for the alu pipe (a quarter of the producer warps' instructions in the layer-0 kernel were this loop)
        if (mbar_try_wait(bar, parity)) return;
        if (mbar_try_wait(bar, parity)) return;
        if (mbar_try_wait(bar, parity)) return;
        if (mbar_try_wait(bar, parity)) return;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_4d(void *smem, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// plain (1-D) bulk copy global -> shared, completion on an mbarrier; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_load(void *smem, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *smem, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem, const CUtensorMap *map, uint64_t *bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
// multicast forms: the box lands at the same shared-memory offset of every CTA in `mask` (cluster ranks) and
// completes transaction bytes on the same-offset mbarrier of each of them
__device__ __forceinline__ void tma_load_4d_mc(void *smem, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3, uint16_t mask)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5, %6, %7}], [%2], %3;" ::"r"(
            smem_u32(smem)),
        "l"(map), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void *smem, const CUtensorMap *map, uint64_t *bar, int c0, int c1, uint16_t mask)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(
                     smem_u32(smem)),
                 "l"(map), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1)
                 : "memory");
}
// tcgen05.commit that arrives on the same-offset mbarrier of every CTA in `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, const void *smem, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(smem_u32(smem)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *smem, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(smem)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// Ampere-style asynchronous global->shared copies (LDGSTS), L1-cached (.ca: neighbouring pixels share taps)
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void st_shared32(uint32_t dst, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst), "r"(v) : "memory"); }
__device__ __forceinline__ void st_shared128(uint32_t dst, uint4 v)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One lane of a CONVERGED warp (all 32 lanes must execute this).  Keeping the single-thread tcgen05 / TMA issue inside
// warp-uniform control flow lets the compiler keep descriptors in uniform registers; under `if (lane == 0)` it wraps
// every UTCIMMA in an ELECT / R2UR / BRA.U.ANY loop and the issuing thread, not the tensor pipe, becomes the limit.
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *slot)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"   // same asm block: the registers are only defined after the wait
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
// split form for software pipelining: issue the load, do other work, then wait.  The wait names the registers as
// read-write operands so that no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&v)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]));
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr)
{
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n\ttcgen05.wait::ld.sync.aligned;" : "=r"(v) : "r"(taddr));
    return v;
}

// 16x256b shape (cute Copy_Traits<SM100_TMEM_LOAD_16dp256b{1,2}x>): a warp reads 16 TMEM lanes starting at the lane in
// taddr (its own quarter, +0 or +16); thread t gets lane t/4 in {r0, r1} = columns 2(t%4), 2(t%4)+1 and lane t/4 + 8 in
// {r2, r3}; .x2 repeats the pattern 8 columns further in {r4..r7}.  These helpers read BOTH 16-lane halves of the
// quarter, so thread t ends up with lanes {t/4, t/4+8} (v0) and {t/4+16, t/4+24} (v1).
__device__ __forceinline__ void tmem_ldq(uint32_t taddr, uint32_t (&v0)[8], uint32_t (&v1)[8])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%16];\n\t"
        "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%8, %9, %10, %11, %12, %13, %14, %15}, [%17];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(v0[0]), "=r"(v0[1]), "=r"(v0[2]), "=r"(v0[3]), "=r"(v0[4]), "=r"(v0[5]), "=r"(v0[6]), "=r"(v0[7]), "=r"(v1[0]), "=r"(v1[1]),
          "=r"(v1[2]), "=r"(v1[3]), "=r"(v1[4]), "=r"(v1[5]), "=r"(v1[6]), "=r"(v1[7])
        : "r"(taddr), "r"(taddr + (16u << 16)));
}
// the same plus 8 columns at tsum (.x1) into s0 / s1 (the activation-sum columns), one wait for all four loads
__device__ __forceinline__ void tmem_ldq_first(uint32_t tsum, uint32_t taddr, uint32_t (&s0)[4], uint32_t (&s1)[4], uint32_t (&v0)[8],
                                               uint32_t (&v1)[8])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%24];\n\t"
        "tcgen05.ld.sync.aligned.16x256b.x1.b32 {%4, %5, %6, %7}, [%25];\n\t"
        "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%8, %9, %10, %11, %12, %13, %14, %15}, [%26];\n\t"
        "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%16, %17, %18, %19, %20, %21, %22, %23}, [%27];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(s0[0]), "=r"(s0[1]), "=r"(s0[2]), "=r"(s0[3]), "=r"(s1[0]), "=r"(s1[1]), "=r"(s1[2]), "=r"(s1[3]), "=r"(v0[0]), "=r"(v0[1]),
          "=r"(v0[2]), "=r"(v0[3]), "=r"(v0[4]), "=r"(v0[5]), "=r"(v0[6]), "=r"(v0[7]), "=r"(v1[0]), "=r"(v1[1]), "=r"(v1[2]), "=r"(v1[3]),
          "=r"(v1[4]), "=r"(v1[5]), "=r"(v1[6]), "=r"(v1[7])
        : "r"(tsum), "r"(tsum + (16u << 16)), "r"(taddr), "r"(taddr + (16u << 16)));
}

// ---------------------------------------------------------------------------------------------
// cta_group::2 (CTA pair of a cluster; forms follow cute/arch/{mma_sm100_umma,copy_sm100_tma,tmem_allocator_sm100}.hpp)
// ---------------------------------------------------------------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t *slot)   // one warp of EACH CTA of the pair, same slot offset
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t addr)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
// M = 256 (128 rows from each CTA's A tile), B rows split between the two CTAs' tiles; issued by the leader CTA only
__device__ __forceinline__ void umma_i8_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// arrives (once the pair's prior MMAs have completed) on the same-offset mbarrier of every CTA in `mask`
__device__ __forceinline__ void umma_commit_2cta(uint64_t *bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
// shared::cluster address of `p` (an address of MY shared memory) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void *p, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-D TMA load into MY shared memory whose completion is signalled on an mbarrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_2cta(void *smem, const CUtensorMap *map, uint32_t bar_cluster_addr, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(smem)),
                 "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
                 : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_m(int m, int n) { return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

// kind::i8 instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 (2) [4,6), a/b format 0 = UINT8
// [7,10)/[10,13), a/b K-major (0) [15]/[16], N>>3 [17,23), M>>4 [24,29); M is always 128 here.
__host__ __device__ constexpr uint32_t make_idesc(int n) { return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout type [61,64): 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B.
// ROWB = bytes per row of the swizzle atom (128 / 64 / 32); SBO = 8 rows.
template <int ROWB>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    constexpr uint64_t layout = ROWB == 128 ? 2 : (ROWB == 64 ? 4 : 6);
    constexpr uint64_t sbo = (8 * ROWB) >> 4;
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

}  // namespace yqtc
