// probe: cycles per tcgen05.mma kind::i8 (M = 128, K = 32, SS) with NO-SWIZZLE K-major descriptors of the geometries the rows
// kernel uses: A = overlapping windows over 144-byte rows (LBO 16 or 2688, SBO 144) against aligned forms, B = the canonical
// 8-row x 16-byte core-matrix tiles (LBO 128, SBO 256).  One CTA per SM, one issuing thread, B tiles cycled like the kernel does.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../yolo_quantization_b200/csrc/yq_tc_ptx.cuh"
using namespace yqtc;

__device__ __forceinline__ uint64_t desc_ns(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(128, 1) probe(int N, uint32_t lbo, uint32_t sbo, uint32_t astep, int nb, int iters, long long *out)
{
    extern __shared__ __align__(16) uint8_t raw[];
    uint8_t *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int t = threadIdx.x;
    for (int i = t; i < 200 * 1024 / 4; i += 128) ((uint32_t *)smem)[i] = 0x01010101u;
    if (t == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (t < 32) tmem_alloc<512>(&slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    long long t0 = 0, t1 = 0;
    if (t == 0) {
        const uint32_t idesc = (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 64 * 1024;
        const uint32_t bsub = (uint32_t)N * 32;
        uint64_t da[12], db[12];      // descriptors in registers: the timed loop is nothing but MMA issues
#pragma unroll
        for (int m = 0; m < 12; ++m) {
            da[m] = desc_ns(a0 + (m % 6) * astep, lbo, sbo);
            db[m] = desc_ns(b0 + (m % nb) * bsub, 128, 256);
        }
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int m = 0; m < 12; ++m) umma_i8(tm, da[m], db[m], idesc, 1u);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        t1 = clock64();
    }
    if (t == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (t < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

int main()
{
    long long *d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    const int iters = 500;
    struct Cfg { const char *name; uint32_t lbo, sbo, astep; } cfgs[] = {
        {"rows c>=16: LBO 2688 (E->O plane), SBO 144, windows step 144/16", 2688, 144, 160},
        {"             LBO 2688, SBO 128 (aligned rows), step 128", 2688, 128, 128},
        {"             LBO 2688, SBO 256, step 256", 2688, 256, 256},
        {"rows c=4:    LBO 16 (overlapping K chunks), SBO 144, step 144", 16, 144, 144},
        {"             LBO 16, SBO 128, step 128", 16, 128, 128},
        {"canonical:   LBO 128, SBO 256, step 4096", 128, 256, 4096},
    };
    for (int N : {64, 144, 160})
        for (int nb : {1, 12})
            for (auto &c : cfgs) {
                probe<<<148, 128, 220 * 1024>>>(N, c.lbo, c.sbo, c.astep, nb, iters, d);
                long long h = 0;
                cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                printf("N %3d, %2d filter tiles, A %-66s: %s %.1f clk per MMA\n", N, nb, c.name, e == cudaSuccess ? "" : cudaGetErrorString(e), (double)h / (iters * 12.0));
            }
    return 0;
}
