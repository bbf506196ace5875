// probe: tcgen05.mma (kind::i8, M=128, K=32, SS) rate while bulk copies stream global -> shared memory into the same SM,
// i.e. how much the tensor core's operand reads and the copy engine's writes compete for shared-memory bandwidth.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../yolo_quantization_b200/csrc/yq_tc_ptx.cuh"
using namespace yqtc;

__device__ __forceinline__ void bulk_g2s(void *smem, const void *g, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)), "l"(g), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(int N, int iters, int ring, int chunk, int accmask, int tilemask, int commit_every, const uint8_t *src, size_t src_bytes, long long *out)
{
    extern __shared__ __align__(16) uint8_t raw[];
    uint8_t *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar, cbar[8];
    __shared__ uint32_t slot;
    __shared__ volatile int done;
    const int t = threadIdx.x;
    for (int i = t; i < 96 * 1024 / 4; i += 128) ((uint32_t *)smem)[i] = 0x01010101u;
    if (t == 0) {
        mbar_init(&bar, 1);
        for (int i = 0; i < 8; ++i) mbar_init(&cbar[i], 1);
        done = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (t < 32) tmem_alloc<512>(&slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (t < 32) {
        const uint32_t idesc = (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 32 * 1024;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (elect_one()) {
                const uint64_t da = make_desc<128>(a0 + (i & tilemask) * 16384), db = make_desc<128>(b0 + (i & tilemask) * 32768);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_i8(tm + (i & accmask) * 256, da + 2 * k, db + 2 * k, idesc, 1u);
                if (commit_every && (i % commit_every) == 0) umma_commit(&cbar[7]);
            }
        }
        if (elect_one()) umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        done = 1;
        if (t == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    } else if (t == 32 && ring > 0) {
        // copy engine load: `ring` bulk copies of `chunk` bytes in flight into a scratch region above the operands
        uint8_t *scratch = smem + 96 * 1024;
        long long bytes = 0;
        const size_t base = ((size_t)blockIdx.x * 7919 * 4096) % (src_bytes / 2);
        uint32_t issued = 0, waited = 0;
        for (int i = 0; i < ring; ++i, ++issued) {
            mbar_expect_tx(&cbar[i], chunk);
            bulk_g2s(scratch + i * chunk, src + base + ((size_t)issued * chunk) % (src_bytes / 2), chunk, &cbar[i]);
        }
        while (!done) {
            const int s = waited % ring;
            mbar_wait(&cbar[s], (waited / ring) & 1);
            ++waited;
            bytes += chunk;
            mbar_expect_tx(&cbar[s], chunk);
            bulk_g2s(scratch + s * chunk, src + base + ((size_t)issued * chunk) % (src_bytes / 2), chunk, &cbar[s]);
            ++issued;
        }
        for (; waited < issued; ++waited) mbar_wait(&cbar[waited % ring], (waited / ring) & 1);
        if (blockIdx.x == 0) out[1] = bytes;
    }
    tc_fence_before();
    __syncthreads();
    if (t < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

int main()
{
    long long *d; cudaMalloc(&d, 16);
    uint8_t *src; const size_t src_bytes = 64u << 20; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    const int iters = 4000;
    for (int N : {16, 48, 64, 80, 96, 112, 128, 144, 160, 192, 256})
        for (int ce : {0, 1}) {
            cudaMemset(d, 0, 16);
            probe<<<148, 128, 225 * 1024>>>(N, iters, 0, 16384, 1, 1, ce, src, src_bytes, d);
            long long h[2] = {0, 0};
            cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("N %3d, commit per 4 MMAs %d: %s  %.1f clk per MMA (N/2 = %d)\n", N, ce, cudaGetErrorString(e), (double)h[0] / (iters * 4.0), N / 2);
        }
    return 0;
}
