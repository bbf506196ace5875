// probe: cycles per tcgen05.mma kind::i8 (M=128, K=32, SS) as a function of N, of the number of accumulators the stream
// alternates between, and of how many distinct smem operand tiles it cycles through.  One CTA per SM, one issuing warp.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../yolo_quantization_b200/csrc/yq_tc_ptx.cuh"
using namespace yqtc;

__global__ void __launch_bounds__(128, 1) probe(int N, int nacc, int ntiles, int iters, long long *out)
{
    extern __shared__ __align__(16) uint8_t raw[];
    uint8_t *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int t = threadIdx.x;
    for (int i = t; i < 160 * 1024 / 4; i += 128) ((uint32_t *)smem)[i] = 0x01010101u;
    if (t == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (t < 32) tmem_alloc<512>(&slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    long long t0 = 0, t1 = 0;
    if (t < 32) {
        const uint32_t idesc = (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 64 * 1024;
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (elect_one()) {
                const int tile = i % ntiles;
                const uint64_t da = make_desc<128>(a0 + tile * 16384), db = make_desc<128>(b0 + tile * 32768 / 2);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_i8(tm + (i % nacc) * 256, da + 2 * k, db + 2 * k, idesc, 1u);
            }
        }
        if (elect_one()) umma_commit(&bar);
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
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    for (int ctas : {1, 148})
        for (int N : {64, 128, 144, 256})
            for (int nacc : {1, 2})
                for (int ntiles : {1, 4}) {
                    probe<<<ctas, 128, 200 * 1024>>>(N, nacc, ntiles, iters, d);
                    long long h = 0;
                    cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                    printf("ctas %3d N %3d acc %d tiles %d: %s  %.1f clk per MMA (floor N/2 = %d)\n", ctas, N, nacc, ntiles, cudaGetErrorString(e),
                           (double)h / (iters * 4.0), N / 2);
                }
    return 0;
}
