// probe: the chip's sustained tcgen05.mma kind::i8 rate (TOP/s, wall clock by CUDA events), all SMs issuing back to back.
//   cta_group::1  M = 128, N = 256, K = 32 per MMA, one CTA per SM           (what the flat / flat2 / per-tap kernels issue)
//   cta_group::2  M = 256, N = 256, K = 32 per MMA, one CTA pair per TPC     (what flat2x issues)
// Operands are 0x01 bytes in shared memory (SWIZZLE_128B K-major descriptors), accumulators alternate between two TMEM
// column blocks.  Prints one JSON line: {"int8_tops_cta1": ..., "int8_tops_cta2": ..., "sm_count": ..., "ms": ...}.
// This is the denominator bench.py uses for `roofline` of tensor-bound launches (written to profiles/ by tools/round_check.sh).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../yolo_quantization_b200/csrc/yq_tc_ptx.cuh"
using namespace yqtc;

template <int PAIR>
__global__ void __launch_bounds__(128, 1) peak(int iters)
{
    extern __shared__ __align__(16) uint8_t raw[];
    uint8_t *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int t = threadIdx.x;
    for (int i = t; i < 96 * 1024 / 4; i += 128) ((uint32_t *)smem)[i] = 0x01010101u;
    if (t == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (t < 32) {
        if (PAIR) tmem_alloc_2cta<512>(&slot);
        else tmem_alloc<512>(&slot);
    }
    fence_proxy_async();
    tc_fence_before();
    if (PAIR) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    const bool leader = !PAIR || cluster_ctarank() == 0;
    if (t < 32 && leader) {
        // cta_group::1: A = 128 rows x 128 B, B = 256 rows x 128 B.  cta_group::2: each CTA supplies 128 A rows and 128 of the 256 B rows.
        const uint32_t idesc = PAIR ? make_idesc_m(256, 256) : make_idesc_m(128, 256);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 32 * 1024;
        for (int i = 0; i < iters; ++i) {
            if (elect_one()) {
                const uint64_t da = make_desc<128>(a0 + (i & 1) * 16384), db = make_desc<128>(b0 + (i & 1) * 32768);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (PAIR) umma_i8_2cta(tm + (i & 1) * 256, da + 2 * k, db + 2 * k, idesc, 1u);
                    else umma_i8(tm + (i & 1) * 256, da + 2 * k, db + 2 * k, idesc, 1u);
                }
            }
        }
        if (elect_one()) {
            if (PAIR) umma_commit_2cta(&bar, 1);
            else umma_commit(&bar);
        }
        mbar_wait(&bar, 0);
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();
    else __syncthreads();
    if (t < 32) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_2cta<512>(tm);
        else tmem_dealloc<512>(tm);
    }
}

template <int PAIR>
static double run(int ctas, int iters)
{
    auto kern = peak<PAIR>;
    const int smem = 100 * 1024;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = PAIR ? 2 : 1;
    at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        cudaLaunchKernelEx(&cfg, kern, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) {
            fprintf(stderr, "probe_i8_peak: %s\n", cudaGetErrorString(cudaGetLastError()));
            return -1;
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, iters = 40000;
    const double ms1 = run<0>(sms, iters);
    const double ms2 = run<1>(sms & ~1, iters);
    // ops per MMA = 2 * M * N * K
    const double ops1 = (double)sms * iters * 4 * 2.0 * 128 * 256 * 32;
    const double ops2 = (double)(sms / 2) * iters * 4 * 2.0 * 256 * 256 * 32;
    printf("{\"int8_tops_cta1\": %.1f, \"int8_tops_cta2\": %.1f, \"sm_count\": %d, \"ms_cta1\": %.3f, \"ms_cta2\": %.3f, \"iters\": %d, "
           "\"what\": \"tcgen05.mma kind::i8 SS, N=256 K=32, M=128 (cta_group::1) / M=256 (cta_group::2), all SMs, best of 4 timed launches\"}\n",
           ops1 / (ms1 * 1e-3) / 1e12, ops2 / (ms2 * 1e-3) / 1e12, sms, ms1, ms2, iters);
    return 0;
}
