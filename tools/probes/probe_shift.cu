// probe: does a SWIZZLE_128B K-major UMMA descriptor whose start address is shifted by whole 128-byte rows read the
// rows one expects (a) with base_offset = (start >> 7) & 7, (b) with base_offset = 0 ?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../yolo_quantization_b200/csrc/yq_tc_ptx.cuh"
using namespace yqtc;

__device__ __forceinline__ uint64_t desc128(uint32_t saddr, uint32_t base_off)
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | ((uint64_t)(base_off & 7) << 49) | (2ull << 61);
}

__global__ void probe(int shift, int use_bo, int *out)
{
    extern __shared__ __align__(16) uint8_t raw[];
    uint8_t *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint8_t *A = smem;                 // 160 rows x 128 B
    uint8_t *B = smem + 160 * 128;     // 16 rows x 128 B   (20480 = 20 * 1024: aligned)
    uint64_t *bar = (uint64_t *)(B + 16 * 128);
    uint32_t *slot = (uint32_t *)(bar + 1);
    const int t = threadIdx.x;
    for (int i = t; i < 160 * 128; i += 128) {
        int row = i / 128, col = i % 128;
        uint8_t v = (uint8_t)((row * 7 + col * 13) & 0xff);
        A[row * 128 + (((col / 16) ^ (row % 8)) * 16) + col % 16] = v;
    }
    for (int i = t; i < 16 * 128; i += 128) {
        int row = i / 128, col = i % 128;
        uint8_t v = (col == row) ? 1 : 0;
        B[row * 128 + (((col / 16) ^ (row % 8)) * 16) + col % 16] = v;
    }
    if (t == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (t < 32) tmem_alloc<32>(slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tm = *slot;
    if (t == 0) {
        uint32_t sa = smem_u32(A) + shift * 128;
        umma_i8(tm, desc128(sa, use_bo ? ((sa >> 7) & 7) : 0), desc128(smem_u32(B), 0), make_idesc(16), 0);
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    uint32_t v[16];
    tmem_ld16(tm + ((uint32_t)((t / 32) * 32) << 16), v);
    int bad = 0;
    for (int n = 0; n < 16; ++n) {
        int expect = ((t + shift) * 7 + n * 13) & 0xff;
        if ((int)v[n] != expect) ++bad;
    }
    out[t] = bad;
    if (t == 0) { out[128] = v[0]; out[129] = v[1]; }
    tc_fence_before();
    __syncthreads();
    if (t < 32) { tc_fence_after(); tmem_dealloc<32>(tm); }
}

int main()
{
    int *d; cudaMalloc(&d, 130 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    for (int bo = 0; bo < 2; ++bo)
        for (int shift = 0; shift <= 17; ++shift) {
            probe<<<1, 128, 40000>>>(shift, bo, d);
            int h[130];
            cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
            int rows_bad = 0;
            for (int i = 0; i < 128; ++i) rows_bad += h[i] != 0;
            printf("base_offset %s shift %2d: %s rows_bad=%d (v0=%d v1=%d)\n", bo ? "set" : "0  ", shift, cudaGetErrorString(e), rows_bad, h[128], h[129]);
        }
    return 0;
}
