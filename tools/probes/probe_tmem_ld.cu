// probe: tcgen05.ld throughput per SM by shape (32x32b vs 16x256b), by how many warps load at once (4 = one CTA's four lane
// quarters, 8 / 16 = several co-resident CTAs), with the wait after every load or after a batch.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../yolo_quantization_b200/csrc/yq_tc_ptx.cuh"
using namespace yqtc;

template <int SHAPE>   // 0: 32x32b.x16 (2 KB per warp)   1: 2 x 16x256b.x2 = both halves (2 KB per warp)   2: 32x32b.x32 (4 KB)  3: 2 x 16x256b.x4 (4 KB)
__device__ __forceinline__ uint32_t do_ld(uint32_t ta)
{
    uint32_t acc = 0;
    if (SHAPE == 0) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\ttcgen05.wait::ld.sync.aligned;"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                       "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(ta));
#pragma unroll
        for (int i = 0; i < 16; ++i) acc ^= v[i];
    } else if (SHAPE == 1) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%16];\n\t"
                     "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%17];\n\ttcgen05.wait::ld.sync.aligned;"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                       "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(ta), "r"(ta + (16u << 16)));
#pragma unroll
        for (int i = 0; i < 16; ++i) acc ^= v[i];
    } else if (SHAPE == 2) {
        uint32_t v[32];
        tmem_ld32(ta, v);   // includes the wait
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[i];
    } else {
        uint32_t v[32];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%32];\n\t"
                     "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%33];\n\t"
                     "tcgen05.wait::ld.sync.aligned;"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                       "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                       "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                       "=r"(v[31])
                     : "r"(ta), "r"(ta + (16u << 16)));
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[i];
    }
    return acc;
}

template <int SHAPE>
__global__ void __launch_bounds__(128, 4) probe(int iters, long long *out, uint32_t *sink)
{
    __shared__ uint32_t slot;
    const int t = threadIdx.x, warp = t >> 5;
    if (t < 32) tmem_alloc<128>(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tq = slot + ((uint32_t)(warp * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) acc ^= do_ld<SHAPE>(tq + (i & 1) * 32);
    const long long t1 = clock64();
    if (acc == 0x12345678u) sink[0] = acc;
    if (t == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (t < 32) { tc_fence_after(); tmem_dealloc<128>(slot); }
}

template <int SHAPE>
void run(const char *name, int bytes_per_warp, long long *d, uint32_t *sink)
{
    const int iters = 4000;
    for (int per_sm : {1, 2, 4}) {
        probe<SHAPE><<<148 * per_sm, 128>>>(iters, d, sink);
        long long h = 0;
        cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        const double clk = (double)h / iters;
        printf("%-22s %d CTAs/SM (%2d warps): %s  %.1f clk per load+wait, %.1f B/clk/SM\n", name, per_sm, 4 * per_sm, cudaGetErrorString(e), clk,
               bytes_per_warp * 4.0 * per_sm / clk);
    }
}

int main()
{
    long long *d; cudaMalloc(&d, 8);
    uint32_t *sink; cudaMalloc(&sink, 4);
    run<0>("32x32b.x16", 2048, d, sink);
    run<1>("2 x 16x256b.x2", 2048, d, sink);
    run<2>("32x32b.x32", 4096, d, sink);
    run<3>("2 x 16x256b.x4", 4096, d, sink);
    return 0;
}
