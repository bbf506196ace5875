// probe: which 3-D TMA box / coordinate combinations work on a small u8 [planes][h][w] tensor (one config per process:
// an illegal instruction poisons the context).  usage: probe_tma3d box0 box1 box2 c0 c1 c2 [w h planes]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../yolo_quantization_b200/csrc/yq_tc_ptx.cuh"
using namespace yqtc;

__global__ void k(const __grid_constant__ CUtensorMap tm, int bytes, int c0, int c1, int c2, uint32_t *out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&bar, (uint32_t)bytes);
        tma_load_3d(smem, &tm, &bar, c0, c1, c2);
        mbar_wait(&bar, 0);
        uint32_t s = 0;
        for (int i = 0; i < bytes; ++i) s += smem[i];
        out[0] = s;
    }
}

int main(int argc, char **argv)
{
    int b0 = atoi(argv[1]), b1 = atoi(argv[2]), b2 = atoi(argv[3]), c0 = atoi(argv[4]), c1 = atoi(argv[5]), c2 = atoi(argv[6]);
    int w = argc > 7 ? atoi(argv[7]) : 64, h = argc > 8 ? atoi(argv[8]) : 48, planes = argc > 9 ? atoi(argv[9]) : 6;
    uint8_t *d; cudaMalloc(&d, (size_t)w * h * planes); cudaMemset(d, 1, (size_t)w * h * planes);
    uint32_t *out; cudaMalloc(&out, 4);
    typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                           const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)planes}, strides[2] = {(cuuint64_t)w, (cuuint64_t)w * h};
    cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2}, es[3] = {1, 1, 1};
    CUresult r = ((Fn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %d %d %d at %d %d %d on %d x %d x %d: encode %d, ", b0, b1, b2, c0, c1, c2, w, h, planes, (int)r);
    k<<<1, 32, 16384>>>(m, b0 * b1 * b2, c0, c1, c2, out);
    uint32_t s = 0;
    cudaError_t e = cudaMemcpy(&s, out, 4, cudaMemcpyDeviceToHost);
    printf("%s, sum of the box = %u (in-bounds bytes are 1)\n", cudaGetErrorString(e), s);
    return 0;
}
