// yq_conv_tc.cu -- tcgen05 (kind::i8) implicit-GEMM convolution flavour.  (placeholder: not yet enabled)
#include "yq_common.h"

int yq_tc_supported(const yq_conv_layer *) { return 0; }
int yq_tc_prepare(yq_conv_layer *) { return -1; }
void yq_tc_free(yq_conv_layer *) {}
int yq_tc_forward(yq_conv_layer *, const uint8_t *, uint8_t *, float *, int32_t *, int, cudaStream_t)
{
    return yq::fail("tcgen05 flavour not built");
}
