// K3 (tensor-core variant, HVPR_MEM_BF16_RESCORE) — placeholder until the tcgen05 kernel lands:
// reports HVPR_ERR_UNSUPPORTED so callers fail loudly instead of silently using another path.
#include "common.cuh"
#include <cuda_bf16.h>

namespace hvpr {
__global__ void pack_bf16_kernel(const float *__restrict__ W, int M, int C, int Mpad, __nv_bfloat16 *__restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)Mpad * C) return;
    int r = (int)(i / C);
    out[i] = __float2bfloat16_rn(r < M ? W[i] : 0.0f);
}
}  // namespace hvpr
using namespace hvpr;

int hvpr_mem_attn_tc_init() { return HVPR_OK; }
size_t hvpr_mem_attn_tc_workspace_bytes(int64_t, int) { return 0; }
int hvpr_mem_pack_bf16_impl(const float *W, int M, int C, void *out, cudaStream_t stream) {
    int Mpad = (M + 255) / 256 * 256;
    int64_t n = (int64_t)Mpad * C;
    pack_bf16_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, stream>>>(W, M, C, Mpad, (__nv_bfloat16 *)out);
    HVPR_CHECK_LAUNCH();
    return HVPR_OK;
}
int hvpr_mem_attn_tc(const float *, const int32_t *, int64_t, const float *, const void *, int, int, int, float *,
                     int32_t *, void *, size_t, cudaStream_t) {
    return HVPR_ERR_UNSUPPORTED;
}
