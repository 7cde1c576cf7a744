// C-ABI glue: status strings, one-time init, memory-attention dispatch.
#include "common.cuh"
#include <string.h>

int hvpr_pfn_init();
int hvpr_mem_attn_fp32_init();
int hvpr_mem_attn_fp32(const float *, const int32_t *, int64_t, const float *, int, int, int, float *, int32_t *, cudaStream_t);
int hvpr_mem_attn_tc_init();
size_t hvpr_mem_attn_tc_workspace_bytes(int64_t n_rows_max, int M);
int hvpr_mem_attn_tc(const float *, const int32_t *, int64_t, const float *, const void *, int, int, int, float *,
                     int32_t *, void *, size_t, const HvprZeroFill *, cudaStream_t);
int hvpr_conv_init();
int hvpr_mem_train_init();
int hvpr_mem_pack_bf16_impl(const float *, int, int, void *, cudaStream_t);

namespace hvpr {
int num_sms() {
    static int cached[64] = {0};                 // 0 = not queried yet; racing first calls write the same value
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { (void)cudaGetLastError(); return kMaxSMs; }
    int n = cached[dev];
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { (void)cudaGetLastError(); return kMaxSMs; }
        if (n > kMaxSMs) n = kMaxSMs;            // workspaces are sized for kMaxSMs persistent blocks
        cached[dev] = n;
    }
    return n;
}
static thread_local char g_cuda_err[256] = "";
void set_cuda_error(cudaError_t e) {
    strncpy(g_cuda_err, cudaGetErrorString(e), sizeof(g_cuda_err) - 1);
    g_cuda_err[sizeof(g_cuda_err) - 1] = 0;
}
}  // namespace hvpr

extern "C" const char *hvpr_strerror(int status) {
    switch (status) {
        case HVPR_OK: return "ok";
        case HVPR_ERR_ARG: return "invalid argument";
        case HVPR_ERR_UNSUPPORTED: return "unsupported configuration";
        case HVPR_ERR_WORKSPACE: return "workspace too small";
        case HVPR_ERR_CUDA: return "CUDA error (see hvpr_last_cuda_error)";
        default: return "unknown status";
    }
}
extern "C" const char *hvpr_last_cuda_error(void) { return hvpr::g_cuda_err; }
extern "C" int hvpr_version(void) { return 100; }

extern "C" int hvpr_init(void) {
    int s;
    if ((s = hvpr_pfn_init()) != HVPR_OK) return s;
    if ((s = hvpr_mem_attn_fp32_init()) != HVPR_OK) return s;
    if ((s = hvpr_mem_attn_tc_init()) != HVPR_OK) return s;
    if ((s = hvpr_conv_init()) != HVPR_OK) return s;
    if ((s = hvpr_mem_train_init()) != HVPR_OK) return s;
    return HVPR_OK;
}

extern "C" size_t hvpr_mem_attn_workspace_bytes(int64_t n_rows_max, int M, int precision_mode) {
    if (precision_mode == HVPR_MEM_BF16_RESCORE) return hvpr_mem_attn_tc_workspace_bytes(n_rows_max, M);
    return 0;
}

extern "C" int hvpr_mem_pack_bf16(const float *mem_weight, int M, int C, void *mem_weight_bf16, void *stream) {
    if (!mem_weight || !mem_weight_bf16 || M <= 0 || C <= 0) return HVPR_ERR_ARG;
    return hvpr_mem_pack_bf16_impl(mem_weight, M, C, mem_weight_bf16, (cudaStream_t)stream);
}

extern "C" int hvpr_mem_attn(const float *pillars, const int32_t *n_pillars_dev, int64_t n_rows_max,
                             const float *mem_weight, const void *mem_weight_bf16, int M, int C, int k,
                             int precision_mode, float *readout, int32_t *topk_idx_out, void *workspace,
                             size_t workspace_bytes, const HvprZeroFill *zero_fill, void *stream) {
    if (n_rows_max < 0 || !mem_weight) return HVPR_ERR_ARG;
    if (zero_fill) {
        if (zero_fill->n < 0 || zero_fill->n > 4) return HVPR_ERR_ARG;
        for (int r = 0; r < zero_fill->n; ++r)
            if (zero_fill->bytes[r] && (!zero_fill->ptr[r] || ((uintptr_t)zero_fill->ptr[r] | zero_fill->bytes[r]) % 16)) return HVPR_ERR_ARG;
    }
    // the tcgen05 kernel zeroes the ranges itself, beside its own work; every other path does it with plain memsets
    auto memset_ranges = [&]() -> int {
        if (zero_fill)
            for (int r = 0; r < zero_fill->n; ++r)
                if (zero_fill->bytes[r]) HVPR_CHECK_CUDA(cudaMemsetAsync(zero_fill->ptr[r], 0, zero_fill->bytes[r], (cudaStream_t)stream));
        return HVPR_OK;
    };
    if (n_rows_max == 0) return memset_ranges();
    if (!pillars || !readout) return HVPR_ERR_ARG;
    if (precision_mode == HVPR_MEM_FP32) {
        const int s = memset_ranges();
        if (s != HVPR_OK) return s;
        return hvpr_mem_attn_fp32(pillars, n_pillars_dev, n_rows_max, mem_weight, M, C, k, readout, topk_idx_out,
                                  (cudaStream_t)stream);
    }
    if (precision_mode == HVPR_MEM_BF16_RESCORE) {
        if (!mem_weight_bf16) return HVPR_ERR_ARG;
        return hvpr_mem_attn_tc(pillars, n_pillars_dev, n_rows_max, mem_weight, mem_weight_bf16, M, C, k, readout,
                                topk_idx_out, workspace, workspace_bytes, zero_fill, (cudaStream_t)stream);
    }
    return HVPR_ERR_ARG;
}

// ---- debug helpers (not part of the public header): raw runtime memset / D2D copy, used to probe copy-engine overlap
extern "C" int hvpr_dbg_memset_async(void *dst, int value, size_t bytes, void *stream) {
    HVPR_CHECK_CUDA(cudaMemsetAsync(dst, value, bytes, (cudaStream_t)stream));
    return HVPR_OK;
}
extern "C" int hvpr_dbg_memcpy_d2d_async(void *dst, const void *src, size_t bytes, void *stream) {
    HVPR_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return HVPR_OK;
}
