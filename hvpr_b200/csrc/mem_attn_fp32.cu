// K3 (exact-fp32 variant, HVPR_MEM_FP32) — MemoryUnit_Agg.forward eval branch
// (pcdet/models/backbones_2d/map_to_bev/memory_module.py:60-77):
//   logits = pillars @ W^T (:64) ; softmax over M is monotone so top-k on the logits == top-k on the scores (:65-66) ;
//   gather the k items (:67) ; a = softmax_k(<item, pillar>) (:70-72) ; out = sum_k a_k item_k (:73-74).
// The (rows, M) score matrix the reference materialises (`att`, never read: pointpillar_scatter.py:201,212) stays
// in shared memory.  This SIMT kernel is the precision reference for the tcgen05 variant in mem_attn_tc.cu.
#include "common.cuh"
#include <math.h>

namespace hvpr {

constexpr int kMaRows = 16;       // pillar rows per block
constexpr int kMaThreads = 256;
constexpr int kMaChunk = 2048;    // logits of one column chunk kept in shared memory: 16 x 2048 fp32 = 128 KB
constexpr int kMaC = 64;

struct MaSmem {
    float p[kMaRows][kMaC];
    float logit[kMaRows][kMaChunk];
};

// Any M: the columns are walked in chunks of kMaChunk; every row keeps its running top-k as a sorted list across the lanes of
// its warp (lane i = i-th largest so far) and a chunk only contributes while its maximum still beats the k-th entry.
// Ties: the lower index wins (inside a chunk by the argmax rule, across chunks because a later equal value is not inserted).
__global__ void __launch_bounds__(kMaThreads) mem_attn_fp32_kernel(const float *__restrict__ pillars,
                                                                   const int32_t *__restrict__ n_pillars_dev,
                                                                   int64_t n_rows_max,
                                                                   const float *__restrict__ W, int M, int k,
                                                                   float *__restrict__ readout,
                                                                   int32_t *__restrict__ topk_idx_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MaSmem &S = *reinterpret_cast<MaSmem *>(smem_raw);
    int64_t nP = n_pillars_dev ? (int64_t)*n_pillars_dev : n_rows_max;
    if (nP > n_rows_max) nP = n_rows_max;
    const int64_t r0 = (int64_t)blockIdx.x * kMaRows;
    if (r0 >= nP) return;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int kRowsPerWarp = kMaRows / (kMaThreads / 32);

    for (int i = t; i < kMaRows * kMaC; i += kMaThreads) {
        const int r = i / kMaC, c = i % kMaC;
        S.p[r][c] = (r0 + r < nP) ? pillars[(r0 + r) * kMaC + c] : 0.0f;
    }
    float top_val[kRowsPerWarp];     // lane kk keeps the kk-th selected logit / index of row warp + rr * 8
    int top_idx[kRowsPerWarp];
    int filled[kRowsPerWarp];          // live entries of the list (warp-uniform)
#pragma unroll
    for (int rr = 0; rr < kRowsPerWarp; ++rr) { top_val[rr] = -INFINITY; top_idx[rr] = 0; filled[rr] = 0; }

    for (int m0 = 0; m0 < M; m0 += kMaChunk) {
        const int mc = (M - m0) < kMaChunk ? (M - m0) : kMaChunk;      // live columns of this chunk
        const int mpad = (mc + 31) & ~31;
        __syncthreads();                                                // previous chunk's readers are done (and S.p is visible)
        // logits: each thread owns memory items j = t, t+256, ...; the item row sits in registers
        for (int j = t; j < mpad; j += kMaThreads) {
            if (j < mc) {
                float w[kMaC];
                const float4 *wr = reinterpret_cast<const float4 *>(W + (int64_t)(m0 + j) * kMaC);
#pragma unroll
                for (int c4 = 0; c4 < kMaC / 4; ++c4) {
                    float4 v = __ldg(wr + c4);
                    w[4 * c4] = v.x; w[4 * c4 + 1] = v.y; w[4 * c4 + 2] = v.z; w[4 * c4 + 3] = v.w;
                }
#pragma unroll 4
                for (int r = 0; r < kMaRows; ++r) {
                    const float4 *pr = reinterpret_cast<const float4 *>(S.p[r]);
                    float acc = 0.0f;
#pragma unroll
                    for (int c4 = 0; c4 < kMaC / 4; ++c4) {
                        float4 v = pr[c4];
                        acc = fmaf(w[4 * c4], v.x, acc); acc = fmaf(w[4 * c4 + 1], v.y, acc);
                        acc = fmaf(w[4 * c4 + 2], v.z, acc); acc = fmaf(w[4 * c4 + 3], v.w, acc);
                    }
                    S.logit[r][j] = acc;
                }
            } else {
#pragma unroll
                for (int r = 0; r < kMaRows; ++r) S.logit[r][j] = -INFINITY;
            }
        }
        __syncthreads();

        // merge this chunk into the running top-k: one warp per row
#pragma unroll
        for (int rr = 0; rr < kRowsPerWarp; ++rr) {
            const int r = warp + rr * (kMaThreads / 32);
            if (r0 + r >= nP) continue;   // warp-uniform
            float *L = S.logit[r];
            for (int kk = 0; kk < k; ++kk) {
                float bv = -INFINITY; int bi = 0x7fffffff;
                for (int j = lane; j < mpad; j += 32) {
                    const float v = L[j];
                    if (v > bv) { bv = v; bi = j; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                if (bi == 0x7fffffff) break;                            // nothing finite left in this chunk (warp-uniform)
                const float kth = __shfl_sync(0xffffffffu, top_val[rr], k - 1);
                if (filled[rr] >= k && !(bv > kth)) break;              // the list is full and this chunk cannot improve it any more
                if (filled[rr] < k) ++filled[rr];
                if (lane == (bi & 31)) L[bi] = -INFINITY;
                // sorted insert: entries >= the new value stay, the rest shift one lane down (the k-th falls off)
                const int pos = __popc(__ballot_sync(0xffffffffu, lane < k && top_val[rr] >= bv));
                const float up_v = __shfl_up_sync(0xffffffffu, top_val[rr], 1);
                const int up_i = __shfl_up_sync(0xffffffffu, top_idx[rr], 1);
                if (lane == pos) { top_val[rr] = bv; top_idx[rr] = m0 + bi; }
                else if (lane > pos) { top_val[rr] = up_v; top_idx[rr] = up_i; }
                __syncwarp();
            }
        }
    }

    // softmax over the k selected logits (memory_module.py:72) + weighted readout
#pragma unroll
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int r = warp + rr * (kMaThreads / 32);
        const int64_t row = r0 + r;
        if (row >= nP) continue;
        const float my_val = top_val[rr];
        const int my_idx = top_idx[rr];
        float mx = (lane < k) ? my_val : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float e = (lane < k) ? expf(my_val - mx) : 0.0f;
        float sum = e;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float a = e / sum;
        float o0 = 0.0f, o1 = 0.0f;
        for (int kk = 0; kk < k; ++kk) {
            const float ak = __shfl_sync(0xffffffffu, a, kk);
            const int ik = __shfl_sync(0xffffffffu, my_idx, kk);
            o0 = fmaf(ak, __ldg(W + (int64_t)ik * kMaC + lane), o0);
            o1 = fmaf(ak, __ldg(W + (int64_t)ik * kMaC + 32 + lane), o1);
        }
        readout[row * kMaC + lane] = o0;
        readout[row * kMaC + 32 + lane] = o1;
        if (topk_idx_out && lane < k) topk_idx_out[row * k + lane] = my_idx;
    }
}

}  // namespace hvpr

using namespace hvpr;

int hvpr_mem_attn_fp32_init() {
    cudaError_t e = cudaFuncSetAttribute(mem_attn_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MaSmem));
    if (e != cudaSuccess) { set_cuda_error(e); return HVPR_ERR_CUDA; }
    return HVPR_OK;
}

int hvpr_mem_attn_fp32(const float *pillars, const int32_t *n_pillars_dev, int64_t n_rows_max, const float *W, int M,
                       int C, int k, float *readout, int32_t *topk_idx_out, cudaStream_t stream) {
    if (C != kMaC || M < k || k < 1 || k > 32) return HVPR_ERR_UNSUPPORTED;
    if ((uintptr_t)W % 16) return HVPR_ERR_ARG;
    const int blocks = (int)ceil_div64(n_rows_max, kMaRows);
    mem_attn_fp32_kernel<<<blocks, kMaThreads, sizeof(MaSmem), stream>>>(pillars, n_pillars_dev, n_rows_max, W, M, k,
                                                                        readout, topk_idx_out);
    HVPR_CHECK_LAUNCH();
    return HVPR_OK;
}
