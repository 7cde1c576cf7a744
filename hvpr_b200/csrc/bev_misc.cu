// N1 helpers around the tcgen05 convolutions (conv_tc.cu):
//   hvpr_nchw_to_nhwc_bf16   fp32 NCHW canvases (the module-API output of K4) -> NHWC bf16 activations
//   hvpr_attention_gate      SpatialAttention.forward's gate (pcdet/models/backbones_2d/spatial_attention.py:53-61):
//                            ChannelPool (max, mean over channels, :43-45) -> 3x3 conv 2->1 (+bias) + BN(1) -> sigmoid.
//                            The gate depends on the scale branch only, so it is evaluated once per level and reused by the
//                            three SFM iterations (base_bev_backbone.py:286-290 re-evaluates the same numbers).
#include "common.cuh"
#include <cuda_bf16.h>

namespace hvpr {

__global__ void nchw_to_nhwc_bf16_kernel(const float *__restrict__ in, int c, int64_t hw, __nv_bfloat16 *__restrict__ out, int cs) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z, c0 = blockIdx.y * 32;
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;      // 32 x 8
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int cc = c0 + ty + 8 * k;
        const int64_t p = p0 + tx;
        tile[ty + 8 * k][tx] = (cc < c && p < hw) ? in[((int64_t)n * c + cc) * hw + p] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t p = p0 + ty + 8 * k;
        const int cc = c0 + tx;
        if (p < hw && cc < c) out[((int64_t)n * hw + p) * cs + cc] = __float2bfloat16_rn(tile[tx][ty + 8 * k]);
    }
}

// one thread per pixel: max and mean over the c real channels of an NHWC bf16 tensor -> pooled (pixels, 2) fp32
__global__ void channel_pool_kernel(const __nv_bfloat16 *__restrict__ y, int64_t npix, int cs, int c, float2 *__restrict__ pooled) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(y + p * cs);
    float mx = -INFINITY, sum = 0.0f;
    for (int j = 0; j < c / 8; ++j) {
        const uint4 u = __ldg(src + j);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xFFFF0000u);
            mx = fmaxf(mx, fmaxf(lo, hi));
            sum += lo; sum += hi;
        }
    }
    pooled[p] = make_float2(mx, sum / (float)c);
}

struct GateW { float w[18]; float b; };   // folded conv+BN: w[(ch*3 + dy)*3 + dx]

__global__ void gate_conv_kernel(const float2 *__restrict__ pooled, int n, int h, int w, GateW G, float *__restrict__ gate) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t npix = (int64_t)n * h * w;
    if (p >= npix) return;
    const int x = (int)(p % w), yy = (int)((p / w) % h);
    float a = G.b;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int sy = yy + dy - 1, sx = x + dx - 1;
            if (sy >= 0 && sy < h && sx >= 0 && sx < w) {
                const float2 v = __ldg(pooled + p + (int64_t)(dy - 1) * w + (dx - 1));
                a = fmaf(G.w[dy * 3 + dx], v.x, a);
                a = fmaf(G.w[9 + dy * 3 + dx], v.y, a);
            }
        }
    gate[p] = 1.0f / (1.0f + expf(-a));
}

}  // namespace hvpr
using namespace hvpr;

extern "C" int hvpr_nchw_to_nhwc_bf16(const float *in, int n, int c, int h, int w, void *out, int out_cs, void *stream) {
    if (!in || !out || n <= 0 || c <= 0 || h <= 0 || w <= 0 || out_cs < c) return HVPR_ERR_ARG;
    const int64_t hw = (int64_t)h * w;
    dim3 grid((unsigned)ceil_div64(hw, 32), (unsigned)ceil_div64(c, 32), (unsigned)n), block(32, 8);
    nchw_to_nhwc_bf16_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, c, hw, (__nv_bfloat16 *)out, out_cs);
    HVPR_CHECK_LAUNCH();
    return HVPR_OK;
}

extern "C" int hvpr_attention_gate(const void *y_nhwc_bf16, int n, int h, int w, int cs, int c, const float *w18_host,
                                   float bias, float *pooled_ws, float *gate_out, void *stream) {
    if (!y_nhwc_bf16 || !w18_host || !pooled_ws || !gate_out || n <= 0 || h <= 0 || w <= 0) return HVPR_ERR_ARG;
    if (c <= 0 || c % 8 || cs % 8 || cs < c || (uintptr_t)y_nhwc_bf16 % 16 || (uintptr_t)pooled_ws % 8) return HVPR_ERR_UNSUPPORTED;
    const int64_t npix = (int64_t)n * h * w;
    GateW G;
    for (int i = 0; i < 18; ++i) G.w[i] = w18_host[i];
    G.b = bias;
    const unsigned blocks = (unsigned)ceil_div64(npix, 256);
    channel_pool_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)y_nhwc_bf16, npix, cs, c, (float2 *)pooled_ws);
    HVPR_CHECK_LAUNCH();
    gate_conv_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float2 *)pooled_ws, n, h, w, G, gate_out);
    HVPR_CHECK_LAUNCH();
    return HVPR_OK;
}

// ---- N2: AnchorHeadSingle eval — generate_predicted_boxes (pcdet/models/dense_heads/anchor_head_template.py:293-340)
// with ResidualCoder.decode_torch (pcdet/utils/box_coder_utils.py:45-77) and the direction-classifier fix-up
// (common_utils.limit_period, pcdet/utils/common_utils.py:20-23).  One thread per (frame, pixel, anchor); the three 1x1 head
// convolutions ran as ONE tcgen05 GEMM whose fp32 NHWC output row holds [cls A*C | box A*7 | dir A*bins | pad].
// Separately rounded mul/add (no FMA contraction), as the torch expression evaluates them.
namespace hvpr {
__global__ void __launch_bounds__(256) head_decode_kernel(const float *__restrict__ head, int cs, const float *__restrict__ anchors,
                                                          int64_t hw, int n_batch, int A, int C, int cls_off, int box_off, int dir_off,
                                                          int bins, float dir_offset, float dir_limit_offset, float period,
                                                          float *__restrict__ cls_out, float *__restrict__ box_out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)n_batch * hw * A;
    if (i >= total) return;
    const int a = (int)(i % A);
    const int64_t bp = i / A;                 // frame * hw + pixel
    const int64_t pix = bp % hw;
    const float *src = head + bp * cs;
    for (int c = 0; c < C; ++c) cls_out[i * C + c] = src[cls_off + a * C + c];
    const float *t = src + box_off + a * 7;
    const float *an = anchors + (pix * A + a) * 7;
    const float xa = an[0], ya = an[1], za = an[2], dxa = an[3], dya = an[4], dza = an[5], ra = an[6];
    const float diag = sqrtf(__fadd_rn(__fmul_rn(dxa, dxa), __fmul_rn(dya, dya)));
    float *o = box_out + i * 7;
    o[0] = __fadd_rn(__fmul_rn(t[0], diag), xa);
    o[1] = __fadd_rn(__fmul_rn(t[1], diag), ya);
    o[2] = __fadd_rn(__fmul_rn(t[2], dza), za);
    o[3] = __fmul_rn(expf(t[3]), dxa);
    o[4] = __fmul_rn(expf(t[4]), dya);
    o[5] = __fmul_rn(expf(t[5]), dza);
    float rg = __fadd_rn(t[6], ra);
    if (dir_off >= 0) {
        const float *d = src + dir_off + a * bins;
        int label = 0;
        float best = d[0];
        for (int b = 1; b < bins; ++b) if (d[b] > best) { best = d[b]; label = b; }       // torch.max: first maximum wins
        const float val = __fsub_rn(rg, dir_offset);
        const float rot = __fsub_rn(val, __fmul_rn(floorf(__fadd_rn(__fdiv_rn(val, period), dir_limit_offset)), period));
        rg = __fadd_rn(__fadd_rn(rot, dir_offset), __fmul_rn(period, (float)label));
    }
    o[6] = rg;
}
}  // namespace hvpr

extern "C" int hvpr_head_decode(const float *head_nhwc, int n, int h, int w, int cs, int A, int C, int cls_off, int box_off,
                                int dir_off, int num_dir_bins, const float *anchors, float dir_offset, float dir_limit_offset,
                                float *cls_out, float *box_out, void *stream) {
    if (!head_nhwc || !anchors || !cls_out || !box_out || n <= 0 || h <= 0 || w <= 0 || A <= 0 || C <= 0) return HVPR_ERR_ARG;
    if (cls_off < 0 || box_off < 0 || cls_off + A * C > cs || box_off + A * 7 > cs) return HVPR_ERR_ARG;
    if (dir_off >= 0 && (num_dir_bins < 1 || dir_off + A * num_dir_bins > cs)) return HVPR_ERR_ARG;
    const int64_t total = (int64_t)n * h * w * A;
    const float period = (float)(2.0 * 3.14159265358979323846 / (double)(num_dir_bins > 0 ? num_dir_bins : 1));
    hvpr::head_decode_kernel<<<(unsigned)hvpr::ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(
        head_nhwc, cs, anchors, (int64_t)h * w, n, A, C, cls_off, box_off, dir_off, num_dir_bins, dir_offset, dir_limit_offset, period,
        cls_out, box_out);
    HVPR_CHECK_LAUNCH();
    return HVPR_OK;
}
