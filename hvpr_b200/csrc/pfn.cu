// K2 — fused pillar feature net: PillarVFE_Scale.forward (pcdet/models/backbones_3d/vfe/pillar_vfe.py:184-221),
// PFNLayer.forward (:29-49) x2 with eval-mode BN folded into the bias-free Linear, and the 5->16->32 scale MLP (:213-216).
//
// Point-parallel: one thread per REAL point (pillars average ~4 points of 32 slots), activations staged in shared
// memory for the per-pillar max.  Zero-padded slots are handled analytically as ONE virtual row per pillar with
// n < max_points: zero input -> bias-free Linear -> 0 -> BN shift b0 -> ReLU(b0) takes part in the layer-0 max, and
// layer 1 sees [ReLU(b0) | x_max] for that row.  Layer 1 is split W1 = [W1a | W1b]:
// W1b.x_max + b1 is evaluated once per pillar, W1a.x once per real point.
// Weights live in the kernel's constant bank (passed by value, 12 KB) so the unrolled FFMAs take them as immediate
// constant operands.
#include "common.cuh"

namespace hvpr {

constexpr int kPfnThreads = 128;
constexpr int kPfnG = 32;        // pillars per block
constexpr int kPfnStride = 84;   // floats per staged point row: 64 layer-1 pre-activations + 16 layer-0 activations (+pad)

struct PfnParams {
    HvprPfnWeights w;
    float rb0[16];   // ReLU(b0): layer-0 activation of a zero-padded row
    float v1[64];    // W1a . ReLU(b0)
};

struct PfnSmem {
    int n[kPfnG];
    int poff[kPfnG + 1];
    float mean[kPfnG][4];
    float ctr[kPfnG][4];
    float part[4][kPfnG][3];
    float h[kPfnG][17];
    alignas(16) float xmax[kPfnG][20];   // running max of the layer-0 activations (16 used; float4 rows)
    alignas(16) float m1[kPfnG][68];     // running max of W1a.x (64 used; padded: lane = pillar access is conflict-free)
    alignas(16) float buf[kPfnThreads * kPfnStride];
};

__device__ __forceinline__ int find_pillar(const int *poff, int q) {
    int lo = 0;   // largest pl in [0, G) with poff[pl] <= q
#pragma unroll
    for (int s = kPfnG / 2; s > 0; s >>= 1)
        if (poff[lo + s] <= q) lo += s;
    return lo;
}


// Per-pillar phases run with lane = pillar and the channel range split over the 4 warps.  The warp index is folded
// into a template parameter so every weight index is a compile-time constant (uniform-register operands, no LDC).
template <int Q>
__device__ __forceinline__ void pfn_seed(const PfnParams &P, PfnSmem &S, int pl, bool padded) {
#pragma unroll
    for (int cc = 0; cc < 16; ++cc) S.m1[pl][Q * 16 + cc] = padded ? P.v1[Q * 16 + cc] : -INFINITY;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) S.xmax[pl][Q * 4 + kk] = padded ? P.rb0[Q * 4 + kk] : 0.0f;
}
template <int Q>
__device__ __forceinline__ void pfn_scale_hidden(const PfnParams &P, PfnSmem &S, int pl, const float (&in)[5]) {
#pragma unroll
    for (int uu = 0; uu < 4; ++uu) {
        float a = P.w.bs0[Q * 4 + uu];
#pragma unroll
        for (int i = 0; i < 5; ++i) a = fmaf(P.w.ws0[Q * 4 + uu][i], in[i], a);
        S.h[pl][Q * 4 + uu] = fmaxf(a, 0.0f);
    }
}
template <int Q>
__device__ __forceinline__ void pfn_scale_out(const PfnParams &P, PfnSmem &S, int pl, float (&o)[8]) {
    float hh[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) hh[u] = S.h[pl][u];
#pragma unroll
    for (int oo = 0; oo < 8; ++oo) {
        float a = P.w.bs1[Q * 8 + oo];
#pragma unroll
        for (int u = 0; u < 16; ++u) a = fmaf(P.w.ws1[Q * 8 + oo][u], hh[u], a);
        o[oo] = fmaxf(a, 0.0f);
    }
}
template <int Q>
__device__ __forceinline__ void pfn_finish(const PfnParams &P, PfnSmem &S, int pl, float (&o)[16]) {
    float xm[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) xm[k] = S.xmax[pl][k];
#pragma unroll
    for (int cc = 0; cc < 16; ++cc) {
        float a = P.w.b1[Q * 16 + cc];
#pragma unroll
        for (int k = 0; k < 16; ++k) a = fmaf(P.w.w1b[Q * 16 + cc][k], xm[k], a);
        o[cc] = fmaxf(S.m1[pl][Q * 16 + cc] + a, 0.0f);
    }
}
#define PFN_DISPATCH_Q(q, CALL)            \
    switch (q) {                           \
        case 0: { constexpr int Q = 0; CALL; } break; \
        case 1: { constexpr int Q = 1; CALL; } break; \
        case 2: { constexpr int Q = 2; CALL; } break; \
        default: { constexpr int Q = 3; CALL; } break; \
    }

// max over a pillar's points commutes with the per-pillar constant and the ReLU:
//     max_p ReLU(W1a.x_p + c) = ReLU(max_p(W1a.x_p) + c),      c = b1 + W1b.x_max
// so ONE pass over the real points produces both x_max (layer 0) and max_p(W1a.x_p); c is applied per pillar afterwards.
template <bool kScale>
__global__ void __launch_bounds__(kPfnThreads) pfn_kernel(const __grid_constant__ PfnParams P,
                                                          const float *__restrict__ voxels,
                                                          const int32_t *__restrict__ num_points,
                                                          const int32_t *__restrict__ coords,
                                                          const int32_t *__restrict__ n_pillars_dev,
                                                          int64_t n_rows_max, int T, float vx, float vy, float vz,
                                                          float x_off, float y_off, float z_off,
                                                          float *__restrict__ feats, float *__restrict__ scale_out,
                                                          float *__restrict__ mask_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PfnSmem &S = *reinterpret_cast<PfnSmem *>(smem_raw);
    const int t = threadIdx.x, lane = t & 31, q4 = t >> 5;
    int64_t nP = n_pillars_dev ? (int64_t)*n_pillars_dev : n_rows_max;
    if (nP > n_rows_max) nP = n_rows_max;
    const int64_t g0 = (int64_t)blockIdx.x * kPfnG;
    if (g0 >= nP) return;
    const float4 *vox4 = reinterpret_cast<const float4 *>(voxels);

    // ---- phase 0: counts, exclusive scan (one warp), pillar centres ------------------------------------------
    if (t < kPfnG) {
        const int64_t p = g0 + t;
        int n = 0;
        if (p < nP) {
            n = num_points[p];
            n = n < 0 ? 0 : (n > T ? T : n);
            const int4 c = __ldg(reinterpret_cast<const int4 *>(coords) + p);   // [b, z, y, x]
            // coords*voxel + offset: mul then add, separately rounded (pillar_vfe.py:191-193)
            S.ctr[t][0] = __fadd_rn(__fmul_rn((float)c.w, vx), x_off);
            S.ctr[t][1] = __fadd_rn(__fmul_rn((float)c.z, vy), y_off);
            S.ctr[t][2] = __fadd_rn(__fmul_rn((float)c.y, vz), z_off);
        }
        S.n[t] = n;
        int inc = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        S.poff[t] = inc - n;
        if (t == kPfnG - 1) S.poff[kPfnG] = inc;
    }
    __syncthreads();
    const int total = S.poff[kPfnG];

    // ---- phase 1: per-pillar mean (4 threads per pillar, fixed combination order -> deterministic) ------------
    {
        const int pl = lane, n = S.n[pl];
        const int64_t p = g0 + pl;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        for (int j = q4; j < n; j += 4) {
            float4 v = __ldg(vox4 + p * T + j);
            sx += v.x; sy += v.y; sz += v.z;
        }
        S.part[q4][pl][0] = sx; S.part[q4][pl][1] = sy; S.part[q4][pl][2] = sz;
        if (mask_out && p < nP)
            for (int j = q4; j < T; j += 4) mask_out[p * T + j] = (j < n) ? 1.0f : 0.0f;
        // seeds of the running maxima: the virtual zero-padded row when the pillar has padding, else the identity
        const bool padded = n < T;
        PFN_DISPATCH_Q(q4, pfn_seed<Q>(P, S, pl, padded));
    }
    __syncthreads();
    if (t < kPfnG) {
        const float nf = (float)S.n[t];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float s = ((S.part[0][t][a] + S.part[1][t][a]) + S.part[2][t][a]) + S.part[3][t][a];
            S.mean[t][a] = __fdiv_rn(s, nf);                    // pillar_vfe.py:187 (no guard, as the reference)
        }
    }
    __syncthreads();

    // ---- phase 1b: scale MLP on [n, |mean|, mean_xyz]  (pillar_vfe.py:213-216) --------------------------------
    if (kScale) {
        const int pl = lane;
        const float mx = S.mean[pl][0], my = S.mean[pl][1], mz = S.mean[pl][2];
        const float in[5] = {(float)S.n[pl], sqrtf(mx * mx + my * my + mz * mz), mx, my, mz};
        PFN_DISPATCH_Q(q4, pfn_scale_hidden<Q>(P, S, pl, in));
        __syncthreads();
        float o[8];
        PFN_DISPATCH_Q(q4, pfn_scale_out<Q>(P, S, pl, o));
        const int64_t p = g0 + pl;
        if (p < nP) {
            float4 *dst = reinterpret_cast<float4 *>(scale_out + p * 32 + q4 * 8);
            dst[0] = make_float4(o[0], o[1], o[2], o[3]);
            dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
    }

    // ---- single pass over the real points: layer 0 (10->16, ReLU) and W1a.x (16->64), staged for the pillar max -----
    for (int c0 = 0; c0 < total; c0 += kPfnThreads) {
        const int qpt = c0 + t;
        if (qpt < total) {
            const int pl = find_pillar(S.poff, qpt);
            const int j = qpt - S.poff[pl];
            const float4 v = __ldg(vox4 + (g0 + pl) * T + j);
            float f[10];
            f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
            f[4] = v.x - S.mean[pl][0]; f[5] = v.y - S.mean[pl][1]; f[6] = v.z - S.mean[pl][2];
            f[7] = v.x - S.ctr[pl][0];  f[8] = v.y - S.ctr[pl][1];  f[9] = v.z - S.ctr[pl][2];
            float x0[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                float a = P.w.b0[k];
#pragma unroll
                for (int i = 0; i < 10; ++i) a = fmaf(P.w.w0[k][i], f[i], a);
                x0[k] = fmaxf(a, 0.0f);
            }
            float4 *row = reinterpret_cast<float4 *>(S.buf + t * kPfnStride);
#pragma unroll
            for (int c4 = 0; c4 < 16; ++c4) {
                float y[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float a = 0.0f;
#pragma unroll
                    for (int k = 0; k < 16; ++k) a = fmaf(P.w.w1a[c4 * 4 + e][k], x0[k], a);
                    y[e] = a;
                }
                row[c4] = make_float4(y[0], y[1], y[2], y[3]);
            }
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) row[16 + k4] = make_float4(x0[4 * k4], x0[4 * k4 + 1], x0[4 * k4 + 2], x0[4 * k4 + 3]);
        }
        __syncthreads();
        const int c1e = min(c0 + kPfnThreads, total);
        const int p_lo = find_pillar(S.poff, c0), p_hi = find_pillar(S.poff, c1e - 1);
        // pillar maxima: 20 float4 columns (16 x layer-1 pre-activations, 4 x layer-0 activations) x 6 pillar slices
        if (t < 120) {
            const int col4 = t % 20, slice = t / 20;
            const int np = p_hi - p_lo + 1, pps = (np + 5) / 6;
            const int pa = p_lo + slice * pps, pb = min(pa + pps, p_hi + 1);
            for (int pl = pa; pl < pb; ++pl) {
                const int a = max(S.poff[pl], c0), b = min(S.poff[pl + 1], c1e);
                float4 *dst = (col4 < 16) ? reinterpret_cast<float4 *>(&S.m1[pl][col4 * 4])
                                          : reinterpret_cast<float4 *>(&S.xmax[pl][(col4 - 16) * 4]);
                float4 m = *dst;
                const float4 *src = reinterpret_cast<const float4 *>(S.buf + (a - c0) * kPfnStride) + col4;
                for (int j = a; j < b; ++j, src += kPfnStride / 4) {
                    const float4 v = *src;
                    m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
                }
                *dst = m;
            }
        }
        __syncthreads();
    }

    // ---- per pillar: c = b1 + W1b.x_max ; pillar_features = ReLU(max + c) -------------------------------------------
    {
        const int pl = lane;
        float o[16];
        PFN_DISPATCH_Q(q4, pfn_finish<Q>(P, S, pl, o));
        const int64_t p = g0 + pl;
        if (p < nP) {
            float4 *dst = reinterpret_cast<float4 *>(feats + p * 64 + q4 * 16);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) dst[c4] = make_float4(o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2], o[4 * c4 + 3]);
        }
    }
}

}  // namespace hvpr

using namespace hvpr;

int hvpr_pfn_init() {
    cudaError_t e;
    e = cudaFuncSetAttribute(pfn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PfnSmem));
    if (e != cudaSuccess) { set_cuda_error(e); return HVPR_ERR_CUDA; }
    e = cudaFuncSetAttribute(pfn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PfnSmem));
    if (e != cudaSuccess) { set_cuda_error(e); return HVPR_ERR_CUDA; }
    return HVPR_OK;
}

extern "C" int hvpr_pfn(const float *voxels, const int32_t *num_points, const int32_t *coords,
                        const int32_t *n_pillars_dev, int64_t n_rows_max, int max_points,
                        const HvprPfnWeights *weights_host, const HvprGeom *geom, float x_off, float y_off, float z_off,
                        float *pillar_features, float *scale_out, float *mask_out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!weights_host || !geom || n_rows_max < 0) return HVPR_ERR_ARG;
    if (n_rows_max == 0) return HVPR_OK;
    if (!voxels || !num_points || !coords || !pillar_features) return HVPR_ERR_ARG;
    if (max_points < 1 || max_points > 32) return HVPR_ERR_UNSUPPORTED;
    if (((uintptr_t)voxels | (uintptr_t)coords | (uintptr_t)scale_out | (uintptr_t)pillar_features) % 16) return HVPR_ERR_ARG;
    PfnParams P;
    P.w = *weights_host;
    for (int k = 0; k < 16; ++k) P.rb0[k] = P.w.b0[k] > 0.f ? P.w.b0[k] : 0.f;
    for (int c = 0; c < 64; ++c) {
        float a = 0.f;
        for (int k = 0; k < 16; ++k) a = fmaf(P.w.w1a[c][k], P.rb0[k], a);
        P.v1[c] = a;
    }
    const int blocks = (int)ceil_div64(n_rows_max, kPfnG);
    if (scale_out)
        pfn_kernel<true><<<blocks, kPfnThreads, sizeof(PfnSmem), stream>>>(
            P, voxels, num_points, coords, n_pillars_dev, n_rows_max, max_points, geom->vs[0], geom->vs[1],
            geom->vs[2], x_off, y_off, z_off, pillar_features, scale_out, mask_out);
    else
        pfn_kernel<false><<<blocks, kPfnThreads, sizeof(PfnSmem), stream>>>(
            P, voxels, num_points, coords, n_pillars_dev, n_rows_max, max_points, geom->vs[0], geom->vs[1],
            geom->vs[2], x_off, y_off, z_off, pillar_features, scale_out, mask_out);
    HVPR_CHECK_LAUNCH();
    return HVPR_OK;
}
