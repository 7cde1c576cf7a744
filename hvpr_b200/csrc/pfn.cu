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

#ifndef HVPR_PFN_STRIDED
#define HVPR_PFN_STRIDED 1
#endif
#ifndef HVPR_PFN_PREFETCH
#define HVPR_PFN_PREFETCH 1
#endif
#ifndef HVPR_PFN_LOWREG_MINB
#define HVPR_PFN_LOWREG_MINB 3   // 5 (96 regs, no spills) lets K1 co-run beside PFN + fill in the streaming step, which measured WORSE: 0.741 vs 0.685 ms
#endif
constexpr int kPfnThreads = 128;
constexpr int kPfnG = 32;        // pillars per group (one group per block iteration)
constexpr int kPfnXS = 20;       // row stride (floats) of the layer-0 activation staging: conflict-free fragment loads
constexpr int kPfnYS = 72;       // row stride of the per-warp accumulator staging: conflict-free 64-bit fragment stores

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
    alignas(16) uint32_t xmax[kPfnG][kPfnXS];   // running max of the layer-0 activations (>= 0: float bits order as uints)
    alignas(16) uint32_t m1[kPfnG][68];         // running max of W1a.x as order-preserving keys
    alignas(16) uint32_t bfrag[2][8][2][2][2][32];  // [W1a|W1b][n-tile][k-step][reg][hi|lo][lane] tf32 B fragments
    alignas(16) float xs[4][32][kPfnXS];        // per-warp layer-0 activations, one row per point
    int pid[4][32];                             // per-warp pillar of each staged point (-1: none)
    float b1s[64];                              // layer-1 BN shift, lane-indexed in the epilogue
    alignas(16) float ys[4][16][kPfnYS];        // per-warp W1a.x tile (16 points x 64 channels)
};

__device__ __forceinline__ int find_pillar(const int *poff, int q) {
    int lo = 0;   // largest pl in [0, G) with poff[pl] <= q
#pragma unroll
    for (int s = kPfnG / 2; s > 0; s >>= 1)
        if (poff[lo + s] <= q) lo += s;
    return lo;
}
__device__ __forceinline__ uint32_t pfn_key(float f) {   // order-preserving float -> uint
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float pfn_unkey(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// D += A(16x8, row) * B(8x8, col), tf32 inputs, fp32 accumulate (legacy warp-level tensor-core path, HMMA in SASS)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 3xTF32: x = hi + lo with hi = tf32(x), lo = tf32(x - hi); A.B ~= Alo.Bhi + Ahi.Blo + Ahi.Bhi (error ~2^-21 relative)
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}

// Per-pillar phases run with lane = pillar and the channel range split over the 4 warps.  The warp index is folded
// into a template parameter so every weight index is a compile-time constant (uniform-register operands, no LDC).
template <int Q>
__device__ __forceinline__ void pfn_seed(const PfnParams &P, PfnSmem &S, int pl, bool padded) {
#pragma unroll
    for (int cc = 0; cc < 16; ++cc) S.m1[pl][Q * 16 + cc] = pfn_key(padded ? P.v1[Q * 16 + cc] : -INFINITY);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) S.xmax[pl][Q * 4 + kk] = __float_as_uint(padded ? P.rb0[Q * 4 + kk] : 0.0f);
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
// Both 16->64 contractions (W1a.x per point, W1b.x_max per pillar) run on the tensor cores as 3xTF32 m16n8k8 MMAs.
template <bool kScale, bool kFragRegs>
__global__ void __launch_bounds__(kPfnThreads, kFragRegs ? 3 : HVPR_PFN_LOWREG_MINB) pfn_kernel(const __grid_constant__ PfnParams P,
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
    const int gid = lane >> 2, tig = lane & 3;          // mma fragment coordinates
    int64_t nP = n_pillars_dev ? (int64_t)*n_pillars_dev : n_rows_max;
    if (nP > n_rows_max) nP = n_rows_max;
    const int64_t ngroups = (nP + kPfnG - 1) / kPfnG;
    const float4 *vox4 = reinterpret_cast<const float4 *>(voxels);

    // ---- once per (persistent) block: tf32 hi/lo B fragments of W1a and W1b ------------------------------------------
    // B(k, n) = W[n][k];  b0: k = 8*ks + tig, b1: k = 8*ks + tig + 4;  n = 8*nt + gid
    for (int e = t; e < 2 * 8 * 2 * 2 * 32; e += kPfnThreads) {
        const int ln = e & 31, reg = (e >> 5) & 1, ks = (e >> 6) & 1, nt = (e >> 7) & 7, mat = e >> 10;
        const int nn = 8 * nt + (ln >> 2), kk = 8 * ks + (ln & 3) + 4 * reg;
        const float wv = mat ? P.w.w1b[nn][kk] : P.w.w1a[nn][kk];
        uint32_t hi, lo;
        split_tf32(wv, hi, lo);
        S.bfrag[mat][nt][ks][reg][0][ln] = hi;
        S.bfrag[mat][nt][ks][reg][1][ln] = lo;
    }

    if (t < 64) S.b1s[t] = P.w.b1[t];
    __syncthreads();
    // W1a fragments live in registers for the whole (persistent) block: no shared-memory traffic in the MMA loop
    uint32_t wah[kFragRegs ? 8 : 1][2][2], wal[kFragRegs ? 8 : 1][2][2];
    if (kFragRegs) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int rg = 0; rg < 2; ++rg) { wah[nt][ks][rg] = S.bfrag[0][nt][ks][rg][0][lane]; wal[nt][ks][rg] = S.bfrag[0][nt][ks][rg][1][lane]; }
    }

    for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        // Pillar of group slot pl.  Slots are strided (pl * ngroups + grp), not consecutive: first-seen order puts the crowded
        // near-field pillars (32 real points instead of ~4) into the lowest rows of every frame, so consecutive groups
        // were 8x heavier there and the static round-robin left the blocks that drew them running alone at the end.
#if HVPR_PFN_STRIDED
#define PFN_ROW(pl) ((int64_t)(pl) * ngroups + grp)
#else
#define PFN_ROW(pl) (grp * kPfnG + (pl))
#endif
        __syncthreads();      // previous group's readers are done (and the fragments above are visible)

        // Prefetch what the NEXT group of this block touches first (count, coords, points 0..15 of every pillar row): the
        // mean phase below was the largest stall of the kernel (ncu: 15 % of the samples on its first-touch DRAM loads,
        // another 5 % on the barrier behind the count loads), and a persistent block knows its next group.
#if HVPR_PFN_PREFETCH
        {
            const int64_t grp_next = grp + gridDim.x;
            const int64_t pn = (int64_t)(t >> 2) * (HVPR_PFN_STRIDED ? ngroups : 1) + (HVPR_PFN_STRIDED ? grp_next : grp_next * kPfnG);
            if (grp_next < ngroups && pn < nP) {
                const void *a = (t & 3) == 0 ? (const void *)(vox4 + pn * T)
                              : (t & 3) == 1 ? (const void *)(num_points + pn)
                              : (t & 3) == 2 ? (const void *)(coords + pn * 4) : (const void *)(vox4 + pn * T + 8);
#if HVPR_PFN_PREFETCH == 2
                asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
#else
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
#endif
            }
        }
#endif
        // ---- phase 0: counts, exclusive scan (one warp), pillar centres --------------------------------------
        if (t < kPfnG) {
            const int64_t p = PFN_ROW(t);
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

        // ---- phase 1: per-pillar mean (4 threads per pillar, fixed combination order -> deterministic) --------
        {
            const int pl = lane, n = S.n[pl];
            const int64_t p = PFN_ROW(pl);
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

        // ---- phase 1b: scale MLP on [n, |mean|, mean_xyz]  (pillar_vfe.py:213-216) ----------------------------
        if (kScale) {
            const int pl = lane;
            const float mx = S.mean[pl][0], my = S.mean[pl][1], mz = S.mean[pl][2];
            const float in[5] = {(float)S.n[pl], sqrtf(mx * mx + my * my + mz * mz), mx, my, mz};
            PFN_DISPATCH_Q(q4, pfn_scale_hidden<Q>(P, S, pl, in));
            __syncthreads();
            float o[8];
            PFN_DISPATCH_Q(q4, pfn_scale_out<Q>(P, S, pl, o));
            const int64_t p = PFN_ROW(pl);
            if (p < nP) {
                float4 *dst = reinterpret_cast<float4 *>(scale_out + p * 32 + q4 * 8);
                dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                dst[1] = make_float4(o[4], o[5], o[6], o[7]);
            }
        }

        // ---- single pass over the real points; every warp works on its own 32 points, no block-wide barriers ---
        float (*xs)[kPfnXS] = S.xs[q4];
        float (*ys)[kPfnYS] = S.ys[q4];
        int *pid = S.pid[q4];
        for (int c0 = q4 * 32; c0 < total; c0 += kPfnThreads) {
            const int qpt = c0 + lane;
            // layer 0 (10 -> 16, ReLU) on CUDA cores, one point per lane
            float x0[16];
            int pl = -1;
            if (qpt < total) {
                pl = find_pillar(S.poff, qpt);
                const int j = qpt - S.poff[pl];
                const float4 v = __ldg(vox4 + PFN_ROW(pl) * T + j);
                float f[10];
                f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
                f[4] = v.x - S.mean[pl][0]; f[5] = v.y - S.mean[pl][1]; f[6] = v.z - S.mean[pl][2];
                f[7] = v.x - S.ctr[pl][0];  f[8] = v.y - S.ctr[pl][1];  f[9] = v.z - S.ctr[pl][2];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    float a = P.w.b0[k];
#pragma unroll
                    for (int i = 0; i < 10; ++i) a = fmaf(P.w.w0[k][i], f[i], a);
                    x0[k] = fmaxf(a, 0.0f);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) x0[k] = 0.0f;
            }
            __syncwarp();                                   // previous iteration's readers of xs / pid / ys are done
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
                *reinterpret_cast<float4 *>(&xs[lane][4 * k4]) = make_float4(x0[4 * k4], x0[4 * k4 + 1], x0[4 * k4 + 2], x0[4 * k4 + 3]);
            pid[lane] = pl;
            // bit r: row r starts a new pillar run (points arrive pillar by pillar).  The segmented maxima below test this
            // register instead of loading and comparing the pillar id of every row (a dependent LDS -> compare -> branch per row)
            const int pl_prev = __shfl_up_sync(0xffffffffu, pl, 1);
            const uint32_t bnd = __ballot_sync(0xffffffffu, lane == 0 || pl != pl_prev);
            __syncwarp();

            // pillar max of the layer-0 activations: lane = (half, column); 16 rows each
            {
                const int k = lane & 15, r0 = (lane >> 4) * 16;
                const uint32_t b16 = bnd >> r0;
                int cur = pid[r0];
                float acc = 0.0f;
                float v[16];
#pragma unroll
                for (int r = 0; r < 16; ++r) v[r] = xs[r0 + r][k];      // all loads in flight before the dependent max chain
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    if (r > 0 && ((b16 >> r) & 1u)) {
                        if (cur >= 0) atomicMax(&S.xmax[cur][k], __float_as_uint(acc));
                        cur = pid[r0 + r]; acc = v[r];
                    } else acc = fmaxf(acc, v[r]);
                }
                if (cur >= 0) atomicMax(&S.xmax[cur][k], __float_as_uint(acc));
            }

            // W1a . x on the tensor cores, one 16-point m-tile at a time
#pragma unroll 1
            for (int mt = 0; mt < 2; ++mt) {
                if (c0 + 16 * mt >= total) break;           // warp-uniform: this 16-point m-tile holds no real point
                uint32_t ahi[2][4], alo[2][4];
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    split_tf32(xs[16 * mt + gid][8 * ks + tig], ahi[ks][0], alo[ks][0]);
                    split_tf32(xs[16 * mt + gid + 8][8 * ks + tig], ahi[ks][1], alo[ks][1]);
                    split_tf32(xs[16 * mt + gid][8 * ks + tig + 4], ahi[ks][2], alo[ks][2]);
                    split_tf32(xs[16 * mt + gid + 8][8 * ks + tig + 4], ahi[ks][3], alo[ks][3]);
                }
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint32_t bh0 = kFragRegs ? wah[kFragRegs ? nt : 0][ks][0] : S.bfrag[0][nt][ks][0][0][lane];
                        const uint32_t bl0 = kFragRegs ? wal[kFragRegs ? nt : 0][ks][0] : S.bfrag[0][nt][ks][0][1][lane];
                        const uint32_t bh1 = kFragRegs ? wah[kFragRegs ? nt : 0][ks][1] : S.bfrag[0][nt][ks][1][0][lane];
                        const uint32_t bl1 = kFragRegs ? wal[kFragRegs ? nt : 0][ks][1] : S.bfrag[0][nt][ks][1][1][lane];
                        mma_tf32(c, alo[ks], bh0, bh1);
                        mma_tf32(c, ahi[ks], bl0, bl1);
                        mma_tf32(c, ahi[ks], bh0, bh1);
                    }
                    *reinterpret_cast<float2 *>(&ys[gid][8 * nt + 2 * tig]) = make_float2(c[0], c[1]);
                    *reinterpret_cast<float2 *>(&ys[gid + 8][8 * nt + 2 * tig]) = make_float2(c[2], c[3]);
                }
                __syncwarp();
                // pillar max over the 16 staged points: lane owns channels (2*lane, 2*lane+1)
                {
                    const uint32_t b16 = bnd >> (16 * mt);               // warp-uniform
                    int cur = pid[16 * mt];
                    float2 acc = make_float2(-INFINITY, -INFINITY);
#pragma unroll
                    for (int rb = 0; rb < 16; rb += 8) {
                        float2 v[8];
#pragma unroll
                        for (int r = 0; r < 8; ++r) v[r] = *reinterpret_cast<const float2 *>(&ys[rb + r][2 * lane]);
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            if (rb + r > 0 && ((b16 >> (rb + r)) & 1u)) {
                                if (cur >= 0) { atomicMax(&S.m1[cur][2 * lane], pfn_key(acc.x)); atomicMax(&S.m1[cur][2 * lane + 1], pfn_key(acc.y)); }
                                cur = pid[16 * mt + rb + r]; acc = v[r];
                            } else { acc.x = fmaxf(acc.x, v[r].x); acc.y = fmaxf(acc.y, v[r].y); }
                        }
                    }
                    if (cur >= 0) { atomicMax(&S.m1[cur][2 * lane], pfn_key(acc.x)); atomicMax(&S.m1[cur][2 * lane + 1], pfn_key(acc.y)); }
                }
                __syncwarp();
            }
        }
        __syncthreads();

        // ---- per pillar: c = b1 + W1b.x_max (tensor cores) ; pillar_features = ReLU(max + c) --------------------
        // warp q4 produces channels [16*q4, 16*q4 + 16) of all 32 pillars: 2 m-tiles x 2 n-tiles
        {
            const float (*xm)[kPfnXS] = reinterpret_cast<const float (*)[kPfnXS]>(S.xmax);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                uint32_t ahi[2][4], alo[2][4];
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    split_tf32(xm[16 * mt + gid][8 * ks + tig], ahi[ks][0], alo[ks][0]);
                    split_tf32(xm[16 * mt + gid + 8][8 * ks + tig], ahi[ks][1], alo[ks][1]);
                    split_tf32(xm[16 * mt + gid][8 * ks + tig + 4], ahi[ks][2], alo[ks][2]);
                    split_tf32(xm[16 * mt + gid + 8][8 * ks + tig + 4], ahi[ks][3], alo[ks][3]);
                }
#pragma unroll
                for (int nn = 0; nn < 2; ++nn) {
                    const int nt = 2 * q4 + nn;
                    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint32_t bh0 = S.bfrag[1][nt][ks][0][0][lane], bl0 = S.bfrag[1][nt][ks][0][1][lane];
                        const uint32_t bh1 = S.bfrag[1][nt][ks][1][0][lane], bl1 = S.bfrag[1][nt][ks][1][1][lane];
                        mma_tf32(c, alo[ks], bh0, bh1);
                        mma_tf32(c, ahi[ks], bl0, bl1);
                        mma_tf32(c, ahi[ks], bh0, bh1);
                    }
                    const int ch = 8 * nt + 2 * tig;
#pragma unroll
                    for (int hrow = 0; hrow < 2; ++hrow) {
                        const int pl = 16 * mt + gid + 8 * hrow;
                        const int64_t p = PFN_ROW(pl);
                        const float o0 = fmaxf(pfn_unkey(S.m1[pl][ch]) + (c[2 * hrow] + S.b1s[ch]), 0.0f);
                        const float o1 = fmaxf(pfn_unkey(S.m1[pl][ch + 1]) + (c[2 * hrow + 1] + S.b1s[ch + 1]), 0.0f);
                        if (p < nP) *reinterpret_cast<float2 *>(feats + p * 64 + ch) = make_float2(o0, o1);
                    }
                }
            }
        }
    }
}

}  // namespace hvpr

using namespace hvpr;

int hvpr_pfn_init() {
    cudaError_t e;
    e = cudaFuncSetAttribute(pfn_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PfnSmem));
    if (e != cudaSuccess) { set_cuda_error(e); return HVPR_ERR_CUDA; }
    e = cudaFuncSetAttribute(pfn_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PfnSmem));
    if (e != cudaSuccess) { set_cuda_error(e); return HVPR_ERR_CUDA; }
    e = cudaFuncSetAttribute(pfn_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PfnSmem));
    if (e != cudaSuccess) { set_cuda_error(e); return HVPR_ERR_CUDA; }
    e = cudaFuncSetAttribute(pfn_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PfnSmem));
    if (e != cudaSuccess) { set_cuda_error(e); return HVPR_ERR_CUDA; }
    return HVPR_OK;
}

extern "C" int hvpr_pfn(const float *voxels, const int32_t *num_points, const int32_t *coords,
                        const int32_t *n_pillars_dev, int64_t n_rows_max, int max_points,
                        const HvprPfnWeights *weights_host, const HvprGeom *geom, float x_off, float y_off, float z_off,
                        float *pillar_features, float *scale_out, float *mask_out, const HvprLaunchCfg *launch, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!weights_host || !geom || n_rows_max < 0) return HVPR_ERR_ARG;
    if (n_rows_max == 0) return HVPR_OK;
    if (!voxels || !num_points || !coords || !pillar_features) return HVPR_ERR_ARG;
    if (max_points < 1 || max_points > 32) return HVPR_ERR_UNSUPPORTED;
    const int bps = (launch && launch->blocks_per_sm != 0) ? launch->blocks_per_sm : 3;
    const bool frag_regs = !(launch && launch->variant != 0);
    if (bps < 1 || bps > 3 || (launch && (launch->variant < 0 || launch->variant > 1))) return HVPR_ERR_ARG;
    if (((uintptr_t)voxels | (uintptr_t)coords | (uintptr_t)scale_out | (uintptr_t)pillar_features) % 16) return HVPR_ERR_ARG;
    PfnParams P;
    P.w = *weights_host;
    for (int k = 0; k < 16; ++k) P.rb0[k] = P.w.b0[k] > 0.f ? P.w.b0[k] : 0.f;
    for (int c = 0; c < 64; ++c) {
        float a = 0.f;
        for (int k = 0; k < 16; ++k) a = fmaf(P.w.w1a[c][k], P.rb0[k], a);
        P.v1[c] = a;
    }
    int64_t want = ceil_div64(n_rows_max, kPfnG);
    const int64_t cap = (int64_t)num_sms() * bps;                              // persistent blocks
    const int blocks = (int)(want < cap ? want : cap);
#define HVPR_PFN_LAUNCH(SC, FR)                                                                                  \
    pfn_kernel<SC, FR><<<blocks, kPfnThreads, sizeof(PfnSmem), stream>>>(                                         \
        P, voxels, num_points, coords, n_pillars_dev, n_rows_max, max_points, geom->vs[0], geom->vs[1], geom->vs[2], \
        x_off, y_off, z_off, pillar_features, scale_out, mask_out)
    if (scale_out) { if (frag_regs) HVPR_PFN_LAUNCH(true, true); else HVPR_PFN_LAUNCH(true, false); }
    else { if (frag_regs) HVPR_PFN_LAUNCH(false, true); else HVPR_PFN_LAUNCH(false, false); }
    HVPR_CHECK_LAUNCH();
    return HVPR_OK;
}
