// K2 — fused pillar feature net: PillarVFE_Scale.forward (pcdet/models/backbones_3d/vfe/pillar_vfe.py:184-221),
// PFNLayer.forward (:29-49) x2 with eval-mode BN folded into the bias-free Linear, and the 5->16->32 scale MLP (:213-216).
//
// Point-parallel: one lane per REAL point (pillars average ~4 points of 32 slots), activations staged in warp-private
// shared memory for the per-pillar max; a task of 32 pillar rows is owned by one warp (no block barriers).  Zero-padded slots are handled analytically as ONE virtual row per pillar with
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
#define HVPR_PFN_LOWREG_MINB 2
#endif
#ifndef HVPR_PFN_WARPS
#define HVPR_PFN_WARPS 4
#endif
constexpr int kPfnWarps = HVPR_PFN_WARPS;   // 5 is 3 % faster alone, but two 5-warp blocks hold 224 KB of shared memory and starve the
                                            // voxelizer kernels that share the SMs in the streaming schedule (0.746 vs 0.667 ms per step)
constexpr int kPfnThreads = 32 * kPfnWarps;
constexpr int kPfnG = 32;        // pillars per task; a task belongs to ONE warp (no block barrier anywhere in the task loop)
constexpr int kPfnXS = 20;       // row stride (floats) of the layer-0 activation staging: conflict-free fragment loads
constexpr int kPfnYS = 72;       // row stride of the accumulator staging: conflict-free 64-bit fragment stores
constexpr int kPfnMS = 64;       // row stride of the per-pillar running maxima of W1a.x (lane l owns columns 2l, 2l+1: conflict-free)
// pillars per sub-task: with the W1a fragments in registers (the stand-alone shape) the point pass and the epilogue run in two halves
// of 16 pillars, so that the maxima buffer is 4 KB instead of 8 and THREE blocks of four warps fit an SM (0.154 -> 0.139 ms alone);
// the low-register shape used beside the canvas fill in the streaming schedule keeps whole tasks (its two blocks per SM are set
// by the fill's needs, and halves only add partially filled 32-point chunks: 0.180 -> 0.195 ms)
template <bool kFragRegs> struct PfnSub { static constexpr int kH = kFragRegs ? 16 : 32; };

struct PfnParams {
    HvprPfnWeights w;
    float rb0[16];   // ReLU(b0): layer-0 activation of a zero-padded row
    float v1[64];    // W1a . ReLU(b0)
};

// private to one warp
template <int kPfnH>
struct PfnWarpSmemT {
    int poff[kPfnG + 1];
    int pid[32];                                // pillar of each staged point (-1: none)
    float mean[kPfnG][4];
    float ctr[kPfnG][4];
    alignas(16) uint32_t xmax[kPfnG][kPfnXS];   // running max of the layer-0 activations (>= 0: float bits order as uints)
    alignas(16) float xs[32][kPfnXS];           // layer-0 activations of the current 32 points, one row per point
    alignas(16) float ys[16][kPfnYS];           // W1a.x tile (16 points x 64 channels)
    alignas(16) float m1[kPfnH][kPfnMS];        // running max over a pillar's points of W1a.x, current half task
};
constexpr int kPfnFragWords = 8 * 2 * 2 * 2 * 32;   // one matrix: [n-tile][k-step][reg][hi|lo][lane] tf32 B fragments
// kFragRegs: the W1a fragments live in registers, only W1b's sit in shared memory (8 KB instead of 16)
template <bool kFragRegs>
struct PfnSmemT {
    alignas(16) uint32_t bfrag[kFragRegs ? 1 : 2][8][2][2][2][32];  // [W1a|W1b] or [W1b]
    float b1s[64];                                  // layer-1 BN shift
    PfnWarpSmemT<PfnSub<kFragRegs>::kH> w[kPfnWarps];
};

__device__ __forceinline__ int find_pillar(const int *poff, int q) {
    int lo = 0;   // largest pl in [0, G) with poff[pl] <= q
#pragma unroll
    for (int s = kPfnG / 2; s > 0; s >>= 1)
        if (poff[lo + s] <= q) lo += s;
    return lo;
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

// max over a pillar's points commutes with the per-pillar constant and the ReLU:
//     max_p ReLU(W1a.x_p + c) = ReLU(max_p(W1a.x_p) + c),      c = b1 + W1b.x_max
// so ONE pass over the real points produces both x_max (layer 0) and max_p(W1a.x_p); c is applied per pillar afterwards.
// Both 16->64 contractions (W1a.x per point, W1b.x_max per pillar) run on the tensor cores as 3xTF32 m16n8k8 MMAs.
//
// Round 2: a task (32 strided pillar rows) is owned by ONE WARP from its counts to its outputs.  Round 1 gave a task to a block
// of four warps that split the channels of every per-pillar phase and met at six block barriers per task (ncu: 19 % barrier
// stalls, 12 warps per SM, IPC 1.2), took the pillar maxima with shared-memory atomicMax on order-preserving keys (a pillar
// could span warps) and summed the mean with four threads per pillar.  Now: lane = pillar for the per-pillar phases (counts,
// centre, mean, scale MLP), lane = point for the point pass, the running maxima of W1a.x are plain floats that only one lane
// ever touches (lane l owns channels 2l, 2l+1 and carries its running value in registers across m-tiles), and the only
// synchronisation is __syncwarp.
template <bool kScale, bool kFragRegs>
__global__ void __launch_bounds__(kPfnThreads, kFragRegs ? 3 : HVPR_PFN_LOWREG_MINB) pfn_kernel(const __grid_constant__ PfnParams P,
                                                          const float *__restrict__ voxels,
                                                          const int32_t *__restrict__ num_points,
                                                          const int32_t *__restrict__ coords,
                                                          const int32_t *__restrict__ n_pillars_dev,
                                                          int64_t n_rows_max, int T, float vx, float vy, float vz,
                                                          float x_off, float y_off, float z_off,
                                                          float *__restrict__ feats, float *__restrict__ scale_out,
                                                          float *__restrict__ mask_out, const uint4 *__restrict__ frag_image) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using PfnSmem = PfnSmemT<kFragRegs>;
    PfnSmem &S = *reinterpret_cast<PfnSmem *>(smem_raw);
    constexpr int kB = kFragRegs ? 0 : 1;               // index of W1b's fragments in S.bfrag
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int gid = lane >> 2, tig = lane & 3;          // mma fragment coordinates
    int64_t nP = n_pillars_dev ? (int64_t)*n_pillars_dev : n_rows_max;
    if (nP > n_rows_max) nP = n_rows_max;
    const int64_t ngroups = (nP + kPfnG - 1) / kPfnG;
    const float4 *vox4 = reinterpret_cast<const float4 *>(voxels);

    // ---- once per (persistent) block: tf32 hi/lo B fragments of W1a and W1b ------------------------------------------
    // B(k, n) = W[n][k];  b0: k = 8*ks + tig, b1: k = 8*ks + tig + 4;  n = 8*nt + gid
    uint32_t wah[kFragRegs ? 8 : 1][2][2], wal[kFragRegs ? 8 : 1][2][2];     // W1a fragments in registers (kFragRegs)
    if (frag_image) {
        // packed once per weight version by hvpr_pfn_pack: coalesced copies.  Building the image here reads the by-value weights
        // with a different constant address per lane (serialised, cold constant cache): ncu put 11 % of the kernel's stall samples
        // on that prologue, repeated by every block
        const uint4 *src = frag_image + (kFragRegs ? kPfnFragWords / 4 : 0);
        uint4 *dst = reinterpret_cast<uint4 *>(&S.bfrag[0][0][0][0][0][0]);
        for (int e = t; e < (int)(sizeof(S.bfrag) / 16); e += kPfnThreads) dst[e] = __ldg(src + e);
        if (kFragRegs) {
            const uint32_t *img = reinterpret_cast<const uint32_t *>(frag_image);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                    for (int rg = 0; rg < 2; ++rg) {
                        const int base = (((nt * 2 + ks) * 2 + rg) * 2) * 32 + lane;
                        wah[kFragRegs ? nt : 0][ks][rg] = __ldg(img + base); wal[kFragRegs ? nt : 0][ks][rg] = __ldg(img + base + 32);
                    }
        }
    } else {
        for (int e = t; e < 2 * kPfnFragWords / 2; e += kPfnThreads) {
            const int ln = e & 31, reg = (e >> 5) & 1, ks = (e >> 6) & 1, nt = (e >> 7) & 7, mat = e >> 10;
            const int nn = 8 * nt + (ln >> 2), kk = 8 * ks + (ln & 3) + 4 * reg;
            if (kFragRegs && mat == 0) continue;
            const float wv = mat ? P.w.w1b[nn][kk] : P.w.w1a[nn][kk];
            uint32_t hi, lo;
            split_tf32(wv, hi, lo);
            S.bfrag[kFragRegs ? 0 : mat][nt][ks][reg][0][ln] = hi;
            S.bfrag[kFragRegs ? 0 : mat][nt][ks][reg][1][ln] = lo;
        }
        if (kFragRegs) {
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                    for (int rg = 0; rg < 2; ++rg)
                        split_tf32(P.w.w1a[8 * nt + gid][8 * ks + tig + 4 * rg], wah[kFragRegs ? nt : 0][ks][rg], wal[kFragRegs ? nt : 0][ks][rg]);
        }
    }
    if (t < 64) S.b1s[t] = P.w.b1[t];
    __syncthreads();                                    // the only block barrier of the kernel
    constexpr int kPfnH = PfnSub<kFragRegs>::kH;
    PfnWarpSmemT<kPfnH> &W = S.w[wid];
    const int64_t warp0 = (int64_t)blockIdx.x * kPfnWarps + wid, nwarps = (int64_t)gridDim.x * kPfnWarps;

    for (int64_t grp = warp0; grp < ngroups; grp += nwarps) {
        // Pillar of task slot pl.  Slots are strided (pl * ngroups + grp), not consecutive: first-seen order puts the crowded
        // near-field pillars (32 real points instead of ~4) into the lowest rows of every frame, so consecutive groups
        // were 8x heavier there and a static round-robin leaves the warps that drew them running alone at the end.
#if HVPR_PFN_STRIDED
#define PFN_ROW(pl) ((int64_t)(pl) * ngroups + grp)
#else
#define PFN_ROW(pl) (grp * kPfnG + (pl))
#endif
        __syncwarp();      // previous task's readers of this warp's staging are done

        // Prefetch what the NEXT task of this warp touches first (count, coords, first points of every pillar row)
#if HVPR_PFN_PREFETCH
        {
            const int64_t grp_next = grp + nwarps;
            const int64_t pn = HVPR_PFN_STRIDED ? ((int64_t)lane * ngroups + grp_next) : (grp_next * kPfnG + lane);
            if (grp_next < ngroups && pn < nP) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(vox4 + pn * T));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(num_points + pn));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(coords + pn * 4));
            }
        }
#endif
        // ---- per-pillar phase, lane = pillar: count, centre, exclusive scan, mean, maxima seeds, scale MLP -------------
        const int64_t p = PFN_ROW(lane);
        int n = 0;
        if (p < nP) {
            n = num_points[p];
            n = n < 0 ? 0 : (n > T ? T : n);
            const int4 c = __ldg(reinterpret_cast<const int4 *>(coords) + p);   // [b, z, y, x]
            // coords*voxel + offset: mul then add, separately rounded (pillar_vfe.py:191-193)
            W.ctr[lane][0] = __fadd_rn(__fmul_rn((float)c.w, vx), x_off);
            W.ctr[lane][1] = __fadd_rn(__fmul_rn((float)c.z, vy), y_off);
            W.ctr[lane][2] = __fadd_rn(__fmul_rn((float)c.y, vz), z_off);
        }
        int inc = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        W.poff[lane] = inc - n;
        if (lane == 31) W.poff[kPfnG] = inc;
        const int total = __shfl_sync(0xffffffffu, inc, 31);
        // mean over the real points in slot order (the zero padding adds nothing): pillar_vfe.py:187, no guard as the reference
        float mx, my, mz;
        {
            float sx = 0.f, sy = 0.f, sz = 0.f;
            const int nmax = __reduce_max_sync(0xffffffffu, n);
            // eight slots per round trip: the loads of a batch are issued together, the sums stay in slot order
            for (int j0 = 0; j0 < nmax; j0 += 8) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = (j0 + u < n) ? __ldg(vox4 + p * T + j0 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 8; ++u) { sx += v[u].x; sy += v[u].y; sz += v[u].z; }
            }
            const float nf = (float)n;
            mx = __fdiv_rn(sx, nf); my = __fdiv_rn(sy, nf); mz = __fdiv_rn(sz, nf);
            W.mean[lane][0] = mx; W.mean[lane][1] = my; W.mean[lane][2] = mz;
        }
        if (mask_out && p < nP)
            for (int j = 0; j < T; ++j) mask_out[p * T + j] = (j < n) ? 1.0f : 0.0f;
        // seeds of the running maxima: the virtual zero-padded row when the pillar has padding, else the identity
        {
            const bool padded = n < T;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
                *reinterpret_cast<uint4 *>(&W.xmax[lane][4 * k4]) =
                    padded ? make_uint4(__float_as_uint(P.rb0[4 * k4]), __float_as_uint(P.rb0[4 * k4 + 1]),
                                        __float_as_uint(P.rb0[4 * k4 + 2]), __float_as_uint(P.rb0[4 * k4 + 3]))
                           : make_uint4(0u, 0u, 0u, 0u);
        }
        // scale MLP on [n, |mean|, mean_xyz]  (pillar_vfe.py:213-216)
        if (kScale) {
            const float in[5] = {(float)n, sqrtf(mx * mx + my * my + mz * mz), mx, my, mz};
            float hh[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                float a = P.w.bs0[u];
#pragma unroll
                for (int i = 0; i < 5; ++i) a = fmaf(P.w.ws0[u][i], in[i], a);
                hh[u] = fmaxf(a, 0.0f);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float o[8];
#pragma unroll
                for (int oo = 0; oo < 8; ++oo) {
                    float a = P.w.bs1[q * 8 + oo];
#pragma unroll
                    for (int u = 0; u < 16; ++u) a = fmaf(P.w.ws1[q * 8 + oo][u], hh[u], a);
                    o[oo] = fmaxf(a, 0.0f);
                }
                if (p < nP) {
                    float4 *dst = reinterpret_cast<float4 *>(scale_out + p * 32 + q * 8);
                    dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                    dst[1] = make_float4(o[4], o[5], o[6], o[7]);
                }
            }
        }
        __syncwarp();

        // ---- the task's two halves of 16 pillars: single pass over the half's real points, 32 at a time, lane = point ----
        (void)total;
#pragma unroll 1
        for (int hf = 0; hf < kPfnG / kPfnH; ++hf) {
        const int c_begin = W.poff[kPfnH * hf], c_end = W.poff[kPfnH * hf + kPfnH];
        // seeds of this half's running maxima of W1a.x: the virtual zero-padded row when the pillar has padding, else the identity;
        // lane = (pillar of the half, 32-channel half of its row)
        {
            // lane = (pillar of the sub-task, channel range): 16 pillars x 2 ranges of 32 channels, or 32 pillars x all 64 channels
            constexpr int kRanges = 32 / kPfnH, kChan = 64 / kRanges;
            const int pl_s = lane % kPfnH, cb0 = (lane / kPfnH) * kChan;
            const int n_p = __shfl_sync(0xffffffffu, n, kPfnH * hf + pl_s);
            const bool padded_p = n_p < T;
#pragma unroll
            for (int c4 = 0; c4 < kChan / 4; ++c4)
                *reinterpret_cast<float4 *>(&W.m1[pl_s][cb0 + 4 * c4]) =
                    padded_p ? make_float4(P.v1[cb0 + 4 * c4], P.v1[cb0 + 4 * c4 + 1], P.v1[cb0 + 4 * c4 + 2], P.v1[cb0 + 4 * c4 + 3])
                             : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
        __syncwarp();
        int run_pl = -1;                                     // pillar (of the half) whose running W1a.x maximum this lane carries
        float2 run = make_float2(-INFINITY, -INFINITY);      // channels (2*lane, 2*lane+1)
        for (int c0 = c_begin; c0 < c_end; c0 += 32) {
            const int qpt = c0 + lane;
            // layer 0 (10 -> 16, ReLU) on CUDA cores, one point per lane
            float x0[16];
            int pl = -1;
            if (qpt < c_end) {
                pl = find_pillar(W.poff, qpt);
                const int j = qpt - W.poff[pl];
                const float4 v = __ldg(vox4 + PFN_ROW(pl) * T + j);
                float f[10];
                f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
                f[4] = v.x - W.mean[pl][0]; f[5] = v.y - W.mean[pl][1]; f[6] = v.z - W.mean[pl][2];
                f[7] = v.x - W.ctr[pl][0];  f[8] = v.y - W.ctr[pl][1];  f[9] = v.z - W.ctr[pl][2];
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
                *reinterpret_cast<float4 *>(&W.xs[lane][4 * k4]) = make_float4(x0[4 * k4], x0[4 * k4 + 1], x0[4 * k4 + 2], x0[4 * k4 + 3]);
            W.pid[lane] = pl;
            // bit r: row r starts a new pillar run (points arrive pillar by pillar)
            const int pl_prev = __shfl_up_sync(0xffffffffu, pl, 1);
            const uint32_t bnd = __ballot_sync(0xffffffffu, lane == 0 || pl != pl_prev);
            __syncwarp();

            // pillar max of the layer-0 activations: lane = (half, column); 16 rows each.  The two halves may meet in one pillar,
            // hence the shared-memory atomicMax (values >= 0: the float bits order as unsigned integers)
            {
                const int k = lane & 15, r0 = (lane >> 4) * 16;
                const uint32_t b16 = bnd >> r0;
                int cur = W.pid[r0];
                float acc = 0.0f;
                float v[16];
#pragma unroll
                for (int r = 0; r < 16; ++r) v[r] = W.xs[r0 + r][k];      // all loads in flight before the dependent max chain
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    if (r > 0 && ((b16 >> r) & 1u)) {
                        if (cur >= 0) atomicMax(&W.xmax[cur][k], __float_as_uint(acc));
                        cur = W.pid[r0 + r]; acc = v[r];
                    } else acc = fmaxf(acc, v[r]);
                }
                if (cur >= 0) atomicMax(&W.xmax[cur][k], __float_as_uint(acc));
            }

            // W1a . x on the tensor cores, one 16-point m-tile at a time
#pragma unroll 1
            for (int mt = 0; mt < 2; ++mt) {
                if (c0 + 16 * mt >= c_end) break;           // warp-uniform: this 16-point m-tile holds no real point
                uint32_t ahi[2][4], alo[2][4];
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    split_tf32(W.xs[16 * mt + gid][8 * ks + tig], ahi[ks][0], alo[ks][0]);
                    split_tf32(W.xs[16 * mt + gid + 8][8 * ks + tig], ahi[ks][1], alo[ks][1]);
                    split_tf32(W.xs[16 * mt + gid][8 * ks + tig + 4], ahi[ks][2], alo[ks][2]);
                    split_tf32(W.xs[16 * mt + gid + 8][8 * ks + tig + 4], ahi[ks][3], alo[ks][3]);
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
                    *reinterpret_cast<float2 *>(&W.ys[gid][8 * nt + 2 * tig]) = make_float2(c[0], c[1]);
                    *reinterpret_cast<float2 *>(&W.ys[gid + 8][8 * nt + 2 * tig]) = make_float2(c[2], c[3]);
                }
                __syncwarp();
                // pillar max over the 16 staged points: lane owns channels (2*lane, 2*lane+1) and carries (run_pl, run) in registers
                // across m-tiles; a finished run is folded into m1 with a plain read-max-write (no other lane touches those words)
                {
                    const uint32_t b16 = bnd >> (16 * mt);               // warp-uniform
#pragma unroll
                    for (int rb = 0; rb < 16; rb += 8) {
                        float2 v[8];
#pragma unroll
                        for (int r = 0; r < 8; ++r) v[r] = *reinterpret_cast<const float2 *>(&W.ys[rb + r][2 * lane]);
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            if ((b16 >> (rb + r)) & 1u) {                // warp-uniform: a new pillar run starts at this row
                                if (run_pl >= 0) {
                                    float2 *m = reinterpret_cast<float2 *>(&W.m1[run_pl][2 * lane]);
                                    const float2 old = *m;
                                    *m = make_float2(fmaxf(old.x, run.x), fmaxf(old.y, run.y));
                                }
                                run_pl = W.pid[16 * mt + rb + r] - kPfnH * hf; run = v[r];
                            } else { run.x = fmaxf(run.x, v[r].x); run.y = fmaxf(run.y, v[r].y); }
                        }
                    }
                }
                __syncwarp();
            }
        }
        if (run_pl >= 0) {
            float2 *m = reinterpret_cast<float2 *>(&W.m1[run_pl][2 * lane]);
            const float2 old = *m;
            *m = make_float2(fmaxf(old.x, run.x), fmaxf(old.y, run.y));
        }
        __syncwarp();

        // ---- the half's 16 pillars: c = b1 + W1b.x_max (tensor cores) ; pillar_features = ReLU(max + c) --------------------
        {
            const float (*xm)[kPfnXS] = reinterpret_cast<const float (*)[kPfnXS]>(W.xmax);
#pragma unroll 1
            for (int mt = hf * (kPfnH / 16); mt < (hf + 1) * (kPfnH / 16); ++mt) {     // m-tiles (16 pillars each) of this sub-task
            uint32_t ahi[2][4], alo[2][4];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                split_tf32(xm[16 * mt + gid][8 * ks + tig], ahi[ks][0], alo[ks][0]);
                split_tf32(xm[16 * mt + gid + 8][8 * ks + tig], ahi[ks][1], alo[ks][1]);
                split_tf32(xm[16 * mt + gid][8 * ks + tig + 4], ahi[ks][2], alo[ks][2]);
                split_tf32(xm[16 * mt + gid + 8][8 * ks + tig + 4], ahi[ks][3], alo[ks][3]);
            }
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const uint32_t bh0 = S.bfrag[kB][nt][ks][0][0][lane], bl0 = S.bfrag[kB][nt][ks][0][1][lane];
                    const uint32_t bh1 = S.bfrag[kB][nt][ks][1][0][lane], bl1 = S.bfrag[kB][nt][ks][1][1][lane];
                    mma_tf32(c, alo[ks], bh0, bh1);
                    mma_tf32(c, ahi[ks], bl0, bl1);
                    mma_tf32(c, ahi[ks], bh0, bh1);
                }
                const int ch = 8 * nt + 2 * tig;
                const float2 bb = *reinterpret_cast<const float2 *>(&S.b1s[ch]);
#pragma unroll
                for (int hrow = 0; hrow < 2; ++hrow) {
                    const int plh = 16 * mt + gid + 8 * hrow - kPfnH * hf;      // pillar within the sub-task
                    const int64_t pr = PFN_ROW(16 * mt + gid + 8 * hrow);
                    const float2 mm = *reinterpret_cast<const float2 *>(&W.m1[plh][ch]);
                    const float o0 = fmaxf(mm.x + (c[2 * hrow] + bb.x), 0.0f);
                    const float o1 = fmaxf(mm.y + (c[2 * hrow + 1] + bb.y), 0.0f);
                    if (pr < nP) *reinterpret_cast<float2 *>(feats + pr * 64 + ch) = make_float2(o0, o1);
                }
            }
            }   // mt
        }
        __syncwarp();                                            // the next half re-seeds m1
        }   // hf
    }
}

// the fragment image of the kernel above, written to global memory once per weight version (hvpr_pfn_pack)
__global__ void __launch_bounds__(256) pfn_pack_kernel(const __grid_constant__ PfnParams P, uint32_t *__restrict__ out) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 2 * 8 * 2 * 2 * 32; e += gridDim.x * blockDim.x) {
        const int ln = e & 31, reg = (e >> 5) & 1, ks = (e >> 6) & 1, nt = (e >> 7) & 7, mat = e >> 10;
        const int nn = 8 * nt + (ln >> 2), kk = 8 * ks + (ln & 3) + 4 * reg;
        const float wv = mat ? P.w.w1b[nn][kk] : P.w.w1a[nn][kk];
        uint32_t hi, lo;
        split_tf32(wv, hi, lo);
        // same element order as PfnSmem::bfrag [mat][nt][ks][reg][hi|lo][lane]
        const int base = ((((mat * 8 + nt) * 2 + ks) * 2 + reg) * 2) * 32 + ln;
        out[base] = hi;
        out[base + 32] = lo;
    }
}

}  // namespace hvpr

using namespace hvpr;

int hvpr_pfn_init() {
    cudaError_t e;
    e = cudaFuncSetAttribute(pfn_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PfnSmemT<true>));
    if (e != cudaSuccess) { set_cuda_error(e); return HVPR_ERR_CUDA; }
    e = cudaFuncSetAttribute(pfn_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PfnSmemT<true>));
    if (e != cudaSuccess) { set_cuda_error(e); return HVPR_ERR_CUDA; }
    e = cudaFuncSetAttribute(pfn_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PfnSmemT<false>));
    if (e != cudaSuccess) { set_cuda_error(e); return HVPR_ERR_CUDA; }
    e = cudaFuncSetAttribute(pfn_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PfnSmemT<false>));
    if (e != cudaSuccess) { set_cuda_error(e); return HVPR_ERR_CUDA; }
    return HVPR_OK;
}

extern "C" size_t hvpr_pfn_packed_bytes(void) { return 2 * kPfnFragWords * sizeof(uint32_t); }

extern "C" int hvpr_pfn_pack(const HvprPfnWeights *weights_host, void *packed_dev, void *stream_) {
    if (!weights_host || !packed_dev || (uintptr_t)packed_dev % 16) return HVPR_ERR_ARG;
    PfnParams P;
    P.w = *weights_host;
    for (int k = 0; k < 16; ++k) P.rb0[k] = 0.f;
    for (int c = 0; c < 64; ++c) P.v1[c] = 0.f;
    pfn_pack_kernel<<<8, 256, 0, (cudaStream_t)stream_>>>(P, (uint32_t *)packed_dev);
    HVPR_CHECK_LAUNCH();
    return HVPR_OK;
}

extern "C" int hvpr_pfn(const float *voxels, const int32_t *num_points, const int32_t *coords,
                        const int32_t *n_pillars_dev, int64_t n_rows_max, int max_points,
                        const HvprPfnWeights *weights_host, const HvprGeom *geom, float x_off, float y_off, float z_off,
                        float *pillar_features, float *scale_out, float *mask_out, const void *weights_packed,
                        const HvprLaunchCfg *launch, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!weights_host || !geom || n_rows_max < 0) return HVPR_ERR_ARG;
    if (n_rows_max == 0) return HVPR_OK;
    if (!voxels || !num_points || !coords || !pillar_features) return HVPR_ERR_ARG;
    if (max_points < 1 || max_points > 32) return HVPR_ERR_UNSUPPORTED;
    const bool frag_regs = !(launch && launch->variant != 0);
    // shared memory admits three blocks per SM with the W1a fragments in registers (69 KB per block), two otherwise (77 KB)
    const int bps = (launch && launch->blocks_per_sm != 0) ? launch->blocks_per_sm : (frag_regs ? 3 : 2);
    if (bps < 1 || bps > 3 || (launch && (launch->variant < 0 || launch->variant > 1))) return HVPR_ERR_ARG;
    if (((uintptr_t)voxels | (uintptr_t)coords | (uintptr_t)scale_out | (uintptr_t)pillar_features | (uintptr_t)weights_packed) % 16) return HVPR_ERR_ARG;
    PfnParams P;
    P.w = *weights_host;
    for (int k = 0; k < 16; ++k) P.rb0[k] = P.w.b0[k] > 0.f ? P.w.b0[k] : 0.f;
    for (int c = 0; c < 64; ++c) {
        float a = 0.f;
        for (int k = 0; k < 16; ++k) a = fmaf(P.w.w1a[c][k], P.rb0[k], a);
        P.v1[c] = a;
    }
    int64_t want = ceil_div64(ceil_div64(n_rows_max, kPfnG), kPfnWarps);
    const int64_t cap = (int64_t)num_sms() * bps;                              // persistent blocks
    const int blocks = (int)(want < cap ? want : cap);
#define HVPR_PFN_LAUNCH(SC, FR)                                                                                  \
    pfn_kernel<SC, FR><<<blocks, kPfnThreads, sizeof(PfnSmemT<FR>), stream>>>(                                         \
        P, voxels, num_points, coords, n_pillars_dev, n_rows_max, max_points, geom->vs[0], geom->vs[1], geom->vs[2], \
        x_off, y_off, z_off, pillar_features, scale_out, mask_out, (const uint4 *)weights_packed)
    if (scale_out) { if (frag_regs) HVPR_PFN_LAUNCH(true, true); else HVPR_PFN_LAUNCH(true, false); }
    else { if (frag_regs) HVPR_PFN_LAUNCH(false, true); else HVPR_PFN_LAUNCH(false, false); }
    HVPR_CHECK_LAUNCH();
    return HVPR_OK;
}
