// N1 — the convolutions of BaseBEVBackbone_Scale (pcdet/models/backbones_2d/base_bev_backbone.py:150-213,
// eval forward :280-315) as ONE persistent, warp-specialised tcgen05 implicit-GEMM kernel:
//
//   3x3 stride-1 / stride-2 Conv2d (+ZeroPad2d(1)) + folded BN + ReLU            blocks[i], sfmblocks_down[i], scale_layers[i]
//   ... with the SpatialAttention gate and the residual fused in the epilogue     x_att = gate * sfm(x_att) + x_att   (:286-290)
//   ConvTranspose2d(k = s, stride = s) + folded BN + ReLU as a 1x1 GEMM with      deblocks[i] (:180-188), written straight into
//   N = s*s*C_out and a pixel-shuffle store                                       its channel slice of the fp32 NCHW concat (:298)
//
// GEMM view: M = 128 output pixels (a bx x by patch of one image), N = bn output channels, K = taps x C_in.
// Activations are NHWC bf16.  The A operand of tap (dy,dx) is the SAME patch shifted by (dy-1, dx-1): a 4-D TMA tensor-map
// load {64 ch, bx, by, 1} with signed coordinates, out-of-bounds pixels zero-filled by the TMA unit (= the zero padding),
// landing as the 128-B-swizzled K-major tile tcgen05.mma wants.  Stride 2 uses four "parity" tensor maps over the same
// buffer (pixel strides doubled, base shifted by the parity), so a tap is again a plain shifted box — no im2col buffer, no
// wasted MACs.  Weights are pre-packed (bf16, pre-swizzled, BN scale folded in) and streamed as verbatim TMA row boxes.
// Accumulators live in TMEM, double-buffered: the epilogue of tile i overlaps the MMAs of tile i+1.
//   warp 0: TMA producer     warp 1: TMEM alloc + MMA issuer     warps 4-11: epilogue (one accumulator row per thread, two
//   warps per TMEM lane quadrant interleaving the 32-column chunks)
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>

namespace hvpr {

constexpr int kCvThreads = 384;         // warp 0 producer, warp 1 MMA issuer, warps 4-11 epilogue (two warps per TMEM lane quadrant)
constexpr int kCvABytes = 128 * 128;         // 128 pixels x 64 channels bf16
constexpr int kCvPipeBytes = 216 * 1024;     // operand ring(s)
constexpr int kCvMaxStages = 8;
constexpr int kCvMaxN = 2048;                // GEMM columns (bias staged in shared memory)

// Optional cycle accounting (compile with -DHVPR_CV_PROFILE, `make prof`): per CTA and role, cycles spent at each wait site,
// accumulated into P.prof[blockIdx.x][16].  Never enabled in the shipped library.
#ifdef HVPR_CV_PROFILE
#define CVP_DECL long long cvp_t = 0, cvp_acc[4] = {0, 0, 0, 0}
#define CVP_B() cvp_t = clock64()
#define CVP_E(i) cvp_acc[i] += clock64() - cvp_t
#define CVP_DUMP(base) do { if (P.prof) for (int i_ = 0; i_ < 4; ++i_) atomicAdd(reinterpret_cast<unsigned long long *>(P.prof) + (size_t)blockIdx.x * 16 + (base) + i_, (unsigned long long)cvp_acc[i_]); } while (0)
#else
#define CVP_DECL
#define CVP_B()
#define CVP_E(i)
#define CVP_DUMP(base)
#endif

struct ConvParams {
    alignas(64) CUtensorMap tmap[4];
    alignas(64) CUtensorMap tmap_w;       // the packed (pre-swizzled) weight image viewed as [rows][64] bf16, copied verbatim
    const uint8_t *wpk;
    const float *bias;
    const float *gate;
    const __nv_bfloat16 *residual;
    void *out;
    long long *prof;          // cycle-accounting sink (profile build only), else null
    int n_img, h_out, w_out;
    int log2_bx, tiles_x, tiles_y;
    int ntaps, kblocks, bn, n_tiles, nstages, n_total;
    int halo;                 // 3x3 stride-1: three column-shifted patches per k-block feed all 9 taps (see cv_body)
    int msub;                 // 128-pixel sub-tiles per CTA and tile (2: two A patches share every weight block)
    int pair;                 // 1: CTA-pair kernel (a tile spans the patches of both CTAs)
    int relu, res_cs;
    int out_mode, out_cs, out_c_off;
    int up, cout, out_h, out_w, out_ctot;
    int8_t tap_map[16], tap_ox[16], tap_oy[16];
};

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t cv_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cv_mbar_init(uint64_t *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cv_smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void cv_mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cv_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void cv_mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cv_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cv_mbar_wait(uint64_t *b, uint32_t parity) {
    const uint32_t a = cv_smem_u32(b);
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (!done && spin > (1u << 26)) __trap();             // watchdog: a protocol bug must not hang the GPU
    }
}
__device__ __forceinline__ void cv_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cv_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// K-major, 128-byte swizzle, rows of 128 B, 8-row atoms 1024 B apart (same encoding as mem_attn_tc.cu)
__device__ __forceinline__ uint64_t cv_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void cv_tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void cv_tmem_ld32_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                      "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                      "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                      "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]));
}
// explicit shared-space vector load: the bias table sits behind a rounded-up dynamic-shared pointer, which the compiler can only
// address generically (32 generic LD.E per chunk were 40 % of the epilogue's stall samples)
__device__ __forceinline__ float4 cv_lds_f4(uint32_t saddr) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint32_t cv_pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ float cv_bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float cv_bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

struct CvTile { int nt, img, x0, y0; };
__device__ __forceinline__ CvTile cv_decode(const ConvParams &P, int tile) {
    CvTile t;
    t.nt = tile % P.n_tiles;                 // N tiles fastest: CTAs running together share the A patch through L2
    int m = tile / P.n_tiles;
    const int tx = m % P.tiles_x; m /= P.tiles_x;
    const int ty = m % P.tiles_y;
    t.img = m / P.tiles_y;
    t.x0 = tx << P.log2_bx;
    t.y0 = ty * (128 >> P.log2_bx) * P.msub * (P.pair + 1);
    return t;
}

// Epilogue of 32 accumulator columns of one row (one output pixel): bias (folded BN shift), ReLU, gate * v + residual, then
// either the bf16 NHWC store or the pixel-shuffle store of the transposed convolution.
__device__ __forceinline__ void cv_epilogue_chunk(const ConvParams &P, const CvTile &t, uint32_t bias_s, bool valid, int64_t pix,
                                                  float g, int x, int y, const uint32_t (&r)[32], int ch) {
                const int col0 = t.nt * P.bn + ch;
                float v[32];
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 b4 = cv_lds_f4(bias_s + (uint32_t)(col0 + 4 * i4) * 4u);
                    v[4 * i4 + 0] = __uint_as_float(r[4 * i4 + 0]) + b4.x; v[4 * i4 + 1] = __uint_as_float(r[4 * i4 + 1]) + b4.y;
                    v[4 * i4 + 2] = __uint_as_float(r[4 * i4 + 2]) + b4.z; v[4 * i4 + 3] = __uint_as_float(r[4 * i4 + 3]) + b4.w;
                }
                if (P.relu) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
                }
                if (valid && P.out_mode == 0) {
                    if (P.residual) {
                        const uint4 *rp = reinterpret_cast<const uint4 *>(P.residual + pix * P.res_cs + col0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 u = __ldg(rp + j);
                            v[8 * j + 0] = fmaf(g, v[8 * j + 0], cv_bf16_lo(u.x)); v[8 * j + 1] = fmaf(g, v[8 * j + 1], cv_bf16_hi(u.x));
                            v[8 * j + 2] = fmaf(g, v[8 * j + 2], cv_bf16_lo(u.y)); v[8 * j + 3] = fmaf(g, v[8 * j + 3], cv_bf16_hi(u.y));
                            v[8 * j + 4] = fmaf(g, v[8 * j + 4], cv_bf16_lo(u.z)); v[8 * j + 5] = fmaf(g, v[8 * j + 5], cv_bf16_hi(u.z));
                            v[8 * j + 6] = fmaf(g, v[8 * j + 6], cv_bf16_lo(u.w)); v[8 * j + 7] = fmaf(g, v[8 * j + 7], cv_bf16_hi(u.w));
                        }
                    } else if (P.gate) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] *= g;
                    }
                    uint4 *op = reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(P.out) + pix * P.out_cs + P.out_c_off + col0);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        op[j] = make_uint4(cv_pack_bf16(v[8 * j], v[8 * j + 1]), cv_pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                           cv_pack_bf16(v[8 * j + 4], v[8 * j + 5]), cv_pack_bf16(v[8 * j + 6], v[8 * j + 7]));
                } else if (valid && P.out_mode == 3) {
                    // transposed conv, channels-last: GEMM column = (dy*up + dx)*cout + co, so a 32-column chunk is 32 consecutive
                    // channels of ONE output pixel (y*up + dy, x*up + dx) -> 64 contiguous bytes of the bf16 NHWC tensor
                    const int sub = col0 / P.cout, co0 = col0 % P.cout;
                    const int64_t opix = ((int64_t)t.img * P.out_h + (y * P.up + sub / P.up)) * P.out_w + (x * P.up + sub % P.up);
                    uint4 *op = reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(P.out) + opix * P.out_cs + P.out_c_off + co0);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        op[j] = make_uint4(cv_pack_bf16(v[8 * j], v[8 * j + 1]), cv_pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                           cv_pack_bf16(v[8 * j + 4], v[8 * j + 5]), cv_pack_bf16(v[8 * j + 6], v[8 * j + 7]));
                } else if (valid && P.out_mode == 2) {
                    // fp32 NHWC (dense-head logits / box deltas must not be rounded to bf16)
                    float4 *op = reinterpret_cast<float4 *>(reinterpret_cast<float *>(P.out) + pix * P.out_cs + P.out_c_off + col0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                } else if (valid) {
                    // pixel shuffle of the transposed conv: GEMM column = (dy*cout + co)*up + dx  ->  fp32 NCHW slice.
                    // dx is the fastest column index, so a thread owns `up` horizontally adjacent output pixels of each
                    // channel and neighbouring lanes (x, x+1, ...) extend the run: full 32-B sectors for every up.
                    const int per_dy = P.cout * P.up;
                    const int dy = col0 / per_dy, co0 = (col0 % per_dy) / P.up;
                    float *op = reinterpret_cast<float *>(P.out) +
                                (((int64_t)t.img * P.out_ctot + P.out_c_off + co0) * P.out_h + (y * P.up + dy)) * P.out_w + x * P.up;
                    const int64_t cstride = (int64_t)P.out_h * P.out_w;
                    if (P.up == 4) {
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            __stcs(reinterpret_cast<float4 *>(op + c * cstride), make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
                    } else if (P.up == 2) {
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            __stcs(reinterpret_cast<float2 *>(op + c * cstride), make_float2(v[2 * c], v[2 * c + 1]));
                    } else if (P.up == 1) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) __stcs(op + c * cstride, v[c]);
                    }
                }
}

// One accumulator row over the bn columns of the tile.  The TMEM loads are software-pipelined: chunk c+1 is in flight while
// chunk c is converted and stored.  Warp-collective (tcgen05.ld): every lane of the warp must call it.
__device__ __forceinline__ void cv_epilogue_rows(const ConvParams &P, const CvTile &t, uint32_t bias_s, int x, int y,
                                                 uint32_t taddr, int half) {
    const bool valid = (x < P.w_out) && (y < P.h_out);
    const int64_t pix = ((int64_t)t.img * P.h_out + y) * P.w_out + x;
    const float g = (P.gate && valid) ? __ldg(P.gate + pix) : 1.0f;
    // the two warps of a lane quadrant interleave the 32-column chunks: this warp owns chunks half, half + 2, half + 4, ...
    const int c0 = half * 32;
    if (c0 >= P.bn) return;                               // warp-uniform (bn == 32: the second warp has nothing to do)
    uint32_t ra[32], rb[32];
    __syncwarp();
    cv_tmem_ld32_issue(taddr + (uint32_t)c0, ra);
    for (int ch = c0; ch < P.bn; ch += 128) {
        cv_tmem_ld32_wait(ra);
        const bool more = ch + 64 < P.bn;                 // warp-uniform
        if (more) cv_tmem_ld32_issue(taddr + (uint32_t)(ch + 64), rb);
        cv_epilogue_chunk(P, t, bias_s, valid, pix, g, x, y, ra, ch);
        __syncwarp();                                     // re-converge after the guarded stores before the next collective op
        if (more) {
            cv_tmem_ld32_wait(rb);
            if (ch + 128 < P.bn) cv_tmem_ld32_issue(taddr + (uint32_t)(ch + 128), ra);
            cv_epilogue_chunk(P, t, bias_s, valid, pix, g, x, y, rb, ch + 64);
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------- cluster / CTA-pair helpers
__device__ __forceinline__ uint32_t cv_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cv_mapa(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cv_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Accumulator hand-back to the MMA issuer (possibly in the peer CTA).  RELAXED on purpose: the only accesses that must be ordered
// before it are the tcgen05.ld reads, which tcgen05.fence::before_thread_sync already orders; a .release here would also wait for
// the epilogue's global stores to drain (measured: 20 % of all stall samples of the kernel sat on this one instruction).
__device__ __forceinline__ void cv_mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA tensor-map loads.  kPair: the cta_group::2 form, whose mbarrier operand may live in the peer CTA (the pair's leader).
template <bool kPair>
__device__ __forceinline__ void cv_tma_4d(void *dst, const CUtensorMap *map, uint32_t bar_addr, int c0, int c1, int c2, int c3) {
    if (kPair)
        asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(cv_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(cv_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
template <bool kPair>
__device__ __forceinline__ void cv_tma_2d(void *dst, const CUtensorMap *map, uint32_t bar_addr, int c0, int c1) {
    if (kPair)
        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(cv_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_addr), "r"(c0), "r"(c1) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(cv_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_addr), "r"(c0), "r"(c1) : "memory");
}
template <bool kPair>
__device__ __forceinline__ void cv_commit(uint64_t *bar) {       // kPair: arrive on `bar` in BOTH CTAs of the pair
    if (kPair)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(cv_smem_u32(bar)), "h"((uint16_t)3) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(cv_smem_u32(bar)) : "memory");
}
template <bool kPair>
__device__ __forceinline__ void cv_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (kPair)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// ---------------------------------------------------------------------------------------------- the kernel body
// kMsub  128-pixel sub-tiles per CTA and tile (2: two A patches share every weight block; needs bn <= 128)
// kHalo  3x3 stride-1 only: three column-shifted 8 x (rows+2) patches per k-block (dx = -1, 0, +1) feed all nine taps: tap (dy,dx)
//        is patch dx read from a start address dy image rows (dy x 1024 B) further down — every 8-row UMMA group stays one aligned
//        1024-B swizzle atom.  3.4x instead of 9x the activation bytes.  (A single (8+2)-wide halo patch with starts shifted by
//        128-B rows also reads back correctly — the tensor core applies the swizzle XOR to ABSOLUTE shared-memory address bits, so
//        base_offset stays 0, measured with tools/dev/halo_diag.py — but those unaligned groups run 5-25 % slower.)
// kPair  CTA pair (cluster of 2, tcgen05 cta_group::2): UMMA M = 256 — each CTA owns the patches of its half of the tile (its
//        accumulator rows stay in its own TMEM) and loads only HALF of every weight block; the pair's tensor cores share the
//        halves.  Per SM that halves the weight bytes pulled from L2 and the B-operand bytes read from shared memory.
//        rank 0 (leader): TMA producer, MMA issuer, epilogue      rank 1: TMA producer, epilogue
//        full[s] / hfull[s] live in the leader: its producer posts the byte count of BOTH CTAs with one local arrive and both
//        CTAs' TMA loads complete_tx on it (a remote arrive per stage would put a cluster round trip on the producer's
//        critical path — measured: 2x slower); empty[s] / hempty / tfull[acc] are signalled in both CTAs by a multicast
//        tcgen05.commit; tempty[acc] lives in the leader and counts the epilogue warps of the pair.
// The single MMA-issuing thread is the critical path of the whole CTA: its loop carries no divisions, no runtime inner loops and
// precomputed descriptors (measured: a runtime sub-tile loop and it % nstages in this loop cost 12 %).
template <int kMsub, bool kHalo, bool kPair>
__device__ __forceinline__ void cv_body(const ConvParams &P) {
    extern __shared__ uint8_t cv_smem_raw[];
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(cv_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *pipe = base;
    float *bias_s = reinterpret_cast<float *>(base + kCvPipeBytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(base + kCvPipeBytes + kCvMaxN * 4);
    uint64_t *empty = full + kCvMaxStages;
    uint64_t *tfull = empty + kCvMaxStages;
    uint64_t *tempty = tfull + 2;
    uint64_t *hfull = tempty + 2;
    uint64_t *hempty = hfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(hempty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = kPair ? cv_cluster_rank() : 0u;
    const int tile0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    constexpr uint32_t kCtas = kPair ? 2u : 1u;
    constexpr int kHaloPatch = 8 * (16 * kMsub + 2) * 128;               // one column-shifted patch: 8 px x (rows + 2) x 64 ch bf16
    constexpr int kHaloBox = 3 * kHaloPatch;                              // patches for dx = -1, 0, +1
    constexpr int kHaloBytes = kHalo ? kHaloBox : 0;                      // multiple of 1024 (8 px x 128 B per image row)
    constexpr int kABytes = kHalo ? 0 : kMsub * kCvABytes;
    const int b_rows = kPair ? P.bn >> 1 : P.bn;                          // weight rows this CTA loads per block
    const int stage_bytes = kABytes + b_rows * 128;
    uint8_t *ring = pipe + 2 * kHaloBytes;                                // [halo slot 0][halo slot 1][operand ring]
    const int nst = P.nstages;
    const int kiters = P.ntaps * P.kblocks;
    const int total_tiles = P.n_img * P.tiles_y * P.tiles_x * P.n_tiles;
    const int by_cta = (128 >> P.log2_bx) * kMsub;                        // image rows this CTA owns per tile

    if (tid == 0) {
        for (int s = 0; s < kCvMaxStages; ++s) { cv_mbar_init(&full[s], 1); cv_mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) {
            cv_mbar_init(&tfull[a], 1); cv_mbar_init(&tempty[a], 8 * kCtas);
            cv_mbar_init(&hfull[a], 1); cv_mbar_init(&hempty[a], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < P.n_total; i += kCvThreads) bias_s[i] = P.bias ? P.bias[i] : 0.0f;
    if (warp == 1) {
        if (kPair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(cv_smem_u32(tmem_slot)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(cv_smem_u32(tmem_slot)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    cv_fence_before();
    __syncthreads();
    if (kPair) cv_cluster_sync();
    cv_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== producer: activation boxes (TMA tensor map) + my rows of the packed weight blocks, signalled on the leader =====
        if (lane == 0) {
            int s = 0, hs = 0;
            uint32_t ph = 0, hph = 0;
            CVP_DECL;
            for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                const CvTile t = cv_decode(P, tile);
                const int y0 = t.y0 + (int)rank * by_cta;
                const int wrow0 = t.nt * kiters * P.bn + (int)rank * b_rows;      // row of the packed image viewed as [rows][64]
                if (kHalo) {
                    for (int kb = 0; kb < P.kblocks; ++kb) {
                        cv_mbar_wait(&hempty[hs], hph ^ 1);
                        if (rank == 0) cv_mbar_expect_tx(&hfull[hs], kCtas * (uint32_t)kHaloBox);
                        const uint32_t hbar = kPair ? cv_mapa(cv_smem_u32(&hfull[hs]), 0) : cv_smem_u32(&hfull[hs]);
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx)
                            cv_tma_4d<kPair>(pipe + (size_t)hs * kHaloBytes + dx * kHaloPatch, &P.tmap[0], hbar, kb * 64, t.x0 - 1 + dx, y0 - 1, t.img);
                        if (++hs == 2) { hs = 0; hph ^= 1; }
                        for (int tap = 0; tap < 9; ++tap) {
                            CVP_B();
                            cv_mbar_wait(&empty[s], ph ^ 1);
                            CVP_E(0);
                            if (rank == 0) cv_mbar_expect_tx(&full[s], kCtas * (uint32_t)stage_bytes);
                            cv_tma_2d<kPair>(ring + (size_t)s * stage_bytes, &P.tmap_w, kPair ? cv_mapa(cv_smem_u32(&full[s]), 0) : cv_smem_u32(&full[s]),
                                             0, wrow0 + (tap * P.kblocks + kb) * P.bn);
                            if (++s == nst) { s = 0; ph ^= 1; }
                        }
                    }
                } else {
                    for (int tap = 0; tap < P.ntaps; ++tap) {
                        const CUtensorMap *map = &P.tmap[P.tap_map[tap]];
                        const int x = t.x0 + P.tap_ox[tap], y = y0 + P.tap_oy[tap];
                        for (int kb = 0; kb < P.kblocks; ++kb) {
                            CVP_B();
                            cv_mbar_wait(&empty[s], ph ^ 1);
                            CVP_E(0);
                            uint8_t *dst = ring + (size_t)s * stage_bytes;
                            const uint32_t lfull = kPair ? cv_mapa(cv_smem_u32(&full[s]), 0) : cv_smem_u32(&full[s]);
                            if (rank == 0) cv_mbar_expect_tx(&full[s], kCtas * (uint32_t)stage_bytes);
                            cv_tma_4d<kPair>(dst, map, lfull, kb * 64, x, y, t.img);
                            cv_tma_2d<kPair>(dst + kABytes, &P.tmap_w, lfull, 0, wrow0 + (tap * P.kblocks + kb) * P.bn);
                            if (++s == nst) { s = 0; ph ^= 1; }
                        }
                    }
                }
            }
            CVP_DUMP(0);
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (the pair's leader only) ====================================================================
        if (lane == 0 && rank == 0) {
            // idesc: D=f32 (1<<4), A=bf16 (1<<7), B=bf16 (1<<10), K-major both, N>>3 at [17,23), M>>4 at [24,29)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.bn >> 3) << 17) |
                                   ((uint32_t)((kPair ? 256 : 128) >> 4) << 24);
            const uint64_t ring_desc = cv_desc_sw128(cv_smem_u32(ring));
            const uint32_t stage_units = (uint32_t)stage_bytes >> 4;          // descriptor address units (16 B)
            const uint32_t bn = (uint32_t)P.bn;
            int s = 0, hs = 0;
            uint32_t ph = 0, hph = 0, ti = 0, soff = 0;
            CVP_DECL;
            for (int tile = tile0; tile < total_tiles; tile += tile_step, ++ti) {
                const uint32_t acc = ti & 1;
                CVP_B();
                cv_mbar_wait(&tempty[acc], ((ti >> 1) & 1) ^ 1);                  // the epilogues drained this accumulator
                CVP_E(0);
                cv_fence_after();
                const uint32_t d = tmem_base + acc * 256u;
                uint32_t accum = 0;
                if (kHalo) {
                    for (int kb = 0; kb < P.kblocks; ++kb) {
                        CVP_B();
                        cv_mbar_wait(&hfull[hs], hph);
                        CVP_E(3);
                        const uint32_t ha = cv_smem_u32(pipe + (size_t)hs * kHaloBytes);
                        for (int dy = 0; dy < 3; ++dy)
                            for (int dx = 0; dx < 3; ++dx) {
                                CVP_B();
                                cv_mbar_wait(&full[s], ph);
                                CVP_E(1);
                                CVP_B();
                                cv_fence_after();
                                const uint64_t bdesc = ring_desc + soff;
#pragma unroll
                                for (int m = 0; m < kMsub; ++m) {
                                    const uint64_t adesc = cv_desc_sw128(ha + (uint32_t)(dx * kHaloPatch + (m * 16 + dy) * 1024));
#pragma unroll
                                    for (int kk = 0; kk < 4; ++kk)
                                        cv_mma<kPair>(d + (uint32_t)m * bn, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                                                      accum | (uint32_t)(kk > 0));
                                }
                                accum = 1;
                                cv_commit<kPair>(&empty[s]);
                                CVP_E(2);
                                soff += stage_units;
                                if (++s == nst) { s = 0; ph ^= 1; soff = 0; }
                            }
                        cv_commit<kPair>(&hempty[hs]);
                        if (++hs == 2) { hs = 0; hph ^= 1; }
                    }
                } else {
                    for (int ki = 0; ki < kiters; ++ki) {
                        CVP_B();
                        cv_mbar_wait(&full[s], ph);
                        CVP_E(1);
                        CVP_B();
                        cv_fence_after();
                        const uint64_t adesc0 = ring_desc + soff, bdesc = adesc0 + (uint64_t)(kABytes >> 4);
#pragma unroll
                        for (int m = 0; m < kMsub; ++m) {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)        // +32 B along K inside the 128-B swizzle atom
                                cv_mma<kPair>(d + (uint32_t)m * bn, adesc0 + (uint64_t)(m * (kCvABytes >> 4) + kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                                              accum | (uint32_t)(kk > 0));
                        }
                        accum = 1;
                        cv_commit<kPair>(&empty[s]);
                        CVP_E(2);
                        soff += stage_units;
                        if (++s == nst) { s = 0; ph ^= 1; soff = 0; }
                    }
                }
                cv_commit<kPair>(&tfull[acc]);
            }
            CVP_DUMP(4);
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===== epilogue: my 128 x kMsub accumulator rows =============================================================
        const int q = warp & 3, half = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const int px = row & ((1 << P.log2_bx) - 1), py = row >> P.log2_bx;
        const uint32_t tempty_addr = kPair ? cv_mapa(cv_smem_u32(&tempty[0]), 0) : cv_smem_u32(&tempty[0]);
        uint32_t ti = 0;
        CVP_DECL;
        for (int tile = tile0; tile < total_tiles; tile += tile_step, ++ti) {
            const CvTile t = cv_decode(P, tile);
            const uint32_t acc = ti & 1;
            CVP_B();
            cv_mbar_wait(&tfull[acc], (ti >> 1) & 1);
            CVP_E(0);
            CVP_B();
            cv_fence_after();
#pragma unroll
            for (int m = 0; m < kMsub; ++m)
                cv_epilogue_rows(P, t, cv_smem_u32(bias_s), t.x0 + px, t.y0 + (int)rank * by_cta + m * (128 >> P.log2_bx) + py,
                                 tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256u + (uint32_t)(m * P.bn), half);
            cv_fence_before();
            __syncwarp();
            if (lane == 0) cv_mbar_arrive_cluster(tempty_addr + acc * 8u);
            CVP_E(1);
        }
        if (warp == 4 && lane == 0) { CVP_DUMP(8); }
    }

    cv_fence_before();
    __syncthreads();
    if (kPair) cv_cluster_sync();           // the peer may still multicast into my barriers / read my shared memory until here
    if (warp == 1) {
        cv_fence_after();
        if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

template <int kMsub, bool kHalo>
__global__ void __launch_bounds__(kCvThreads, 1) conv_tc_kernel(const __grid_constant__ ConvParams P) { cv_body<kMsub, kHalo, false>(P); }
template <int kMsub, bool kHalo>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kCvThreads, 1) conv_tc2_kernel(const __grid_constant__ ConvParams P) { cv_body<kMsub, kHalo, true>(P); }

// folded fp32 weights (n_total, taps, cin) -> bf16 image [n-tile][tap][k-block][bn rows x 128 B], 128-B swizzled
__global__ void conv_pack_kernel(const float *__restrict__ w, int n_total, int taps, int cin, int bn, uint8_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // one 16-byte piece (8 input channels) per thread
    const int kblocks = cin / 64;
    const int64_t total = (int64_t)n_total * taps * kblocks * 8;
    if (i >= total) return;
    const int j = (int)(i & 7);
    int64_t rest = i >> 3;
    const int kb = (int)(rest % kblocks); rest /= kblocks;
    const int tap = (int)(rest % taps);
    const int n = (int)(rest / taps);
    const float *src = w + ((int64_t)n * taps + tap) * cin + kb * 64 + j * 8;
    uint4 pk;
    pk.x = cv_pack_bf16(src[0], src[1]); pk.y = cv_pack_bf16(src[2], src[3]);
    pk.z = cv_pack_bf16(src[4], src[5]); pk.w = cv_pack_bf16(src[6], src[7]);
    const int nt = n / bn, r = n % bn;
    const size_t off = ((size_t)(nt * taps + tap) * kblocks + kb) * (size_t)(bn * 128) + (size_t)(r >> 3) * 1024 +
                       (size_t)(r & 7) * 128 + (size_t)((j ^ (r & 7)) << 4);
    *reinterpret_cast<uint4 *>(out + off) = pk;
}

}  // namespace hvpr
using namespace hvpr;

typedef CUresult (*CvEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CvEncodeFn cv_encode_fn() {
    static CvEncodeFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (CvEncodeFn)p;
    }
    return fn;
}

static size_t cv_smem_bytes() { return 1024 + kCvPipeBytes + kCvMaxN * 4 + (2 * kCvMaxStages + 8) * 8 + 16; }

int hvpr_conv_init() {
    const void *fns[8] = {(const void *)conv_tc_kernel<1, false>, (const void *)conv_tc_kernel<2, false>,
                          (const void *)conv_tc_kernel<1, true>, (const void *)conv_tc_kernel<2, true>,
                          (const void *)conv_tc2_kernel<1, false>, (const void *)conv_tc2_kernel<2, false>,
                          (const void *)conv_tc2_kernel<1, true>, (const void *)conv_tc2_kernel<2, true>};
    for (int i = 0; i < 8; ++i) {
        cudaError_t e = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cv_smem_bytes());
        if (e != cudaSuccess) { set_cuda_error(e); return HVPR_ERR_CUDA; }
    }
    return HVPR_OK;
}

// debug knob (not part of the public header): 0 = automatic tile policy, 1 / 2 = force the number of 128-pixel sub-tiles
static int g_cv_force_msub = 0;
static int g_cv_halo_off = 1;
extern "C" int hvpr_dbg_conv_force_msub(int msub) {
    if (msub < 0 || msub > 2) return HVPR_ERR_ARG;
    g_cv_force_msub = msub;
    return HVPR_OK;
}
// knob: 0 = automatic choice between the single-CTA and the CTA-pair (cta_group::2) kernel, 1 = never pair, 2 = always pair
static long long *g_cv_prof = nullptr;
extern "C" int hvpr_dbg_conv_prof(void *buf) { g_cv_prof = (long long *)buf; return HVPR_OK; }   // (grid, 16) int64, profile build
static int g_cv_pair_mode = 0;
extern "C" int hvpr_dbg_conv_pair(int mode) {
    if (mode < 0 || mode > 2) return HVPR_ERR_ARG;
    g_cv_pair_mode = mode;
    return HVPR_OK;
}
// knob: 1 (default) = one TMA box per tap; 0 = 3x3 stride-1 layers read all nine taps out of three column-shifted patches
extern "C" int hvpr_dbg_conv_halo_off(int off) { g_cv_halo_off = off ? 1 : 0; return HVPR_OK; }

extern "C" size_t hvpr_conv_packed_bytes(int n_total, int taps, int c_in) { return (size_t)n_total * taps * c_in * 2; }

extern "C" int hvpr_conv_pack_weights(const float *w_ntc, int n_total, int taps, int c_in, int bn, void *out_packed, void *stream) {
    if (!w_ntc || !out_packed || n_total <= 0 || taps <= 0 || c_in <= 0) return HVPR_ERR_ARG;
    if (c_in % 64 || bn < 32 || bn > 256 || (bn & (bn - 1)) || n_total % bn || n_total > kCvMaxN) return HVPR_ERR_UNSUPPORTED;
    if ((uintptr_t)out_packed % 16 || (uintptr_t)w_ntc % 4) return HVPR_ERR_ARG;
    const int64_t total = (int64_t)n_total * taps * (c_in / 64) * 8;
    conv_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w_ntc, n_total, taps, c_in, bn,
                                                                                      (uint8_t *)out_packed);
    HVPR_CHECK_LAUNCH();
    return HVPR_OK;
}

extern "C" int hvpr_conv2d(const HvprConvArgs *a, void *stream) {
    if (!a || !a->in || !a->w_packed || !a->out) return HVPR_ERR_ARG;
    if (a->n <= 0 || a->h_in <= 0 || a->w_in <= 0) return HVPR_ERR_ARG;
    if (a->c_in <= 0 || a->c_in % 64 || a->in_cs < a->c_in || a->in_cs % 8) return HVPR_ERR_UNSUPPORTED;
    if (a->bn < 32 || a->bn > 256 || (a->bn & (a->bn - 1)) || a->n_total % a->bn || a->n_total > kCvMaxN) return HVPR_ERR_UNSUPPORTED;
    if (!((a->ksize == 3 && (a->stride == 1 || a->stride == 2)) || (a->ksize == 1 && a->stride == 1))) return HVPR_ERR_UNSUPPORTED;
    if (a->stride == 2 && ((a->h_in | a->w_in) & 1)) return HVPR_ERR_UNSUPPORTED;
    if (((uintptr_t)a->in | (uintptr_t)a->w_packed | (uintptr_t)a->out | (uintptr_t)a->residual) % 16) return HVPR_ERR_ARG;
    CvEncodeFn enc = cv_encode_fn();
    if (!enc) return HVPR_ERR_UNSUPPORTED;

    ConvParams P;
    memset(&P, 0, sizeof(P));
    P.n_img = a->n;
    P.prof = g_cv_prof;
    P.h_out = a->h_in / a->stride;
    P.w_out = a->w_in / a->stride;
    // patch shape: the power-of-two split of 128 pixels that wastes the fewest out-of-image pixels (ties: wider rows)
    // two 128-pixel sub-tiles per tile when the column tile is narrow: halves the weight traffic per MAC
    P.msub = (a->bn <= 128 && (int64_t)a->n * P.h_out * P.w_out >= 2 * 128 * (int64_t)num_sms()) ? 2 : 1;
    if (g_cv_force_msub == 1 || (g_cv_force_msub == 2 && a->bn <= 128)) P.msub = g_cv_force_msub;
    // automatic policy (measured per layer with tools/dev/backbone_bench.py, 432 x 496 canvases, batch 8): the pair kernel wins
    // 6-9 % on the 256-column 3x3 layers (weight traffic and operand reads halved per SM); narrow layers and the transposed
    // convolutions (short K loops, store-bound epilogue) are faster on the single-CTA kernel
    const bool pair = g_cv_pair_mode == 2 ||
                      (g_cv_pair_mode == 0 && a->bn == 256 && a->ksize == 3 && a->out_mode == 0 &&
                       (int64_t)a->n * P.h_out * P.w_out >= 256 * (int64_t)(num_sms() / 2) * 4);
    P.pair = pair ? 1 : 0;
    P.halo = (a->ksize == 3 && a->stride == 1 && !g_cv_halo_off) ? 1 : 0;
    if (P.halo) P.msub = 1;      // two double-buffered patch triples of a 256-pixel tile (2 x 102 KB) do not fit beside the weight ring
    if (pair && (a->bn > 128 || g_cv_force_msub == 1)) P.msub = 1;
    int best = -1; int64_t best_cost = 0;
    // transposed-conv epilogue: a warp's store covers bx*up consecutive floats per output row -> keep patches >= 16 pixels wide
    const int l_min = P.halo ? 3 : ((a->out_mode == 1 && P.w_out >= 16) ? 4 : 0);
    for (int l = l_min; l <= (P.halo ? 3 : 7); ++l) {      // halo path: patches are 8 pixels wide
        const int bx = 1 << l, by = (128 >> l) * P.msub * (P.pair + 1);
        const int64_t cost = ceil_div64(P.w_out, bx) * bx * (ceil_div64(P.h_out, by) * by);
        if (best < 0 || cost < best_cost || (cost == best_cost && l <= 4)) { best = l; best_cost = cost; }
    }
    P.log2_bx = best;
    const int bx = 1 << best, by = (128 >> best) * P.msub * (P.pair + 1);      // rows of a whole tile
    const int by_cta = (128 >> best) * P.msub;                                    // rows one CTA loads
    P.tiles_x = (int)ceil_div64(P.w_out, bx);
    P.tiles_y = (int)ceil_div64(P.h_out, by);
    P.kblocks = a->c_in / 64;
    P.bn = a->bn;
    P.n_total = a->n_total;
    P.n_tiles = a->n_total / a->bn;
    const int halo_bytes = P.halo ? 3 * 8 * (16 * P.msub + 2) * 128 : 0;
    const int stage_bytes = (P.halo ? 0 : P.msub * kCvABytes) + (pair ? a->bn * 64 : a->bn * 128);
    P.nstages = (kCvPipeBytes - 2 * halo_bytes) / stage_bytes;
    if (P.nstages > kCvMaxStages) P.nstages = kCvMaxStages;
    P.wpk = (const uint8_t *)a->w_packed;
    P.bias = a->bias;
    P.gate = a->gate;
    P.residual = (const __nv_bfloat16 *)a->residual;
    P.res_cs = a->res_cs;
    P.relu = a->relu;
    P.out = a->out;
    P.out_mode = a->out_mode;
    P.out_cs = a->out_cs;
    P.out_c_off = a->out_c_off;
    if (a->out_mode == 0 || a->out_mode == 2) {
        if (a->out_cs % 8 || a->out_c_off % 8 || a->out_c_off + a->n_total > a->out_cs) return HVPR_ERR_ARG;
        if (a->out_mode == 2 && (a->residual || a->gate)) return HVPR_ERR_ARG;
        if (a->residual && (a->res_cs % 8 || a->res_cs < a->n_total)) return HVPR_ERR_ARG;
    } else if (a->out_mode == 1 || a->out_mode == 3) {
        if (a->out_mode == 3 && (a->out_cs % 8 || a->out_c_off % 8 || a->out_c_off + a->c_out > a->out_cs)) return HVPR_ERR_ARG;
        if (a->up != 1 && a->up != 2 && a->up != 4) return HVPR_ERR_UNSUPPORTED;
        if (a->c_out <= 0 || a->c_out % 32 || a->n_total != a->up * a->up * a->c_out) return HVPR_ERR_ARG;
        if ((a->out_mode == 1 && a->out_c_off + a->c_out > a->out_ctot) || a->residual || a->gate) return HVPR_ERR_ARG;
        P.up = a->up; P.cout = a->c_out; P.out_ctot = a->out_ctot;
        P.out_h = P.h_out * a->up; P.out_w = P.w_out * a->up;
    } else return HVPR_ERR_ARG;

    // tensor maps over the NHWC bf16 input: dims (channel, x, y, image)
    const cuuint32_t box[4] = {64u, (cuuint32_t)(P.halo ? 8 : bx), (cuuint32_t)(P.halo ? by_cta + 2 : by_cta), 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const uint64_t pix_b = (uint64_t)a->in_cs * 2u;
    const int nmaps = (a->stride == 2) ? 4 : 1;
    for (int m = 0; m < nmaps; ++m) {
        const int pyy = m >> 1, pxx = m & 1;
        void *gbase = (uint8_t *)a->in + ((uint64_t)pyy * a->w_in + pxx) * pix_b;
        const cuuint64_t dims[4] = {(cuuint64_t)a->c_in, (cuuint64_t)(a->w_in / a->stride), (cuuint64_t)(a->h_in / a->stride),
                                    (cuuint64_t)a->n};
        const cuuint64_t strides[3] = {pix_b * a->stride, pix_b * a->w_in * a->stride, pix_b * a->w_in * a->h_in};
        CUresult r = enc(&P.tmap[m], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, gbase, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return HVPR_ERR_ARG;
    }
    P.ntaps = a->ksize * a->ksize;
    for (int dy = 0; dy < a->ksize; ++dy)
        for (int dx = 0; dx < a->ksize; ++dx) {
            const int tpi = dy * a->ksize + dx;
            const int oy = (a->ksize == 3) ? dy - 1 : 0, ox = (a->ksize == 3) ? dx - 1 : 0;
            if (a->stride == 2) {      // input pixel 2*o + off  ->  parity (off & 1), half-resolution offset (off - parity) / 2
                const int ppy = oy & 1, ppx = ox & 1;
                P.tap_map[tpi] = (int8_t)(ppy * 2 + ppx);
                P.tap_oy[tpi] = (int8_t)((oy - ppy) / 2);
                P.tap_ox[tpi] = (int8_t)((ox - ppx) / 2);
            } else {
                P.tap_map[tpi] = 0; P.tap_oy[tpi] = (int8_t)oy; P.tap_ox[tpi] = (int8_t)ox;
            }
        }
    const int64_t total_tiles = (int64_t)P.n_img * P.tiles_x * P.tiles_y * P.n_tiles;
    {
        // weights: the pre-swizzled image as a plain [rows][64] bf16 tensor (no TMA swizzle: rows are copied verbatim); a CTA pulls
        // bn rows of every block, or bn/2 in the pair kernel
        const cuuint64_t wdims[2] = {64u, (cuuint64_t)a->n_total * (cuuint64_t)(P.ntaps * P.kblocks)};
        const cuuint64_t wstr[1] = {128u};
        const cuuint32_t wbox[2] = {64u, (cuuint32_t)(pair ? a->bn / 2 : a->bn)};
        const cuuint32_t wes[2] = {1u, 1u};
        CUresult r = enc(&P.tmap_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)a->w_packed, wdims, wstr, wbox, wes,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return HVPR_ERR_ARG;
    }
    const size_t smem = cv_smem_bytes();
    cudaStream_t st = (cudaStream_t)stream;
    if (pair) {
        const unsigned grid = 2u * (unsigned)(total_tiles < num_sms() / 2 ? total_tiles : num_sms() / 2);
        if (P.msub == 2) { if (P.halo) conv_tc2_kernel<2, true><<<grid, kCvThreads, smem, st>>>(P); else conv_tc2_kernel<2, false><<<grid, kCvThreads, smem, st>>>(P); }
        else { if (P.halo) conv_tc2_kernel<1, true><<<grid, kCvThreads, smem, st>>>(P); else conv_tc2_kernel<1, false><<<grid, kCvThreads, smem, st>>>(P); }
    } else {
        const unsigned grid = (unsigned)(total_tiles < num_sms() ? total_tiles : num_sms());
        if (P.msub == 2) { if (P.halo) conv_tc_kernel<2, true><<<grid, kCvThreads, smem, st>>>(P); else conv_tc_kernel<2, false><<<grid, kCvThreads, smem, st>>>(P); }
        else { if (P.halo) conv_tc_kernel<1, true><<<grid, kCvThreads, smem, st>>>(P); else conv_tc_kernel<1, false><<<grid, kCvThreads, smem, st>>>(P); }
    }
    HVPR_CHECK_LAUNCH();
    return HVPR_OK;
}
