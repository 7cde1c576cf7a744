// K4 — BEV canvas gather-fill: PointPillarScatter_Agg_Memory_1_scale.forward eval branch
// (pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py:169-220) and PointPillarScatter.forward (:14-37).
//
// The reference zero-fills both canvases, then scatters 160 strided 4-byte stores per pillar and finally copies
// everything again in torch.stack.  Here the canvas is walked in its final NCHW layout and every element is written
// exactly once — feature or 0 — with 128-bit streaming stores; the pillar row for a cell comes from the dense
// cell->row map that the voxelizer already produced (or hvpr_build_cell_map for foreign coords).
// Each (frame, cell) holds at most one pillar, so the result is order-independent and bitwise reproducible.
#include "common.cuh"
#include <cuda_bf16.h>

namespace hvpr {

struct BevSrc {
    const float *feat;   // (rows, C)
    float *out;          // (B, Ctot, cells) canvas this source writes into
    int C;               // channels of this source
    int Ctot;            // channels of the destination canvas
    int c_off;           // first destination channel
};
struct BevArgs {
    BevSrc src[3];
    int chunk_src[16];   // channel chunk -> source
    int chunk_c0[16];    // channel chunk -> first channel within the source
    int chunk_nc[16];    // channels in the chunk (multiple of 4)
};
constexpr int kBevChunk = 32;
#ifndef HVPR_BEV_BPS
#define HVPR_BEV_BPS 0
#endif
constexpr int kBevThreads = 128;   // small blocks: they slot in beside the register-heavy PFN blocks of the next batch

// Persistent form: gridDim.x blocks walk the (x-block, frame, channel-chunk) items with a grid stride and load the cell->row
// map of their NEXT item before they write the current one, so the only DRAM round trip of an item (the map) is off the
// critical path and a couple of small blocks per SM keep the store stream going.  That matters in the streaming mode, where
// the register-heavy PFN blocks of the next batch leave room for just two fill blocks per SM: with one-item blocks every item
// paid the map latency plus a block launch and the fill dropped below DRAM saturation (HvprLaunchCfg{2, 0}: 0.723 -> 0.682 ms
// per streaming step).  Alone, one block per item (n_items blocks, the default) is faster — 0.202 vs 0.227-0.244 ms: the
// hardware block scheduler balances occupied and empty items, the static grid stride does not.
#ifndef HVPR_BEV_MINB
#define HVPR_BEV_MINB 8
#endif
// kZeroedLanes != 0: the canvas already holds zeros (hvpr_mem_attn's zero_fill wrote them while it computed the readout), so
// only aligned runs of kZeroedLanes threads (2 / 4 / 8 = 32 / 64 / 128 bytes per channel) that hold a pillar are written —
// whole DRAM sectors, so that no partially written sector has to be merged with its old contents.  A template parameter: the
// write-everything form must keep its 64 registers without spills (8 blocks per SM).
template <int kZeroedLanes>
__global__ void __launch_bounds__(kBevThreads, kZeroedLanes ? 6 : HVPR_BEV_MINB) bev_fill_kernel(const __grid_constant__ BevArgs A,
                                                       const int32_t *__restrict__ cell_map, int64_t cells,
                                                       int xblocks, int n_frames, int64_t n_items) {
    constexpr int zeroed_lanes = kZeroedLanes;
    const unsigned lane_group = zeroed_lanes ? (((1u << zeroed_lanes) - 1u) << ((threadIdx.x & 31) & ~(zeroed_lanes - 1))) : 0u;
    const int64_t groups = cells >> 2;
    int64_t item = blockIdx.x;
    if (item >= n_items) return;
    auto load_map = [&](int64_t it, int64_t &g, int &f, int &ch) -> int4 {
        const int xb = (int)(it % xblocks);
        const int64_t r = it / xblocks;
        f = (int)(r % n_frames); ch = (int)(r / n_frames);
        g = (int64_t)xb * kBevThreads + threadIdx.x;                    // group of 4 consecutive cells
        return g < groups ? __ldg(reinterpret_cast<const int4 *>(cell_map + (int64_t)f * cells) + g) : make_int4(-1, -1, -1, -1);
    };
    int64_t g; int f, ch;
    int4 m = load_map(item, g, f, ch);
    while (true) {
        const int64_t next = item + gridDim.x;
        int64_t g_n = 0; int f_n = 0, ch_n = 0;
        int4 m_n = make_int4(-1, -1, -1, -1);
        if (next < n_items) m_n = load_map(next, g_n, f_n, ch_n);
        const bool live = g < groups;
        const BevSrc &s = A.src[A.chunk_src[ch]];
        const int c0 = A.chunk_c0[ch], nc = A.chunk_nc[ch];
        float4 *out = reinterpret_cast<float4 *>(s.out + ((int64_t)f * s.Ctot + s.c_off + c0) * cells) + g;
        const bool occupied = (m.x & m.y & m.z & m.w) != -1;   // row ids are >= 0, empties are exactly -1
        const unsigned occ = zeroed_lanes ? __ballot_sync(0xffffffffu, occupied) : (unsigned)__any_sync(0xffffffffu, occupied);
        if (zeroed_lanes && (occ & lane_group) == 0u) {
            // nothing to do: these cells are empty and already zero
        } else if (occ == 0u) {
            if (live) {
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
                for (int c = 0; c < nc; ++c) st_stream_f4(out + (int64_t)c * groups, z);
            }
        } else if (live) {
            const float *fa = s.feat + c0;
            const int C = s.C;
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
            for (int c = 0; c < nc; c += 4) {
                const float4 ra = m.x >= 0 ? __ldg(reinterpret_cast<const float4 *>(fa + (int64_t)m.x * C + c)) : z4;
                const float4 rb = m.y >= 0 ? __ldg(reinterpret_cast<const float4 *>(fa + (int64_t)m.y * C + c)) : z4;
                const float4 rc = m.z >= 0 ? __ldg(reinterpret_cast<const float4 *>(fa + (int64_t)m.z * C + c)) : z4;
                const float4 rd = m.w >= 0 ? __ldg(reinterpret_cast<const float4 *>(fa + (int64_t)m.w * C + c)) : z4;
                st_stream_f4(out + (int64_t)(c + 0) * groups, make_float4(ra.x, rb.x, rc.x, rd.x));
                st_stream_f4(out + (int64_t)(c + 1) * groups, make_float4(ra.y, rb.y, rc.y, rd.y));
                st_stream_f4(out + (int64_t)(c + 2) * groups, make_float4(ra.z, rb.z, rc.z, rd.z));
                st_stream_f4(out + (int64_t)(c + 3) * groups, make_float4(ra.w, rb.w, rc.w, rd.w));
            }
        }
        if (next >= n_items) break;
        item = next; m = m_n; g = g_n; f = f_n; ch = ch_n;
    }
}

// generic fallback (cells or channel counts not multiples of 4): one thread per (cell), scalar stores
__global__ void __launch_bounds__(256) bev_fill_scalar_kernel(const float *__restrict__ feat, int C, int Ctot, int c_off,
                                                              float *__restrict__ outp,
                                                              const int32_t *__restrict__ cell_map, int64_t cells) {
    const int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (cell >= cells) return;
    const int32_t r = cell_map[(int64_t)f * cells + cell];
    float *out = outp + ((int64_t)f * Ctot + c_off) * cells + cell;
    for (int c = 0; c < C; ++c) out[(int64_t)c * cells] = r >= 0 ? __ldg(feat + (int64_t)r * C + c) : 0.0f;
}

// Channels-last bf16 variant for the native backbone (conv_tc.cu consumes NHWC bf16): one thread per (cell, 16-byte piece).
// A cell's pillar row [feat_a | feat_b] (and [feat_s | 0-pad]) is contiguous in this layout, so the fill is a row copy with an
// fp32 -> bf16 rounding — 2.4x fewer bytes than the fp32 NCHW canvases and no transposition pass in front of the first conv.
__global__ void __launch_bounds__(256) bev_fill_nhwc_kernel(const float *__restrict__ fa, int ca, const float *__restrict__ fb, int cb,
                                                            const int32_t *__restrict__ cell_map, int64_t total_cells,
                                                            uint4 *__restrict__ out, int out_cs) {
    const int pieces = out_cs >> 3;                                  // 8 bf16 per 16-byte piece
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= total_cells * pieces) return;
    const int64_t cell = i / pieces;
    const int c0 = (int)(i - cell * pieces) << 3;
    const int32_t r = __ldg(cell_map + cell);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r >= 0 && c0 < ca + cb) {
        const float *src = (c0 < ca) ? fa + (int64_t)r * ca + c0 : fb + (int64_t)r * cb + (c0 - ca);
        const float4 lo = __ldg(reinterpret_cast<const float4 *>(src)), hi = __ldg(reinterpret_cast<const float4 *>(src) + 1);
        __nv_bfloat162 p0 = __floats2bfloat162_rn(lo.x, lo.y), p1 = __floats2bfloat162_rn(lo.z, lo.w);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(hi.x, hi.y), p3 = __floats2bfloat162_rn(hi.z, hi.w);
        v = make_uint4(*reinterpret_cast<uint32_t *>(&p0), *reinterpret_cast<uint32_t *>(&p1),
                       *reinterpret_cast<uint32_t *>(&p2), *reinterpret_cast<uint32_t *>(&p3));
    }
    out[i] = v;
}

__global__ void cell_map_kernel(const int32_t *__restrict__ coords, const int32_t *__restrict__ n_pillars_dev,
                                int64_t n_rows_max, int B, int nx, int ny, int32_t *__restrict__ cell_map) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t nP = n_pillars_dev ? (int64_t)*n_pillars_dev : n_rows_max;
    if (nP > n_rows_max) nP = n_rows_max;
    if (r >= nP) return;
    const int4 c = __ldg(reinterpret_cast<const int4 *>(coords) + r);      // [b, z, y, x]
    if (c.x < 0 || c.x >= B || c.z < 0 || c.z >= ny || c.w < 0 || c.w >= nx) return;
    const int64_t idx = (int64_t)c.y + (int64_t)c.z * nx + c.w;            // pointpillar_scatter.py:192 (nz == 1 -> z == 0)
    if (idx < 0 || idx >= (int64_t)nx * ny) return;
    cell_map[(int64_t)c.x * nx * ny + idx] = (int32_t)r;
}

}  // namespace hvpr

using namespace hvpr;

extern "C" int hvpr_bev_fill(const float *feat_a, int ca, const float *feat_b, int cb, const float *feat_s, int cs,
                             const int32_t *cell_map, int n_frames, int nx, int ny,
                             float *spatial, float *spatial_scale, const HvprLaunchCfg *launch, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const int bev_bps = launch ? launch->blocks_per_sm : HVPR_BEV_BPS;
    if (bev_bps < 0 || bev_bps > 16) return HVPR_ERR_ARG;
    const int variant = launch ? launch->variant : 0;
    if (variant < 0 || variant > 3) return HVPR_ERR_ARG;
    const int zeroed_lanes = variant ? (1 << variant) : 0;    // 1 / 2 / 3 -> runs of 2 / 4 / 8 threads = 32 / 64 / 128 bytes
    if (!cell_map || !spatial || !feat_a || ca <= 0 || cb < 0 || cs < 0 || n_frames <= 0 || nx <= 0 || ny <= 0) return HVPR_ERR_ARG;
    if ((cb > 0 && !feat_b) || (cs > 0 && (!feat_s || !spatial_scale))) return HVPR_ERR_ARG;
    const int64_t cells = (int64_t)nx * ny;
    const bool fast = (cells % 4 == 0) && (ca % 4 == 0) && (cb % 4 == 0) && (cs % 4 == 0) &&
                      (((uintptr_t)feat_a | (uintptr_t)feat_b | (uintptr_t)feat_s | (uintptr_t)cell_map |
                        (uintptr_t)spatial | (uintptr_t)spatial_scale) % 16 == 0);
    if (fast) {
        BevArgs A;
        A.src[0] = BevSrc{feat_a, spatial, ca, ca + cb, 0};
        A.src[1] = BevSrc{feat_b, spatial, cb, ca + cb, ca};
        A.src[2] = BevSrc{feat_s, spatial_scale, cs, cs, 0};
        int n = 0;
        for (int s = 0; s < 3; ++s)
            for (int c0 = 0; c0 < A.src[s].C; c0 += kBevChunk) {
                if (n >= 16) return HVPR_ERR_UNSUPPORTED;
                A.chunk_src[n] = s; A.chunk_c0[n] = c0;
                A.chunk_nc[n] = A.src[s].C - c0 < kBevChunk ? A.src[s].C - c0 : kBevChunk;
                ++n;
            }
        for (int i = n; i < 16; ++i) { A.chunk_src[i] = 0; A.chunk_c0[i] = 0; A.chunk_nc[i] = 0; }
        const int xblocks = (int)ceil_div64(cells / 4, kBevThreads);
        const int64_t n_items = (int64_t)xblocks * n_frames * n;
        const int64_t cap = bev_bps > 0 ? (int64_t)num_sms() * bev_bps : n_items;
        const unsigned nb = (unsigned)(n_items < cap ? n_items : cap);
        switch (zeroed_lanes) {
            case 0: bev_fill_kernel<0><<<nb, kBevThreads, 0, stream>>>(A, cell_map, cells, xblocks, n_frames, n_items); break;
            case 2: bev_fill_kernel<2><<<nb, kBevThreads, 0, stream>>>(A, cell_map, cells, xblocks, n_frames, n_items); break;
            case 4: bev_fill_kernel<4><<<nb, kBevThreads, 0, stream>>>(A, cell_map, cells, xblocks, n_frames, n_items); break;
            default: bev_fill_kernel<8><<<nb, kBevThreads, 0, stream>>>(A, cell_map, cells, xblocks, n_frames, n_items); break;
        }
        HVPR_CHECK_LAUNCH();
    } else {
        // odd shapes: the scalar kernels rewrite every element (a canvas that is already zero is simply overwritten)
        dim3 grid((unsigned)ceil_div64(cells, 256), (unsigned)n_frames);
        bev_fill_scalar_kernel<<<grid, 256, 0, stream>>>(feat_a, ca, ca + cb, 0, spatial, cell_map, cells);
        HVPR_CHECK_LAUNCH();
        if (cb > 0) { bev_fill_scalar_kernel<<<grid, 256, 0, stream>>>(feat_b, cb, ca + cb, ca, spatial, cell_map, cells); HVPR_CHECK_LAUNCH(); }
        if (cs > 0) { bev_fill_scalar_kernel<<<grid, 256, 0, stream>>>(feat_s, cs, cs, 0, spatial_scale, cell_map, cells); HVPR_CHECK_LAUNCH(); }
    }
    return HVPR_OK;
}

extern "C" int hvpr_bev_fill_nhwc_bf16(const float *feat_a, int ca, const float *feat_b, int cb, const float *feat_s, int cs,
                                       const int32_t *cell_map, int n_frames, int nx, int ny,
                                       void *spatial_nhwc, int spatial_cs, void *scale_nhwc, int scale_cs, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!cell_map || !spatial_nhwc || !feat_a || ca <= 0 || cb < 0 || cs < 0 || n_frames <= 0 || nx <= 0 || ny <= 0) return HVPR_ERR_ARG;
    if ((cb > 0 && !feat_b) || (cs > 0 && (!feat_s || !scale_nhwc))) return HVPR_ERR_ARG;
    if (ca % 8 || cb % 8 || cs % 8 || spatial_cs % 8 || scale_cs % 8 || spatial_cs < ca + cb || (cs > 0 && scale_cs < cs)) return HVPR_ERR_UNSUPPORTED;
    if (((uintptr_t)feat_a | (uintptr_t)feat_b | (uintptr_t)feat_s | (uintptr_t)spatial_nhwc | (uintptr_t)scale_nhwc) % 16) return HVPR_ERR_ARG;
    const int64_t total = (int64_t)n_frames * nx * ny;
    bev_fill_nhwc_kernel<<<(unsigned)ceil_div64(total * (spatial_cs / 8), 256), 256, 0, stream>>>(
        feat_a, ca, feat_b, cb, cell_map, total, (uint4 *)spatial_nhwc, spatial_cs);
    HVPR_CHECK_LAUNCH();
    if (cs > 0) {
        bev_fill_nhwc_kernel<<<(unsigned)ceil_div64(total * (scale_cs / 8), 256), 256, 0, stream>>>(
            feat_s, cs, nullptr, 0, cell_map, total, (uint4 *)scale_nhwc, scale_cs);
        HVPR_CHECK_LAUNCH();
    }
    return HVPR_OK;
}

extern "C" int hvpr_build_cell_map(const int32_t *coords, const int32_t *n_pillars_dev, int64_t n_rows_max,
                                   int n_frames, int nx, int ny, int32_t *cell_map, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!cell_map || n_rows_max < 0 || n_frames <= 0 || nx <= 0 || ny <= 0) return HVPR_ERR_ARG;
    if (n_rows_max > 0 && (!coords || (uintptr_t)coords % 16)) return HVPR_ERR_ARG;
    HVPR_CHECK_CUDA(cudaMemsetAsync(cell_map, 0xFF, sizeof(int32_t) * (size_t)n_frames * nx * ny, stream));
    if (n_rows_max > 0) {
        cell_map_kernel<<<(unsigned)ceil_div64(n_rows_max, 256), 256, 0, stream>>>(coords, n_pillars_dev, n_rows_max,
                                                                                  n_frames, nx, ny, cell_map);
        HVPR_CHECK_LAUNCH();
    }
    return HVPR_OK;
}
