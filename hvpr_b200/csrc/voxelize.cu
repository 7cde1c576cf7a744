// K1 — point -> pillar voxelization, bit-exact with the serial first-seen loop of spconv's VoxelGenerator
// (call site pcdet/datasets/processor/data_processor.py:50-67; loop witness tools/vis.py:23-50) and collated as
// pcdet/datasets/dataset.py:159-166.
//
// Parallel restatement (validated against the serial loop in NumPy: oracle/voxelize.py::voxelize_np):
//   hash    per point: cell = floor((p-lo)/vs) per axis in IEEE fp32 (no FMA, no reciprocal);  dense per-frame table
//           (a perfect hash for nz=1 pillar grids):  first[cell] = atomicMin(point index), cnt[cell] += 1.
//           Grids whose dense table does not fit (nz > 1 at fine resolution; more than 2^31 cells) go through an OPEN-ADDRESSING
//           table instead: 64-bit cell keys, 2 x max_frame_points slots per frame (load <= 0.5), linear probing with atomicCAS;
//           every later kernel indexes the table by SLOT, so only the hash kernel and the coordinate decode differ.  Which slot a
//           cell lands in depends on the race, the outputs (first index / count per cell, ranks) do not.
//   assign  single-pass exclusive scan of (is_first, cnt) in point order (block aggregates published with a ready bit)
//           -> first-seen voxel rank + CSR segment offset; caps applied on the rank
//   fill    unordered CSR fill of point indices per kept voxel
//   gather  a warp per 32 strided voxel slots: keep the 32 lowest point indices (bitonic sort / merge in registers) = arrival
//           order of the serial loop; write the 512-byte zero-padded voxel row, coords, count and the cell->row map
// All atomics are integer min/add, so the result does not depend on scheduling.
#include "common.cuh"
#include <limits.h>

namespace hvpr {

constexpr int kScanThreads = 256;
#ifndef HVPR_SCAN_ITEMS
#define HVPR_SCAN_ITEMS 8
#endif
constexpr int kScanItems = HVPR_SCAN_ITEMS;                       // points per thread in count/assign
constexpr int kScanTile = kScanThreads * kScanItems;

struct VoxWorkspace {
    long long *keys;    // hashed mode only: [B*slots] cell id of each slot (-1 = free)
    int2 *table;        // [B*cells] {first, cnt}   (hashed mode: cells = slots per frame)
    int32_t *cellbuf;   // [n_total]
    unsigned long long *agg;  // [B*blocks_per_frame] block aggregates {sum_cnt : ready flag | n_first}, zeroed by init
    int32_t *ticket;    // [B] scan-block tickets (dispatch order), zeroed by init
    int32_t *vox_cell;  // [B*max_vox]
    int32_t *seg_off;   // [B*max_vox]
    int32_t *cursor;    // [B*max_vox]
    int32_t *csr;       // [n_total]
    int32_t *frame_nvox;// [B]
    int32_t *istar;     // [B]
    size_t bytes;
};

static VoxWorkspace carve(void *base, int64_t n_total, int B, int64_t cells, int max_vox, int64_t blocks_per_frame, bool hashed) {
    VoxWorkspace w;
    size_t off = 0;
    auto take = [&](size_t bytes) { void *p = base ? (char *)base + off : nullptr; off += align_up(bytes, 256); return p; };
    w.keys = hashed ? (long long *)take(sizeof(long long) * (size_t)B * cells) : nullptr;
    w.table = (int2 *)take(sizeof(int2) * (size_t)B * cells);
    w.cellbuf = (int32_t *)take(sizeof(int32_t) * (size_t)n_total);
    w.agg = (unsigned long long *)take(sizeof(unsigned long long) * (size_t)B * blocks_per_frame);
    w.ticket = (int32_t *)take(sizeof(int32_t) * (size_t)B);
    w.vox_cell = (int32_t *)take(sizeof(int32_t) * (size_t)B * max_vox);
    w.seg_off = (int32_t *)take(sizeof(int32_t) * (size_t)B * max_vox);
    w.cursor = (int32_t *)take(sizeof(int32_t) * (size_t)B * max_vox);
    w.csr = (int32_t *)take(sizeof(int32_t) * (size_t)n_total);
    w.frame_nvox = (int32_t *)take(sizeof(int32_t) * (size_t)B);
    w.istar = (int32_t *)take(sizeof(int32_t) * (size_t)B);
    w.bytes = off;
    return w;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void vox_init_kernel(long long *keys, int2 *table, int64_t n_table, int32_t *cell_map, int64_t n_map,
                                int32_t *cursor, int64_t n_cursor, unsigned long long *agg, int64_t n_agg,
                                int32_t *frame_nvox, int32_t *istar, int32_t *ticket, int B) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // table entries are 8 B; write two per thread as one 16-B store
    int4 *t4 = reinterpret_cast<int4 *>(table);
    int64_t n4 = n_table / 2;
    for (int64_t j = i; j < n4; j += stride) t4[j] = make_int4(INT_MAX, 0, INT_MAX, 0);
    if (i == 0 && (n_table & 1)) table[n_table - 1] = make_int2(INT_MAX, 0);
    if (keys)
        for (int64_t j = i; j < n_table; j += stride) keys[j] = -1ll;
    if (cell_map) {
        int4 *m4 = reinterpret_cast<int4 *>(cell_map);
        int64_t nm4 = n_map / 4;
        for (int64_t j = i; j < nm4; j += stride) m4[j] = make_int4(-1, -1, -1, -1);
        for (int64_t j = nm4 * 4 + i; j < n_map; j += stride) cell_map[j] = -1;
    }
    for (int64_t j = i; j < n_cursor; j += stride) cursor[j] = 0;
    for (int64_t j = i; j < n_agg; j += stride) agg[j] = 0ull;
    if (i < B) { frame_nvox[i] = 0; istar[i] = INT_MAX; ticket[i] = 0; }
}

// cell index of one point; -1 when outside the grid.  IEEE fp32 sub, div, floor — exactly the CPU arithmetic.
__device__ __forceinline__ int32_t point_cell(float x, float y, float z, const HvprGeom &g) {
    float cx = floorf(__fdiv_rn(__fsub_rn(x, g.lo[0]), g.vs[0]));
    float cy = floorf(__fdiv_rn(__fsub_rn(y, g.lo[1]), g.vs[1]));
    float cz = floorf(__fdiv_rn(__fsub_rn(z, g.lo[2]), g.vs[2]));
    bool ok = (cx >= 0.0f) && (cx < (float)g.grid[0]) && (cy >= 0.0f) && (cy < (float)g.grid[1]) &&
              (cz >= 0.0f) && (cz < (float)g.grid[2]);   // NaN fails every comparison -> rejected
    if (!ok) return -1;
    return ((int32_t)cz * g.grid[1] + (int32_t)cy) * g.grid[0] + (int32_t)cx;
}

// 64-bit cell id for grids beyond 2^31 cells (same fp32 arithmetic)
__device__ __forceinline__ long long point_cell64(float x, float y, float z, const HvprGeom &g) {
    float cx = floorf(__fdiv_rn(__fsub_rn(x, g.lo[0]), g.vs[0]));
    float cy = floorf(__fdiv_rn(__fsub_rn(y, g.lo[1]), g.vs[1]));
    float cz = floorf(__fdiv_rn(__fsub_rn(z, g.lo[2]), g.vs[2]));
    bool ok = (cx >= 0.0f) && (cx < (float)g.grid[0]) && (cy >= 0.0f) && (cy < (float)g.grid[1]) &&
              (cz >= 0.0f) && (cz < (float)g.grid[2]);
    if (!ok) return -1ll;
    return ((long long)cz * g.grid[1] + (long long)cy) * g.grid[0] + (long long)cx;
}
// open addressing, linear probing: the slot that holds `cell` in this frame's key table (claims a free one on first sight).
// slots is a power of two and at least twice the frame's point count, so a free slot always exists.
__device__ __forceinline__ int32_t hash_slot(long long *keys, int64_t slots, long long cell) {
    const unsigned long long mask = (unsigned long long)slots - 1ull;
    unsigned long long s = (((unsigned long long)cell * 0x9E3779B97F4A7C15ull) >> 24) & mask;
    for (;;) {
        const long long prev = (long long)atomicCAS(reinterpret_cast<unsigned long long *>(keys + s), ~0ull, (unsigned long long)cell);
        if (prev == -1ll || prev == cell) return (int32_t)s;
        s = (s + 1ull) & mask;
    }
}

// grid: (blocks over max_frame_points, B).  Point index stored in the table is frame-local.
// kVoxPPT points per thread, block-strided (coalesced), all loads issued before the first use: the kernel is one
// dependent chain per point (load -> cell -> two REDs; ncu: long_scoreboard + drain, 11 % issue activity).  Measured for
// hash + fill on the headline batch: 1 / 2 / 4 / 8 points per thread = 0.1076 / 0.1045 / 0.106 / 0.1081 ms for all of K1 —
// per-thread latency is not the bound (the scattered REDs into the 13.7 MB table are), so 2 it is.
#ifndef HVPR_VOX_PPT
#define HVPR_VOX_PPT 2
#endif
constexpr int kVoxPPT = HVPR_VOX_PPT;
template <bool kVec4, bool kHashed>
__global__ void __launch_bounds__(256) vox_hash_kernel(const float *__restrict__ pts, int stride, int xyz_col,
                                                       const int32_t *__restrict__ frame_off, int32_t frame_cap, HvprGeom g,
                                                       int64_t cells, long long *keys, int2 *table, int32_t *__restrict__ cellbuf) {
    const int f = blockIdx.y;
    const int32_t start = frame_off[f];
    const int32_t n = min(frame_off[f + 1] - start, frame_cap);      // a frame longer than the caller's bound is cut, identically in every kernel
    const int32_t i0 = blockIdx.x * (256 * kVoxPPT) + threadIdx.x;
    if (i0 >= n) return;
    float x[kVoxPPT], y[kVoxPPT], z[kVoxPPT];
#pragma unroll
    for (int k = 0; k < kVoxPPT; ++k) {
        const int32_t i = i0 + k * 256;
        x[k] = y[k] = z[k] = 0.f;
        if (i < n) {
            const int64_t gi = (int64_t)start + i;
            if (kVec4) {
                float4 p = __ldg(reinterpret_cast<const float4 *>(pts) + gi);
                x[k] = p.x; y[k] = p.y; z[k] = p.z;
            } else {
                const float *p = pts + gi * stride + xyz_col;
                x[k] = __ldg(p); y[k] = __ldg(p + 1); z[k] = __ldg(p + 2);
            }
        }
    }
    int2 *tab = table + (int64_t)f * cells;
#pragma unroll
    for (int k = 0; k < kVoxPPT; ++k) {
        const int32_t i = i0 + k * 256;
        if (i < n) {
            int32_t c;
            if (kHashed) {
                const long long c64 = point_cell64(x[k], y[k], z[k], g);
                c = c64 < 0 ? -1 : hash_slot(keys + (int64_t)f * cells, cells, c64);
            } else {
                c = point_cell(x[k], y[k], z[k], g);
            }
            cellbuf[(int64_t)start + i] = c;
            if (c >= 0) {
                atomicMin(&tab[c].x, i);
                atomicAdd(&tab[c].y, 1);
            }
        }
    }
}

// Exclusive scan in point order -> voxel rank / CSR offset for every first point.  Single pass: a block sums its own
// (is_first, cnt-of-first) pairs, publishes the pair as ONE 64-bit word carrying a ready bit, then adds up the words of
// the blocks in front of it in the same frame (a frame has a few dozen to a few hundred blocks, so a flat sum replaces a
// look-back chain).  Blocks take their position from a per-frame ticket, so every block a block waits for is already
// running, whatever order the hardware dispatches them in.
constexpr unsigned long long kAggReady = 0x80000000ull;
__global__ void __launch_bounds__(kScanThreads) vox_assign_kernel(const int32_t *__restrict__ frame_off, int32_t frame_cap, int64_t cells,
                                                                  int2 *table, const int32_t *__restrict__ cellbuf,
                                                                  unsigned long long *agg, int32_t *ticket,
                                                                  int blocks_per_frame,
                                                                  int max_vox, int32_t *__restrict__ vox_cell,
                                                                  int32_t *__restrict__ seg_off,
                                                                  int32_t *__restrict__ frame_nvox,
                                                                  int32_t *__restrict__ istar) {
    const int f = blockIdx.y;
    const int32_t start = frame_off[f];
    const int32_t n = min(frame_off[f + 1] - start, frame_cap);
    if ((int64_t)blockIdx.x * kScanTile >= n) return;       // as many blocks take a ticket as the frame has tiles
    __shared__ int s_a[kScanThreads / 32], s_b[kScanThreads / 32], s_pa[kScanThreads / 32], s_pb[kScanThreads / 32];
    __shared__ int s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(&ticket[f], 1);
    __syncthreads();
    const int tile = s_tile;
    const int32_t base = tile * kScanTile;
    unsigned long long *fagg = agg + (int64_t)f * blocks_per_frame;

    int2 *tab = table + (int64_t)f * cells;
    int32_t cell[kScanItems];
    int cnt[kScanItems];
    bool isf[kScanItems];
    int32_t cc[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int32_t i = base + threadIdx.x * kScanItems + k;
        cc[k] = (i < n) ? cellbuf[(int64_t)start + i] : -1;
    }
    int ta = 0, tb = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int32_t i = base + threadIdx.x * kScanItems + k;
        isf[k] = false; cnt[k] = 0; cell[k] = -1;
        if (cc[k] >= 0) {
            const int2 e = tab[cc[k]];
            if (e.x == i) { isf[k] = true; cnt[k] = e.y; cell[k] = cc[k]; ta += 1; tb += e.y; }
        }
    }
    // block exclusive scan of (ta, tb)
    int ia = ta, ib = tb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int va = __shfl_up_sync(0xffffffffu, ia, o), vb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += va; ib += vb; }
    }
    if (lane == 31) { s_a[warp] = ia; s_b[warp] = ib; }
    __syncthreads();
    int wa = 0, wb = 0, tot_a = 0, tot_b = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        if (w < warp) { wa += s_a[w]; wb += s_b[w]; }
        tot_a += s_a[w]; tot_b += s_b[w];
    }
    if (threadIdx.x == 0)
        atomicExch(&fagg[tile], ((unsigned long long)(uint32_t)tot_b << 32) | kAggReady | (unsigned long long)(uint32_t)tot_a);

    // sum of the aggregates of the tiles in front (published before their owners wait, so this cannot deadlock)
    int pa = 0, pb = 0;
    for (int j = threadIdx.x; j < tile; j += kScanThreads) {
        unsigned long long v;
        do { v = *reinterpret_cast<volatile unsigned long long *>(fagg + j); } while (!(v & kAggReady));
        pa += (int)(v & 0x7fffffffull); pb += (int)(v >> 32);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { pa += __shfl_xor_sync(0xffffffffu, pa, o); pb += __shfl_xor_sync(0xffffffffu, pb, o); }
    if (lane == 0) { s_pa[warp] = pa; s_pb[warp] = pb; }
    __syncthreads();
    int pref_a = 0, pref_b = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) { pref_a += s_pa[w]; pref_b += s_pb[w]; }

    int rank = pref_a + wa + (ia - ta);
    int off = pref_b + wb + (ib - tb);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (isf[k]) {
            int32_t i = base + threadIdx.x * kScanItems + k;
            if (rank < max_vox) {
                vox_cell[(int64_t)f * max_vox + rank] = cell[k];
                seg_off[(int64_t)f * max_vox + rank] = off;
                tab[cell[k]].x = -(rank + 1);         // cell -> voxel id, read by fill (next launch)
            } else if (rank == max_vox) {
                istar[f] = i;                          // first point that would open voxel #max_vox (break mode)
            }
            rank += 1; off += cnt[k];
        }
    }
    // the block holding the frame's last point publishes the (capped) voxel count
    if (base + kScanTile >= n && threadIdx.x == kScanThreads - 1) {
        int total = rank;   // last thread's running rank == inclusive total
        frame_nvox[f] = total < max_vox ? total : max_vox;
    }
}

// kVoxPPT points per thread, stage by stage (cell -> voxel id -> cursor claim + segment offset -> CSR store) so that the
// four dependent memory round trips of kVoxPPT points overlap.
__global__ void __launch_bounds__(256) vox_fill_kernel(const int32_t *__restrict__ frame_off, int32_t frame_cap, int64_t cells,
                                                       const int2 *__restrict__ table,
                                                       const int32_t *__restrict__ cellbuf, int max_vox,
                                                       const int32_t *__restrict__ seg_off, int32_t *cursor,
                                                       int32_t *__restrict__ csr, const int32_t *__restrict__ istar,
                                                       int break_mode) {
    const int f = blockIdx.y;
    const int32_t start = frame_off[f];
    int32_t n = min(frame_off[f + 1] - start, frame_cap);
    if (break_mode) { const int32_t is = istar[f]; n = is < n ? is : n; }       // points from istar on are dropped
    const int32_t i0 = blockIdx.x * (256 * kVoxPPT) + threadIdx.x;
    if (i0 >= n) return;
    const int2 *tab = table + (int64_t)f * cells;
    int32_t c[kVoxPPT], v[kVoxPPT];
#pragma unroll
    for (int k = 0; k < kVoxPPT; ++k) {
        const int32_t i = i0 + k * 256;
        c[k] = (i < n) ? cellbuf[(int64_t)start + i] : -1;
    }
#pragma unroll
    for (int k = 0; k < kVoxPPT; ++k) {
        const int32_t e = (c[k] >= 0) ? tab[c[k]].x : 0;
        v[k] = (e < 0) ? -e - 1 : -1;                      // e >= 0: the cell belongs to a voxel beyond the cap
    }
    int pos[kVoxPPT], so[kVoxPPT];
#pragma unroll
    for (int k = 0; k < kVoxPPT; ++k) {
        pos[k] = 0; so[k] = 0;
        if (v[k] >= 0) {
            pos[k] = atomicAdd(&cursor[(int64_t)f * max_vox + v[k]], 1);
            so[k] = seg_off[(int64_t)f * max_vox + v[k]];
        }
    }
#pragma unroll
    for (int k = 0; k < kVoxPPT; ++k)
        if (v[k] >= 0) csr[(int64_t)start + so[k] + pos[k]] = i0 + k * 256;
}

__device__ __forceinline__ int32_t bitonic_sort32_asc(int32_t x, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            int32_t y = __shfl_xor_sync(0xffffffffu, x, j);
            bool up = ((lane & k) == 0);          // ascending block
            bool lower = ((lane & j) == 0);
            x = (lower == up) ? min(x, y) : max(x, y);
        }
    }
    return x;
}
// x holds a bitonic sequence across the warp -> ascending
__device__ __forceinline__ int32_t bitonic_merge32_asc(int32_t x, int lane) {
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        int32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        x = ((lane & j) == 0) ? min(x, y) : max(x, y);
    }
    return x;
}

// ascending bitonic sort of the first K lanes (K = power of two >= number of real entries; the rest hold INT_MAX)
__device__ __forceinline__ int32_t bitonic_sort_bounded_asc(int32_t x, int lane, int K) {
    for (int k = 2; k <= K; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int32_t y = __shfl_xor_sync(0xffffffffu, x, j);
            const bool up = ((lane & k) == 0);
            const bool lower = ((lane & j) == 0);
            x = (lower == up) ? min(x, y) : max(x, y);
        }
    }
    return x;
}

// One warp handles 32 voxel slots of one frame: lane <-> voxel for the metadata, then the voxels are materialised four
// at a time so that the CSR loads, the point gathers and the 512-byte row stores of four voxels are in flight together
// (each voxel needs two dependent loads).  The 32 slots are NOT consecutive: lane l of warp w takes slot
// (l / kGatherRun * warps_per_frame + w) * kGatherRun + l % kGatherRun.  First-seen order puts the crowded near-field
// pillars (n > 32: extra CSR chunks + sort/merge rounds, ~10x the work of a 1..4-point pillar) into the lowest ranks of
// every frame, so consecutive slots gave a few warps 32 heavy pillars each and the kernel waited for them (ncu: SMs
// active 57 % of the elapsed cycles); strided runs of kGatherRun slots give every warp at most kGatherRun of them and the same
// number of live voxels: 60 -> 37 us on the headline batch (runs of 1 and 4 measure the same; 4 keeps 16-byte metadata
// segments).
#ifndef HVPR_GATHER_RUN
#define HVPR_GATHER_RUN 4
#endif
constexpr int kGatherRun = HVPR_GATHER_RUN;      // consecutive slots per run (power of two <= 32)
#ifndef HVPR_GATHER_LOCKSTEP
#define HVPR_GATHER_LOCKSTEP 1
#endif
constexpr int kGatherGroup = 4;
template <bool kVec4>
__global__ void __launch_bounds__(256) vox_gather_kernel(const float *__restrict__ pts, int stride, int xyz_col,
                                                         const int32_t *__restrict__ frame_off, int B, HvprGeom g,
                                                         int64_t cells, int max_vox, int max_points,
                                                         const long long *__restrict__ keys,
                                                         const int32_t *__restrict__ vox_cell,
                                                         const int32_t *__restrict__ seg_off,
                                                         const int32_t *__restrict__ cursor,
                                                         const int32_t *__restrict__ csr,
                                                         const int32_t *__restrict__ frame_nvox,
                                                         float *__restrict__ voxels, int32_t *__restrict__ coords,
                                                         int32_t *__restrict__ num_points,
                                                         int32_t *__restrict__ voxel_offsets,
                                                         int32_t *__restrict__ cell_map) {
    __shared__ int32_t s_base[65], s_nvox[64], s_start[64];
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int f = 0; f < B; ++f) { const int nv = frame_nvox[f]; s_base[f] = acc; s_nvox[f] = nv; s_start[f] = frame_off[f]; acc += nv; }
        s_base[B] = acc;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x <= B) voxel_offsets[threadIdx.x] = s_base[threadIdx.x];

    const int runs_per_frame = (max_vox + kGatherRun - 1) / kGatherRun;
    const int wpf = (runs_per_frame + 32 / kGatherRun - 1) / (32 / kGatherRun);      // warps per frame
    const int64_t total_warps = (int64_t)B * wpf;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float4 *vox4 = reinterpret_cast<float4 *>(voxels);
    for (int64_t gw = warp0; gw < total_warps; gw += nwarps) {
        // ---- lane <-> voxel metadata -------------------------------------------------------------------------
        const int f = (int)(gw / wpf);
        const int wl = (int)(gw - (int64_t)f * wpf);
        const int run = (lane / kGatherRun) * wpf + wl;
        const int v = run * kGatherRun + (lane % kGatherRun);
        int n = 0, off = 0, cell = 0;
        int64_t row = 0;
        if (run < runs_per_frame && v < s_nvox[f]) {
            const int64_t w = (int64_t)f * max_vox + v;
            n = cursor[w];
            off = s_start[f] + seg_off[w];
            cell = vox_cell[w];
            row = (int64_t)s_base[f] + v;
        }
        const uint32_t live = __ballot_sync(0xffffffffu, n > 0);
        if (live == 0) continue;
        const int kept_mine = n < max_points ? n : max_points;
        if (n > 0) {
            // hashed mode: `cell` is a slot of the frame's key table, the key is the cell id
            const long long cid = keys ? keys[(int64_t)f * cells + cell] : (long long)cell;
            const int32_t cx = (int32_t)(cid % g.grid[0]);
            const int32_t cy = (int32_t)((cid / g.grid[0]) % g.grid[1]);
            const int32_t cz = (int32_t)(cid / ((long long)g.grid[0] * g.grid[1]));
            reinterpret_cast<int4 *>(coords)[row] = make_int4(f, cz, cy, cx);
            num_points[row] = kept_mine;
            if (cell_map) cell_map[(int64_t)f * cells + cell] = (int32_t)row;
        }
        // ---- materialise the rows, kGatherGroup voxels at a time -----------------------------------------------
        // live voxels form a prefix of every frame's slot range, so walking set bits keeps groups dense
        uint32_t todo = live;
        while (todo) {
            int src[kGatherGroup], nn[kGatherGroup], oo[kGatherGroup];
            int64_t rr[kGatherGroup];
            int32_t idx[kGatherGroup];
#pragma unroll
            for (int u = 0; u < kGatherGroup; ++u) {
                src[u] = todo ? (__ffs(todo) - 1) : -1;
                if (todo) todo &= todo - 1;
                const int sl = src[u] < 0 ? 0 : src[u];
                nn[u] = src[u] < 0 ? 0 : __shfl_sync(0xffffffffu, n, sl);
                oo[u] = __shfl_sync(0xffffffffu, off, sl);
                rr[u] = __shfl_sync(0xffffffffu, row, sl);
                idx[u] = (lane < nn[u]) ? __ldg(csr + oo[u] + lane) : INT_MAX;          // CSR loads of the group in flight
            }
#if HVPR_GATHER_LOCKSTEP
            {   // the four voxels of the group go through ONE bitonic network sized for the largest of them (lane predicates and
                // loop control shared, four independent shuffle chains in flight); rows with fewer entries hold INT_MAX above them
                int nmax = nn[0];
#pragma unroll
                for (int u = 1; u < kGatherGroup; ++u) nmax = nn[u] > nmax ? nn[u] : nmax;
                if (nmax > 1) {                                                          // warp-uniform
                    int K = 2;
                    while (K < nmax && K < 32) K <<= 1;
                    for (int k = 2; k <= K; k <<= 1)
                        for (int j = k >> 1; j > 0; j >>= 1) {
                            const bool take_min = (((lane & k) == 0) == ((lane & j) == 0));
#pragma unroll
                            for (int u = 0; u < kGatherGroup; ++u) {
                                const int32_t y = __shfl_xor_sync(0xffffffffu, idx[u], j);
                                idx[u] = take_min ? min(idx[u], y) : max(idx[u], y);
                            }
                        }
                }
            }
#pragma unroll
            for (int u = 0; u < kGatherGroup; ++u)
                for (int c0 = 32; c0 < nn[u]; c0 += 32) {                                // > 32 candidates: keep the 32 lowest
                    int32_t e = (c0 + lane < nn[u]) ? __ldg(csr + oo[u] + c0 + lane) : INT_MAX;
                    const int32_t worst = __shfl_sync(0xffffffffu, idx[u], 31);
                    if (!__any_sync(0xffffffffu, e < worst)) continue;
                    e = bitonic_sort32_asc(e, lane);
                    const int32_t er = __shfl_sync(0xffffffffu, e, 31 - lane);
                    idx[u] = bitonic_merge32_asc(min(idx[u], er), lane);
                }
#else
#pragma unroll
            for (int u = 0; u < kGatherGroup; ++u) {
                if (nn[u] > 1) {                                                         // warp-uniform
                    int K = 2;
                    while (K < nn[u] && K < 32) K <<= 1;
                    idx[u] = bitonic_sort_bounded_asc(idx[u], lane, K);
                    for (int c0 = 32; c0 < nn[u]; c0 += 32) {                            // > 32 candidates: keep the 32 lowest
                        int32_t e = (c0 + lane < nn[u]) ? __ldg(csr + oo[u] + c0 + lane) : INT_MAX;
                        const int32_t worst = __shfl_sync(0xffffffffu, idx[u], 31);
                        if (!__any_sync(0xffffffffu, e < worst)) continue;
                        e = bitonic_sort32_asc(e, lane);
                        const int32_t er = __shfl_sync(0xffffffffu, e, 31 - lane);
                        idx[u] = bitonic_merge32_asc(min(idx[u], er), lane);
                    }
                }
            }
#endif
            float4 pv[kGatherGroup];
#pragma unroll
            for (int u = 0; u < kGatherGroup; ++u) pv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            // point gathers of the group in flight (idx is frame-local)
#pragma unroll
            for (int u = 0; u < kGatherGroup; ++u) {
                const int kept = nn[u] < max_points ? nn[u] : max_points;
                if (lane < kept) {
                    const int64_t gi = (int64_t)s_start[f] + idx[u];
                    if (kVec4) pv[u] = __ldg(reinterpret_cast<const float4 *>(pts) + gi);
                    else {
                        const float *q = pts + gi * stride + xyz_col;
                        pv[u] = make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < kGatherGroup; ++u)
                if (src[u] >= 0 && lane < max_points) vox4[rr[u] * max_points + lane] = pv[u];
        }
    }
}

__global__ void frame_offsets_kernel(const float *__restrict__ pts, int64_t n, int stride, int B,
                                     int32_t *__restrict__ off) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i > n) return;
    // boundary between point i-1 and i: every frame id in (b_prev, b_cur] starts at i
    int b_prev = (i == 0) ? -1 : (int)pts[(i - 1) * stride];
    int b_cur = (i == n) ? B : (int)pts[i * stride];
    if (b_cur > B) b_cur = B;
    for (int f = b_prev + 1; f <= b_cur; ++f)
        if (f >= 0 && f <= B) off[f] = (int32_t)i;
}

}  // namespace hvpr

using namespace hvpr;

// the dense {first, count} table is used while it stays below this size; beyond it (or beyond 2^31 cells) the open-addressing table
constexpr size_t kDenseTableLimit = (size_t)4 << 30;
static bool dense_allowed(int64_t cells, int n_frames) {
    return cells <= INT_MAX && (size_t)cells * (size_t)(n_frames > 0 ? n_frames : 1) * sizeof(int2) <= kDenseTableLimit;
}
static int64_t hash_slots(int64_t max_frame_points) {           // power of two >= 2 x points of the largest frame
    int64_t s = 1024;
    while (s < 2 * max_frame_points) s <<= 1;
    return s;
}

extern "C" size_t hvpr_voxelize_workspace_bytes(int64_t n_total, int n_frames, const HvprGeom *geom, int max_voxels) {
    if (!geom || n_total < 0 || n_frames < 0 || max_voxels < 0) return 0;
    int64_t cells = (int64_t)geom->grid[0] * geom->grid[1] * geom->grid[2];
    int64_t bpf = ceil_div64(n_total > 0 ? n_total : 1, kScanTile);
    // enough for either table: the dense one when it is allowed, the open-addressing one always (a frame may hold all n_total points)
    const size_t hashed = carve(nullptr, n_total, n_frames, hash_slots(n_total), max_voxels, bpf, true).bytes;
    const size_t dense = dense_allowed(cells, n_frames) ? carve(nullptr, n_total, n_frames, cells, max_voxels, bpf, false).bytes : 0;
    return (dense > hashed ? dense : hashed) + 256;
}

extern "C" int hvpr_voxelize(const float *points, int64_t n_total, int pts_stride, int xyz_col,
                             const int32_t *frame_offsets, int n_frames, int64_t max_frame_points,
                             const HvprGeom *geom, int max_points, int max_voxels, int overflow_mode,
                             float *voxels, int32_t *coords, int32_t *num_points, int32_t *voxel_offsets,
                             int32_t *cell_map, void *workspace, size_t workspace_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!geom || !frame_offsets || !voxels || !coords || !num_points || !voxel_offsets || !workspace) return HVPR_ERR_ARG;
    if (n_total < 0 || n_frames <= 0 || pts_stride < 4 || xyz_col < 0 || xyz_col + 4 > pts_stride) return HVPR_ERR_ARG;
    if (n_total > 0 && !points) return HVPR_ERR_ARG;
    if (max_points < 1 || max_points > 32 || max_voxels < 1) return HVPR_ERR_UNSUPPORTED;
    if (n_frames > 64) return HVPR_ERR_UNSUPPORTED;
    const bool force_hash = (overflow_mode & HVPR_VOXELIZE_FORCE_HASH) != 0;
    overflow_mode &= ~HVPR_VOXELIZE_FORCE_HASH;
    if (overflow_mode != HVPR_OVERFLOW_CONTINUE && overflow_mode != HVPR_OVERFLOW_BREAK) return HVPR_ERR_ARG;
    const int64_t grid_cells = (int64_t)geom->grid[0] * geom->grid[1] * geom->grid[2];
    if (grid_cells <= 0 || n_total > INT_MAX || geom->grid[0] <= 0 || geom->grid[1] <= 0 || geom->grid[2] <= 0) return HVPR_ERR_UNSUPPORTED;
    if (max_frame_points <= 0 || max_frame_points > n_total) max_frame_points = n_total;
    const bool hashed = force_hash || !dense_allowed(grid_cells, n_frames);
    if (hashed && cell_map) return HVPR_ERR_ARG;             // the dense cell -> row map only exists for grids that have a dense table
    // table stride per frame: the grid's cells, or the slots of the open-addressing table
    const int64_t cells = hashed ? hash_slots(max_frame_points) : grid_cells;
    const int32_t frame_cap = (int32_t)max_frame_points;     // grids are sized from this bound; the kernels clamp every frame to it

    const int64_t bpf_ws = ceil_div64(n_total > 0 ? n_total : 1, kScanTile);
    // 256-B align the caller's pointer
    uintptr_t basep = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
    VoxWorkspace w = carve((void *)basep, n_total, n_frames, cells, max_voxels, bpf_ws, hashed);
    if (w.bytes + (basep - (uintptr_t)workspace) > workspace_bytes) return HVPR_ERR_WORKSPACE;

    {
        int64_t work = (int64_t)n_frames * cells / 2;
        int blocks = (int)(ceil_div64(work, 256) < 148 * 8 ? (ceil_div64(work, 256) > 0 ? ceil_div64(work, 256) : 1) : 148 * 8);
        vox_init_kernel<<<blocks, 256, 0, stream>>>(w.keys, w.table, (int64_t)n_frames * cells, cell_map,
                                                   (int64_t)n_frames * cells, w.cursor,
                                                   (int64_t)n_frames * max_voxels, w.agg, (int64_t)n_frames * bpf_ws,
                                                   w.frame_nvox, w.istar, w.ticket, n_frames);
        HVPR_CHECK_LAUNCH();
    }
    const bool vec4 = (pts_stride == 4 && xyz_col == 0 && ((uintptr_t)points % 16 == 0));
    if (max_frame_points > 0) {
        dim3 gridp((unsigned)ceil_div64(max_frame_points, 256 * kVoxPPT), (unsigned)n_frames);
        if (hashed) {
            if (vec4) vox_hash_kernel<true, true><<<gridp, 256, 0, stream>>>(points, pts_stride, xyz_col, frame_offsets, frame_cap, *geom, cells, w.keys, w.table, w.cellbuf);
            else vox_hash_kernel<false, true><<<gridp, 256, 0, stream>>>(points, pts_stride, xyz_col, frame_offsets, frame_cap, *geom, cells, w.keys, w.table, w.cellbuf);
        } else {
            if (vec4) vox_hash_kernel<true, false><<<gridp, 256, 0, stream>>>(points, pts_stride, xyz_col, frame_offsets, frame_cap, *geom, cells, w.keys, w.table, w.cellbuf);
            else vox_hash_kernel<false, false><<<gridp, 256, 0, stream>>>(points, pts_stride, xyz_col, frame_offsets, frame_cap, *geom, cells, w.keys, w.table, w.cellbuf);
        }
        HVPR_CHECK_LAUNCH();
        const int bpf = (int)ceil_div64(max_frame_points, kScanTile);
        dim3 grids((unsigned)bpf, (unsigned)n_frames);
        vox_assign_kernel<<<grids, kScanThreads, 0, stream>>>(frame_offsets, frame_cap, cells, w.table, w.cellbuf, w.agg, w.ticket, (int)bpf_ws,
                                                             max_voxels, w.vox_cell, w.seg_off, w.frame_nvox, w.istar);
        HVPR_CHECK_LAUNCH();
        vox_fill_kernel<<<gridp, 256, 0, stream>>>(frame_offsets, frame_cap, cells, w.table, w.cellbuf, max_voxels, w.seg_off,
                                                  w.cursor, w.csr, w.istar, overflow_mode == HVPR_OVERFLOW_BREAK);
        HVPR_CHECK_LAUNCH();
    }
    {
        const int64_t runs = ceil_div64(max_voxels, kGatherRun);
        int64_t want = ceil_div64((int64_t)n_frames * ceil_div64(runs, 32 / kGatherRun), 8);   // 8 warps per block, 32 slots per warp
        int blocks = (int)(want < 148 * 16 ? (want > 0 ? want : 1) : 148 * 16);
        if (vec4) vox_gather_kernel<true><<<blocks, 256, 0, stream>>>(points, pts_stride, xyz_col, frame_offsets, n_frames, *geom, cells, max_voxels, max_points, w.keys, w.vox_cell, w.seg_off, w.cursor, w.csr, w.frame_nvox, voxels, coords, num_points, voxel_offsets, cell_map);
        else vox_gather_kernel<false><<<blocks, 256, 0, stream>>>(points, pts_stride, xyz_col, frame_offsets, n_frames, *geom, cells, max_voxels, max_points, w.keys, w.vox_cell, w.seg_off, w.cursor, w.csr, w.frame_nvox, voxels, coords, num_points, voxel_offsets, cell_map);
        HVPR_CHECK_LAUNCH();
    }
    return HVPR_OK;
}

extern "C" int hvpr_frame_offsets(const float *points5, int64_t n_total, int pts_stride, int n_frames,
                                  int32_t *frame_offsets, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!frame_offsets || n_total < 0 || n_frames <= 0 || pts_stride < 1) return HVPR_ERR_ARG;
    if (n_total > 0 && !points5) return HVPR_ERR_ARG;
    if (n_total >= INT_MAX) return HVPR_ERR_UNSUPPORTED;
    int blocks = (int)ceil_div64(n_total + 1, 256);
    frame_offsets_kernel<<<blocks, 256, 0, stream>>>(points5, n_total, pts_stride, n_frames, frame_offsets);
    HVPR_CHECK_LAUNCH();
    return HVPR_OK;
}
