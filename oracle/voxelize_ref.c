/*
 * ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object.  hvpr_b200/ never does.
 *
 * CPU restatement of the point->voxel loop the reference calls at
 *   pcdet/datasets/processor/data_processor.py:50-67   (spconv.utils.VoxelGenerator[V2].generate)
 * spconv is an un-vendored, un-pinned third-party dependency (setup.py:41 bare 'spconv';
 * README.md:26-27 "install from traveller59/spconv"); its source is NOT in /root/reference.
 * The loop below restates the published spconv `points_to_voxel` algorithm and follows the
 * in-tree numba twin of the same loop, tools/vis.py:23-50:
 *     vis.py:26-29  grid = round((hi - lo) / voxel_size)
 *     vis.py:36-41  per axis j in x,y,z: c = floor((p[j]-lo[j])/vs[j]) in fp32; reject c<0 || c>=grid[j];
 *                   coordinate stored reversed (z,y,x)
 *     vis.py:44-50  dense coor_to_voxelidx table initialised to -1; unseen cell -> id = voxel_num;
 *                   if voxel_num >= max_voxels: `break` (vis.py:47-48; spconv 1.0 numba) — spconv >= 1.1
 *                   C++ uses `continue` instead (overflow_mode below)
 * plus the per-voxel append that vis.py replaces with a height map:
 *     num = num_points_per_voxel[id]; if (num < max_points) { voxels[id,num] = point; num++ }
 *
 * PINNED on reference-run code for everything tools/vis.py:8-60 contains: that numba loop is executed in the build
 * container (oracle/ref_loader.load_vis_voxel_kernel) and this function reproduces its coor_to_voxelidx table and its
 * per-cell point counts bev_map[-1] array-for-array (tests/test_oracle_cpu.py, tests/golden/vis_kernel_pins.json).
 * UNPINNED against upstream spconv itself (not vendored / installed): the `continue` mode and the max_points append
 * are checked against an independent dict-based Python model (oracle/voxelize.py::voxelize_py) only.
 * One deliberate difference: NaN coordinates are rejected here; in the reference loop `c < 0 or c >= grid` is false
 * for NaN and the int cast indexes out of bounds (undefined behaviour).
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off, no -ffast-math: fp32 sub/div/floor must be IEEE).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* returns the number of voxels produced, or -1 on bad arguments.
 *   points      : (n, stride) fp32, xyz in columns col0..col0+2, `nfeat` features copied from col0
 *   range       : lo_x lo_y lo_z hi_x hi_y hi_z (fp32)          vsize: vx vy vz (fp32)
 *   overflow_mode: 0 = continue (spconv>=1.1 / VoxelGeneratorV2), 1 = break (vis.py:47-48, spconv 1.0)
 *   voxels      : (max_voxels, max_points, nfeat) fp32, must be zero-filled by the caller (np.zeros in spconv)
 *   coords      : (max_voxels, 3) int32 (z,y,x)
 *   num_points  : (max_voxels,) int32, zero-filled
 *   table       : (gz*gy*gx) int32 scratch, filled with -1 by the caller
 *   point_voxel : optional (n,) int32: voxel id each point was STORED in, -1 if dropped   (may be NULL)
 *   point_slot  : optional (n,) int32: slot within the voxel, -1 if dropped                (may be NULL)
 */
int hvpr_oracle_voxelize(const float *points, int64_t n, int stride, int col0, int nfeat,
                         const float *range, const float *vsize,
                         int max_points, int max_voxels, int overflow_mode,
                         float *voxels, int32_t *coords, int32_t *num_points, int32_t *table,
                         int32_t *point_voxel, int32_t *point_slot)
{
    if (!points || !range || !vsize || !voxels || !coords || !num_points || !table) return -1;
    int32_t grid[3];
    for (int j = 0; j < 3; ++j) {
        float g = (range[3 + j] - range[j]) / vsize[j];        /* vis.py:26 (fp32, as spconv casts to points.dtype) */
        grid[j] = (int32_t)nearbyintf(g);                       /* vis.py:29 np.round (half-to-even) */
    }
    int voxel_num = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = points + i * (int64_t)stride + col0;
        if (point_voxel) point_voxel[i] = -1;
        if (point_slot) point_slot[i] = -1;
        int32_t coor[3];
        int failed = 0;
        for (int j = 0; j < 3; ++j) {                           /* vis.py:36-41 */
            float c = floorf((p[j] - range[j]) / vsize[j]);
            if (c < 0.0f || c >= (float)grid[j] || c != c) { failed = 1; break; }  /* NaN rejected: int cast would be UB */
            coor[2 - j] = (int32_t)c;
        }
        if (failed) continue;
        int64_t cell = ((int64_t)coor[0] * grid[1] + coor[1]) * grid[0] + coor[2];
        int32_t vid = table[cell];                              /* vis.py:44 */
        if (vid == -1) {
            vid = voxel_num;
            if (voxel_num >= max_voxels) {                      /* vis.py:47-48 */
                if (overflow_mode == 1) break; else continue;
            }
            voxel_num += 1;
            table[cell] = vid;
            coords[vid * 3 + 0] = coor[0];
            coords[vid * 3 + 1] = coor[1];
            coords[vid * 3 + 2] = coor[2];
        }
        int32_t num = num_points[vid];
        if (num < max_points) {
            memcpy(voxels + ((int64_t)vid * max_points + num) * nfeat, p, sizeof(float) * (size_t)nfeat);
            if (point_voxel) point_voxel[i] = vid;
            if (point_slot) point_slot[i] = num;
            num_points[vid] = num + 1;
        }
    }
    return voxel_num;
}

/* grid size helper so Python and C agree on the rounding (data_processor.py:56-57, vis.py:26-29) */
void hvpr_oracle_grid(const float *range, const float *vsize, int32_t *grid)
{
    for (int j = 0; j < 3; ++j) grid[j] = (int32_t)nearbyintf((range[3 + j] - range[j]) / vsize[j]);
}
