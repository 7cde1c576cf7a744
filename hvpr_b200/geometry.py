"""Pillar-grid geometry shared by the voxelizer and the VFE / map_to_bev modules.

Follows the reference's conventions:
  * `point_cloud_range` is cast to fp32 (pcdet/datasets/dataset.py:25) and `voxel_size` to the points' dtype
    (spconv casts it), so every derived number here is an fp32 rounding of the YAML value.
  * grid_size = round((hi - lo) / voxel_size)  (pcdet/datasets/processor/data_processor.py:56-57,
    tools/vis.py:26-29), ordered (nx, ny, nz).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Geometry:
    point_cloud_range: tuple  # (lo_x, lo_y, lo_z, hi_x, hi_y, hi_z)
    voxel_size: tuple         # (vx, vy, vz)
    max_points_per_voxel: int = 32
    max_voxels: int = 40000

    @property
    def range_f32(self) -> np.ndarray:
        return np.asarray(self.point_cloud_range, dtype=np.float32)

    @property
    def voxel_f32(self) -> np.ndarray:
        return np.asarray(self.voxel_size, dtype=np.float32)

    @property
    def grid_size(self) -> tuple:
        """(nx, ny, nz) — fp32 arithmetic then round-half-even, as the voxelizer itself does."""
        r, v = self.range_f32, self.voxel_f32
        g = np.round((r[3:] - r[:3]) / v).astype(np.int64)
        return int(g[0]), int(g[1]), int(g[2])

    @property
    def cells_per_frame(self) -> int:
        nx, ny, nz = self.grid_size
        return nx * ny * nz


# G1: the shipped cfg (tools/cfgs/kitti_models/hvpr.yaml:5,23-28) — 296 x 248 x 1
G1 = Geometry((0.0, -19.84, -2.5, 47.36, 19.84, 0.5), (0.16, 0.16, 3.0), 32, 40000)
# G2: BASELINE.json configs[1] — classic PointPillars KITTI geometry, 432 x 496 x 1
G2 = Geometry((0.0, -39.68, -3.0, 69.12, 39.68, 1.0), (0.16, 0.16, 4.0), 32, 40000)
# G3: BASELINE.json configs[3] "extended range" as concretised by SURVEY.md §8d cfg 4 — 640 x 640 x 1, 80k pillars
G3 = Geometry((0.0, -51.2, -3.0, 102.4, 51.2, 1.0), (0.16, 0.16, 4.0), 32, 80000)

GEOMETRIES = {"G1": G1, "G2": G2, "G3": G3}
