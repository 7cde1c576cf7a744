"""Drop-in boundary #1 — the voxelizer.

`VoxelGenerator` keeps spconv's constructor and `generate()` contract as used by the reference at
pcdet/datasets/processor/data_processor.py:50-67; `voxelize_batch(batch_dict)` is the fast path that consumes the
collated `points` (sum N, 5) [b,x,y,z,r] already on the GPU (pcdet/datasets/dataset.py:161-166) and fills `voxels`,
`voxel_coords` [b,z,y,x] and `voxel_num_points` in exactly the order `collate_batch` would (dataset.py:159-166).
All arithmetic runs in hvpr_b200/csrc/voxelize.cu through the C ABI; there is no CPU path here.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .geometry import Geometry


class VoxelizerOutput:
    """Device-resident result with capacity-sized buffers; `voxel_offsets[-1]` is the live pillar count."""
    __slots__ = ("voxels", "coords", "num_points", "voxel_offsets", "cell_map", "n_frames", "max_rows")

    def __init__(self, voxels, coords, num_points, voxel_offsets, cell_map, n_frames):
        self.voxels, self.coords, self.num_points = voxels, coords, num_points
        self.voxel_offsets, self.cell_map, self.n_frames = voxel_offsets, cell_map, n_frames
        self.max_rows = voxels.shape[0]

    @property
    def n_pillars_dev(self):
        return self.voxel_offsets[self.n_frames:self.n_frames + 1]


class Voxelizer:
    """GPU voxelizer with preallocated buffers for up to `max_frames` frames of `max_total_points` points."""

    def __init__(self, geom: Geometry, overflow: str = "continue", device=None, table: str = "auto"):
        """table: "auto" = dense {first, count} table while it fits (every pillar grid of the reference's configs), else the
        open-addressing hash table; "hash" forces the latter (no dense cell map is produced then)."""
        assert table in ("auto", "hash")
        self.geom = geom
        self.overflow = overflow
        self.table = table
        self.device = torch.device(device if device is not None else "cuda")
        self._geom_c = _lib.make_geom(geom.range_f32, geom.voxel_f32, geom.grid_size)
        self._ws = None
        self._ws_key = None
        _lib.lib()

    def _workspace(self, n_total, n_frames):
        key = (int(n_total), int(n_frames))
        if self._ws is None or self._ws_key[0] < key[0] or self._ws_key[1] < key[1]:
            import ctypes
            nbytes = _lib.lib().hvpr_voxelize_workspace_bytes(key[0], key[1], ctypes.byref(self._geom_c),
                                                              self.geom.max_voxels)
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self._ws_key = key
        return self._ws

    def uses_hash_table(self, n_frames: int) -> bool:
        cells = self.geom.cells_per_frame
        return self.table == "hash" or cells > 2 ** 31 - 1 or cells * max(n_frames, 1) * 8 > (4 << 30)

    def alloc_output(self, n_frames: int, with_cell_map: bool = True) -> VoxelizerOutput:
        g, dev = self.geom, self.device
        with_cell_map = with_cell_map and not self.uses_hash_table(n_frames)
        rows = n_frames * g.max_voxels
        return VoxelizerOutput(
            torch.empty((rows, g.max_points_per_voxel, 4), dtype=torch.float32, device=dev),
            torch.empty((rows, 4), dtype=torch.int32, device=dev),
            torch.empty((rows,), dtype=torch.int32, device=dev),
            torch.empty((n_frames + 1,), dtype=torch.int32, device=dev),
            torch.empty((n_frames, g.cells_per_frame), dtype=torch.int32, device=dev) if with_cell_map else None,
            n_frames)

    def run(self, points: torch.Tensor, frame_offsets: torch.Tensor, n_frames: int, max_frame_points: int = 0,
            out: VoxelizerOutput | None = None, xyz_col: int | None = None) -> VoxelizerOutput:
        """points (n_total, 4|5) fp32 CUDA; frame_offsets (n_frames+1,) int32 CUDA.  Enqueues on the current stream."""
        import ctypes
        assert points.is_cuda and points.dtype == torch.float32 and points.is_contiguous()
        assert frame_offsets.dtype == torch.int32 and frame_offsets.is_cuda
        _lib.init_device()
        n_total, stride = points.shape
        if xyz_col is None:
            xyz_col = 1 if stride == 5 else 0
        if out is None:
            out = self.alloc_output(n_frames)
        ws = self._workspace(n_total, n_frames)
        st = _lib.lib().hvpr_voxelize(
            _lib.ptr(points), n_total, stride, xyz_col, _lib.ptr(frame_offsets), n_frames, int(max_frame_points),
            ctypes.byref(self._geom_c), self.geom.max_points_per_voxel, self.geom.max_voxels,
            _lib.OVERFLOW[self.overflow] | (_lib.VOXELIZE_FORCE_HASH if self.table == "hash" else 0), _lib.ptr(out.voxels), _lib.ptr(out.coords), _lib.ptr(out.num_points),
            _lib.ptr(out.voxel_offsets), _lib.ptr(out.cell_map), _lib.ptr(ws), ws.numel(), _lib.cur_stream())
        _lib.check(st, "hvpr_voxelize")
        return out

    def frame_offsets_from_batch_column(self, points5: torch.Tensor, n_frames: int) -> torch.Tensor:
        off = torch.empty((n_frames + 1,), dtype=torch.int32, device=points5.device)
        st = _lib.lib().hvpr_frame_offsets(_lib.ptr(points5), points5.shape[0], points5.shape[1], n_frames,
                                           _lib.ptr(off), _lib.cur_stream())
        _lib.check(st, "hvpr_frame_offsets")
        return off

    def voxelize_batch(self, batch_dict: dict, exact_shapes: bool = True) -> dict:
        """Fill batch_dict['voxels'|'voxel_coords'|'voxel_num_points'] from batch_dict['points'] (sum N,5) on the GPU.
        exact_shapes=True slices to the live pillar count (one 4-byte D2H read, API compatibility with collate_batch);
        False keeps capacity-sized tensors plus batch_dict['num_pillars_dev'] (no host sync; graph-capturable)."""
        pts = batch_dict["points"]
        B = int(batch_dict["batch_size"])
        off = self.frame_offsets_from_batch_column(pts, B)
        out = self.run(pts, off, B)
        batch_dict["voxel_offsets"] = out.voxel_offsets
        batch_dict["cell_map"] = out.cell_map
        batch_dict["num_pillars_dev"] = out.n_pillars_dev
        if exact_shapes:
            P = int(out.voxel_offsets[B].item())
            batch_dict["voxels"], batch_dict["voxel_coords"] = out.voxels[:P], out.coords[:P]
            batch_dict["voxel_num_points"] = out.num_points[:P]
        else:
            batch_dict["voxels"], batch_dict["voxel_coords"] = out.voxels, out.coords
            batch_dict["voxel_num_points"] = out.num_points
        return batch_dict


class VoxelGenerator:
    """spconv.utils.VoxelGenerator look-alike (ctor kwargs and generate() as at data_processor.py:50-67)."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000, overflow="continue",
                 device="cuda"):
        self._geom = Geometry(tuple(float(x) for x in np.asarray(point_cloud_range, dtype=np.float32)),
                              tuple(float(x) for x in np.asarray(voxel_size, dtype=np.float32)),
                              int(max_num_points), int(max_voxels))
        self._vox = Voxelizer(self._geom, overflow, device)
        self.voxel_size = np.asarray(voxel_size, dtype=np.float32)
        self.point_cloud_range = np.asarray(point_cloud_range, dtype=np.float32)
        self.grid_size = np.asarray(self._geom.grid_size, dtype=np.int64)
        self.max_num_points, self.max_voxels = int(max_num_points), int(max_voxels)

    def generate(self, points, max_voxels=None):
        """points (N, >=4) numpy fp32 -> (voxels (P,T,4), coordinates (P,3) [z,y,x] int32, num_points (P,) int32)."""
        assert max_voxels is None or max_voxels == self.max_voxels
        p = torch.from_numpy(np.ascontiguousarray(points[:, :4], dtype=np.float32)).to(self._vox.device)
        off = torch.tensor([0, p.shape[0]], dtype=torch.int32, device=p.device)
        out = self._vox.run(p, off, 1)
        P = int(out.voxel_offsets[1].item())
        return (out.voxels[:P].cpu().numpy(), out.coords[:P, 1:].contiguous().cpu().numpy(),
                out.num_points[:P].cpu().numpy())


class VoxelGeneratorV2(VoxelGenerator):
    """spconv's V2 returns a dict (data_processor.py:63-65)."""

    def generate(self, points, max_voxels=None):
        v, c, n = super().generate(points, max_voxels)
        return {"voxels": v, "coordinates": c, "num_points_per_voxel": n, "voxel_num": len(n)}
