"""ORACLE (test infrastructure) — voxelizer restatements.  See oracle/voxelize_ref.c for the citation block.

Three independent statements of the same semantics:
  voxelize_c   — the C loop (fast; the CPU baseline leg times this one)
  voxelize_py  — a dict-based pure-Python model written from the prose spec (small inputs only)
  voxelize_np  — the *parallel* formulation the CUDA kernels use (first-index min -> is_first -> exclusive
                 scan rank -> in-cell ordinal slot -> caps), in NumPy; proves the reformulation is equivalent
                 to the serial loop before any GPU is involved.
All return (voxels (P,T,F) f32, coords (P,3) i32 [z,y,x], num_points (P,) i32).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c(force: bool = False) -> str:
    so = os.path.join(_HERE, "_build", "liboracle_vox.so")
    src = os.path.join(_HERE, "voxelize_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "all"], stdout=subprocess.DEVNULL)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c())
        _LIB.hvpr_oracle_voxelize.restype = ctypes.c_int
        _LIB.hvpr_oracle_voxelize.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p]
        _LIB.hvpr_oracle_grid.restype = None
        _LIB.hvpr_oracle_grid.argtypes = [ctypes.c_void_p] * 3
    return _LIB


def grid_size(pc_range, voxel_size):
    r = np.ascontiguousarray(pc_range, dtype=np.float32)
    v = np.ascontiguousarray(voxel_size, dtype=np.float32)
    g = np.zeros(3, dtype=np.int32)
    _lib().hvpr_oracle_grid(r.ctypes.data, v.ctypes.data, g.ctypes.data)
    return int(g[0]), int(g[1]), int(g[2])


def voxelize_c(points, pc_range, voxel_size, max_points=32, max_voxels=40000, overflow="continue",
               return_assignment=False):
    """points (N, >=3) fp32, all columns are copied as features (spconv copies the whole row)."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n, nfeat = pts.shape
    r = np.ascontiguousarray(pc_range, dtype=np.float32)
    v = np.ascontiguousarray(voxel_size, dtype=np.float32)
    gx, gy, gz = grid_size(r, v)
    voxels = np.zeros((max_voxels, max_points, nfeat), dtype=np.float32)
    coords = np.zeros((max_voxels, 3), dtype=np.int32)
    nump = np.zeros((max_voxels,), dtype=np.int32)
    table = np.full((gz * gy * gx,), -1, dtype=np.int32)
    pv = np.empty((n,), dtype=np.int32) if return_assignment else None
    ps = np.empty((n,), dtype=np.int32) if return_assignment else None
    mode = {"continue": 0, "break": 1}[overflow]
    p = _lib().hvpr_oracle_voxelize(
        pts.ctypes.data, n, nfeat, 0, nfeat, r.ctypes.data, v.ctypes.data, max_points, max_voxels, mode,
        voxels.ctypes.data, coords.ctypes.data, nump.ctypes.data, table.ctypes.data,
        pv.ctypes.data if pv is not None else None, ps.ctypes.data if ps is not None else None)
    assert p >= 0
    out = (voxels[:p], coords[:p], nump[:p])
    return out + (pv, ps) if return_assignment else out


def cell_table_c(points, pc_range, voxel_size, max_voxels=40000, overflow="break", max_points=None):
    """The C loop run WITHOUT a payload (nfeat = 0): -> (coor_to_voxelidx table (nz,ny,nx) int32, per-cell point counts
    (nz,ny,nx) int32, coords (P,3)).  With max_points=None the counts are uncapped — exactly what the reference's in-tree
    loop tools/vis.py:8-60 leaves in `coor_to_voxelidx` and `bev_map[-1]`; this is the function pinned against it."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n = pts.shape[0]
    r = np.ascontiguousarray(pc_range, dtype=np.float32)
    v = np.ascontiguousarray(voxel_size, dtype=np.float32)
    gx, gy, gz = grid_size(r, v)
    dummy = np.zeros((1,), dtype=np.float32)
    coords = np.zeros((max_voxels, 3), dtype=np.int32)
    nump = np.zeros((max_voxels,), dtype=np.int32)
    table = np.full((gz * gy * gx,), -1, dtype=np.int32)
    p = _lib().hvpr_oracle_voxelize(
        pts.ctypes.data, n, pts.shape[1], 0, 0, r.ctypes.data, v.ctypes.data,
        int(max_points) if max_points is not None else max(n, 1), max_voxels, {"continue": 0, "break": 1}[overflow],
        dummy.ctypes.data, coords.ctypes.data, nump.ctypes.data, table.ctypes.data, None, None)
    assert p >= 0
    counts = np.zeros_like(table)
    c = coords[:p]
    counts[(c[:, 0] * gy + c[:, 1]) * gx + c[:, 2]] = nump[:p]
    return table.reshape(gz, gy, gx), counts.reshape(gz, gy, gx), c


def voxelize_py(points, pc_range, voxel_size, max_points=32, max_voxels=40000, overflow="continue"):
    """Dict-based model written from the prose spec — deliberately NOT a transliteration
    of the C loop: cells are keyed by tuple, voxels are Python lists, caps are applied when reading out."""
    pts = np.asarray(points, dtype=np.float32)
    lo = np.asarray(pc_range, dtype=np.float32)[:3]
    hi = np.asarray(pc_range, dtype=np.float32)[3:]
    vs = np.asarray(voxel_size, dtype=np.float32)
    grid = np.round((hi - lo) / vs).astype(np.int64)
    order, members = [], {}
    for i in range(pts.shape[0]):
        c = np.floor((pts[i, :3] - lo) / vs)          # fp32 arithmetic (all operands are np.float32)
        ok = True
        for j in range(3):                             # x, then y, then z
            if not (c[j] >= 0 and c[j] < grid[j]):
                ok = False
                break
        if not ok:
            continue
        key = (int(c[2]), int(c[1]), int(c[0]))        # stored reversed: z, y, x
        if key not in members:
            if len(order) >= max_voxels:
                if overflow == "break":
                    break
                continue
            members[key] = []
            order.append(key)
        if len(members[key]) < max_points:
            members[key].append(i)
    P = len(order)
    voxels = np.zeros((P, max_points, pts.shape[1]), dtype=np.float32)
    coords = np.zeros((P, 3), dtype=np.int32)
    nump = np.zeros((P,), dtype=np.int32)
    for v, key in enumerate(order):
        idx = members[key]
        voxels[v, :len(idx)] = pts[idx]
        coords[v] = key
        nump[v] = len(idx)
    return voxels, coords, nump


def voxelize_np(points, pc_range, voxel_size, max_points=32, max_voxels=40000, overflow="continue"):
    """NumPy model of the PARALLEL formulation implemented by hvpr_b200/csrc/voxelize.cu."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n = pts.shape[0]
    lo = np.asarray(pc_range, dtype=np.float32)[:3]
    hi = np.asarray(pc_range, dtype=np.float32)[3:]
    vs = np.asarray(voxel_size, dtype=np.float32)
    grid = np.round((hi - lo) / vs).astype(np.int64)
    with np.errstate(invalid="ignore"):
        c = np.floor((pts[:, :3] - lo) / vs)
        valid = np.all((c >= 0) & (c < grid.astype(np.float32)), axis=1)
    ci = np.where(valid[:, None], c, 0).astype(np.int64)
    cell = (ci[:, 2] * grid[1] + ci[:, 1]) * grid[0] + ci[:, 0]
    cell = np.where(valid, cell, -1)
    ncell = int(grid.prod())
    idx = np.arange(n, dtype=np.int64)
    first = np.full(ncell, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(first, cell[valid], idx[valid])                      # pass 1: atomicMin of point index
    is_first = valid & (first[np.where(valid, cell, 0)] == idx)
    rank = np.cumsum(is_first) - is_first                               # pass 2: exclusive scan
    cell_rank = np.full(ncell, -1, dtype=np.int64)
    cell_rank[cell[is_first]] = rank[is_first]
    vox = np.where(valid, cell_rank[np.where(valid, cell, 0)], -1)
    keep = valid & (vox < max_voxels)
    if overflow == "break":
        over = np.nonzero(is_first & (rank == max_voxels))[0]
        if over.size:
            keep &= idx < over[0]
    # slot = ordinal among kept points of the same voxel, in point order
    kidx = idx[keep]
    order = np.lexsort((kidx, vox[keep]))
    sv = vox[keep][order]
    start = np.r_[0, np.nonzero(np.diff(sv))[0] + 1] if sv.size else np.zeros(0, dtype=np.int64)
    seg_start = np.repeat(start, np.diff(np.r_[start, sv.size])) if sv.size else start
    slot_sorted = np.arange(sv.size) - seg_start
    P = int(min(is_first.sum(), max_voxels))
    voxels = np.zeros((P, max_points, pts.shape[1]), dtype=np.float32)
    coords = np.zeros((P, 3), dtype=np.int32)
    nump = np.zeros((P,), dtype=np.int32)
    sel = slot_sorted < max_points
    voxels[sv[sel], slot_sorted[sel]] = pts[kidx[order][sel]]
    np.add.at(nump, sv[sel], 1)
    fsel = is_first & (rank < max_voxels)
    coords[rank[fsel]] = np.stack([ci[fsel, 2], ci[fsel, 1], ci[fsel, 0]], axis=1)
    return voxels, coords, nump


def voxelize_batch(frames, pc_range, voxel_size, max_points=32, max_voxels=40000, overflow="continue",
                   fn=voxelize_c):
    """Per-frame voxelization followed by the reference's collate (pcdet/datasets/dataset.py:159-166):
    voxels / counts concatenated, coords left-padded with the frame index -> [b, z, y, x]."""
    vs, cs, ns = [], [], []
    for b, f in enumerate(frames):
        v, c, k = fn(f, pc_range, voxel_size, max_points, max_voxels, overflow)
        vs.append(v)
        cs.append(np.pad(c, ((0, 0), (1, 0)), mode="constant", constant_values=b))
        ns.append(k)
    return np.concatenate(vs, 0), np.concatenate(cs, 0).astype(np.int32), np.concatenate(ns, 0)
