"""ORACLE (test infrastructure) — import the reference's OWN hot-path modules from /root/reference.

Only usable where /root/reference exists (the build container); the GPU box never has it, so nothing in the
`-m gpu` tests, smoke() or bench.py calls this.  It is used by oracle/make_golden.py (to mint tests/golden/) and by
tests/test_oracle_cpu.py (CPU, live-reference cases skipped when the tree is absent) to pin oracle/hybrid.py and
oracle/voxelize_ref.c.

Nothing is copied: sources are read, patched IN MEMORY and exec'd into fresh module objects:
  (1) delete memory_module.py:75  (stray prose line -> SyntaxError)
  (2) pointpillar_scatter.py:133,200  self.memory(pillars.t(), self.k) -> self.memory(pillars.t(), None, self.k)
      (forward(self, input1, input2, k) at memory_module.py:29; the eval branch ignores input2, :61)
  (3) `from .memory_module import MemoryUnit_Agg` (pointpillar_scatter.py:3) resolved to the patched module
pillar_vfe.py is loaded unmodified under a stub package.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF = os.environ.get("HVPR_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "pcdet/models/backbones_3d/vfe/pillar_vfe.py"))


class Cfg(dict):
    """attribute-style cfg (the reference only does attribute reads: pillar_vfe.py:131-139, pointpillar_scatter.py:53-59)"""
    __getattr__ = dict.__getitem__


VFE_CFG = Cfg(NAME="PillarVFE_Scale", WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, USE_NORM=True,
              NUM_FILTERS=[32, 64], NUM_SCALE_FEATURES=[16, 32])                       # hvpr.yaml:69-75
BEV_CFG = Cfg(NAME="PointPillarScatter_Agg_Memory_1_scale", NUM_BEV_FEATURES=128, NUM_PT_FEATURES=64,
              NUM_SCALE_FEATURES=32, NUM_COORD_POINTS=3, NUM_K=20, NUM_M=2000, SHRINK_TH=0.0025)  # hvpr.yaml:77-85

_CACHE = {}


def load():
    """-> namespace with PFNLayer, PillarVFE, PillarVFE_Scale, MemoryUnit_Agg, PointPillarScatter,
    PointPillarScatter_Agg_Memory_1_scale  (the reference's classes)."""
    if "ns" in _CACHE:
        return _CACHE["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    sys.dont_write_bytecode = True
    vfe_dir = os.path.join(REF, "pcdet/models/backbones_3d/vfe")
    pkg = types.ModuleType("_hvpr_ref_vfe")
    pkg.__path__ = [vfe_dir]
    sys.modules["_hvpr_ref_vfe"] = pkg
    mods = {}
    for name in ("vfe_template", "pillar_vfe"):
        spec = importlib.util.spec_from_file_location("_hvpr_ref_vfe." + name, os.path.join(vfe_dir, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = m
        spec.loader.exec_module(m)
        mods[name] = m

    bev_dir = os.path.join(REF, "pcdet/models/backbones_2d/map_to_bev")
    src_mm = open(os.path.join(bev_dir, "memory_module.py")).read().split("\n")
    assert "Mem, (TxM) x (MxC) = TxC" in src_mm[74], "reference memory_module.py:75 changed"
    del src_mm[74]                                                                     # patch (1)
    mm = types.ModuleType("_hvpr_ref_memory_module")
    exec(compile("\n".join(src_mm), "memory_module.py[patched]", "exec"), mm.__dict__)
    sys.modules["_hvpr_ref_memory_module"] = mm

    src_sc = open(os.path.join(bev_dir, "pointpillar_scatter.py")).read()
    assert src_sc.count("self.memory(pillars.t(), self.k)") == 2
    src_sc = src_sc.replace("self.memory(pillars.t(), self.k)", "self.memory(pillars.t(), None, self.k)")   # patch (2)
    src_sc = src_sc.replace("from .memory_module import MemoryUnit_Agg",
                            "from _hvpr_ref_memory_module import MemoryUnit_Agg")                          # patch (3)
    sc = types.ModuleType("_hvpr_ref_pointpillar_scatter")
    exec(compile(src_sc, "pointpillar_scatter.py[patched]", "exec"), sc.__dict__)

    ns = types.SimpleNamespace(
        PFNLayer=mods["pillar_vfe"].PFNLayer, PillarVFE=mods["pillar_vfe"].PillarVFE,
        PillarVFE_Scale=mods["pillar_vfe"].PillarVFE_Scale, MemoryUnit_Agg=mm.MemoryUnit_Agg,
        PointPillarScatter=sc.PointPillarScatter,
        PointPillarScatter_Agg_Memory_1_scale=sc.PointPillarScatter_Agg_Memory_1_scale)
    _CACHE["ns"] = ns
    return ns


BACKBONE_CFG = Cfg(NAME="BaseBEVBackbone_Scale", LAYER_NUMS=[3, 3, 3], SFM_LAYER_NUMS=[3, 3, 3], LAYER_STRIDES=[1, 2, 2],
                   NUM_FILTERS=[128, 256, 512], NUM_SCALE_FILTERS=[32, 64, 128], UPSAMPLE_STRIDES=[1, 2, 4],
                   NUM_UPSAMPLE_FILTERS=[128, 128, 128])                                # hvpr.yaml:87-95


def load_backbone():
    """-> namespace with BaseBEVBackbone, BaseBEVBackbone_Scale, SpatialAttention (row N1 of SURVEY.md §8f).

    One in-memory patch (breakage B4): base_bev_backbone.py:220 instantiates `SpatialAttention` without
    importing it; the class lives next door in spatial_attention.py:47-63 and is injected into the module namespace.
    """
    if "bb" in _CACHE:
        return _CACHE["bb"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    sys.dont_write_bytecode = True
    d = os.path.join(REF, "pcdet/models/backbones_2d")
    sa = types.ModuleType("_hvpr_ref_spatial_attention")
    exec(compile(open(os.path.join(d, "spatial_attention.py")).read(), "spatial_attention.py", "exec"), sa.__dict__)
    bb = types.ModuleType("_hvpr_ref_base_bev_backbone")
    bb.__dict__["SpatialAttention"] = sa.SpatialAttention                               # patch B4
    exec(compile(open(os.path.join(d, "base_bev_backbone.py")).read(), "base_bev_backbone.py", "exec"), bb.__dict__)
    ns = types.SimpleNamespace(BaseBEVBackbone=bb.BaseBEVBackbone, BaseBEVBackbone_Scale=bb.BaseBEVBackbone_Scale,
                               SpatialAttention=sa.SpatialAttention)
    _CACHE["bb"] = ns
    return ns


def load_vis_voxel_kernel():
    """-> the reference's OWN in-tree point->voxel loop, `_points_to_bevmap_reverse_kernel` (tools/vis.py:8-60), jitted by
    numba exactly as the reference declares it (`@numba.jit(nopython=True)`).

    tools/vis.py cannot be imported as a module (it pulls cv2 / matplotlib / the broken pcdet package at :1-6), so the
    source text of that ONE function (decorator line to the line before `def points_to_bev`) is exec'd in memory with
    `numba` and `np` in scope — nothing is copied into the repo.  This loop is the same lineage as spconv's
    `points_to_voxel` (floor / per-axis reject in x,y,z order / reversed coor / dense coor_to_voxelidx table / first-seen
    id / `break` at max_voxels); the per-voxel point append is replaced by a height map, and `bev_map[-1]` counts the
    points of every cell.  It pins: cell arithmetic, reject order, reversed coords, first-seen rank, the `break` cap and
    per-cell point counts of oracle/voxelize_ref.c on code the reference itself ships and runs.
    """
    if "vis" in _CACHE:
        return _CACHE["vis"]
    path = os.path.join(REF, "tools/vis.py")
    if not os.path.isfile(path):
        raise RuntimeError("reference tree not present at %s" % REF)
    import numba
    import numpy as np
    lines = open(path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("@numba.jit"))
    assert lines[start + 1].startswith("def _points_to_bevmap_reverse_kernel("), "reference tools/vis.py:8-9 changed"
    end = next(i for i, l in enumerate(lines) if l.startswith("def points_to_bev("))
    ns = {"numba": numba, "np": np}
    exec(compile("\n".join(lines[start:end]), "tools/vis.py[%d:%d]" % (start + 1, end), "exec"), ns)
    _CACHE["vis"] = ns["_points_to_bevmap_reverse_kernel"]
    return _CACHE["vis"]


def run_vis_voxel_kernel(points, pc_range, voxel_size, max_voxels):
    """Calls the reference loop the way its own caller does (tools/vis.py:86-104: dtype casts, DHW table of -1, zero
    bev_map with one extra plane, height_lowers).  -> (coor_to_voxelidx (nz,ny,nx) int32, per-cell point counts (ny,nx) int32)."""
    import numpy as np
    k = load_vis_voxel_kernel()
    points = np.ascontiguousarray(points, dtype=np.float32)
    vs = np.array(voxel_size, dtype=points.dtype)                                      # vis.py:86-87
    cr = np.array(pc_range, dtype=points.dtype)                                        # vis.py:88-89
    shape = tuple(np.round((cr[3:] - cr[:3]) / vs).astype(np.int32).tolist())[::-1]    # vis.py:90-92
    table = -np.ones(shape=shape, dtype=np.int32)                                      # vis.py:93
    bshape = list(shape)
    bshape[0] += 1                                                                     # vis.py:95-96
    lowers = np.linspace(cr[2], cr[5], shape[0], endpoint=False)                       # vis.py:97-98
    bev = np.zeros(shape=bshape, dtype=points.dtype)                                   # vis.py:101
    k(points, vs, cr, table, bev, lowers, False, max_voxels)                           # vis.py:102-104
    return table, bev[-1].astype(np.int32)
