"""hvpr_b200 — B200-native (sm_100a) hybrid voxel-point encoding front end of HVPR.

Host side mirrors the reference's OpenPCDet interfaces for this path; the arithmetic lives in
hvpr_b200/csrc/*.cu behind the C ABI of include/hvpr_b200.h.  No CPU fallback exists.
"""
from . import config, geometry, synth  # noqa: F401
from .geometry import G1, G2, G3, Geometry  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # torch-dependent modules are imported lazily so `import hvpr_b200` stays cheap
    import importlib
    if name in ("vfe", "map_to_bev", "voxelizer", "frontend", "backbone", "dense_head", "post_process", "pipeline", "detector", "_lib"):
        return importlib.import_module("." + name, __name__)
    if name in ("Voxelizer", "VoxelGenerator", "VoxelGeneratorV2"):
        return getattr(importlib.import_module(".voxelizer", __name__), name)
    if name in ("PillarVFE", "PillarVFE_Scale"):
        return getattr(importlib.import_module(".vfe", __name__), name)
    if name in ("PointPillarScatter", "PointPillarScatter_Agg_Memory_1_scale", "MemoryUnit_Agg"):
        return getattr(importlib.import_module(".map_to_bev", __name__), name)
    if name in ("BaseBEVBackbone", "BaseBEVBackbone_Scale"):
        return getattr(importlib.import_module(".backbone", __name__), name)
    if name == "MixAnchor_Memory":
        return importlib.import_module(".detector", __name__).MixAnchor_Memory
    if name == "AnchorHeadSingle":
        return importlib.import_module(".dense_head", __name__).AnchorHeadSingle
    if name == "HybridFrontEnd":
        return importlib.import_module(".frontend", __name__).HybridFrontEnd
    raise AttributeError(name)
