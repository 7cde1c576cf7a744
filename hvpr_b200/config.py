"""Attribute-style cfg shim.  The reference reads cfg entries as attributes of an EasyDict built from YAML
(pcdet/config.py:51-80); its modules only ever do attribute reads (pillar_vfe.py:131-139,152;
pointpillar_scatter.py:53-59), so any dict subclass with __getattr__ is a drop-in."""
from __future__ import annotations


class Cfg(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def get(self, k, default=None):          # EasyDict-compatible
        return dict.get(self, k, default)


# tools/cfgs/kitti_models/hvpr.yaml:69-75
HVPR_VFE_CFG = Cfg(NAME="PillarVFE_Scale", WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, USE_NORM=True,
                   NUM_FILTERS=[32, 64], NUM_SCALE_FEATURES=[16, 32])
# tools/cfgs/kitti_models/hvpr.yaml:77-85
HVPR_BEV_CFG = Cfg(NAME="PointPillarScatter_Agg_Memory_1_scale", NUM_BEV_FEATURES=128, NUM_PT_FEATURES=64,
                   NUM_SCALE_FEATURES=32, NUM_COORD_POINTS=3, NUM_K=20, NUM_M=2000, SHRINK_TH=0.0025)


def load_yaml_model_cfg(path: str):
    """Read MODEL.VFE / MODEL.MAP_TO_BEV (and the voxelizer entry of DATA_CONFIG.DATA_PROCESSOR) from an
    OpenPCDet-style YAML such as tools/cfgs/kitti_models/hvpr.yaml, with plain `yaml` (easydict is not needed)."""
    import yaml
    with open(path) as f:
        y = yaml.safe_load(f)
    out = {"VFE": Cfg(y["MODEL"]["VFE"]), "MAP_TO_BEV": Cfg(y["MODEL"]["MAP_TO_BEV"])}
    dc = y.get("DATA_CONFIG", {})
    out["POINT_CLOUD_RANGE"] = dc.get("POINT_CLOUD_RANGE")
    for p in dc.get("DATA_PROCESSOR", []):
        if p.get("NAME") == "transform_points_to_voxels":
            out["VOXELIZER"] = Cfg(p)
    return out
