"""ORACLE (test infrastructure) — mint the committed fixtures under tests/golden/.

Run in the build container, where /root/reference exists:   python -m oracle.make_golden
  tests/golden/small_<name>.npz  full tensors for small cases.  Voxelizer outputs come from the C restatement
                                 (cross-checked here against the independent dict model); VFE / memory / BEV tensors
                                 come from the REFERENCE'S OWN modules (oracle/ref_loader.py), not from oracle/hybrid.py.
  tests/golden/voxel_hashes.json sha256 of the integer / bit-copied voxelizer outputs for the full-size synthetic
                                 frames (G1/G2 x U/L x continue/break, N=120k; G3 L N=300k).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hvpr_b200 import synth                      # noqa: E402
from hvpr_b200.geometry import G1, G2, G3, Geometry   # noqa: E402
from oracle import hybrid, ref_loader             # noqa: E402
from oracle import voxelize as ov                 # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

TINY = Geometry((0.0, -3.2, -3.0, 6.4, 3.2, 1.0), (0.16, 0.16, 4.0), 32, 300)      # 40 x 40 x 1
TINY_T5 = Geometry((0.0, -3.2, -3.0, 6.4, 3.2, 1.0), (0.16, 0.16, 4.0), 5, 10000)


def sha(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.shape).encode() + str(a.dtype).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def small_case(name, geom, n, batch, dist, overflow, wseed):
    frames = synth.make_batch(dist, n, geom.point_cloud_range, batch, first_frame=100, edge_cases=True)
    # pin the C voxelizer against the dict model on this very input
    for f in frames:
        a = ov.voxelize_c(f, geom.range_f32, geom.voxel_f32, geom.max_points_per_voxel, geom.max_voxels, overflow)
        b = ov.voxelize_py(f, geom.range_f32, geom.voxel_f32, geom.max_points_per_voxel, geom.max_voxels, overflow)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), "C voxelizer != dict model"
    vox, coords, nump = ov.voxelize_batch(frames, geom.range_f32, geom.voxel_f32, geom.max_points_per_voxel,
                                          geom.max_voxels, overflow)
    w = hybrid.random_weights(wseed)
    ns = ref_loader.load()
    vfe = ns.PillarVFE_Scale(ref_loader.VFE_CFG, 4, list(geom.voxel_size), geom.range_f32).eval()
    bev = ns.PointPillarScatter_Agg_Memory_1_scale(ref_loader.BEV_CFG, grid_size=geom.grid_size).eval()
    vfe.load_state_dict({k[4:]: v for k, v in w.items() if k.startswith("vfe.")}, strict=False)
    bev.load_state_dict({"memory.weight": w["map_to_bev_module.memory.weight"]})
    # the reference sees fp32 coords / counts on the device (E4)
    bd = dict(voxels=torch.from_numpy(vox).clone(), voxel_num_points=torch.from_numpy(nump).float(),
              voxel_coords=torch.from_numpy(coords).float(), batch_size=batch)
    with torch.no_grad():
        bd = bev(vfe(bd))
        readout = torch.cat([bev.memory(bd["pillar_features"][bd["voxel_coords"][:, 0] == b], None, bev.k)["output"]
                             for b in range(batch)], 0)
    np.savez_compressed(
        os.path.join(GOLD, "small_%s.npz" % name),
        geom_range=geom.range_f32, geom_voxel=geom.voxel_f32, max_points=geom.max_points_per_voxel,
        max_voxels=geom.max_voxels, overflow=overflow, wseed=wseed,
        points=np.concatenate(frames, 0), frame_sizes=np.array([len(f) for f in frames]),
        voxels=vox, voxel_coords=coords, voxel_num_points=nump,
        pillar_features=bd["pillar_features"].numpy(), pillar_scale_features=bd["pillar_scale_features"].numpy(),
        memory_readout=readout.numpy(),
        spatial_features=bd["spatial_features"].numpy(), spatial_scale_features=bd["spatial_scale_features"].numpy())
    print("small", name, "P", len(nump), "K", int(nump.sum()))


def hashes():
    out = {}
    cases = [("G1", G1, 120000), ("G2", G2, 120000)]
    for gname, g, n in cases:
        for dist in "UL":
            f = synth.make_frame(dist, n, g.point_cloud_range, 1024, edge_cases=True)
            for mode in ("continue", "break"):
                v, c, k = ov.voxelize_c(f, g.range_f32, g.voxel_f32, g.max_points_per_voxel, g.max_voxels, mode)
                out["%s/%s/%d/%s" % (gname, dist, n, mode)] = dict(P=int(len(k)), K=int(k.sum()),
                                                                  sha256=sha(v.view(np.int32), c, k))
    f = synth.make_frame("L", 300000, G3.point_cloud_range, 1024, edge_cases=True)
    v, c, k = ov.voxelize_c(f, G3.range_f32, G3.voxel_f32, 32, G3.max_voxels, "continue")
    out["G3/L/300000/continue"] = dict(P=int(len(k)), K=int(k.sum()), sha256=sha(v.view(np.int32), c, k))
    with open(os.path.join(GOLD, "voxel_hashes.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    assert ref_loader.available(), "needs /root/reference"
    os.makedirs(GOLD, exist_ok=True)
    small_case("tiny_continue", TINY, 1500, 2, "L", "continue", 1)
    small_case("tiny_break_cap", TINY, 2500, 2, "U", "break", 2)
    small_case("tiny_t5", TINY_T5, 3000, 1, "L", "continue", 3)
    hashes()
