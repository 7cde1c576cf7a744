"""ORACLE (test infrastructure) — mint tests/golden/vis_kernel_pins.json from the REFERENCE'S OWN voxel loop.

Runs `_points_to_bevmap_reverse_kernel` (tools/vis.py:8-60, loaded by oracle/ref_loader.load_vis_voxel_kernel) on
seeded synthetic frames and stores, per case, the sha256 of
  * the coor_to_voxelidx table (cell -> first-seen voxel id, -1 empty; `break` at max_voxels), and
  * the per-cell point counts bev_map[-1] (uncapped) and the same clipped to max_points = 32,
so that the C restatement (CPU tests, any box) and the CUDA voxelizer's cell map / counts (GPU box, no reference tree)
are both checked against numbers the reference's code produced.   Usage (build container only):
    python -m oracle.make_golden_vis
"""
import json
import os

import numpy as np

from hvpr_b200 import synth
from hvpr_b200.geometry import G1, G2, G3
from oracle import ref_loader

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [(g, d, n, mv) for g in ("G1", "G2") for d in ("U", "L") for n in (120000,) for mv in (5000, 40000)]
CASES += [("G3", "L", 300000, 80000), ("G2", "L", 16384, 40000)]
GEOM = {"G1": G1, "G2": G2, "G3": G3}


def case_frame(gname, dist, n):
    g = GEOM[gname]
    return synth.make_frame(dist, n, g.point_cloud_range, 1024, edge_cases=True)


def digest(a) -> str:
    import hashlib
    a = np.ascontiguousarray(a)
    return hashlib.sha256(str(a.shape).encode() + str(a.dtype).encode() + a.tobytes()).hexdigest()


def main():
    out = {}
    for gname, dist, n, mv in CASES:
        g = GEOM[gname]
        f = case_frame(gname, dist, n)
        table, counts = ref_loader.run_vis_voxel_kernel(f, g.range_f32, g.voxel_f32, mv)
        key = "%s/%s/%d/%d" % (gname, dist, n, mv)
        out[key] = {"P": int(table.max() + 1), "points_counted": int(counts.sum()),
                    "table_sha256": digest(table.reshape(-1)), "counts_sha256": digest(counts.reshape(-1)),
                    "counts_cap32_sha256": digest(np.minimum(counts, 32).reshape(-1))}
        print(key, out[key]["P"], out[key]["points_counted"])
    with open(os.path.join(HERE, "..", "tests", "golden", "vis_kernel_pins.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
