"""Shared helpers for the parity tests (imports the ORACLE — allowed only under tests/)."""
import hashlib
import os

import numpy as np
import torch

from hvpr_b200.geometry import Geometry

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TOL_FP32 = 1e-4     # north_star: pillar / BEV features within 1e-4 relative (fp32 path)
TOL_BF16 = 1e-2     # north_star: within 1e-2 for the bf16 memory-attention variant


def sha(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.shape).encode() + str(a.dtype).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def rel_err(a: torch.Tensor, b: torch.Tensor):
    """(max|a-b| / max|b| , ||a-b||_2 / ||b||_2) — the metric fixed in SURVEY.md §8d."""
    a, b = a.double().cpu(), b.double().cpu()
    d = (a - b)
    return float(d.abs().max() / b.abs().max().clamp_min(1e-30)), float(d.norm() / b.norm().clamp_min(1e-30))


def load_small(name):
    z = np.load(os.path.join(GOLDEN, "small_%s.npz" % name))
    geom = Geometry(tuple(float(x) for x in z["geom_range"]), tuple(float(x) for x in z["geom_voxel"]),
                    int(z["max_points"]), int(z["max_voxels"]))
    sizes = z["frame_sizes"]
    off = np.r_[0, np.cumsum(sizes)]
    frames = [z["points"][off[i]:off[i + 1]] for i in range(len(sizes))]
    return z, geom, frames, str(z["overflow"]), int(z["wseed"])


def to_dev(frames, device="cuda"):
    pts = torch.from_numpy(np.ascontiguousarray(np.concatenate(frames, 0))).to(device)
    off = torch.tensor(np.r_[0, np.cumsum([len(f) for f in frames])], dtype=torch.int32, device=device)
    return pts, off
