"""Synthetic KITTI-range LiDAR frames (host-side input generation; not on the accelerated path).

Exactly the generators SURVEY.md Appendix A specifies, so occupancy statistics are reproducible:
points are fp32 rows [x, y, z, intensity]; seeds are `1024 + frame_index` (1024 echoes tools/test.py:47).
"""
from __future__ import annotations

import numpy as np


def frame_uniform(n: int, pc_range, seed: int) -> np.ndarray:
    """U: uniform in the range box."""
    rng = np.random.default_rng(seed)
    r = np.asarray(pc_range, dtype=np.float64)
    xyz = rng.uniform(r[0:3], r[3:6], size=(n, 3))
    inten = rng.uniform(0, 1, (n, 1))
    return np.concatenate([xyz, inten], axis=1).astype(np.float32)


def frame_lidar(n: int, pc_range, seed: int) -> np.ndarray:
    """L: LiDAR-like — dense near field, ground plane at z=-1.73, sparse far field."""
    rng = np.random.default_rng(seed)
    r = np.asarray(pc_range, dtype=np.float64)
    out, total = [], 0
    while total < n:
        m = 2 * n
        az = rng.uniform(-0.3 * np.pi, 0.3 * np.pi, m)
        el = np.deg2rad(rng.uniform(-24.8, 2.0, m))
        d_g = np.where(el < 0, 1.73 / np.maximum(np.tan(-el), 1e-3), 1e9)
        d_o = rng.gamma(shape=2.0, scale=12.0, size=m)
        d = np.minimum(d_g, d_o)
        x, y, z = d * np.cos(el) * np.cos(az), d * np.cos(el) * np.sin(az), d * np.sin(el)
        p = np.stack([x, y, z, rng.uniform(0, 1, m)], axis=1).astype(np.float32)
        keep = (p[:, 0] >= r[0]) & (p[:, 0] <= r[3]) & (p[:, 1] >= r[1]) & (p[:, 1] <= r[4])
        p = p[keep]
        out.append(p)
        total += len(p)
    return np.ascontiguousarray(np.concatenate(out, axis=0)[:n])


def inject_edge_cases(p: np.ndarray, pc_range) -> np.ndarray:
    """Boundary points / duplicates used by the parity tests (last paragraph)."""
    p = p.copy()
    r = np.asarray(pc_range, dtype=np.float32)
    p[::997, 0] = r[3]                       # x == hi  -> rejected (c == grid)
    p[1::991, 1] = r[1]                      # y == lo  -> cell 0
    p[2::983, 2] = r[5]                      # z == hi  -> rejected
    p[3::977] = p[0]                         # exact duplicates
    p[4::971, 0] = r[3] - np.float32(1e-6)   # just inside
    return p


def make_frame(dist: str, n: int, pc_range, seed: int, edge_cases: bool = False) -> np.ndarray:
    f = {"U": frame_uniform, "L": frame_lidar}[dist](n, pc_range, seed)
    return inject_edge_cases(f, pc_range) if edge_cases else f


def make_batch(dist: str, n: int, pc_range, batch: int, first_frame: int = 0, edge_cases: bool = False):
    """Returns (list of per-frame (n,4) arrays)."""
    return [make_frame(dist, n, pc_range, 1024 + first_frame + b, edge_cases) for b in range(batch)]


def collate_points(frames) -> np.ndarray:
    """pcdet/datasets/dataset.py:161-166 — left-pad every frame's points with the frame index -> (sum N, 5)."""
    cols = [np.pad(f, ((0, 0), (1, 0)), mode="constant", constant_values=i) for i, f in enumerate(frames)]
    return np.ascontiguousarray(np.concatenate(cols, axis=0).astype(np.float32))


def random_frontend_weights(seed: int = 0, num_filters=(32, 64), num_scale=(16, 32), in_feat=10, mem=(2000, 64),
                            randomize_bn: bool = True, vfe_scale: bool = True):
    """Random-init weights under the reference's state_dict names (nn.Linear Kaiming-uniform; memory U(+-1/sqrt(C)),
    memory_module.py:23-25) with BN affine + running stats randomised (default BN hides the padded-row term,
    SURVEY.md §8c).  Synthetic INPUT data for benchmarks and tests — there is no network for checkpoints."""
    import math

    import torch
    g = torch.Generator().manual_seed(seed)
    w = {}

    def lin(o, i):
        b = 1.0 / math.sqrt(i)            # kaiming_uniform_(a=sqrt(5)) bound == 1/sqrt(fan_in)
        return (torch.rand(o, i, generator=g) * 2 - 1) * b

    def bn(prefix, c):
        if randomize_bn:
            w[prefix + ".weight"] = torch.rand(c, generator=g) * 1.0 + 0.5
            w[prefix + ".bias"] = torch.randn(c, generator=g) * 0.5
            w[prefix + ".running_mean"] = torch.randn(c, generator=g) * 0.5
            w[prefix + ".running_var"] = torch.rand(c, generator=g) * 1.5 + 0.25
        else:
            w[prefix + ".weight"] = torch.ones(c)
            w[prefix + ".bias"] = torch.zeros(c)
            w[prefix + ".running_mean"] = torch.zeros(c)
            w[prefix + ".running_var"] = torch.ones(c)

    filt = [in_feat] + list(num_filters)
    for i in range(len(filt) - 1):
        last = i >= len(filt) - 2
        o = filt[i + 1] if last else filt[i + 1] // 2            # pillar_vfe.py:18-19
        w["vfe.pfn_layers.%d.linear.weight" % i] = lin(o, filt[i])
        bn("vfe.pfn_layers.%d.norm" % i, o)
    if vfe_scale:
        sc = [5] + list(num_scale)
        for i in range(len(sc) - 1):
            w["vfe.pfn_scale_layers.%d.0.weight" % i] = lin(sc[i + 1], sc[i])
            bn("vfe.pfn_scale_layers.%d.1" % i, sc[i + 1])
    if mem is not None:
        stdv = 1.0 / math.sqrt(mem[1])
        w["map_to_bev_module.memory.weight"] = (torch.rand(mem[0], mem[1], generator=g) * 2 - 1) * stdv
    return w
