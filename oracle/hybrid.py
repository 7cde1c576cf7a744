"""ORACLE (test infrastructure) — CPU restatement of the reference's VFE / memory attention / BEV scatter.

Written with the same torch CPU fp32 ops the reference itself dispatches, as plain functions over a weight dict
whose keys are the reference's state_dict names:
    vfe.pfn_layers.{0,1}.linear.weight, vfe.pfn_layers.{0,1}.norm.{weight,bias,running_mean,running_var}
    vfe.pfn_scale_layers.{0,1}.0.weight, vfe.pfn_scale_layers.{0,1}.1.{weight,bias,running_mean,running_var}
    map_to_bev_module.memory.weight
Pinned against the reference's own modules (oracle/ref_loader.py) by tests/test_oracle_cpu.py::test_oracle_vs_live_reference and the
fixtures under tests/golden/.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-3          # pillar_vfe.py:23, :162


def random_weights(seed: int = 0, num_filters=(32, 64), num_scale=(16, 32), in_feat=10, mem=(2000, 64),
                   randomize_bn: bool = True, vfe_scale: bool = True):
    """Seeded random-init weights under the reference's state_dict names (the generator lives with the other synthetic
    inputs in hvpr_b200/synth.py so that bench.py's GPU arm does not have to import this package)."""
    from hvpr_b200 import synth
    return synth.random_frontend_weights(seed, num_filters, num_scale, in_feat, mem, randomize_bn, vfe_scale)


def _bn_eval(x, w, prefix):
    return F.batch_norm(x, w[prefix + ".running_mean"], w[prefix + ".running_var"],
                        w[prefix + ".weight"], w[prefix + ".bias"], False, 0.0, BN_EPS)


def pfn_layer(x, w, i, last):
    """PFNLayer.forward, pillar_vfe.py:29-49 (the 50 000-row chunking at :30-35 is numerically identical)."""
    x = F.linear(x, w["vfe.pfn_layers.%d.linear.weight" % i])                                  # :37
    x = _bn_eval(x.permute(0, 2, 1), w, "vfe.pfn_layers.%d.norm" % i).permute(0, 2, 1)         # :39
    x = F.relu(x)                                                                              # :41
    x_max = torch.max(x, dim=1, keepdim=True)[0]                                               # :42
    if last:
        return x_max                                                                           # :45
    return torch.cat([x, x_max.repeat(1, x.shape[1], 1)], dim=2)                               # :47-48


def pillar_vfe(voxels, num_points, coords, w, voxel_size, pc_range, scale: bool = True):
    """PillarVFE_Scale.forward pillar_vfe.py:184-221 (scale=True) / PillarVFE.forward :94-124 (scale=False).
    voxels (P,T,4) f32; num_points (P,); coords (P,4) [b,z,y,x] (any numeric dtype, as E4).
    -> pillar_features (P,C), pillar_scale_features (P,32) or None, pillar_mask (P,T,1)."""
    voxels = voxels.float()
    vx, vy, vz = voxel_size[0], voxel_size[1], voxel_size[2]
    x_off = vx / 2 + pc_range[0]                                                               # :169-171
    y_off = vy / 2 + pc_range[1]
    z_off = vz / 2 + pc_range[2]
    n = num_points.type_as(voxels)
    points_mean = voxels[:, :, :3].sum(dim=1, keepdim=True) / n.view(-1, 1, 1)                 # :187
    f_cluster = voxels[:, :, :3] - points_mean                                                 # :188
    f_center = torch.zeros_like(voxels[:, :, :3])                                              # :190-193
    f_center[:, :, 0] = voxels[:, :, 0] - (coords[:, 3].to(voxels.dtype).unsqueeze(1) * vx + x_off)
    f_center[:, :, 1] = voxels[:, :, 1] - (coords[:, 2].to(voxels.dtype).unsqueeze(1) * vy + y_off)
    f_center[:, :, 2] = voxels[:, :, 2] - (coords[:, 1].to(voxels.dtype).unsqueeze(1) * vz + z_off)
    feats = torch.cat([voxels, f_cluster, f_center], dim=-1)                                   # :196,203 (USE_ABSLOTE_XYZ)
    T = feats.shape[1]
    mask = (num_points.int().unsqueeze(1) > torch.arange(T, dtype=torch.int).view(1, -1))      # :176-182
    mask = mask.unsqueeze(-1).type_as(voxels)                                                  # :207
    feats = feats * mask                                                                       # :208
    nl = sum(1 for k in w if k.startswith("vfe.pfn_layers.") and k.endswith(".linear.weight"))
    for i in range(nl):
        feats = pfn_layer(feats, w, i, last=(i == nl - 1))                                     # :209-210
    feats = feats.squeeze(1)                                                                   # :211 (E8: keep the pillar axis)
    if not scale:
        return feats, None, mask
    d_mean = torch.norm(points_mean, 2, 2, keepdim=True)                                       # :213
    sf = torch.cat((n.unsqueeze(1), d_mean.squeeze(1), points_mean.squeeze(1)), dim=-1)        # :214
    for i in range(2):                                                                         # :215-216
        sf = F.linear(sf, w["vfe.pfn_scale_layers.%d.0.weight" % i])
        sf = F.relu(_bn_eval(sf, w, "vfe.pfn_scale_layers.%d.1" % i))
    return feats, sf, mask


def memory_attention(pillars, mem_weight, k=20, return_indices=False):
    """MemoryUnit_Agg.forward eval branch, memory_module.py:60-77.  pillars (nv,C) -> (nv,C)."""
    nv, d = pillars.shape
    score = F.softmax(F.linear(pillars, mem_weight), dim=1)                                    # :64-65
    _, indices = torch.topk(score, k, dim=1)                                                   # :66
    memory_positive = mem_weight[indices]                                                      # :67
    p = pillars.unsqueeze(1).expand(nv, k, d)                                                  # :70
    agg = F.softmax((memory_positive * p).sum(dim=2), dim=1)                                   # :71-72
    out = (agg.unsqueeze(2).expand(nv, k, d) * memory_positive).sum(dim=1)                     # :73-74
    return (out, indices) if return_indices else out


def scatter_plain(pillar_features, coords, batch_size, nx, ny):
    """PointPillarScatter.forward, pointpillar_scatter.py:14-37."""
    C = pillar_features.shape[1]
    canv = torch.zeros(batch_size, C, ny * nx, dtype=pillar_features.dtype)
    for b in range(batch_size):
        m = coords[:, 0] == b
        tc = coords[m]
        idx = (tc[:, 1] + tc[:, 2] * nx + tc[:, 3]).long()                                     # :27-28
        canv[b][:, idx] = pillar_features[m].t()                                               # :31
    return canv.view(batch_size, C, ny, nx)


def scatter_agg_memory(pillar_features, pillar_scale_features, coords, mem_weight, batch_size, nx, ny, k=20):
    """PointPillarScatter_Agg_Memory_1_scale.forward eval branch, pointpillar_scatter.py:169-220.
    -> spatial_features (B,2C,ny,nx), spatial_scale_features (B,Cs,ny,nx), memory readout (sum P, C)."""
    C, Cs = pillar_features.shape[1], pillar_scale_features.shape[1]
    canv = torch.zeros(batch_size, 2 * C, ny * nx, dtype=pillar_features.dtype)
    canv_s = torch.zeros(batch_size, Cs, ny * nx, dtype=pillar_features.dtype)
    readout = torch.zeros_like(pillar_features)
    for b in range(batch_size):                                                                # :178
        m = coords[:, 0] == b                                                                  # :190
        tc = coords[m]
        idx = (tc[:, 1] + tc[:, 2] * nx + tc[:, 3]).long()                                     # :192-193
        pil = pillar_features[m]
        out = memory_attention(pil, mem_weight, k)                                             # :200
        readout[m] = out
        canv[b][:, idx] = torch.cat((pil.t(), out.t()), dim=0)                                 # :204,207
        canv_s[b][:, idx] = pillar_scale_features[m].t()                                       # :208
    return canv.view(batch_size, 2 * C, ny, nx), canv_s.view(batch_size, Cs, ny, nx), readout


def frontend(frames, geom, w, overflow="continue", voxelize=None):
    """Whole path on CPU: list of (N,4) frames -> dict with every intermediate, reference semantics."""
    from . import voxelize as ov
    import numpy as np
    vox, coords, nump = ov.voxelize_batch(frames, geom.range_f32, geom.voxel_f32, geom.max_points_per_voxel,
                                          geom.max_voxels, overflow, fn=voxelize or ov.voxelize_c)
    nx, ny, nz = geom.grid_size
    tv, tc, tn = torch.from_numpy(vox), torch.from_numpy(coords), torch.from_numpy(nump)
    with torch.no_grad():
        pf, psf, mask = pillar_vfe(tv, tn, tc, w, list(geom.voxel_size), geom.range_f32)
        sp, sps, ro = scatter_agg_memory(pf, psf, tc, w["map_to_bev_module.memory.weight"], len(frames), nx, ny)
    return dict(voxels=tv, voxel_coords=tc, voxel_num_points=tn, pillar_features=pf, pillar_scale_features=psf,
                pillar_mask=mask, memory_readout=ro, spatial_features=sp, spatial_scale_features=sps)
