"""ORACLE (test infrastructure) — row N4 of SURVEY.md §8f: the TRAINING-branch hybrid aggregation, forward only.

Restated line by line from the reference (same torch CPU fp32 ops), each function citing what it follows:
  get_score            PointPillarScatter_Agg_Memory_1_scale.get_score   pointpillar_scatter.py:67-83
  hard_shrink_relu     memory_module.py:85-87
  memory_train         MemoryUnit_Agg.forward, training branch            memory_module.py:31-59
  scatter_train        PointPillarScatter_Agg_Memory_1_scale.forward, training branch   pointpillar_scatter.py:87-167
  mem_loss             AnchorHeadTemplate.get_mem_loss                    anchor_head_template.py:262-275

Pinning (tests/test_oracle_cpu.py, wherever /root/reference exists): get_score and memory_train are bit-identical to the reference's
OWN get_score method and MemoryUnit_Agg(train mode).forward, which run unmodified.  scatter_train CANNOT be pinned as a whole: the
reference's training forward calls `self.memory(pillars.t(), self.k)` (pointpillar_scatter.py:133) against the signature
`forward(self, input1, input2, k)` (memory_module.py:29) — a TypeError as shipped — and the tensor the memory unit's docstring asks
for as input2 ("k positive point features for each pillar", memory_module.py:30, shape (nv, k, d) at :33) is computed inside
get_score (:76, `points_positive`) but never returned.  The restatement passes exactly that tensor; everything else follows the
source.  PointNet++ (the producer of `point_features`) is absent from the reference tree, so point features are inputs.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def get_score(points, pillars_t, k=20, return_positive=False):
    """points (np, d), pillars_t (d, nv) -> output (nv, d) [, points_positive (nv, k, d), indices (k, nv)]   — :67-83"""
    score = torch.matmul(points, pillars_t)                                          # :72  np x nv
    score = F.softmax(score, dim=0)                                                  # :73
    _, indices = torch.topk(score.detach(), k, dim=0)                                # :75
    points_positive = points[indices.detach()].permute(1, 0, 2)                      # :76  nv x k x d
    agg_weight = torch.matmul(pillars_t.t().unsqueeze(1), points_positive.permute(0, 2, 1)).squeeze()   # :77
    if agg_weight.dim() == 1:                                                        # nv == 1: .squeeze() also drops the pillar axis
        agg_weight = agg_weight.unsqueeze(0)
    agg_weight = F.softmax(agg_weight, dim=1)                                        # :79
    output = (agg_weight.detach().unsqueeze(2) * points_positive).sum(dim=1)         # :80-81
    return (output, points_positive, indices) if return_positive else output


def hard_shrink_relu(x, lambd=0.0, epsilon=1e-12):
    return (F.relu(x - lambd) * x) / (torch.abs(x - lambd) + epsilon)                # memory_module.py:85-87


def memory_train(input1, input2, weight, k=20, shrink_thres=0.0025):
    """input1 (nv, d) pillars, input2 (nv, k, d) positive point features, weight (M, d) -> output (nv, d)   — memory_module.py:31-59"""
    nv, d = input1.size()
    points = input2.reshape(-1, d)                                                   # :34
    mem_trans = weight.permute(1, 0)                                                 # :35
    att_weight = F.softmax(F.linear(points, weight), dim=1)                          # :37-38
    if shrink_thres > 0:                                                             # :41-45
        att_weight = hard_shrink_relu(att_weight, lambd=shrink_thres)
        att_weight = F.normalize(att_weight, p=1, dim=1)
    memory_positive = F.linear(att_weight, mem_trans).reshape(nv, k, d)              # :49-50
    pillars = input1.unsqueeze(1).expand(nv, k, d)                                   # :53
    agg_weight = F.softmax((memory_positive * pillars).sum(dim=2), dim=1)            # :54-55
    return (agg_weight.detach().unsqueeze(2).expand(nv, k, d) * memory_positive).sum(dim=1)   # :56-57


def scatter_train(pillar_features, pillar_scale_features, coords, point_features, point_coords, mem_weight, batch_size, nx, ny,
                  k=20, shrink_thres=0.0025):
    """pointpillar_scatter.py:87-167 (see the module docstring for the one repaired call)."""
    C, Cs = pillar_features.shape[1], pillar_scale_features.shape[1]
    sp = torch.zeros(batch_size, 2 * C, ny * nx)
    sp_pt = torch.zeros(batch_size, 2 * C, ny * nx)
    sps = torch.zeros(batch_size, Cs, ny * nx)
    pos_pt, pos_mem = [], []
    for b in range(batch_size):                                                      # :103
        m = coords[:, 0] == b                                                        # :122
        mp = point_coords[:, 0] == b                                                 # :123
        tc = coords[m]
        idx = (tc[:, 1] + tc[:, 2] * nx + tc[:, 3]).long()                           # :126-127
        pil = pillar_features[m]
        points = point_features[mp]                                                  # :135
        out_pt, positive, _ = get_score(points, pil.t(), k, return_positive=True)    # :136
        out_mem = memory_train(pil, positive, mem_weight, k, shrink_thres)           # :137 (repaired arity)
        sp_pt[b][:, idx] = torch.cat((pil.t(), out_pt.t()), dim=0)                   # :141-142
        sp[b][:, idx] = torch.cat((pil.t(), out_mem.t()), dim=0)                     # :144-146
        sps[b][:, idx] = pillar_scale_features[m].t()                                # :147
        pos_pt.append(out_pt)
        pos_mem.append(out_mem)
    return dict(spatial_features=sp.view(batch_size, 2 * C, ny, nx), spatial_features_point=sp_pt.view(batch_size, 2 * C, ny, nx),
                spatial_scale_features=sps.view(batch_size, Cs, ny, nx), point_positive_features=torch.cat(pos_pt, 0),
                memory_positive_features=torch.cat(pos_mem, 0))                      # :156-167


def mem_loss(memory, target, mem_weight=1.0):
    """anchor_head_template.py:262-275: MSE(memory, target) / target.shape[0] * LOSS_WEIGHTS['mem_weight']"""
    return F.mse_loss(memory, target.detach()) / int(target.shape[0]) * mem_weight
