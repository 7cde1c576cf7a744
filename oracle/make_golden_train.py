"""ORACLE (test infrastructure) — mint tests/golden/train_small.npz from the REFERENCE'S OWN training-branch code (row N4):
`PointPillarScatter_Agg_Memory_1_scale.get_score` (pointpillar_scatter.py:67-83) and `MemoryUnit_Agg.forward` in training mode
(memory_module.py:31-59), both run unmodified through oracle/ref_loader.  A small memory (M = 256, shrink 0.01) keeps the
fixture small; inputs are scaled so that several memory items survive the hard shrinkage.
    python -m oracle.make_golden_train        (build container only)
"""
import os

import numpy as np
import torch

from oracle import ref_loader

HERE = os.path.dirname(os.path.abspath(__file__))


def inputs(seed=5, nv=50, npts=300, M=256, d=64):
    g = torch.Generator().manual_seed(seed)
    pillars = torch.rand(nv, d, generator=g) * 1.5
    points = torch.randn(npts, d, generator=g) * 1.2
    weight = (torch.rand(M, d, generator=g) * 2 - 1) * 0.6
    return pillars, points, weight


def main():
    ns = ref_loader.load()
    cfg = ref_loader.Cfg(dict(ref_loader.BEV_CFG), NUM_M=256, SHRINK_TH=0.01)
    bev = ns.PointPillarScatter_Agg_Memory_1_scale(cfg, grid_size=(32, 32, 1))
    pillars, points, weight = inputs()
    with torch.no_grad():
        bev.memory.weight.copy_(weight)
        gs = bev.get_score(points, pillars.t())
        _, idx = torch.topk(torch.nn.functional.softmax(points @ pillars.t(), dim=0), cfg.NUM_K, dim=0)     # :73-75, to export input2
        positive = points[idx].permute(1, 0, 2).contiguous()                                                # :76
        mem = bev.memory.train()(pillars, positive, cfg.NUM_K)
    att = mem["att"]
    print("surviving items per point: mean %.1f, max %d" % (float((att > 0).sum(1).float().mean()), int((att > 0).sum(1).max())))
    np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "train_small.npz"),
                        pillars=pillars.numpy(), points=points.numpy(), weight=weight.numpy(), k=cfg.NUM_K, shrink=cfg.SHRINK_TH,
                        get_score_output=gs["output"].numpy(), positive=positive.numpy(), memory_output=mem["output"].numpy())


if __name__ == "__main__":
    main()
