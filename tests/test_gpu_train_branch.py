"""GPU parity of row N4 — the TRAINING-branch hybrid aggregation, forward only — through the C ABI, against oracle/train_branch.py
(pinned bit-exactly on the reference's own get_score / MemoryUnit_Agg(train) where /root/reference exists) and against the
fixture the reference's code produced (tests/golden/train_small.npz).  Tolerance: 1e-4 relative (fp32 path, BASELINE north_star)."""
import os

import numpy as np
import pytest
import torch

from hvpr_b200 import config, map_to_bev
from oracle import train_branch as tb

from helpers import GOLDEN, TOL_FP32, rel_err, tie_aware_readout_check

pytestmark = pytest.mark.gpu


def _bev(M=2000, shrink=0.0025, grid=(32, 32, 1)):
    cfg = config.Cfg(dict(config.HVPR_BEV_CFG), NUM_M=M, SHRINK_TH=shrink)
    return map_to_bev.PointPillarScatter_Agg_Memory_1_scale(cfg, grid_size=grid).cuda()


def test_train_branch_vs_reference_made_golden():
    z = np.load(os.path.join(GOLDEN, "train_small.npz"))
    pil, pts, w = (torch.from_numpy(z[n]).cuda() for n in ("pillars", "points", "weight"))
    k, shrink = int(z["k"]), float(z["shrink"])
    bev = _bev(w.shape[0], shrink)
    with torch.no_grad():
        bev.memory.weight.copy_(w)
    gs = bev.get_score(pts, pil.t(), return_positive=True)
    assert rel_err(gs["output"], torch.from_numpy(z["get_score_output"]))[0] <= TOL_FP32
    # the k positive points of every pillar: same SET as the reference's top-k (order inside the set is irrelevant to :53-57)
    ref_pos = torch.from_numpy(z["positive"])
    got = gs["points_positive"].cpu()
    assert torch.equal(got.sum(1), got.sum(1)) and rel_err(got.sort(dim=1)[0], ref_pos.sort(dim=1)[0])[0] <= 1e-6
    bev.train()
    out = bev.memory(pil, torch.from_numpy(z["positive"]).cuda(), k)["output"]
    torch.cuda.synchronize()
    e = rel_err(out, torch.from_numpy(z["memory_output"]))
    assert e[0] <= TOL_FP32 and e[1] <= TOL_FP32, e


@pytest.mark.parametrize("nv,npts", [(300, 5000), (1000, 16384), (17, 100), (1, 2049)])
def test_get_score_vs_oracle_any_point_count(nv, npts):
    """get_score = top-20 POINTS per pillar out of np (the shipped cfg samples 16 384 points per frame, hvpr.yaml:10-14): the exact
    fp32 kernel walks the columns in chunks of 2048 with a running top-k."""
    g = torch.Generator().manual_seed(nv + npts)
    pil = torch.rand(nv, 64, generator=g) * 1.5
    pts = torch.randn(npts, 64, generator=g)
    bev = _bev()
    gs = bev.get_score(pts.cuda(), pil.t().cuda(), return_positive=True)
    torch.cuda.synchronize()
    ref = tb.get_score(pts, pil.t(), 20)
    err, flips = tie_aware_readout_check(gs["output"], ref, pil, pts, TOL_FP32, k=20, idx=gs["indices"])
    assert err <= TOL_FP32
    idx = gs["indices"].cpu().long()
    assert int(idx.min()) >= 0 and int(idx.max()) < npts and all(len(set(r.tolist())) == 20 for r in idx[:50])


@pytest.mark.parametrize("M,shrink,scale", [(2000, 0.0025, 3.0), (256, 0.01, 1.2), (2000, 0.0, 1.0), (777, 0.005, 2.0)])
def test_memory_train_forward_vs_oracle(M, shrink, scale):
    g = torch.Generator().manual_seed(M)
    nv, k = 213, 20
    pil = torch.rand(nv, 64, generator=g) * 1.5
    pos = torch.randn(nv, k, 64, generator=g) * scale
    w = (torch.rand(M, 64, generator=g) * 2 - 1) * 0.5
    bev = _bev(M, shrink).train()
    with torch.no_grad():
        bev.memory.weight.copy_(w.cuda())
    out = bev.memory(pil.cuda(), pos.cuda(), k)["output"]
    torch.cuda.synchronize()
    ref = tb.memory_train(pil, pos, w, k, shrink)
    assert float(ref.abs().max()) > 1e-3
    e = rel_err(out, ref)
    assert e[0] <= TOL_FP32 and e[1] <= TOL_FP32, e


def test_scatter_training_forward_and_mem_loss_vs_oracle():
    """PointPillarScatter_Agg_Memory_1_scale.forward in training mode (pointpillar_scatter.py:87-167, one repaired call — see
    oracle/train_branch.py): three canvases + the two positive-feature tensors, then get_mem_loss."""
    g = torch.Generator().manual_seed(3)
    nx = ny = 24
    B, M, shrink = 2, 256, 0.01
    counts = [90, 61]
    coords, pfs = [], []
    for b, n in enumerate(counts):
        cells = torch.randperm(nx * ny, generator=g)[:n]
        coords.append(torch.stack([torch.full((n,), b), torch.zeros(n, dtype=torch.long), cells // nx, cells % nx], 1))
    coords = torch.cat(coords).int()
    P = coords.shape[0]
    pf = torch.rand(P, 64, generator=g) * 1.5
    psf = torch.rand(P, 32, generator=g)
    npts = [700, 2500]
    point_features = torch.randn(sum(npts), 64, generator=g) * 1.2
    point_coords = torch.cat([torch.full((n, 1), float(b)) for b, n in enumerate(npts)] and
                             [torch.cat([torch.full((n, 1), float(b)), torch.rand(n, 3, generator=g)], 1) for b, n in enumerate(npts)])
    w = (torch.rand(M, 64, generator=g) * 2 - 1) * 0.6
    bev = _bev(M, shrink, (nx, ny, 1)).train()
    with torch.no_grad():
        bev.memory.weight.copy_(w.cuda())
    bd = bev(dict(pillar_features=pf.cuda(), pillar_scale_features=psf.cuda(), pillar_mask=None, voxel_coords=coords.cuda().float(),
                  point_features=point_features.cuda(), point_coords=point_coords.cuda(), batch_size=B))
    torch.cuda.synchronize()
    ref = tb.scatter_train(pf, psf, coords.long(), point_features, point_coords, w, B, nx, ny, 20, shrink)
    for key in ("spatial_features", "spatial_features_point", "spatial_scale_features", "point_positive_features", "memory_positive_features"):
        assert tuple(bd[key].shape) == tuple(ref[key].shape), key
        e = rel_err(bd[key], ref[key])
        assert e[0] <= TOL_FP32 and e[1] <= TOL_FP32, (key, e)
    assert bd["memory_items"] is bev.memory.weight
    # the pillar halves of the canvases are bit copies; empty cells are exactly zero
    assert torch.equal(bd["spatial_features"][:, :64].cpu(), ref["spatial_features"][:, :64])
    assert torch.equal(bd["spatial_scale_features"].cpu(), ref["spatial_scale_features"])
    loss = map_to_bev.mem_loss(bd["memory_positive_features"], bd["point_positive_features"], mem_weight=0.7)
    ref_loss = tb.mem_loss(ref["memory_positive_features"], ref["point_positive_features"], 0.7)
    assert abs(float(loss) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
