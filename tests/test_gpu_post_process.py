"""GPU parity of row N3 (score threshold -> top-k -> rotated-BEV NMS -> first POST_MAXSIZE) against oracle/post_process.py, through
the C ABI.  The reference's own NMS op is not in the tree (parity unpinned upstream); the oracle is an independent float64 restatement,
so kept INDEX LISTS must be identical whenever no pair's IoU lies within 1e-4 of the threshold (the oracle reports that margin)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(cls_list, box_list, cfg, normalized=False):
    from hvpr_b200.config import Cfg
    from hvpr_b200.post_process import PostProcessor
    pp = PostProcessor(Cfg(SCORE_THRESH=cfg["SCORE_THRESH"], NMS_CONFIG=Cfg(MULTI_CLASSES_NMS=False, NMS_TYPE="nms_gpu",
                                                                          NMS_THRESH=cfg["NMS_THRESH"], NMS_PRE_MAXSIZE=cfg["NMS_PRE_MAXSIZE"],
                                                                          NMS_POST_MAXSIZE=cfg["NMS_POST_MAXSIZE"])))
    bd = {"batch_cls_preds": torch.from_numpy(np.stack(cls_list)).cuda(), "batch_box_preds": torch.from_numpy(np.stack(box_list)).cuda(),
          "cls_preds_normalized": normalized}
    out = pp.post_processing(bd)
    torch.cuda.synchronize()
    return out


def _good_seed(start, n, cfg, **kw):
    from oracle import post_process as op
    for seed in range(start, start + 20):
        cls, box = op.random_detections(seed, n, **kw)
        sel, sc, lab, margin = op.post_process_frame(cls, box, cfg, return_margin=True)
        if margin > 1e-4:
            return cls, box, sel, sc, lab
    raise AssertionError("no seed with a clear IoU margin")


@pytest.mark.parametrize("cfg", [
    dict(SCORE_THRESH=0.1, NMS_PRE_MAXSIZE=4096, NMS_POST_MAXSIZE=500, NMS_THRESH=0.1),        # hvpr.yaml:136-148
    dict(SCORE_THRESH=0.05, NMS_PRE_MAXSIZE=1000, NMS_POST_MAXSIZE=100, NMS_THRESH=0.5),
    dict(SCORE_THRESH=0.3, NMS_PRE_MAXSIZE=100, NMS_POST_MAXSIZE=10, NMS_THRESH=0.01),
])
def test_post_processing_matches_oracle(cfg):
    frames = [_good_seed(10, 20000, cfg), _good_seed(40, 20000, cfg, n_clusters=40)]
    out = _run([f[0] for f in frames], [f[1] for f in frames], cfg)
    for (cls, box, sel, sc, lab), o in zip(frames, out):
        got = o["pred_anchor_index"].cpu().numpy()
        assert got.tolist() == sel.tolist()                                        # same boxes, same (descending score) order
        assert np.allclose(o["pred_scores"].cpu().numpy(), sc, rtol=1e-6, atol=1e-7)
        assert np.array_equal(o["pred_boxes"].cpu().numpy(), box[sel])
        assert np.array_equal(o["pred_labels"].cpu().numpy(), lab)


def test_post_processing_edge_cases():
    from oracle import post_process as op
    cfg = dict(op.POST_CFG)
    # (1) no candidate at all, (2) fewer than one 64-box block, (3) more candidates than NMS_PRE_MAXSIZE with many equal scores
    cls0, box0 = op.random_detections(3, 5000)
    cls0[:] = -9.0
    cls1, box1 = op.random_detections(4, 5000)
    cls1[:] = -9.0
    cls1[[7, 4000, 4999, 123], 0] = [2.0, 1.0, 3.0, 0.5]
    box1[[7, 4000, 4999, 123], 0] = [5.0, 25.0, 45.0, 65.0]                         # far apart: nothing suppressed
    cls2, box2 = op.random_detections(5, 5000, n_clusters=5000)
    cls2[:] = 1.25                                                                  # 5000 identical scores: ties -> lower index first
    out = _run([cls0, cls1, cls2], [box0, box1, box2], cfg)
    assert len(out[0]["pred_scores"]) == 0
    assert out[1]["pred_anchor_index"].cpu().tolist() == [4999, 7, 4000, 123]
    sel2, sc2, _ = op.post_process_frame(cls2, box2, cfg)
    assert out[2]["pred_anchor_index"].cpu().tolist() == sel2.tolist() and sel2.max() < 4096


def test_post_processing_multi_class_labels_and_normalized_scores():
    from oracle import post_process as op
    cfg = dict(op.POST_CFG, NMS_THRESH=0.3)
    rng = np.random.default_rng(6)
    _, box = op.random_detections(7, 8000)
    prob = rng.uniform(0.0, 1.0, (8000, 3)).astype(np.float32) ** 4                 # already-normalised scores, 3 classes
    sel, sc, lab, margin = op.post_process_frame(prob, box, cfg, normalized=True, return_margin=True)
    assert margin > 1e-5
    out = _run([prob], [box], cfg, normalized=True)[0]
    assert out["pred_anchor_index"].cpu().tolist() == sel.tolist()
    assert np.array_equal(out["pred_labels"].cpu().numpy(), lab)


def test_points_to_detections_pipeline():
    """cfg 5 (full single-stage inference): raw points -> front end -> backbone -> head -> post-processing in one CUDA graph; the
    detections must equal the oracle's post-processing of the PIPELINE's own head output (each stage is pinned separately)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import load_small, to_dev
    from hvpr_b200 import synth
    from hvpr_b200.pipeline import HVPR_HEAD_CFG, HVPR_POST_CFG, FrontEndWithBackbone
    from oracle import backbone as ob, dense_head as od, hybrid, post_process as op
    z, geom, frames, overflow, wseed = load_small("tiny_continue")
    pipe = FrontEndWithBackbone(geom, overflow=overflow, head_cfg=HVPR_HEAD_CFG, post_cfg=HVPR_POST_CFG)
    pipe.frontend.load_reference_weights(hybrid.random_weights(wseed))
    pipe.backbone_2d.load_state_dict({k: torch.from_numpy(v) for k, v in ob.random_backbone_weights(21).items()}, strict=False)
    wh = od.random_head_weights(22)
    wh["conv_cls.bias"] = wh["conv_cls.bias"] + 1.0                     # enough candidates above the 0.1 threshold
    pipe.dense_head.load_state_dict({k: torch.from_numpy(v) for k, v in wh.items()})
    bd = pipe({"points": torch.from_numpy(synth.collate_points(frames)).cuda(), "batch_size": len(frames)})
    torch.cuda.synchronize()
    cls, box = bd["batch_cls_preds"].cpu().numpy(), bd["batch_box_preds"].cpu().numpy()
    assert len(bd["pred_dicts"]) == len(frames)
    total = 0
    for f, pd in enumerate(bd["pred_dicts"]):
        sel, sc, lab, margin = op.post_process_frame(cls[f], box[f], return_margin=True)
        if margin <= 1e-5:
            continue
        assert np.array_equal(pd["pred_boxes"].cpu().numpy(), box[f][sel])
        assert np.allclose(pd["pred_scores"].cpu().numpy(), sc, rtol=1e-6, atol=1e-7)
        total += len(sel)
    assert total > 0


def test_detector_object_loads_reference_names_and_matches_the_pipeline():
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import load_small
    from hvpr_b200 import synth
    from hvpr_b200.detector import MixAnchor_Memory
    from oracle import backbone as ob, dense_head as od, hybrid, post_process as op
    z, geom, frames, overflow, wseed = load_small("tiny_continue")
    det = MixAnchor_Memory(geom, overflow=overflow)
    wh = od.random_head_weights(22)
    wh["conv_cls.bias"] = wh["conv_cls.bias"] + 1.0
    state = dict(hybrid.random_weights(wseed))
    state.update({"backbone_2d." + k: torch.from_numpy(v) for k, v in ob.random_backbone_weights(21).items()})
    state.update({"dense_head." + k: torch.from_numpy(v) for k, v in wh.items()})
    state["backbone_3d.SA_modules.0.mlps.0.0.weight"] = torch.zeros(16, 4, 1, 1)      # train-only PointNet++ weights are ignored
    det.load_reference_state(state)
    assert {k.split(".")[0] for k in det.state_dict()} == {"vfe", "map_to_bev_module", "backbone_2d", "dense_head"}
    pred_dicts, recall_dicts, bd = det({"points": torch.from_numpy(synth.collate_points(frames)).cuda(), "batch_size": len(frames)})
    torch.cuda.synchronize()
    assert len(pred_dicts) == len(frames) and recall_dicts == {}
    cls, box = bd["batch_cls_preds"].cpu().numpy(), bd["batch_box_preds"].cpu().numpy()
    n = 0
    for f, pd in enumerate(pred_dicts):
        sel, sc, lab, margin = op.post_process_frame(cls[f], box[f], return_margin=True)
        if margin > 1e-5:
            assert np.array_equal(pd["pred_boxes"].cpu().numpy(), box[f][sel]) and np.array_equal(pd["pred_labels"].cpu().numpy(), lab)
            n += len(sel)
    assert n > 0
    with pytest.raises(KeyError):
        det.load_reference_state({"vfe.bogus": torch.zeros(1)})


def test_nms_exact_duplicates_flipped_duplicates_and_shared_edge_lines():
    """coincident / collinear edges.  Exact duplicates, heading + pi duplicates and same-heading boxes shifted along their
    own axes share whole edge lines; the boundary-integral IoU must count those edges exactly once (round 1 counted them 0, 1 or 2
    times depending on fp32 jitter, so 5-9 % of exact duplicates survived NMS).  Boxes sit up to 70 m from the origin, where an ulp
    of the world coordinate is 8e-6 m."""
    from oracle import post_process as op
    cfg = dict(op.POST_CFG)                                          # NMS_THRESH 0.1
    rng = np.random.default_rng(77)
    nb = 400
    gx, gy = np.meshgrid(np.linspace(3.0, 66.0, 20), np.linspace(-37.0, 37.0, 20))
    base = np.zeros((nb, 7), np.float32)
    base[:, 0] = gx.ravel() + rng.uniform(-0.2, 0.2, nb); base[:, 1] = gy.ravel() + rng.uniform(-0.2, 0.2, nb)
    base[:, 2] = -1.0
    base[:, 3] = rng.uniform(0.7, 1.4, nb); base[:, 4] = rng.uniform(0.4, 0.8, nb); base[:, 5] = 1.5      # small boxes: clusters stay apart
    base[:, 6] = rng.uniform(-np.pi, np.pi, nb)
    base[::7, 6] = 0.0; base[3::7, 6] = np.float32(np.pi / 2)        # axis-aligned ones too
    c, s_ = np.cos(base[:, 6]), np.sin(base[:, 6])
    def shifted(along, across):
        b = base.copy()
        b[:, 0] += (along * base[:, 3] * c - across * base[:, 4] * s_).astype(np.float32)
        b[:, 1] += (along * base[:, 3] * s_ + across * base[:, 4] * c).astype(np.float32)
        return b
    flip = base.copy(); flip[:, 6] += np.float32(np.pi)
    groups = [base, base.copy(), flip, shifted(0.3, 0.0), shifted(0.0, 0.5), shifted(0.97, 0.0), shifted(0.0, -0.96)]
    expect_kept = [True, False, False, False, False, True, True]    # IoU 1, 1, 0.54, 0.33 -> suppressed; 0.015, 0.02 -> kept
    box = np.concatenate(groups, 0).astype(np.float32)
    # scores: the base box of every cluster is the best one, companions get distinct lower scores
    cls = np.concatenate([np.full(nb, 3.0) - 0.001 * np.arange(nb)] +
                         [np.full(nb, 2.0 - 0.2 * g) - 0.001 * np.arange(nb) for g in range(1, len(groups))]).astype(np.float32)[:, None]
    sel, sc, lab, margin = op.post_process_frame(cls, box, cfg, return_margin=True)
    assert margin > 1e-3                                             # the construction keeps every IoU far from the threshold
    exp = set()
    for g, keep in enumerate(expect_kept):
        if keep:
            exp.update(range(g * nb, (g + 1) * nb))
    # post_max is 500: compare the oracle's list (truncated the same way) and check it is made of expected survivors only
    assert set(sel.tolist()) <= exp
    out = _run([cls], [box], cfg)[0]
    got = out["pred_anchor_index"].cpu().numpy()
    assert got.tolist() == sel.tolist()
    # and with room for every survivor: all duplicates / flipped duplicates / overlapping shifts are gone, all near-touching ones stay
    cfg2 = dict(cfg, NMS_POST_MAXSIZE=4096)
    sel2 = op.post_process_frame(cls, box, cfg2)[0]
    got2 = _run([cls], [box], cfg2)[0]["pred_anchor_index"].cpu().numpy()
    assert got2.tolist() == sel2.tolist() and set(got2.tolist()) == exp


def test_multi_classes_nms_branch_matches_oracle():
    """MULTI_CLASSES_NMS (detector3d_template.py:214-233, model_nms_utils.py:28-65): class-by-class NMS, results concatenated."""
    from hvpr_b200.config import Cfg
    from hvpr_b200.post_process import PostProcessor
    from oracle import post_process as op
    cfg = dict(SCORE_THRESH=0.1, NMS_PRE_MAXSIZE=1000, NMS_POST_MAXSIZE=50, NMS_THRESH=0.1)
    rng = np.random.default_rng(5)
    frames = []
    for seed in (60, 61):
        cls, box = op.random_detections(seed, 6000, n_clusters=60)
        cls3 = np.concatenate([cls, rng.normal(-2.5, 1.5, (6000, 2)).astype(np.float32)], 1)
        frames.append((cls3, box))
    pp = PostProcessor(Cfg(SCORE_THRESH=cfg["SCORE_THRESH"], NMS_CONFIG=Cfg(MULTI_CLASSES_NMS=True, NMS_TYPE="nms_gpu", NMS_THRESH=cfg["NMS_THRESH"],
                                                                          NMS_PRE_MAXSIZE=cfg["NMS_PRE_MAXSIZE"], NMS_POST_MAXSIZE=cfg["NMS_POST_MAXSIZE"])))
    out = pp.post_processing({"batch_cls_preds": torch.from_numpy(np.stack([f[0] for f in frames])).cuda(),
                              "batch_box_preds": torch.from_numpy(np.stack([f[1] for f in frames])).cuda(), "cls_preds_normalized": False})
    torch.cuda.synchronize()
    for (cls3, box), o in zip(frames, out):
        rb, rs, rl = op.multi_classes_frame(cls3, box, cfg)
        assert np.array_equal(o["pred_labels"].cpu().numpy(), rl)
        assert np.array_equal(o["pred_boxes"].cpu().numpy(), rb)
        assert np.allclose(o["pred_scores"].cpu().numpy(), rs, rtol=1e-6, atol=1e-7)
        assert set(rl.tolist()) == {1, 2, 3}


def test_boxes_iou3d_and_recall_record_match_oracle():
    """generate_recall_record (detector3d_template.py:277-318): pairwise 3-D IoU + per-threshold recall counts, zero-padded gt rows stripped."""
    from hvpr_b200 import _lib
    from hvpr_b200.config import Cfg
    from hvpr_b200.post_process import PostProcessor
    from oracle import post_process as op
    rng = np.random.default_rng(9)
    gt = np.zeros((12, 8), np.float32)
    gt[:9, 0] = rng.uniform(5, 60, 9); gt[:9, 1] = rng.uniform(-30, 30, 9); gt[:9, 2] = rng.uniform(-1.5, -0.5, 9)
    gt[:9, 3:6] = [3.9, 1.6, 1.56]; gt[:9, 6] = rng.uniform(-3, 3, 9); gt[:9, 7] = 1
    pred = np.concatenate([gt[:6, :7] + rng.normal(0, 0.08, (6, 7)).astype(np.float32),          # close hits
                           gt[6:8, :7] + np.array([1.2, 0.5, 0.3, 0, 0, 0, 0.4], np.float32),    # poor hits
                           rng.uniform(-5, 5, (5, 7)).astype(np.float32) + np.array([30, 0, -1, 4, 2, 1.5, 0], np.float32)], 0).astype(np.float32)
    pred[:, 3:6] = np.abs(pred[:, 3:6]) + 0.1
    a, b = torch.from_numpy(pred).cuda(), torch.from_numpy(gt[:9, :7].copy()).cuda()
    iou = torch.empty((len(pred), 9), device="cuda")
    _lib.check(_lib.lib().hvpr_boxes_iou3d(_lib.ptr(a), len(pred), _lib.ptr(b), 9, _lib.ptr(iou), _lib.cur_stream()))
    ref = np.array([[op.iou3d(p.astype(np.float64), g.astype(np.float64)) for g in gt[:9, :7]] for p in pred])
    assert np.abs(iou.cpu().numpy() - ref).max() <= 2e-5
    pp = PostProcessor(Cfg(SCORE_THRESH=0.1, RECALL_THRESH_LIST=[0.3, 0.5, 0.7],
                           NMS_CONFIG=Cfg(MULTI_CLASSES_NMS=False, NMS_TYPE="nms_gpu", NMS_THRESH=0.1, NMS_PRE_MAXSIZE=4096, NMS_POST_MAXSIZE=500)))
    rec = pp.generate_recall_record(a, {}, 0, {"gt_boxes": torch.from_numpy(gt).cuda()[None]}, [0.3, 0.5, 0.7])
    exp = op.recall_record(pred, gt[:, :7])
    assert rec["gt"] == exp["gt"] == 9
    for t in (0.3, 0.5, 0.7):
        assert rec["rcnn_%s" % t] == exp["rcnn_%s" % t], (t, rec, exp)
    assert rec["rcnn_0.3"] >= 6 and pp.generate_recall_record(a, {}, 0, {}, None) == {}
