"""ORACLE (test infrastructure) — CPU restatement of row N3: Detector3DTemplate.post_processing, class-agnostic
branch (pcdet/models/detectors/detector3d_template.py:168-260) + class_agnostic_nms (pcdet/models/model_utils/model_nms_utils.py:6-25).

PARITY UNPINNED for the NMS itself: the reference calls `iou3d_nms_utils.nms_gpu`, a CUDA op declared in setup.py:53-62 whose source is
not in the tree and which cannot run here.  This file restates the published algorithm independently of hvpr_b200/csrc/nms.cu: float64,
world-frame convex clipping of rectangle A by the four edges of rectangle B (the CUDA kernel clips in B's local frame in fp32), greedy
suppression in descending score order.  The surrounding selection logic (sigmoid, max over classes, score threshold, top-k, first
NMS_POST_MAXSIZE, label = argmax + 1) follows the Python source line by line.  Never imported by hvpr_b200/.
"""
from __future__ import annotations

import numpy as np

POST_CFG = dict(SCORE_THRESH=0.1, NMS_PRE_MAXSIZE=4096, NMS_POST_MAXSIZE=500, NMS_THRESH=0.1)      # hvpr.yaml:136-148


def corners(box):
    x, y, dx, dy, r = box[0], box[1], box[3], box[4], box[6]
    c, s = np.cos(r), np.sin(r)
    local = np.array([[dx / 2, dy / 2], [-dx / 2, dy / 2], [-dx / 2, -dy / 2], [dx / 2, -dy / 2]])
    return local @ np.array([[c, s], [-s, c]]) + np.array([x, y])            # counter-clockwise


def clip(poly, a, b):
    """keep the part of convex polygon `poly` on the left of the directed edge a -> b"""
    out = []
    n = len(poly)
    for i in range(n):
        p, q = poly[i], poly[(i + 1) % n]
        sp = (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])
        sq = (b[0] - a[0]) * (q[1] - a[1]) - (b[1] - a[1]) * (q[0] - a[0])
        if sp >= 0:
            out.append(p)
        if (sp > 0 and sq < 0) or (sp < 0 and sq > 0):
            t = sp / (sp - sq)
            out.append(p + t * (q - p))
    return out


def iou_bev(b1, b2):
    poly = list(corners(b1))
    cb = corners(b2)
    for i in range(4):
        if len(poly) < 3:
            return 0.0
        poly = clip(poly, cb[i], cb[(i + 1) % 4])
    if len(poly) < 3:
        return 0.0
    p = np.array(poly)
    inter = 0.5 * abs(np.sum(p[:, 0] * np.roll(p[:, 1], -1) - np.roll(p[:, 0], -1) * p[:, 1]))
    return inter / max(b1[3] * b1[4] + b2[3] * b2[4] - inter, 1e-8)


def post_process_frame(cls_preds, box_preds, cfg=POST_CFG, normalized=False, return_margin=False):
    """cls_preds (N, C) logits, box_preds (N, 7) -> (selected anchor indices, scores, labels) of one frame."""
    cls = cls_preds.astype(np.float64)
    prob = cls if normalized else 1.0 / (1.0 + np.exp(-cls))                  # :206-207
    scores = prob.max(-1)                                                     # :241
    labels = prob.argmax(-1) + 1
    keep_mask = scores.astype(np.float32) >= np.float32(cfg["SCORE_THRESH"])  # model_nms_utils.py:8-11
    idx = np.nonzero(keep_mask)[0]
    if idx.size == 0:
        return (np.zeros(0, np.int64),) * 3 + ((1.0,) if return_margin else ())
    order = np.lexsort((idx, -scores[idx].astype(np.float32)))                # topk, ties -> lower index first
    idx = idx[order][: cfg["NMS_PRE_MAXSIZE"]]                                # :15
    boxes = box_preds[idx].astype(np.float64)
    n = len(idx)
    suppressed = np.zeros(n, bool)
    kept = []
    margin = 1.0
    rad = 0.5 * np.hypot(boxes[:, 3], boxes[:, 4])
    for i in range(n):
        if suppressed[i]:
            continue
        kept.append(i)
        if len(kept) >= cfg["NMS_POST_MAXSIZE"]:
            break
        d = np.hypot(boxes[i + 1:, 0] - boxes[i, 0], boxes[i + 1:, 1] - boxes[i, 1])
        for j in np.nonzero((d <= rad[i + 1:] + rad[i]) & ~suppressed[i + 1:])[0] + i + 1:
            v = iou_bev(boxes[i], boxes[j])
            margin = min(margin, abs(v - cfg["NMS_THRESH"]))
            if v > cfg["NMS_THRESH"]:
                suppressed[j] = True
    sel = idx[np.array(kept[: cfg["NMS_POST_MAXSIZE"]], dtype=np.int64)]       # :20
    out = (sel, scores[sel], labels[sel])
    return out + (margin,) if return_margin else out


def random_detections(seed, n, n_clusters=150, extent=(0.0, -39.68, 69.12, 39.68)):
    """clustered car-sized boxes + logits: dense overlaps inside a cluster, none between clusters"""
    rng = np.random.default_rng(seed)
    cx = rng.uniform(extent[0], extent[2], n_clusters)
    cy = rng.uniform(extent[1], extent[3], n_clusters)
    which = rng.integers(0, n_clusters, n)
    box = np.zeros((n, 7), np.float32)
    box[:, 0] = cx[which] + rng.normal(0, 0.6, n)
    box[:, 1] = cy[which] + rng.normal(0, 0.6, n)
    box[:, 2] = -1.0
    box[:, 3] = 3.9 * rng.uniform(0.85, 1.15, n)
    box[:, 4] = 1.6 * rng.uniform(0.85, 1.15, n)
    box[:, 5] = 1.56
    box[:, 6] = rng.uniform(-np.pi, np.pi, n)
    cls = rng.normal(-2.5, 1.5, (n, 1)).astype(np.float32)
    return cls, box


def multi_classes_frame(cls_preds, box_preds, cfg=POST_CFG, normalized=False):
    """model_nms_utils.multi_classes_nms (:28-65) for one frame and one head: class by class the class-agnostic procedure on that class's
    score column; labels k + 1 (see hvpr_b200/post_process.py for the reference's off-by-one in the single-head label mapping)."""
    boxes, scores, labels = [], [], []
    for k in range(cls_preds.shape[1]):
        sel, sc, _ = post_process_frame(cls_preds[:, k:k + 1], box_preds, cfg, normalized)
        boxes.append(box_preds[sel]); scores.append(sc); labels.append(np.full(len(sel), k + 1, np.int64))
    return np.concatenate(boxes, 0), np.concatenate(scores, 0), np.concatenate(labels, 0)


def iou3d(b1, b2):
    """published boxes_iou3d_gpu: rotated-BEV overlap x height overlap over the union volume (z = box centre), float64"""
    inter_bev = iou_bev(b1, b2)
    a1, a2 = b1[3] * b1[4], b2[3] * b2[4]
    inter = inter_bev * (a1 + a2) / (1.0 + inter_bev)                       # invert iou = i / (a1 + a2 - i)
    top = min(b1[2] + b1[5] / 2, b2[2] + b2[5] / 2)
    bot = max(b1[2] - b1[5] / 2, b2[2] - b2[5] / 2)
    o3d = inter * max(top - bot, 0.0)
    return o3d / max(b1[3] * b1[4] * b1[5] + b2[3] * b2[4] * b2[5] - o3d, 1e-6)


def recall_record(pred_boxes, gt_boxes, thresh_list=(0.3, 0.5, 0.7)):
    """generate_recall_record (detector3d_template.py:277-318), single-stage case, one frame -> {'gt': n, 'rcnn_t': count}"""
    gt = np.asarray(gt_boxes, np.float64)
    k = len(gt) - 1
    while k > 0 and gt[k].sum() == 0:
        k -= 1
    gt = gt[:k + 1]
    out = {"gt": len(gt)}
    for t in thresh_list:
        out["rcnn_%s" % str(t)] = 0
    if len(gt) and len(pred_boxes):
        m = np.array([[iou3d(np.asarray(p, np.float64), g) for g in gt] for p in pred_boxes])
        best = m.max(0)
        for t in thresh_list:
            out["rcnn_%s" % str(t)] = int((best > t).sum())
    return out
