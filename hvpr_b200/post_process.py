"""Row N3 — single-stage post-processing with the reference's semantics: Detector3DTemplate.post_processing,
class-agnostic branch (pcdet/models/detectors/detector3d_template.py:168-260) + class_agnostic_nms
(pcdet/models/model_utils/model_nms_utils.py:6-25) + the rotated-BEV NMS the reference delegates to its (absent) `iou3d_nms` CUDA op.
All frames of the batch in four launches (hvpr_b200/csrc/nms.cu) through the C ABI; no CPU fallback."""
from __future__ import annotations

import torch

from . import _lib


def _get(cfg, k, default=None):
    return cfg.get(k, default) if hasattr(cfg, "get") else getattr(cfg, k, default)


class PostProcessor:
    def __init__(self, post_process_cfg):
        nms = _get(post_process_cfg, "NMS_CONFIG")
        self.multi_classes = bool(_get(nms, "MULTI_CLASSES_NMS", False))
        self.recall_thresh_list = list(_get(post_process_cfg, "RECALL_THRESH_LIST", [0.3, 0.5, 0.7]) or [])
        self.score_thresh = float(_get(post_process_cfg, "SCORE_THRESH"))
        self.nms_thresh = float(_get(nms, "NMS_THRESH"))
        self.pre_max = int(_get(nms, "NMS_PRE_MAXSIZE"))
        self.post_max = int(_get(nms, "NMS_POST_MAXSIZE"))
        if self.pre_max > 4096:
            raise NotImplementedError("NMS_PRE_MAXSIZE <= 4096")
        self._bufs = {}

    def run(self, cls_preds, box_preds, cls_normalized=False):
        """cls_preds (B, N, C), box_preds (B, N, 7) fp32 CUDA -> fixed-capacity (B, post_max, ...) boxes / scores / labels / anchor
        index + per-frame counts, all on the device, no host sync (graph-capturable)."""
        if not cls_preds.is_cuda:
            raise _lib.HvprError("PostProcessor needs CUDA tensors; there is no CPU path")
        _lib.init_device()
        B, N, C = cls_preds.shape
        dev, L = cls_preds.device, _lib.lib()
        key = (B, N, str(dev))
        if key not in self._bufs:
            nb = L.hvpr_post_process_workspace_bytes(B, N)
            self._bufs[key] = dict(ws=torch.empty(nb, dtype=torch.uint8, device=dev),
                                   boxes=torch.zeros(B, self.post_max, 7, device=dev), scores=torch.zeros(B, self.post_max, device=dev),
                                   labels=torch.zeros(B, self.post_max, dtype=torch.int32, device=dev),
                                   index=torch.zeros(B, self.post_max, dtype=torch.int32, device=dev),
                                   count=torch.zeros(B, dtype=torch.int32, device=dev))
        b = self._bufs[key]
        st = L.hvpr_post_process(_lib.ptr(cls_preds.contiguous()), _lib.ptr(box_preds.contiguous()), B, N, C, int(cls_normalized),
                                 self.score_thresh, self.pre_max, self.post_max, self.nms_thresh, _lib.ptr(b["boxes"]),
                                 _lib.ptr(b["scores"]), _lib.ptr(b["labels"]), _lib.ptr(b["index"]), _lib.ptr(b["count"]),
                                 _lib.ptr(b["ws"]), b["ws"].numel(), _lib.cur_stream())
        _lib.check(st, "hvpr_post_process")
        return b

    def _multi_classes(self, batch_dict):
        """MULTI_CLASSES_NMS branch (detector3d_template.py:214-233 -> model_nms_utils.multi_classes_nms :28-65): per head and per class
        k: threshold on that class's scores, top NMS_PRE_MAXSIZE, NMS, first NMS_POST_MAXSIZE; results concatenated class by class.
        Every (head, class) pass runs the same four kernels on a one-column score view.  Labels: the head's `multihead_label_mapping`
        entry, or k + 1 for a single head (the reference builds `arange(1, num_class)` there, one entry short of its own assert at :222)."""
        cls, box = batch_dict["batch_cls_preds"], batch_dict["batch_box_preds"]
        normalized = bool(batch_dict.get("cls_preds_normalized", False))
        heads = cls if isinstance(cls, (list, tuple)) else [cls]
        mapping = batch_dict.get("multihead_label_mapping") if isinstance(cls, (list, tuple)) else None
        B = heads[0].shape[0]
        per_frame = [dict(b=[], s=[], l=[]) for _ in range(B)]
        start = 0
        for h, hc in enumerate(heads):
            hb = box[:, start:start + hc.shape[1]].contiguous()                     # :224
            for k in range(hc.shape[2]):
                out = self.run(hc[:, :, k:k + 1].contiguous(), hb, normalized)
                counts = out["count"].cpu().tolist()
                label = int(mapping[h][k]) if mapping is not None else k + 1
                for i, n in enumerate(counts):
                    per_frame[i]["b"].append(out["boxes"][i, :n].clone()); per_frame[i]["s"].append(out["scores"][i, :n].clone())
                    per_frame[i]["l"].append(torch.full((n,), label, dtype=torch.long, device=hc.device))
            start += hc.shape[1]
        return [{"pred_boxes": torch.cat(f["b"], 0), "pred_scores": torch.cat(f["s"], 0), "pred_labels": torch.cat(f["l"], 0)} for f in per_frame]

    def generate_recall_record(self, box_preds, recall_dict, batch_index, data_dict=None, thresh_list=None):
        """detector3d_template.py:277-318 for a single-stage detector (no 'rois'): count the ground-truth boxes of frame `batch_index`
        whose best 3-D IoU with a prediction exceeds each threshold.  One hvpr_boxes_iou3d launch + one small read-back."""
        if data_dict is None or "gt_boxes" not in data_dict:
            return recall_dict
        thresh_list = self.recall_thresh_list if thresh_list is None else thresh_list
        gt = data_dict["gt_boxes"][batch_index]
        if len(recall_dict) == 0:
            recall_dict = {"gt": 0}
            for t in thresh_list:
                recall_dict["roi_%s" % str(t)] = 0
                recall_dict["rcnn_%s" % str(t)] = 0
        k = gt.shape[0] - 1                                                        # strip the zero padding of the collated gt (:291-294)
        rows = gt.abs().sum(dim=1).cpu()
        while k > 0 and float(rows[k]) == 0:
            k -= 1
        gt = gt[:k + 1]
        if gt.shape[0] > 0:
            if box_preds.shape[0] > 0:
                a, b = box_preds[:, :7].contiguous().float(), gt[:, :7].contiguous().float()
                iou = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
                _lib.check(_lib.lib().hvpr_boxes_iou3d(_lib.ptr(a), a.shape[0], _lib.ptr(b), b.shape[0], _lib.ptr(iou), _lib.cur_stream()), "hvpr_boxes_iou3d")
                best = iou.max(dim=0)[0]
                for t in thresh_list:
                    recall_dict["rcnn_%s" % str(t)] += int((best > t).sum().item())
            recall_dict["gt"] += int(gt.shape[0])
        return recall_dict

    def post_processing_with_recall(self, batch_dict):
        """-> (pred_dicts, recall_dict), the first two results of Detector3DTemplate.post_processing (:168-275)."""
        pred_dicts = self.post_processing(batch_dict)
        recall_dict = {}
        for i, p in enumerate(pred_dicts):
            recall_dict = self.generate_recall_record(p["pred_boxes"], recall_dict, i, batch_dict, self.recall_thresh_list)
        return pred_dicts, recall_dict

    def post_processing(self, batch_dict):
        """-> pred_dicts: [{'pred_boxes' (K,7), 'pred_scores' (K), 'pred_labels' (K)}] per frame, as detector3d_template.py:255-260
        (one device->host read of the per-frame counts, like the reference's data-dependent indexing)."""
        if self.multi_classes:
            return self._multi_classes(batch_dict)
        b = self.run(batch_dict["batch_cls_preds"], batch_dict["batch_box_preds"], bool(batch_dict.get("cls_preds_normalized", False)))
        counts = b["count"].cpu().tolist()
        # fresh tensors, like the reference: the (B, post_max, ...) buffers of run() are overwritten by the next call
        return [{"pred_boxes": b["boxes"][i, :k].clone(), "pred_scores": b["scores"][i, :k].clone(), "pred_labels": b["labels"][i, :k].long(),
                 "pred_anchor_index": b["index"][i, :k].long()} for i, k in enumerate(counts)]
