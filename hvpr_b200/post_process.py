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
        if _get(nms, "MULTI_CLASSES_NMS", False):
            raise NotImplementedError("PostProcessor (B200): the class-agnostic NMS branch (hvpr.yaml:143-148) only")
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

    def post_processing(self, batch_dict):
        """-> pred_dicts: [{'pred_boxes' (K,7), 'pred_scores' (K), 'pred_labels' (K)}] per frame, as detector3d_template.py:255-260
        (one device->host read of the per-frame counts, like the reference's data-dependent indexing)."""
        b = self.run(batch_dict["batch_cls_preds"], batch_dict["batch_box_preds"], bool(batch_dict.get("cls_preds_normalized", False)))
        counts = b["count"].cpu().tolist()
        # fresh tensors, like the reference: the (B, post_max, ...) buffers of run() are overwritten by the next call
        return [{"pred_boxes": b["boxes"][i, :k].clone(), "pred_scores": b["scores"][i, :k].clone(), "pred_labels": b["labels"][i, :k].long(),
                 "pred_anchor_index": b["index"][i, :k].long()} for i, k in enumerate(counts)]
