"""The single-stage detector the reference builds from hvpr.yaml (MODEL.NAME MixAnchor_Memory, pcdet/models/detectors/pointpillar.py:
eval forward = every module but the train-only PointNet++ branch, then post_processing), assembled from the B200 modules.

Sub-modules carry the reference's attribute names — `vfe`, `map_to_bev_module`, `backbone_2d`, `dense_head`
(Detector3DTemplate.module_topology, detector3d_template.py:30-33) — so a reference checkpoint's `model_state` loads unchanged
(`load_reference_state` drops the `backbone_3d.*` PointNet++ weights, which the eval path never touches, pointpillar.py:54).
The forward runs the whole chain as ONE CUDA graph: GPU voxelization of the collated points replaces the dataset's CPU voxelizer.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .pipeline import HVPR_BACKBONE_CFG, HVPR_HEAD_CFG, HVPR_POST_CFG, FrontEndWithBackbone


class MixAnchor_Memory(nn.Module):
    def __init__(self, geom, backbone_cfg=HVPR_BACKBONE_CFG, head_cfg=HVPR_HEAD_CFG, post_cfg=HVPR_POST_CFG, device="cuda",
                 **frontend_kwargs):
        super().__init__()
        pipe = FrontEndWithBackbone(geom, backbone_cfg=backbone_cfg, device=device, head_cfg=head_cfg, post_cfg=post_cfg, **frontend_kwargs)
        object.__setattr__(self, "_pipe", pipe)              # not registered: its sub-modules appear once, under the reference's names
        self.vfe = pipe.frontend.vfe
        self.map_to_bev_module = pipe.frontend.map_to_bev_module
        self.backbone_2d = pipe.backbone_2d
        self.dense_head = pipe.dense_head
        self.eval()

    def load_reference_state(self, model_state: dict):
        """model_state of a reference checkpoint (checkpoint['model_state'], detector3d_template.py:320-346)."""
        sd = {k: v for k, v in model_state.items() if not k.startswith("backbone_3d.")}
        r = self.load_state_dict(sd, strict=False)
        missing = [k for k in r.missing_keys if not k.endswith("num_batches_tracked")]
        if missing or r.unexpected_keys:
            raise KeyError("checkpoint does not match the detector: missing %s, unexpected %s" % (missing[:5], r.unexpected_keys[:5]))
        return self

    @torch.no_grad()
    def forward(self, batch_dict):
        """batch_dict['points'] (sum N, 5) [b,x,y,z,r] on the GPU + 'batch_size' -> (pred_dicts, recall_dicts, batch_dict),
        the eval return of pointpillar.py:52-56; recall_dicts follows generate_recall_record (detector3d_template.py:277-318) when
        batch_dict carries 'gt_boxes', else it is empty."""
        if self.training:
            raise NotImplementedError("hvpr_b200.MixAnchor_Memory implements the eval branch (pointpillar.py:52-56) only")
        batch_dict = self._pipe(batch_dict)
        recall = {}
        if "gt_boxes" in batch_dict:
            for i, p in enumerate(batch_dict["pred_dicts"]):
                recall = self._pipe.post.generate_recall_record(p["pred_boxes"], recall, i, batch_dict)
        return batch_dict["pred_dicts"], recall, batch_dict
