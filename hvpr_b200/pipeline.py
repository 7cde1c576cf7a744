"""Front end + 2-D backbone as one object: raw points -> spatial_features_2d.

Rows (a)-(e) of SURVEY.md §8 (K1 voxelize, K2 PFN, K3 memory attention, K4 canvas fill) followed by row N1
(BaseBEVBackbone_Scale on the tcgen05 convolution kernel).  Between the two the canvases stay channels-last bf16
(K4 writes them in that layout directly), so the fp32 NCHW canvases of the module API (1.1 GB per 8-frame batch at
432 x 496) are never materialised and no transposition pass runs.  The whole chain is captured in one CUDA graph.
"""
from __future__ import annotations

import torch

from . import _lib
from .backbone import BaseBEVBackbone_Scale
from .config import Cfg
from .dense_head import AnchorHeadSingle
from .post_process import PostProcessor
from .frontend import HybridFrontEnd

# tools/cfgs/kitti_models/hvpr.yaml:87-95
HVPR_BACKBONE_CFG = Cfg(NAME="BaseBEVBackbone_Scale", LAYER_NUMS=[3, 3, 3], SFM_LAYER_NUMS=[3, 3, 3], LAYER_STRIDES=[1, 2, 2],
                        NUM_FILTERS=[128, 256, 512], NUM_SCALE_FILTERS=[32, 64, 128], UPSAMPLE_STRIDES=[1, 2, 4],
                        NUM_UPSAMPLE_FILTERS=[128, 128, 128])


# tools/cfgs/kitti_models/hvpr.yaml:96-118, with feature_map_stride 1: the backbone emits full-resolution features and the shipped
# stride 2 cannot be reshaped onto them (breakage B10)
HVPR_HEAD_CFG = Cfg(NAME="AnchorHeadSingle", CLASS_AGNOSTIC=False, USE_DIRECTION_CLASSIFIER=True, DIR_OFFSET=0.78539,
                    DIR_LIMIT_OFFSET=0.0, NUM_DIR_BINS=2,
                    ANCHOR_GENERATOR_CONFIG=[dict(class_name="Car", anchor_sizes=[[3.9, 1.6, 1.56]], anchor_rotations=[0, 1.57],
                                                  anchor_bottom_heights=[-1.78], align_center=False, feature_map_stride=1,
                                                  matched_threshold=0.6, unmatched_threshold=0.45)])


# tools/cfgs/kitti_models/hvpr.yaml:136-148
HVPR_POST_CFG = Cfg(RECALL_THRESH_LIST=[0.3, 0.5, 0.7], SCORE_THRESH=0.1, OUTPUT_RAW_SCORE=False, EVAL_METRIC="kitti",
                    NMS_CONFIG=Cfg(MULTI_CLASSES_NMS=False, NMS_TYPE="nms_gpu", NMS_THRESH=0.1, NMS_PRE_MAXSIZE=4096, NMS_POST_MAXSIZE=500))


class FrontEndWithBackbone(torch.nn.Module):
    def __init__(self, geom, backbone_cfg=HVPR_BACKBONE_CFG, device="cuda", head_cfg=None, post_cfg=None, **frontend_kwargs):
        super().__init__()
        self.frontend = HybridFrontEnd(geom, device=device, **frontend_kwargs) if frontend_kwargs else HybridFrontEnd(geom, device=device)
        self.backbone_2d = BaseBEVBackbone_Scale(backbone_cfg, self.frontend.map_to_bev_module.num_bev_features).to(device).eval()
        # optional row N2: the dense head consumes the backbone's features channels-last, the fp32 NCHW tensor is never written
        self.dense_head = None
        if head_cfg is not None:
            self.dense_head = AnchorHeadSingle(head_cfg, self.backbone_2d.num_bev_features, 1, ["Car"], geom.grid_size,
                                               geom.point_cloud_range).to(device).eval()
        # optional row N3: score threshold / top-k / rotated NMS on the head's output (needs head_cfg)
        self.post = PostProcessor(post_cfg) if (post_cfg is not None and head_cfg is not None) else None
        self._p = None

    def weights_version(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def plan(self, n_frames: int, n_total_points: int, max_frame_points: int = 0):
        """Buffers (and, at the first run, one CUDA graph) for up to `n_total_points` points in `n_frames` frames.  The kernels take
        the live extents from the device-side frame offsets, so any batch that fits the capacity replays the same graph."""
        fe = self.frontend
        p = fe.plan(n_frames, n_total_points, max_frame_points, use_graph=False)
        nx, ny, _ = fe.geom.grid_size
        # the fp32 NCHW canvases of the module API are not needed on this path
        p.spatial = p.spatial_scale = None
        bf = dict(dtype=torch.bfloat16, device=p.points.device)
        p.x_nhwc = torch.zeros((n_frames, ny, nx, 128), **bf)
        p.y_nhwc = torch.zeros((n_frames, ny, nx, 64), **bf)
        if self.dense_head is not None:
            p.f2d_nhwc = torch.empty((n_frames, ny, nx, self.backbone_2d.num_bev_features), **bf)
        p.graph = None
        self._p = p
        return p

    def _enqueue(self, p):
        fe = self.frontend
        nx, ny, _ = fe.geom.grid_size
        vox = fe.voxelizer.run(p.points, p.frame_offsets, p.B, p.max_frame_points, out=p.vox)
        nP = vox.n_pillars_dev
        fe.vfe.run(vox.voxels, vox.num_points, vox.coords, nP, out=p.pillar_features, scale_out=p.pillar_scale)
        fe.map_to_bev_module.run_nhwc(p.pillar_features, p.pillar_scale, vox.cell_map, p.B, nP, p.readout, p.x_nhwc, p.y_nhwc)
        if self.dense_head is None:
            p.out = self.backbone_2d.run_nhwc(p.x_nhwc, p.y_nhwc, p.B, ny, nx)
        else:
            self.backbone_2d.run_nhwc(p.x_nhwc, p.y_nhwc, p.B, ny, nx, out_nhwc=p.f2d_nhwc)
            p.cls_preds, p.box_preds = self.dense_head.run_nhwc(p.f2d_nhwc, p.B, ny, nx)
            if self.post is not None:
                p.det = self.post.run(p.cls_preds, p.box_preds, cls_normalized=False)

    def kernel_launches_per_run(self) -> int:
        bb = self.backbone_2d
        convs = sum(1 + n for n in bb.layer_nums) + sum(bb.sfm_layer_nums) + 2 * len(bb.num_filters)     # blocks + sfm + scale + deblock
        head = (2 if self.dense_head is not None else 0) + (4 if self.post is not None else 0)            # head GEMM + decode, 4 NMS kernels
        return 5 + 1 + 1 + 2 + convs + 2 * len(bb.num_filters) + head                                     # K1 x5, K2, K3, K4 x2, convs, gate x2/level

    @torch.no_grad()
    def run(self):
        """points / frame_offsets already resident in plan.points -> plan.out = spatial_features_2d (B,384,ny,nx) fp32."""
        p = self._p
        _lib.init_device()
        wkey = self.weights_version()
        if p.graph is not None and p.graph_wkey != wkey:
            p.graph = None                                # load_state_dict() after the first run: fold / pack / capture again
        if p.graph is None:
            p.graph_wkey = wkey
            fe = self.frontend
            fe.vfe._weights_packed(p.points.device)
            if fe.map_to_bev_module.memory.precision == "bf16_rescore":
                fe.map_to_bev_module.memory._packed_bf16()
            self.backbone_2d._ensure_packed(p.points.device)
            if self.dense_head is not None:
                self.dense_head._ensure_packed(p.points.device)
                self.dense_head.anchors(p.points.device)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._enqueue(p)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                self._enqueue(p)
            p.graph = gr
        p.graph.replay()
        return p

    @torch.no_grad()
    def forward(self, batch_dict: dict) -> dict:
        """batch_dict['points'] (sum N, 5) [b,x,y,z,r] on the GPU + 'batch_size'  ->  adds 'spatial_features_2d'."""
        pts = batch_dict["points"]
        if not pts.is_cuda:
            raise _lib.HvprError("FrontEndWithBackbone needs CUDA tensors; there is no CPU path")
        B = int(batch_dict["batch_size"])
        n = pts.shape[0]
        if self._p is None or self._p.B != B or self._p.n_total < n:
            # plan for a CAPACITY (25 % headroom, 64 Ki granularity): real collated batches differ in point count every time, and
            # re-planning reallocates > 1.5 GB at G2 / B = 8 and re-captures the graph
            self.plan(B, max(65536, -(-(n + n // 4) // 65536) * 65536))
        p = self._p
        p.points[:n].copy_(pts[:, 1:5])
        st = _lib.lib().hvpr_frame_offsets(_lib.ptr(pts.contiguous()), n, 5, B, _lib.ptr(p.frame_offsets), _lib.cur_stream())
        _lib.check(st, "hvpr_frame_offsets")
        self.run()
        if self.dense_head is None:
            batch_dict["spatial_features_2d"] = p.out.clone()       # the reference returns fresh tensors; plan.out is overwritten by the next call
        else:
            batch_dict["batch_cls_preds"], batch_dict["batch_box_preds"] = p.cls_preds.clone(), p.box_preds.clone()
            batch_dict["cls_preds_normalized"] = False
            if self.post is not None:                    # pred_dicts as Detector3DTemplate.post_processing returns them (one D2H of the counts)
                counts = p.det["count"].cpu().tolist()
                batch_dict["pred_dicts"] = [{"pred_boxes": p.det["boxes"][i, :k].clone(), "pred_scores": p.det["scores"][i, :k].clone(),
                                             "pred_labels": p.det["labels"][i, :k].long()} for i, k in enumerate(counts)]
        return batch_dict
