"""Row N2 — AnchorHeadSingle with the reference's constructor signature, parameter names (`conv_cls`, `conv_box`,
`conv_dir_cls`) and batch_dict contract (pcdet/models/dense_heads/anchor_head_single.py:6-33 ctor, :109-145 eval forward;
anchor_head_template.py:38-47 anchors, :293-340 generate_predicted_boxes).

Eval forward only: the three 1x1 convolutions run as ONE tcgen05 GEMM (hvpr_conv2d, bf16 operands, fp32 accumulate, fp32 NHWC
output so logits / box deltas are not rounded), followed by hvpr_head_decode (ResidualCoder decode + direction fix-up).
Anchors are a constant (H*W*A, 7) tensor built once with the reference's arithmetic (torch.arange in fp32).  No CPU fallback.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .backbone import BaseBEVBackbone_Scale, _ConvLayer


def _cfg_get(cfg, k, default=None):
    return cfg.get(k, default) if hasattr(cfg, "get") else getattr(cfg, k, default)


def build_anchors(anchor_cfgs, grid_size, point_cloud_range, device):
    """(ny, nx, A, 7) fp32 anchors [x, y, z, dx, dy, dz, r] of ONE anchor set, A = sizes x rotations, ordered (size, rotation)
    exactly as anchor_generator.py:17-60 lays them out after its permute; centres from torch.arange in fp32 like the reference."""
    if len(anchor_cfgs) != 1:
        raise NotImplementedError("AnchorHeadSingle (B200): one anchor set / class (hvpr.yaml:103-113)")
    c = anchor_cfgs[0]
    stride = int(_cfg_get(c, "feature_map_stride"))
    gx, gy = int(grid_size[0]) // stride, int(grid_size[1]) // stride
    r = [float(v) for v in point_cloud_range]
    if _cfg_get(c, "align_center", False):
        sx, sy = (r[3] - r[0]) / gx, (r[4] - r[1]) / gy
        ox, oy = sx / 2, sy / 2
    else:
        sx, sy = (r[3] - r[0]) / (gx - 1), (r[4] - r[1]) / (gy - 1)
        ox, oy = 0, 0
    xs = torch.arange(r[0] + ox, r[3] + 1e-5, step=sx, dtype=torch.float32)
    ys = torch.arange(r[1] + oy, r[4] + 1e-5, step=sy, dtype=torch.float32)
    heights = list(_cfg_get(c, "anchor_bottom_heights"))
    if len(heights) != 1:
        raise NotImplementedError("one anchor_bottom_height per class")
    sizes = torch.tensor(_cfg_get(c, "anchor_sizes"), dtype=torch.float32)
    rots = torch.tensor(_cfg_get(c, "anchor_rotations"), dtype=torch.float32)
    A = sizes.shape[0] * rots.shape[0]
    an = torch.empty(len(ys), len(xs), A, 7, dtype=torch.float32)
    an[..., 0] = xs[None, :, None]
    an[..., 1] = ys[:, None, None]
    an[..., 3:6] = sizes.repeat_interleave(rots.shape[0], 0)[None, None]
    an[..., 6] = rots.repeat(sizes.shape[0])[None, None]
    an[..., 2] = torch.tensor(heights[0], dtype=torch.float32) + an[..., 5] / 2          # bottom height -> box centre (:56)
    return an.to(device).contiguous()


class AnchorHeadSingle(nn.Module):
    def __init__(self, model_cfg, input_channels, num_class, class_names, grid_size, point_cloud_range,
                 predict_boxes_when_training=True):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_class = num_class
        self.class_names = class_names
        self.grid_size = [int(v) for v in grid_size]
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.input_channels = input_channels
        acfg = _cfg_get(model_cfg, "ANCHOR_GENERATOR_CONFIG")
        self._anchor_cfgs = list(acfg)
        a0 = self._anchor_cfgs[0]
        self.num_anchors_per_location = len(_cfg_get(a0, "anchor_sizes")) * len(_cfg_get(a0, "anchor_rotations")) * \
            len(_cfg_get(a0, "anchor_bottom_heights"))
        self.code_size = 7                                           # ResidualCoder, box_coder_utils.py:5-8
        A = self.num_anchors_per_location
        self.conv_cls = nn.Conv2d(input_channels, A * num_class, kernel_size=1)
        self.conv_box = nn.Conv2d(input_channels, A * self.code_size, kernel_size=1)
        self.use_dir = _cfg_get(model_cfg, "USE_DIRECTION_CLASSIFIER", None) is not None
        self.num_dir_bins = int(_cfg_get(model_cfg, "NUM_DIR_BINS", 2)) if self.use_dir else 0
        self.conv_dir_cls = nn.Conv2d(input_channels, A * self.num_dir_bins, kernel_size=1) if self.use_dir else None
        pi = 0.01                                                    # anchor_head_single.py:35-38
        nn.init.constant_(self.conv_cls.bias, -np.log((1 - pi) / pi))
        nn.init.normal_(self.conv_box.weight, mean=0, std=0.001)
        if input_channels % 64:
            raise NotImplementedError("AnchorHeadSingle (B200): input_channels must be a multiple of 64")
        self._packed = None
        self._packed_key = None
        self._anchors = None
        self._bufs = {}

    # ------------------------------------------------------------------------------------------ weights / anchors
    def _layout(self):
        A = self.num_anchors_per_location
        cls_off, box_off = 0, A * self.num_class
        dir_off = box_off + A * self.code_size if self.use_dir else -1
        used = box_off + A * self.code_size + (A * self.num_dir_bins if self.use_dir else 0)
        return cls_off, box_off, dir_off, (used + 31) // 32 * 32

    def _ensure_packed(self, dev):
        key = (tuple((p.data_ptr(), p._version) for p in self.parameters()), str(dev))
        if self._packed is not None and self._packed_key == key:
            return self._packed
        _lib.init_device()
        cls_off, box_off, dir_off, n_total = self._layout()
        convs = [self.conv_cls, self.conv_box] + ([self.conv_dir_cls] if self.use_dir else [])
        w = torch.zeros(n_total, 1, self.input_channels, dtype=torch.float32, device=dev)
        b = torch.zeros(n_total, dtype=torch.float32, device=dev)
        row = 0
        for c in convs:
            n = c.weight.shape[0]
            w[row:row + n, 0] = c.weight.detach().float().reshape(n, -1).to(dev)
            b[row:row + n] = c.bias.detach().float().to(dev)
            row += n
        lay = _ConvLayer()
        lay.n_total, lay.bn, lay.c_in, lay.ksize, lay.stride, lay.up, lay.c_out = n_total, 32, self.input_channels, 1, 1, 1, n_total
        L = _lib.lib()
        lay.wpk = torch.empty(L.hvpr_conv_packed_bytes(n_total, 1, self.input_channels), dtype=torch.uint8, device=dev)
        lay.bias = b
        _lib.check(L.hvpr_conv_pack_weights(_lib.ptr(w), n_total, 1, self.input_channels, lay.bn, _lib.ptr(lay.wpk), _lib.cur_stream()),
                   "hvpr_conv_pack_weights")
        self._packed, self._packed_key = lay, key
        return lay

    def anchors(self, dev):
        if self._anchors is None or self._anchors.device != torch.device(dev):
            self._anchors = build_anchors(self._anchor_cfgs, self.grid_size, self.point_cloud_range, dev)
        return self._anchors

    # ------------------------------------------------------------------------------------------ forward
    def run_nhwc(self, x_nhwc, B, H, W):
        """x_nhwc (B,H,W,cs>=C) bf16 -> batch_cls_preds (B, H*W*A, num_class), batch_box_preds (B, H*W*A, 7) fp32."""
        dev = x_nhwc.device
        lay, an, L = self._ensure_packed(dev), self.anchors(dev), _lib.lib()
        if an.shape[0] != H or an.shape[1] != W:
            raise _lib.HvprError("anchor map %dx%d (grid_size // feature_map_stride) does not match the %dx%d feature map "
                                 "(hvpr.yaml ships feature_map_stride 2 with a full-resolution backbone: breakage B10)"
                                 % (an.shape[0], an.shape[1], H, W))
        cls_off, box_off, dir_off, n_total = self._layout()
        A = self.num_anchors_per_location
        key = (B, H, W, str(dev))
        if key not in self._bufs:
            self._bufs[key] = (torch.empty(B, H, W, n_total, dtype=torch.float32, device=dev),
                               torch.empty(B, H * W * A, self.num_class, dtype=torch.float32, device=dev),
                               torch.empty(B, H * W * A, self.code_size, dtype=torch.float32, device=dev))
        head, cls_out, box_out = self._bufs[key]
        BaseBEVBackbone_Scale._conv(lay, x_nhwc, B, H, W, head, relu=False, out_mode=2)
        st = L.hvpr_head_decode(_lib.ptr(head), B, H, W, n_total, A, self.num_class, cls_off, box_off, dir_off, self.num_dir_bins,
                                _lib.ptr(an), float(_cfg_get(self.model_cfg, "DIR_OFFSET", 0.0)),
                                float(_cfg_get(self.model_cfg, "DIR_LIMIT_OFFSET", 0.0)), _lib.ptr(cls_out), _lib.ptr(box_out),
                                _lib.cur_stream())
        _lib.check(st, "hvpr_head_decode")
        return cls_out, box_out

    def forward(self, data_dict):
        if self.training:
            raise NotImplementedError("hvpr_b200.AnchorHeadSingle implements the eval branch (anchor_head_single.py:109-145) only")
        x = data_dict["spatial_features_2d"]
        if not x.is_cuda:
            raise _lib.HvprError("AnchorHeadSingle needs CUDA tensors; there is no CPU path")
        _lib.init_device()
        B, C, H, W = x.shape
        key = ("in", B, H, W, str(x.device))
        if key not in self._bufs:
            self._bufs[key] = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=x.device)
        x_nhwc = self._bufs[key]
        _lib.check(_lib.lib().hvpr_nchw_to_nhwc_bf16(_lib.ptr(x.contiguous().float()), B, C, H, W, _lib.ptr(x_nhwc), C,
                                                     _lib.cur_stream()), "nchw_to_nhwc")
        cls, box = self.run_nhwc(x_nhwc, B, H, W)
        data_dict["batch_cls_preds"] = cls.clone()                     # fresh tensors (run_nhwc() reuses per-shape buffers)
        data_dict["batch_box_preds"] = box.clone()
        data_dict["cls_preds_normalized"] = False                      # anchor_head_single.py:143
        return data_dict


__all__ = {"AnchorHeadSingle": AnchorHeadSingle}
