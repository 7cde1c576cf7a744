"""ORACLE (test infrastructure) — CPU restatement of row N1: BaseBEVBackbone_Scale eval forward
(pcdet/models/backbones_2d/base_bev_backbone.py:280-315, layers built at :150-220) with the SpatialAttention gate
(pcdet/models/backbones_2d/spatial_attention.py:47-63).  fp32 torch functional ops on the CPU, no nn.Module.

Pinned: tests/golden/backbone_tiny.npz holds the output of the REFERENCE'S OWN module (oracle/ref_loader.load_backbone,
one in-memory patch: B4) for the seeded weights / inputs below; tests/test_oracle_cpu.py checks this restatement
against it.  Never imported by hvpr_b200/.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3                      # base_bev_backbone.py:163,170,177 / spatial_attention.py:29

CFG = dict(LAYER_NUMS=[3, 3, 3], SFM_LAYER_NUMS=[3, 3, 3], LAYER_STRIDES=[1, 2, 2], NUM_FILTERS=[128, 256, 512],
           NUM_SCALE_FILTERS=[32, 64, 128], UPSAMPLE_STRIDES=[1, 2, 4], NUM_UPSAMPLE_FILTERS=[128, 128, 128])  # hvpr.yaml:87-95


def _bn(rng, prefix, c, w):
    w[prefix + ".weight"] = rng.uniform(0.6, 1.4, c).astype(np.float32)
    w[prefix + ".bias"] = rng.uniform(-0.2, 0.2, c).astype(np.float32)
    w[prefix + ".running_mean"] = rng.uniform(-0.2, 0.2, c).astype(np.float32)
    w[prefix + ".running_var"] = rng.uniform(0.5, 1.5, c).astype(np.float32)


# the plain backbone as configured for PointPillars (tools/cfgs/kitti_models/pointpillar.yaml BACKBONE_2D)
PLAIN_CFG = dict(LAYER_NUMS=[3, 5, 5], LAYER_STRIDES=[2, 2, 2], NUM_FILTERS=[64, 128, 256], UPSAMPLE_STRIDES=[1, 2, 4],
                 NUM_UPSAMPLE_FILTERS=[128, 128, 128])


def random_backbone_weights(seed: int, cfg=CFG, input_channels: int = 128, with_scale: bool = True):
    """Reference parameter names/shapes (state_dict of BaseBEVBackbone_Scale).  Conv weights ~ U(+-sqrt(3/fan_in)) keep
    the activation scale ~1 through 7 layers; BN affine and running stats are randomised (default BN is almost identity).
    numpy PCG64 streams are stable across versions, so the GPU box regenerates the same tensors from the seed."""
    rng = np.random.default_rng(seed)
    w = {}
    nf = cfg["NUM_FILTERS"]
    nsf = cfg["NUM_SCALE_FILTERS"] if with_scale else [0] * len(nf)
    cin = [input_channels] + nf[:-1]
    cin_s = [input_channels // 4] + nsf[:-1]

    def conv(name, co, ci, kh, kw):
        a = np.sqrt(3.0 / (ci * kh * kw))
        w[name] = rng.uniform(-a, a, (co, ci, kh, kw)).astype(np.float32)

    for i in range(len(nf)):
        conv("blocks.%d.1.weight" % i, nf[i], cin[i], 3, 3)
        _bn(rng, "blocks.%d.2" % i, nf[i], w)
        for k in range(cfg["LAYER_NUMS"][i]):
            conv("blocks.%d.%d.weight" % (i, 4 + 3 * k), nf[i], nf[i], 3, 3)
            _bn(rng, "blocks.%d.%d" % (i, 5 + 3 * k), nf[i], w)
        if with_scale:
            conv("sfmblocks_down.%d.0.weight" % i, nf[i], nf[i], 3, 3)
            _bn(rng, "sfmblocks_down.%d.1" % i, nf[i], w)
        s = cfg["UPSAMPLE_STRIDES"][i]
        a = np.sqrt(3.0 / nf[i])
        w["deblocks.%d.0.weight" % i] = rng.uniform(-a, a, (nf[i], cfg["NUM_UPSAMPLE_FILTERS"][i], s, s)).astype(np.float32)
        _bn(rng, "deblocks.%d.1" % i, cfg["NUM_UPSAMPLE_FILTERS"][i], w)
        if with_scale:
            conv("scale_layers.%d.1.weight" % i, nsf[i], cin_s[i], 3, 3)
            _bn(rng, "scale_layers.%d.2" % i, nsf[i], w)
    if with_scale:
        w["attention.spatial.conv.weight"] = rng.uniform(-0.5, 0.5, (1, 2, 3, 3)).astype(np.float32)
        w["attention.spatial.conv.bias"] = rng.uniform(-0.2, 0.2, 1).astype(np.float32)
        _bn(rng, "attention.spatial.norm", 1, w)
    return w


def random_canvases(seed: int, batch: int, h: int, w: int, occupancy: float = 0.15):
    """BEV-like inputs: mostly-empty canvases with non-negative features at the occupied cells (post-ReLU pillars)."""
    rng = np.random.default_rng(seed)
    occ = rng.random((batch, 1, h, w)) < occupancy
    spatial = (np.abs(rng.standard_normal((batch, 128, h, w))) * occ).astype(np.float32)
    scale = (np.abs(rng.standard_normal((batch, 32, h, w))) * occ).astype(np.float32)
    return spatial, scale


def _t(w, k):
    return torch.from_numpy(np.asarray(w[k]))


def _conv_bn_relu(x, w, conv_key, bn_prefix, stride, relu=True, bias_key=None):
    # ZeroPad2d(1) + padding=0 (:153-158) and padding=1 (:167) are the same arithmetic
    x = F.conv2d(x, _t(w, conv_key), _t(w, bias_key) if bias_key else None, stride=stride, padding=1)
    x = F.batch_norm(x, _t(w, bn_prefix + ".running_mean"), _t(w, bn_prefix + ".running_var"), _t(w, bn_prefix + ".weight"),
                     _t(w, bn_prefix + ".bias"), False, 0.0, BN_EPS)
    return F.relu(x) if relu else x


def attention_gate(y, w):
    """spatial_attention.py:53-61 — gate depends on the scale branch only (x is multiplied afterwards)."""
    pooled = torch.cat((y.max(1, keepdim=True)[0], y.mean(1, keepdim=True)), 1)          # ChannelPool :43-45
    a = _conv_bn_relu(pooled, w, "attention.spatial.conv.weight", "attention.spatial.norm", 1, relu=False,
                      bias_key="attention.spatial.conv.bias")
    return torch.sigmoid(a)


def backbone_forward(w, spatial, scale, cfg=CFG, return_levels: bool = False):
    """base_bev_backbone.py:280-315 (eval branch).  spatial (B,128,H,W), scale (B,32,H,W) fp32 -> (B,384,H,W).
    scale=None: the plain BaseBEVBackbone.forward (:62-102) — no scale branch, no SFM loop."""
    x = torch.from_numpy(np.asarray(spatial)).float()
    y = torch.from_numpy(np.asarray(scale)).float() if scale is not None else None
    ups, levels = [], []
    with torch.no_grad():
        for i in range(len(cfg["NUM_FILTERS"])):
            s = cfg["LAYER_STRIDES"][i]
            x = _conv_bn_relu(x, w, "blocks.%d.1.weight" % i, "blocks.%d.2" % i, s)               # :283, layers :152-165
            for k in range(cfg["LAYER_NUMS"][i]):
                x = _conv_bn_relu(x, w, "blocks.%d.%d.weight" % (i, 4 + 3 * k), "blocks.%d.%d" % (i, 5 + 3 * k), 1)
            xa = x
            if y is not None:
                y = _conv_bn_relu(y, w, "scale_layers.%d.1.weight" % i, "scale_layers.%d.2" % i, s)   # :284, layers :204-213
                gate = attention_gate(y, w)
                for _ in range(cfg["SFM_LAYER_NUMS"][i]):                                            # :286-290
                    xa = gate * _conv_bn_relu(xa, w, "sfmblocks_down.%d.0.weight" % i, "sfmblocks_down.%d.1" % i, 1) + xa
            levels.append(xa)
            us = cfg["UPSAMPLE_STRIDES"][i]
            u = F.conv_transpose2d(xa, _t(w, "deblocks.%d.0.weight" % i), stride=us)             # :293-294, layers :180-188
            p = "deblocks.%d.1" % i
            u = F.relu(F.batch_norm(u, _t(w, p + ".running_mean"), _t(w, p + ".running_var"), _t(w, p + ".weight"),
                                    _t(w, p + ".bias"), False, 0.0, BN_EPS))
            ups.append(u)
        out = torch.cat(ups, 1)                                                                   # :298-299
    if return_levels:
        return out.numpy(), [l.numpy() for l in levels]
    return out.numpy()
