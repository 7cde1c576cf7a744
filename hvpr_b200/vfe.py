"""Drop-in boundary #2a — VFE modules with the reference's constructor signature, parameter names and batch_dict
contract (pcdet/models/backbones_3d/vfe/pillar_vfe.py:52-221, vfe_template.py:4-22, vfe/__init__.py:5-10).

forward() runs hvpr_b200/csrc/pfn.cu through the C ABI (eval mode only — this build accelerates the inference path;
training raises).  A reference `model_state` loads unchanged: `pfn_layers.N.linear.weight`, `pfn_layers.N.norm.*`,
`pfn_scale_layers.N.0.weight`, `pfn_scale_layers.N.1.*`.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib


class VFETemplate(nn.Module):
    def __init__(self, model_cfg, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg

    def get_output_feature_dim(self):
        raise NotImplementedError

    def forward(self, **kwargs):
        raise NotImplementedError


class PFNLayer(nn.Module):
    """Parameter container with the reference's layout (pillar_vfe.py:8-27); the math is fused into pfn.cu."""

    def __init__(self, in_channels, out_channels, use_norm=True, last_layer=False):
        super().__init__()
        self.last_vfe = last_layer
        self.use_norm = use_norm
        if not self.last_vfe:
            out_channels = out_channels // 2
        if self.use_norm:
            self.linear = nn.Linear(in_channels, out_channels, bias=False)
            self.norm = nn.BatchNorm1d(out_channels, eps=1e-3, momentum=0.01)
        else:
            self.linear = nn.Linear(in_channels, out_channels, bias=True)
        self.part = 50000

    def folded(self):
        """(W', b') with eval-mode BN merged: y = W'x + b' (float64 on the host, cast to fp32)."""
        w = self.linear.weight.detach().double().cpu()
        if self.use_norm:
            s = self.norm.weight.detach().double().cpu() / torch.sqrt(self.norm.running_var.detach().double().cpu() + self.norm.eps)
            b = self.norm.bias.detach().double().cpu() - self.norm.running_mean.detach().double().cpu() * s
            return (w * s[:, None]).float(), b.float()
        return w.float(), self.linear.bias.detach().double().cpu().float()


def _fold_seq(seq):
    lin, bn = seq[0], seq[1]
    w = lin.weight.detach().double().cpu()
    s = bn.weight.detach().double().cpu() / torch.sqrt(bn.running_var.detach().double().cpu() + bn.eps)
    b = bn.bias.detach().double().cpu() - bn.running_mean.detach().double().cpu() * s
    return (w * s[:, None]).float(), b.float()


class _PillarVFEBase(VFETemplate):
    _HAS_SCALE = False

    def __init__(self, model_cfg, num_point_features, voxel_size, point_cloud_range):
        super().__init__(model_cfg=model_cfg)
        self.use_norm = self.model_cfg.USE_NORM
        self.with_distance = self.model_cfg.WITH_DISTANCE
        self.use_absolute_xyz = self.model_cfg.USE_ABSLOTE_XYZ
        self._raw_point_features = num_point_features
        num_point_features += 6 if self.use_absolute_xyz else 3
        if self.with_distance:
            num_point_features += 1
        self.num_filters = self.model_cfg.NUM_FILTERS
        assert len(self.num_filters) > 0
        num_filters = [num_point_features] + list(self.num_filters)
        pfn_layers = []
        for i in range(len(num_filters) - 1):
            pfn_layers.append(PFNLayer(num_filters[i], num_filters[i + 1], self.use_norm,
                                       last_layer=(i >= len(num_filters) - 2)))
        self.pfn_layers = nn.ModuleList(pfn_layers)

        if self._HAS_SCALE:
            self.num_scale_features = self.model_cfg.NUM_SCALE_FEATURES
            assert len(self.num_scale_features) > 0
            nsf = [5] + list(self.num_scale_features)
            self.pfn_scale_layers = nn.ModuleList()
            for i in range(len(nsf) - 1):
                self.pfn_scale_layers.append(nn.Sequential(
                    nn.Linear(nsf[i], nsf[i + 1], bias=False),
                    nn.BatchNorm1d(nsf[i + 1], eps=1e-3, momentum=0.01),
                    nn.ReLU()))

        # same Python expressions as pillar_vfe.py:166-171 (E6)
        self.voxel_x = voxel_size[0]
        self.voxel_y = voxel_size[1]
        self.voxel_z = voxel_size[2]
        self.x_offset = self.voxel_x / 2 + point_cloud_range[0]
        self.y_offset = self.voxel_y / 2 + point_cloud_range[1]
        self.z_offset = self.voxel_z / 2 + point_cloud_range[2]
        self._geom_c = _lib.make_geom(point_cloud_range[:3], voxel_size, (0, 0, 0))
        self._wcache = None
        self._wkey = None
        self.emit_mask = True        # pillar_vfe.py:220 writes batch_dict['pillar_mask']; nothing downstream reads it at eval
        self._check_supported()

    def _check_supported(self):
        ok = (self._raw_point_features == 4 and self.use_absolute_xyz and not self.with_distance and self.use_norm
              and list(self.num_filters) == [32, 64]
              and (not self._HAS_SCALE or list(self.num_scale_features) == [16, 32]))
        if not ok:
            raise NotImplementedError(
                "hvpr_b200 PFN kernel is built for the shipped HVPR cfg (tools/cfgs/kitti_models/hvpr.yaml:69-75): "
                "4 point features, USE_ABSLOTE_XYZ, no distance, USE_NORM, NUM_FILTERS [32,64], NUM_SCALE_FEATURES [16,32]")

    def get_output_feature_dim(self):
        return self.num_filters[-1]

    # -- folded weights, cached against parameter versions (no device sync on the steady-state path) --------------
    def _weights(self):
        key = tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))
        if self._wcache is None or key != self._wkey:
            W = _lib.HvprPfnWeights()
            w0, b0 = self.pfn_layers[0].folded()
            w1, b1 = self.pfn_layers[1].folded()

            def put(field, t):
                flat = t.contiguous().view(-1).tolist()
                arr = getattr(W, field)
                assert len(flat) == len(arr), (field, len(flat), len(arr))
                for i, v in enumerate(flat):
                    arr[i] = v
            put("w0", w0); put("b0", b0)
            put("w1a", w1[:, :16]); put("w1b", w1[:, 16:]); put("b1", b1)
            if self._HAS_SCALE:
                ws0, bs0 = _fold_seq(self.pfn_scale_layers[0])
                ws1, bs1 = _fold_seq(self.pfn_scale_layers[1])
                put("ws0", ws0); put("bs0", bs0); put("ws1", ws1); put("bs1", bs1)
            self._wcache, self._wkey = W, key
            self._wpacked = None
        return self._wcache

    def _weights_packed(self, device):
        """Device image of the tensor-core weight fragments (hvpr_pfn_pack), rebuilt when the folded weights change."""
        W = self._weights()
        if getattr(self, "_wpacked", None) is None or self._wpacked.device != device:
            buf = torch.empty(int(_lib.lib().hvpr_pfn_packed_bytes()), dtype=torch.uint8, device=device)
            _lib.check(_lib.lib().hvpr_pfn_pack(ctypes.byref(W), _lib.ptr(buf), _lib.cur_stream()), "hvpr_pfn_pack")
            self._wpacked = buf
        return self._wpacked

    def invalidate_weights(self):
        """Drop the folded-weight cache (needed only after raw `.data` edits, which bypass tensor version counters)."""
        self._wcache = None

    def run(self, voxels, num_points, coords, n_pillars_dev=None, out=None, scale_out=None, mask_out=None, launch=None):
        """Enqueue the fused PFN on the current stream.  int32 coords/counts; returns (features, scale, mask)."""
        _lib.init_device()
        rows, T = voxels.shape[0], voxels.shape[1]
        dev = voxels.device
        if out is None:
            out = torch.empty((rows, 64), dtype=torch.float32, device=dev)
        if self._HAS_SCALE and scale_out is None:
            scale_out = torch.empty((rows, 32), dtype=torch.float32, device=dev)
        st = _lib.lib().hvpr_pfn(
            _lib.ptr(voxels), _lib.ptr(num_points), _lib.ptr(coords), _lib.ptr(n_pillars_dev), rows, T,
            ctypes.byref(self._weights()), ctypes.byref(self._geom_c),
            float(self.x_offset), float(self.y_offset), float(self.z_offset),
            _lib.ptr(out), _lib.ptr(scale_out) if self._HAS_SCALE else None, _lib.ptr(mask_out), _lib.ptr(self._weights_packed(dev)),
            _lib.launch_cfg(launch),
            _lib.cur_stream())
        _lib.check(st, "hvpr_pfn")
        return out, scale_out, mask_out

    def forward(self, batch_dict, **kwargs):
        if self.training:
            raise NotImplementedError("hvpr_b200 VFE modules implement the inference (eval) path only")
        voxels, num_points, coords = batch_dict["voxels"], batch_dict["voxel_num_points"], batch_dict["voxel_coords"]
        if not voxels.is_cuda:
            raise _lib.HvprError("hvpr_b200 has no CPU path: batch_dict tensors must be on a CUDA device")
        voxels = voxels.contiguous().float()
        # upstream load_data_to_gpu floats everything (E4): accept int or fp32 coords / counts
        num_points_i = num_points if num_points.dtype == torch.int32 else num_points.to(torch.int32)
        coords_i = coords if coords.dtype == torch.int32 else coords.to(torch.int32)
        coords_i, num_points_i = coords_i.contiguous(), num_points_i.contiguous()
        rows, T = voxels.shape[0], voxels.shape[1]
        mask = torch.empty((rows, T, 1), dtype=torch.float32, device=voxels.device) if self.emit_mask else None
        feats, scale, _ = self.run(voxels, num_points_i, coords_i, batch_dict.get("num_pillars_dev"), mask_out=mask)
        batch_dict["pillar_features"] = feats
        if self._HAS_SCALE:
            batch_dict["pillar_scale_features"] = scale
        if mask is not None:
            batch_dict["pillar_mask"] = mask
        return batch_dict


class PillarVFE(_PillarVFEBase):
    """pillar_vfe.py:52-124"""
    _HAS_SCALE = False


class PillarVFE_Scale(_PillarVFEBase):
    """pillar_vfe.py:127-221"""
    _HAS_SCALE = True


__all__ = {
    "VFETemplate": VFETemplate,
    "PillarVFE": PillarVFE,
    "PillarVFE_Scale": PillarVFE_Scale,
}
