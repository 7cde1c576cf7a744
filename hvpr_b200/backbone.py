"""Row N1 — BaseBEVBackbone_Scale with the reference's constructor signature, parameter names and
batch_dict contract (pcdet/models/backbones_2d/base_bev_backbone.py:116-222 ctor, :280-315 eval forward; registry
backbones_2d/__init__.py:3-6), plus SpatialAttention's parameters (spatial_attention.py:47-52).

forward() (eval only) runs every convolution as the tcgen05 implicit-GEMM kernel of hvpr_b200/csrc/conv_tc.cu through the
C ABI: NHWC bf16 activations, eval-mode BN folded into bf16 weights + fp32 bias, fp32 accumulation in TMEM, the attention
gate and the residual of the SFM loop fused into the epilogue, the three transposed convolutions written straight into
their channel slices of the fp32 NCHW `spatial_features_2d`.  A reference `model_state` loads unchanged
(`blocks.L.{1,4,7,10}.weight`, `blocks.L.{2,5,8,11}.*`, `sfmblocks_down.L.{0,1}.*`, `scale_layers.L.{1,2}.*`,
`deblocks.L.{0,1}.*`, `attention.spatial.{conv,norm}.*`).  There is no CPU / eager fallback.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib

BN_EPS, BN_MOM = 1e-3, 0.01


def _slots(*mods):
    """nn.Sequential used purely as a parameter container: indices match the reference's layer lists."""
    return nn.Sequential(*mods)


def _bn(c):
    return nn.BatchNorm2d(c, eps=BN_EPS, momentum=BN_MOM)


class _SpatialGateParams(nn.Module):
    """attention.spatial.conv / attention.spatial.norm (spatial_attention.py:10-33,51)."""

    def __init__(self):
        super().__init__()
        self.spatial = nn.Module()
        self.spatial.conv = nn.Conv2d(2, 1, kernel_size=3, stride=1, padding=1)
        self.spatial.norm = nn.BatchNorm2d(1, eps=BN_EPS, momentum=BN_MOM)


def _fold(conv_w, bn, conv_b=None):
    """eval-mode BN merged into the preceding conv, in float64: y = (W*s) x + (beta + (b - mean)*s)."""
    s = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    b0 = conv_b.detach().double() if conv_b is not None else 0.0
    return conv_w.detach().double(), s, bn.bias.detach().double() + (b0 - bn.running_mean.detach().double()) * s


def _bn_tile(n_total: int) -> int:
    return 256 if n_total % 256 == 0 else (128 if n_total % 128 == 0 else (64 if n_total % 64 == 0 else 32))


class _ConvLayer:
    """One packed GEMM: weights image, bias, geometry."""

    __slots__ = ("wpk", "bias", "n_total", "bn", "c_in", "ksize", "stride", "up", "c_out")


class BaseBEVBackbone_Scale(nn.Module):
    """Eval forward of base_bev_backbone.py:280-315; the plain BaseBEVBackbone (:62-102, no scale branch / SFM loop) is the
    subclass below with _WITH_SCALE = False."""
    _WITH_SCALE = True

    def __init__(self, model_cfg, input_channels):
        super().__init__()
        self.model_cfg = model_cfg
        g = (lambda k: model_cfg.get(k, None)) if hasattr(model_cfg, "get") else (lambda k: getattr(model_cfg, k, None))
        layer_nums, strides, filters = list(g("LAYER_NUMS") or []), list(g("LAYER_STRIDES") or []), list(g("NUM_FILTERS") or [])
        assert len(layer_nums) == len(strides) == len(filters)
        self.sfm_layer_nums = list(g("SFM_LAYER_NUMS") or [])
        up_strides, up_filters = list(g("UPSAMPLE_STRIDES") or []), list(g("NUM_UPSAMPLE_FILTERS") or [])
        scale_filters = list(g("NUM_SCALE_FILTERS") or [])
        if not self._WITH_SCALE:
            scale_filters, self.sfm_layer_nums = [0] * len(filters), [0] * len(filters)
        if not (len(up_strides) == len(up_filters) == len(scale_filters) == len(self.sfm_layer_nums) == len(filters)):
            raise NotImplementedError("BaseBEVBackbone[_Scale] (B200): one deblock (and scale layer / SFM count) per level "
                                      "(hvpr.yaml:87-95 layout) is what the kernels are wired for")
        if any(s not in (1, 2) for s in strides) or any((not float(u).is_integer()) or u < 1 for u in up_strides):
            raise NotImplementedError("LAYER_STRIDES in {1,2} and integer UPSAMPLE_STRIDES >= 1 only")
        self.layer_nums, self.layer_strides, self.num_filters = layer_nums, strides, filters
        self.upsample_strides = [int(u) for u in up_strides]
        self.num_upsample_filters, self.num_scale_filters = up_filters, scale_filters
        self.input_channels = input_channels
        c_in = [input_channels] + filters[:-1]
        c_in_s = [input_channels // 4] + scale_filters[:-1]

        self.sfmblocks_down, self.sfmblocks_up = nn.ModuleList(), nn.ModuleList()
        self.scale_layers, self.blocks, self.deblocks = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for i, nf in enumerate(filters):
            mods = [nn.ZeroPad2d(1), nn.Conv2d(c_in[i], nf, 3, stride=strides[i], padding=0, bias=False), _bn(nf), nn.ReLU()]
            for _ in range(layer_nums[i]):
                mods += [nn.Conv2d(nf, nf, 3, padding=1, bias=False), _bn(nf), nn.ReLU()]
            self.blocks.append(_slots(*mods))
            u = self.upsample_strides[i]
            self.deblocks.append(_slots(nn.ConvTranspose2d(nf, up_filters[i], u, stride=u, bias=False), _bn(up_filters[i]), nn.ReLU()))
            if self._WITH_SCALE:
                self.sfmblocks_down.append(_slots(nn.Conv2d(nf, nf, 3, padding=1, bias=False), _bn(nf), nn.ReLU()))
                self.scale_layers.append(_slots(nn.ZeroPad2d(1), nn.Conv2d(c_in_s[i], scale_filters[i], 3, stride=strides[i],
                                                                           padding=0, bias=False), _bn(scale_filters[i]), nn.ReLU()))
        self.num_bev_features = sum(up_filters)
        if self._WITH_SCALE:
            self.attention = _SpatialGateParams()
        else:
            del self.sfmblocks_down, self.sfmblocks_up, self.scale_layers      # the plain backbone has no such sub-modules
        self._packed = None
        self._packed_key = None
        self._plans = {}

    # ------------------------------------------------------------------------------------------ weights
    def invalidate_weights(self):
        self._packed = None

    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))

    def _pack_conv(self, w, scale, shift, dev) -> _ConvLayer:
        """w (Cout, Cin, kh, kw) float64 -> (Cout, taps, Cin_pad) fp32 -> packed bf16 image."""
        co, ci, kh, kw = w.shape
        ci_pad = (ci + 63) // 64 * 64
        wn = (w * scale[:, None, None, None]).permute(0, 2, 3, 1).reshape(co, kh * kw, ci)
        full = torch.zeros(co, kh * kw, ci_pad, dtype=torch.float32, device=dev)
        full[:, :, :ci] = wn.float()
        return self._pack(full, shift.float().to(dev), co, kh * kw, ci_pad, kh, dev)

    def _pack(self, w_ntc, bias, n_total, taps, ci_pad, ksize, dev) -> _ConvLayer:
        L = _lib.lib()
        lay = _ConvLayer()
        lay.n_total, lay.bn, lay.c_in, lay.ksize = n_total, _bn_tile(n_total), ci_pad, ksize
        lay.wpk = torch.empty(L.hvpr_conv_packed_bytes(n_total, taps, ci_pad), dtype=torch.uint8, device=dev)
        lay.bias = bias.contiguous()
        w_ntc = w_ntc.contiguous()
        _lib.check(L.hvpr_conv_pack_weights(_lib.ptr(w_ntc), n_total, taps, ci_pad, lay.bn, _lib.ptr(lay.wpk),
                                            _lib.cur_stream()), "hvpr_conv_pack_weights")
        lay.stride, lay.up, lay.c_out = 1, 1, n_total
        return lay

    def _ensure_packed(self, dev):
        key = (self._weights_key(), str(dev))
        if self._packed is not None and self._packed_key == key:
            return self._packed
        _lib.init_device()
        P = {"blocks": [], "sfm": [], "scale": [], "de": [], "de_nhwc": []}
        for i in range(len(self.num_filters)):
            seq = self.blocks[i]
            convs = []
            for j in range(1, len(seq), 3):
                lay = self._pack_conv(*_fold(seq[j].weight, seq[j + 1]), dev)
                lay.stride = seq[j].stride[0]
                convs.append(lay)
            P["blocks"].append(convs)
            if self._WITH_SCALE:
                P["sfm"].append(self._pack_conv(*_fold(self.sfmblocks_down[i][0].weight, self.sfmblocks_down[i][1]), dev))
                sl = self._pack_conv(*_fold(self.scale_layers[i][1].weight, self.scale_layers[i][2]), dev)
                sl.stride = self.scale_layers[i][1].stride[0]
                P["scale"].append(sl)
            # ConvTranspose2d(k = s, stride = s): GEMM column (dy*Cout + co)*s + dx  <-  W[ci, co, dy, dx]
            w, s, shift = _fold(self.deblocks[i][0].weight, self.deblocks[i][1])
            ci, co, u, _ = w.shape
            wn = (w * s[None, :, None, None]).permute(2, 1, 3, 0).reshape(u * u * co, 1, ci).float().to(dev)
            de = self._pack(wn, shift.float().repeat_interleave(u).repeat(u).to(dev), u * u * co, 1, ci, 1, dev)
            de.up, de.c_out = u, co
            P["de"].append(de)
            # channels-last variant (out_mode 3, for a native consumer of bf16 NHWC features): column (dy*s + dx)*Cout + co
            wn = (w * s[None, :, None, None]).permute(2, 3, 1, 0).reshape(u * u * co, 1, ci).float().to(dev)
            de2 = self._pack(wn, shift.float().repeat(u * u).to(dev), u * u * co, 1, ci, 1, dev)
            de2.up, de2.c_out = u, co
            P["de_nhwc"].append(de2)
        if self._WITH_SCALE:
            a = self.attention.spatial
            w, s, shift = _fold(a.conv.weight, a.norm, a.conv.bias)
            P["gate_w"] = (ctypes.c_float * 18)(*[float(v) for v in (w * s[:, None, None, None]).reshape(-1).cpu()])
            P["gate_b"] = float(shift.reshape(-1)[0].cpu())
        self._packed, self._packed_key = P, key
        return P

    # ------------------------------------------------------------------------------------------ buffers
    def _plan(self, B, H, W, dev):
        key = (B, H, W, str(dev))
        if key in self._plans:
            return self._plans[key]
        tot = 1
        for s in self.layer_strides:
            tot *= s
        if H % tot or W % tot:
            raise NotImplementedError("canvas %dx%d is not divisible by the total stride %d" % (H, W, tot))
        bf = dict(dtype=torch.bfloat16, device=dev)
        pl = {"x_in": torch.zeros(B, H, W, (self.input_channels + 63) // 64 * 64, **bf), "lv": []}
        if self._WITH_SCALE:
            pl["y_in"] = torch.zeros(B, H, W, (self.input_channels // 4 + 63) // 64 * 64, **bf)
        h, w = H, W
        for i, nf in enumerate(self.num_filters):
            h, w = h // self.layer_strides[i], w // self.layer_strides[i]
            lv = {"h": h, "w": w, "a": torch.empty(B, h, w, nf, **bf), "b": torch.empty(B, h, w, nf, **bf)}
            if self._WITH_SCALE:
                ycs = (self.num_scale_filters[i] + 63) // 64 * 64
                lv.update({"c": torch.empty(B, h, w, nf, **bf),
                           "y": torch.zeros(B, h, w, ycs, **bf),               # pad channels stay zero
                           "pooled": torch.empty(B, h, w, 2, dtype=torch.float32, device=dev),
                           "gate": torch.empty(B, h, w, dtype=torch.float32, device=dev)})
            pl["lv"].append(lv)
        # every deblock must land on one common resolution (they are concatenated, :298 / :96)
        ho, wo = pl["lv"][0]["h"] * self.upsample_strides[0], pl["lv"][0]["w"] * self.upsample_strides[0]
        if any(lv["h"] * u != ho or lv["w"] * u != wo for lv, u in zip(pl["lv"], self.upsample_strides)):
            raise NotImplementedError("deblock outputs must share one resolution (UPSAMPLE_STRIDES vs LAYER_STRIDES)")
        pl["out"] = torch.empty(B, self.num_bev_features, ho, wo, dtype=torch.float32, device=dev)
        self._plans[key] = pl
        return pl

    @staticmethod
    def _conv(lay, src, n, h_in, w_in, dst, *, relu=True, gate=None, residual=None, out_mode=0, out_c_off=0, out_ctot=0):
        a = _lib.HvprConvArgs()
        a.in_, a.n, a.h_in, a.w_in, a.in_cs, a.c_in = src.data_ptr(), n, h_in, w_in, src.shape[-1], lay.c_in
        a.ksize, a.stride, a.w_packed, a.n_total, a.bn = lay.ksize, lay.stride, lay.wpk.data_ptr(), lay.n_total, lay.bn
        a.bias, a.relu = lay.bias.data_ptr(), int(relu)
        a.gate = gate.data_ptr() if gate is not None else None
        a.residual = residual.data_ptr() if residual is not None else None
        a.res_cs = residual.shape[-1] if residual is not None else 0
        a.out_mode, a.out = out_mode, dst.data_ptr()
        a.out_cs = dst.shape[-1] if out_mode in (0, 2, 3) else 0
        a.out_c_off, a.up, a.c_out, a.out_ctot = out_c_off, lay.up, lay.c_out, out_ctot
        _lib.check(_lib.lib().hvpr_conv2d(ctypes.byref(a), _lib.cur_stream()), "hvpr_conv2d")

    # ------------------------------------------------------------------------------------------ forward
    def run_nhwc(self, x_in, y_in, B, H, W, out_nhwc=None):
        """x_in (B,H,W,>=C) / y_in (B,H,W,>=C/4) NHWC bf16 (zero pad channels) -> spatial_features_2d (B,384,H,W) fp32, or — when
        `out_nhwc` (B,Ho,Wo,>=384) bf16 is given — the same features channels-last for a native consumer (the dense head)."""
        dev = x_in.device
        P, pl, L = self._ensure_packed(dev), self._plan(B, H, W, dev), _lib.lib()
        st = _lib.cur_stream()
        x, y, h, w = x_in, y_in, H, W
        c_off = 0
        for i in range(len(self.num_filters)):
            lv = pl["lv"][i]
            cur, other = lv["a"], lv["b"]
            for j, lay in enumerate(P["blocks"][i]):                                   # :283
                self._conv(lay, x, B, h if j == 0 else lv["h"], w if j == 0 else lv["w"], cur)
                x, cur, other = cur, other, cur
            xa = x
            if self._WITH_SCALE:
                self._conv(P["scale"][i], y, B, h, w, lv["y"])                          # :284
                y = lv["y"]
                _lib.check(L.hvpr_attention_gate(_lib.ptr(y), B, lv["h"], lv["w"], y.shape[-1], self.num_scale_filters[i],
                                                 P["gate_w"], P["gate_b"], _lib.ptr(lv["pooled"]), _lib.ptr(lv["gate"]), st),
                           "hvpr_attention_gate")
                # the SFM chain works on a copy-free side branch: x (the blocks' output) also feeds the next level (:283)
                ring = (cur, lv["c"])
                for k in range(self.sfm_layer_nums[i]):                                # :286-290
                    self._conv(P["sfm"][i], xa, B, lv["h"], lv["w"], ring[k & 1], gate=lv["gate"], residual=xa)
                    xa = ring[k & 1]
            h, w = lv["h"], lv["w"]
            if out_nhwc is None:
                self._conv(P["de"][i], xa, B, h, w, pl["out"], out_mode=1, out_c_off=c_off, out_ctot=self.num_bev_features)  # :293-299
            else:
                self._conv(P["de_nhwc"][i], xa, B, h, w, out_nhwc, out_mode=3, out_c_off=c_off)
            c_off += self.num_upsample_filters[i]
        return pl["out"] if out_nhwc is None else out_nhwc

    def forward(self, data_dict):
        if self.training:
            raise NotImplementedError("hvpr_b200 backbones accelerate the eval forward (base_bev_backbone.py:62-102, :280-315); "
                                      "training is out of scope")
        sp = data_dict["spatial_features"]
        if not sp.is_cuda:
            raise _lib.HvprError("%s needs CUDA tensors; there is no CPU path" % type(self).__name__)
        _lib.init_device()
        B, C, H, W = sp.shape
        pl, L, st = self._plan(B, H, W, sp.device), _lib.lib(), _lib.cur_stream()
        sp = sp.contiguous().float()
        _lib.check(L.hvpr_nchw_to_nhwc_bf16(_lib.ptr(sp), B, C, H, W, _lib.ptr(pl["x_in"]), pl["x_in"].shape[-1], st), "nchw_to_nhwc")
        if self._WITH_SCALE:
            sc = data_dict["spatial_scale_features"].contiguous().float()
            _lib.check(L.hvpr_nchw_to_nhwc_bf16(_lib.ptr(sc), B, sc.shape[1], H, W, _lib.ptr(pl["y_in"]), pl["y_in"].shape[-1], st), "nchw_to_nhwc")
        # a fresh tensor, like the reference module: run_nhwc() writes a persistent per-shape buffer that the next call overwrites
        data_dict["spatial_features_2d"] = self.run_nhwc(pl["x_in"], pl.get("y_in"), B, H, W).clone()
        return data_dict


class BaseBEVBackbone(BaseBEVBackbone_Scale):
    """The plain 2-D backbone (base_bev_backbone.py:6-102): blocks -> deblocks -> concat, same kernels, no scale branch."""
    _WITH_SCALE = False


__all__ = {"BaseBEVBackbone": BaseBEVBackbone, "BaseBEVBackbone_Scale": BaseBEVBackbone_Scale}      # backbones_2d/__init__.py:3-6
