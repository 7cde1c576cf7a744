"""GPU parity of row N1 (BaseBEVBackbone_Scale on tcgen05) against the oracle / golden fixture, through the C ABI.

Tolerances (stated here because north_star gives none for this row): activations and weights are bf16 with fp32
accumulation, so a single layer must match an fp32 convolution of the SAME bf16-rounded operands to bf16 output rounding
(2^-8 relative to the tensor maximum), and the whole 24-convolution backbone must match the fp32 reference within
TOL_BACKBONE = 2e-2 (max-norm, relative to max|ref|) and 1e-2 in relative L2.
"""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import GOLDEN, rel_err

pytestmark = pytest.mark.gpu

TOL_LAYER = 2.0 ** -8
TOL_BACKBONE = 2e-2
TOL_BACKBONE_L2 = 1e-2


def _lib():
    from hvpr_b200 import _lib
    _lib.init_device()
    return _lib


def _pack(w_ntc, bn):
    L = _lib()
    n_total, taps, cin = w_ntc.shape
    out = torch.empty(L.lib().hvpr_conv_packed_bytes(n_total, taps, cin), dtype=torch.uint8, device="cuda")
    L.check(L.lib().hvpr_conv_pack_weights(L.ptr(w_ntc.contiguous()), n_total, taps, cin, bn, L.ptr(out), L.cur_stream()))
    return out


def _conv_call(x_nhwc, wpk, n_total, bn, bias, ksize, stride, c_in, out, *, relu=True, gate=None, residual=None,
               out_mode=0, out_c_off=0, up=1, c_out=0, out_ctot=0):
    L = _lib()
    a = L.HvprConvArgs()
    n, h, w, cs = x_nhwc.shape
    a.in_, a.n, a.h_in, a.w_in, a.in_cs, a.c_in = x_nhwc.data_ptr(), n, h, w, cs, c_in
    a.ksize, a.stride, a.w_packed, a.n_total, a.bn = ksize, stride, wpk.data_ptr(), n_total, bn
    a.bias, a.relu = (bias.data_ptr() if bias is not None else None), int(relu)
    a.gate = gate.data_ptr() if gate is not None else None
    a.residual = residual.data_ptr() if residual is not None else None
    a.res_cs = residual.shape[-1] if residual is not None else 0
    a.out_mode, a.out = out_mode, out.data_ptr()
    a.out_cs = out.shape[-1] if out_mode == 0 else 0
    a.out_c_off, a.up, a.c_out, a.out_ctot = out_c_off, up, c_out or n_total, out_ctot
    L.check(L.lib().hvpr_conv2d(ctypes.byref(a), L.cur_stream()), "hvpr_conv2d")
    torch.cuda.synchronize()


def _rand_case(seed, n, h, w, cin, cout, k):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(n, cin, h, w, generator=g).cuda().bfloat16()
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda().bfloat16()
    b = torch.randn(cout, generator=g).cuda() * 0.1
    return x, wt, b


@pytest.mark.parametrize("n,h,w,cin,cout,stride,bn", [
    (1, 16, 8, 64, 32, 1, 32),          # exactly one 128-pixel patch
    (2, 20, 28, 128, 128, 1, 128),      # ragged patches: TMA zero fill on every side + masked stores
    (1, 24, 40, 128, 256, 2, 256),      # stride 2 through the four parity tensor maps
    (2, 12, 12, 256, 256, 1, 256),
    (1, 10, 36, 64, 64, 2, 64),
    (1, 9, 130, 128, 128, 1, 128),      # wider than one patch row, odd height
])
def test_conv3x3_layer_matches_fp32_conv_of_same_bf16_operands(n, h, w, cin, cout, stride, bn):
    x, wt, b = _rand_case(n * 1000 + h, n, h, w, cin, cout, 3)
    ref = F.relu(F.conv2d(x.float(), wt.float(), b, stride=stride, padding=1))
    w_ntc = wt.float().permute(0, 2, 3, 1).reshape(cout, 9, cin)
    wpk = _pack(w_ntc, bn)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    out = torch.full((n, h // stride, w // stride, cout + 8), 7.0, dtype=torch.bfloat16, device="cuda")   # +8: channel stride > C
    _conv_call(x_nhwc, wpk, cout, bn, b, 3, stride, cin, out)
    got = out[..., :cout].permute(0, 3, 1, 2).float()
    assert rel_err(got, ref)[0] <= TOL_LAYER, rel_err(got, ref)
    assert torch.all(out[..., cout:] == 7.0)            # pad channels untouched


@pytest.mark.parametrize("n,h,w,cin,cout,stride,bn", [(2, 20, 28, 128, 128, 1, 128), (1, 24, 40, 64, 64, 2, 64), (1, 37, 50, 64, 32, 1, 32)])
def test_conv3x3_two_subtile_mode(n, h, w, cin, cout, stride, bn):
    """bn <= 128 layers on large canvases process two 128-pixel patches per weight block; force that policy on a small case."""
    L = _lib()
    if stride == 2 and (h % 2 or w % 2):
        h, w = h + h % 2, w + w % 2
    x, wt, b = _rand_case(99 + h, n, h, w, cin, cout, 3)
    ref = F.relu(F.conv2d(x.float(), wt.float(), b, stride=stride, padding=1))
    wpk = _pack(wt.float().permute(0, 2, 3, 1).reshape(cout, 9, cin), bn)
    out = torch.empty(n, h // stride, w // stride, cout, dtype=torch.bfloat16, device="cuda")
    assert L.lib().hvpr_dbg_conv_force_msub(2) == 0
    try:
        _conv_call(x.permute(0, 2, 3, 1).contiguous(), wpk, cout, bn, b, 3, stride, cin, out)
    finally:
        L.lib().hvpr_dbg_conv_force_msub(0)
    assert rel_err(out.permute(0, 3, 1, 2).float(), ref)[0] <= TOL_LAYER


@pytest.mark.parametrize("msub", [1, 2])
@pytest.mark.parametrize("n,h,w,cin,cout,bn", [(2, 20, 28, 128, 128, 128), (1, 33, 19, 256, 256, 256), (1, 16, 8, 64, 32, 32)])
def test_conv3x3_halo_operand_path(msub, n, h, w, cin, cout, bn):
    """Optional operand path for stride-1 3x3 layers: three column-shifted 8 x (rows+2) patches per k-block feed all nine taps
    through row-shifted (still 1024-B aligned) UMMA descriptors: 3.4x instead of 9x the activation traffic; off by default."""
    L = _lib()
    x, wt, b = _rand_case(7 + h, n, h, w, cin, cout, 3)
    ref = F.relu(F.conv2d(x.float(), wt.float(), b, padding=1))
    wpk = _pack(wt.float().permute(0, 2, 3, 1).reshape(cout, 9, cin), bn)
    out = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
    L.lib().hvpr_dbg_conv_halo_off(0)
    L.lib().hvpr_dbg_conv_force_msub(msub)
    try:
        _conv_call(x.permute(0, 2, 3, 1).contiguous(), wpk, cout, bn, b, 3, 1, cin, out)
    finally:
        L.lib().hvpr_dbg_conv_halo_off(1)
        L.lib().hvpr_dbg_conv_force_msub(0)
    assert rel_err(out.permute(0, 3, 1, 2).float(), ref)[0] <= TOL_LAYER


@pytest.mark.parametrize("n,h,w,cin,cout,stride,bn,ksize", [
    (1, 16, 16, 64, 64, 1, 64, 3),         # one super-tile, one CTA pair
    (2, 20, 28, 128, 128, 1, 128, 3),
    (1, 40, 24, 256, 256, 1, 256, 3),
    (1, 24, 40, 128, 256, 2, 256, 3),
    (3, 9, 70, 64, 32, 1, 32, 3),
])
@pytest.mark.parametrize("msub", [0, 2])
def test_conv_cta_pair_kernel(n, h, w, cin, cout, stride, bn, ksize, msub):
    """cta_group::2 variant: UMMA M = 256 across a CTA pair, each CTA loading half of every weight block."""
    L = _lib()
    x, wt, b = _rand_case(17 + h, n, h, w, cin, cout, ksize)
    ref = F.relu(F.conv2d(x.float(), wt.float(), b, stride=stride, padding=1))
    wpk = _pack(wt.float().permute(0, 2, 3, 1).reshape(cout, 9, cin), bn)
    out = torch.empty(n, h // stride, w // stride, cout, dtype=torch.bfloat16, device="cuda")
    assert L.lib().hvpr_dbg_conv_pair(2) == 0
    L.lib().hvpr_dbg_conv_force_msub(msub)          # 2: two sub-tiles per CTA when bn <= 128 (M = 512 per pair and weight block)
    try:
        _conv_call(x.permute(0, 2, 3, 1).contiguous(), wpk, cout, bn, b, 3, stride, cin, out)
    finally:
        L.lib().hvpr_dbg_conv_pair(0); L.lib().hvpr_dbg_conv_force_msub(0)
    assert rel_err(out.permute(0, 3, 1, 2).float(), ref)[0] <= TOL_LAYER


@pytest.mark.parametrize("msub", [1, 2])
@pytest.mark.parametrize("n,h,w,cin,cout,bn", [(2, 40, 28, 128, 128, 128), (1, 33, 19, 256, 256, 256), (1, 70, 9, 64, 64, 64)])
def test_conv_cta_pair_kernel_with_halo_operands(msub, n, h, w, cin, cout, bn):
    """CTA pair + halo patches: each CTA loads its patch triple per k-block and half of every weight block (least L2 traffic)."""
    L = _lib()
    x, wt, b = _rand_case(23 + h, n, h, w, cin, cout, 3)
    ref = F.relu(F.conv2d(x.float(), wt.float(), b, padding=1))
    wpk = _pack(wt.float().permute(0, 2, 3, 1).reshape(cout, 9, cin), bn)
    out = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
    L.lib().hvpr_dbg_conv_pair(2); L.lib().hvpr_dbg_conv_halo_off(0); L.lib().hvpr_dbg_conv_force_msub(msub)
    try:
        _conv_call(x.permute(0, 2, 3, 1).contiguous(), wpk, cout, bn, b, 3, 1, cin, out)
    finally:
        L.lib().hvpr_dbg_conv_pair(0); L.lib().hvpr_dbg_conv_halo_off(1); L.lib().hvpr_dbg_conv_force_msub(0)
    assert rel_err(out.permute(0, 3, 1, 2).float(), ref)[0] <= TOL_LAYER


def test_bev_fill_nhwc_bf16_matches_the_nchw_fill():
    """K4 in channels-last bf16 (what the backbone consumes) == bf16 rounding of the fp32 NCHW canvases."""
    L = _lib()
    torch.manual_seed(3)
    B, nx, ny, P = 2, 24, 20, 150
    cells = torch.randperm(B * nx * ny)[:P]
    cmap = torch.full((B * nx * ny,), -1, dtype=torch.int32)
    cmap[cells] = torch.arange(P, dtype=torch.int32)
    cmap = cmap.cuda()
    fa, fb, fs = torch.randn(P, 64).cuda(), torch.randn(P, 64).cuda(), torch.randn(P, 32).cuda()
    sp = torch.empty(B, 128, ny, nx, device="cuda"); sc = torch.empty(B, 32, ny, nx, device="cuda")
    L.check(L.lib().hvpr_bev_fill(L.ptr(fa), 64, L.ptr(fb), 64, L.ptr(fs), 32, L.ptr(cmap), B, nx, ny, L.ptr(sp), L.ptr(sc), None, L.cur_stream()))
    xo = torch.full((B, ny, nx, 128), 9.0, dtype=torch.bfloat16, device="cuda")
    yo = torch.full((B, ny, nx, 64), 9.0, dtype=torch.bfloat16, device="cuda")
    L.check(L.lib().hvpr_bev_fill_nhwc_bf16(L.ptr(fa), 64, L.ptr(fb), 64, L.ptr(fs), 32, L.ptr(cmap), B, nx, ny,
                                            L.ptr(xo), 128, L.ptr(yo), 64, L.cur_stream()))
    assert torch.equal(xo, sp.permute(0, 2, 3, 1).bfloat16())
    assert torch.equal(yo[..., :32], sc.permute(0, 2, 3, 1).bfloat16()) and torch.all(yo[..., 32:] == 0)


def test_conv_gate_and_residual_epilogue():
    n, h, w, c = 2, 12, 20, 128
    x, wt, b = _rand_case(5, n, h, w, c, c, 3)
    gate = torch.rand(n, h, w, device="cuda")
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    ref = gate[:, None] * F.relu(F.conv2d(x.float(), wt.float(), b, padding=1)) + x.float()
    wpk = _pack(wt.float().permute(0, 2, 3, 1).reshape(c, 9, c), 128)
    out = torch.empty(n, h, w, c, dtype=torch.bfloat16, device="cuda")
    _conv_call(x_nhwc, wpk, c, 128, b, 3, 1, c, out, gate=gate, residual=x_nhwc)
    assert rel_err(out.permute(0, 3, 1, 2).float(), ref)[0] <= TOL_LAYER


@pytest.mark.parametrize("up,cin", [(1, 128), (2, 256), (4, 512)])
def test_transposed_conv_pixel_shuffle_into_nchw_slice(up, cin):
    n, h, w, cout, ctot, coff = 2, 6, 10, 128, 384, 128
    g = torch.Generator(device="cpu").manual_seed(up)
    x = torch.randn(n, cin, h, w, generator=g).cuda().bfloat16()
    wt = (torch.randn(cin, cout, up, up, generator=g) / cin ** 0.5).cuda().bfloat16()
    b = torch.randn(cout, generator=g).cuda() * 0.1
    ref = F.relu(F.conv_transpose2d(x.float(), wt.float(), b, stride=up))
    w_ntc = wt.float().permute(2, 1, 3, 0).reshape(up * up * cout, 1, cin)      # column (dy*cout + co)*up + dx
    bn = 256 if (up * up * cout) % 256 == 0 else 128
    wpk = _pack(w_ntc, bn)
    out = torch.full((n, ctot, h * up, w * up), -3.0, device="cuda")
    _conv_call(x.permute(0, 2, 3, 1).contiguous(), wpk, up * up * cout, bn, b.repeat_interleave(up).repeat(up), 1, 1, cin, out,
               out_mode=1, out_c_off=coff, up=up, c_out=cout, out_ctot=ctot)
    assert rel_err(out[:, coff:coff + cout], ref)[0] <= 1e-5         # fp32 output: only the summation order differs
    assert torch.all(out[:, :coff] == -3.0) and torch.all(out[:, coff + cout:] == -3.0)


def test_attention_gate_matches_oracle():
    from oracle import backbone as ob
    L = _lib()
    w = ob.random_backbone_weights(3)
    n, h, wd, c = 2, 10, 14, 32
    y = torch.rand(n, c, h, wd).bfloat16()
    ref = ob.attention_gate(y.float(), w)[:, 0]
    cw, s = torch.from_numpy(w["attention.spatial.conv.weight"]).double(), None
    s = torch.from_numpy(w["attention.spatial.norm.weight"]).double() / torch.sqrt(
        torch.from_numpy(w["attention.spatial.norm.running_var"]).double() + 1e-3)
    shift = torch.from_numpy(w["attention.spatial.norm.bias"]).double() + (
        torch.from_numpy(w["attention.spatial.conv.bias"]).double() - torch.from_numpy(w["attention.spatial.norm.running_mean"]).double()) * s
    w18 = (ctypes.c_float * 18)(*[float(v) for v in (cw * s).reshape(-1)])
    y_nhwc = torch.zeros(n, h, wd, 64, dtype=torch.bfloat16, device="cuda")
    y_nhwc[..., :c] = y.permute(0, 2, 3, 1).cuda()
    pooled = torch.empty(n, h, wd, 2, device="cuda")
    gate = torch.empty(n, h, wd, device="cuda")
    L.check(L.lib().hvpr_attention_gate(L.ptr(y_nhwc), n, h, wd, 64, c, w18, float(shift[0]), L.ptr(pooled), L.ptr(gate),
                                        L.cur_stream()))
    assert rel_err(gate, ref)[0] <= 1e-5


def test_nchw_to_nhwc_bf16_is_exact_rounding():
    L = _lib()
    x = torch.randn(2, 32, 7, 45, device="cuda")
    out = torch.zeros(2, 7, 45, 64, dtype=torch.bfloat16, device="cuda")
    L.check(L.lib().hvpr_nchw_to_nhwc_bf16(L.ptr(x), 2, 32, 7, 45, L.ptr(out), 64, L.cur_stream()))
    assert torch.equal(out[..., :32], x.permute(0, 2, 3, 1).bfloat16()) and torch.all(out[..., 32:] == 0)


def _backbone(wseed):
    from hvpr_b200.backbone import BaseBEVBackbone_Scale
    from hvpr_b200.config import Cfg
    from oracle import backbone as ob
    m = BaseBEVBackbone_Scale(Cfg(NAME="BaseBEVBackbone_Scale", **ob.CFG), 128).cuda().eval()
    w = ob.random_backbone_weights(wseed)
    missing = m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    assert not missing.unexpected_keys and all(k.endswith("num_batches_tracked") for k in missing.missing_keys)
    return m, w


def test_backbone_matches_reference_golden():
    """tests/golden/backbone_tiny.npz = output of the reference's own module (oracle/make_golden_backbone.py)."""
    from oracle import backbone as ob
    z = np.load(os.path.join(GOLDEN, "backbone_tiny.npz"))
    B, H, W = (int(v) for v in z["shape"])
    m, _ = _backbone(int(z["wseed"]))
    spatial, scale = ob.random_canvases(int(z["xseed"]), B, H, W)
    with torch.no_grad():
        out = m({"spatial_features": torch.from_numpy(spatial).cuda(), "spatial_scale_features": torch.from_numpy(scale).cuda()})
    got, ref = out["spatial_features_2d"], torch.from_numpy(z["spatial_features_2d"])
    assert got.shape == ref.shape and got.dtype == torch.float32
    e = rel_err(got, ref)
    assert e[0] <= TOL_BACKBONE and e[1] <= TOL_BACKBONE_L2, e


def test_backbone_matches_oracle_on_a_ragged_batch():
    """batch of 2, 40 x 72 canvas: patches straddle image borders at every level (40/4 = 10 rows, 72/4 = 18 columns)."""
    from oracle import backbone as ob
    m, w = _backbone(11)
    spatial, scale = ob.random_canvases(12, 2, 40, 72)
    ref = torch.from_numpy(ob.backbone_forward(w, spatial, scale))
    with torch.no_grad():
        out = m({"spatial_features": torch.from_numpy(spatial).cuda(), "spatial_scale_features": torch.from_numpy(scale).cuda()})
    e = rel_err(out["spatial_features_2d"], ref)
    assert e[0] <= TOL_BACKBONE and e[1] <= TOL_BACKBONE_L2, e


def test_points_to_spatial_features_2d_pipeline_vs_oracle_chain():
    """raw points -> K1..K4 (channels-last bf16 canvases) -> backbone, one CUDA graph, vs oracle front end + oracle backbone."""
    from helpers import load_small, to_dev
    from hvpr_b200.pipeline import FrontEndWithBackbone
    from oracle import backbone as ob, hybrid
    z, geom, frames, overflow, wseed = load_small("tiny_continue")
    w_fe, w_bb = hybrid.random_weights(wseed), ob.random_backbone_weights(21)
    o = hybrid.frontend(frames, geom, w_fe, overflow)
    ref = torch.from_numpy(ob.backbone_forward(w_bb, o["spatial_features"].numpy(), o["spatial_scale_features"].numpy()))
    pipe = FrontEndWithBackbone(geom, overflow=overflow)
    pipe.frontend.load_reference_weights({k: torch.from_numpy(np.asarray(v)) if not torch.is_tensor(v) else v for k, v in w_fe.items()})
    pipe.backbone_2d.load_state_dict({k: torch.from_numpy(v) for k, v in w_bb.items()}, strict=False)
    pts, off = to_dev(frames)
    p = pipe.plan(len(frames), pts.shape[0])
    p.points.copy_(pts); p.frame_offsets.copy_(off)
    for _ in range(2):                      # graph replay must be idempotent
        pipe.run()
    torch.cuda.synchronize()
    e = rel_err(p.out, ref)
    assert tuple(p.out.shape) == tuple(ref.shape)
    assert e[0] <= TOL_BACKBONE and e[1] <= TOL_BACKBONE_L2, e
    # module-style entry point gives the same tensor
    from hvpr_b200 import synth
    bd = pipe({"points": torch.from_numpy(synth.collate_points(frames)).cuda(), "batch_size": len(frames)})
    torch.cuda.synchronize()
    assert torch.equal(bd["spatial_features_2d"], p.out)


def test_full_size_backbone_is_invariant_to_the_kernel_policy():
    """BASELINE-size canvases (296 x 248, batch 2): the automatic policy (CTA-pair kernel on the 256-column layers, two sub-tiles
    on the 128-column level) must agree with the plain single-CTA / one-sub-tile kernels — same operands, same K order — and
    the result must be a sane feature map (size-independent property check; the oracle comparison runs on small canvases)."""
    from oracle import backbone as ob
    L = _lib()
    m, _ = _backbone(5)
    B, H, W = 2, 248, 296
    spatial, scale = ob.random_canvases(9, B, H, W, occupancy=0.12)
    bd = {"spatial_features": torch.from_numpy(spatial).cuda(), "spatial_scale_features": torch.from_numpy(scale).cuda()}
    with torch.no_grad():
        auto = m(dict(bd))["spatial_features_2d"].clone()
        L.lib().hvpr_dbg_conv_pair(1); L.lib().hvpr_dbg_conv_force_msub(1)
        try:
            plain = m(dict(bd))["spatial_features_2d"].clone()
        finally:
            L.lib().hvpr_dbg_conv_pair(0); L.lib().hvpr_dbg_conv_force_msub(0)
    torch.cuda.synchronize()
    assert auto.shape == (B, 384, H, W) and torch.isfinite(auto).all() and float(auto.min()) >= 0.0      # ReLU outputs
    assert float(auto.abs().max()) > 0.1
    e = rel_err(auto, plain)
    assert e[0] <= 2.0 ** -7 and e[1] <= 1e-3, e


def test_plain_backbone_matches_reference_golden():
    """The plain BaseBEVBackbone (PointPillars layout: 64 channels in, every level stride 2) on the same kernels vs the
    reference module's output; exercises a stride-2 first level and deblocks that land on half the input resolution."""
    from hvpr_b200.backbone import BaseBEVBackbone
    from hvpr_b200.config import Cfg
    from oracle import backbone as ob
    z = np.load(os.path.join(GOLDEN, "backbone_plain_tiny.npz"))
    w = ob.random_backbone_weights(int(z["wseed"]), ob.PLAIN_CFG, 64, with_scale=False)
    m = BaseBEVBackbone(Cfg(NAME="BaseBEVBackbone", **ob.PLAIN_CFG), 64).cuda().eval()
    r = m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    assert not r.unexpected_keys and all(k.endswith("num_batches_tracked") for k in r.missing_keys)
    with torch.no_grad():
        out = m({"spatial_features": torch.from_numpy(z["spatial_features"]).cuda()})["spatial_features_2d"]
    ref = torch.from_numpy(z["spatial_features_2d"])
    assert tuple(out.shape) == tuple(ref.shape)
    e = rel_err(out, ref)
    assert e[0] <= TOL_BACKBONE and e[1] <= TOL_BACKBONE_L2, e


# ---- row N2: AnchorHeadSingle -------------------------------------------------------------------------------------------------
def _head(seed, grid, rng):
    from hvpr_b200.config import Cfg
    from hvpr_b200.dense_head import AnchorHeadSingle
    from oracle import dense_head as od
    m = AnchorHeadSingle(Cfg(**od.HEAD_CFG), 384, 1, ["Car"], list(grid), rng).cuda().eval()
    w = od.random_head_weights(seed)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    return m, w


@pytest.mark.parametrize("B,H,W", [(2, 20, 24), (1, 37, 50)])
def test_dense_head_matches_oracle(B, H, W):
    """conv_cls / conv_box / conv_dir_cls as one tcgen05 GEMM + decode vs the oracle (whose decode / anchors / limit_period are the
    reference's own functions).  The GEMM operands are bf16 (fp32 accumulate, fp32 output): logits and box deltas carry ~2^-8 of
    their scale, so boxes are compared with an absolute tolerance of 2e-2 of the anchor diagonal / size, and anchors whose two
    direction logits are closer than 0.05 are excluded from the heading check (the argmax may legitimately flip there)."""
    from oracle import dense_head as od
    rng = [0, -39.68, -3, 69.12, 39.68, 1]
    m, w = _head(7, (W, H, 1), rng)
    x = np.abs(np.random.default_rng(8).standard_normal((B, 384, H, W))).astype(np.float32) * (np.random.default_rng(9).random((B, 1, H, W)) < 0.6)
    x = x.astype(np.float32)
    cls_ref, box_ref, (_, _, dir_raw) = od.head_forward(w, x, od.HEAD_CFG, (W, H, 1), rng, return_raw=True)
    with torch.no_grad():
        out = m({"spatial_features_2d": torch.from_numpy(x).cuda()})
    torch.cuda.synchronize()
    cls, box = out["batch_cls_preds"].cpu().numpy(), out["batch_box_preds"].cpu().numpy()
    assert cls.shape == cls_ref.shape and box.shape == box_ref.shape and out["cls_preds_normalized"] is False
    assert np.abs(cls - cls_ref).max() <= 2e-2 * max(1.0, np.abs(cls_ref).max())
    diag = float(np.sqrt(3.9 ** 2 + 1.6 ** 2))
    assert np.abs(box[..., :2] - box_ref[..., :2]).max() <= 2e-2 * diag
    assert np.abs(box[..., 2] - box_ref[..., 2]).max() <= 2e-2 * 1.56
    assert np.abs(box[..., 3:6] / box_ref[..., 3:6] - 1).max() <= 2e-2
    d = dir_raw.reshape(B, -1, 2)
    clear = np.abs(d[..., 0] - d[..., 1]) > 0.05
    dr = np.abs(box[..., 6] - box_ref[..., 6])[clear]
    dr = np.minimum(dr, np.abs(dr - np.pi))          # a heading within 2e-2 of a period boundary may wrap by one period (pi)
    assert dr.max() <= 2e-2 and clear.mean() > 0.8


def test_dense_head_rejects_the_shipped_anchor_stride_mismatch():
    """hvpr.yaml puts anchors on a stride-2 map under a full-resolution backbone (B10): fail loudly instead of mis-indexing."""
    from hvpr_b200 import _lib
    from hvpr_b200.config import Cfg
    from hvpr_b200.dense_head import AnchorHeadSingle
    from oracle import dense_head as od
    cfg = dict(od.HEAD_CFG, ANCHOR_GENERATOR_CONFIG=[dict(od.HEAD_CFG["ANCHOR_GENERATOR_CONFIG"][0], feature_map_stride=2)])
    m = AnchorHeadSingle(Cfg(**cfg), 384, 1, ["Car"], [48, 40, 1], [0, -39.68, -3, 69.12, 39.68, 1]).cuda().eval()
    with pytest.raises(_lib.HvprError):
        m({"spatial_features_2d": torch.zeros(1, 384, 40, 48, device="cuda")})


def test_points_to_boxes_pipeline_vs_oracle_chain():
    """raw points -> front end -> backbone (channels-last features, fp32 NCHW never written) -> dense head, one CUDA graph."""
    from helpers import load_small, to_dev
    from hvpr_b200.pipeline import HVPR_HEAD_CFG, FrontEndWithBackbone
    from oracle import backbone as ob, dense_head as od, hybrid
    z, geom, frames, overflow, wseed = load_small("tiny_continue")
    w_fe, w_bb, w_hd = hybrid.random_weights(wseed), ob.random_backbone_weights(21), od.random_head_weights(22)
    o = hybrid.frontend(frames, geom, w_fe, overflow)
    f2d = ob.backbone_forward(w_bb, o["spatial_features"].numpy(), o["spatial_scale_features"].numpy())
    cls_ref, box_ref, (_, _, dir_raw) = od.head_forward(w_hd, f2d, od.HEAD_CFG, geom.grid_size, list(geom.point_cloud_range), return_raw=True)
    pipe = FrontEndWithBackbone(geom, overflow=overflow, head_cfg=HVPR_HEAD_CFG)
    pipe.frontend.load_reference_weights(w_fe)
    pipe.backbone_2d.load_state_dict({k: torch.from_numpy(v) for k, v in w_bb.items()}, strict=False)
    pipe.dense_head.load_state_dict({k: torch.from_numpy(v) for k, v in w_hd.items()})
    pts, off = to_dev(frames)
    p = pipe.plan(len(frames), pts.shape[0])
    p.points.copy_(pts); p.frame_offsets.copy_(off)
    for _ in range(2):
        pipe.run()
    torch.cuda.synchronize()
    cls, box = p.cls_preds.cpu().numpy(), p.box_preds.cpu().numpy()
    assert cls.shape == cls_ref.shape and box.shape == box_ref.shape
    # the features feeding the head already carry the backbone's bf16 tolerance (2e-2 of their maximum)
    fscale = float(np.abs(f2d).max())
    assert np.abs(cls - cls_ref).max() <= 3e-2 * max(1.0, np.abs(cls_ref).max(), fscale)
    assert np.abs(box[..., :2] - box_ref[..., :2]).max() <= 5e-2 * float(np.sqrt(3.9 ** 2 + 1.6 ** 2))
    assert np.abs(box[..., 3:6] / box_ref[..., 3:6] - 1).max() <= 5e-2


def test_pipeline_plans_for_a_capacity_and_follows_weight_reloads():
    """(1) collated batches differ in point count every time -> the module entry point plans ONCE for a capacity and
    replays the same graph (no per-batch reallocation / re-capture); (2) a load_state_dict() after the first forward must reach
    the captured graph (folded PFN weights are baked in by value); (3) outputs are fresh tensors, not views of plan buffers."""
    from helpers import load_small
    from hvpr_b200 import synth
    from hvpr_b200.pipeline import FrontEndWithBackbone
    from oracle import backbone as ob, hybrid
    z, geom, frames, overflow, wseed = load_small("tiny_continue")
    w_a, w_b, w_bb = hybrid.random_weights(wseed), hybrid.random_weights(wseed + 1), ob.random_backbone_weights(21)

    def make(w):
        pipe = FrontEndWithBackbone(geom, overflow=overflow)
        pipe.frontend.load_reference_weights(w)
        pipe.backbone_2d.load_state_dict({k: torch.from_numpy(v) for k, v in w_bb.items()}, strict=False)
        return pipe

    batch1 = frames
    batch2 = [f[: max(1, len(f) * 2 // 3)] for f in frames]           # fewer points: fits the planned capacity
    bd = lambda fr: {"points": torch.from_numpy(synth.collate_points(fr)).cuda(), "batch_size": len(fr)}
    pipe = make(w_a)
    out1 = pipe(bd(batch1))["spatial_features_2d"]
    plan = pipe._p
    out2 = pipe(bd(batch2))["spatial_features_2d"]
    assert pipe._p is plan and plan.graph is not None                 # same buffers, same graph
    assert out1.data_ptr() != out2.data_ptr() and not torch.equal(out1, out2)
    ref2 = make(w_a)(bd(batch2))["spatial_features_2d"]
    assert torch.equal(out2, ref2)
    ref1 = make(w_a)(bd(batch1))["spatial_features_2d"]
    assert torch.equal(out1, ref1)                                    # out1 was not clobbered by the second call
    # weight reload after the first forward
    pipe.frontend.load_reference_weights(w_b)
    out3 = pipe(bd(batch2))["spatial_features_2d"]
    ref3 = make(w_b)(bd(batch2))["spatial_features_2d"]
    torch.cuda.synchronize()
    assert torch.equal(out3, ref3) and not torch.equal(out3, out2)
