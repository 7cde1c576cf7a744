"""CPU suite: the C-ABI library loads and exports every symbol include/hvpr_b200.h declares (no compute calls),
and the host-side mirror of the reference interface behaves (cfg shim, registries, state_dict names, BN folding,
loud failure without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import hvpr_b200
from hvpr_b200 import _lib, config
from hvpr_b200.geometry import G1, G2
from oracle import hybrid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "hvpr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(hvpr_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    assert sorted(_lib.SYMBOLS) == declared
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared:
        assert hasattr(L, s), s
    assert _lib.lib().hvpr_version() >= 100
    assert _lib.lib().hvpr_strerror(0) == b"ok" and b"workspace" in _lib.lib().hvpr_strerror(-3)


def test_struct_layouts_match_header(tmp_path):
    # the ctypes mirrors against the C compiler's view of include/hvpr_b200.h
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "hvpr_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(HvprGeom), sizeof(HvprPfnWeights), sizeof(HvprLaunchCfg), '
                   'sizeof(HvprZeroFill), offsetof(HvprZeroFill, bytes), offsetof(HvprZeroFill, n), sizeof(HvprConvArgs), offsetof(HvprLaunchCfg, variant)); return 0; }\n')
    exe = tmp_path / "layout"
    import subprocess
    subprocess.run(["gcc", "-I", os.path.join(os.path.dirname(_lib._HERE), "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    Z = _lib.HvprZeroFill
    assert got == [ctypes.sizeof(_lib.HvprGeom), ctypes.sizeof(_lib.HvprPfnWeights), ctypes.sizeof(_lib.HvprLaunchCfg), ctypes.sizeof(Z),
                   Z.bytes.offset, Z.n.offset, ctypes.sizeof(_lib.HvprConvArgs), _lib.HvprLaunchCfg.variant.offset], got
    assert ctypes.sizeof(_lib.HvprGeom) == 36
    assert ctypes.sizeof(_lib.HvprPfnWeights) == 4 * (160 + 16 + 1024 + 1024 + 64 + 80 + 16 + 512 + 32)


def test_voxelizer_workspace_covers_both_table_layouts():
    """The workspace query is valid for the dense table AND the open-addressing one, and stays small for a grid of 2.7e9 cells."""
    L = _lib.lib()
    g2 = _lib.make_geom(G2.range_f32, G2.voxel_f32, G2.grid_size)
    dense = L.hvpr_voxelize_workspace_bytes(960000, 8, ctypes.byref(g2), 40000)
    assert dense > 8 * 432 * 496 * 8                                    # holds the dense {first, count} table of 8 frames
    huge = _lib.make_geom((0.0, -39.68, -3.0), (0.02, 0.02, 0.02), (3456, 3968, 200))
    hashed = L.hvpr_voxelize_workspace_bytes(960000, 8, ctypes.byref(huge), 40000)
    assert 0 < hashed < (1 << 30)                                        # 2 x 960k -> 2^21 slots x 16 B x 8 frames + point-sized arrays


def test_workspace_query_needs_no_gpu():
    g = _lib.make_geom(G2.range_f32, G2.voxel_f32, G2.grid_size)
    n = _lib.lib().hvpr_voxelize_workspace_bytes(8 * 120000, 8, ctypes.byref(g), 40000)
    assert 8 * 432 * 496 * 8 < n < 400 * 2 ** 20           # covers the open-addressing table too (2^21 slots x 16 B x 8 frames)
    assert _lib.lib().hvpr_voxelize_workspace_bytes(-1, 8, ctypes.byref(g), 40000) == 0


def test_argument_validation_without_gpu():
    L = _lib.lib()
    g = _lib.make_geom(G2.range_f32, G2.voxel_f32, G2.grid_size)
    # null pointers / bad sizes are rejected before any CUDA call
    assert L.hvpr_voxelize(None, 10, 4, 0, None, 1, 0, ctypes.byref(g), 32, 100, 0, None, None, None, None, None, None, 0, None) == -1
    assert L.hvpr_bev_fill(None, 64, None, 0, None, 0, None, 1, 432, 496, None, None, None, None) == -1
    assert L.hvpr_mem_attn(None, None, -1, None, None, 2000, 64, 20, 0, None, None, None, 0, None, None) == -1


def test_registries_and_state_dict_names():
    from hvpr_b200 import map_to_bev, vfe
    assert set(vfe.__all__) >= {"VFETemplate", "PillarVFE", "PillarVFE_Scale"}
    assert set(map_to_bev.__all__) >= {"PointPillarScatter", "PointPillarScatter_Agg_Memory_1_scale"}
    m = vfe.__all__["PillarVFE_Scale"](model_cfg=config.HVPR_VFE_CFG, num_point_features=4,
                                       point_cloud_range=G1.range_f32, voxel_size=list(G1.voxel_size))
    assert m.get_output_feature_dim() == 64
    b = map_to_bev.__all__["PointPillarScatter_Agg_Memory_1_scale"](model_cfg=config.HVPR_BEV_CFG, grid_size=G1.grid_size)
    assert b.num_bev_features == 128
    w = hybrid.random_weights(0)
    names = {k[4:] for k in w if k.startswith("vfe.")}
    assert names <= set(m.state_dict().keys())
    assert set(b.state_dict().keys()) == {"memory.weight"}
    assert tuple(b.memory.weight.shape) == (2000, 64)
    assert float(b.memory.weight.detach().abs().max()) <= 0.125 + 1e-6          # memory_module.py:23-25
    # shapes of the reference's parameters
    sd = m.state_dict()
    assert tuple(sd["pfn_layers.0.linear.weight"].shape) == (16, 10)
    assert tuple(sd["pfn_layers.1.linear.weight"].shape) == (64, 32)
    assert tuple(sd["pfn_scale_layers.0.0.weight"].shape) == (16, 5)
    assert tuple(sd["pfn_scale_layers.1.0.weight"].shape) == (32, 16)


def test_unsupported_cfg_is_loud():
    from hvpr_b200 import vfe
    bad = config.Cfg(config.HVPR_VFE_CFG); bad["NUM_FILTERS"] = [64]
    with pytest.raises(NotImplementedError):
        vfe.PillarVFE(bad, 4, list(G1.voxel_size), G1.range_f32)


def test_bn_folding_matches_oracle():
    """The folded (W', b') the kernel receives reproduce the oracle's Linear+BN on CPU."""
    from hvpr_b200 import vfe
    w = hybrid.random_weights(5)
    m = vfe.PillarVFE_Scale(config.HVPR_VFE_CFG, 4, list(G1.voxel_size), G1.range_f32).eval()
    m.load_state_dict({k[4:]: v for k, v in w.items() if k.startswith("vfe.")}, strict=False)
    W = m._weights()
    w0 = torch.tensor(list(W.w0)).view(16, 10); b0 = torch.tensor(list(W.b0))
    x = torch.randn(100, 10)
    ref = hybrid._bn_eval(torch.nn.functional.linear(x, w["vfe.pfn_layers.0.linear.weight"]).unsqueeze(-1), w,
                          "vfe.pfn_layers.0.norm").squeeze(-1)
    torch.testing.assert_close(x @ w0.t() + b0, ref, rtol=1e-5, atol=1e-5)
    w1 = torch.cat([torch.tensor(list(W.w1a)).view(64, 16), torch.tensor(list(W.w1b)).view(64, 16)], 1)
    x = torch.randn(100, 32)
    ref = hybrid._bn_eval(torch.nn.functional.linear(x, w["vfe.pfn_layers.1.linear.weight"]).unsqueeze(-1), w,
                          "vfe.pfn_layers.1.norm").squeeze(-1)
    torch.testing.assert_close(x @ w1.t() + torch.tensor(list(W.b1)), ref, rtol=1e-5, atol=1e-5)
    # cache invalidation on weight update
    with torch.no_grad():
        m.pfn_layers[0].linear.weight.mul_(2.0)      # (raw `.data` edits bypass version counters: call invalidate_weights())
    W2 = m._weights()
    assert abs(W2.w0[0] - 2 * w0.view(-1)[0].item()) < 1e-6


def test_virtual_padded_row_restatement():
    """Host model of the kernel's formulation (real points only + one virtual padded row + split W1) equals the oracle."""
    from helpers import load_small
    z, geom, frames, overflow, wseed = load_small("tiny_continue")
    w = hybrid.random_weights(wseed)
    vox = torch.from_numpy(z["voxels"]); n = torch.from_numpy(z["voxel_num_points"]); c = torch.from_numpy(z["voxel_coords"])
    ref, _, _ = hybrid.pillar_vfe(vox, n, c, w, list(geom.voxel_size), geom.range_f32)
    from hvpr_b200 import vfe
    m = vfe.PillarVFE_Scale(config.HVPR_VFE_CFG, 4, list(geom.voxel_size), geom.range_f32).eval()
    m.load_state_dict({k[4:]: v for k, v in w.items() if k.startswith("vfe.")}, strict=False)
    W = m._weights()
    w0 = torch.tensor(list(W.w0)).view(16, 10); b0 = torch.tensor(list(W.b0))
    w1a = torch.tensor(list(W.w1a)).view(64, 16); w1b = torch.tensor(list(W.w1b)).view(64, 16); b1 = torch.tensor(list(W.b1))
    out = torch.empty_like(ref)
    for p in range(vox.shape[0]):
        k = int(n[p]); pts = vox[p, :k]
        mean = vox[p, :, :3].sum(0) / k
        ctr = torch.tensor([c[p, 3] * geom.voxel_size[0] + m.x_offset, c[p, 2] * geom.voxel_size[1] + m.y_offset,
                            c[p, 1] * geom.voxel_size[2] + m.z_offset], dtype=torch.float32)
        f = torch.cat([pts, pts[:, :3] - mean, pts[:, :3] - ctr], 1)
        x0 = torch.relu(f @ w0.t() + b0)
        rb0 = torch.relu(b0)
        xmax = x0.max(0)[0] if k == 32 else torch.maximum(x0.max(0)[0], rb0)
        c1 = w1b @ xmax + b1
        y = torch.relu(x0 @ w1a.t() + c1).max(0)[0]
        if k < 32:
            y = torch.maximum(y, torch.relu(w1a @ rb0 + c1))
        out[p] = y
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)


def test_cpu_tensors_fail_loudly():
    from hvpr_b200 import map_to_bev, vfe
    m = vfe.PillarVFE_Scale(config.HVPR_VFE_CFG, 4, list(G1.voxel_size), G1.range_f32).eval()
    bd = dict(voxels=torch.zeros(3, 32, 4), voxel_num_points=torch.ones(3), voxel_coords=torch.zeros(3, 4))
    with pytest.raises(_lib.HvprError):
        m(bd)
    b = map_to_bev.PointPillarScatter(config.Cfg(NUM_BEV_FEATURES=64), grid_size=G1.grid_size)
    with pytest.raises(_lib.HvprError):
        b(dict(pillar_features=torch.zeros(3, 64), voxel_coords=torch.zeros(3, 4), batch_size=1))
    m.train()
    with pytest.raises(NotImplementedError):
        m(bd)


def test_yaml_cfg_loader(tmp_path):
    y = tmp_path / "m.yaml"
    y.write_text("DATA_CONFIG:\n  POINT_CLOUD_RANGE: [0, -19.84, -2.5, 47.36, 19.84, 0.5]\n  DATA_PROCESSOR:\n"
                 "    - NAME: transform_points_to_voxels\n      VOXEL_SIZE: [0.16, 0.16, 3]\n      MAX_POINTS_PER_VOXEL: 32\n"
                 "      MAX_NUMBER_OF_VOXELS: {train: 16000, test: 40000}\n"
                 "MODEL:\n  VFE: {NAME: PillarVFE_Scale, WITH_DISTANCE: false, USE_ABSLOTE_XYZ: true, USE_NORM: true,"
                 " NUM_FILTERS: [32, 64], NUM_SCALE_FEATURES: [16, 32]}\n"
                 "  MAP_TO_BEV: {NAME: PointPillarScatter_Agg_Memory_1_scale, NUM_BEV_FEATURES: 128, NUM_PT_FEATURES: 64,"
                 " NUM_SCALE_FEATURES: 32, NUM_COORD_POINTS: 3, NUM_K: 20, NUM_M: 2000, SHRINK_TH: 0.0025}\n")
    c = config.load_yaml_model_cfg(str(y))
    assert c["VFE"].NUM_FILTERS == [32, 64] and c["MAP_TO_BEV"].NUM_K == 20
    assert c["VOXELIZER"].MAX_NUMBER_OF_VOXELS["test"] == 40000


def test_synth_generators_are_deterministic():
    from hvpr_b200 import synth
    a = synth.make_frame("L", 5000, G2.point_cloud_range, 1024)
    b = synth.make_frame("L", 5000, G2.point_cloud_range, 1024)
    assert a.dtype == np.float32 and a.shape == (5000, 4) and np.array_equal(a, b)
    c = synth.collate_points([a, b])
    assert c.shape == (10000, 5) and c[4999, 0] == 0 and c[5000, 0] == 1


# ---- row N1: backbone host side and conv C ABI (no GPU needed) -------------------------------------------------------
def test_conv_args_struct_layout_and_validation_without_gpu():
    assert ctypes.sizeof(_lib.HvprConvArgs) == 128
    assert _lib.HvprConvArgs.w_packed.offset == 40 and _lib.HvprConvArgs.out.offset == 96 and _lib.HvprConvArgs.out_ctot.offset == 120
    L = _lib.lib()
    assert L.hvpr_conv_packed_bytes(128, 9, 128) == 128 * 9 * 128 * 2
    assert L.hvpr_conv2d(None, None) == -1
    a = _lib.HvprConvArgs()                                    # all-null arguments are rejected before any CUDA call
    assert L.hvpr_conv2d(ctypes.byref(a), None) == -1
    assert L.hvpr_conv_pack_weights(None, 128, 9, 128, 128, None, None) == -1
    buf = ctypes.create_string_buffer(64)
    assert L.hvpr_conv_pack_weights(buf, 128, 9, 100, 128, buf, None) == -2     # c_in must be a multiple of 64
    assert L.hvpr_conv_pack_weights(buf, 96, 9, 128, 128, buf, None) == -2      # n_total must be a multiple of bn
    assert L.hvpr_attention_gate(None, 1, 8, 8, 64, 32, None, 0.0, None, None, None) == -1
    assert L.hvpr_nchw_to_nhwc_bf16(None, 1, 32, 8, 8, None, 64, None) == -1
    assert L.hvpr_bev_fill_nhwc_bf16(None, 64, None, 64, None, 32, None, 1, 8, 8, None, 128, None, 64, None) == -1


def test_backbone_registry_state_dict_names_and_loud_failures():
    from hvpr_b200.backbone import BaseBEVBackbone_Scale
    from hvpr_b200.pipeline import HVPR_BACKBONE_CFG
    from oracle import backbone as ob
    m = BaseBEVBackbone_Scale(HVPR_BACKBONE_CFG, 128)
    assert m.num_bev_features == 384
    w = ob.random_backbone_weights(0)
    sd = m.state_dict()
    assert {k for k in sd if not k.endswith("num_batches_tracked")} == set(w)
    assert all(tuple(sd[k].shape) == w[k].shape for k in w)
    r = m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    assert not r.unexpected_keys
    with pytest.raises(_lib.HvprError):                        # CPU tensors: no fallback
        m.eval()({"spatial_features": torch.zeros(1, 128, 8, 8), "spatial_scale_features": torch.zeros(1, 32, 8, 8)})
    with pytest.raises(NotImplementedError):                   # training branch is out of scope
        m.train()({"spatial_features": torch.zeros(1, 128, 8, 8), "spatial_scale_features": torch.zeros(1, 32, 8, 8)})
    with pytest.raises(NotImplementedError):                   # unsupported layouts are refused at construction
        BaseBEVBackbone_Scale(config.Cfg(HVPR_BACKBONE_CFG, LAYER_STRIDES=[1, 3, 2]), 128)
    if ob is not None:                                         # same names as the reference's own module when the tree is present
        from oracle import ref_loader
        if ref_loader.available():
            ref = ref_loader.load_backbone().BaseBEVBackbone_Scale(ref_loader.BACKBONE_CFG, 128)
            assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}


def test_backbone_bn_folding_matches_oracle_layer():
    """float64 BN folding of one conv block reproduces conv -> BN(eval) of the oracle (CPU, fp32 tolerance)."""
    import torch.nn.functional as F
    from hvpr_b200.backbone import BaseBEVBackbone_Scale, _fold
    from hvpr_b200.pipeline import HVPR_BACKBONE_CFG
    from oracle import backbone as ob
    w = ob.random_backbone_weights(4)
    m = BaseBEVBackbone_Scale(HVPR_BACKBONE_CFG, 128).eval()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    cw, s, shift = _fold(m.blocks[1][1].weight, m.blocks[1][2])
    x = torch.randn(1, 128, 6, 8)
    got = F.conv2d(x.double(), cw * s[:, None, None, None], shift, stride=2, padding=1)
    ref = ob._conv_bn_relu(x, w, "blocks.1.1.weight", "blocks.1.2", 2, relu=False)
    assert float((got - ref.double()).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_dense_head_module_names_anchors_and_loud_failures():
    from hvpr_b200.dense_head import AnchorHeadSingle, build_anchors
    from oracle import dense_head as od
    rng = [0, -39.68, -3, 69.12, 39.68, 1]
    m = AnchorHeadSingle(config.Cfg(**od.HEAD_CFG), 384, 1, ["Car"], [48, 40, 1], rng)
    w = od.random_head_weights(0)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: v.shape for k, v in w.items()}
    assert m.num_anchors_per_location == 2 and m._layout() == (0, 2, 16, 32)
    assert torch.equal(build_anchors(od.HEAD_CFG["ANCHOR_GENERATOR_CONFIG"], [48, 40, 1], rng, "cpu"),
                       od.generate_anchors(od.HEAD_CFG, (48, 40, 1), rng))
    with pytest.raises(_lib.HvprError):
        m.eval()({"spatial_features_2d": torch.zeros(1, 384, 40, 48)})
    with pytest.raises(NotImplementedError):
        m.train()({"spatial_features_2d": torch.zeros(1, 384, 40, 48)})
    assert _lib.lib().hvpr_head_decode(None, 1, 1, 1, 32, 2, 1, 0, 2, 16, 2, None, 0.0, 0.0, None, None, None) == -1


def test_post_process_abi_and_oracle_geometry_without_gpu():
    L = _lib.lib()
    assert L.hvpr_post_process_workspace_bytes(0, 10) == 0
    n = L.hvpr_post_process_workspace_bytes(8, 428544)
    assert 8 * 428544 * 8 < n < 64 * 2 ** 20
    assert L.hvpr_post_process(None, None, 1, 10, 1, 0, 0.1, 4096, 500, 0.1, None, None, None, None, None, None, 0, None) == -1
    buf = ctypes.create_string_buffer(64)
    assert L.hvpr_post_process(buf, buf, 1, 10, 1, 0, 0.1, 8192, 500, 0.1, buf, buf, buf, buf, buf, buf, 64, None) == -2   # pre > 4096
    assert L.hvpr_post_process(buf, buf, 1, 10, 1, 0, 0.1, 4096, 500, 0.1, buf, buf, buf, buf, buf, buf, 64, None) == -3   # workspace
    from hvpr_b200.post_process import PostProcessor
    assert PostProcessor(config.Cfg(SCORE_THRESH=0.1, NMS_CONFIG=config.Cfg(MULTI_CLASSES_NMS=True, NMS_THRESH=0.1, NMS_PRE_MAXSIZE=4096,
                                                                            NMS_POST_MAXSIZE=500))).multi_classes      # round 2: built
    with pytest.raises(NotImplementedError):
        PostProcessor(config.Cfg(SCORE_THRESH=0.1, NMS_CONFIG=config.Cfg(MULTI_CLASSES_NMS=False, NMS_THRESH=0.1, NMS_PRE_MAXSIZE=8192,
                                                                         NMS_POST_MAXSIZE=500)))
    # 3-D IoU oracle: two unit-height boxes, half overlapping in x, same z -> 1/3; shifted by half their height -> (1/2 * 1/2) / (2 - 1/4)
    b1z, b2z = np.array([0, 0, 0, 4, 2, 1, 0.0]), np.array([2, 0, 0.5, 4, 2, 1, 0.0])
    from oracle import post_process as op3
    assert abs(op3.iou3d(b1z, b1z + np.array([2, 0, 0, 0, 0, 0, 0])) - 1.0 / 3.0) < 1e-9
    assert abs(op3.iou3d(b1z, b2z) - (4 * 0.5) / (16 - 2)) < 1e-9
    # oracle geometry: known answers of the rotated IoU
    from oracle import post_process as op
    b = np.array([0, 0, 0, 4, 2, 1, 0.3])
    assert abs(op.iou_bev(b, b) - 1.0) < 1e-12 and op.iou_bev(b, b + np.array([10, 0, 0, 0, 0, 0, 0])) == 0.0
    b1, b2 = np.array([0, 0, 0, 4, 2, 1, 0.0]), np.array([2, 0, 0, 4, 2, 1, 0.0])
    assert abs(op.iou_bev(b1, b2) - 1.0 / 3.0) < 1e-12
    sq, rot = np.array([0, 0, 0, 2, 2, 1, 0.0]), np.array([0, 0, 0, 2, 2, 1, np.pi / 4])
    assert abs(op.iou_bev(sq, rot) - (8 * (np.sqrt(2) - 1)) / (8 - 8 * (np.sqrt(2) - 1))) < 1e-9       # regular octagon of two unit squares


def test_detector_state_dict_uses_the_reference_module_names():
    """MixAnchor_Memory exposes vfe / map_to_bev_module / backbone_2d / dense_head (detector3d_template.py:30-33): the union of the
    per-module reference state dicts, each under its reference prefix, and nothing else."""
    if True:
        # the detector allocates its voxelizer on the device; the key layout is checked through the parts on CPU
        from hvpr_b200 import map_to_bev, vfe
        from hvpr_b200.backbone import BaseBEVBackbone_Scale
        from hvpr_b200.dense_head import AnchorHeadSingle
        from hvpr_b200.pipeline import HVPR_BACKBONE_CFG, HVPR_HEAD_CFG
        from oracle import backbone as ob, dense_head as od
        parts = {"vfe": vfe.PillarVFE_Scale(config.HVPR_VFE_CFG, 4, list(G2.voxel_size), G2.range_f32),
                 "map_to_bev_module": map_to_bev.PointPillarScatter_Agg_Memory_1_scale(config.HVPR_BEV_CFG, grid_size=G2.grid_size),
                 "backbone_2d": BaseBEVBackbone_Scale(HVPR_BACKBONE_CFG, 128),
                 "dense_head": AnchorHeadSingle(HVPR_HEAD_CFG, 384, 1, ["Car"], G2.grid_size, G2.point_cloud_range)}
        keys = {"%s.%s" % (p, k) for p, m in parts.items() for k in m.state_dict() if not k.endswith("num_batches_tracked")}
        want = (set(hybrid.random_weights(0)) | {"backbone_2d." + k for k in ob.random_backbone_weights(0)} |
                {"dense_head." + k for k in od.random_head_weights(0)})
        assert keys == want


# ---------------------------------------------------------------------------------------------- work-assignment maps
@pytest.mark.parametrize("max_vox,run", [(40000, 4), (80000, 4), (16000, 4), (37, 4), (1, 4), (130, 1), (4099, 4)])
def test_gather_slot_assignment_covers_every_slot_once(max_vox, run):
    """vox_gather_kernel (hvpr_b200/csrc/voxelize.cu): lane l of warp w takes slot (l // run * wpf + w) * run + l % run of its frame.
    The strided runs must tile [0, max_vox) exactly once for any cap, including caps that are not multiples of 32 or of the run."""
    runs = -(-max_vox // run)
    wpf = -(-runs // (32 // run))
    lane = np.arange(32)
    seen = np.zeros(max_vox, dtype=np.int32)
    for w in range(wpf):
        r = (lane // run) * wpf + w
        v = r * run + lane % run
        ok = (r < runs) & (v < max_vox)
        np.add.at(seen, v[ok], 1)
    assert (seen == 1).all()


@pytest.mark.parametrize("n_rows", [1, 31, 32, 33, 1000, 201833])
def test_pfn_group_assignment_covers_every_row_once(n_rows):
    """pfn_kernel (hvpr_b200/csrc/pfn.cu): slot pl of group grp is pillar row pl * ngroups + grp (strided, not consecutive)."""
    ngroups = -(-n_rows // 32)
    pl, grp = np.meshgrid(np.arange(32), np.arange(ngroups), indexing="ij")
    rows = (pl.astype(np.int64) * ngroups + grp).ravel()
    rows = rows[rows < n_rows]
    assert len(rows) == n_rows and len(np.unique(rows)) == n_rows


@pytest.mark.parametrize("n,tile", [(1, 2048), (2047, 2048), (2048, 2048), (120000, 2048), (300000, 2048), (5000, 64)])
def test_single_pass_assign_scan_model(n, tile):
    """vox_assign_kernel (hvpr_b200/csrc/voxelize.cu): every tile publishes (n_first, sum of the first points' cell counts) as ONE 64-bit
    word — count in the high half, ready bit 31 and n_first in the low half — and takes the sum of the words in front of it as its
    prefix.  NumPy restatement of that arithmetic against a plain exclusive scan (what oracle/voxelize.py::voxelize_np uses)."""
    rng = np.random.default_rng(n)
    is_first = rng.random(n) < 0.25
    cnt = np.where(is_first, rng.integers(1, 400, n), 0).astype(np.int64)
    ready = np.uint64(0x80000000)
    words = []
    for t0 in range(0, n, tile):
        a, b = int(is_first[t0:t0 + tile].sum()), int(cnt[t0:t0 + tile].sum())
        assert a < 2 ** 31 and b < 2 ** 32
        words.append((np.uint64(b) << np.uint64(32)) | ready | np.uint64(a))
    rank = np.empty(n, dtype=np.int64); off = np.empty(n, dtype=np.int64)
    for ti, t0 in enumerate(range(0, n, tile)):
        w = np.array(words[:ti], dtype=np.uint64)
        assert ((w & ready) != 0).all()
        pa = int((w & np.uint64(0x7FFFFFFF)).sum()); pb = int((w >> np.uint64(32)).sum())
        f, c = is_first[t0:t0 + tile], cnt[t0:t0 + tile]
        rank[t0:t0 + tile] = pa + np.cumsum(f) - f
        off[t0:t0 + tile] = pb + np.cumsum(c) - c
    assert (rank == np.cumsum(is_first) - is_first).all() and (off == np.cumsum(cnt) - cnt).all()


def test_bench_workloads_partition_frames_and_numa_binding_is_harmless():
    """bench.py host logic (no GPU): cfg3 splits its 64 frames frame-wise over the ranks (every frame exactly once, frame i -> rank
    i mod W), cfg2 / cfg4 / cfg5 give every rank 8 distinct frames; bind_numa never raises and leaves a usable affinity."""
    import argparse
    import bench
    for W in (1, 2, 4, 8):
        seen = []
        for r in range(W):
            geom, ids, pts, total = bench.workload(argparse.Namespace(config="cfg3"), r, W)
            assert total == 64 and pts == 120000 and geom.grid_size == (432, 496, 1)
            assert ids == list(range(r, 64, W))
            seen += ids
        assert sorted(seen) == list(range(64))
        for cfg, grid, npts in (("cfg2", (432, 496, 1), 120000), ("cfg4", (640, 640, 1), 300000), ("cfg5", (432, 496, 1), 120000)):
            all_ids = [bench.workload(argparse.Namespace(config=cfg), r, W)[1] for r in range(W)]
            assert all(len(i) == 8 for i in all_ids) and len({x for i in all_ids for x in i}) == 8 * W
            g, _, p, total = bench.workload(argparse.Namespace(config=cfg), 0, W)
            assert g.grid_size == grid and p == npts and total == 8 * W
    before = os.sched_getaffinity(0)
    info = bench.bind_numa(0, 8)
    assert "numa_nodes" in info and len(os.sched_getaffinity(0)) >= 1
    os.sched_setaffinity(0, before)
    a = bench.algorithmic_bytes(120000, 25229, 102786, 432, 496)
    assert a["voxelize"] == 16 * 120000 + 532 * 25229 and a["bev_fill"] == 4 * 160 * 432 * 496 + 4 * 432 * 496 + 640 * 25229
