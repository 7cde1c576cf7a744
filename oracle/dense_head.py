"""ORACLE (test infrastructure) — CPU restatement of row N2: AnchorHeadSingle eval forward
(pcdet/models/dense_heads/anchor_head_single.py:109-145) = three 1x1 convolutions + generate_predicted_boxes
(anchor_head_template.py:293-340) with ResidualCoder.decode_torch (pcdet/utils/box_coder_utils.py:45-77), AnchorGenerator
(dense_heads/target_assigner/anchor_generator.py:17-60) and common_utils.limit_period (pcdet/utils/common_utils.py:20-23).

Pin status: the reference's head CLASS cannot be imported (package-relative imports pull the absent iou3d_nms / roiaware CUDA ops,
breakages B5/B6), so the assembly is restated from the source; its three numeric building blocks — decode_torch, generate_anchors
(with the hard-coded `.cuda()` calls patched out in memory) and limit_period — are the reference's OWN functions whenever
/root/reference is present (tests/test_oracle_cpu.py checks this restatement against them).  Never imported by hvpr_b200/.

B10: the shipped hvpr.yaml puts the anchors on a stride-2 map while the backbone emits full resolution; the reference
would fail in `.view(batch_size, num_anchors, -1)`.  Tests use feature_map_stride = 1, which is the only consistent reading.
"""
from __future__ import annotations

import math
import os
import types

import numpy as np
import torch
import torch.nn.functional as F

# hvpr.yaml:96-118
HEAD_CFG = dict(USE_DIRECTION_CLASSIFIER=True, DIR_OFFSET=0.78539, DIR_LIMIT_OFFSET=0.0, NUM_DIR_BINS=2,
                ANCHOR_GENERATOR_CONFIG=[dict(class_name="Car", anchor_sizes=[[3.9, 1.6, 1.56]], anchor_rotations=[0, 1.57],
                                              anchor_bottom_heights=[-1.78], align_center=False, feature_map_stride=1,
                                              matched_threshold=0.6, unmatched_threshold=0.45)])


def generate_anchors(cfg, grid_size, point_cloud_range):
    """anchor_generator.py:17-60 + anchor_head_template.py:40-47 for one class -> (ny, nx, A, 7) fp32 [x,y,z,dx,dy,dz,r]."""
    c = cfg["ANCHOR_GENERATOR_CONFIG"][0]
    gx, gy = int(grid_size[0]) // c["feature_map_stride"], int(grid_size[1]) // c["feature_map_stride"]
    r = point_cloud_range
    if c.get("align_center", False):
        xs, ys = (r[3] - r[0]) / gx, (r[4] - r[1]) / gy
        xo, yo = xs / 2, ys / 2
    else:
        xs, ys = (r[3] - r[0]) / (gx - 1), (r[4] - r[1]) / (gy - 1)
        xo, yo = 0, 0
    x = torch.arange(r[0] + xo, r[3] + 1e-5, step=xs, dtype=torch.float32)
    y = torch.arange(r[1] + yo, r[4] + 1e-5, step=ys, dtype=torch.float32)
    z = torch.tensor(c["anchor_bottom_heights"], dtype=torch.float32)
    sizes = torch.tensor(c["anchor_sizes"], dtype=torch.float32)
    rots = torch.tensor(c["anchor_rotations"], dtype=torch.float32)
    out = torch.zeros(len(y), len(x), len(z) * len(sizes) * len(rots), 7)
    a = 0
    for zi in range(len(z)):                       # meshgrid order [x, y, z] -> permute(2,1,0,...) -> view: (z, y, x, size, rot)
        assert len(z) == 1, "one bottom height (hvpr.yaml)"
        for si in range(len(sizes)):
            for ri in range(len(rots)):
                out[:, :, a, 0] = x[None, :]
                out[:, :, a, 1] = y[:, None]
                out[:, :, a, 2] = z[zi]
                out[:, :, a, 3:6] = sizes[si]
                out[:, :, a, 6] = rots[ri]
                a += 1
    out[..., 2] += out[..., 5] / 2                  # :56 shift to box centres
    return out


def decode(box_encodings, anchors):
    """ResidualCoder.decode_torch (box_coder_utils.py:45-77), code_size 7."""
    xa, ya, za, dxa, dya, dza, ra = torch.split(anchors, 1, dim=-1)
    xt, yt, zt, dxt, dyt, dzt, rt = torch.split(box_encodings, 1, dim=-1)
    diagonal = torch.sqrt(dxa ** 2 + dya ** 2)
    return torch.cat([xt * diagonal + xa, yt * diagonal + ya, zt * dza + za, torch.exp(dxt) * dxa, torch.exp(dyt) * dya,
                      torch.exp(dzt) * dza, rt + ra], dim=-1)


def limit_period(val, offset=0.5, period=math.pi):
    return val - torch.floor(val / period + offset) * period            # common_utils.py:20-23


def random_head_weights(seed: int, input_channels=384, num_class=1, A=2, bins=2):
    rng = np.random.default_rng(seed)
    s = math.sqrt(3.0 / input_channels)
    w = {"conv_cls.weight": rng.uniform(-s, s, (A * num_class, input_channels, 1, 1)).astype(np.float32),
         "conv_cls.bias": rng.uniform(-2.0, 0.5, A * num_class).astype(np.float32),
         "conv_box.weight": (0.3 * rng.uniform(-s, s, (A * 7, input_channels, 1, 1))).astype(np.float32),
         "conv_box.bias": rng.uniform(-0.1, 0.1, A * 7).astype(np.float32),
         "conv_dir_cls.weight": rng.uniform(-s, s, (A * bins, input_channels, 1, 1)).astype(np.float32),
         "conv_dir_cls.bias": rng.uniform(-0.3, 0.3, A * bins).astype(np.float32)}
    return w


def head_forward(w, spatial_features_2d, cfg, grid_size, point_cloud_range, return_raw=False):
    """anchor_head_single.py:109-145 (eval) -> batch_cls_preds (B, N, C), batch_box_preds (B, N, 7)."""
    x = torch.from_numpy(np.asarray(spatial_features_2d)).float()
    t = lambda k: torch.from_numpy(w[k])
    with torch.no_grad():
        cls = F.conv2d(x, t("conv_cls.weight"), t("conv_cls.bias")).permute(0, 2, 3, 1).contiguous()          # :112-118
        box = F.conv2d(x, t("conv_box.weight"), t("conv_box.bias")).permute(0, 2, 3, 1).contiguous()
        dirp = F.conv2d(x, t("conv_dir_cls.weight"), t("conv_dir_cls.bias")).permute(0, 2, 3, 1).contiguous() \
            if cfg.get("USE_DIRECTION_CLASSIFIER") else None
        B = x.shape[0]
        anchors = generate_anchors(cfg, grid_size, point_cloud_range)
        assert anchors.shape[0] == x.shape[2] and anchors.shape[1] == x.shape[3], "B10: anchor map != feature map"
        num_anchors = anchors.view(-1, 7).shape[0]                                                            # :312
        batch_anchors = anchors.view(1, -1, 7).repeat(B, 1, 1)
        batch_cls = cls.view(B, num_anchors, -1).float()
        batch_box = decode(box.view(B, num_anchors, -1), batch_anchors)                                       # :318
        if dirp is not None:                                                                                  # :320-332
            dir_offset, dir_limit_offset = cfg["DIR_OFFSET"], cfg["DIR_LIMIT_OFFSET"]
            dir_labels = torch.max(dirp.view(B, num_anchors, -1), dim=-1)[1]
            period = 2 * np.pi / cfg["NUM_DIR_BINS"]
            dir_rot = limit_period(batch_box[..., 6] - dir_offset, dir_limit_offset, period)
            batch_box[..., 6] = dir_rot + dir_offset + period * dir_labels.to(batch_box.dtype)
    if return_raw:
        return batch_cls.numpy(), batch_box.numpy(), (cls.numpy(), box.numpy(), dirp.numpy() if dirp is not None else None)
    return batch_cls.numpy(), batch_box.numpy()


# ---- the reference's own building blocks (only where /root/reference exists) -------------------------------------------------
def load_reference_pieces(ref_root=os.environ.get("HVPR_REFERENCE", "/root/reference")):
    ns = types.SimpleNamespace()
    src = open(os.path.join(ref_root, "pcdet/utils/box_coder_utils.py")).read()
    m = types.ModuleType("_hvpr_ref_box_coder")
    exec(compile(src, "box_coder_utils.py", "exec"), m.__dict__)
    ns.ResidualCoder = m.ResidualCoder
    src = open(os.path.join(ref_root, "pcdet/models/dense_heads/target_assigner/anchor_generator.py")).read()
    assert src.count(".cuda()") == 2
    src = src.replace(".cuda()", "")                    # in-memory patch: the generator hard-codes the device
    m = types.ModuleType("_hvpr_ref_anchor_generator")
    exec(compile(src, "anchor_generator.py[patched]", "exec"), m.__dict__)
    ns.AnchorGenerator = m.AnchorGenerator
    src = open(os.path.join(ref_root, "pcdet/utils/common_utils.py")).read()
    a = src.index("def check_numpy_to_torch")
    b = src.index("def drop_info_with_name")
    m = types.ModuleType("_hvpr_ref_common_utils")
    exec(compile("import numpy as np\nimport torch\n" + src[a:b], "common_utils.py[excerpt]", "exec"), m.__dict__)
    ns.limit_period = m.limit_period
    return ns
