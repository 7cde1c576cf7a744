"""Dev probe — row N1 throughput: BaseBEVBackbone_Scale on the tcgen05 conv kernel, per layer and whole, next to the
same network evaluated by cuDNN (torch eager, bf16 channels_last) as the library baseline.
    python tools/dev/backbone_bench.py [--geom G2] [--batch 8] [--iters 5] [--no-cudnn]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hvpr_b200 import G1, G2, G3, _lib                      # noqa: E402
from hvpr_b200.backbone import BaseBEVBackbone_Scale        # noqa: E402
from hvpr_b200.config import Cfg                            # noqa: E402

PAIR_DEFAULT = int(os.environ.get("HVPR_CONV_PAIR", "0"))      # hvpr_dbg_conv_pair mode used for the whole-network timing
CFG = dict(LAYER_NUMS=[3, 3, 3], SFM_LAYER_NUMS=[3, 3, 3], LAYER_STRIDES=[1, 2, 2], NUM_FILTERS=[128, 256, 512],
           NUM_SCALE_FILTERS=[32, 64, 128], UPSAMPLE_STRIDES=[1, 2, 4], NUM_UPSAMPLE_FILTERS=[128, 128, 128])


def flops(m, B, H, W):
    tot, h, w, cin, cs = 0, H, W, 128, 32
    for i, nf in enumerate(m.num_filters):
        h, w = h // m.layer_strides[i], w // m.layer_strides[i]
        px = B * h * w
        tot += 2 * 9 * cin * nf * px + (m.layer_nums[i] + m.sfm_layer_nums[i]) * 2 * 9 * nf * nf * px
        tot += 2 * 9 * cs * m.num_scale_filters[i] * px
        tot += 2 * nf * m.num_upsample_filters[i] * m.upsample_strides[i] ** 2 * px
        cin, cs = nf, m.num_scale_filters[i]
    return tot


def cudnn_forward(m, x, y):
    """the same eval forward through torch's own layers (the parameter containers are real nn modules)"""
    ups = []
    a = m.attention.spatial
    for i in range(len(m.blocks)):
        x = m.blocks[i](x)
        y = m.scale_layers[i](y)
        pooled = torch.cat((y.max(1, keepdim=True)[0], y.mean(1, keepdim=True)), 1)
        gate = torch.sigmoid(a.norm(a.conv(pooled)))
        xa = x
        for _ in range(m.sfm_layer_nums[i]):
            xa = gate * m.sfmblocks_down[i](xa) + xa
        ups.append(m.deblocks[i](xa))
    return torch.cat(ups, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--geom", default="G2")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--no-cudnn", action="store_true")
    a = ap.parse_args()
    g = {"G1": G1, "G2": G2, "G3": G3}[a.geom]
    W, H = g.grid_size[0], g.grid_size[1]
    B = a.batch
    torch.manual_seed(0)
    m = BaseBEVBackbone_Scale(Cfg(NAME="BaseBEVBackbone_Scale", **CFG), 128).cuda().eval()
    for p in m.parameters():
        if p.dim() == 4:
            torch.nn.init.uniform_(p, -(3.0 / (p.shape[1] * p.shape[2] * p.shape[3])) ** 0.5, (3.0 / (p.shape[1] * p.shape[2] * p.shape[3])) ** 0.5)
    _lib.init_device()
    _lib.lib().hvpr_dbg_conv_pair(PAIR_DEFAULT)
    occ = (torch.rand(B, H, W, 1, device="cuda") < 0.12)
    x_in = (torch.randn(B, H, W, 128, device="cuda").abs() * occ).bfloat16()
    y_in = torch.zeros(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
    y_in[..., :32] = (torch.randn(B, H, W, 32, device="cuda").abs() * occ).bfloat16()
    with torch.no_grad():
        for _ in range(2):
            m.run_nhwc(x_in, y_in, B, H, W)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            m.run_nhwc(x_in, y_in, B, H, W)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
    fl = flops(m, B, H, W)
    res = {"geom": a.geom, "batch": B, "canvas": [H, W], "gflop_per_batch": fl / 1e9, "ms_per_batch": ms,
           "tflops": fl / ms / 1e9, "frames_per_sec": B / ms * 1e3}
    # per-layer: time each distinct conv shape alone
    layers = []
    P, pl = m._packed, m._plan(B, H, W, x_in.device)
    h, w, src = H, W, x_in
    with torch.no_grad():
        for i in range(3):
            lv = pl["lv"][i]
            for name, lay, s, hi, wi in (("first", P["blocks"][i][0], src, h, w), ("body", P["blocks"][i][1], lv["a"], lv["h"], lv["w"]),
                                         ("sfm", P["sfm"][i], lv["a"], lv["h"], lv["w"])):
                kw = dict(gate=lv["gate"], residual=s) if name == "sfm" else {}
                dst = lv["b"]
                ts = {}
                for mode in ("single", "halo", "pair", "pair_halo"):   # A/B inside one run: boxes and thermal state differ between runs
                    _lib.lib().hvpr_dbg_conv_halo_off(int("halo" not in mode))
                    _lib.lib().hvpr_dbg_conv_pair(2 if "pair" in mode else 1)
                    for _ in range(2):
                        m._conv(lay, s, B, hi, wi, dst, **kw)
                    e0.record()
                    for _ in range(a.iters):
                        m._conv(lay, s, B, hi, wi, dst, **kw)
                    e1.record()
                    torch.cuda.synchronize()
                    ts[mode] = e0.elapsed_time(e1) / a.iters
                _lib.lib().hvpr_dbg_conv_halo_off(1)
                _lib.lib().hvpr_dbg_conv_pair(PAIR_DEFAULT)
                t = ts["single"]
                f = 2 * 9 * lay.c_in * lay.n_total * B * lv["h"] * lv["w"]
                layers.append({"level": i, "layer": name, "cin": lay.c_in, "cout": lay.n_total, "stride": lay.stride,
                               "hw": [lv["h"], lv["w"]], "ms": round(t, 4), "tflops": round(f / t / 1e9, 1),
                               "ms_halo": round(ts["halo"], 4), "ms_pair": round(ts["pair"], 4), "ms_pair_halo": round(ts["pair_halo"], 4)})
            de = P["de"][i]
            tde = {}
            for mode in ("single", "pair"):
                _lib.lib().hvpr_dbg_conv_pair(2 if mode == "pair" else 1)
                for _ in range(2):
                    m._conv(de, lv["a"], B, lv["h"], lv["w"], pl["out"], out_mode=1, out_c_off=128 * i, out_ctot=384)
                e0.record()
                for _ in range(a.iters):
                    m._conv(de, lv["a"], B, lv["h"], lv["w"], pl["out"], out_mode=1, out_c_off=128 * i, out_ctot=384)
                e1.record()
                torch.cuda.synchronize()
                tde[mode] = e0.elapsed_time(e1) / a.iters
            _lib.lib().hvpr_dbg_conv_pair(PAIR_DEFAULT)
            for _ in range(2):
                m._conv(de, lv["a"], B, lv["h"], lv["w"], pl["out"], out_mode=1, out_c_off=128 * i, out_ctot=384)
            e0.record()
            for _ in range(a.iters):
                m._conv(de, lv["a"], B, lv["h"], lv["w"], pl["out"], out_mode=1, out_c_off=128 * i, out_ctot=384)
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / a.iters
            layers.append({"level": i, "layer": "deblock", "cin": de.c_in, "cout": de.n_total, "ms": round(t, 4),
                           "tflops": round(2 * de.c_in * de.n_total * B * lv["h"] * lv["w"] / t / 1e9, 1),
                           "out_gbps": round(B * 128 * H * W * 4 / t / 1e6, 1), "ms_single": round(tde["single"], 4), "ms_pair": round(tde["pair"], 4)})
            h, w, src = lv["h"], lv["w"], lv["a"]
    res["layers"] = layers
    if not a.no_cudnn:
        mb = m.to(memory_format=torch.channels_last).bfloat16()
        xc = x_in[..., :128].permute(0, 3, 1, 2)            # NCHW view over NHWC memory = channels_last
        yc = y_in[..., :32].permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        torch.backends.cudnn.benchmark = True
        with torch.no_grad():
            for _ in range(2):
                cudnn_forward(mb, xc, yc)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(a.iters):
                cudnn_forward(mb, xc, yc)
            e1.record()
            torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / a.iters
        res["cudnn_bf16_channels_last"] = {"ms_per_batch": t, "tflops": fl / t / 1e9, "note": "bf16 output, no fp32 NCHW concat"}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
