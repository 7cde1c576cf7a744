"""Dev probe — interleaved A/B of the conv kernel variants on three representative layers (median of rounds)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hvpr_b200 import G2, _lib                      # noqa: E402
from hvpr_b200.backbone import BaseBEVBackbone_Scale        # noqa: E402
from hvpr_b200.config import Cfg                            # noqa: E402
from tools.dev.backbone_bench import CFG                    # noqa: E402

B = 8
W, H = G2.grid_size[0], G2.grid_size[1]
m = BaseBEVBackbone_Scale(Cfg(NAME="BaseBEVBackbone_Scale", **CFG), 128).cuda().eval()
_lib.init_device()
L = _lib.lib()
x_in = torch.randn(B, H, W, 128, device="cuda").abs().bfloat16()
y_in = torch.zeros(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
MODES = {"single": (1, 1, 0), "pair": (2, 1, 0), "halo": (1, 0, 0), "pair_halo": (2, 0, 0)}
with torch.no_grad():
    m.run_nhwc(x_in, y_in, B, H, W)
    torch.cuda.synchronize()
    P, pl = m._packed, m._plan(B, H, W, x_in.device)
    lv0, lv1, lv2 = pl["lv"]
    cases = [("L0 body 128->128", P["blocks"][0][1], lv0["a"], H, W, lv0["b"]),
             ("L1 body 256->256", P["blocks"][1][1], lv1["a"], lv1["h"], lv1["w"], lv1["b"]),
             ("L2 body 512->512", P["blocks"][2][1], lv2["a"], lv2["h"], lv2["w"], lv2["b"]),
             ("L0 scale 32->32 (64-ch rows)", P["scale"][0], y_in, H, W, lv0["y"])]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for label, lay, src, h, w, dst in cases:
        res = {k: [] for k in MODES}
        for rnd in range(7):
            for k, (pair, halo_off, msub) in MODES.items():
                L.hvpr_dbg_conv_pair(pair); L.hvpr_dbg_conv_halo_off(halo_off); L.hvpr_dbg_conv_force_msub(msub)
                m._conv(lay, src, B, h, w, dst)
                e0.record()
                for _ in range(5):
                    m._conv(lay, src, B, h, w, dst)
                e1.record()
                torch.cuda.synchronize()
                res[k].append(e0.elapsed_time(e1) / 5)
        print(label, {k: round(float(np.median(v)), 4) for k, v in res.items()})
    L.hvpr_dbg_conv_pair(0); L.hvpr_dbg_conv_halo_off(1); L.hvpr_dbg_conv_force_msub(0)
