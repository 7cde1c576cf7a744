"""Dev probe — level-0 and level-1 body convs once per operand path (halo vs per-tap), for ncu."""
import os
import sys

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
x_in = torch.randn(B, H, W, 128, device="cuda").abs().bfloat16()
y_in = torch.zeros(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
with torch.no_grad():
    m.run_nhwc(x_in, y_in, B, H, W)
    torch.cuda.synchronize()
    P, pl = m._packed, m._plan(B, H, W, x_in.device)
    lv0, lv1, lv2 = pl["lv"]
    torch.cuda.nvtx.range_push("probe")
    for off in (0, 1):
        _lib.lib().hvpr_dbg_conv_halo_off(off)
        m._conv(P["blocks"][0][1], lv0["a"], B, H, W, lv0["b"])
        m._conv(P["blocks"][1][1], lv1["a"], B, lv1["h"], lv1["w"], lv1["b"])
    torch.cuda.nvtx.range_pop()
    torch.cuda.synchronize()
print("done")
