"""Dev probe — cycle accounting of the conv kernels (needs `make -C hvpr_b200/csrc prof`): where does each role wait?"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hvpr_b200 import G2, _lib                      # noqa: E402
_lib.LIB_PATH = _lib.LIB_PATH.replace("libhvpr_b200.so", "libhvpr_b200_prof.so")
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
names = ["prod.wait_empty", "-", "-", "-", "mma.wait_tempty", "mma.wait_full", "mma.issue+commit", "mma.wait_hfull",
         "epi.wait_tfull", "epi.work", "-", "-"]
with torch.no_grad():
    m.run_nhwc(x_in, y_in, B, H, W)
    torch.cuda.synchronize()
    P, pl = m._packed, m._plan(B, H, W, x_in.device)
    lv0, lv1, lv2 = pl["lv"]
    cases = [("L0 body 128->128", P["blocks"][0][1], lv0["a"], H, W, lv0["b"]),
             ("L1 body 256->256", P["blocks"][1][1], lv1["a"], lv1["h"], lv1["w"], lv1["b"])]
    cases.append(("L2 deblock 512->2048 (up 4)", P["de"][2], lv2["a"], lv2["h"], lv2["w"], None))
    cases.append(("L0 deblock 128->128 (up 1)", P["de"][0], lv0["a"], H, W, None))
    for label, lay, src, h, w, dst in cases:
        for pair in (1, 2):
            L.hvpr_dbg_conv_pair(pair)
            prof = torch.zeros((148, 16), dtype=torch.int64, device="cuda")
            L.hvpr_dbg_conv_prof(_lib.ptr(prof))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if dst is None:
                m._conv(lay, src, B, h, w, pl["out"], out_mode=1, out_c_off=0, out_ctot=384)
            else:
                m._conv(lay, src, B, h, w, dst)
            e1.record()
            torch.cuda.synchronize()
            L.hvpr_dbg_conv_prof(None)
            ms = e0.elapsed_time(e1)
            pr = prof.cpu().numpy().astype(np.float64)
            rows = pr[::2] if pair == 2 else pr          # leader CTAs carry the MMA numbers
            print("%s  %s  %.3f ms (~%.0f kcycles @1.9GHz)" % (label, "PAIR" if pair == 2 else "single", ms, ms * 1.9e3))
            for i, n in enumerate(names):
                if n != "-":
                    print("    %-18s %9.1f kcycles (all CTAs mean %.1f)" % (n, rows[:, i].mean() / 1e3, pr[:, i].mean() / 1e3))
    L.hvpr_dbg_conv_pair(0)
