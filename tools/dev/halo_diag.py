"""Dev probe (written for the first, single-patch halo variant; the shipped halo path uses aligned column-shifted patches) —
what does tcgen05.mma read through a shifted halo descriptor?  Identity weights on ONE tap make the
output a copy of the A operand; inputs encode the pixel id (pass A) or the 16-byte chunk id (pass B)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from hvpr_b200 import _lib     # noqa: E402

_lib.init_device()
L = _lib.lib()
L.hvpr_dbg_conv_halo_off(0)          # the halo operand path is off by default
n, h, w, c = 1, 16, 8, 64


def run(x_nhwc, tap):
    wt = torch.zeros(c, 9, c, device="cuda")
    wt[:, tap, :] = torch.eye(c, device="cuda")
    wpk = torch.empty(L.hvpr_conv_packed_bytes(c, 9, c), dtype=torch.uint8, device="cuda")
    _lib.check(L.hvpr_conv_pack_weights(_lib.ptr(wt), c, 9, c, 64, _lib.ptr(wpk), _lib.cur_stream()))
    out = torch.zeros(n, h, w, c, dtype=torch.bfloat16, device="cuda")
    a = _lib.HvprConvArgs()
    a.in_, a.n, a.h_in, a.w_in, a.in_cs, a.c_in = x_nhwc.data_ptr(), n, h, w, c, c
    a.ksize, a.stride, a.w_packed, a.n_total, a.bn = 3, 1, wpk.data_ptr(), c, 64
    a.bias, a.relu, a.out_mode, a.out, a.out_cs = None, 0, 0, out.data_ptr(), c
    a.up, a.c_out = 1, c
    _lib.check(L.hvpr_conv2d(ctypes.byref(a), _lib.cur_stream()))
    torch.cuda.synchronize()
    return out.float()


pid = torch.arange(h * w, device="cuda").float().reshape(1, h, w, 1).expand(1, h, w, c).contiguous().bfloat16()
cid = (torch.arange(c, device="cuda") // 8).float().reshape(1, 1, 1, c).expand(1, h, w, c).contiguous().bfloat16() + 1
for mode in (0,):
  for tap in (4, 3, 5, 1, 7, 0, 8):
      dy, dx = tap // 3 - 1, tap % 3 - 1
      op = run(pid, tap)[0]          # (h, w, c): pixel id seen per output pixel and channel
      oc = run(cid, tap)[0]
      exp = torch.full((h, w), -1.0, device="cuda")
      ys, xs = torch.meshgrid(torch.arange(h, device="cuda"), torch.arange(w, device="cuda"), indexing="ij")
      sy, sx = ys + dy, xs + dx
      ok = (sy >= 0) & (sy < h) & (sx >= 0) & (sx < w)
      exp = torch.where(ok, (sy * w + sx).float(), torch.zeros_like(exp))
      good_p = (op[..., 0] == exp)
      print("tap", tap, "dy,dx", dy, dx, "pixel-id ok rows: %d/128" % int(good_p.sum()), " chunk-id ok: %d/%d" % (int(((oc[..., ::8] == torch.arange(1, 9, device='cuda').float()) | ~ok[..., None]).all(-1).sum()), h * w))
      if tap == 4:
          for y in range(0, 4):
              print("   y=%d pix seen (ch0):" % y, [int(v) for v in op[y, :, 0].tolist()], " exp:", [int(v) for v in exp[y].tolist()])
              print("        chunk ids seen at x=0..7 (ch 0,8,..56):", [[int(v) for v in oc[y, xx, ::8].tolist()] for xx in range(0, 8, 3)])
