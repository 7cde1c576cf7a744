import ctypes, sys, torch
sys.path.insert(0, '.')
from hvpr_b200 import _lib
variant = sys.argv[1] if len(sys.argv) > 1 else ""
if variant:
    _lib.LIB_PATH = _lib.LIB_PATH.replace("libhvpr_b200.so", "libhvpr_b200_bev%s.so" % variant)
_lib.init_device(); L = _lib.lib()
B, nx, ny, P = 8, 432, 496, 25229 * 8
g = torch.Generator(device="cuda").manual_seed(0)
cm = torch.full((B, nx * ny), -1, dtype=torch.int32, device="cuda")
for b in range(B):
    idx = torch.randperm(nx * ny, device="cuda", generator=g)[:P // B]
    cm[b, idx] = torch.arange(b * (P // B), (b + 1) * (P // B), dtype=torch.int32, device="cuda")
fa, fb, fs = torch.randn(P, 64, device="cuda"), torch.randn(P, 64, device="cuda"), torch.randn(P, 32, device="cuda")
sp = torch.empty((B, 128, ny, nx), device="cuda"); sps = torch.empty((B, 32, ny, nx), device="cuda")
import os
if os.environ.get('EMPTY'): cm.fill_(-1)          # EMPTY=1: zero-fill only (no occupied cells), isolates the store stream from the gathers
def run():
    _lib.check(L.hvpr_bev_fill(_lib.ptr(fa), 64, _lib.ptr(fb), 64, _lib.ptr(fs), 32, _lib.ptr(cm), B, nx, ny, _lib.ptr(sp), _lib.ptr(sps), _lib.launch_cfg((knob, 0)) if knob else None, _lib.cur_stream()))
nbytes = 4 * 160 * nx * ny * B + 4 * nx * ny * B + 640 * P
ref = None
for knob in (0, 16, 8, 4, 3, 2, 1):
    for _ in range(5): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    chk = (float(sp.double().sum()), float(sps.double().sum()))
    if ref is None: ref = chk
    print("variant", variant, "blocks/SM", knob, "ms", round(ms, 4), "GB/s", round(nbytes / ms / 1e6, 1), "same" if chk == ref else "DIFFERENT")
