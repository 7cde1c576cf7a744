import ctypes, sys, numpy as np, torch
sys.path.insert(0, '.')
from hvpr_b200 import _lib, synth
v = sys.argv[1] if len(sys.argv) > 1 else ''
if v: _lib.LIB_PATH = _lib.LIB_PATH.replace('libhvpr_b200.so', 'libhvpr_b200_%s.so' % v)
from hvpr_b200.geometry import G2
from hvpr_b200.frontend import HybridFrontEnd
from oracle import hybrid
_lib.init_device(); L = _lib.lib()
w = hybrid.random_weights(0)
B, N = 8, 120000
frames = synth.make_batch("L", N, G2.point_cloud_range, B)
pts = torch.from_numpy(np.concatenate(frames, 0)).cuda(); off = torch.tensor(np.r_[0, np.cumsum([N] * B)], dtype=torch.int32).cuda()
import itertools
for ((bps, low), bev), order in itertools.product([((b, l), bev) for l in (0, 1) for b in (2, 3) for bev in (0, 2)], ("fork",)):
    fe = HybridFrontEnd(G2).load_reference_weights(w)
    fe.stream_pfn_knob = (bps, low); fe.stream_bev_knob = (bev, 0) if bev else None; fe.stream_k1_order = order
    sp = fe.plan_stream(B, B * N, N)
    for sl in range(3):
        sp.in_points[sl].copy_(pts); sp.in_offsets[sl].copy_(off)
    fe.stream_prime()
    for _ in range(10): fe.stream_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): fe.stream_step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 100
    print("pfn knob", (bps, low), "bev knob", bev, "K1", order, "stream ms/step", round(ms, 4), "fps", round(B / ms * 1e3))
