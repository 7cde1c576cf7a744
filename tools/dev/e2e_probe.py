"""Dev probe: where the gap between the device-resident streaming step and the end-to-end step (pinned host input, D2H of the
pillar offsets) comes from: no host I/O | H2D only | H2D + D2H."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from hvpr_b200 import _lib, synth
from hvpr_b200.geometry import G2
from hvpr_b200.frontend import HybridFrontEnd
from oracle import hybrid
_lib.init_device()
w = hybrid.random_weights(0)
B, N = 8, 120000
frames = synth.make_batch("L", N, G2.point_cloud_range, B)
host_pts = torch.from_numpy(np.concatenate(frames, 0)).pin_memory()
host_off = torch.tensor(np.r_[0, np.cumsum([N] * B)], dtype=torch.int32).pin_memory()
host_cnt = torch.zeros(B + 1, dtype=torch.int32).pin_memory()
fe = HybridFrontEnd(G2).load_reference_weights(w)
sp = fe.plan_stream(B, B * N, N)
for sl in range(3):
    sp.in_points[sl].copy_(host_pts); sp.in_offsets[sl].copy_(host_off)      # every slot holds a batch (the resident leg re-voxelizes them)
fe.stream_prime((host_pts, host_off), (host_pts, host_off))
def timed(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    fe.stream_wait_outputs()
    t_host = (time.perf_counter() - t0) / n * 1e3
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 4), round(t_host, 4)
print("resident      (gpu ms/step, host enqueue ms/step)", timed(lambda: fe.stream_step()))
print("H2D           ", timed(lambda: fe.stream_step(host_pts, host_off)))
print("H2D + D2H     ", timed(lambda: fe.stream_step(host_pts, host_off, host_cnt)))
