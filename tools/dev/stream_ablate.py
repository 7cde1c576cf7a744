"""Ablation of the streaming step: drop one stage from the captured graph and time the step."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from hvpr_b200 import synth
from hvpr_b200.geometry import G2
from hvpr_b200.frontend import HybridFrontEnd
from oracle import hybrid
w = hybrid.random_weights(0)
B, N = 8, 120000
frames = synth.make_batch("L", N, G2.point_cloud_range, B)
pts = torch.from_numpy(np.concatenate(frames, 0)).cuda(); off = torch.tensor(np.r_[0, np.cumsum([N] * B)], dtype=torch.int32).cuda()
def run(drop):
    fe = HybridFrontEnd(G2).load_reference_weights(w)
    sp = fe.plan_stream(B, B * N, N)
    for sl in range(len(sp.in_points)):
        sp.in_points[sl].copy_(pts); sp.in_offsets[sl].copy_(off)
    orig = {n: getattr(fe, n) for n in ("_stage_vox", "_stage_pfn", "_stage_bev")}
    # warm everything first with the real stages (stream_prime captures with whatever is bound at that time)
    state = {"capturing": False}
    def wrap(name):
        f = orig[name]
        def g(*a, **k):
            if name in drop and torch.cuda.is_current_stream_capturing():
                return None
            return f(*a, **k)
        return g
    for n in orig: setattr(fe, n, wrap(n))
    fe.stream_prime()
    for _ in range(10): fe.stream_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): fe.stream_step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 100
for drop in ((), ("_stage_vox",), ("_stage_pfn",), ("_stage_vox", "_stage_pfn")):
    print("dropped", drop or "nothing", "-> ms/step", round(run(set(drop)), 4))
