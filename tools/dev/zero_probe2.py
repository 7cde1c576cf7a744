import sys, numpy as np, torch
sys.path.insert(0, '.')
from hvpr_b200 import _lib, synth
v = sys.argv[1] if len(sys.argv) > 1 else ''
if v: _lib.LIB_PATH = _lib.LIB_PATH.replace('libhvpr_b200.so', 'libhvpr_b200_%s.so' % v)
from hvpr_b200.geometry import G2
from hvpr_b200.frontend import HybridFrontEnd
from oracle import hybrid
fe = HybridFrontEnd(G2).load_reference_weights(hybrid.random_weights(0))
B, N = 8, 120000
frames = synth.make_batch("L", N, G2.point_cloud_range, B)
p = fe.plan(B, B * N, N, use_graph=False)
p.points.copy_(torch.from_numpy(np.concatenate(frames, 0))); p.frame_offsets.copy_(torch.tensor(np.r_[0, np.cumsum([N] * B)], dtype=torch.int32))
m = fe.map_to_bev_module
fe.run(); torch.cuda.synchronize()
ro2 = torch.empty_like(p.readout)
flat = p.spatial.view(-1)
def k3(nbytes=0): m.memory.run(p.pillar_features, 20, p.vox.n_pillars_dev, out=ro2, zero_fill=[flat[:nbytes // 4]] if nbytes else None)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print(v or "default", "K3", round(timeit(k3), 4), " ".join("| %d MB: %.4f" % (mb, timeit(lambda: k3(mb << 20))) for mb in (256, 512, 836)))
