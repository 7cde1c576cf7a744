import sys, numpy as np, torch
sys.path.insert(0, '.')
from hvpr_b200 import _lib, synth
v = sys.argv[1] if len(sys.argv) > 1 else ''
if v: _lib.LIB_PATH = _lib.LIB_PATH.replace('libhvpr_b200.so', 'libhvpr_b200_%s.so' % v)
from hvpr_b200.geometry import G2
from hvpr_b200.frontend import HybridFrontEnd
from oracle import hybrid
w = hybrid.random_weights(0)
fe = HybridFrontEnd(G2).load_reference_weights(w)
B, N = 8, 120000
frames = synth.make_batch("L", N, G2.point_cloud_range, B)
p = fe.plan(B, B * N, N, use_graph=False)
p.points.copy_(torch.from_numpy(np.concatenate(frames, 0))); p.frame_offsets.copy_(torch.tensor(np.r_[0, np.cumsum([N] * B)], dtype=torch.int32))
fe.run(); torch.cuda.synchronize()
ro2 = torch.empty_like(p.readout)
m = fe.map_to_bev_module
def k3(): m.memory.run(p.pillar_features, 20, p.vox.n_pillars_dev, out=ro2)
def k4():
    _lib.check(_lib.lib().hvpr_bev_fill(_lib.ptr(p.pillar_features), 64, _lib.ptr(p.readout), 64, _lib.ptr(p.pillar_scale), 32,
               _lib.ptr(p.vox.cell_map), B, m.nx, m.ny, _lib.ptr(p.spatial), _lib.ptr(p.spatial_scale), _lib.launch_cfg(K4CFG), _lib.cur_stream()))
hi = torch.cuda.Stream(priority=-1); lo = torch.cuda.Stream()
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
main = torch.cuda.current_stream()
K4CFG = None
print("variant", v or "default", "K3 alone", round(timeit(k3), 4))
def both(first_k3=True):
    hi.wait_stream(main); lo.wait_stream(main)
    if first_k3:
        with torch.cuda.stream(hi): k3()
        with torch.cuda.stream(lo): k4()
    else:
        with torch.cuda.stream(lo): k4()
        with torch.cuda.stream(hi): k3()
    main.wait_stream(hi); main.wait_stream(lo)
for cfg in (None, (1, 0), (2, 0), (4, 0)):
    K4CFG = cfg
    print("K4 cfg", cfg, "alone", round(timeit(k4), 4), "K3||K4 K3 first", round(timeit(lambda: both(True)), 4), "K4 first", round(timeit(lambda: both(False)), 4))
