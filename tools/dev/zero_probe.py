"""K3 with the canvas zero fill as a side job (bulk async stores from the producer thread) + K4 writing only occupied runs,
against K3 + write-everything K4.  Usage: python tools/dev/zero_probe.py [variant-lib-suffix]"""
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
for dist in ("L", "U"):
    frames = synth.make_batch(dist, N, G2.point_cloud_range, B)
    p = fe.plan(B, B * N, N, use_graph=False)
    p.points.copy_(torch.from_numpy(np.concatenate(frames, 0))); p.frame_offsets.copy_(torch.tensor(np.r_[0, np.cumsum([N] * B)], dtype=torch.int32))
    m = fe.map_to_bev_module
    m.fused_zero_fill = 0
    fe.run(); torch.cuda.synchronize()
    ref_sp, ref_sps = p.spatial.clone(), p.spatial_scale.clone()
    ro2 = torch.empty_like(p.readout)
    def k3(zf=False): m.memory.run(p.pillar_features, 20, p.vox.n_pillars_dev, out=ro2, zero_fill=[p.spatial, p.spatial_scale] if zf else None)
    def k4(cfg=None):
        _lib.check(_lib.lib().hvpr_bev_fill(_lib.ptr(p.pillar_features), 64, _lib.ptr(p.readout), 64, _lib.ptr(p.pillar_scale), 32,
                   _lib.ptr(p.vox.cell_map), B, m.nx, m.ny, _lib.ptr(p.spatial), _lib.ptr(p.spatial_scale), _lib.launch_cfg(cfg), _lib.cur_stream()))
    def timeit(fn, reps=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    print(dist, "pillars", int(p.vox.voxel_offsets[-1]), "| K3", round(timeit(k3), 4), "| K3+zero", round(timeit(lambda: k3(True)), 4),
          "| K4 full", round(timeit(k4), 4), "| K3 -> K4 full", round(timeit(lambda: (k3(), k4())), 4))
    for var in (1, 2, 3):
        for bps in (0, 4):
            t4 = timeit(lambda: k4((bps, var)))
            tt = timeit(lambda: (k3(True), k4((bps, var))))
            p.spatial.fill_(float("nan")); p.spatial_scale.fill_(float("nan"))
            k3(True); k4((bps, var)); torch.cuda.synchronize()
            ok = torch.equal(p.spatial, ref_sp) and torch.equal(p.spatial_scale, ref_sps)
            print("   variant", var, "bps", bps, "| K4 occupied-only", round(t4, 4), "| K3+zero -> K4", round(tt, 4), "| bits equal", ok)
