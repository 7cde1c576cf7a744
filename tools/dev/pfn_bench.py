"""Dev probe — K2 (PFN) alone on the headline batch; optional library variant: python tools/dev/pfn_bench.py [variant]"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from hvpr_b200 import _lib
v = sys.argv[1] if len(sys.argv) > 1 else ""
if v: _lib.LIB_PATH = _lib.LIB_PATH.replace("libhvpr_b200.so", "libhvpr_b200_%s.so" % v)
from hvpr_b200 import synth
from hvpr_b200.geometry import G2
from hvpr_b200.frontend import HybridFrontEnd
fe = HybridFrontEnd(G2).load_reference_weights(synth.random_frontend_weights(0))
B, N = 8, 120000
frames = synth.make_batch("L", N, G2.point_cloud_range, B)
p = fe.plan(B, B * N, N, use_graph=False)
p.points.copy_(torch.from_numpy(np.concatenate(frames, 0))); p.frame_offsets.copy_(torch.tensor(np.r_[0, np.cumsum([N] * B)], dtype=torch.int32))
fe.run(); torch.cuda.synchronize()
vox = p.vox
def k2(): fe.vfe.run(vox.voxels, vox.num_points, vox.coords, vox.n_pillars_dev, out=p.pillar_features, scale_out=p.pillar_scale, launch=(BPS, lowreg))
res = {}
for BPS, lowreg in ((2, 0), (3, 0), (2, 1), (3, 1)):
    for _ in range(5): k2()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): k2()
    e1.record(); torch.cuda.synchronize()
    res["bps%d_lowreg%d" % (BPS, lowreg)] = round(e0.elapsed_time(e1) / 30, 4)
print("variant", v or "default", "K2 ms", res, "checksum", float(p.pillar_features.double().sum()))
