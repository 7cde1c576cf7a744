"""Probe: does a runtime memset / D2D copy overlap a kernel that occupies every SM (mem_attn_tc)?"""
import ctypes, sys, numpy as np, torch
sys.path.insert(0, '.')
from hvpr_b200 import _lib, synth
from hvpr_b200.geometry import G2
from hvpr_b200.frontend import HybridFrontEnd
from oracle import hybrid
_lib.init_device(); L = _lib.lib()
L.hvpr_dbg_memset_async.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
L.hvpr_dbg_memcpy_d2d_async.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
w = hybrid.random_weights(0)
fe = HybridFrontEnd(G2).load_reference_weights(w)
B, N = 8, 120000
frames = synth.make_batch("L", N, G2.point_cloud_range, B)
p = fe.plan(B, B * N, N, use_graph=False)
p.points.copy_(torch.from_numpy(np.concatenate(frames, 0))); p.frame_offsets.copy_(torch.tensor(np.r_[0, np.cumsum([N] * B)], dtype=torch.int32))
fe.run(); torch.cuda.synchronize()
canvas = p.spatial; nbytes = canvas.numel() * 4 + p.spatial_scale.numel() * 4
big = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
zeros = torch.zeros(64 << 20, dtype=torch.uint8, device="cuda")
side = torch.cuda.Stream()
def k3():
    fe.map_to_bev_module.memory.run(p.pillar_features, 20, p.vox.n_pillars_dev, out=p.readout)
def memset(stream):
    _lib.check(L.hvpr_dbg_memset_async(ctypes.c_void_p(big.data_ptr()), 0, nbytes, ctypes.c_void_p(stream.cuda_stream)))
def d2d(stream):
    off = 0
    while off < nbytes:
        n = min(zeros.numel(), nbytes - off)
        _lib.check(L.hvpr_dbg_memcpy_d2d_async(ctypes.c_void_p(big.data_ptr() + off), ctypes.c_void_p(zeros.data_ptr()), n, ctypes.c_void_p(stream.cuda_stream)))
        off += n
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
main = torch.cuda.current_stream()
print("K3 alone ms", timeit(k3))
print("memset alone ms (%.2f GB)" % (nbytes / 1e9), timeit(lambda: memset(main)))
print("d2d-from-64MB-zeros alone ms", timeit(lambda: d2d(main)))
def both(op):
    side.wait_stream(main)
    op(side)
    k3()
    main.wait_stream(side)
print("K3 || memset ms", timeit(lambda: both(memset)))
print("K3 || d2d ms", timeit(lambda: both(d2d)))
def both_rev(op):      # start K3 first
    k3()
    side.wait_stream(main) if False else None
    op(side)
    main.wait_stream(side)
print("K3 then-launched memset on side ms", timeit(lambda: both_rev(memset)))
