"""GPU parity suite (-m gpu): every CUDA kernel, called through the C ABI (ctypes), against the CPU oracle on the same
seeded inputs, against the committed golden fixtures (made by the reference's own modules), and — at BASELINE's full
sizes — through size-independent properties.  Bars (BASELINE.json north_star):
    voxel coords / counts / point->voxel assignment (and the bit-copied payload): bit-exact
    BEV fill given its inputs: bitwise
    pillar / BEV features: <= 1e-4 relative (fp32 path), <= 1e-2 (bf16 memory-attention variant)
Nothing here reads /root/reference.
"""
import json
import os

import numpy as np
import pytest
import torch

from hvpr_b200 import synth
from hvpr_b200.geometry import G1, G2, G3, Geometry
from oracle import hybrid
from oracle import voxelize as ov

from helpers import GOLDEN, TOL_BF16, TOL_FP32, load_small, rel_err, sha, tie_aware_readout_check, to_dev

pytestmark = pytest.mark.gpu


def _frontend(geom, w, overflow="continue", mem_precision="bf16_rescore"):
    from hvpr_b200.frontend import HybridFrontEnd
    fe = HybridFrontEnd(geom, overflow=overflow, mem_precision=mem_precision)
    fe.load_reference_weights(w)
    return fe


def _voxelize_gpu(frames, geom, overflow):
    from hvpr_b200.voxelizer import Voxelizer
    vz = Voxelizer(geom, overflow)
    pts, off = to_dev(frames)
    out = vz.run(pts, off, len(frames), max(len(f) for f in frames) if frames else 0)
    torch.cuda.synchronize()
    vo = out.voxel_offsets.cpu().numpy()
    P = int(vo[-1])
    return out, vo, out.voxels[:P].cpu().numpy(), out.coords[:P].cpu().numpy(), out.num_points[:P].cpu().numpy()


def _assert_vox_equal(frames, geom, overflow):
    out, vo, v, c, n = _voxelize_gpu(frames, geom, overflow)
    rv, rc, rn = ov.voxelize_batch(frames, geom.range_f32, geom.voxel_f32, geom.max_points_per_voxel,
                                   geom.max_voxels, overflow)
    assert int(vo[-1]) == len(rn)
    assert np.array_equal(n, rn)
    assert np.array_equal(c, rc)
    assert np.array_equal(v.view(np.int32), rv.view(np.int32))       # payload is a bit copy, padding is zero
    # cell map: exactly the kept cells, pointing at their rows
    nx, ny, nz = geom.grid_size
    cm = out.cell_map.cpu().numpy()
    exp = np.full_like(cm, -1)
    exp[rc[:, 0], (rc[:, 1] * ny + rc[:, 2]) * nx + rc[:, 3]] = np.arange(len(rn), dtype=np.int32)
    assert np.array_equal(cm, exp)
    return v, c, n


# ------------------------------------------------------------------------------------------------ K1 voxelize
@pytest.mark.parametrize("name", ["tiny_continue", "tiny_break_cap", "tiny_t5"])
def test_voxelize_golden_small(name):
    z, geom, frames, overflow, _ = load_small(name)
    out, vo, v, c, n = _voxelize_gpu(frames, geom, overflow)
    assert np.array_equal(v.view(np.int32), z["voxels"].view(np.int32))
    assert np.array_equal(c, z["voxel_coords"]) and np.array_equal(n, z["voxel_num_points"])


@pytest.mark.parametrize("mode", ["continue", "break"])
@pytest.mark.parametrize("gname,dist,n,B", [("G1", "U", 120000, 1), ("G1", "L", 120000, 2), ("G2", "U", 120000, 2),
                                            ("G2", "L", 120000, 3), ("G2", "L", 16384, 4)])
def test_voxelize_vs_oracle_full_size(gname, dist, n, B, mode):
    g = {"G1": G1, "G2": G2}[gname]
    frames = synth.make_batch(dist, n, g.point_cloud_range, B, first_frame=3, edge_cases=True)
    _assert_vox_equal(frames, g, mode)


def test_voxelize_golden_hashes():
    with open(os.path.join(GOLDEN, "voxel_hashes.json")) as fh:
        gold = json.load(fh)
    G = {"G1": G1, "G2": G2, "G3": G3}
    for key, ref in gold.items():
        gname, dist, n, mode = key.split("/")
        g = G[gname]
        f = synth.make_frame(dist, int(n), g.point_cloud_range, 1024, edge_cases=True)
        out, vo, v, c, k = _voxelize_gpu([f], g, mode)
        assert (len(k), int(k.sum())) == (ref["P"], ref["K"]), key
        assert sha(v.view(np.int32), c[:, 1:].copy(), k) == ref["sha256"], key


def test_voxelize_vs_reference_loop_pins():
    """K1's cell map and per-pillar counts against fixtures minted by the REFERENCE'S OWN voxel loop (tools/vis.py:8-60 run
    under numba by oracle/make_golden_vis.py; tests/golden/vis_kernel_pins.json): `break` mode, caps 5 000 / 40 000 / 80 000."""
    from oracle.make_golden_vis import GEOM, case_frame, digest
    with open(os.path.join(GOLDEN, "vis_kernel_pins.json")) as fh:
        pins = json.load(fh)
    for key, ref in pins.items():
        gname, dist, n, mv = key.split("/")
        g0 = GEOM[gname]
        g = Geometry(g0.point_cloud_range, g0.voxel_size, 32, int(mv))
        f = case_frame(gname, dist, int(n))
        out, vo, v, c, k = _voxelize_gpu([f], g, "break")
        assert len(k) == ref["P"], key
        cm = out.cell_map.cpu().numpy().reshape(-1)
        assert digest(cm) == ref["table_sha256"], key
        nx, ny, nz = g.grid_size
        counts = np.zeros(ny * nx, dtype=np.int32)
        counts[c[:, 2] * nx + c[:, 3]] = k
        assert digest(counts) == ref["counts_cap32_sha256"], key


def test_voxelize_ragged_and_empty_frames():
    g = G2
    frames = [synth.make_frame("L", 5000, g.point_cloud_range, 1), np.zeros((0, 4), np.float32),
              synth.make_frame("U", 777, g.point_cloud_range, 2), synth.make_frame("L", 1, g.point_cloud_range, 3)]
    _assert_vox_equal(frames, g, "continue")
    # all points outside the range -> zero pillars
    far = np.full((1000, 4), 1e6, np.float32)
    out, vo, v, c, n = _voxelize_gpu([far], g, "continue")
    assert int(vo[-1]) == 0


def test_voxelize_all_points_one_pillar_and_tiny_caps():
    g = Geometry(G2.point_cloud_range, G2.voxel_size, 32, 40000)
    p = synth.make_frame("U", 20000, (10.09, 0.97, -1.0, 10.23, 1.11, 0.0), 5)     # inside one 0.16 m cell
    v, c, n = _assert_vox_equal([p], g, "continue")
    assert len(n) == 1 and n[0] == 32 and np.array_equal(v[0], p[:32])              # first 32 in arrival order
    # max_voxels = 1, max_points = 1
    g1 = Geometry(G2.point_cloud_range, G2.voxel_size, 1, 1)
    f = synth.make_frame("L", 3000, g1.point_cloud_range, 9)
    _assert_vox_equal([f], g1, "continue")
    _assert_vox_equal([f], g1, "break")


def test_voxelize_truncates_frames_longer_than_the_stated_bound():
    """max_frame_points is a promise of the caller; a frame that breaks it is cut to its first max_frame_points points
    by every kernel alike (it used to lose its last scan tile and come out with ZERO pillars)."""
    from hvpr_b200.voxelizer import Voxelizer
    g = G2
    frames = [synth.make_frame("L", 9000, g.point_cloud_range, 31), synth.make_frame("U", 3000, g.point_cloud_range, 32),
              synth.make_frame("L", 5000, g.point_cloud_range, 33)]
    cap = 4100
    for mode in ("continue", "break"):
        vz = Voxelizer(g, mode)
        pts, off = to_dev(frames)
        out = vz.run(pts, off, len(frames), cap)
        torch.cuda.synchronize()
        P = int(out.voxel_offsets[-1])
        rv, rc, rn = ov.voxelize_batch([f[:cap] for f in frames], g.range_f32, g.voxel_f32, 32, g.max_voxels, mode)
        assert P == len(rn) and P > 0
        assert np.array_equal(out.num_points[:P].cpu().numpy(), rn) and np.array_equal(out.coords[:P].cpu().numpy(), rc)
        assert np.array_equal(out.voxels[:P].cpu().numpy().view(np.int32), rv.view(np.int32))


def _voxelize_hash(frames, geom, overflow, table):
    from hvpr_b200.voxelizer import Voxelizer
    vz = Voxelizer(geom, overflow, table=table)
    pts, off = to_dev(frames)
    out = vz.run(pts, off, len(frames), max(len(f) for f in frames))
    torch.cuda.synchronize()
    P = int(out.voxel_offsets[-1])
    return out, out.voxels[:P].cpu().numpy(), out.coords[:P].cpu().numpy(), out.num_points[:P].cpu().numpy()


@pytest.mark.parametrize("mode", ["continue", "break"])
def test_voxelize_open_addressing_table_equals_dense_table(mode):
    """The open-addressing hash table (64-bit cell keys, linear probing; used when the dense table does not fit) gives the same bits
    as the dense table and the oracle — pillar grids with the cap hit (U) and not hit (L), ragged / empty frames."""
    for g, frames in ((G2, synth.make_batch("U", 120000, G2.point_cloud_range, 2, first_frame=5, edge_cases=True)),
                      (G1, synth.make_batch("L", 60000, G1.point_cloud_range, 3, first_frame=8, edge_cases=True) + [np.zeros((0, 4), np.float32)])):
        out_h, v, c, n = _voxelize_hash(frames, g, mode, "hash")
        assert out_h.cell_map is None
        rv, rc, rn = ov.voxelize_batch(frames, g.range_f32, g.voxel_f32, 32, g.max_voxels, mode)
        assert np.array_equal(n, rn) and np.array_equal(c, rc) and np.array_equal(v.view(np.int32), rv.view(np.int32))
        _, v2, c2, n2 = _voxelize_hash(frames, g, mode, "auto")
        assert np.array_equal(v.view(np.int32), v2.view(np.int32)) and np.array_equal(c, c2) and np.array_equal(n, n2)


def test_voxelize_3d_grid_and_grid_beyond_2_pow_31_cells():
    """nz > 1: a SECOND-style 3-D voxel grid (dense table and hash table agree with the oracle), and a grid of 2.7e9 cells that only
    the open-addressing table can hold — checked against the dict-based oracle (no dense table anywhere)."""
    g3 = Geometry((0.0, -39.68, -3.0, 69.12, 39.68, 1.0), (0.16, 0.16, 0.1), 5, 16000)            # 432 x 496 x 40
    assert g3.grid_size == (432, 496, 40)
    frames = synth.make_batch("L", 40000, g3.point_cloud_range, 2, first_frame=3, edge_cases=True)
    rv, rc, rn = ov.voxelize_batch(frames, g3.range_f32, g3.voxel_f32, 5, 16000, "continue")
    assert len(set(rc[:, 1].tolist())) > 5                                                         # several z layers are occupied
    for table in ("auto", "hash"):
        _, v, c, n = _voxelize_hash(frames, g3, "continue", table)
        assert np.array_equal(n, rn) and np.array_equal(c, rc) and np.array_equal(v.view(np.int32), rv.view(np.int32)), table
    huge = Geometry((0.0, -39.68, -3.0, 69.12, 39.68, 1.0), (0.02, 0.02, 0.02), 8, 4000)           # 3456 x 3968 x 200 = 2.74e9 cells
    nx, ny, nz = huge.grid_size
    assert nx * ny * nz > 2 ** 31
    f = synth.make_frame("L", 6000, huge.point_cloud_range, 77, edge_cases=True)
    f[100:140] = f[100]                                                                            # a crowded cell (cap 8)
    from hvpr_b200.voxelizer import Voxelizer
    assert Voxelizer(huge, "continue").uses_hash_table(1)
    _, v, c, n = _voxelize_hash([f], huge, "continue", "auto")
    pv, pc, pn = ov.voxelize_py(f, huge.range_f32, huge.voxel_f32, 8, 4000, "continue")
    assert np.array_equal(n, pn) and np.array_equal(c[:, 1:], pc) and np.array_equal(v.view(np.int32), pv.view(np.int32))
    assert int(n.max()) == 8 and len(n) == 4000


def test_voxelize_collated_points_with_batch_column():
    """batch_dict path: (sum N, 5) [b,x,y,z,r] exactly as collate_batch pads it (dataset.py:161-166)."""
    from hvpr_b200.voxelizer import Voxelizer
    g = G1
    frames = synth.make_batch("L", 20000, g.point_cloud_range, 3, edge_cases=True)
    bd = dict(points=torch.from_numpy(synth.collate_points(frames)).cuda(), batch_size=3)
    bd = Voxelizer(g).voxelize_batch(bd)
    rv, rc, rn = ov.voxelize_batch(frames, g.range_f32, g.voxel_f32, 32, g.max_voxels)
    assert np.array_equal(bd["voxel_coords"].cpu().numpy(), rc)
    assert np.array_equal(bd["voxel_num_points"].cpu().numpy(), rn)
    assert np.array_equal(bd["voxels"].cpu().numpy().view(np.int32), rv.view(np.int32))


def test_voxelgenerator_dropin():
    """spconv-style ctor + generate(), tuple and dict flavours (data_processor.py:50-67)."""
    from hvpr_b200.voxelizer import VoxelGenerator, VoxelGeneratorV2
    g = G1
    f = synth.make_frame("L", 16384, g.point_cloud_range, 4)
    vg = VoxelGenerator(voxel_size=[0.16, 0.16, 3], point_cloud_range=g.range_f32, max_num_points=32, max_voxels=40000)
    v, c, n = vg.generate(f)
    rv, rc, rn = ov.voxelize_c(f, g.range_f32, g.voxel_f32, 32, 40000)
    assert np.array_equal(v.view(np.int32), rv.view(np.int32)) and np.array_equal(c, rc) and np.array_equal(n, rn)
    assert tuple(vg.grid_size) == g.grid_size
    d = VoxelGeneratorV2(voxel_size=[0.16, 0.16, 3], point_cloud_range=g.range_f32, max_num_points=32,
                         max_voxels=40000).generate(f)
    assert np.array_equal(d["coordinates"], rc) and np.array_equal(d["num_points_per_voxel"], rn)


def test_voxelize_is_deterministic():
    g = G2
    frames = synth.make_batch("U", 120000, g.point_cloud_range, 2)
    a = _voxelize_gpu(frames, g, "continue")
    b = _voxelize_gpu(frames, g, "continue")
    for x, y in zip(a[2:], b[2:]):
        assert np.array_equal(x.view(np.int32) if x.dtype == np.float32 else x, y.view(np.int32) if y.dtype == np.float32 else y)


# ------------------------------------------------------------------------------------------------ K2 PFN
def _pfn_case(geom, frames, wseed, scale=True):
    from hvpr_b200 import config, vfe
    w = hybrid.random_weights(wseed)
    rv, rc, rn = ov.voxelize_batch(frames, geom.range_f32, geom.voxel_f32, geom.max_points_per_voxel, geom.max_voxels)
    tv, tc, tn = torch.from_numpy(rv), torch.from_numpy(rc), torch.from_numpy(rn)
    with torch.no_grad():
        ref = hybrid.pillar_vfe(tv, tn, tc, w, list(geom.voxel_size), geom.range_f32, scale=scale)
    cls = vfe.PillarVFE_Scale if scale else vfe.PillarVFE
    m = cls(config.HVPR_VFE_CFG, 4, list(geom.voxel_size), geom.range_f32).cuda().eval()
    m.load_state_dict({k[4:]: v for k, v in w.items() if k.startswith("vfe.") and (scale or "pfn_layers" in k)}, strict=False)
    # the reference feeds fp32 coords / counts (E4); the module must accept them
    bd = dict(voxels=tv.cuda(), voxel_num_points=tn.float().cuda(), voxel_coords=tc.float().cuda())
    with torch.no_grad():
        bd = m(bd)
    torch.cuda.synchronize()
    return ref, bd


@pytest.mark.parametrize("gname,dist,n", [("G1", "L", 30000), ("G2", "L", 120000), ("G2", "U", 120000)])
def test_pfn_vs_oracle(gname, dist, n):
    g = {"G1": G1, "G2": G2}[gname]
    frames = synth.make_batch(dist, n, g.point_cloud_range, 2, first_frame=11, edge_cases=True)
    (rf, rs, rm), bd = _pfn_case(g, frames, 21)
    e1, e2 = rel_err(bd["pillar_features"], rf)
    assert e1 <= TOL_FP32 and e2 <= TOL_FP32, (e1, e2)
    e1, e2 = rel_err(bd["pillar_scale_features"], rs)
    assert e1 <= TOL_FP32 and e2 <= TOL_FP32, (e1, e2)
    assert torch.equal(bd["pillar_mask"].cpu(), rm)
    # element-wise too: |a-b| <= 1e-4 * (|b| + max|b| * 1e-2)
    d = (bd["pillar_features"].cpu() - rf).abs()
    assert bool((d <= 1e-4 * (rf.abs() + 1e-2 * rf.abs().max())).all())


def test_pfn_plain_pillarvfe():
    g = G1
    frames = synth.make_batch("L", 20000, g.point_cloud_range, 1, first_frame=5)
    (rf, rs, rm), bd = _pfn_case(g, frames, 4, scale=False)
    assert "pillar_scale_features" not in bd
    e1, e2 = rel_err(bd["pillar_features"], rf)
    assert e1 <= TOL_FP32 and e2 <= TOL_FP32


def test_pfn_dense_pillars_and_single_pillar():
    """pillars with n == 32 (no padded row), n == 1, and a batch with a single pillar (E8: output keeps (P,64))."""
    g = Geometry(G2.point_cloud_range, G2.voxel_size, 32, 40000)
    dense = synth.make_frame("U", 40000, (5.0, 0.0, -2.0, 6.6, 1.6, 0.0), 5)        # 100 cells x ~400 pts
    (rf, rs, rm), bd = _pfn_case(g, [dense], 8)
    assert int((torch.from_numpy(ov.voxelize_c(dense, g.range_f32, g.voxel_f32)[2]) == 32).sum()) > 50
    assert rel_err(bd["pillar_features"], rf)[0] <= TOL_FP32
    one = synth.make_frame("U", 7, (10.09, 0.97, -1.0, 10.23, 1.11, 0.0), 5)
    (rf, rs, rm), bd = _pfn_case(g, [one], 8)
    assert tuple(bd["pillar_features"].shape) == (1, 64)
    assert rel_err(bd["pillar_features"], rf)[0] <= TOL_FP32 and rel_err(bd["pillar_scale_features"], rs)[0] <= TOL_FP32


@pytest.mark.parametrize("name", ["tiny_continue", "tiny_break_cap", "tiny_t5"])
def test_pfn_golden_small(name):
    """against tensors the REFERENCE'S OWN PillarVFE_Scale produced (tests/golden)."""
    from hvpr_b200 import config, vfe
    z, geom, frames, overflow, wseed = load_small(name)
    w = hybrid.random_weights(wseed)
    m = vfe.PillarVFE_Scale(config.HVPR_VFE_CFG, 4, list(geom.voxel_size), geom.range_f32).cuda().eval()
    m.load_state_dict({k[4:]: v for k, v in w.items() if k.startswith("vfe.")}, strict=False)
    bd = dict(voxels=torch.from_numpy(z["voxels"]).cuda(), voxel_num_points=torch.from_numpy(z["voxel_num_points"]).cuda(),
              voxel_coords=torch.from_numpy(z["voxel_coords"]).cuda())
    with torch.no_grad():
        bd = m(bd)
    assert rel_err(bd["pillar_features"], torch.from_numpy(z["pillar_features"]))[0] <= TOL_FP32
    assert rel_err(bd["pillar_scale_features"], torch.from_numpy(z["pillar_scale_features"]))[0] <= TOL_FP32


# ------------------------------------------------------------------------------------------------ K3 memory attention
def _mem_inputs(P, seed):
    g = torch.Generator().manual_seed(seed)
    pil = torch.relu(torch.randn(P, 64, generator=g) * 1.5 + 0.3)        # non-negative, like post-ReLU pillar features
    W = (torch.rand(2000, 64, generator=g) * 2 - 1) / 8.0
    return pil, W


@pytest.mark.parametrize("precision,tol", [("fp32", TOL_FP32), ("bf16_rescore", TOL_FP32)])
def test_mem_attn_vs_oracle(precision, tol):
    from hvpr_b200.map_to_bev import MemoryUnit_Agg
    pil, W = _mem_inputs(20011, 3)
    with torch.no_grad():
        ref, ridx = hybrid.memory_attention(pil, W, 20, return_indices=True)
    m = MemoryUnit_Agg(2000, 64).cuda().eval()
    m.precision = precision
    with torch.no_grad():
        m.weight.copy_(W)
    idx = torch.full((pil.shape[0], 20), -1, dtype=torch.int32, device="cuda")
    out = m.run(pil.cuda(), 20, topk_idx_out=idx)
    torch.cuda.synchronize()
    # top-k is discontinuous: rows whose 20th/21st logits tie within fp32 noise may legitimately pick the other item
    err, frac = tie_aware_readout_check(out, ref, pil, W, tol, idx=idx, ref_idx=ridx)
    assert err <= tol
    assert rel_err(out, ref)[1] <= 5e-3              # L2 over the whole tensor, tie rows included
    a = torch.sort(idx.cpu().long(), 1)[0]
    b = torch.sort(ridx, 1)[0]
    assert float((a != b).any(1).float().mean()) <= 2e-3
    assert bool((a[:, 1:] != a[:, :-1]).all()) and int(a.min()) >= 0 and int(a.max()) < 2000   # 20 distinct valid items


def test_mem_attn_tc_many_tiles_dead_and_degenerate_rows():
    """The tensor-core kernel's dynamic tile schedule (548 tiles: every CTA claims tiles beyond its prologue), a device-side row
    count below the buffer size (the last live tile is partial, the tiles after it are dead) and degenerate all-zero pillar rows
    (every logit ties: candidate overflow -> exact full-scan path).  Live rows match the oracle, dead rows are never written."""
    from hvpr_b200.map_to_bev import MemoryUnit_Agg
    rows, live = 70001, 65003
    pil, W = _mem_inputs(rows, 11)
    zero_rows = list(range(700, live, 1499)) + list(range(12790, 12810))      # isolated ones + a run across a tile boundary
    pil[zero_rows] = 0.0
    with torch.no_grad():
        ref = hybrid.memory_attention(pil[:live], W, 20)
    m = MemoryUnit_Agg(2000, 64).cuda().eval()
    m.precision = "bf16_rescore"
    with torch.no_grad():
        m.weight.copy_(W)
    n_dev = torch.tensor([live], dtype=torch.int32, device="cuda")
    out = torch.full((rows, 64), float("nan"), device="cuda")
    idx = torch.full((rows, 20), -1, dtype=torch.int32, device="cuda")
    for _ in range(2):                                   # twice: the tile counter must restart from zero on every launch
        out.fill_(float("nan"))
        m.run(pil.cuda(), 20, n_pillars_dev=n_dev, out=out, topk_idx_out=idx)
    torch.cuda.synchronize()
    assert bool(torch.isnan(out[live:]).all()) and int(idx[live:].max()) == -1
    got = out[:live].cpu()
    assert bool(torch.isfinite(got).all())
    zr = torch.tensor(zero_rows)
    # an all-zero row: 20 equal logits -> uniform weights over the 20 lowest-numbered items (torch.topk's pick is unspecified)
    assert bool((idx[:live].cpu()[zr].long() == torch.arange(20)).all())
    assert float((got[zr] - W[:20].mean(0)).abs().max()) <= 1e-6
    keep = torch.ones(live, dtype=torch.bool); keep[zr] = False
    err, _ = tie_aware_readout_check(got[keep], ref[keep], pil[:live][keep], W, TOL_FP32, idx=idx[:live].cpu()[keep])
    assert err <= TOL_FP32
    a = torch.sort(idx[:live].cpu().long()[keep], 1)[0]
    assert bool((a[:, 1:] != a[:, :-1]).all()) and int(a.min()) >= 0 and int(a.max()) < 2000


def test_mem_attn_tc_gemm_logits():
    """tcgen05 plumbing in isolation: the TMEM accumulators equal a bf16-input / fp32-accumulate matmul
    (descriptors, 128-B swizzle, instruction descriptor, tcgen05.ld lane mapping)."""
    import ctypes
    from hvpr_b200 import _lib
    _lib.init_device()
    L = _lib.lib()
    pil, W = _mem_inputs(1000, 5)            # 7.8 tiles: exercises the ragged last tile
    pd, Wd = pil.cuda(), W.cuda()
    wpk = torch.empty((2048, 64), dtype=torch.bfloat16, device="cuda")
    _lib.check(L.hvpr_mem_pack_bf16(_lib.ptr(Wd), 2000, 64, _lib.ptr(wpk), _lib.cur_stream()))
    nb = L.hvpr_mem_attn_workspace_bytes(1000, 2000, 1)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    logits = torch.full((1000, 2048), float("nan"), device="cuda")
    out = torch.empty((1000, 64), device="cuda")
    idx = torch.full((1000, 20), -1, dtype=torch.int32, device="cuda")
    L.hvpr_dbg_mem_attn_logits.restype = ctypes.c_int
    L.hvpr_dbg_mem_attn_logits.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                           ctypes.c_void_p, ctypes.c_void_p]
    _lib.check(L.hvpr_dbg_mem_attn_logits(_lib.ptr(pd), 1000, _lib.ptr(Wd), _lib.ptr(wpk), 2000, _lib.ptr(out),
                                          _lib.ptr(idx), _lib.ptr(ws), nb, _lib.ptr(logits), _lib.cur_stream()))
    torch.cuda.synchronize()
    ref = pil.bfloat16().float() @ W.bfloat16().float().t()
    got = logits.cpu()
    assert bool(torch.isinf(got[:, 2000:]).all()) and bool((got[:, 2000:] < 0).all())      # padded items masked
    assert float((got[:, :2000] - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    with torch.no_grad():
        r = hybrid.memory_attention(pil, W, 20)
    tie_aware_readout_check(out, r, pil, W, TOL_FP32, idx=idx)


def test_mem_attn_module_forward_signature():
    """MemoryUnit_Agg.forward(input1, input2, k) -> {'output','att'} (memory_module.py:29,77)."""
    from hvpr_b200.map_to_bev import MemoryUnit_Agg
    pil, W = _mem_inputs(333, 4)
    m = MemoryUnit_Agg(2000, 64).cuda().eval()
    with torch.no_grad():
        m.weight.copy_(W)
        r = m(pil.cuda(), None, 20)
        ref = hybrid.memory_attention(pil, W, 20)
    assert rel_err(r["output"], ref)[0] <= TOL_FP32


@pytest.mark.parametrize("mem_dim,k", [(500, 10), (200, 20), (2000, 8), (3000, 20)])
def test_mem_attn_module_other_memory_shapes_use_the_exact_kernel(mem_dim, k):
    """Memory sizes / k outside the tcgen05 kernel's build (k = 20, 384 <= M <= 2048) are served by the exact fp32 CUDA kernel
    without the caller having to switch `precision` (memory_module.py:60-77 for any mem_dim / k)."""
    from hvpr_b200.map_to_bev import MemoryUnit_Agg
    gen = torch.Generator().manual_seed(mem_dim + k)
    pil = torch.randn(777, 64, generator=gen).relu()
    W = (torch.rand(mem_dim, 64, generator=gen) - 0.5) / 4
    m = MemoryUnit_Agg(mem_dim, 64).cuda().eval()
    assert m.precision == "bf16_rescore"
    with torch.no_grad():
        m.weight.copy_(W)
        r = m(pil.cuda(), None, k)
        ref = hybrid.memory_attention(pil, W, k)
    tie_aware_readout_check(r["output"], ref, pil, W, TOL_FP32, k=k)


# ------------------------------------------------------------------------------------------------ K3 side job: zero fill
@pytest.mark.parametrize("mode", ["bf16_rescore", "fp32"])
def test_mem_attn_zero_fill_side_job(mode):
    """hvpr_mem_attn(zero_fill=...) leaves every listed range all-zero — ranges that are not multiples of the 8 KB store unit,
    more units than chunk iterations, a single tile, no live rows at all — never writes past a range, and the readout is the
    same bits as without the side job."""
    from hvpr_b200 import _lib, map_to_bev
    w = hybrid.random_weights(0)
    mem = map_to_bev.MemoryUnit_Agg(2000, 64, 0.0025).cuda()
    mem.weight.data.copy_(w["map_to_bev_module.memory.weight"])
    mem.precision = mode
    gen = torch.Generator().manual_seed(5)
    for rows, live in ((70000, None), (100, None), (40000, 0), (300, 129)):
        x = torch.randn(rows, 64, generator=gen).cuda()
        n_dev = None if live is None else torch.tensor([live], dtype=torch.int32, device="cuda")
        base = mem.run(x, 20, n_dev).clone()
        sizes = (16, 8192 * 3 + 4112, 50_000_000, 8192)               # bytes; every range is guarded by 64 poisoned bytes on both sides
        bufs = [torch.full((n + 128,), 0xAB, dtype=torch.uint8, device="cuda") for n in sizes]
        out = mem.run(x, 20, n_dev, zero_fill=[b[64:64 + n] for b, n in zip(bufs, sizes)])
        torch.cuda.synchronize()
        nl = rows if live is None else live
        assert torch.equal(out[:nl], base[:nl]), (rows, live)
        for b, n in zip(bufs, sizes):
            assert int(b[64:64 + n].max()) == 0, (rows, live, n)
            assert bool((b[:64] == 0xAB).all()) and bool((b[64 + n:] == 0xAB).all()), (rows, live, n)
    # argument checks: misaligned pointer / size, too many ranges
    L = _lib.lib()
    z = _lib.HvprZeroFill(); z.n = 5
    args = lambda zf: (_lib.ptr(x), None, 300, _lib.ptr(mem.weight.detach()), None, 2000, 64, 20, _lib.MEM_FP32, _lib.ptr(out), None, None, 0,
                       zf, _lib.cur_stream())
    import ctypes
    assert L.hvpr_mem_attn(*args(ctypes.byref(z))) == -1
    z.n = 1; z.ptr[0] = bufs[0].data_ptr() + 4; z.bytes[0] = 16
    assert L.hvpr_mem_attn(*args(ctypes.byref(z))) == -1
    z.ptr[0] = bufs[0].data_ptr(); z.bytes[0] = 24
    assert L.hvpr_mem_attn(*args(ctypes.byref(z))) == -1


# ------------------------------------------------------------------------------------------------ K4 BEV fill
@pytest.mark.parametrize("gname", ["G1", "G2"])
def test_bev_fill_bitwise(gname):
    from hvpr_b200 import config, map_to_bev
    g = {"G1": G1, "G2": G2}[gname]
    nx, ny, _ = g.grid_size
    frames = synth.make_batch("L", 60000, g.point_cloud_range, 3, first_frame=2)
    rv, rc, rn = ov.voxelize_batch(frames, g.range_f32, g.voxel_f32)
    P = len(rn)
    gen = torch.Generator().manual_seed(0)
    pf, ro, ps = torch.randn(P, 64, generator=gen), torch.randn(P, 64, generator=gen), torch.randn(P, 32, generator=gen)
    tc = torch.from_numpy(rc)
    ref = torch.zeros(3, 128, ny * nx); refs = torch.zeros(3, 32, ny * nx)
    for b in range(3):
        m = tc[:, 0] == b
        idx = (tc[m, 1] + tc[m, 2] * nx + tc[m, 3]).long()
        ref[b][:, idx] = torch.cat([pf[m].t(), ro[m].t()], 0)
        refs[b][:, idx] = ps[m].t()
    from hvpr_b200 import _lib
    _lib.init_device()
    cm = torch.empty((3, nx * ny), dtype=torch.int32, device="cuda")
    coords = tc.cuda()
    _lib.check(_lib.lib().hvpr_build_cell_map(_lib.ptr(coords), None, P, 3, nx, ny, _lib.ptr(cm), _lib.cur_stream()))
    sp = torch.full((3, 128, ny, nx), float("nan"), device="cuda"); sps = torch.full((3, 32, ny, nx), float("nan"), device="cuda")
    a, b_, s = pf.cuda(), ro.cuda(), ps.cuda()
    _lib.check(_lib.lib().hvpr_bev_fill(_lib.ptr(a), 64, _lib.ptr(b_), 64, _lib.ptr(s), 32, _lib.ptr(cm), 3, nx, ny,
                                        _lib.ptr(sp), _lib.ptr(sps), None, _lib.cur_stream()))
    torch.cuda.synchronize()
    assert torch.equal(sp.cpu().view(3, 128, -1), ref) and torch.equal(sps.cpu().view(3, 32, -1), refs)
    # persistent form (grid-stride over the items, next item's map prefetched): 1, 2 and 8 blocks per SM, same bits
    for bps in (1, 2, 8):
        sp.fill_(float("nan")); sps.fill_(float("nan"))
        _lib.check(_lib.lib().hvpr_bev_fill(_lib.ptr(a), 64, _lib.ptr(b_), 64, _lib.ptr(s), 32, _lib.ptr(cm), 3, nx, ny,
                                            _lib.ptr(sp), _lib.ptr(sps), _lib.launch_cfg((bps, 0)), _lib.cur_stream()))
        torch.cuda.synchronize()
        assert torch.equal(sp.cpu().view(3, 128, -1), ref) and torch.equal(sps.cpu().view(3, 32, -1), refs), bps
    # canvas-is-zero form: only 32 / 64 / 128-byte runs that hold a pillar are written; on zeroed canvases the bits are the same,
    # with and without the persistent grid
    for variant in (1, 2, 3):
        for bps in (0, 2):
            sp.zero_(); sps.zero_()
            _lib.check(_lib.lib().hvpr_bev_fill(_lib.ptr(a), 64, _lib.ptr(b_), 64, _lib.ptr(s), 32, _lib.ptr(cm), 3, nx, ny,
                                                _lib.ptr(sp), _lib.ptr(sps), _lib.launch_cfg((bps, variant)), _lib.cur_stream()))
            torch.cuda.synchronize()
            assert torch.equal(sp.cpu().view(3, 128, -1), ref) and torch.equal(sps.cpu().view(3, 32, -1), refs), (variant, bps)
    # ... and it really leaves empty runs alone (that is the point): a poisoned canvas keeps its poison exactly there
    sp.fill_(float("nan")); sps.fill_(float("nan"))
    _lib.check(_lib.lib().hvpr_bev_fill(_lib.ptr(a), 64, _lib.ptr(b_), 64, _lib.ptr(s), 32, _lib.ptr(cm), 3, nx, ny,
                                        _lib.ptr(sp), _lib.ptr(sps), _lib.launch_cfg((0, 1)), _lib.cur_stream()))
    torch.cuda.synchronize()
    occ8 = (cm.view(3, -1, 8) >= 0).any(-1).cpu()                            # 8 cells = 32 bytes per channel
    got = sp.cpu().view(3, 128, -1, 8)
    assert bool(torch.isnan(got[:, 0][~occ8]).all()) and not bool(torch.isnan(got[:, 0][occ8]).any())
    assert _lib.lib().hvpr_bev_fill(_lib.ptr(a), 64, _lib.ptr(b_), 64, _lib.ptr(s), 32, _lib.ptr(cm), 3, nx, ny,
                                    _lib.ptr(sp), _lib.ptr(sps), _lib.launch_cfg((0, 4)), _lib.cur_stream()) == -1
    for bad in (17, -1):       # the launch shape is validated per call (there is no process-wide knob any more)
        assert _lib.lib().hvpr_bev_fill(_lib.ptr(a), 64, _lib.ptr(b_), 64, _lib.ptr(s), 32, _lib.ptr(cm), 3, nx, ny,
                                        _lib.ptr(sp), _lib.ptr(sps), _lib.launch_cfg((bad, 0)), _lib.cur_stream()) == -1
    # vanilla PointPillarScatter module, fp32 coords, no batch_size key (falls back to the reference formula)
    mod = map_to_bev.PointPillarScatter(config.Cfg(NUM_BEV_FEATURES=64), grid_size=g.grid_size)
    bd = mod(dict(pillar_features=a, voxel_coords=coords.float()))
    assert torch.equal(bd["spatial_features"].cpu().view(3, 64, -1), ref[:, :64])


def test_bev_fill_odd_shapes_fallback():
    from hvpr_b200 import _lib
    _lib.init_device()
    nx, ny, B, P = 37, 29, 2, 300
    gen = torch.Generator().manual_seed(1)
    cells = torch.stack([torch.randperm(nx * ny, generator=gen)[:P // 2] for _ in range(B)])
    coords = torch.zeros(P, 4, dtype=torch.int32)
    for b in range(B):
        sl = slice(b * P // 2, (b + 1) * P // 2)
        coords[sl, 0] = b; coords[sl, 2] = (cells[b] // nx).int(); coords[sl, 3] = (cells[b] % nx).int()
    pf = torch.randn(P, 6, generator=gen)
    ref = torch.zeros(B, 6, ny * nx)
    for b in range(B):
        m = coords[:, 0] == b
        ref[b][:, (coords[m, 2] * nx + coords[m, 3]).long()] = pf[m].t()
    cm = torch.empty((B, nx * ny), dtype=torch.int32, device="cuda")
    cd, pd = coords.cuda(), pf.cuda()
    _lib.check(_lib.lib().hvpr_build_cell_map(_lib.ptr(cd), None, P, B, nx, ny, _lib.ptr(cm), _lib.cur_stream()))
    out = torch.full((B, 6, ny, nx), float("nan"), device="cuda")
    _lib.check(_lib.lib().hvpr_bev_fill(_lib.ptr(pd), 6, None, 0, None, 0, _lib.ptr(cm), B, nx, ny, _lib.ptr(out), None, None, _lib.cur_stream()))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu().view(B, 6, -1), ref)


# ------------------------------------------------------------------------------------------------ whole path
@pytest.mark.parametrize("name", ["tiny_continue", "tiny_break_cap", "tiny_t5"])
def test_frontend_golden_small(name):
    """raw points -> BEV canvases vs tensors the REFERENCE'S OWN modules produced (tests/golden)."""
    z, geom, frames, overflow, wseed = load_small(name)
    fe = _frontend(geom, hybrid.random_weights(wseed), overflow)
    bd = fe(dict(points=torch.from_numpy(synth.collate_points(frames)).cuda(), batch_size=len(frames)))
    torch.cuda.synchronize()
    assert np.array_equal(bd["voxel_coords"].cpu().numpy(), z["voxel_coords"])
    assert np.array_equal(bd["voxel_num_points"].cpu().numpy(), z["voxel_num_points"])
    assert np.array_equal(bd["voxels"].cpu().numpy().view(np.int32), z["voxels"].view(np.int32))
    for k in ("pillar_features", "pillar_scale_features", "memory_readout", "spatial_features", "spatial_scale_features"):
        e1, e2 = rel_err(bd[k], torch.from_numpy(z[k]))
        assert e1 <= TOL_FP32 and e2 <= TOL_FP32, (k, e1, e2)
        assert tuple(bd[k].shape) == z[k].shape


@pytest.mark.parametrize("mem_precision", ["bf16_rescore", "fp32"])
@pytest.mark.parametrize("gname,dist", [("G1", "L"), ("G2", "L"), ("G2", "U")])
def test_frontend_planned_graph_vs_oracle(gname, dist, mem_precision):
    """cfg 1/2 shape: 120k-point frames through the planned, CUDA-graph-replayed path vs the oracle."""
    g = {"G1": G1, "G2": G2}[gname]
    B, N = 2, 120000
    frames = synth.make_batch(dist, N, g.point_cloud_range, B)
    w = hybrid.random_weights(31)
    o = hybrid.frontend(frames, g, w)
    fe = _frontend(g, w, mem_precision=mem_precision)
    p = fe.plan(B, B * N, N)
    pts, off = to_dev(frames)
    p.points.copy_(pts); p.frame_offsets.copy_(off)
    for _ in range(3):                                   # replays must be idempotent
        fe.run()
    torch.cuda.synchronize()
    P = int(p.vox.voxel_offsets[-1])
    assert P == o["voxel_coords"].shape[0]
    assert torch.equal(p.vox.coords[:P].cpu(), o["voxel_coords"])
    assert torch.equal(p.vox.num_points[:P].cpu(), o["voxel_num_points"])
    assert torch.equal(p.vox.voxels[:P].cpu().view(torch.int32), o["voxels"].view(torch.int32))
    for a, k in ((p.pillar_features[:P], "pillar_features"), (p.pillar_scale[:P], "pillar_scale_features"),
                 (p.spatial_scale, "spatial_scale_features"), (p.spatial[:, :64], "spatial_features")):
        ref = o[k][:, :64] if k == "spatial_features" else o[k]
        e1, e2 = rel_err(a, ref)
        assert e1 <= TOL_FP32 and e2 <= TOL_FP32, (k, e1, e2)
    # memory readout (canvas channels 64..127): tie-aware, see helpers.tie_aware_readout_check
    err, frac = tie_aware_readout_check(p.readout[:P], o["memory_readout"], o["pillar_features"],
                                        w["map_to_bev_module.memory.weight"], TOL_FP32)
    assert err <= TOL_FP32
    assert rel_err(p.spatial[:, 64:], o["spatial_features"][:, 64:])[1] <= 1e-2
    # size-independent properties: canvas is zero exactly off the occupied cells, and occupied columns equal the rows
    nx, ny, _ = g.grid_size
    sp = p.spatial.view(B, 128, -1)
    occ = torch.zeros(B, ny * nx, dtype=torch.bool, device="cuda")
    c = p.vox.coords[:P].long()
    occ[c[:, 0], c[:, 2] * nx + c[:, 3]] = True
    assert float(sp.abs().sum(1)[~occ].max()) == 0.0
    cols = sp[c[:, 0], :, c[:, 2] * nx + c[:, 3]]
    assert torch.equal(cols[:, :64], p.pillar_features[:P]) and torch.equal(cols[:, 64:], p.readout[:P])


def test_frontend_module_chain_matches_planned_path():
    """batch_dict module API (voxelize_batch -> VFE -> map_to_bev) == planned graph path, bitwise."""
    g = G1
    B, N = 3, 40000
    frames = synth.make_batch("L", N, g.point_cloud_range, B)
    w = hybrid.random_weights(2)
    fe = _frontend(g, w)
    bd = fe(dict(points=torch.from_numpy(synth.collate_points(frames)).cuda(), batch_size=B))
    p = fe.plan(B, B * N, N)
    pts, off = to_dev(frames)
    p.points.copy_(pts); p.frame_offsets.copy_(off)
    fe.run(); torch.cuda.synchronize()
    assert torch.equal(bd["spatial_features"], p.spatial) and torch.equal(bd["spatial_scale_features"], p.spatial_scale)
    assert bd["pillar_mask"].shape == (bd["voxels"].shape[0], 32, 1)


def test_frontend_run_host_end_to_end():
    g = G1
    B, N = 2, 30000
    frames = synth.make_batch("L", N, g.point_cloud_range, B)
    w = hybrid.random_weights(2)
    fe = _frontend(g, w)
    p = fe.plan(B, B * N, N)
    hp = torch.from_numpy(np.concatenate(frames, 0)).pin_memory()
    ho = torch.tensor([0, N, 2 * N], dtype=torch.int32).pin_memory()
    hc = torch.zeros(B + 1, dtype=torch.int32).pin_memory()
    fe.run_host(hp, ho, hc); torch.cuda.synchronize()
    o = hybrid.frontend(frames, g, w)
    assert int(hc[-1]) == o["voxel_coords"].shape[0]
    assert rel_err(p.spatial, o["spatial_features"])[0] <= TOL_FP32


# ------------------------------------------------------------------------------------------------ BASELINE configs 3 / 4
def test_config4_dense_frames_extended_range():
    """BASELINE.json configs[3]: 300k-point frames, 640x640 grid, 80k max pillars (SURVEY §8d cfg 4) — voxelization
    bit-exact vs the oracle, features within tolerance, canvases consistent with the rows."""
    g = G3
    B, N = 2, 300000
    frames = synth.make_batch("L", N, g.point_cloud_range, B, first_frame=40)
    w = hybrid.random_weights(5)
    fe = _frontend(g, w)
    p = fe.plan(B, B * N, N)
    pts, off = to_dev(frames)
    p.points.copy_(pts); p.frame_offsets.copy_(off)
    fe.run(); torch.cuda.synchronize()
    P = int(p.vox.voxel_offsets[-1])
    rv, rc, rn = ov.voxelize_batch(frames, g.range_f32, g.voxel_f32, 32, g.max_voxels)
    assert P == len(rn)
    assert np.array_equal(p.vox.coords[:P].cpu().numpy(), rc) and np.array_equal(p.vox.num_points[:P].cpu().numpy(), rn)
    assert np.array_equal(p.vox.voxels[:P].cpu().numpy().view(np.int32), rv.view(np.int32))
    with torch.no_grad():
        pf, psf, _ = hybrid.pillar_vfe(torch.from_numpy(rv), torch.from_numpy(rn), torch.from_numpy(rc), w,
                                       list(g.voxel_size), g.range_f32)
        ro = hybrid.memory_attention(pf, w["map_to_bev_module.memory.weight"], 20)
    assert rel_err(p.pillar_features[:P], pf)[0] <= TOL_FP32 and rel_err(p.pillar_scale[:P], psf)[0] <= TOL_FP32
    tie_aware_readout_check(p.readout[:P], ro, pf, w["map_to_bev_module.memory.weight"], TOL_FP32)
    nx, ny, _ = g.grid_size
    c = p.vox.coords[:P].long()
    cols = p.spatial.view(B, 128, -1)[c[:, 0], :, c[:, 2] * nx + c[:, 3]]
    assert torch.equal(cols[:, :64], p.pillar_features[:P]) and torch.equal(cols[:, 64:], p.readout[:P])
    assert int((p.spatial.view(B, 128, -1).abs().sum(1) != 0).sum()) <= P
    assert tuple(p.spatial.shape) == (B, 128, 640, 640) and tuple(p.spatial_scale.shape) == (B, 32, 640, 640)


def test_config3_many_frames_equals_per_frame_runs():
    """BASELINE.json configs[2]: a big batch sharded frame-wise gives the same bits as running its frames alone —
    the property the multi-GPU partition (frame i -> rank i mod W) relies on."""
    g = G2
    N = 30000
    frames = synth.make_batch("L", N, g.point_cloud_range, 16, first_frame=60)
    w = hybrid.random_weights(6)
    fe = _frontend(g, w)
    p = fe.plan(16, 16 * N, N)
    pts, off = to_dev(frames)
    p.points.copy_(pts); p.frame_offsets.copy_(off)
    fe.run(); torch.cuda.synchronize()
    whole = p.spatial.clone(); whole_s = p.spatial_scale.clone()
    vo = p.vox.voxel_offsets.cpu().numpy()
    fe2 = _frontend(g, w)
    p2 = fe2.plan(4, 4 * N, N)
    for r in range(4):                                   # "rank" r owns frames r, r+4, r+8, r+12
        mine = frames[r::4]
        pts, off = to_dev(mine)
        p2.points.copy_(pts); p2.frame_offsets.copy_(off)
        fe2.run(); torch.cuda.synchronize()
        assert torch.equal(p2.spatial, whole[r::4]) and torch.equal(p2.spatial_scale, whole_s[r::4])
        vo2 = p2.vox.voxel_offsets.cpu().numpy()
        assert list(np.diff(vo2)) == list(np.diff(vo)[r::4])


def test_streaming_mode_is_bit_identical_to_single_stream():
    """Software-pipelined batches (K1 of batch k+1 on a side stream while K2-K4 of batch k run) produce the same bits as
    the plain path, for alternating DIFFERENT batches fed from pinned host memory."""
    g = G1
    B, N = 2, 40000
    w = hybrid.random_weights(9)
    batches = [synth.make_batch("L", N, g.point_cloud_range, B, first_frame=10 * i) for i in range(3)]
    fe = _frontend(g, w)
    ref = []
    p = fe.plan(B, B * N, N)
    for fr in batches:
        pts, off = to_dev(fr)
        p.points.copy_(pts); p.frame_offsets.copy_(off)
        fe.run(); torch.cuda.synchronize()
        ref.append((p.spatial.clone(), p.spatial_scale.clone(), p.vox.voxel_offsets.clone()))
    sp = fe.plan_stream(B, B * N, N)
    host = [(torch.from_numpy(np.concatenate(fr, 0)).pin_memory(),
             torch.tensor(np.r_[0, np.cumsum([len(f) for f in fr])], dtype=torch.int32).pin_memory()) for fr in batches]
    cnt = torch.zeros(B + 1, dtype=torch.int32).pin_memory()
    order = [0, 1, 2, 0, 2, 1, 1, 0, 2, 2]
    fe.stream_prime(host[order[0]], host[order[1]])
    for i, b in enumerate(order):
        nb = order[i + 2] if i + 2 < len(order) else 0            # the batch voxelized by this step finishes 2 steps later
        fe.stream_step(host[nb][0], host[nb][1], cnt)
        torch.cuda.synchronize()
        assert torch.equal(sp.spatial, ref[b][0]) and torch.equal(sp.spatial_scale, ref[b][1]), (i, b)
        assert torch.equal(cnt, ref[b][2].cpu())


@pytest.mark.parametrize("variant", [1, 3])
def test_fused_zero_fill_front_end_is_bit_identical(variant):
    """FUSED_ZERO_FILL: the memory kernel zero-fills both canvases as a side job and the fill writes only occupied runs — same
    canvases, bit for bit, as the write-everything path; plain, graph-captured and streaming, with poisoned canvases in front."""
    g = G1
    B, N = 2, 40000
    w = hybrid.random_weights(9)
    batches = [synth.make_batch("L", N, g.point_cloud_range, B, first_frame=10 * i) for i in range(2)]
    fe = _frontend(g, w)
    p = fe.plan(B, B * N, N, use_graph=False)
    ref = []
    for fr in batches:
        pts, off = to_dev(fr)
        p.points.copy_(pts); p.frame_offsets.copy_(off)
        fe.run(); torch.cuda.synchronize()
        ref.append((p.spatial.clone(), p.spatial_scale.clone()))
    fe2 = _frontend(g, w)
    fe2.map_to_bev_module.fused_zero_fill = variant
    for use_graph in (False, True):
        p2 = fe2.plan(B, B * N, N, use_graph=use_graph)
        for rep in range(2):
            for fr, r in zip(batches, ref):
                pts, off = to_dev(fr)
                p2.points.copy_(pts); p2.frame_offsets.copy_(off)
                p2.spatial.fill_(float("nan")); p2.spatial_scale.fill_(float("nan"))
                fe2.run(); torch.cuda.synchronize()
                assert torch.equal(p2.spatial, r[0]) and torch.equal(p2.spatial_scale, r[1]), (use_graph, rep)
    sp = fe2.plan_stream(B, B * N, N)
    host = [(torch.from_numpy(np.concatenate(fr, 0)).pin_memory(),
             torch.tensor(np.r_[0, np.cumsum([len(f) for f in fr])], dtype=torch.int32).pin_memory()) for fr in batches]
    cnt = torch.zeros(B + 1, dtype=torch.int32).pin_memory()
    order = [0, 1, 1, 0, 0, 1]
    fe2.stream_prime(host[order[0]], host[order[1]])
    for i, b in enumerate(order):
        nb = order[i + 2] if i + 2 < len(order) else 0
        sp.spatial.fill_(float("nan")); sp.spatial_scale.fill_(float("nan"))
        torch.cuda.synchronize()
        fe2.stream_step(host[nb][0], host[nb][1], cnt)
        torch.cuda.synchronize()
        assert torch.equal(sp.spatial, ref[b][0]) and torch.equal(sp.spatial_scale, ref[b][1]), (i, b)


def test_two_front_ends_two_streams_with_different_launch_shapes():
    """SURVEY §8b #3 (no global state, re-entrant across streams): two front ends in ONE process, on two streams, with
    different HvprLaunchCfg shapes for the PFN and the canvas fill, interleaved call by call, produce the bits each one
    produces alone.  (Round 1 had process-global hvpr_tune_* knobs that two front ends could race on.)"""
    from hvpr_b200 import _lib
    g = G1
    w1, w2 = hybrid.random_weights(21), hybrid.random_weights(22)
    fr1 = synth.make_batch("L", 30000, g.point_cloud_range, 2, first_frame=40)
    fr2 = synth.make_batch("U", 25000, g.point_cloud_range, 2, first_frame=50)
    shapes = [((3, 0), None), ((1, 1), (2, 0))]
    alone = []
    for fr, w, (pfn_cfg, bev_cfg) in ((fr1, w1, shapes[0]), (fr2, w2, shapes[1])):
        fe = _frontend(g, w)
        p = fe.plan(2, sum(len(f) for f in fr), max(len(f) for f in fr), use_graph=False)
        pts, off = to_dev(fr)
        p.points.copy_(pts); p.frame_offsets.copy_(off)
        fe.run(); torch.cuda.synchronize()
        alone.append((p.spatial.clone(), p.spatial_scale.clone()))
    fes, plans, streams = [], [], [torch.cuda.Stream(), torch.cuda.Stream()]
    for fr, w in ((fr1, w1), (fr2, w2)):
        fe = _frontend(g, w)
        p = fe.plan(2, sum(len(f) for f in fr), max(len(f) for f in fr), use_graph=False)
        pts, off = to_dev(fr)
        p.points.copy_(pts); p.frame_offsets.copy_(off)
        fes.append(fe); plans.append(p)
    torch.cuda.synchronize()
    for rep in range(3):
        for p in plans:
            p.spatial.fill_(float("nan")); p.spatial_scale.fill_(float("nan"))
        torch.cuda.synchronize()
        # stage by stage, alternating between the two front ends / streams, each with its own launch shape
        voxs = []
        for i in (0, 1):
            with torch.cuda.stream(streams[i]):
                voxs.append(fes[i].voxelizer.run(plans[i].points, plans[i].frame_offsets, 2, plans[i].max_frame_points, out=plans[i].vox))
        for i in (1, 0):
            with torch.cuda.stream(streams[i]):
                v, p = voxs[i], plans[i]
                fes[i].vfe.run(v.voxels, v.num_points, v.coords, v.n_pillars_dev, out=p.pillar_features,
                               scale_out=p.pillar_scale, launch=shapes[i][0])
        for i in (0, 1):
            with torch.cuda.stream(streams[i]):
                v, p, m = voxs[i], plans[i], fes[i].map_to_bev_module
                m.memory.run(p.pillar_features, m.k, v.n_pillars_dev, out=p.readout)
                _lib.check(_lib.lib().hvpr_bev_fill(_lib.ptr(p.pillar_features), 64, _lib.ptr(p.readout), 64, _lib.ptr(p.pillar_scale), 32,
                                                    _lib.ptr(v.cell_map), 2, m.nx, m.ny, _lib.ptr(p.spatial), _lib.ptr(p.spatial_scale),
                                                    _lib.launch_cfg(shapes[i][1]), _lib.cur_stream()))
        torch.cuda.synchronize()
        for i in (0, 1):
            assert torch.equal(plans[i].spatial, alone[i][0]) and torch.equal(plans[i].spatial_scale, alone[i][1]), (rep, i)
