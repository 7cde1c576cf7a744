#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json: VFE frames/sec @120k pts/frame on B200.

One "step" = one pass of the hot path (voxelize -> fused PFN -> memory attention -> BEV fill) over one batch of
synthetic frames: BASELINE.json configs[1] = G2 (pillar 0.16x0.16x4 m, 432x496 grid, 32 pts/pillar, 40k max pillars),
B = 8 frames of 120 000 points per GPU.  N GPUs => N ranks (torchrun), each with its own 8 frames (weak scaling, no
collective on the data path; NCCL only reduces the timings).

    python bench.py [--gpus N] [--steps K] [--warmup W]           our arm (BASELINE.json configs[1], "cfg2")
    python bench.py --impl reference ...                           the reference's CPU path (oracle port) on host cores
    python bench.py --config cfg3|cfg4 [--dist U]                  the other measured configs of BASELINE.json / SURVEY.md §8d:
        cfg3 = 64 frames split frame-wise over the ranks (strong scaling; N = 1 runs all 64), cfg4 = G3 640x640 grid,
        300 000 points per frame, 80 000 max pillars, 8 frames per GPU; --dist U = uniform frames (the 40 000-pillar cap is hit)

Prints ONE JSON line (rank 0).  At N = 1 the line also carries `next_rows.bev_backbone`: the widened rows N1-N3 of SURVEY.md §8f measured on
the same batch — BaseBEVBackbone_Scale on the tcgen05 conv kernel (ms, TFLOP/s vs the measured bf16 peak, the same network through cuDNN as the
library baseline), the dense head, the NMS post-processing, and points -> features / boxes / detections end to end (`--no-backbone` skips it).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "vfe_frames_per_sec_120k_pts"
UNIT = "frames/s"
FRAMES_PER_GPU = 8
POINTS_PER_FRAME = 120000
DIST = "L"
NOMINAL_HBM_GBS = 8000.0          # BASELINE.json: "roughly 8 TB/s"; fractions are quoted against the measured AND the nominal peak
NOMINAL_BF16_TFLOPS = 2250.0

# BASELINE.json configs[1..3] as concretised by SURVEY.md §8d
WORKLOADS = {
    "cfg2": dict(geom="G2", frames_per_gpu=8, frames_total=None, points=120000, scaling="weak",
                 what="BASELINE.json configs[1]: batch of 8 frames per GPU"),
    "cfg3": dict(geom="G2", frames_per_gpu=None, frames_total=64, points=120000, scaling="strong",
                 what="BASELINE.json configs[2]: 64 frames sharded frame-wise over the ranks (frame i -> rank i mod W)"),
    "cfg4": dict(geom="G3", frames_per_gpu=8, frames_total=None, points=300000, scaling="weak",
                 what="BASELINE.json configs[3]: dense 300k-point frames, extended range (640x640 grid), 80k max pillars"),
    "cfg5": dict(geom="G2", frames_per_gpu=8, frames_total=None, points=120000, scaling="weak",
                 what="BASELINE.json configs[4]: full single-stage inference (hybrid VFE + BEV 2D backbone + anchor head + rotated-IoU NMS)"),
}


def workload(args, rank, world):
    """-> (geom, frame ids of this rank, points per frame, frames per step over all ranks)"""
    from hvpr_b200 import sharding
    from hvpr_b200.geometry import GEOMETRIES
    wl = WORKLOADS[args.config]
    geom = GEOMETRIES[wl["geom"]]
    if wl["frames_total"]:
        ids = sharding.frames_of_rank(wl["frames_total"], rank, world)
        total = wl["frames_total"]
    else:
        ids = sharding.weak_scaling_frames(wl["frames_per_gpu"], rank)
        total = wl["frames_per_gpu"] * world
    return geom, ids, wl["points"], total


def pinned_like(arr, write_combined):
    """Host staging buffer of a batch: page-locked (torch pin_memory) or page-locked + write-combined (cudaHostAllocWriteCombined:
    no CPU cache snooping on the PCIe reads — the buffer is written once by the CPU and only read by the GPU's copy engine)."""
    import ctypes
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if not write_combined:
        return t.pin_memory()
    rt = None
    for name in ("libcudart.so.12", "libcudart.so"):
        for d in [os.path.join(os.path.dirname(torch.__file__), "lib"), "/usr/local/cuda/lib64", ""]:
            try:
                rt = ctypes.CDLL(os.path.join(d, name) if d else name)
                break
            except OSError:
                continue
        if rt is not None:
            break
    if rt is None:
        return t.pin_memory()
    ptr = ctypes.c_void_p()
    nbytes = max(t.numel() * t.element_size(), 16)
    if rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04)) != 0 or not ptr.value:
        return t.pin_memory()
    buf = (ctypes.c_byte * nbytes).from_address(ptr.value)
    out = torch.frombuffer(buf, dtype=t.dtype, count=t.numel()).view(t.shape)      # lives until process exit (never freed: a bench buffer)
    out.copy_(t)
    return out


def bind_numa(local_rank, world):
    """Before any pinned allocation: run this rank (and first-touch its pinned buffers) on ONE NUMA node — the GPU's own when
    NVML can tell them apart, else ranks are spread round-robin over the nodes.  Round 1 left every rank's pinned batch on
    node 0 and the 8-GPU end-to-end efficiency dropped to 0.93 while the device-resident one stayed at 0.99."""
    import glob
    nodes = []
    for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*"), key=lambda s: int(s.rsplit("node", 1)[1])):
        try:
            cpus = set()
            for part in open(os.path.join(d, "cpulist")).read().strip().split(","):
                if part:
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
            if cpus:
                nodes.append(cpus)
        except Exception:
            pass
    allowed = os.sched_getaffinity(0)
    nodes = [n & allowed for n in nodes if n & allowed]
    info = {"numa_nodes": len(nodes), "bound_node": None, "how": "single node or no topology: not bound"}
    if len(nodes) < 2:
        return info
    target, how = None, ""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(max(n) for n in nodes) + 64) // 64)
        gpu_cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1} & allowed
        hits = [i for i, n in enumerate(nodes) if n & gpu_cpus]
        if len(hits) == 1:
            target, how = hits[0], "NVML cpu affinity of the GPU"
    except Exception:
        pass
    if target is None:
        target, how = (local_rank * len(nodes)) // max(world, 1) % len(nodes), "round-robin over nodes (GPU affinity spans several)"
    try:
        os.sched_setaffinity(0, nodes[target])
        info.update(bound_node=target, how=how, cpus=len(nodes[target]))
    except Exception as e:
        info["how"] = "sched_setaffinity failed: %r" % (e,)
    return info


def algorithmic_bytes(N, P, K, nx, ny):
    """SURVEY.md §8d per-frame figures (fp32 = i32 = 4 B)."""
    return {
        "voxelize": 16 * N + 532 * P,
        "pfn": 16 * K + 404 * P,
        "mem_attn": 512 * P + 512000,
        "bev_fill": 4 * 160 * nx * ny + 4 * nx * ny + 640 * P,
    }


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region (pynvml; falls back to nvidia-smi)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}
            while not self._stop.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.002)
        except Exception as e:  # pragma: no cover
            self.reasons.add("sampler_error:%s" % type(e).__name__)

    def __enter__(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=2)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_frames(geom, ids, points, dist):
    from hvpr_b200 import synth
    return [synth.make_frame(dist, points, geom.point_cloud_range, 1024 + i) for i in ids]


# --------------------------------------------------------------------------------------------------- CPU arms
def cpu_reference_run(geom, frames, w, steps, warmup, frames_per_step, workers=4, components=False, gpu_check=None):
    """The reference's CPU path, as the oracle port (oracle/: the same torch CPU fp32 ops as the reference's modules, pinned
    bit-exactly on them where /root/reference exists; the GPU box has no reference tree, and the reference ships nothing to
    compile): spconv-style voxelizer single-threaded per frame (one frame per DataLoader worker, `--workers 4` default,
    tools/test.py:25) + PillarVFE_Scale + PointPillarScatter_Agg_Memory_1_scale on torch CPU with every host thread.
    components=True adds BASELINE.md §3.4's breakdown (t_vox single thread / x4 workers / x n_cores, t_vfe, t_bev);
    gpu_check(frames) -> dict of GPU tensors: the same frames through the CUDA path, compared with the oracle's outputs."""
    import concurrent.futures as cf
    import torch
    from oracle import hybrid
    from oracle import voxelize as ov
    ncores = len(os.sched_getaffinity(0))
    torch.set_num_threads(ncores)
    nx, ny, _ = geom.grid_size
    pool = cf.ThreadPoolExecutor(max_workers=min(workers, ncores))
    vox1 = lambda f: ov.voxelize_c(f, geom.range_f32, geom.voxel_f32, geom.max_points_per_voxel, geom.max_voxels, "continue")
    last = {}

    def collate(outs):
        vox = np.concatenate([o[0] for o in outs], 0)
        coords = np.concatenate([np.pad(o[1], ((0, 0), (1, 0)), constant_values=b) for b, o in enumerate(outs)], 0)
        nump = np.concatenate([o[2] for o in outs], 0)
        return torch.from_numpy(vox), torch.from_numpy(coords), torch.from_numpy(nump)

    def step(i):
        fs = [frames[(i * frames_per_step + j) % len(frames)] for j in range(frames_per_step)]
        tv, tc, tn = collate(list(pool.map(vox1, fs)))
        with torch.no_grad():
            pf, psf, _ = hybrid.pillar_vfe(tv, tn, tc, w, list(geom.voxel_size), geom.range_f32)
            sp, sps, ro = hybrid.scatter_agg_memory(pf, psf, tc, w["map_to_bev_module.memory.weight"], len(fs), nx, ny)
        last.update(frames=fs, pf=pf, ro=ro, sp=sp, sps=sps, tc=tc, tn=tn)
        return float(sp[0, 0, 0, 0])

    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    dt = time.perf_counter() - t0
    res = {"fps": frames_per_step * steps / dt, "ms_per_step": dt / steps * 1e3, "cores": ncores}
    if components:
        fs = [frames[j % len(frames)] for j in range(max(frames_per_step, 2))]
        t0 = time.perf_counter(); outs = [vox1(f) for f in fs]; t_vox1 = (time.perf_counter() - t0) / len(fs)
        many = [frames[j % len(frames)] for j in range(max(8, min(2 * ncores, 32)))]

        def pooled(nw):
            with cf.ThreadPoolExecutor(max_workers=nw) as ex:
                t0 = time.perf_counter(); list(ex.map(vox1, many)); return (time.perf_counter() - t0) / len(many)
        t_vox4, t_voxn = pooled(min(4, ncores)), pooled(ncores)
        tv, tc, tn = collate(outs)
        with torch.no_grad():
            t0 = time.perf_counter()
            pf, psf, _ = hybrid.pillar_vfe(tv, tn, tc, w, list(geom.voxel_size), geom.range_f32)
            t_vfe = (time.perf_counter() - t0) / len(fs)
            t0 = time.perf_counter()
            hybrid.scatter_agg_memory(pf, psf, tc, w["map_to_bev_module.memory.weight"], len(fs), nx, ny)
            t_bev = (time.perf_counter() - t0) / len(fs)
        res["components"] = {
            "t_vox_ms_per_frame": {"1_thread": t_vox1 * 1e3, "4_workers": t_vox4 * 1e3, "%d_cores" % ncores: t_voxn * 1e3},
            "t_vfe_ms_per_frame": t_vfe * 1e3, "t_bev_ms_per_frame": t_bev * 1e3,
            "frames_per_sec": {"vox_1_thread": 1.0 / (t_vox1 + t_vfe + t_bev), "vox_4_workers": 1.0 / (t_vox4 + t_vfe + t_bev),
                               "vox_all_cores": 1.0 / (t_voxn + t_vfe + t_bev)},
            "note": "frames/s = 1 / (t_vox + t_vfe + t_bev) as SURVEY.md §8d defines it; t_vfe / t_bev = oracle/hybrid.py (the reference's torch "
                    "CPU ops) on all %d cores" % ncores}
    if gpu_check is not None and last:
        # the oracle as CHECKER: the last sample step's frames through the CUDA path vs the oracle's tensors
        g = gpu_check(last["frames"])
        scale = float(last["ro"].abs().max())
        row_err = (g["readout"].double().cpu() - last["ro"].double()).abs().max(1)[0] / max(scale, 1e-30)
        off = row_err > 1e-4
        rel = lambda a, b: float((a.double().cpu() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
        res["parity_sample"] = {
            "frames": len(last["frames"]), "pillars": int(last["tn"].shape[0]),
            "coords_counts_bit_exact": bool(torch.equal(g["coords"].cpu(), last["tc"].int()) and torch.equal(g["num_points"].cpu(), last["tn"].int())),
            "pillar_features_max_rel_err": rel(g["pillar_features"], last["pf"]),
            "readout_rows_off_by_more_than_1e-4": int(off.sum()), "readout_tie_row_fraction": float(off.double().mean()),
            "readout_max_rel_err_other_rows": float(row_err[~off].max()) if (~off).any() else 0.0,
            "spatial_scale_features_bit_exact_given_inputs": rel(g["spatial_scale"], last["sps"]) <= 1e-4,
            "note": "top-20 is discontinuous: rows whose 20th/21st logits tie within fp32 summation noise legitimately pick another item "
                    "(tests/helpers.py::tie_aware_readout_check proves the near-tie in fp64 for every such row)"}
    return res


def run_reference_arm(args):
    from hvpr_b200 import sharding
    from oracle import hybrid
    rank, _, world = sharding.dist_env()
    if rank != 0:
        return 0
    geom, ids, points, total = workload(args, 0, 1)
    fps_step = 2
    frames = make_frames(geom, ids[:8], points, args.dist)
    w = hybrid.random_weights(0)
    steps, warmup = args.steps, args.warmup
    r = cpu_reference_run(geom, frames, w, steps, warmup, fps_step)
    fps, ms, ncores = r["fps"], r["ms_per_step"], r["cores"]
    sample = "%d frames/step out of the workload's frames (%s, %d pts/frame, dist %s)" % (fps_step, WORKLOADS[args.config]["geom"], points, args.dist)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": WORKLOADS[args.config]["scaling"], "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(geom, args, total, world),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": ncores, "kind": "port", "sample": sample,
                         "what": "oracle/ port of the reference's CPU path: C voxelizer (1 thread per frame, 4 workers) + the reference's torch "
                                 "CPU ops for VFE / memory / scatter on all cores; /root/reference is 100 % Python and absent on the GPU box"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_config(geom, args, frames_total, world):
    nx, ny, _ = geom.grid_size
    wl = WORKLOADS[args.config]
    fpg = frames_total // max(world, 1)
    return {"workload": "%s — HVPR KITTI cfg, %s pillars %.2fx%.2fx%g m, %dx%d grid, 32 pts/pillar, %dk max pillars; "
                        "%d synthetic %s frames x %d pts per step over %d GPU(s) (%d per GPU); voxelize+PFN+memory attention+BEV fill"
                        % (wl["what"], wl["geom"], geom.voxel_size[0], geom.voxel_size[1], geom.voxel_size[2], nx, ny, geom.max_voxels // 1000,
                           frames_total, {"L": "LiDAR-like", "U": "uniform"}[args.dist], wl["points"], world, fpg),
            "name": args.config, "frames_per_gpu": fpg, "frames_per_step": frames_total, "points_per_frame": wl["points"],
            "distribution": args.dist, "weights": "random-init (seed 0), BN stats randomised",
            "l2": "each step streams %.1f GB of canvas writes per GPU (> 8x the 126 MB L2), so inputs are evicted between steps"
                  % (fpg * 160 * nx * ny * 4 / 1e9),
            "streaming": "batches are software-pipelined 3 deep over CUDA streams: K1 voxelize of batch k+2 and K2 PFN of batch k+1 "
                         "overlap K3/K4 of batch k; every batch runs the same 8 kernels (value_single_stream = no overlap); a batch's "
                         "canvases are complete two steps after it was submitted (latency = 3 x ms_per_step)"}


# --------------------------------------------------------------------------------------------------- GPU arm
def backbone_flops(bb, B, H, W):
    tot, h, w, cin, cs = 0, H, W, bb.input_channels, bb.input_channels // 4
    for i, nf in enumerate(bb.num_filters):
        h, w = h // bb.layer_strides[i], w // bb.layer_strides[i]
        px = B * h * w
        tot += 2 * 9 * cin * nf * px + (bb.layer_nums[i] + bb.sfm_layer_nums[i]) * 2 * 9 * nf * nf * px
        tot += 2 * 9 * cs * bb.num_scale_filters[i] * px
        tot += 2 * nf * bb.num_upsample_filters[i] * bb.upsample_strides[i] ** 2 * px
        cin, cs = nf, bb.num_scale_filters[i]
    return tot


def bench_backbone(geom, w, host_pts, host_off, B, N, dev, mem_precision, steps=10, warmup=3):
    """points -> spatial_features_2d (front end + backbone in one CUDA graph) and the backbone alone, same batch as the headline."""
    import torch
    from hvpr_b200.pipeline import FrontEndWithBackbone
    nx, ny, _ = geom.grid_size
    torch.manual_seed(0)
    pipe = FrontEndWithBackbone(geom, device=dev, mem_precision=mem_precision)
    pipe.frontend.load_reference_weights(w)
    p = pipe.plan(B, B * N, N)
    p.points.copy_(host_pts)
    p.frame_offsets.copy_(host_off)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        pipe.run()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        pipe.run()
    e1.record()
    torch.cuda.synchronize()
    ms_pipe = e0.elapsed_time(e1) / steps
    bb = pipe.backbone_2d
    t = 0.0
    for i in range(steps + warmup):
        flush.zero_()                       # activations (>= 439 MB per layer) already exceed L2; flush anyway
        e0.record()
        bb.run_nhwc(p.x_nhwc, p.y_nhwc, B, ny, nx)
        e1.record()
        torch.cuda.synchronize()
        if i >= warmup:
            t += e0.elapsed_time(e1) / steps
    fl = backbone_flops(bb, B, ny, nx)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = peaks.get("bf16_tflops_sustained", 1376.4)
    out = {"component": "BaseBEVBackbone_Scale (hvpr.yaml:87-95), bf16 tcgen05 implicit GEMM, fp32 accumulate",
           "ms_per_batch": t, "frames_per_sec": B / (t * 1e-3), "gflop_per_batch": fl / 1e9,
           "roofline": {"bound": "tensor", "achieved": fl / (t * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                        "frac": fl / (t * 1e-3) / 1e12 / peak, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained"},
           "gpu_launches_per_batch": pipe.kernel_launches_per_run() - 9,
           "points_to_spatial_features_2d": {"ms_per_batch": ms_pipe, "frames_per_sec": B / (ms_pipe * 1e-3),
                                             "gpu_launches_per_batch": pipe.kernel_launches_per_run()}}
    # row N2 on top: points -> boxes (dense head fed with channels-last features; the fp32 NCHW feature map is never written)
    try:
        from hvpr_b200.pipeline import HVPR_HEAD_CFG, HVPR_POST_CFG
        pipe2 = FrontEndWithBackbone(geom, device=dev, mem_precision=mem_precision, head_cfg=HVPR_HEAD_CFG, post_cfg=HVPR_POST_CFG)
        pipe2.frontend.load_reference_weights(w)
        p2 = pipe2.plan(B, B * N, N)
        p2.points.copy_(host_pts)
        p2.frame_offsets.copy_(host_off)
        pipe2.run()
        torch.cuda.synchronize()
        # random-init logits never reach the 0.1 threshold (conv_cls.bias = -4.6): shift the bias so that ~1 % of the anchors pass,
        # i.e. NMS_PRE_MAXSIZE (4096) is saturated in every frame - the worst case for the NMS kernels
        with torch.no_grad():
            q99 = torch.quantile(p2.cls_preds.flatten()[:: 16].float(), 0.99)
            pipe2.dense_head.conv_cls.bias += float(-2.1972246 - q99)        # logit(0.1) = -2.197
        p2.graph = None
        for _ in range(warmup):
            pipe2.run()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            pipe2.run()
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / steps
        th = 0.0
        for i in range(steps + warmup):
            e0.record()
            pipe2.dense_head.run_nhwc(p2.f2d_nhwc, B, ny, nx)
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                th += e0.elapsed_time(e1) / steps
        tn = 0.0
        for i in range(steps + warmup):
            e0.record()
            pipe2.post.run(p2.cls_preds, p2.box_preds)
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                tn += e0.elapsed_time(e1) / steps
        out["dense_head"] = {"component": "AnchorHeadSingle eval (one tcgen05 GEMM for conv_cls/box/dir + decode), %d anchors per frame"
                                          % (ny * nx * pipe2.dense_head.num_anchors_per_location),
                             "ms_per_batch": th}
        out["post_processing"] = {"component": "score threshold 0.1 -> top-4096 -> rotated-BEV NMS 0.1 -> 500; head bias calibrated so that ~1 % of "
                                               "the 428 544 anchors per frame pass the threshold (NMS_PRE_MAXSIZE saturated)", "ms_per_batch": tn,
                                  "candidates_after_nms_per_frame": [int(v) for v in p2.det["count"].cpu().tolist()]}
        out["points_to_detections"] = {"ms_per_batch": ms2, "frames_per_sec": B / (ms2 * 1e-3),
                                       "gpu_launches_per_batch": pipe2.kernel_launches_per_run()}
        del pipe2, p2
    except Exception as e:
        out["dense_head"] = {"error": repr(e)[:200]}
    # library baseline on the same box: the same network through torch's own layers (cuDNN, bf16 channels_last, eager).
    # The module's nn.Sequential parameter containers are ordinary torch layers, so they can simply be called.
    # accuracy of the bf16-operand backbone at FULL size against the same network in fp32 (torch layers, TF32 off) on the same
    # (bf16-quantised) canvases, one frame: north_star gives no tolerance for this row; the tests state 2e-2 max-norm / 1e-2 L2
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False

        def torch_forward(m, x, y):
            ups, a = [], m.attention.spatial
            for i in range(len(m.blocks)):
                x = m.blocks[i](x)
                y = m.scale_layers[i](y)
                gate = torch.sigmoid(a.norm(a.conv(torch.cat((y.max(1, keepdim=True)[0], y.mean(1, keepdim=True)), 1))))
                xa = x
                for _ in range(m.sfm_layer_nums[i]):
                    xa = gate * m.sfmblocks_down[i](xa) + xa
                ups.append(m.deblocks[i](xa))
            return torch.cat(ups, 1)
        with torch.no_grad():
            ours = bb.run_nhwc(p.x_nhwc[:1].contiguous(), p.y_nhwc[:1].contiguous(), 1, ny, nx).float().clone()
            ref32 = torch_forward(bb, p.x_nhwc[:1, ..., :128].permute(0, 3, 1, 2).float().contiguous(),
                                  p.y_nhwc[:1, ..., :32].permute(0, 3, 1, 2).float().contiguous())
            d = (ours - ref32).double()
            out["accuracy_vs_fp32"] = {"what": "spatial_features_2d of one G2-size frame: bf16-operand tcgen05 backbone vs the same torch layers in fp32 "
                                               "(cuDNN, TF32 off) on identical canvases", "max_abs_err_over_max_abs_ref": float(d.abs().max() / ref32.abs().max()),
                                       "rel_l2": float(d.norm() / ref32.double().norm()), "stated_tolerance": {"max_norm": 2e-2, "rel_l2": 1e-2}}
            del ours, ref32, d
    except Exception as e:
        out["accuracy_vs_fp32"] = {"error": repr(e)[:200]}
    try:
        torch.backends.cudnn.benchmark = True
        mb = bb.to(memory_format=torch.channels_last).bfloat16()
        xc = p.x_nhwc[..., :128].permute(0, 3, 1, 2)
        yc = p.y_nhwc[..., :32].permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)

        def cudnn_forward():
            x, y, ups, a = xc, yc, [], mb.attention.spatial
            for i in range(len(mb.blocks)):
                x = mb.blocks[i](x)
                y = mb.scale_layers[i](y)
                gate = torch.sigmoid(a.norm(a.conv(torch.cat((y.max(1, keepdim=True)[0], y.mean(1, keepdim=True)), 1))))
                xa = x
                for _ in range(mb.sfm_layer_nums[i]):
                    xa = gate * mb.sfmblocks_down[i](xa) + xa
                ups.append(mb.deblocks[i](xa))
            return torch.cat(ups, 1)
        with torch.no_grad():
            for _ in range(2):
                cudnn_forward()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                cudnn_forward()
            e1.record()
            torch.cuda.synchronize()
        tc = e0.elapsed_time(e1) / 5
        out["library_baseline"] = {"what": "same network via cuDNN (torch eager, bf16 channels_last, cudnn.benchmark), bf16 output",
                                   "ms_per_batch": tc, "tflops": fl / (tc * 1e-3) / 1e12, "speedup": tc / t}
    except Exception as e:
        out["library_baseline"] = {"error": repr(e)[:200]}
    del pipe, p, flush
    torch.cuda.empty_cache()
    return out


def library_baseline(geom, vox, P_total, B, w, dev, reps=5):
    """SURVEY §8d / BASELINE.md §3.5: the reference's OWN formulation of the stages run by PyTorch's library kernels (cuBLAS SGEMM,
    ATen elementwise / reduce / topk / index_put) on the same B200, same batch, CUDA-event timed — the number each hand-written
    kernel has to beat.  A plain torch restatement of pillar_vfe.py:184-221, memory_module.py:60-77 and
    pointpillar_scatter.py:169-220 (fp32, TF32 off); the voxelizer has no library counterpart (spconv is CPU code)."""
    import torch
    import torch.nn.functional as F
    nx, ny, _ = geom.grid_size
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    wd = {k: v.to(dev) for k, v in w.items()}
    voxels, coords, nump = vox.voxels[:P_total], vox.coords[:P_total], vox.num_points[:P_total]
    vs, rng = [float(v) for v in geom.voxel_size], [float(v) for v in geom.point_cloud_range]
    offs = [vs[i] / 2 + rng[i] for i in range(3)]

    def bn(x, prefix):
        return F.batch_norm(x, wd[prefix + ".running_mean"], wd[prefix + ".running_var"], wd[prefix + ".weight"], wd[prefix + ".bias"],
                            False, 0.0, 1e-3)

    def pfn_layer(x, i, last):
        x = F.linear(x, wd["vfe.pfn_layers.%d.linear.weight" % i])
        x = bn(x.permute(0, 2, 1), "vfe.pfn_layers.%d.norm" % i).permute(0, 2, 1)
        x = F.relu(x)
        xm = torch.max(x, dim=1, keepdim=True)[0]
        return xm if last else torch.cat([x, xm.repeat(1, x.shape[1], 1)], dim=2)

    def vfe():
        n = nump.to(voxels.dtype).view(-1, 1, 1)
        mean = voxels[:, :, :3].sum(dim=1, keepdim=True) / n
        f_cluster = voxels[:, :, :3] - mean
        c = coords.to(voxels.dtype)
        f_center = torch.stack([voxels[:, :, 0] - (c[:, 3:4] * vs[0] + offs[0]), voxels[:, :, 1] - (c[:, 2:3] * vs[1] + offs[1]),
                                voxels[:, :, 2] - (c[:, 1:2] * vs[2] + offs[2])], dim=2)
        feats = torch.cat([voxels, f_cluster, f_center], dim=-1)
        mask = (torch.arange(voxels.shape[1], device=dev).view(1, -1) < nump.view(-1, 1)).unsqueeze(-1).to(voxels.dtype)
        feats = feats * mask
        x = pfn_layer(pfn_layer(feats, 0, False), 1, True).squeeze(1)
        sc = torch.cat([n.view(-1, 1), mean.norm(dim=2), mean.squeeze(1)], dim=1)
        for i in range(2):
            sc = F.relu(bn(F.linear(sc, wd["vfe.pfn_scale_layers.%d.0.weight" % i]), "vfe.pfn_scale_layers.%d.1" % i))
        return x, sc

    W = wd["map_to_bev_module.memory.weight"]

    def memory(pil):
        score = F.softmax(F.linear(pil, W), dim=1)
        _, idx = torch.topk(score, 20, dim=1)
        mem = W[idx]
        agg = F.softmax((mem * pil.unsqueeze(1)).sum(dim=2), dim=1)
        return (agg.unsqueeze(2) * mem).sum(dim=1)

    def mem_and_scatter(pf, psf):
        sp_l, sc_l = [], []
        for b in range(B):
            m = coords[:, 0] == b
            tc = coords[m]
            idx = (tc[:, 1] + tc[:, 2] * nx + tc[:, 3]).long()
            pil = pf[m]
            out = memory(pil)
            canvas = torch.zeros(128, nx * ny, device=dev)
            canvas_s = torch.zeros(32, nx * ny, device=dev)
            canvas[:, idx] = torch.cat((pil.t(), out.t()), dim=0)
            canvas_s[:, idx] = psf[m].t()
            sp_l.append(canvas); sc_l.append(canvas_s)
        return torch.stack(sp_l, 0).view(B, 128, ny, nx), torch.stack(sc_l, 0).view(B, 32, ny, nx)

    def mem_only(pf):
        return [memory(pf[coords[:, 0] == b]) for b in range(B)]

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t = {"pfn": 0.0, "mem_attn+bev_fill": 0.0, "mem_attn": 0.0}
    with torch.no_grad():
        for it in range(reps + 2):
            ev[0].record(); pf, psf = vfe(); ev[1].record(); sp, sps = mem_and_scatter(pf, psf); ev[2].record(); mem_only(pf); ev[3].record()
            torch.cuda.synchronize()
            if it >= 2:
                t["pfn"] += ev[0].elapsed_time(ev[1]) / reps
                t["mem_attn+bev_fill"] += ev[1].elapsed_time(ev[2]) / reps
                t["mem_attn"] += ev[2].elapsed_time(ev[3]) / reps
            del sp, sps
    t["bev_fill"] = max(t["mem_attn+bev_fill"] - t["mem_attn"], 0.0)
    out = {"what": "the reference's formulation of PFN / memory attention / scatter through PyTorch's library kernels (cuBLAS + ATen, fp32, "
                   "TF32 off, eager) on this GPU, same %d-frame batch, device-resident, CUDA events; voxelization excluded (no library path)" % B,
           "ms": t, "ms_pfn_mem_bev": t["pfn"] + t["mem_attn+bev_fill"],
           "frames_per_sec_excluding_voxelize": B / ((t["pfn"] + t["mem_attn+bev_fill"]) * 1e-3)}
    del wd
    torch.cuda.empty_cache()
    return out


def run_gpu_arm(args):
    import torch
    from hvpr_b200 import sharding
    from hvpr_b200.frontend import HybridFrontEnd
    from hvpr_b200 import synth

    rank, local_rank, world = sharding.dist_env()
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    numa = bind_numa(local_rank, world)                  # before the first pinned allocation
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        dist = sharding.init_process_group("nccl", device_id=dev)

    geom, ids, N, frames_total = workload(args, rank, world)
    nx, ny, _ = geom.grid_size
    B = len(ids)
    frames = make_frames(geom, ids, N, args.dist)
    w = synth.random_frontend_weights(0)     # synthetic weights under the reference's state_dict names
    fe = HybridFrontEnd(geom, mem_precision=args.mem_precision, device=dev).load_reference_weights(w)
    p = fe.plan(B, B * N, N, use_graph=not args.no_graph)
    host_pts = pinned_like(np.concatenate(frames, 0), args.host_alloc == "wc")
    host_off = torch.tensor(np.r_[0, np.cumsum([len(f) for f in frames])], dtype=torch.int32).pin_memory()
    host_cnt = torch.zeros(B + 1, dtype=torch.int32).pin_memory()
    p.points.copy_(host_pts)
    p.frame_offsets.copy_(host_off)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.current_stream()
    W_ = max(args.warmup, 3)
    K = args.steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(step_fn, finish=None, n=None, per_step=False):
        """n (default: exactly K) steps between barriers -> total ms.  per_step=True also records one event per step and returns
        (total, median, p95) of the step time; the contract numbers are taken WITHOUT the per-step events."""
        n = n or K
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)] if per_step else None
        barrier()
        e0.record(stream)
        for i in range(n):
            step_fn()
            if per_step:
                marks[i].record(stream)
        if finish is not None:
            finish()
        e1.record(stream)
        barrier()
        total = e0.elapsed_time(e1)
        if not per_step:
            return total
        per = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(n - 1))
        return total, per[len(per) // 2], per[int(0.95 * (len(per) - 1))]

    # ---- reference point: one batch at a time, single stream (graph replay of the 8-kernel chain) ---------------------
    for _ in range(W_):
        fe.run()
    ms_serial = sharding.max_over_ranks(timed(fe.run), dev) / K

    # ---- device-resident throughput (`value`): streaming mode, inputs of both slots already in HBM ---------------------
    sp = fe.plan_stream(B, B * N, N)
    for sl in range(len(sp.in_points)):
        sp.in_points[sl].copy_(host_pts)
        sp.in_offsets[sl].copy_(host_off)
    fe.stream_prime()
    for _ in range(W_):
        fe.stream_step()
    barrier()
    with ClockSampler(local_rank) as clk:
        t_dev = timed(fe.stream_step)
        n_stats = max(K, 100)                    # SURVEY §8d asks for >= 100 timed iterations and the median: a second region, one event per step
        _, med_dev, p95_dev = timed(fe.stream_step, n=n_stats, per_step=True)
    ms_total = sharding.max_over_ranks(t_dev, dev)
    ms_step = ms_total / K
    med_dev = sharding.max_over_ranks(med_dev, dev)
    value = frames_total * K / (ms_total * 1e-3)

    # ---- end-to-end through the public call with HOST buffers (`e2e`): H2D of every batch inside the timed region -------
    fe.stream_prime((host_pts, host_off), (host_pts, host_off))
    for _ in range(max(3, W_ // 2)):
        fe.stream_step(host_pts, host_off, host_cnt)
    t_e2e = timed(lambda: fe.stream_step(host_pts, host_off, host_cnt), finish=fe.stream_wait_outputs)
    ms_e2e = sharding.max_over_ranks(t_e2e, dev)
    e2e_value = frames_total * K / (ms_e2e * 1e-3)
    h2d = host_pts.numel() * 4 + host_off.numel() * 4
    d2h = host_cnt.numel() * 4
    P_total = int(host_cnt[-1])
    h2d_gbs_ranks = sharding.gather_scalars(h2d / (t_e2e / K * 1e-3) / 1e9, dev)

    peak, peak_src = measured_peaks()
    peaks_json = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    line = {
        "metric": METRIC if N == POINTS_PER_FRAME else "vfe_frames_per_sec_%dk_pts" % (N // 1000), "value": value, "unit": UNIT, "n_gpus": world,
        "steps": K, "warmup": W_, "ms_per_step": ms_step, "higher_is_better": True, "scaling": WORKLOADS[args.config]["scaling"],
        "vs_baseline": None,
        "dtype": "f32" if args.mem_precision == "fp32" else "f32 (bf16 tensor-core candidate GEMM in memory attention)",
        "data": "synthetic", "config": workload_config(geom, args, frames_total, world),
        "clocks": clk.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / K,
                "h2d_gbs_per_rank": [round(v, 2) for v in h2d_gbs_ranks], "host_alloc": args.host_alloc,
                "note": "host pinned points -> H2D (copy stream, overlapped with the previous batch) -> kernel chain -> D2H of "
                        "per-frame pillar offsets; BEV canvases stay in HBM for the 2D backbone, as in the reference "
                        "(base_bev_backbone.py:281-282); h2d_gbs_per_rank = input bytes / step time (the link is shared with nothing else)"},
        "gpu_launches": fe.kernel_launches_per_run() * K,
        "mem_precision": args.mem_precision,
        "ms_per_step_median": med_dev, "ms_per_step_p95": p95_dev, "value_from_median_step": frames_total / (med_dev * 1e-3),
        "median_over_steps": n_stats,
        "latency_ms_streaming": 3 * ms_step,
        "ms_per_step_single_stream": ms_serial,
        "value_single_stream": frames_total / (ms_serial * 1e-3),
        "numa": numa,
    }

    if rank == 0:
        # ---- per-kernel timing (eager, CUDA events between stages, same stream) -> roofline --------------------
        per = {"voxelize": [], "pfn": [], "mem_attn": [], "bev_fill": []}
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        vox = p.vox
        nP = vox.n_pillars_dev
        reps = min(max(K, 20), 100)
        from hvpr_b200 import _lib
        for it in range(reps + 3):
            evs[0].record(stream)
            fe.voxelizer.run(p.points, p.frame_offsets, B, N, out=vox)
            evs[1].record(stream)
            fe.vfe.run(vox.voxels, vox.num_points, vox.coords, nP, out=p.pillar_features, scale_out=p.pillar_scale)
            evs[2].record(stream)
            fe.map_to_bev_module.memory.run(p.pillar_features, 20, nP, out=p.readout)
            evs[3].record(stream)
            _lib.check(_lib.lib().hvpr_bev_fill(_lib.ptr(p.pillar_features), 64, _lib.ptr(p.readout), 64,
                                                _lib.ptr(p.pillar_scale), 32, _lib.ptr(vox.cell_map), B, nx, ny,
                                                _lib.ptr(p.spatial), _lib.ptr(p.spatial_scale), None, _lib.cur_stream()))
            evs[4].record(stream)
            torch.cuda.synchronize()
            if it >= 3:
                for i, k in enumerate(per):
                    per[k].append(evs[i].elapsed_time(evs[i + 1]))
        per = {k: statistics.median(v) for k, v in per.items()}
        K_total = int(vox.num_points[:P_total].sum())
        alg = algorithmic_bytes(N, P_total / B, K_total / B, nx, ny)
        kern = {}
        for k in per:
            gbs = alg[k] * B / (per[k] * 1e-3) / 1e9
            kern[k] = {"ms": per[k], "alg_bytes": int(alg[k] * B), "gbs": gbs, "frac_hbm": gbs / peak, "frac_hbm_nominal": gbs / NOMINAL_HBM_GBS}
        flops = 256000.0 * P_total
        tf = flops / (per["mem_attn"] * 1e-3) / 1e12
        bf16_peak = peaks_json.get("bf16_tflops", 1590.0)
        kern["mem_attn"].update(gflop=flops / 1e9, tflops=tf, frac_tensor=tf / bf16_peak, frac_tensor_nominal=tf / NOMINAL_BF16_TFLOPS)
        line["kernels"] = kern
        line["kernels_note"] = "median of %d eager launches per stage, CUDA events on the launching stream" % reps
        line["pillars_per_frame"] = P_total / B
        line["kept_points_per_frame"] = K_total / B
        dom = max(per, key=lambda k: per[k])
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(dom)
        if dom == "mem_attn" and args.mem_precision != "fp32":
            line["roofline"] = {"bound": "tensor", "kernel": dom, "achieved": tf, "peak": bf16_peak, "unit": "TFLOP/s", "frac": tf / bf16_peak,
                                "frac_nominal": tf / NOMINAL_BF16_TFLOPS, "traffic": traffic,
                                "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)" if peaks_json else "fallback (B200_PROFILING.md)"}
        else:
            line["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbs"], "peak": peak,
                                "unit": "GB/s", "frac": kern[dom]["frac_hbm"], "frac_nominal": kern[dom]["frac_hbm_nominal"],
                                "traffic": traffic, "peak_source": peak_src}
        total_alg = sum(alg.values()) * B
        line["path_roofline"] = {"alg_bytes_per_step": int(total_alg), "achieved_gbs": total_alg / (ms_step * 1e-3) / 1e9,
                                 "frac_hbm": total_alg / (ms_step * 1e-3) / 1e9 / peak,
                                 "frac_hbm_nominal": total_alg / (ms_step * 1e-3) / 1e9 / NOMINAL_HBM_GBS}
        # ---- library baseline: the reference's formulation through cuBLAS / ATen on this GPU -------------------------------
        if world == 1 and not args.no_library_baseline:
            try:
                lb = library_baseline(geom, vox, P_total, B, w, dev)
                lb["speedup_of_hand_written_kernels"] = {"pfn": lb["ms"]["pfn"] / per["pfn"], "mem_attn": lb["ms"]["mem_attn"] / per["mem_attn"],
                                                        "bev_fill": lb["ms"]["bev_fill"] / per["bev_fill"] if lb["ms"]["bev_fill"] > 0 else None,
                                                        "pfn+mem_attn+bev_fill": lb["ms_pfn_mem_bev"] / (per["pfn"] + per["mem_attn"] + per["bev_fill"])}
                line["library_baseline"] = lb
            except Exception as e:
                line["library_baseline"] = {"error": repr(e)[:300]}
        # ---- next row N1: BaseBEVBackbone_Scale on the tcgen05 conv kernel, reported beside the headline ----
        if world == 1 and not args.no_backbone and B <= 8:
            try:
                line["next_rows"] = {"bev_backbone": bench_backbone(geom, w, host_pts, host_off, B, N, dev, args.mem_precision)}
            except Exception as e:   # the headline line must survive a failure of the widened row
                line["next_rows"] = {"bev_backbone": {"error": repr(e)[:300]}}
        # ---- CPU baseline: the oracle port on this box's host cores, bounded sample ---------------------------
        if world == 1 and not args.no_cpu_baseline:
            def gpu_check(fs):
                q = fe.plan(len(fs), sum(len(f) for f in fs), max(len(f) for f in fs), use_graph=False)
                q.points.copy_(torch.from_numpy(np.ascontiguousarray(np.concatenate(fs, 0))))
                q.frame_offsets.copy_(torch.tensor(np.r_[0, np.cumsum([len(f) for f in fs])], dtype=torch.int32))
                fe.run(); torch.cuda.synchronize()
                Pq = int(q.vox.voxel_offsets[-1])
                return {"coords": q.vox.coords[:Pq], "num_points": q.vox.num_points[:Pq], "pillar_features": q.pillar_features[:Pq],
                        "readout": q.readout[:Pq], "spatial_scale": q.spatial_scale}
            r = cpu_reference_run(geom, frames, w, steps=3, warmup=1, frames_per_step=2, components=True, gpu_check=gpu_check)
            line["cpu_baseline"] = {"value": r["fps"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                    "sample": "3 steps x 2 frames of the same batch (C voxelizer 1 thread/frame, 4 workers + "
                                              "torch CPU VFE/memory/scatter on all cores)",
                                    "what": "oracle/ port: the reference's own torch CPU ops (pinned bit-exactly on its modules where the reference "
                                            "tree exists) + a C restatement of the spconv voxel loop; the reference ships no compilable source",
                                    "components": r.get("components")}
            line["parity_sample"] = r.get("parity_sample")
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def cfg5_cpu_frame(geom, frame, w, wb, wh):
    """The reference's CPU path of the full detector for ONE frame, as far as a CPU can run it: oracle front end + oracle
    BaseBEVBackbone_Scale + oracle AnchorHeadSingle (decode included).  The rotated NMS is NOT part of it: the reference calls
    a CUDA-only op (iou3d_nms `nms_gpu`, setup.py:53-62, source absent) and has no CPU implementation of that step."""
    import torch
    from oracle import backbone as ob, dense_head as od, hybrid
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    t0 = time.perf_counter()
    o = hybrid.frontend([frame], geom, w)
    t1 = time.perf_counter()
    f2d = ob.backbone_forward(wb, o["spatial_features"].numpy(), o["spatial_scale_features"].numpy())
    t2 = time.perf_counter()
    od.head_forward(wh, f2d, od.HEAD_CFG, geom.grid_size, list(geom.point_cloud_range))
    t3 = time.perf_counter()
    return {"front_end_s": t1 - t0, "backbone_s": t2 - t1, "head_s": t3 - t2, "total_s": t3 - t0}


def run_cfg5(args):
    """BASELINE.json configs[4]: points -> detections on N GPUs (frames sharded frame-wise, 8 per GPU, no collective), next to the
    reference's CPU path for the same detector on the host cores (bounded sample: one frame; NMS excluded, see cfg5_cpu_frame)."""
    from hvpr_b200 import sharding
    rank, local_rank, world = sharding.dist_env()
    geom, ids, N, frames_total = workload(args, rank if args.impl == "ours" else 0, world if args.impl == "ours" else 1)
    metric = "full_inference_frames_per_sec_120k_pts"
    if args.impl == "reference":
        if rank != 0:
            return 0
        from oracle import backbone as ob, dense_head as od, hybrid
        frames = make_frames(geom, ids[:2], N, args.dist)
        w, wb, wh = hybrid.random_weights(0), ob.random_backbone_weights(1), od.random_head_weights(3)
        n = max(1, min(args.steps, 3))
        ts = [cfg5_cpu_frame(geom, frames[i % len(frames)], w, wb, wh) for i in range(n + 1)][1:]
        tot = sum(t["total_s"] for t in ts) / len(ts)
        cores = len(os.sched_getaffinity(0))
        line = {"impl": "reference", "metric": metric, "value": 1.0 / tot, "unit": UNIT, "n_gpus": args.gpus, "steps": n, "warmup": 1,
                "ms_per_step": tot * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(geom, args, frames_total, 1),
                "cpu_baseline": {"value": 1.0 / tot, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": "one frame per step (front end + 2D backbone + anchor head on the host cores; the rotated NMS is a CUDA-only "
                                           "op in the reference and is not part of its CPU path)",
                                 "components_s": {k: sum(t[k] for t in ts) / len(ts) for k in ts[0]}},
                "e2e": {"value": 1.0 / tot, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0
    import torch
    from hvpr_b200 import synth
    from hvpr_b200.pipeline import HVPR_HEAD_CFG, HVPR_POST_CFG, FrontEndWithBackbone
    numa = bind_numa(local_rank, world)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = sharding.init_process_group("nccl", device_id=dev) if world > 1 else None
    nx, ny, _ = geom.grid_size
    B = len(ids)
    frames = make_frames(geom, ids, N, args.dist)
    torch.manual_seed(0)
    pipe = FrontEndWithBackbone(geom, device=dev, mem_precision=args.mem_precision, head_cfg=HVPR_HEAD_CFG, post_cfg=HVPR_POST_CFG)
    pipe.frontend.load_reference_weights(synth.random_frontend_weights(0))
    p = pipe.plan(B, B * N, N)
    host_pts = torch.from_numpy(np.ascontiguousarray(np.concatenate(frames, 0))).pin_memory()
    host_off = torch.tensor(np.r_[0, np.cumsum([len(f) for f in frames])], dtype=torch.int32).pin_memory()
    p.points.copy_(host_pts); p.frame_offsets.copy_(host_off)
    pipe.run(); torch.cuda.synchronize()
    with torch.no_grad():       # random-init logits never reach the 0.1 threshold: calibrate the bias so that ~1 % of the anchors pass (NMS_PRE_MAXSIZE saturated)
        q99 = torch.quantile(p.cls_preds.flatten()[::16].float(), 0.99)
        pipe.dense_head.conv_cls.bias += float(-2.1972246 - q99)
    host_cnt = torch.zeros(B, dtype=torch.int32).pin_memory()
    host_box = torch.zeros((B, pipe.post.post_max, 7), dtype=torch.float32).pin_memory()
    host_sc = torch.zeros((B, pipe.post.post_max), dtype=torch.float32).pin_memory()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K, W_ = args.steps, max(args.warmup, 3)
    for _ in range(W_):
        pipe.run()
    barrier()
    with ClockSampler(local_rank) as clk:
        barrier(); e0.record(stream)
        for _ in range(K):
            pipe.run()
        e1.record(stream); barrier()
    ms_total = sharding.max_over_ranks(e0.elapsed_time(e1), dev)

    def e2e_step():
        p.points.copy_(host_pts, non_blocking=True); p.frame_offsets.copy_(host_off, non_blocking=True)
        pipe.run()
        host_cnt.copy_(p.det["count"], non_blocking=True); host_box.copy_(p.det["boxes"], non_blocking=True); host_sc.copy_(p.det["scores"], non_blocking=True)
    for _ in range(3):
        e2e_step()
    barrier(); e0.record(stream)
    for _ in range(K):
        e2e_step()
    e1.record(stream); barrier()
    ms_e2e = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    if rank == 0:
        fl = backbone_flops(pipe.backbone_2d, B, ny, nx)
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        peak = peaks.get("bf16_tflops_sustained", 1376.4)
        ms_step = ms_total / K
        line = {"metric": metric, "value": frames_total * K / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16 operands, fp32 accumulate (2-D backbone / head); f32 front end", "data": "synthetic",
                "config": workload_config(geom, args, frames_total, world), "clocks": clk.summary(),
                "e2e": {"value": frames_total * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": host_pts.numel() * 4 + host_off.numel() * 4,
                        "d2h_bytes_per_step": host_cnt.numel() * 4 + host_box.numel() * 4 + host_sc.numel() * 4, "ms_per_step": ms_e2e / K,
                        "note": "pinned points -> H2D -> one CUDA graph (voxelize ... NMS) -> D2H of the detections (boxes, scores, counts)"},
                "gpu_launches": pipe.kernel_launches_per_run() * K,
                "roofline": {"bound": "tensor", "kernel": "conv_tc (2-D backbone, %d of the step's launches)" % (pipe.kernel_launches_per_run() - 15),
                             "achieved": fl / (ms_step * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s", "frac": fl / (ms_step * 1e-3) / 1e12 / peak,
                             "note": "backbone FLOPs over the WHOLE step time (front end, head and NMS included): a lower bound of the conv kernel's own fraction",
                             "traffic": None},
                "detections_per_frame": [int(v) for v in p.det["count"].cpu().tolist()], "numa": numa}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--config", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--dist", default=DIST, choices=["L", "U"], help="synthetic frame distribution (SURVEY.md Appendix A)")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--host-alloc", default="pinned", choices=["pinned", "wc"],
                    help="host staging buffer of the e2e leg: page-locked, or page-locked + write-combined")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mem-precision", default=os.environ.get("HVPR_MEM_PRECISION", "bf16_rescore"), choices=["fp32", "bf16_rescore"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-backbone", action="store_true", help="skip the next-row (N1) backbone measurement")
    args = ap.parse_args()
    if args.config == "cfg5":
        return run_cfg5(args)
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
