#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json: VFE frames/sec @120k pts/frame on B200.

One "step" = one pass of the hot path (voxelize -> fused PFN -> memory attention -> BEV fill) over one batch of
synthetic frames: BASELINE.json configs[1] = G2 (pillar 0.16x0.16x4 m, 432x496 grid, 32 pts/pillar, 40k max pillars),
B = 8 frames of 120 000 points per GPU.  N GPUs => N ranks (torchrun), each with its own 8 frames (weak scaling, no
collective on the data path; NCCL only reduces the timings).

    python bench.py [--gpus N] [--steps K] [--warmup W]           our arm
    python bench.py --impl reference ...                           the reference's CPU path (oracle port) on host cores

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


def make_frames(geom, rank, n_frames=FRAMES_PER_GPU):
    from hvpr_b200 import sharding, synth
    ids = sharding.weak_scaling_frames(n_frames, rank)
    return [synth.make_frame(DIST, POINTS_PER_FRAME, geom.point_cloud_range, 1024 + i) for i in ids]


# --------------------------------------------------------------------------------------------------- CPU arms
def cpu_reference_run(geom, frames, w, steps, warmup, frames_per_step, workers=4):
    """The reference's CPU path, as the oracle port: spconv-style voxelizer single-threaded per frame (one frame per
    DataLoader worker, `--workers 4` default, tools/test.py:25) + PillarVFE_Scale + PointPillarScatter_Agg_Memory_1_scale
    on torch CPU with every host thread."""
    import concurrent.futures as cf
    import torch
    from oracle import hybrid
    from oracle import voxelize as ov
    ncores = len(os.sched_getaffinity(0))
    torch.set_num_threads(ncores)
    nx, ny, _ = geom.grid_size
    pool = cf.ThreadPoolExecutor(max_workers=min(workers, ncores))

    def step(i):
        fs = [frames[(i * frames_per_step + j) % len(frames)] for j in range(frames_per_step)]
        outs = list(pool.map(lambda f: ov.voxelize_c(f, geom.range_f32, geom.voxel_f32, geom.max_points_per_voxel,
                                                     geom.max_voxels, "continue"), fs))
        vox = np.concatenate([o[0] for o in outs], 0)
        coords = np.concatenate([np.pad(o[1], ((0, 0), (1, 0)), constant_values=b) for b, o in enumerate(outs)], 0)
        nump = np.concatenate([o[2] for o in outs], 0)
        with torch.no_grad():
            tv, tc, tn = torch.from_numpy(vox), torch.from_numpy(coords), torch.from_numpy(nump)
            pf, psf, _ = hybrid.pillar_vfe(tv, tn, tc, w, list(geom.voxel_size), geom.range_f32)
            sp, sps, _ = hybrid.scatter_agg_memory(pf, psf, tc, w["map_to_bev_module.memory.weight"], len(fs), nx, ny)
        return float(sp[0, 0, 0, 0])

    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    dt = time.perf_counter() - t0
    return frames_per_step * steps / dt, dt / steps * 1e3, ncores


def run_reference_arm(args):
    from hvpr_b200 import sharding
    from hvpr_b200.geometry import G2
    from oracle import hybrid
    rank, _, world = sharding.dist_env()
    if rank != 0:
        return 0
    fps_step = 2
    frames = make_frames(G2, 0)
    w = hybrid.random_weights(0)
    steps, warmup = args.steps, args.warmup
    fps, ms, ncores = cpu_reference_run(G2, frames, w, steps, warmup, fps_step)
    sample = "%d frames/step of the %d-frame batch (G2, %d pts/frame, dist %s)" % (fps_step, FRAMES_PER_GPU, POINTS_PER_FRAME, DIST)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(G2),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": ncores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_config(geom):
    nx, ny, _ = geom.grid_size
    return {"workload": "HVPR KITTI cfg, G2 pillars 0.16x0.16x4 m, %dx%d grid, 32 pts/pillar, 40k max pillars; "
                        "batch of %d synthetic LiDAR-like frames x %d pts per GPU; voxelize+PFN+memory attention+BEV fill"
                        % (nx, ny, FRAMES_PER_GPU, POINTS_PER_FRAME),
            "frames_per_gpu": FRAMES_PER_GPU, "points_per_frame": POINTS_PER_FRAME, "distribution": DIST,
            "weights": "random-init (seed 0), BN stats randomised",
            "l2": "each step streams 1.1 GB of canvas writes (>8x the 126 MB L2), so inputs are evicted between steps",
            "streaming": "batches are software-pipelined 3 deep over CUDA streams: K1 voxelize of batch k+2 and K2 PFN of batch k+1 "
                         "overlap K3/K4 of batch k; every batch runs the same 8 kernels (value_single_stream = no overlap)"}


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


def run_gpu_arm(args):
    import torch
    from hvpr_b200 import sharding
    from hvpr_b200.frontend import HybridFrontEnd
    from hvpr_b200.geometry import G2
    from hvpr_b200 import synth

    rank, local_rank, world = sharding.dist_env()
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        dist = sharding.init_process_group("nccl", device_id=dev)

    geom = G2
    nx, ny, _ = geom.grid_size
    B, N = FRAMES_PER_GPU, POINTS_PER_FRAME
    frames = make_frames(geom, rank)
    w = synth.random_frontend_weights(0)     # synthetic weights under the reference's state_dict names
    fe = HybridFrontEnd(geom, mem_precision=args.mem_precision, device=dev).load_reference_weights(w)
    p = fe.plan(B, B * N, N, use_graph=not args.no_graph)
    host_pts = torch.from_numpy(np.ascontiguousarray(np.concatenate(frames, 0))).pin_memory()
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---- reference point: one batch at a time, single stream (graph replay of the 8-kernel chain) ---------------------
    for _ in range(W_):
        fe.run()
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        fe.run()
    e1.record(stream)
    barrier()
    ms_serial = sharding.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps

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
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            fe.stream_step()
        e1.record(stream)
        barrier()
    ms_total = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- end-to-end through the public call with HOST buffers (`e2e`): H2D of every batch inside the timed region -------
    fe.stream_prime((host_pts, host_off), (host_pts, host_off))
    for _ in range(3):
        fe.stream_step(host_pts, host_off, host_cnt)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        fe.stream_step(host_pts, host_off, host_cnt)
    fe.stream_wait_outputs()                      # the last read-back of the pillar offsets is inside the timed region
    e1.record(stream)
    barrier()
    ms_e2e = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    h2d = host_pts.numel() * 4 + host_off.numel() * 4
    d2h = host_cnt.numel() * 4
    P_total = int(host_cnt[-1])

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.mem_precision == "fp32" else "f32 (bf16 tensor-core candidate GEMM in memory attention)",
        "data": "synthetic", "config": workload_config(geom),
        "clocks": clk.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps,
                "note": "host pinned points -> H2D (copy stream, overlapped with the previous batch) -> kernel chain -> D2H of "
                        "per-frame pillar offsets; BEV canvases stay in HBM for the 2D backbone, as in the reference "
                        "(base_bev_backbone.py:281-282)"},
        "gpu_launches": fe.kernel_launches_per_run() * args.steps,
        "mem_precision": args.mem_precision,
        "ms_per_step_single_stream": ms_serial,
        "value_single_stream": world * B / (ms_serial * 1e-3),
    }

    if rank == 0:
        # ---- per-kernel timing (eager, CUDA events between stages, same stream) -> roofline --------------------
        per = {"voxelize": 0.0, "pfn": 0.0, "mem_attn": 0.0, "bev_fill": 0.0}
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        vox = p.vox
        nP = vox.n_pillars_dev
        reps = min(args.steps, 20)
        for it in range(reps + 2):
            evs[0].record(stream)
            fe.voxelizer.run(p.points, p.frame_offsets, B, N, out=vox)
            evs[1].record(stream)
            fe.vfe.run(vox.voxels, vox.num_points, vox.coords, nP, out=p.pillar_features, scale_out=p.pillar_scale)
            evs[2].record(stream)
            fe.map_to_bev_module.memory.run(p.pillar_features, 20, nP, out=p.readout)
            evs[3].record(stream)
            from hvpr_b200 import _lib
            _lib.check(_lib.lib().hvpr_bev_fill(_lib.ptr(p.pillar_features), 64, _lib.ptr(p.readout), 64,
                                                _lib.ptr(p.pillar_scale), 32, _lib.ptr(vox.cell_map), B, nx, ny,
                                                _lib.ptr(p.spatial), _lib.ptr(p.spatial_scale), None, _lib.cur_stream()))
            evs[4].record(stream)
            torch.cuda.synchronize()
            if it >= 2:
                for i, k in enumerate(per):
                    per[k] += evs[i].elapsed_time(evs[i + 1]) / reps
        K_total = int(vox.num_points[:P_total].sum())
        alg = algorithmic_bytes(N, P_total / B, K_total / B, nx, ny)
        peak, peak_src = measured_peaks()
        kern = {}
        for k in per:
            gbs = alg[k] * B / (per[k] * 1e-3) / 1e9
            kern[k] = {"ms": per[k], "alg_bytes": int(alg[k] * B), "gbs": gbs, "frac_hbm": gbs / peak}
        line["kernels"] = kern
        line["pillars_per_frame"] = P_total / B
        line["kept_points_per_frame"] = K_total / B
        dom = max(per, key=lambda k: per[k])
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(dom)
        if dom == "mem_attn" and args.mem_precision != "fp32":
            flops = 256000.0 * P_total
            line["roofline"] = {"bound": "tensor", "kernel": dom, "achieved": flops / (per[dom] * 1e-3) / 1e12,
                                "peak": json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
                                if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1590.0,
                                "unit": "TFLOP/s", "traffic": traffic}
            line["roofline"]["frac"] = line["roofline"]["achieved"] / line["roofline"]["peak"]
        else:
            line["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbs"], "peak": peak,
                                "unit": "GB/s", "frac": kern[dom]["frac_hbm"], "traffic": traffic, "peak_source": peak_src}
        total_alg = sum(alg.values()) * B
        line["path_roofline"] = {"alg_bytes_per_step": int(total_alg), "achieved_gbs": total_alg / (ms_step * 1e-3) / 1e9,
                                 "frac_hbm": total_alg / (ms_step * 1e-3) / 1e9 / peak}
        # ---- next row N1: BaseBEVBackbone_Scale on the tcgen05 conv kernel, reported beside the headline ----
        if world == 1 and not args.no_backbone:
            try:
                line["next_rows"] = {"bev_backbone": bench_backbone(geom, w, host_pts, host_off, B, N, dev, args.mem_precision)}
            except Exception as e:   # the headline line must survive a failure of the widened row
                line["next_rows"] = {"bev_backbone": {"error": repr(e)[:300]}}
        # ---- CPU baseline: the oracle port on this box's host cores, bounded sample ---------------------------
        if world == 1 and not args.no_cpu_baseline:
            fps, ms, ncores = cpu_reference_run(geom, frames, w, steps=3, warmup=1, frames_per_step=2)
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": ncores, "kind": "port",
                                    "sample": "3 steps x 2 frames of the same 8-frame batch (C voxelizer 1 thread/frame + "
                                              "torch CPU VFE/memory/scatter on all cores)"}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mem-precision", default=os.environ.get("HVPR_MEM_PRECISION", "bf16_rescore"), choices=["fp32", "bf16_rescore"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-backbone", action="store_true", help="skip the next-row (N1) backbone measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
