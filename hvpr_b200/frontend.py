"""The whole accelerated path as one object: raw LiDAR frames -> BEV pseudo-image, device-resident, CUDA-graph replayed.

    points -> hvpr_voxelize -> hvpr_pfn -> hvpr_mem_attn -> hvpr_bev_fill
           -> spatial_features (B,128,ny,nx), spatial_scale_features (B,32,ny,nx)

This is the composition the reference performs across a DataLoader worker (spconv voxelizer,
pcdet/datasets/processor/data_processor.py:43-75), collate (dataset.py:148-180), H2D, and two nn.Modules
(pillar_vfe.py:184-221, pointpillar_scatter.py:169-220).  Buffers are allocated once for a fixed batch shape, no
host synchronisation happens inside `run()`, and the kernel chain is captured into a CUDA graph on first use.
"""
from __future__ import annotations

import os

import torch

from . import _lib
from .config import HVPR_BEV_CFG, HVPR_VFE_CFG, Cfg
from .geometry import Geometry
from .map_to_bev import PointPillarScatter_Agg_Memory_1_scale
from .vfe import PillarVFE_Scale
from .voxelizer import Voxelizer


class HybridFrontEnd(torch.nn.Module):
    # HvprLaunchCfg (blocks_per_sm, variant) of the PFN inside the streaming graphs: the low-register variant leaves room for the fill blocks
    stream_pfn_knob = (2, 1)
    # HvprLaunchCfg of the canvas fill inside the streaming graphs (persistent blocks per SM; it shares the SMs with K1 / K2 there)
    stream_bev_knob = None   # one block per item (round 2: the persistent form lost once the PFN blocks shrank; tools/dev/stream_probe.py)
    if os.environ.get("HVPR_STREAM_BEV_BPS"):          # dev override for tools/dev probes
        stream_bev_knob = (int(os.environ["HVPR_STREAM_BEV_BPS"]), 0) if int(os.environ["HVPR_STREAM_BEV_BPS"]) else None
    # where K1 of batch k+2 sits in the step: "fork" (own stream from the start of the step), "before_k3" / "after_k3" / "last" (main stream)
    stream_k1_order = "fork"

    def __init__(self, geom: Geometry, vfe_cfg: Cfg = HVPR_VFE_CFG, bev_cfg: Cfg = HVPR_BEV_CFG,
                 overflow: str = "continue", mem_precision: str = "bf16_rescore", device="cuda"):
        super().__init__()
        self.geom = geom
        self.dev = torch.device(device)
        self.vfe = PillarVFE_Scale(vfe_cfg, 4, list(geom.voxel_size), geom.range_f32)
        self.map_to_bev_module = PointPillarScatter_Agg_Memory_1_scale(bev_cfg, grid_size=geom.grid_size)
        self.map_to_bev_module.memory.precision = mem_precision
        self.voxelizer = Voxelizer(geom, overflow, self.dev)
        self.to(self.dev)
        self.eval()
        self._plan = None

    def load_reference_weights(self, w: dict):
        """w: reference state_dict names ('vfe.*', 'map_to_bev_module.memory.weight')."""
        sd = {k: v for k, v in w.items() if k.startswith("vfe.") or k.startswith("map_to_bev_module.")}
        missing, unexpected = self.load_state_dict(sd, strict=False)
        missing = [m for m in missing if "num_batches_tracked" not in m]
        assert not missing and not unexpected, (missing, unexpected)
        return self

    def weights_version(self):
        """(data_ptr, version) of every parameter / buffer: a captured graph bakes the folded PFN weights in by value and reads the
        packed memory image, so a load_state_dict() after the first run must invalidate it (checked at every run / stream_step)."""
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    # ---------------------------------------------------------------------------------------------------------
    class _Plan:
        pass

    def plan(self, n_frames: int, n_total_points: int, max_frame_points: int = 0, use_graph: bool = True):
        """Allocate every buffer for a fixed batch shape (sized for 180 GB HBM: capacity rows = B * max_voxels)."""
        g, dev = self.geom, self.dev
        nx, ny, _ = g.grid_size
        p = self._Plan()
        p.B, p.n_total, p.max_frame_points = n_frames, n_total_points, max_frame_points
        p.points = torch.empty((n_total_points, 4), dtype=torch.float32, device=dev)
        p.frame_offsets = torch.zeros((n_frames + 1,), dtype=torch.int32, device=dev)
        p.vox = self.voxelizer.alloc_output(n_frames)
        rows = p.vox.max_rows
        p.pillar_features = torch.empty((rows, 64), dtype=torch.float32, device=dev)
        p.pillar_scale = torch.empty((rows, 32), dtype=torch.float32, device=dev)
        p.readout = torch.empty((rows, 64), dtype=torch.float32, device=dev)
        p.spatial = torch.empty((n_frames, 128, ny, nx), dtype=torch.float32, device=dev)
        p.spatial_scale = torch.empty((n_frames, 32, ny, nx), dtype=torch.float32, device=dev)
        p.graph = None
        p.use_graph = use_graph
        self._plan = p
        return p

    def _enqueue(self, p):
        vox = self.voxelizer.run(p.points, p.frame_offsets, p.B, p.max_frame_points, out=p.vox)
        nP = vox.n_pillars_dev
        self.vfe.run(vox.voxels, vox.num_points, vox.coords, nP, out=p.pillar_features, scale_out=p.pillar_scale)
        self.map_to_bev_module.run(p.pillar_features, p.pillar_scale, vox.cell_map, p.B, nP, readout=p.readout,
                                   spatial=p.spatial, spatial_scale=p.spatial_scale)

    def kernel_launches_per_run(self) -> int:
        # init, hash, assign, fill, gather | pfn | mem_attn | bev_fill
        return 8

    @torch.no_grad()
    def run(self):
        """One pass over the planned batch whose points / frame_offsets are already resident in p.points."""
        p = self._plan
        _lib.init_device()
        if not p.use_graph:
            self._enqueue(p)
            return p
        wkey = self.weights_version()
        if p.graph is not None and p.graph_wkey != wkey:
            p.graph = None                                # weights changed since capture: fold / pack / capture again
        if p.graph is None:
            p.graph_wkey = wkey
            self.vfe._weights_packed(self.dev)            # host-side folding + fragment packing happen outside capture
            if self.map_to_bev_module.memory.precision == "bf16_rescore":
                self.map_to_bev_module.memory._packed_bf16()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._enqueue(p)                        # warm-up (workspace allocation, lazy init)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                self._enqueue(p)
            p.graph = gr
        p.graph.replay()
        return p

    @torch.no_grad()
    def run_host(self, points_pinned: torch.Tensor, frame_offsets_pinned: torch.Tensor, counts_pinned: torch.Tensor):
        """End-to-end call with HOST buffers: H2D of the frame batch, the kernel chain, D2H of the per-frame pillar
        offsets (the data-dependent sizes a caller needs to shape `voxels`); all asynchronous on the current stream."""
        p = self._plan
        p.points.copy_(points_pinned, non_blocking=True)
        p.frame_offsets.copy_(frame_offsets_pinned, non_blocking=True)
        self.run()
        counts_pinned.copy_(p.vox.voxel_offsets, non_blocking=True)
        return p


    # ---------------------------------------------------------------------------------------------------------
    # Streaming mode: consecutive batches are software-pipelined, three stages deep, across CUDA streams.
    #   step k (one CUDA graph):   main stream    K3 memory attention -> K4 BEV fill      of batch k
    #                              side stream 1  K2 PFN                                   of batch k+1
    #                              side stream 2  K1 voxelize                              of batch k+2
    #   copy stream (eager):       H2D of batch k+2's points while the previous graph is still running
    # K1 is latency-bound (five small launches), K2 is latency/issue-bound and K4 is HBM-write-bound with spare SM resources,
    # so once the persistent K3 (which occupies whole SMs) retires, the three run concurrently.  Every batch still gets
    # exactly the same kernels on the same data: results are bit-identical to `run()`; they appear two steps after the
    # batch was submitted.
    _NS = 3

    def plan_stream(self, n_frames: int, n_total_points: int, max_frame_points: int = 0):
        g, dev = self.geom, self.dev
        nx, ny, _ = g.grid_size
        NS = self._NS
        p = self._Plan()
        p.B, p.n_total, p.max_frame_points = n_frames, n_total_points, max_frame_points
        p.in_points = [torch.empty((n_total_points, 4), dtype=torch.float32, device=dev) for _ in range(NS)]
        p.in_offsets = [torch.zeros((n_frames + 1,), dtype=torch.int32, device=dev) for _ in range(NS)]
        p.voxs = [self.voxelizer.alloc_output(n_frames) for _ in range(NS)]
        rows = p.voxs[0].max_rows
        p.pfs = [torch.empty((rows, 64), dtype=torch.float32, device=dev) for _ in range(NS)]
        p.pss = [torch.empty((rows, 32), dtype=torch.float32, device=dev) for _ in range(NS)]
        p.readout = torch.empty((rows, 64), dtype=torch.float32, device=dev)
        p.spatial = torch.empty((n_frames, 128, ny, nx), dtype=torch.float32, device=dev)
        p.spatial_scale = torch.empty((n_frames, 32, ny, nx), dtype=torch.float32, device=dev)
        # side1 (PFN of the next batch) gets dispatch priority over the canvas fill it shares the SMs with
        p.side1 = torch.cuda.Stream(device=dev, priority=int(os.environ.get("HVPR_STREAM_PFN_PRIO", "-1")))
        p.side2, p.copy = torch.cuda.Stream(device=dev, priority=int(os.environ.get("HVPR_STREAM_K1_PRIO", "0"))), torch.cuda.Stream(device=dev)
        p.ev_copied = [torch.cuda.Event() for _ in range(NS)]
        p.ev_input_free = [torch.cuda.Event() for _ in range(NS)]
        # per-frame pillar offsets of the finished batch: staged on the device inside the graph, read back on their own stream so
        # that the 36-byte D2H never sits between two graph launches on the main stream (it cost 0.046 ms per step there)
        p.cnt_stage = [torch.zeros((n_frames + 1,), dtype=torch.int32, device=dev) for _ in range(NS)]
        p.copy_out = torch.cuda.Stream(device=dev)
        p.ev_graph_done = [torch.cuda.Event() for _ in range(NS)]
        p.ev_cnt_done = [torch.cuda.Event() for _ in range(NS)]
        p.cnt_pending = [False] * NS
        p.graphs = [None] * NS
        p.k = 0              # slot of the batch finished by the next step
        p.primed = False
        p.vox, p.pillar_features, p.pillar_scale = p.voxs[0], p.pfs[0], p.pss[0]
        self._splan = p
        return p

    def _stage_vox(self, p, slot):      # K1
        self.voxelizer.run(p.in_points[slot], p.in_offsets[slot], p.B, p.max_frame_points, out=p.voxs[slot])

    def _stage_pfn(self, p, slot, launch=None):      # K2
        vox = p.voxs[slot]
        self.vfe.run(vox.voxels, vox.num_points, vox.coords, vox.n_pillars_dev, out=p.pfs[slot], scale_out=p.pss[slot], launch=launch)

    def _stage_bev(self, p, slot, after_k3=None, launch=None):      # K3 + K4
        vox = p.voxs[slot]
        m = self.map_to_bev_module
        fused = m.fused_zero_fill                       # K3 zeroes the canvases beside its own work, K4 writes only occupied runs
        m.memory.run(p.pfs[slot], m.k, vox.n_pillars_dev, out=p.readout, zero_fill=[p.spatial, p.spatial_scale] if fused else None)
        if after_k3 is not None:
            after_k3()
        if fused:
            launch = (launch[0] if launch else 0, fused)
        st = _lib.lib().hvpr_bev_fill(_lib.ptr(p.pfs[slot]), 64, _lib.ptr(p.readout), 64, _lib.ptr(p.pss[slot]), 32,
                                      _lib.ptr(vox.cell_map), p.B, m.nx, m.ny, _lib.ptr(p.spatial),
                                      _lib.ptr(p.spatial_scale), _lib.launch_cfg(launch), _lib.cur_stream())
        _lib.check(st, "hvpr_bev_fill")

    @torch.no_grad()
    def stream_prime(self, batch0=None, batch1=None):
        """Fill the pipeline: batch0 -> voxelized + PFN (slot 0), batch1 -> voxelized (slot 1).  Each batch is a
        (points, frame_offsets) pair of device or pinned-host tensors, or None when the caller already filled
        `in_points[slot]` / `in_offsets[slot]`."""
        p = self._splan
        NS = self._NS
        _lib.init_device()
        for slot, b in ((0, batch0), (1, batch1)):
            if b is not None:
                p.in_points[slot].copy_(b[0], non_blocking=True)
                p.in_offsets[slot].copy_(b[1], non_blocking=True)
        self.vfe._weights_packed(self.dev)
        if self.map_to_bev_module.memory.precision == "bf16_rescore":
            self.map_to_bev_module.memory._packed_bf16()
        wkey = self.weights_version()
        if p.graphs[0] is not None and p.graphs_wkey != wkey:
            p.graphs = [None] * NS                        # weights changed since capture
        if p.graphs[0] is None:                         # warm every kernel up outside capture, then capture the NS phases
            p.graphs_wkey = wkey
            for slot in range(NS):
                self._stage_vox(p, slot); self._stage_pfn(p, slot)
            self._stage_bev(p, 0)
            torch.cuda.synchronize()
            # launch shapes travel with the calls (HvprLaunchCfg): the low-register PFN variant leaves room for the canvas-fill
            # blocks it runs beside (+10 % measured); nothing process-wide is touched, so several front ends can capture at once
            for k in range(NS):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    main = torch.cuda.current_stream()
                    order = self.stream_k1_order
                    p.side2.wait_stream(main)           # fork
                    with torch.cuda.stream(p.side2):
                        p.cnt_stage[k].copy_(p.voxs[k].voxel_offsets)      # slot k is re-voxelized by the NEXT step
                        if order == "fork":
                            self._stage_vox(p, (k + 2) % NS)
                    if order == "before_k3":
                        self._stage_vox(p, (k + 2) % NS)

                    def _fork_pfn(k=k):
                        if order == "after_k3":
                            self._stage_vox(p, (k + 2) % NS)
                        # K3 holds whole SMs; the PFN of the next batch starts when it retires and shares the SMs with K4
                        p.side1.wait_stream(main)
                        with torch.cuda.stream(p.side1):
                            self._stage_pfn(p, (k + 1) % NS, launch=self.stream_pfn_knob)
                    self._stage_bev(p, k, after_k3=_fork_pfn, launch=self.stream_bev_knob)
                    main.wait_stream(p.side1)           # join
                    main.wait_stream(p.side2)
                    if order == "last":
                        self._stage_vox(p, (k + 2) % NS)
                p.graphs[k] = gr
        self._stage_vox(p, 0); self._stage_pfn(p, 0)
        self._stage_vox(p, 1)
        p.k, p.primed = 0, True
        p.vox, p.pillar_features, p.pillar_scale = p.voxs[0], p.pfs[0], p.pss[0]
        return p

    @torch.no_grad()
    def stream_step(self, next_points=None, next_offsets=None, counts_out=None):
        """One pipeline step: finish the oldest batch in flight (its canvases are valid once this step completes), run the
        PFN of the next one and voxelize the batch given here (it is finished two steps later).
        next_points / next_offsets: pinned host or device tensors; None = already resident in `in_points[(k + 2) % 3]`.
        counts_out: optional pinned (B+1,) int32 receiving the finished batch's per-frame pillar offsets; the read-back runs on
        its own stream — call stream_wait_outputs() (or synchronize the device) before reading it on the host."""
        p = self._splan
        NS = self._NS
        assert p.primed, "call stream_prime() first"
        k = p.k
        vs = (k + 2) % NS                                        # slot voxelized by this step
        main = torch.cuda.current_stream()
        if next_points is not None:
            with torch.cuda.stream(p.copy):
                p.copy.wait_event(p.ev_input_free[vs])          # the K1 that last read this slot has finished
                p.in_points[vs].copy_(next_points, non_blocking=True)
                p.in_offsets[vs].copy_(next_offsets, non_blocking=True)
                p.ev_copied[vs].record(p.copy)
            main.wait_event(p.ev_copied[vs])
        if p.cnt_pending[k]:                                     # the read-back of this slot's staging buffer, three steps ago
            main.wait_event(p.ev_cnt_done[k]); p.cnt_pending[k] = False
        p.graphs[k].replay()
        p.ev_input_free[vs].record(main)
        if counts_out is not None:
            p.ev_graph_done[k].record(main)
            with torch.cuda.stream(p.copy_out):
                p.copy_out.wait_event(p.ev_graph_done[k])
                counts_out.copy_(p.cnt_stage[k], non_blocking=True)
                p.ev_cnt_done[k].record(p.copy_out)
            p.cnt_pending[k] = True
        p.vox, p.pillar_features, p.pillar_scale = p.voxs[k], p.pfs[k], p.pss[k]   # the batch finished by this step
        p.k = (k + 1) % NS
        return p

    def stream_wait_outputs(self):
        """Make the current stream wait for every pending counts_out read-back of stream_step()."""
        torch.cuda.current_stream().wait_stream(self._splan.copy_out)

    @torch.no_grad()
    def forward(self, batch_dict: dict) -> dict:
        """batch_dict API: consumes 'points' (sum N,5) [b,x,y,z,r] + 'batch_size', fills every key the reference's
        voxelizer + VFE + map_to_bev would (exact-shaped tensors; one 4-byte D2H read for the pillar count)."""
        batch_dict = self.voxelizer.voxelize_batch(batch_dict, exact_shapes=True)
        batch_dict.pop("num_pillars_dev", None)
        batch_dict = self.vfe(batch_dict)
        batch_dict = self.map_to_bev_module(batch_dict)
        return batch_dict
