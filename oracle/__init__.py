"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU restatement of the reference's algorithm for the hybrid voxel-point encoding front end
(voxelize -> PillarVFE[_Scale] -> MemoryUnit_Agg eval -> PointPillarScatter[_Agg_Memory_1_scale] eval).

Who may import / execute this package:  tests/,  __graft_entry__.smoke(),  bench.py's `cpu_baseline`
leg and `bench.py --impl reference`.  Nothing under hvpr_b200/ imports it; the product path fails loudly
when its CUDA extension is missing instead of falling back to anything here.

Pinning status:
  * voxelizer          — loop structure PINNED ON REFERENCE-RUN CODE: the reference's own in-tree numba voxel loop
                         (`_points_to_bevmap_reverse_kernel`, tools/vis.py:8-60 — same lineage as spconv's points_to_voxel)
                         is executed here (oracle/ref_loader.load_vis_voxel_kernel) and oracle/voxelize_ref.c reproduces its
                         coor_to_voxelidx table and per-cell point counts exactly (cell arithmetic, x->y->z reject order,
                         reversed coords, first-seen ids, the `break` cap): tests/test_oracle_cpu.py::
                         test_voxelizer_oracle_vs_live_reference_loop + fixtures tests/golden/vis_kernel_pins.json
                         (oracle/make_golden_vis.py).  Still unpinned against UPSTREAM spconv (not vendored, not pinned, not
                         installed): the `continue` overflow mode and the 32-point append, which vis.py replaces with a height
                         map, rest on the published algorithm + an independent dict-based model.
  * VFE / memory / BEV — pinned against the reference's OWN Python modules imported from /root/reference in the
                         build container (oracle/ref_loader.py, 3 in-memory patches) — see oracle/make_golden.py and
                         tests/golden/*.npz, and tests/test_oracle_cpu.py::test_oracle_vs_live_reference (runs wherever /root/reference exists).
  * 2-D backbone (N1)  — oracle/backbone.py, pinned against the reference's OWN BaseBEVBackbone_Scale
                         (oracle/ref_loader.load_backbone, one in-memory patch: breakage B4) — oracle/make_golden_backbone.py,
                         tests/golden/backbone_tiny.npz.
  * dense head (N2)    — oracle/dense_head.py: assembly restated (the head class is not importable: B5/B6), its decode / anchor /
                         limit_period building blocks pinned bit-exactly on the reference's own functions (tests/test_oracle_cpu.py).
  * post-processing (N3) — oracle/post_process.py: PARITY UNPINNED for the NMS (the reference's iou3d_nms CUDA op is not in the tree);
                         selection logic restated from the Python source, rotated IoU restated independently in float64.
"""
