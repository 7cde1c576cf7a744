"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU restatement of the reference's algorithm for the hybrid voxel-point encoding front end
(voxelize -> PillarVFE[_Scale] -> MemoryUnit_Agg eval -> PointPillarScatter[_Agg_Memory_1_scale] eval).

Who may import / execute this package:  tests/,  __graft_entry__.smoke(),  bench.py's `cpu_baseline`
leg and `bench.py --impl reference`.  Nothing under hvpr_b200/ imports it; the product path fails loudly
when its CUDA extension is missing instead of falling back to anything here.

Pinning status:
  * voxelizer          — PARITY UNPINNED against upstream spconv (not vendored, not pinned, not installed; the
                         reference has no tests).  Pinned against an independent dict-based model and the in-tree
                         loop witness tools/vis.py:23-50.
  * VFE / memory / BEV — pinned against the reference's OWN Python modules imported from /root/reference in the
                         build container (oracle/ref_loader.py, 3 in-memory patches) — see oracle/make_golden.py and
                         tests/golden/*.npz, and tests/test_oracle_vs_reference.py (runs wherever /root/reference exists).
  * 2-D backbone (N1)  — oracle/backbone.py, pinned against the reference's OWN BaseBEVBackbone_Scale
                         (oracle/ref_loader.load_backbone, one in-memory patch: breakage B4) — oracle/make_golden_backbone.py,
                         tests/golden/backbone_tiny.npz.
  * dense head (N2)    — oracle/dense_head.py: assembly restated (the head class is not importable: B5/B6), its decode / anchor /
                         limit_period building blocks pinned bit-exactly on the reference's own functions (tests/test_oracle_cpu.py).
  * post-processing (N3) — oracle/post_process.py: PARITY UNPINNED for the NMS (the reference's iou3d_nms CUDA op is not in the tree);
                         selection logic restated from the Python source, rotated IoU restated independently in float64.
"""
