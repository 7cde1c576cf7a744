"""ORACLE (test infrastructure) — mint tests/golden/backbone_tiny.npz from the REFERENCE'S OWN BaseBEVBackbone_Scale.

Run in the build container, where /root/reference exists:   python -m oracle.make_golden_backbone
The fixture stores only seeds + the reference's outputs; weights and inputs are regenerated from the seeds by
oracle/backbone.py (numpy PCG64, stable), with a checksum to catch drift.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import backbone as ob, ref_loader   # noqa: E402

WSEED, XSEED, B, H, W = 77, 78, 1, 16, 24


def weights_digest(w) -> str:
    h = hashlib.sha256()
    for k in sorted(w):
        h.update(k.encode() + np.ascontiguousarray(w[k]).tobytes())
    return h.hexdigest()


def main():
    ns = ref_loader.load_backbone()
    w = ob.random_backbone_weights(WSEED)
    m = ns.BaseBEVBackbone_Scale(ref_loader.BACKBONE_CFG, 128).eval()
    missing = m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    assert not missing.unexpected_keys and all(k.endswith("num_batches_tracked") for k in missing.missing_keys), missing
    spatial, scale = ob.random_canvases(XSEED, B, H, W)
    with torch.no_grad():
        out = m({"spatial_features": torch.from_numpy(spatial), "spatial_scale_features": torch.from_numpy(scale)})
    ref = out["spatial_features_2d"].numpy()
    mine, lv = ob.backbone_forward(w, spatial, scale, return_levels=True)
    err = np.abs(mine - ref).max() / np.abs(ref).max()
    print("restatement vs reference module: max rel err %.3g" % err)
    assert err < 1e-5
    path = os.path.join(ROOT, "tests", "golden", "backbone_tiny.npz")
    np.savez_compressed(path, wseed=WSEED, xseed=XSEED, shape=np.array([B, H, W]), weights_sha256=weights_digest(w),
                        spatial_features_2d=ref)
    print("wrote", path, os.path.getsize(path), "bytes; out", ref.shape, "absmax %.3f mean %.3f" % (np.abs(ref).max(), ref.mean()))


def main_plain():
    """the plain BaseBEVBackbone (PointPillars layout: 64 input channels, first level stride 2)"""
    ns = ref_loader.load_backbone()
    w = ob.random_backbone_weights(WSEED + 1, ob.PLAIN_CFG, 64, with_scale=False)
    m = ns.BaseBEVBackbone(ref_loader.Cfg(**ob.PLAIN_CFG), 64).eval()
    missing = m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    assert not missing.unexpected_keys and all(k.endswith("num_batches_tracked") for k in missing.missing_keys), missing
    rng = np.random.default_rng(XSEED + 1)
    spatial = (np.abs(rng.standard_normal((B, 64, 32, 48))) * (rng.random((B, 1, 32, 48)) < 0.2)).astype(np.float32)
    with torch.no_grad():
        ref = m({"spatial_features": torch.from_numpy(spatial)})["spatial_features_2d"].numpy()
    mine = ob.backbone_forward(w, spatial, None, ob.PLAIN_CFG)
    err = np.abs(mine - ref).max() / np.abs(ref).max()
    print("plain restatement vs reference module: max rel err %.3g" % err)
    assert err < 1e-5
    path = os.path.join(ROOT, "tests", "golden", "backbone_plain_tiny.npz")
    np.savez_compressed(path, wseed=WSEED + 1, xseed=XSEED + 1, weights_sha256=weights_digest(w), spatial_features=spatial,
                        spatial_features_2d=ref)
    print("wrote", path, os.path.getsize(path), "bytes; out", ref.shape)


if __name__ == "__main__":
    main()
    main_plain()
