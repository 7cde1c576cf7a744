"""Shared helpers for the parity tests (imports the ORACLE — allowed only under tests/)."""
import hashlib
import os

import numpy as np
import torch

from hvpr_b200.geometry import Geometry

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TOL_FP32 = 1e-4     # north_star: pillar / BEV features within 1e-4 relative (fp32 path)
TOL_BF16 = 1e-2     # north_star: within 1e-2 for the bf16 memory-attention variant


def sha(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.shape).encode() + str(a.dtype).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def rel_err(a: torch.Tensor, b: torch.Tensor):
    """(max|a-b| / max|b| , ||a-b||_2 / ||b||_2) — the metric fixed in SURVEY.md §8d."""
    a, b = a.double().cpu(), b.double().cpu()
    d = (a - b)
    return float(d.abs().max() / b.abs().max().clamp_min(1e-30)), float(d.norm() / b.norm().clamp_min(1e-30))


def load_small(name):
    z = np.load(os.path.join(GOLDEN, "small_%s.npz" % name))
    geom = Geometry(tuple(float(x) for x in z["geom_range"]), tuple(float(x) for x in z["geom_voxel"]),
                    int(z["max_points"]), int(z["max_voxels"]))
    sizes = z["frame_sizes"]
    off = np.r_[0, np.cumsum(sizes)]
    frames = [z["points"][off[i]:off[i + 1]] for i in range(len(sizes))]
    return z, geom, frames, str(z["overflow"]), int(z["wseed"])


def to_dev(frames, device="cuda"):
    pts = torch.from_numpy(np.ascontiguousarray(np.concatenate(frames, 0))).to(device)
    off = torch.tensor(np.r_[0, np.cumsum([len(f) for f in frames])], dtype=torch.int32, device=device)
    return pts, off


def tie_aware_readout_check(out, ref, pillars, W, tol, k=20, idx=None, ref_idx=None, max_flip_frac=2e-3,
                            tie_eps=5e-5):
    """Parity for the memory readout, which is DISCONTINUOUS in its input: top-k swaps an item of weight ~1/k whenever
    the k-th and (k+1)-th logits are closer than the fp32 rounding noise of the 64-term dot product, and any two
    correct fp32 implementations with different summation orders (MKL sgemm, cuBLAS, this kernel) disagree on such rows.
      * rows whose readout is within `tol` (max-norm relative to max|ref|) pass outright;
      * every other row must be a PROVEN near-tie in fp64: (v_k - v_{k+1}) <= tie_eps * max|logit_row|, and, when the
        selected indices are available, the kernel's set must still be a valid top-k up to that noise;
      * such rows must be rare (<= max_flip_frac).
    Returns (max rel err over non-tie rows, fraction of tie rows)."""
    out, ref = out.double().cpu(), ref.double().cpu()
    scale = float(ref.abs().max().clamp_min(1e-30))
    row_err = (out - ref).abs().max(1)[0] / scale
    bad = (row_err > tol).nonzero()[:, 0]
    frac = bad.numel() / max(1, out.shape[0])
    assert frac <= max_flip_frac, "too many rows off: %g" % frac
    if bad.numel():
        L = pillars.double().cpu()[bad] @ W.double().cpu().t()
        top = torch.topk(L, k + 1, dim=1)[0]
        gap = top[:, k - 1] - top[:, k]
        lim = tie_eps * L.abs().max(1)[0]
        assert bool((gap <= lim).all()), "row off by %g without a near-tie (gap %g, limit %g)" % (
            float(row_err[bad].max()), float(gap.max()), float(lim.min()))
        if idx is not None:
            sel = torch.gather(L, 1, idx.cpu().long()[bad])
            assert bool((sel.min(1)[0] >= top[:, k] - lim).all()), "selected set is not a valid top-k"
    good_err = float(row_err[row_err <= tol].max()) if (row_err <= tol).any() else 0.0
    return good_err, frac
