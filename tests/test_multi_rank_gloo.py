"""CPU suite: the N>1 host logic with world_size-2 gloo (frame-wise sharding, timing reduction, reference-arm rank
gating).  No GPU kernels run here; the data path has no collective to test."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import torch
import torch.distributed as dist
from hvpr_b200 import sharding, synth
from hvpr_b200.geometry import G1
rank, local_rank, world = sharding.dist_env()
d = sharding.init_process_group("gloo")
assert d.get_world_size() == world == 2 and d.get_rank() == rank
# strong-scaling partition: frame i -> rank i mod W, every frame exactly once
mine = sharding.frames_of_rank(13, rank, world)
owned = [None, None]
d.all_gather_object(owned, mine)
assert sorted(owned[0] + owned[1]) == list(range(13)) and not (set(owned[0]) & set(owned[1]))
# weak-scaling partition used by bench.py: distinct seeds per rank, same count
w = sharding.weak_scaling_frames(8, rank)
allw = [None, None]
d.all_gather_object(allw, w)
assert allw[0] == list(range(8)) and allw[1] == list(range(8, 16))
# frames of different ranks are different data
f = synth.make_frame("L", 2000, G1.point_cloud_range, 1024 + w[0])
chk = [None, None]
d.all_gather_object(chk, float(f.sum()))
assert chk[0] != chk[1]
# elapsed time is the MAX over ranks; scalars gather in rank order
t = sharding.max_over_ranks(1.0 + rank)
assert t == 2.0
assert sharding.gather_scalars(10.0 * rank) == [0.0, 10.0]
d.barrier()
d.destroy_process_group()
print(json.dumps({"rank": rank, "ok": True}))
'''


def _run_two_ranks(tmp_path, script_text, extra_args=()):
    script = tmp_path / "worker.py"
    script.write_text(script_text)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", str(script), *extra_args]
    return subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=240, cwd=ROOT)


def test_two_rank_gloo_sharding(tmp_path):
    r = _run_two_ranks(tmp_path, WORKER % {"root": ROOT})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    import re
    oks = [json.loads(m) for m in re.findall(r'\{"rank": \d+, "ok": true\}', r.stdout)]
    assert sorted(o["rank"] for o in oks) == [0, 1] and all(o["ok"] for o in oks), r.stdout[-500:]


def test_reference_arm_only_rank0_prints(tmp_path):
    """bench.py --impl reference under torchrun: rank 0 alone runs and prints ONE line, other ranks exit 0 silently."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29542", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-500:]
    d = lines[0]
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_frames_of_rank_properties():
    from hvpr_b200 import sharding
    for n in (0, 1, 7, 64):
        for w in (1, 2, 4, 8):
            parts = [sharding.frames_of_rank(n, r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
