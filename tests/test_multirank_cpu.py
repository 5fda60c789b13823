"""CPU tests of the N > 1 host logic with world_size 2 over gloo: subject sharding, identical phi
starts on every rank, the out-of-band exchange used for the NCCL id, and the reference arm's
rank handling.  (The GPU data path itself is covered by the -m gpu tests and the multi-GPU bench.)"""
import hashlib
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _digest(sh):
    h = hashlib.sha256()
    for t in sh.trials:
        h.update(t.rt.tobytes())
        h.update(t.cell.tobytes())
    h.update(sh.subj0.tobytes())
    return h.hexdigest()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from ggdmc_b200 import workloads as W

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        S = 7  # odd on purpose: shards of 3 and 4
        b, e = W.shard_bounds(S, rank, world)
        sh = W.build_population(2, S, 64, n_replicate=2, subject_begin=b, subject_end=e)
        # the NCCL unique id travels like this in bench.py
        uid = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        info = {"rank": rank, "bounds": (b, e), "phi": hashlib.sha256(sh.phi0.tobytes()).hexdigest(), "digest": _digest(sh),
                "n": len(sh.trials), "uid_ok": uid[0] == bytes(range(128)),
                "per_subject": [hashlib.sha256(t.rt.tobytes()).hexdigest() for t in sh.trials]}
        out = [None] * world
        dist.all_gather_object(out, info)
        import torch
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, out, float(t.item())))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    import torch.multiprocessing as mp
    from ggdmc_b200 import workloads as W

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out, mx in res:
        assert mx == 2.0
        assert [o["rank"] for o in out] == [0, 1]
        assert out[0]["bounds"] == (0, 3) and out[1]["bounds"] == (3, 7)
        assert out[0]["phi"] == out[1]["phi"]  # phi start replicated without communication
        assert all(o["uid_ok"] for o in out)
    full = W.build_population(2, 7, 64, n_replicate=2)
    per_subject = [hashlib.sha256(t.rt.tobytes()).hexdigest() for t in full.trials]
    out = res[0][1]
    assert out[0]["per_subject"] + out[1]["per_subject"] == per_subject  # shards are slices of the 1-rank problem
    assert hashlib.sha256(full.phi0.tobytes()).hexdigest() == out[0]["phi"]


@pytest.mark.parametrize("world,S", [(1, 5), (2, 1024), (3, 10), (8, 1024), (8, 5)])
def test_shard_bounds_partition(world, S):
    from ggdmc_b200 import workloads as W
    cuts = [W.shard_bounds(S, r, world) for r in range(world)]
    assert cuts[0][0] == 0 and cuts[-1][1] == S
    assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
    sizes = [e - b for b, e in cuts]
    assert max(sizes) - min(sizes) <= 1


def test_reference_arm_runs_on_rank0_only():
    """`bench.py --impl reference`: rank 0 prints the JSON line, other ranks exit 0 without work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "trial-likelihoods/s"
