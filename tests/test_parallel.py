"""CPU, world_size 2 over gloo: the batch-shard + token all-gather step of multi-GPU generation
(lina_speech_b200/parallel.py), and bench.py's reference arm under a 2-process launch."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lina_speech_b200.parallel import gather_tokens, shard_range
    Q, b = 2, 3
    lo, hi = shard_range(world * b, rank, world)
    assert (lo, hi) == (rank * b, rank * b + b)
    ok = True
    for step in range(4):
        q_local = (torch.arange(Q * b).view(Q, b, 1) + 100 * rank + 1000 * step).long()
        q_all = gather_tokens(q_local)
        expect = torch.cat([(torch.arange(Q * b).view(Q, b, 1) + 100 * r + 1000 * step) for r in range(world)], dim=1)
        ok = ok and torch.equal(q_all, expect) and q_all.shape == (Q, world * b, 1)
        stop = (q_all == 2).prod(dim=0)                       # the global stop bookkeeping every rank derives
        ok = ok and stop.shape == (world * b, 1)
    ret[rank] = ok
    dist.destroy_process_group()


def test_token_all_gather_world2_gloo():
    world, port = 2, 29611
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def test_shard_range_is_a_partition():
    from lina_speech_b200.parallel import shard_range
    for B in (1, 7, 32, 256):
        for W in (1, 2, 3, 8):
            spans = [shard_range(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_bench_reference_arm_two_process_launch():
    """`torchrun --nproc-per-node 2 bench.py --impl reference`: rank 0 prints one JSON line, rank 1 exits 0."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29613", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0", "--cpu-budget", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and out["unit"] == "tokens/s" and out["value"] > 0
    assert out["cpu_baseline"]["kind"] == "port" and out["e2e"]["h2d_bytes_per_step"] == 0
