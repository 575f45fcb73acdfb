"""CPU, world_size 2, gloo: the multi-GPU host logic of the path — contiguous batch shards per rank, identical
weights, one all_gather of logits + kept indices, max-over-ranks timing reduction.  The per-rank
compute is the oracle port here (no GPU in this container); on the GPU box bench.py runs the same plumbing over
NCCL with the tokred kernels."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ops as O
    torch.manual_seed(0)                                    # identical inputs/weights on every rank
    b, n, c, k = 8, 65, 32, 40
    x = torch.randn(b, n, c)
    scores = torch.stack([torch.randperm(n - 1).float() for _ in range(b)])
    per = b // world
    xs, ss = x[rank * per:(rank + 1) * per], scores[rank * per:(rank + 1) * per]
    out, idx = O.topk_gather(xs, ss, k)                     # shard-local reduction, no data-path collective
    outs = [torch.empty_like(out) for _ in range(world)]
    idxs = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(outs, out)
    dist.all_gather(idxs, idx)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)     # max-over-ranks timing reduction
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        full_out, full_idx = O.topk_gather(x, scores, k)
        ok = torch.equal(torch.cat(outs), full_out) and torch.equal(torch.cat(idxs), full_idx) and float(t) == float(world)
        with open(out_path, "w") as fh:
            fh.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_batch_sharding_gloo_world2(tmp_path):
    out = tmp_path / "result.txt"
    port = 29400 + os.getpid() % 500
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"


def test_reference_arm_runs_on_rank0_only(tmp_path):
    """bench.py --impl reference under a 2-rank launch: rank 1 exits without work, rank 0 prints the JSON line."""
    import json
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1", "--cpu-batch", "2", "--workload", "topk_small_kr0.7_b64"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] in ("reference", "port") and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0
