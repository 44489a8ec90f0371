"""CPU, world_size 2, gloo: the host-side multi-GPU logic (sharding, max-over-ranks timing, DDP wrap with the frozen dead
branch) and bench.py's reference arm under a torchrun-style environment (rank 0 alone prints)."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    from openvivqa_b200 import dist as D

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert D.env_rank_world() == (rank, rank, world)
        # timing reduction: slowest rank wins
        assert D.reduce_max(10.0 + rank) == 10.0 + world - 1
        # batch partition: disjoint cover
        lo, hi = D.shard_range(101, rank, world)
        spans = [None] * world
        dist.all_gather_object(spans, (lo, hi))
        assert spans[0][0] == 0 and spans[-1][1] == 101 and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))

        class Net(torch.nn.Module):                       # stand-in with a declared-unused branch
            def __init__(self):
                super().__init__()
                self.used = torch.nn.Linear(8, 8)
                self.dead = torch.nn.Linear(8, 8)

            def unused_parameters(self):
                return list(self.dead.parameters())

            def forward(self, x):
                return self.used(x)

        torch.manual_seed(0)
        net = Net()
        ddp = D.data_parallel(net)
        assert all(not p.requires_grad for p in net.dead.parameters())
        x = torch.full((4, 8), float(rank + 1))
        ddp(x).sum().backward()                           # would hang/raise on the dead branch without freeze_unused
        g = net.used.weight.grad.clone()
        gs = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(gs, g)
        assert all(torch.equal(gs[0], t) for t in gs)     # averaged gradient identical on every rank
        want = torch.full((8, 8), 4.0 * sum(r + 1 for r in range(world)) / world)
        assert torch.allclose(g, want)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_host_logic():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(30)
    assert res == {0: "ok", 1: "ok"}, res


def test_reference_arm_rank0_only_prints():
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999", OMP_NUM_THREADS="4",
               OVQA_REFERENCE_ARM_BATCH="8", OVQA_REFERENCE_ARM_M4C_BATCH="1")        # (bounded: the arm's default is the workload's batch of 256)
    outs = []
    for rank in (0, 1):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                           env=dict(env, RANK=str(rank), LOCAL_RANK=str(rank)), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-500:]
        outs.append([l for l in r.stdout.splitlines() if l.startswith("{")])
    assert len(outs[0]) == 1 and outs[1] == []
    line = json.loads(outs[0][0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["unit"] == "samples/s"
    assert line["m4c_decode"]["value"] > 0
