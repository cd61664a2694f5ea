"""Single-proof sharding (csrc/shard.cu): the proof made by N ranks together is byte-identical to the proof one
GPU makes alone, on every rank.  The data plane needs GPUs (`-m gpu`, one subprocess per world size through torchrun); the
host-side plan (component -> rank assignment, striped node ranges) is checked on CPU with gloo at world size 2."""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
WORKER = ROOT / "tests" / "dist" / "sharded_prover_worker.py"


def run_worker(world, program, n, port):
    env = dict(os.environ, CM31_ARENA_GIB="12")
    if world == 1:
        cmd = [sys.executable, str(WORKER), str(program), str(n)]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), str(WORKER), str(program), str(n)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))
    # one record per rank; the ranks share a pipe, so two records can end up on one line (no newline in between): match
    # the records themselves, not lines
    import re
    lines = re.findall(r"SHARDED_(?:OK|MISMATCH) rank=\d+ world=\d+ sha256=[0-9a-f]+ .*?collectives=\d+", res.stdout + res.stderr)
    assert res.returncode == 0 and len(lines) == world and all(l.startswith("SHARDED_OK") for l in lines), (res.stdout + res.stderr)[-3000:]
    return lines


@pytest.mark.gpu
@pytest.mark.parametrize("program,n", [(0, 2000), (4, 2)])
def test_sharded_proof_world_1_equals_plain_proof(program, n):
    # the sharded code path (peer-mapped arena, NCCL communicator, striping planner) with a single rank
    run_worker(1, program, n, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("program,n", [(0, 3000), (0, 140_000), (4, 40)])
def test_sharded_proof_equals_single_gpu_proof_on_every_rank(world, program, n):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    lines = run_worker(world, program, n, 29600 + world)
    digests = {l.split("sha256=")[1].split()[0] for l in lines}
    assert len(digests) == 1


# ---------------------------------------------------------------- host-side plan, CPU
def plan(cm, cost, world):
    arr = (C.c_double * len(cost))(*cost)
    out = (C.c_int * len(cost))()
    cm.check(cm.lib().cm31_shard_plan(arr, C.c_size_t(len(cost)), C.c_int(world), out))
    return list(out)


def test_component_assignment_is_balanced_and_deterministic(cm):
    cost = [1 << 21, 1 << 20, 1 << 19, 1 << 19, 1 << 18] + [16.0] * 29
    assert plan(cm, cost, 1) == [0] * 34
    for world in (2, 4, 8):
        owners = plan(cm, cost, world)
        assert owners == plan(cm, cost, world) and all(0 <= o < world for o in owners)
        load = [sum(c for c, o in zip(cost, owners) if o == r) for r in range(world)]
        assert max(load) <= max(max(cost), sum(cost) / world * 4 / 3 + max(cost) / 3)  # LPT bound
        assert owners[0] != owners[1] or world == 1  # the two largest components never share a rank


def test_every_rank_computes_the_same_plan_gloo_world_2(tmp_path):
    # two processes (gloo): each computes the plan for the same costs with its own libcm31 and they compare results
    script = tmp_path / "plan_worker.py"
    script.write_text(f"""
import ctypes as C, importlib, os, sys
sys.path.insert(0, {str(ROOT)!r})
import torch, torch.distributed as dist
dist.init_process_group("gloo")
cm = importlib.import_module("cairo-m_b200")
cost = [float((i * 7919) % 1000 + 1) * (1 << (i % 5)) for i in range(34)]
arr = (C.c_double * 34)(*cost); out = (C.c_int * 34)()
cm.check(cm.lib().cm31_shard_plan(arr, C.c_size_t(34), C.c_int(dist.get_world_size()), out))
mine = torch.tensor(list(out), dtype=torch.int64)
gathered = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
dist.all_gather(gathered, mine)
ok = all(torch.equal(g, mine) for g in gathered) and set(mine.tolist()) == set(range(dist.get_world_size()))
# striped node ranges: disjoint, cover the layer, children of a range stay in the owner's range of the layer below
world, rank = dist.get_world_size(), dist.get_rank()
for log in (11, 15, 22):
    n = 1 << log
    first, count = rank * (n // world), n // world
    below_first, below_count = rank * (2 * n // world), 2 * n // world
    ok = ok and 2 * first == below_first and 2 * (first + count) == below_first + below_count
print("PLAN_OK" if ok else "PLAN_MISMATCH", dist.get_rank(), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
""")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29655", str(script)], capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert res.returncode == 0 and res.stdout.count("PLAN_OK") == 2, (res.stdout + res.stderr)[-2000:]
