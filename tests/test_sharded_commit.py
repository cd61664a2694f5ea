"""Column-sharded commitment (cairo-m_b200/sharded_commit.py): the N>1 data path with its two
collectives (all-to-all of LDE columns into row ranges, all-gather of sub-tree roots) on CPU over
gloo, world sizes 1, 2 and 4 — every rank must reproduce the single-process Merkle root."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def free_port() -> int:
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]
WORKER = ROOT / "tests" / "dist" / "sharded_commit_worker.py"


def run(world, log_size, log_blowup, n_cols, port):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    if world == 1:
        cmd = [sys.executable, str(WORKER), str(log_size), str(log_blowup), str(n_cols)]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(free_port()), str(WORKER), str(log_size), str(log_blowup), str(n_cols)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)


@pytest.mark.parametrize("world,log_size,n_cols", [(1, 6, 5), (2, 6, 5), (2, 8, 2), (4, 7, 9), (4, 5, 3)])
def test_sharded_commit_root_equals_single_process(world, log_size, n_cols):
    res = run(world, log_size, 1, n_cols, 29620 + world + log_size)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    import re
    results = re.findall(r"RESULT rank=(\d+) root_ok=(\w+) rows_ok=(\w+)", res.stdout)
    assert sorted(int(r[0]) for r in results) == list(range(world))
    assert all(r[1] == "True" and r[2] == "True" for r in results)
