"""N>1 plumbing of bench.py on CPU: world_size 2 over gloo (one process per rank, torchrun env),
max-over-ranks timing, whole-job aggregation, rank 0 alone prints; and the reference arm's
rank handling."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def free_port() -> int:
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def torchrun(nproc, *bench_args, timeout=240):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(ROOT / "bench.py"), *bench_args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def json_lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def test_two_ranks_gloo_max_and_aggregate():
    res = torchrun(2, "--gpus", "2", "--dist-selftest")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = json_lines(res.stdout)
    assert len(lines) == 1, "rank 0 alone prints"
    line = lines[0]
    assert line["n_gpus"] == 2
    assert line["ms"] == 20.0  # max(10, 20)
    assert line["value"] == 2 * 1000 * 2 / 0.020


def test_aggregate_value_is_whole_job():
    sys.path.insert(0, str(ROOT))
    import bench
    assert bench.aggregate_value(8, 1 << 22, 3, 100.0) == 8 * (1 << 22) * 3 / 0.1
    assert bench.fib_iterations(22) * 8 + 8 == 1 << 22
