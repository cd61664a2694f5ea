"""Worker of tests/test_sharded_commit.py: one rank of a gloo world running sharded_commit with CPU
ops built on the oracle (test infrastructure), compared with the single-process commit."""
import importlib.util
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from tests import oracle_lib as orc  # noqa: E402

spec = importlib.util.spec_from_file_location("sharded_commit", ROOT / "cairo-m_b200" / "sharded_commit.py")
sc = importlib.util.module_from_spec(spec)
spec.loader.exec_module(sc)


class CpuOps:
    torch = torch

    def empty(self, shape):
        return torch.empty(shape, dtype=torch.int32)

    @staticmethod
    def _np(t):
        return t.numpy().view(np.uint32)

    def interpolate(self, cols, log_size):
        for c in cols:
            self._np(c)[:] = orc.interpolate(self._np(c), log_size)[0]

    def evaluate(self, coeffs, out, log_size, log_eval):
        for c, o in zip(coeffs, out):
            self._np(o)[:] = orc.evaluate(self._np(c), log_size, log_eval)[0]

    def commit_layer(self, log_size, prev, cols):
        mat = np.stack([self._np(c) for c in cols]) if cols else None
        p = self._np(prev.contiguous()) if prev is not None else None
        return torch.from_numpy(orc.commit_on_layer(log_size, p, mat).view(np.int32))

    def sync(self):
        pass


def main():
    log_size, log_blowup, n_cols = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        dist.init_process_group("gloo")
    trace = orc.splitmix64(0x5EED, n_cols << log_size).reshape(n_cols, 1 << log_size)
    lo, hi = sc.column_range(n_cols, world, rank)
    local = [torch.from_numpy(trace[c].copy().view(np.int32)) for c in range(lo, hi)]
    root, rows = sc.sharded_commit(CpuOps(), local, n_cols, log_size, log_blowup, dist if world > 1 else None, rank, world)
    # single-process reference: the same pipeline on all columns
    lde = orc.evaluate(orc.interpolate(trace, log_size), log_size, log_size + log_blowup)
    layer = orc.commit_on_layer(log_size + log_blowup, None, lde)
    for log in range(log_size + log_blowup - 1, -1, -1):
        layer = orc.commit_on_layer(log, layer, None)
    ok = np.array_equal(root.numpy().view(np.uint32), layer.reshape(8))
    m = 1 << (log_size + log_blowup)
    r0 = rank * (m // world)
    ok_rows = all(np.array_equal(rows[c].numpy().view(np.uint32), lde[c, r0:r0 + m // world]) for c in range(n_cols))
    line = f"RESULT rank={rank} root_ok={ok} rows_ok={ok_rows}"
    if world > 1:
        # one writer: concurrent prints of several ranks can interleave inside a line on a shared pipe
        lines = [None] * world
        dist.all_gather_object(lines, line)
        if rank == 0:
            sys.stdout.write("\n".join(lines) + "\n")
            sys.stdout.flush()
        dist.barrier()
    else:
        print(line, flush=True)
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok and ok_rows else 1)


if __name__ == "__main__":
    main()
