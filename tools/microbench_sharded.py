"""Column-sharded commitment microbench (BASELINE configs[3]/[4] shape: synthetic 2^k-row trace,
columns sharded across N GPUs, NCCL all-to-all + root all-gather).

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/microbench_sharded.py \
      [--log-size 22] [--cols 64] [--iters 5]

Prints one JSON line (rank 0): aggregate algorithmic GB/s of the commit pipeline (28n bytes per
column), the per-phase device times (max over ranks) and, at --verify, root equality
with the single-GPU commit of the same trace."""
import argparse
import importlib.util
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
spec = importlib.util.spec_from_file_location("sharded_commit", ROOT / "cairo-m_b200" / "sharded_commit.py")
sc = importlib.util.module_from_spec(spec)
spec.loader.exec_module(sc)

P = (1 << 31) - 1


def make_columns(lo, hi, n, device):
    # seeded per global column index, so every world size commits the same trace
    cols = []
    for c in range(lo, hi):
        g = torch.Generator(device=device)
        g.manual_seed(0xCA1120 + c)
        cols.append(torch.randint(0, P, (n,), generator=g, device=device, dtype=torch.int64).to(torch.int32))
    return cols


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-size", type=int, default=22)
    ap.add_argument("--cols", type=int, default=64)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--mode", default="p2p", choices=["p2p", "nccl"], help="p2p: leaf kernel reads peer LDE buffers; nccl: all-to-all")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    cm = __import__("importlib").import_module("cairo-m_b200")
    cm.check(cm.lib().cm31_set_device(local))
    ops = sc.CudaOps(args.log_size + 2)
    n = 1 << args.log_size
    lo, hi = sc.column_range(args.cols, world, rank)
    times = []
    root = None
    peer = sc.PeerLde(args.cols, args.log_size + 1, dist if world > 1 else None, rank, world) if args.mode == "p2p" else None
    for it in range(args.warmup + args.iters):
        cols = make_columns(lo, hi, n, device)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if peer is not None:
            root = sc.sharded_commit_p2p(ops, peer, cols, args.log_size, 1, dist if world > 1 else None)
        else:
            root, _ = sc.sharded_commit(ops, cols, args.cols, args.log_size, 1, dist if world > 1 else None, rank, world)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        if it >= args.warmup:
            times.append(ms)
    ok = None
    if args.verify:
        # single-GPU commit of the whole trace on rank 0 (needs all columns to fit one GPU)
        if rank == 0:
            full = make_columns(0, args.cols, n, device)
            ref_root, _ = sc.sharded_commit(ops, full, args.cols, args.log_size, 1)
            ok = bool(torch.equal(ref_root.cpu(), root.cpu()))
    if rank == 0:
        ms = sorted(times)[len(times) // 2]
        total_bytes = sc.commit_bytes_per_column(args.log_size, 1) * args.cols
        peaks = ROOT / "MEASURED_PEAKS.json"
        peak = json.loads(peaks.read_text())["hbm_gbs"] if peaks.exists() else 6650.0
        print(json.dumps({"bench": "sharded_commit", "mode": args.mode, "n_gpus": world, "log_size": args.log_size, "cols": args.cols, "ms": ms,
                          "alg_GBps_aggregate": total_bytes / ms / 1e6, "alg_GBps_per_gpu": total_bytes / ms / 1e6 / world,
                          "frac_of_hbm_peak_per_gpu": total_bytes / ms / 1e6 / world / peak,
                          "all_to_all_bytes_per_gpu": 4 * (hi - lo) * (2 * n) * (world - 1) // world, "root_matches_single_gpu": ok,
                          "root": root.cpu().numpy().view("uint32").tolist()}), flush=True)
    if peer is not None:
        if world > 1:
            dist.barrier()
        peer.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
