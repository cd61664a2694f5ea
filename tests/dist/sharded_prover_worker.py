"""One rank of a sharded proof (launched by torchrun, or directly for world size 1): proves the same input first on its own
GPU alone, then as one share of a proof sharded over all ranks (components dealt out over the ranks, row-striped Merkle
layers / DEEP quotients / accumulator sums over NVLink, NCCL for the small joins), and checks that the sharded proof is
byte-identical to the single-GPU one ON EVERY RANK.  Prints one line per rank: "SHARDED_OK rank world sha256 ms".

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/dist/sharded_prover_worker.py [program] [n]
"""
import ctypes as C
import hashlib
import importlib
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    program = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cm = importlib.import_module("cairo-m_b200")
    lib = cm.lib()
    cm.check(lib.cm31_set_device(local))
    h = C.c_void_p()
    cm.check(lib.cm31_test_program_input_create(C.c_uint32(program), C.c_uint32(n), C.byref(h)))
    cm.check(lib.cm31_input_upload(h))
    cap = 1 << 26
    buf = (C.c_uint8 * cap)()
    ln = C.c_size_t()

    def prove():
        cm.check(lib.cm31_prove_cairo_m(h, 16, 80, buf, C.c_size_t(cap), C.byref(ln), None))
        return C.string_at(buf, ln.value)

    alone = prove()
    cm.shard_init(arena_gib=float(os.environ.get("CM31_ARENA_GIB", "16")))
    sharded = prove()
    again = prove()  # the arena is recycled between proofs
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prove()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    ok = sharded == alone and again == alone
    if not ok and rank == 0:  # where do the two proofs part?
        import json

        def js(b):
            raw = (C.c_uint8 * len(b)).from_buffer_copy(b)
            k = C.c_size_t()
            cm.check(lib.cm31_proof_to_json(raw, C.c_size_t(len(b)), None, C.c_size_t(0), C.byref(k)))
            out = C.create_string_buffer(k.value + 1)
            cm.check(lib.cm31_proof_to_json(raw, C.c_size_t(len(b)), out, C.c_size_t(k.value + 1), C.byref(k)))
            return json.loads(out.value.decode())
        a, s_ = js(alone), js(sharded)
        names = list(a["claim"]["opcodes"]) + [k for k in a["claim"] if k != "opcodes"]
        sums = lambda p: [p["interaction_claim"]["opcodes"][k]["claimed_sum"] if k in p["interaction_claim"]["opcodes"] else p["interaction_claim"][k]["claimed_sum"] for k in names]
        print("DIFF claim:", a["claim"] == s_["claim"], "sums differ at:", [n for n, x, y in zip(names, sums(a), sums(s_)) if x != y][:8], flush=True)
        print("DIFF commitments equal:", [x == y for x, y in zip(a["stark_proof"]["commitments"], s_["stark_proof"]["commitments"])], flush=True)
        for t in range(len(a["stark_proof"]["sampled_values"])):
            da = [i for i, (x, y) in enumerate(zip(a["stark_proof"]["sampled_values"][t], s_["stark_proof"]["sampled_values"][t])) if x != y]
            print(f"DIFF sampled tree {t}: {len(da)} of {len(a['stark_proof']['sampled_values'][t])} columns differ, first {da[:6]}", flush=True)
    digest = hashlib.sha256(sharded).hexdigest()
    if world > 1:
        flags = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        ok = ok and bool(flags.item())
    stats = cm.shard_stats()
    print(f"{'SHARDED_OK' if ok else 'SHARDED_MISMATCH'} rank={rank} world={world} sha256={digest[:16]} bytes={len(sharded)} ms={ms:.2f} "
          f"arena_peak_mib={stats['arena_peak_bytes'] >> 20} gathered_mib={stats['bytes_all_gathered'] >> 20} collectives={stats['collectives']}",
          flush=True)
    cm.check(lib.cm31_input_destroy(h))
    cm.check(lib.cm31_shard_finalize())
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
