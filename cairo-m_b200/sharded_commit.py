"""Column-sharded tree commitment across GPUs: the one step of the proving path
with a real data exchange.

`TreeBuilder::extend_evals` + `commit` (external/stwo/crates/prover/src/core/pcs/prover.rs:173-249)
on a trace whose COLUMNS are spread over the ranks:

  1. column-parallel, no communication: interpolate (ICFFT) + low-degree extension (CFFT) of the
     rank's own columns;
  2. the exchange: a Merkle leaf hashes one row of ALL columns (vcs/prover.rs:56-63), so the LDE is
     re-laid out from column shards to row ranges with ONE all-to-all (NCCL over NVLink);
  3. row-parallel, no communication: each rank hashes the leaves of its row range and builds that
     sub-tree down to its root;
  4. ONE all-gather of the G sub-tree roots (32 bytes each); the top log2(G) layers are computed
     redundantly, so every rank ends with the same root — bit-identical to the single-GPU commit.

The math ops are injected (`ops`): `CudaOps` = libcm31 through the C ABI (the product);
tests/test_sharded_commit.py drives the same code over gloo with CPU ops built on the oracle.
All columns of one call have the same size (a synthetic trace / one component); mixed sizes
inject smaller columns lower in the tree and are handled by the single-GPU path only.
"""
from __future__ import annotations

import importlib


class CudaOps:
    """libcm31 kernels on the current CUDA device (torch only owns the buffers)."""

    def __init__(self, max_log_size: int):
        import torch
        self.torch = torch
        self.cm = importlib.import_module("cairo-m_b200")
        self.tw = self.cm.Twiddles(max_log_size)
        self.device = torch.device("cuda", torch.cuda.current_device())

    def empty(self, shape):
        return self.torch.empty(shape, dtype=self.torch.int32, device=self.device)

    def interpolate(self, cols, log_size):
        self.cm.interpolate_batch(cols, log_size, self.tw)

    def evaluate(self, coeffs, out, log_size, log_eval):
        self.cm.evaluate_batch(coeffs, out, log_size, log_eval, self.tw)

    def commit_layer(self, log_size, prev, cols):
        out = self.empty((1 << log_size, 8))
        self.cm.blake2s_commit_layer(log_size, prev, cols, out)
        return out

    def sync(self):
        self.cm.sync()


def _all_to_all(dist, recv, send, rank: int, world: int):
    """recv[p] <- rank p's send[rank].  NCCL: one all_to_all; gloo (CPU tests) has no alltoall, so the
    same exchange is issued as batched point-to-point sends/receives."""
    if dist.get_backend() == "nccl":
        dist.all_to_all(recv, send)
        return
    recv[rank].copy_(send[rank])
    ops = []
    for p in range(world):
        if p != rank:
            ops.append(dist.P2POp(dist.isend, send[p], p))
            ops.append(dist.P2POp(dist.irecv, recv[p], p))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def _ilog2(n: int) -> int:
    assert n > 0 and n & (n - 1) == 0, "power of two expected"
    return n.bit_length() - 1


def column_range(n_cols: int, world: int, rank: int):
    """Contiguous column shard of `rank` (the first n_cols % world ranks get one more)."""
    base, extra = divmod(n_cols, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sharded_commit(ops, local_cols, n_cols_total: int, log_size: int, log_blowup: int, dist=None, rank: int = 0, world: int = 1):
    """Commit a trace of `n_cols_total` columns of 2^log_size rows; this rank holds the evaluations
    of columns column_range(n_cols_total, world, rank) in `local_cols` (consumed: interpolated in place).

    Returns (root: 8-word tensor, rows: list of this rank's row range of every LDE column, in column order).
    """
    torch = ops.torch
    lo, hi = column_range(n_cols_total, world, rank)
    assert len(local_cols) == hi - lo
    log_eval = log_size + log_blowup
    m = 1 << log_eval
    assert world & (world - 1) == 0 and m >= 2 * world, "power-of-two world size, at least two rows per rank"
    log_rows = log_eval - _ilog2(world)
    rows_per_rank = 1 << log_rows

    # 1. column-parallel: interpolate + LDE of the local columns
    ops.interpolate(local_cols, log_size)
    lde = ops.empty((max(1, hi - lo), m))
    lde_cols = [lde[i] for i in range(hi - lo)]
    if lde_cols:
        ops.evaluate(local_cols, lde_cols, log_size, log_eval)
    ops.sync()  # the collective runs on torch's stream

    # 2. all-to-all: my columns' slice q goes to rank q; I receive every rank's columns for my rows
    if world > 1:
        send = [lde[: hi - lo, q * rows_per_rank:(q + 1) * rows_per_rank].contiguous() for q in range(world)]
        recv = []
        for p in range(world):
            plo, phi = column_range(n_cols_total, world, p)
            recv.append(ops.empty((phi - plo, rows_per_rank)))
        _all_to_all(dist, recv, send, rank, world)
        row_block = torch.cat(recv, dim=0)  # (n_cols_total, rows_per_rank), global column order
    else:
        row_block = lde[: hi - lo]
    rows = [row_block[c] for c in range(n_cols_total)]

    # 3. row-parallel: leaves of my row range, then my sub-tree
    layer = ops.commit_layer(log_rows, None, rows)
    for log in range(log_rows - 1, -1, -1):
        layer = ops.commit_layer(log, layer, [])
    ops.sync()

    # 4. all-gather the sub-tree roots; finish the top of the tree redundantly
    if world > 1:
        roots = [ops.empty((1, 8)) for _ in range(world)]
        dist.all_gather(roots, layer.contiguous())
        layer = torch.cat(roots, dim=0)  # layer log2(world), node q = root of rank q's rows
        for log in range(_ilog2(world) - 1, -1, -1):
            layer = ops.commit_layer(log, layer, [])
        ops.sync()
    return layer.reshape(8), rows


class PeerLde:
    """The LDE buffers of all ranks, each visible to every rank's kernels (CUDA IPC over
    NVLink/NVSwitch).  Built once per (column count, size); reused by every commit."""

    def __init__(self, n_cols_total: int, log_eval: int, dist, rank: int, world: int):
        import ctypes as C
        import torch
        self.cm = importlib.import_module("cairo-m_b200")
        self.C, self.torch = C, torch
        self.rank, self.world, self.m = rank, world, 1 << log_eval
        self.n_cols_total = n_cols_total
        lo, hi = column_range(n_cols_total, world, rank)
        self.n_local = hi - lo
        lib = self.cm.lib()
        self.local = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        self.cm.check(lib.cm31_ipc_alloc(C.c_size_t(4 * max(1, self.n_local) * self.m), C.byref(self.local), handle))
        mine = torch.tensor(list(handle), dtype=torch.uint8, device="cuda")
        handles = [torch.empty(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
        if world > 1:
            dist.all_gather(handles, mine)
        else:
            handles = [mine]
        self.base = []
        for p in range(world):
            if p == rank:
                self.base.append(self.local.value)
                continue
            ptr = C.c_void_p()
            h = (C.c_uint8 * 64)(*handles[p].cpu().tolist())
            self.cm.check(lib.cm31_ipc_open(h, C.byref(ptr)))
            self.base.append(ptr.value)

    def local_column(self, i: int) -> int:
        return self.local.value + 4 * i * self.m

    def row_range_columns(self, first_row: int):
        """Device addresses of every global column at `first_row` (remote ones are peer mappings)."""
        out = []
        for p in range(self.world):
            plo, phi = column_range(self.n_cols_total, self.world, p)
            for i in range(phi - plo):
                out.append(self.base[p] + 4 * (i * self.m + first_row))
        return out

    def close(self):
        lib = self.cm.lib()
        for p, b in enumerate(self.base):
            if p != self.rank:
                lib.cm31_ipc_close(self.C.c_void_p(b))
        lib.cm31_ipc_free(self.local)
        self.base = []


def sharded_commit_p2p(ops, peer: PeerLde, local_cols, log_size: int, log_blowup: int, dist=None):
    """Same result as sharded_commit, with the exchange FUSED into the leaf hashing: every rank's
    Merkle leaf kernel reads its row range of all columns straight from the owners' LDE buffers
    (peer loads over NVLink, coalesced 128 B per warp per column) while it hashes — no all-to-all,
    no staging copies; the only collective left is the 32-byte-per-rank root all-gather."""
    torch = ops.torch
    rank, world = peer.rank, peer.world
    log_eval = log_size + log_blowup
    assert peer.m == 1 << log_eval and len(local_cols) == peer.n_local
    log_rows = log_eval - _ilog2(world)
    ops.interpolate(local_cols, log_size)
    if local_cols:
        ops.evaluate(local_cols, [peer.local_column(i) for i in range(peer.n_local)], log_size, log_eval)
    ops.sync()
    if world > 1:
        dist.barrier()  # every owner's LDE is complete before anyone reads it
    layer = ops.commit_layer(log_rows, None, peer.row_range_columns(rank << log_rows))
    for log in range(log_rows - 1, -1, -1):
        layer = ops.commit_layer(log, layer, [])
    ops.sync()
    if world > 1:
        roots = [ops.empty((1, 8)) for _ in range(world)]
        dist.all_gather(roots, layer.contiguous())  # also orders the peers' reads before the buffers are reused
        layer = torch.cat(roots, dim=0)
        for log in range(_ilog2(world) - 1, -1, -1):
            layer = ops.commit_layer(log, layer, [])
        ops.sync()
    return layer.reshape(8)


def commit_bytes_per_column(log_size: int, log_blowup: int) -> int:
    """Algorithmic bytes of the column pipeline: interpolate 8n + LDE 4n+4m + leaf read 4m."""
    n, m = 1 << log_size, 1 << (log_size + log_blowup)
    return 8 * n + 4 * n + 4 * m + 4 * m
