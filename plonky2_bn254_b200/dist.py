"""Multi-GPU host logic: replica sharding of independent proof batches.

The path shards by independent proofs - every proof has its own Fiat-Shamir transcript, so GPU g proves
batch g end to end and there is NO data-path collective. `torch.distributed` (NCCL on the GPU box, gloo in
the CPU tests) is used only for (i) the barrier around the timed region, (ii) the max-over-ranks time and
(iii) gathering the proof blobs / digests to rank 0, which is what the reference's caller (the outer
plonky2 witness generation, single process) would consume.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass

import numpy as np

from . import inputs as I


@dataclass(frozen=True)
class Assignment:
    """Which proof batches a rank proves: batch ids `first, first + world, ...` below `total`."""
    rank: int
    world: int
    total: int

    def batches(self):
        return list(range(self.rank, self.total, self.world))


def batch_seed(config_id: int, batch: int) -> int:
    """Deterministic seed of proof batch `batch` of a config: every rank can regenerate any batch."""
    return I.config_seed(config_id) + 1000 * (batch + 1)


def make_batch(kind: int, instances: int, config_id: int, batch: int):
    return I.make_inputs(kind, instances, batch_seed(config_id, batch))


def prove_assigned(ctx, kind: int, instances: int, config_id: int, assignment: Assignment):
    """Proves this rank's batches on its own context; returns {batch id: proof words (np.uint64)}."""
    out = {}
    for b in assignment.batches():
        inp, ts = make_batch(kind, instances, config_id, b)
        out[b] = ctx.prove(kind, inp, ts).words()
    return out


def proof_digest(words: np.ndarray) -> bytes:
    return hashlib.sha256(np.ascontiguousarray(words, dtype=np.uint64).tobytes()).digest()


def gather_digests(dist, local: dict, total: int, device=None):
    """All ranks contribute {batch: words}; returns on every rank the list of sha256 digests by batch id.
    (Digests, not blobs: 32 B per proof is enough to check completeness and determinism.)"""
    import torch
    buf = torch.zeros((total, 32), dtype=torch.uint8, device=device)
    for b, w in local.items():
        buf[b] = torch.frombuffer(bytearray(proof_digest(w)), dtype=torch.uint8).to(buf.device)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)  # every row is written by exactly one rank
    return [bytes(buf[b].cpu().numpy().tobytes()) for b in range(total)]


def max_over_ranks(dist, seconds: float, device=None) -> float:
    import torch
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- oversized single trace: column-sharded LDE -> all-to-all -> row-sharded Merkle -------------
def shard_columns(total_cols: int, world: int):
    """Contiguous column ranges in rank order; [(first, count)] per rank."""
    return [(total_cols * g // world, total_cols * (g + 1) // world - total_cols * g // world) for g in range(world)]


def dist_commit(ctx, dist, shard, total_cols: int, rate_bits: int, cap_height: int, timings: dict | None = None):
    """PolynomialBatch::from_values of ONE matrix spread over `world` ranks; returns the Merkle cap
    (2^cap_height x 4, torch int64 tensor on the shard's device), identical on every rank and identical to the
    single-GPU commitment.

    `shard`: this rank's columns as a contiguous torch.int64 tensor [cols_g, n] (canonical field elements) on
    the context's device, column ranges as given by `shard_columns`. `dist` is torch.distributed (NCCL on GPUs,
    gloo in the CPU tests) or None for a single rank.

      A  LDE of the local columns (pb254_lde_dev), no communication
      B  ONE all-to-all: rank h receives the row block [h N/G, (h+1) N/G) of every column
      C  row-local leaf hashing (pb254_leaf_hash_rows_dev), digests in natural row order
      D  all-gather of the 32-byte digests (N x 32 B)
      E  every rank builds the subtree under its own cap entries (pb254_merkle_subtree_dev); all-gather of the cap
    """
    import time
    import torch
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    cols_g, n = shard.shape
    N = n << rate_bits
    log_N = N.bit_length() - 1
    log_g = world.bit_length() - 1
    assert (1 << log_g) == world and world <= (1 << cap_height) and N % world == 0
    assert shard.dtype == torch.int64 and shard.is_contiguous()
    counts = [c for _, c in shard_columns(total_cols, world)]
    assert counts[rank] == cols_g
    rows = N // world
    dev = shard.device
    sync = torch.cuda.synchronize if dev.type == "cuda" else (lambda: None)
    t = {}

    def mark(name, t0):
        sync()
        t[name] = (time.perf_counter() - t0) * 1e3

    t0 = time.perf_counter()
    lde = torch.empty((cols_g, N), dtype=torch.int64, device=dev)
    ctx.lde_dev(shard.data_ptr(), cols_g, n, rate_bits, lde.data_ptr())
    mark("lde (column shard)", t0)
    t0 = time.perf_counter()
    if world > 1:
        send = lde.view(cols_g, world, rows).permute(1, 0, 2).contiguous()  # [peer][col][row block]
        del lde
        recv = torch.empty((total_cols, rows), dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv.view(-1), send.view(-1), output_split_sizes=[c * rows for c in counts],
                               input_split_sizes=[cols_g * rows] * world)
        del send
    else:
        recv = lde
    mark("all-to-all (columns -> row blocks)", t0)
    t0 = time.perf_counter()
    dig = torch.empty((rows, 4), dtype=torch.int64, device=dev)
    ctx.leaf_hash_rows_dev(recv.data_ptr(), rows, total_cols, rows, dig.data_ptr())
    mark("leaf hash (row block)", t0)
    t0 = time.perf_counter()
    if world > 1:
        all_dig = torch.empty((N, 4), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_dig, dig)
    else:
        all_dig = dig
    mark("all-gather digests", t0)
    t0 = time.perf_counter()
    log_roots = cap_height - log_g
    roots = torch.empty((1 << log_roots, 4), dtype=torch.int64, device=dev)
    ctx.merkle_subtree_dev(all_dig.data_ptr(), log_N, rank << (log_N - log_g), log_N - log_g, log_roots,
                           roots.data_ptr())
    if world > 1:
        cap = torch.empty((1 << cap_height, 4), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(cap, roots)
    else:
        cap = roots
    mark("subtree + cap all-gather", t0)
    if timings is not None:
        timings.update(t)
    return cap


# ---- ONE proof across the GPUs of a node (pb254_prove_sharded, include/pb254.h) -------------------------------------
class TorchCollectives:
    """The two collectives pb254_prove_sharded calls back for, on raw pointers, through torch.distributed.

    CUDA (NCCL): the pointers are device pointers on the context's GPU; they are wrapped without a copy through
    ``__cuda_array_interface__`` and the collectives are issued with the context's stream current, so they are ordered
    with the prover's kernels like any other work of that stream. CPU (gloo, the hostsim build of the tests): host
    pointers wrapped with ``torch.frombuffer``."""

    def __init__(self, dist, device, torch_stream=None):
        import torch
        self.torch, self.dist, self.device, self.stream = torch, dist, torch.device(device), torch_stream
        self.bytes_all_to_all = 0
        self.bytes_all_gather = 0
        self.calls = 0

    def _view(self, ptr: int, nbytes: int):
        torch = self.torch
        if self.device.type == "cuda":
            class _Buf:
                pass
            b = _Buf()
            b.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
            return torch.as_tensor(b, device=self.device)
        import ctypes
        return torch.frombuffer((ctypes.c_char * nbytes).from_address(ptr), dtype=torch.uint8)

    def _on_stream(self):
        import contextlib
        if self.device.type == "cuda" and self.stream is not None:
            return self.torch.cuda.stream(self.stream)
        return contextlib.nullcontext()

    def all_to_all(self, send: int, recv: int, bytes_per_peer: int):
        world = self.dist.get_world_size()
        with self._on_stream():
            self.dist.all_to_all_single(self._view(recv, bytes_per_peer * world), self._view(send, bytes_per_peer * world))
        self.bytes_all_to_all += bytes_per_peer * (world - 1)
        self.calls += 1

    def all_gather(self, send: int, recv: int, bytes_per_rank: int):
        world = self.dist.get_world_size()
        with self._on_stream():
            self.dist.all_gather_into_tensor(self._view(recv, bytes_per_rank * world), self._view(send, bytes_per_rank))
        self.bytes_all_gather += bytes_per_rank * (world - 1)
        self.calls += 1


def prove_sharded(ctx, dist, kind: int, inputs, timestamps, device, torch_stream=None, config=None, min_rows=1 << 16):
    """One proof of (inputs, timestamps) across all ranks of `dist`; every rank gets the same proof, byte-identical to
    ctx.prove on one GPU. Returns (Proof, TorchCollectives) - the latter carries the exchanged byte counts."""
    coll = TorchCollectives(dist, device, torch_stream)
    pf = ctx.prove_sharded(kind, inputs, timestamps, dist.get_rank(), dist.get_world_size(), coll.all_to_all,
                           coll.all_gather, min_rows=min_rows, config=config)
    return pf, coll
