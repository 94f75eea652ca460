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
