"""Oversized single trace (BASELINE config 5 shape): commit ONE W x 2^log_n matrix across the GPUs of a node with
column-sharded LDE -> NCCL all-to-all -> row-sharded leaf hashing -> digest all-gather -> per-rank subtrees
(plonky2_bn254_b200/dist.py::dist_commit). Run under torchrun; rank 0 prints one JSON line with the per-stage
times (max over ranks) and a digest of the cap, which must not depend on the number of ranks.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_commit_bench.py \
        [--cols 781] [--log-n 22] [--reps 3]
"""
import argparse
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from plonky2_bn254_b200 import dist as D, ffi


def column(c, n, device):
    """column c of the synthetic matrix: canonical field elements (< 2^63 < p), a function of c only"""
    g = torch.Generator(device=device)
    g.manual_seed(0x706232353405 + c)
    return torch.randint(0, 1 << 62, (n,), dtype=torch.int64, device=device, generator=g)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cols", type=int, default=781)
    ap.add_argument("--log-n", type=int, default=22)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=dev)
        dist = dist_mod
    stream = torch.cuda.current_stream()
    ctx = ffi.Context(local, stream=stream.cuda_stream)
    n = 1 << args.log_n
    first, cnt = D.shard_columns(args.cols, world)[rank]
    shard = torch.stack([column(first + c, n, dev) for c in range(cnt)])
    torch.cuda.synchronize()
    best = None
    for rep in range(args.reps):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t = {}
        t0 = time.perf_counter()
        cap = D.dist_commit(ctx, dist, shard, args.cols, 1, 4, timings=t)
        torch.cuda.synchronize()
        t["total"] = (time.perf_counter() - t0) * 1e3
        if dist is not None:  # max over ranks, stage by stage
            keys = sorted(t)
            v = torch.tensor([t[k] for k in keys], dtype=torch.float64, device=dev)
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
            t = dict(zip(keys, v.tolist()))
        if best is None or t["total"] < best["total"]:
            best = t
    if rank == 0:
        N = 2 * n
        bytes_lde = 8 * args.cols * n * 3
        print(json.dumps({
            "what": "dist_commit", "n_gpus": world, "cols": args.cols, "log_n": args.log_n, "rate_bits": 1,
            "ms": {k: round(v, 3) for k, v in best.items()},
            "all_to_all_bytes_per_rank": 8 * cnt * N * (world - 1) // world,
            "digest_allgather_bytes": 32 * N,
            "cap_sha256": hashlib.sha256(cap.cpu().numpy().tobytes()).hexdigest(),
        }), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
