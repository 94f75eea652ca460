"""torchrun: ONE G1 proof of an oversized trace across the GPUs of a node (pb254_prove_sharded over NCCL), BASELINE
config 5 by default (8192 scalar-muls, 2^22 rows x 781 columns).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_prove_bench.py \
        [--instances 8192] [--steps 2] [--check]

Prints one JSON line on rank 0: ms per proof (device time, max over ranks), the stage table, exchanged bytes and, with
--check, whether the proof is byte-identical to the single-GPU proof of rank 0's own context (needs the memory of a
single-GPU proof on top) and accepted by pb254_verify.
"""
import argparse, hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from plonky2_bn254_b200 import dist as D, ffi, inputs as I


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=8192)
    ap.add_argument("--kind", type=int, default=0)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream(device=local)
    ctx = ffi.Context(local, stream=stream.cuda_stream)
    inp, ts = I.make_inputs(args.kind, args.instances, I.config_seed(5))
    n = ctx.L.trace_rows(args.instances, 1 << 16)

    def one():
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        pf, coll = D.prove_sharded(ctx, dist, args.kind, inp, ts, f"cuda:{local}", torch_stream=stream)
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return pf, coll, float(t.item())

    pf, coll, first_ms = one()
    best, stages = None, None
    for _ in range(args.steps):
        pf, coll, ms = one()
        if best is None or ms < best:
            best, stages = ms, ctx.timings()
    w = pf.words()
    sha = hashlib.sha256(w.tobytes()).hexdigest()
    shas = [None] * world
    dist.all_gather_object(shas, sha)
    line = {"what": "one proof across the GPUs of a node (pb254_prove_sharded)", "kind": I.KIND_NAMES[args.kind],
            "instances": args.instances, "trace_rows": int(n), "n_gpus": world, "ms_per_proof": best, "first_call_ms": first_ms,
            "proof_bytes": int(w.size * 8), "identical_on_all_ranks": len(set(shas)) == 1,
            "all_to_all_bytes_sent_per_rank": coll.bytes_all_to_all, "all_gather_bytes_received_per_rank": coll.bytes_all_gather,
            "collective_calls": coll.calls, "stage_ms": {k: round(v, 3) for k, v in stages}}
    ex = sum(v for k, v in stages if "exchange" in k)
    if ex > 0 and world > 1:
        line["exchange_ms"] = ex
        line["all_to_all_gb_s_per_rank"] = coll.bytes_all_to_all / (ex * 1e-3) / 1e9
    if args.check and rank == 0:
        t0 = time.time()
        line["verified"] = bool(ctx.L.verify(args.kind, w, inp, ts))
        line["verify_s"] = time.time() - t0
        try:
            single = ctx.prove(args.kind, inp, ts)
            line["single_gpu_ms"] = sum(ms for _, ms in ctx.timings())
            line["identical_to_single_gpu_proof"] = hashlib.sha256(single.words().tobytes()).hexdigest() == sha
        except Exception as e:  # not enough memory for the single-GPU proof next to the sharded workspace
            line["single_gpu_error"] = str(e)
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
