"""ONE proof across two GPUs (pb254_prove_sharded over NCCL) is byte-identical to the single-GPU proof, which
tests/test_gpu_prover.py pins to the oracle. Needs two GPUs (the driver's single-GPU box skips it;
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu` runs it)."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cases, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from plonky2_bn254_b200 import dist as D, ffi, inputs as I
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    stream = torch.cuda.Stream(device=rank)
    ctx = ffi.Context(rank, stream=stream.cuda_stream)
    out = []
    for kind, k, rate_bits in cases:
        inp, ts = I.make_inputs(kind, k, I.config_seed(70 + kind))
        cfg = None
        if rate_bits != 1:
            cfg = ctx.L.standard_fast_config()
            cfg.rate_bits, cfg.num_query_rounds = rate_bits, 28
        pf, coll = D.prove_sharded(ctx, dist, kind, inp, ts, f"cuda:{rank}", torch_stream=stream, config=cfg)
        w = pf.words()
        single = ctx.prove(kind, inp, ts, config=cfg).words() if rank == 0 else None
        out.append((hashlib.sha256(w.tobytes()).hexdigest(), int(w.size),
                    hashlib.sha256(single.tobytes()).hexdigest() if single is not None else None, coll.bytes_all_to_all))
    torch.cuda.synchronize()
    dist.barrier()
    q.put((rank, out))
    dist.destroy_process_group()


def test_sharded_proof_over_nccl():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    cases = [(2, 3, 1), (0, 130, 1), (1, 2, 1), (2, 5, 3), (2, 300, 1), (1, 129, 1)]
    # 2^16 rows (trace replicated): fq, G2, fq with blow-up 8; >= 2^17 rows (instance-sharded trace generation): G1, fq, G2
    port = _free_port()
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    procs = [ctxmp.Process(target=_worker, args=(r, world, port, cases, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=900) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for i, case in enumerate(cases):
        sha0, size0, single, a2a = res[0][i]
        assert sha0 == single, f"sharded proof differs from the single-GPU proof for {case}"
        assert res[1][i][0] == sha0 and res[1][i][1] == size0 and a2a > 0
