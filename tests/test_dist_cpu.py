"""world_size-2 gloo test of the N > 1 path's host logic (plonky2_bn254_b200/dist.py): replica assignment,
per-batch determinism, digest gathering and the max-over-ranks reduction. The prover context is the
test-only hostsim build (there is no GPU here); the data path has no collective."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from plonky2_bn254_b200 import build, dist as D, ffi, inputs as I
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "4"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = ffi.Context(0, library=ffi.Library(build.HOSTSIM_LIB))
    asg = D.Assignment(rank, world, total)
    local = D.prove_assigned(ctx, I.KIND_FQ, 1, 9, asg)
    digests = D.gather_digests(dist, local, total)
    t = D.max_over_ranks(dist, 1.0 + rank)
    dist.barrier()
    q.put((rank, asg.batches(), [d.hex() for d in digests], t))
    dist.destroy_process_group()


def test_assignment_covers_every_batch_once():
    from plonky2_bn254_b200 import dist as D
    for world in (1, 2, 3, 8):
        for total in (1, 5, 8, 17):
            seen = sorted(b for r in range(world) for b in D.Assignment(r, world, total).batches())
            assert seen == list(range(total))
    assert D.batch_seed(2, 0) != D.batch_seed(2, 1) != D.batch_seed(3, 0)


def test_two_rank_replicas(oracle):
    from plonky2_bn254_b200 import build, dist as D, inputs as I
    build.build_hostsim()
    world, total = 2, 2
    port = _free_port()
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    procs = [ctxmp.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0] and res[1][1] == [1]
    assert res[0][2] == res[1][2]            # every rank sees all digests
    assert res[0][3] == res[1][3] == 2.0     # max over ranks
    # the gathered digest of a batch is that of the oracle's proof of the same batch (bit-identical proofs); one
    # oracle proof is enough here, every batch is compared in the GPU tier
    assert res[0][2][0] != res[0][2][1]
    for b in range(1):
        inp, ts = D.make_batch(I.KIND_FQ, 1, 9, b)
        pf, _, _ = oracle.prove_inputs(I.KIND_FQ, inp, ts)
        assert D.proof_digest(pf.words()).hex() == res[0][2][b]


def _commit_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from plonky2_bn254_b200 import build, dist as D, ffi
    from util import rand_field
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = ffi.Context(0, library=ffi.Library(build.HOSTSIM_LIB))
    rng = np.random.default_rng(77)
    full = rand_field(rng, (21, 1 << 9))            # every rank regenerates the same matrix, keeps its columns
    first, cnt = D.shard_columns(21, world)[rank]
    shard = torch.from_numpy(full[first:first + cnt].view(np.int64).copy())
    cap = D.dist_commit(ctx, dist if world > 1 else None, shard, 21, 1, 4)
    q.put((rank, cap.numpy().view(np.uint64).tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2, 4])
def test_dist_commit_matches_single_commit(oracle, world):
    """Column-sharded LDE -> all-to-all -> row-sharded leaf hashing -> digest all-gather -> per-rank subtrees gives
    the same Merkle cap as PolynomialBatch::from_values on one device (oracle), for 1, 2 and 4 ranks (gloo; the
    prover context is the hostsim build, whose 'device pointers' are host pointers)."""
    from plonky2_bn254_b200 import build
    from util import rand_field
    build.build_hostsim()
    port = _free_port()
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    procs = [ctxmp.Process(target=_commit_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(77)
    full = rand_field(rng, (21, 1 << 9))
    want = oracle.commit(full, 1, 4).tolist()
    for _, cap in res:
        assert cap == want


def _sharded_worker(rank, world, port, kind, k, rate_bits, q):
    sys.path.insert(0, ROOT)
    import hashlib
    import torch.distributed as dist
    from plonky2_bn254_b200 import build, dist as D, ffi, inputs as I
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = str(max(1, 8 // world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = ffi.Context(0, library=ffi.Library(build.HOSTSIM_LIB))
    inp, ts = I.make_inputs(kind, k, I.config_seed(70 + kind))
    cfg = None
    if rate_bits != 1:
        cfg = ctx.L.standard_fast_config()
        cfg.rate_bits, cfg.num_query_rounds = rate_bits, 28
    pf, coll = D.prove_sharded(ctx, dist, kind, inp, ts, "cpu", config=cfg)
    w = pf.words()
    stages = [name for name, _ in ctx.timings()]
    dist.barrier()
    q.put((rank, hashlib.sha256(w.tobytes()).hexdigest(), int(w.size), coll.calls, coll.bytes_all_to_all, stages))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kind,k,rate_bits", [(2, 2, 130, 1), (4, 2, 2, 2)])
def test_sharded_proof_is_the_single_rank_proof(hostsim_ctx, world, kind, k, rate_bits):
    """pb254_prove_sharded (one proof across `world` ranks: column-sharded LDE, all-to-all, row-block leaf hashing,
    row-block quotient with the next-row halo, row-block FRI combination, owner-supplied query rows) gives, on every
    rank, the bytes of the single-rank proof. gloo + the hostsim build; the GPU tier repeats it over NCCL."""
    from plonky2_bn254_b200 import inputs as I
    port = _free_port()
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    procs = [ctxmp.Process(target=_sharded_worker, args=(r, world, port, kind, k, rate_bits, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=900) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import hashlib
    inp, ts = I.make_inputs(kind, k, I.config_seed(70 + kind))
    cfg = None
    if rate_bits != 1:
        cfg = hostsim_ctx.L.standard_fast_config()
        cfg.rate_bits, cfg.num_query_rounds = rate_bits, 28
    want = hostsim_ctx.prove(kind, inp, ts, config=cfg).words()
    for rank, sha, size, calls, a2a, stages in res:
        assert size == want.size and sha == hashlib.sha256(want.tobytes()).hexdigest(), f"rank {rank}"
        assert calls >= 2 * 3 + 3 and a2a > 0          # per matrix: all-to-all, digests (+ tree levels), openings; quotient, FRI, queries
        assert "exchange trace" in stages and "exchange aux" in stages
        assert ("values exchange trace" in stages) == (k == 130)


def _sharded_error_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from plonky2_bn254_b200 import build, dist as D, ffi, inputs as I
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "4"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = ffi.Context(0, library=ffi.Library(build.HOSTSIM_LIB))
    inp, ts = I.make_inputs(I.KIND_FQ, 130, I.config_seed(72))
    inp = inp.copy()
    inp[129, 4:8] = np.uint64(0xFFFFFFFFFFFFFFFF)      # x >= p in an instance that only rank 1 generates
    code = None
    try:
        D.prove_sharded(ctx, dist, I.KIND_FQ, inp, ts, "cpu")
    except ffi.Pb254Error as e:
        code = e.code
    dist.barrier()
    q.put((rank, code))
    dist.destroy_process_group()


def test_sharded_input_error_fails_on_every_rank():
    """A non-canonical coordinate in an instance of rank 1's block: the error code is all-gathered after trace
    generation, so rank 0 returns PB254_E_NOT_CANONICAL too instead of waiting in the next collective."""
    from plonky2_bn254_b200 import build
    build.build_hostsim()
    world, port = 2, _free_port()
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    procs = [ctxmp.Process(target=_sharded_error_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, 3), (1, 3)]
