"""Host orchestration + kernel bodies of the product sources, compiled for the CPU (test-only hostsim
build, tests/hostsim/), against the oracle: transcript, proof assembly and every per-thread kernel body
are the same code the GPU runs. The GPU parity tests proper are the `-m gpu` tests."""
import numpy as np
import pytest

from plonky2_bn254_b200 import inputs as I


def test_hostsim_trace_matches_oracle(hostsim_ctx, oracle):
    for kind, k in ((I.KIND_G1, 2), (I.KIND_FQ, 3)):
        inp, ts = I.make_inputs(kind, k, I.config_seed(30 + kind))
        assert (hostsim_ctx.generate_trace(kind, inp, ts) == oracle.generate_trace(kind, inp, ts)).all()


def test_hostsim_proof_is_byte_identical(hostsim_ctx, oracle, fq_case):
    pf = hostsim_ctx.prove(fq_case["kind"], fq_case["inputs"], fq_case["timestamps"], keep_debug=True)
    w = pf.words()
    ref = fq_case["proof"]
    assert (pf.debug(2) == ref.debug(2)).all()  # challenges
    assert (pf.debug(0) == ref.debug(0)).all()  # auxiliary columns
    assert (pf.debug(1) == ref.debug(1)).all()  # quotient chunks
    assert w.size == fq_case["words"].size and (w == fq_case["words"]).all()
    # the native outputs x^s, read from the trace (pb254_proof_results)
    res = pf.results().reshape(3, -1)
    for i in range(3):
        assert (res[i] == oracle.native_result(fq_case["kind"], fq_case["inputs"][i])).all()
    # prove_trace (the literal prove() signature on a host trace) gives the same bytes
    pf2 = hostsim_ctx.prove_trace(fq_case["kind"], fq_case["trace"])
    assert (pf2.words() == w).all()


def test_hostsim_commit_matches_oracle(hostsim_ctx, oracle):
    from util import rand_field
    rng = np.random.default_rng(11)
    v = rand_field(rng, (9, 1 << 10))
    cap_o, dig_o = oracle.commit(v, 1, 4, want_digests=True)
    cap_h, dig_h = hostsim_ctx.commit(v, 1, 4, want_digests=True)
    assert (cap_o == cap_h).all() and (dig_o == dig_h).all()
    coeffs, lde = oracle.lde_batch(v, 1)
    assert (hostsim_ctx.lde_batch(v, 1) == lde).all()
    assert (hostsim_ctx.lde_batch(coeffs, 1, from_coeffs=True) == lde).all()



def test_prove_many_equals_prove(hostsim_ctx, fq_case):
    """pb254_prove_many (independent proofs round-robin over several contexts, one host thread each) returns, in batch
    order, exactly the proofs pb254_prove gives for each batch (batch 0: the session's oracle proof, byte-identical by
    parity; batch 1: proved again on one context). The failure path is in tests/test_gpu_prover.py."""
    from plonky2_bn254_b200 import ffi
    batches = [(fq_case["inputs"], fq_case["timestamps"]), I.make_inputs(I.KIND_FQ, 3, I.config_seed(121))]
    ctx2 = ffi.Context(0, library=hostsim_ctx.L)
    many = ffi.prove_many([hostsim_ctx, ctx2], I.KIND_FQ, batches)
    ctx2.close()
    assert len(many) == 2
    want = [fq_case["words"], hostsim_ctx.prove(I.KIND_FQ, *batches[1]).words()]
    for b, pf in enumerate(many):
        got = pf.words()
        assert got.size == want[b].size and (got == want[b]).all(), f"batch {b}"
