"""Oracle prover / verifier: the trace and auxiliary columns satisfy every constraint
on the trace domain, the oracle's verifier accepts the oracle's proof and rejects tampered ones, and the
proof bytes are pinned (golden)."""
import hashlib

import numpy as np
import pytest

from util import GL_P


def test_constraints_vanish_on_the_trace_domain(oracle, fq_case):
    """What the quotient argument proves: for every row i of H, all 910 constraints are zero
    (check_constraints-style test of the reference, exp_stark.rs tests)."""
    tr = fq_case["trace"]
    n = tr.shape[1]
    dbg = fq_case["proof"].debug(2)
    betas, gammas = dbg[0:2], dbg[2:4]
    aux = fq_case["proof"].debug(0).reshape(-1, n)
    assert aux.shape[0] == oracle.num_aux(fq_case["kind"])
    g = oracle.lib().orc_gl_root_of_unity(16)
    ginv = pow(g, -1, GL_P)
    alphas = np.array([0x123456789ABCDEF, 0xFEDCBA987654321], dtype=np.uint64)
    rows = [0, 1, 2, 3, 510, 511, 512, 513, 1023, 1535, 1536, 4000, 65534, n - 1]
    for i in rows:
        j = (i + 1) % n
        x = pow(g, i, GL_P)
        acc, cnt = oracle.eval_constraints_base(
            fq_case["kind"], tr[:, i].copy(), tr[:, j].copy(), aux[:, i].copy(), aux[:, j].copy(), betas, gammas,
            alphas, (x - ginv) % GL_P, 1 if i == 0 else 0, 1 if i == n - 1 else 0)
        assert cnt == 910
        assert not acc.any(), i
    # and a corrupted cell breaks them
    bad = tr[:, 5].copy()
    bad[64] ^= np.uint64(1)
    acc, _ = oracle.eval_constraints_base(fq_case["kind"], bad, tr[:, 6].copy(), aux[:, 5].copy(), aux[:, 6].copy(),
                                          betas, gammas, alphas, (pow(g, 5, GL_P) - ginv) % GL_P, 0, 0)
    assert acc.any()


def test_verifier_accepts_and_proof_is_pinned(oracle, fq_case, golden):
    w = fq_case["words"]
    assert oracle.verify(w, fq_case["inputs"], fq_case["timestamps"])
    g = golden["proofs"][0]
    assert int(w.size) == g["words"] and int(w[-1]) == g["pow_witness"]
    assert hashlib.sha256(w.tobytes()).hexdigest() == g["proof_sha256"]


def test_proof_layout(oracle, fq_case):
    """Header {magic, kind, degree_bits, config[7]}, init_challenger_state[12], then C.7 field order."""
    w = fq_case["words"]
    assert int(w[1]) == 2 and int(w[2]) == 16
    assert [int(x) for x in w[3:10]] == [1, 4, 2, 84, 16, 4, 5]
    W, A, Q = 427, 134, 4
    n_open = 2 * (2 * W) + 2 * (2 * A) + 4 + 2 * Q
    arities = 3  # 2^16 -> [4, 4, 4]
    sib = 4 * (17 - 4)
    per_query = (W + sib) + (A + sib) + (Q + sib) + sum(32 + 4 * (17 - 4 * (k + 1) - 4) for k in range(arities))
    final = 2 * (1 << (16 - 12))
    assert w.size == 10 + 12 + 3 * 64 + n_open + arities * 64 + 84 * per_query + final + 1


@pytest.mark.parametrize("where", ["trace_cap", "opening", "query_leaf", "final_poly", "pow"])
def test_verifier_rejects_tampering(oracle, fq_case, where):
    w = fq_case["words"].copy()
    pos = {"trace_cap": 22, "opening": 22 + 192 + 5, "query_leaf": 22 + 192 + 2256 + 192 + 7,
           "final_poly": w.size - 3, "pow": w.size - 1}[where]
    w[pos] = (int(w[pos]) + 1) % GL_P
    with pytest.raises(oracle.OracleError):
        oracle.verify(w, fq_case["inputs"], fq_case["timestamps"])


def test_verifier_rejects_wrong_public_inputs(oracle, fq_case):
    inp = fq_case["inputs"].copy()
    inp[1, 0] ^= np.uint64(1)  # another exponent: the CTL sums no longer match
    with pytest.raises(oracle.OracleError):
        oracle.verify(fq_case["words"], inp, fq_case["timestamps"])
