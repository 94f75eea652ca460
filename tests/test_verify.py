"""pb254_verify (the product's host verifier, csrc/verify.cuh) against proofs from the oracle: accepts valid
proofs of all three STARKs, rejects tampering and wrong public inputs. Runs through the hostsim build here
(verification is host code; the same function is exported by the CUDA library and exercised by the GPU tests)."""
import numpy as np
import pytest

from plonky2_bn254_b200 import ffi, inputs as I
from util import GL_P


@pytest.fixture(scope="module")
def hs_lib():
    from plonky2_bn254_b200 import build
    return ffi.Library(build.build_hostsim())


def test_accepts_oracle_fq_proof(hs_lib, fq_case):
    assert hs_lib.verify(I.KIND_FQ, fq_case["words"], fq_case["inputs"], fq_case["timestamps"])


@pytest.mark.parametrize("pos_name", ["trace_cap", "state", "opening", "query_leaf", "final_poly", "pow"])
def test_rejects_tampering(hs_lib, fq_case, pos_name):
    w = fq_case["words"].copy()
    pos = {"trace_cap": 22, "state": 12, "opening": 22 + 192 + 5, "query_leaf": 22 + 192 + 2256 + 192 + 7,
           "final_poly": w.size - 3, "pow": w.size - 1}[pos_name]
    w[pos] = (int(w[pos]) + 1) % GL_P
    with pytest.raises(ffi.Pb254Error) as e:
        hs_lib.verify(I.KIND_FQ, w, fq_case["inputs"], fq_case["timestamps"])
    assert e.value.code == 7


def test_rejects_wrong_public_inputs(hs_lib, fq_case):
    inp = fq_case["inputs"].copy()
    inp[2, 5] ^= np.uint64(2)  # another base: the native x^s and the input tuple change
    with pytest.raises(ffi.Pb254Error) as e:
        hs_lib.verify(I.KIND_FQ, fq_case["words"], inp, fq_case["timestamps"])
    assert e.value.code == 7 and "cross-table" in str(e.value)
    ts = fq_case["timestamps"] + np.uint64(1)
    with pytest.raises(ffi.Pb254Error):
        hs_lib.verify(I.KIND_FQ, fq_case["words"], fq_case["inputs"], ts)


def test_rejects_truncated_or_foreign_blob(hs_lib, fq_case):
    with pytest.raises(ffi.Pb254Error):
        hs_lib.verify(I.KIND_FQ, fq_case["words"][:-5], fq_case["inputs"], fq_case["timestamps"])
    with pytest.raises(ffi.Pb254Error):
        hs_lib.verify(I.KIND_FQ, np.zeros(100, dtype=np.uint64), fq_case["inputs"], fq_case["timestamps"])


def test_accepts_oracle_g1_proof_and_native_results(hs_lib, oracle):
    """G1: the verifier recomputes s * x + offset natively (g1_generate_ctl_values) for the output CTL."""
    inp, ts = I.make_inputs(I.KIND_G1, 2, I.config_seed(70))
    pf, _, _ = oracle.prove_inputs(I.KIND_G1, inp, ts)
    assert hs_lib.verify(I.KIND_G1, pf.words(), inp, ts)
    bad = inp.copy()
    bad[1, 0] ^= np.uint64(1)
    with pytest.raises(ffi.Pb254Error):
        hs_lib.verify(I.KIND_G1, pf.words(), bad, ts)


def test_non_canonical_public_input_is_an_error_not_a_crash(hs_lib, fq_case):
    """x >= p in the batch handed to the verifier: PB254_E_NOT_CANONICAL (raised inside the parallel native-result
    loop of the hostsim build, which must not let an exception escape an OpenMP region)."""
    inp = fq_case["inputs"].copy()
    inp[1, 4:8] = np.uint64(0xFFFFFFFFFFFFFFFF)
    with pytest.raises(ffi.Pb254Error) as e:
        hs_lib.verify(I.KIND_FQ, fq_case["words"], inp, fq_case["timestamps"])
    assert e.value.code == 3


@pytest.mark.parametrize("field,value", [(6, 1), (7, 0), (5, 1), (3, 2), (4, 3), (8, 3), (9, 4)])
def test_rejects_weakened_header(hs_lib, fq_case, field, value):
    """The security parameters are the verifier's, not the proof's (verifier.rs:32-45 takes the StarkConfig from
    the caller): a header that claims 1 query round / no grinding / 1 challenge / another rate ... is rejected
    before anything else is looked at."""
    w = fq_case["words"].copy()
    w[field] = np.uint64(value)
    with pytest.raises(ffi.Pb254Error) as e:
        hs_lib.verify(I.KIND_FQ, w, fq_case["inputs"], fq_case["timestamps"])
    assert e.value.code == 7 and "header" in str(e.value)


def test_rejects_wrong_kind_and_foreign_config(hs_lib, fq_case):
    """A proof of one STARK presented as another (the caller's buffer is sized for ITS kind: no read past it), and a
    valid proof checked under a stricter caller configuration."""
    w = fq_case["words"]
    g1_inp, g1_ts = I.make_inputs(I.KIND_G1, 3, I.config_seed(71))
    with pytest.raises(ffi.Pb254Error) as e:
        hs_lib.verify(I.KIND_G1, w, g1_inp, g1_ts)
    assert e.value.code == 7
    w2 = w.copy()
    w2[1] = np.uint64(I.KIND_G2)  # the blob claims G2 (36 words per instance) against an Fq buffer (8 words)
    with pytest.raises(ffi.Pb254Error) as e:
        hs_lib.verify(I.KIND_FQ, w2, fq_case["inputs"], fq_case["timestamps"])
    assert e.value.code == 7
    with pytest.raises(ffi.Pb254Error) as e:  # wrong shape for the caller's kind: refused on the Python side
        hs_lib.verify(I.KIND_G2, w2, fq_case["inputs"], fq_case["timestamps"])
    assert e.value.code == 6
    cfg = hs_lib.standard_fast_config()
    cfg.num_query_rounds = 100
    with pytest.raises(ffi.Pb254Error) as e:
        hs_lib.verify(I.KIND_FQ, w, fq_case["inputs"], fq_case["timestamps"], config=cfg)
    assert e.value.code == 7
    assert hs_lib.verify(I.KIND_FQ, w, fq_case["inputs"], fq_case["timestamps"], config=hs_lib.standard_fast_config())


def test_rejects_more_instances_than_periods(hs_lib, fq_case):
    inp, ts = I.make_inputs(I.KIND_FQ, 129, I.config_seed(72))  # a 2^16-row trace holds 128 instances
    with pytest.raises(ffi.Pb254Error) as e:
        hs_lib.verify(I.KIND_FQ, fq_case["words"], inp, ts)
    assert e.value.code == 7
