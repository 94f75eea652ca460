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
    assert hs_lib.verify(fq_case["words"], fq_case["inputs"], fq_case["timestamps"])


@pytest.mark.parametrize("pos_name", ["trace_cap", "state", "opening", "query_leaf", "final_poly", "pow"])
def test_rejects_tampering(hs_lib, fq_case, pos_name):
    w = fq_case["words"].copy()
    pos = {"trace_cap": 22, "state": 12, "opening": 22 + 192 + 5, "query_leaf": 22 + 192 + 2256 + 192 + 7,
           "final_poly": w.size - 3, "pow": w.size - 1}[pos_name]
    w[pos] = (int(w[pos]) + 1) % GL_P
    with pytest.raises(ffi.Pb254Error) as e:
        hs_lib.verify(w, fq_case["inputs"], fq_case["timestamps"])
    assert e.value.code == 7


def test_rejects_wrong_public_inputs(hs_lib, fq_case):
    inp = fq_case["inputs"].copy()
    inp[2, 5] ^= np.uint64(2)  # another base: the native x^s and the input tuple change
    with pytest.raises(ffi.Pb254Error) as e:
        hs_lib.verify(fq_case["words"], inp, fq_case["timestamps"])
    assert e.value.code == 7 and "cross-table" in str(e.value)
    ts = fq_case["timestamps"] + np.uint64(1)
    with pytest.raises(ffi.Pb254Error):
        hs_lib.verify(fq_case["words"], fq_case["inputs"], ts)


def test_rejects_truncated_or_foreign_blob(hs_lib, fq_case):
    with pytest.raises(ffi.Pb254Error):
        hs_lib.verify(fq_case["words"][:-5], fq_case["inputs"], fq_case["timestamps"])
    with pytest.raises(ffi.Pb254Error):
        hs_lib.verify(np.zeros(100, dtype=np.uint64), fq_case["inputs"], fq_case["timestamps"])


def test_accepts_oracle_g1_proof_and_native_results(hs_lib, oracle):
    """G1: the verifier recomputes s * x + offset natively (g1_generate_ctl_values) for the output CTL."""
    inp, ts = I.make_inputs(I.KIND_G1, 2, I.config_seed(70))
    pf, _, _ = oracle.prove_inputs(I.KIND_G1, inp, ts)
    assert hs_lib.verify(pf.words(), inp, ts)
    bad = inp.copy()
    bad[1, 0] ^= np.uint64(1)
    with pytest.raises(ffi.Pb254Error):
        hs_lib.verify(pf.words(), bad, ts)


def test_non_canonical_public_input_is_an_error_not_a_crash(hs_lib, fq_case):
    """x >= p in the batch handed to the verifier: PB254_E_NOT_CANONICAL (raised inside the parallel native-result
    loop of the hostsim build, which must not let an exception escape an OpenMP region)."""
    inp = fq_case["inputs"].copy()
    inp[1, 4:8] = np.uint64(0xFFFFFFFFFFFFFFFF)
    with pytest.raises(ffi.Pb254Error) as e:
        hs_lib.verify(fq_case["words"], inp, fq_case["timestamps"])
    assert e.value.code == 3
