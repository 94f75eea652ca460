import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def gpu_ctx():
    """Product library on cuda:0. Fails loudly (no fallback) if the library or the GPU is missing."""
    from plonky2_bn254_b200 import ffi
    ctx = ffi.Context(0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def hostsim_ctx():
    """Test-only host-simulation build of the same sources (host logic + per-thread kernel bodies)."""
    from plonky2_bn254_b200 import build, ffi
    path = build.build_hostsim()
    ctx = ffi.Context(0, library=ffi.Library(path))
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def fq_case(oracle):
    """One fq_exp batch (3 exponentiations, 2^16 rows) proved by the oracle with debug artefacts; shared by
    the CPU tests (an oracle proof of the smallest admissible trace takes tens of seconds on 8 cores)."""
    from plonky2_bn254_b200 import inputs as I
    inp, ts = I.make_inputs(I.KIND_FQ, 3, I.config_seed(42))
    trace = oracle.generate_trace(I.KIND_FQ, inp, ts)
    proof = oracle.prove(I.KIND_FQ, trace, keep_debug=True)
    return {"kind": I.KIND_FQ, "inputs": inp, "timestamps": ts, "trace": trace, "proof": proof,
            "words": proof.words()}
