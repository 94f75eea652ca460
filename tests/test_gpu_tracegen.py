"""GPU parity of K1/K2 trace generation against the CPU oracle (bit-exact on every cell), through the
C ABI entry point that replaces generate_trace (g1/scalar_mul_stark.rs:55-69 and the G2 / Fq analogues)."""
import numpy as np
import pytest

from plonky2_bn254_b200 import inputs as I

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,k", [(I.KIND_FQ, 5), (I.KIND_G1, 3), (I.KIND_G2, 2), (I.KIND_G1, 130)])
def test_trace_matches_oracle(gpu_ctx, oracle, kind, k):
    inp, ts = I.make_inputs(kind, k, I.config_seed(20 + kind))
    ts = ts + np.uint64(5)
    ref = oracle.generate_trace(kind, inp, ts)
    got = gpu_ctx.generate_trace(kind, inp, ts)
    assert got.shape == ref.shape
    bad = np.argwhere(got != ref)
    assert len(bad) == 0, bad[:8].tolist()


def test_g1_edge_cases(gpu_ctx, oracle):
    """scalar 0, scalar 2^256-1, scalar 1, and offset == x (the a.x == b.x doubling branch in an adding row,
    src/starks/curves/g1/add.rs:76-95)."""
    inp, ts = I.make_inputs(I.KIND_G1, 4, I.config_seed(99))
    inp[0, 0:4] = 0
    inp[1, 0:4] = np.uint64(0xFFFFFFFFFFFFFFFF)
    inp[2, 0:4] = [1, 0, 0, 0]
    inp[3, 12:20] = inp[3, 4:12]  # offset = x
    ref = oracle.generate_trace(I.KIND_G1, inp, ts)
    got = gpu_ctx.generate_trace(I.KIND_G1, inp, ts)
    assert (got == ref).all()


def test_g1_infinity_is_an_error(gpu_ctx):
    """offset = -x makes row 0 add a point to its negative: unsupported by design (g1/add.rs:49-51)."""
    from plonky2_bn254_b200 import ffi
    inp, ts = I.make_inputs(I.KIND_G1, 1, I.config_seed(98))
    x = [int(v) for v in inp[0, 4:12]]
    y = sum(x[4 + i] << (64 * i) for i in range(4))
    ny = (I.BN254_P - y) % I.BN254_P
    inp[0, 12:16] = inp[0, 4:8]
    inp[0, 16:20] = [(ny >> (64 * i)) & I.MASK64 for i in range(4)]
    with pytest.raises(ffi.Pb254Error) as e:
        gpu_ctx.generate_trace(I.KIND_G1, inp, ts)
    assert e.value.code == 2


def test_non_canonical_coordinate_is_an_error(gpu_ctx):
    from plonky2_bn254_b200 import ffi
    inp, ts = I.make_inputs(I.KIND_FQ, 1, I.config_seed(97))
    inp[0, 4:8] = np.uint64(0xFFFFFFFFFFFFFFFF)
    with pytest.raises(ffi.Pb254Error) as e:
        gpu_ctx.generate_trace(I.KIND_FQ, inp, ts)
    assert e.value.code == 3
