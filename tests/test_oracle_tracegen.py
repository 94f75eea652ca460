"""Pins the oracle's witness generation the way the reference's own tests do:
column offsets (row_position_correctness, g1/scalar_mul_view.rs:97-117), witness soundness
(assert_modulus_zero, modular/modulus_zero.rs:121-160) and the native result
(assert_eq!(expected_output, output), g1/scalar_mul_stark.rs:105-108) - here against an independent
big-integer implementation in Python."""
import hashlib

import numpy as np
import pytest

from plonky2_bn254_b200 import inputs as I
from util import GL_P

P = I.BN254_P
P16 = [(P >> (16 * i)) & 0xFFFF for i in range(16)]


def limbs_to_int(l):
    return sum(int(v) << (16 * i) for i, v in enumerate(l))


def words_to_int(w):
    return sum(int(v) << (64 * i) for i, v in enumerate(w))


def signed(v):
    v = int(v)
    return v - GL_P if v > GL_P // 2 else v


def check_modulus_zero_aux(inp, aux80):
    """sign * quot_abs (*) p + (x - 2^16) (lo - 2^29 + 2^16 hi) - input == 0 coefficient-wise (B.2 step 6)."""
    sign = 2 * int(aux80[0]) - 1
    q = [sign * int(v) for v in aux80[1:18]]
    lo, hi = aux80[18:49], aux80[49:80]
    assert all(0 <= int(v) < 65536 for v in aux80)
    ap = [int(lo[i]) - (1 << 29) + (int(hi[i]) << 16) for i in range(31)]
    for k in range(32):
        c = sum(q[i] * P16[k - i] for i in range(17) if 0 <= k - i < 16)
        c += (ap[k - 1] if k >= 1 else 0) - ((1 << 16) * ap[k] if k < 31 else 0)
        c -= inp[k] if k < 31 else 0
        assert c == 0, k


def test_generate_modulus_zero_witness(oracle):
    rng = np.random.default_rng(5)
    for trial in range(20):
        a = [int(x) for x in rng.integers(0, 65536, 16)]
        b = [int(x) for x in rng.integers(0, 65536, 16)]
        va, vb = limbs_to_int(a) % P, limbs_to_int(b) % P
        a = [(va >> (16 * i)) & 0xFFFF for i in range(16)]
        b = [(vb >> (16 * i)) & 0xFFFF for i in range(16)]
        c = va * vb % P
        if trial % 3 == 0:   # a * b - c, positive quotient
            cl = [(c >> (16 * i)) & 0xFFFF for i in range(16)]
            inp = [sum(a[i] * b[k - i] for i in range(16) if 0 <= k - i < 16) - (cl[k] if k < 16 else 0)
                   for k in range(31)]
        elif trial % 3 == 1:  # c - a * b, negative quotient
            cl = [(c >> (16 * i)) & 0xFFFF for i in range(16)]
            inp = [(cl[k] if k < 16 else 0) - sum(a[i] * b[k - i] for i in range(16) if 0 <= k - i < 16)
                   for k in range(31)]
        else:                 # zero quotient: a - a
            inp = [0] * 31
        aux = oracle.gen_modulus_zero(inp)
        check_modulus_zero_aux(inp, aux)
        if trial % 3 == 2:
            assert int(aux[0]) == 0 and not aux[1:18].any()


def test_generate_is_modulus_zero(oracle):
    rng = np.random.default_rng(6)
    a = [int(x) for x in rng.integers(-65535, 65536, 16)]
    z, aux = oracle.gen_is_modulus_zero(a)
    v = limbs_to_int(a) % P
    assert z == 0
    assert limbs_to_int(aux[:16]) == pow(v, -1, P)
    z, aux = oracle.gen_is_modulus_zero([0] * 16)
    assert z == 1 and not aux[:16].any()
    z, _ = oracle.gen_is_modulus_zero(P16)  # the modulus itself is zero mod p
    assert z == 1


def test_column_offsets(oracle):
    """SURVEY.md Appendix A.1-A.3 (the reference pins these with row_position_correctness tests)."""
    assert [oracle.width(k) for k in (0, 1, 2)] == [781, 1295, 427]
    assert [oracle.num_aux(k) for k in (0, 1, 2)] == [456, 906, 134]
    assert [oracle.reg_len(k) for k in (0, 1, 2)] == [32, 64, 16]


def py_scalar_mul_g1(s, x, off):
    F = I._Fq
    acc = (F.one, F.one, F.zero)
    for i in reversed(range(256)):
        acc = I._jac_double(F, acc)
        if (s >> i) & 1:
            acc = I._jac_add_affine(F, acc, x)
    acc = I._jac_add_affine(F, acc, off)
    return I._to_affine(F, acc)


def test_g1_trace_schedule_and_result(oracle, golden):
    g = [t for t in golden["traces"] if t["kind"] == I.KIND_G1][0]
    inp, ts = I.make_inputs(I.KIND_G1, g["instances"], I.config_seed(g["config_id"]))
    tr, res = oracle.generate_trace(I.KIND_G1, inp, ts, want_results=True)
    assert hashlib.sha256(inp.tobytes()).hexdigest() == g["inputs_sha256"]
    assert hashlib.sha256(tr.tobytes()).hexdigest() == g["trace_sha256"]
    assert list(tr.shape) == g["shape"] and res.tolist() == g["result_limbs"]
    for k in range(g["instances"]):
        s = words_to_int(inp[k, 0:4])
        x = (words_to_int(inp[k, 4:8]), words_to_int(inp[k, 8:12]))
        off = (words_to_int(inp[k, 12:16]), words_to_int(inp[k, 16:20]))
        want = py_scalar_mul_g1(s, x, off)
        r0 = 512 * k
        # sum (cols 32..64) at the last row of the instance is s * x + offset
        assert limbs_to_int(tr[32:48, r0 + 511]) == want[0] and limbs_to_int(tr[48:64, r0 + 511]) == want[1]
        assert limbs_to_int(res[k][:16]) == want[0]
        # row 0: double = b = x, a = offset, bits = le_bits(s), is_adding = 1, flags
        assert limbs_to_int(tr[0:16, r0]) == x[0] and limbs_to_int(tr[96:112, r0]) == x[0]
        assert limbs_to_int(tr[64:80, r0]) == off[0]
        assert [int(b) for b in tr[514:770, r0]] == [(s >> i) & 1 for i in range(256)]
        assert tr[770, r0] == 1 and tr[771, r0 + 511] == 1 and tr[776, r0] == 1 and tr[776, r0 + 1] == 0
        assert (tr[772, r0:r0 + 512] == np.arange(512)).all() and (tr[775, r0:r0 + 512] == ts[k]).all()
        # bits rotate left by one on every adding row
        assert [int(b) for b in tr[514:770, r0 + 2]] == [(s >> ((i + 1) % 256)) & 1 for i in range(256)]
        # every modulus-zero witness on a few rows satisfies its identity (x_aux: lambda^2 - a.x - b.x - c.x)
        for r in (r0, r0 + 1, r0 + 2, r0 + 511):
            lam = [int(v) for v in tr[258:274, r]]
            ax, bx, cx = ([int(v) for v in tr[c:c + 16, r]] for c in (64, 96, 128))
            inp31 = [sum(lam[i] * lam[j - i] for i in range(16) if 0 <= j - i < 16)
                     - ((ax[j] + bx[j] + cx[j]) if j < 16 else 0) for j in range(31)]
            check_modulus_zero_aux(inp31, tr[354:434, r])
    # padding rows are zero except the two range-check columns; range_counter saturates; frequencies add up
    used = 512 * g["instances"]
    assert not tr[:779, used:].any()
    assert (tr[780] == np.minimum(np.arange(tr.shape[1]), 65535)).all()
    assert int(tr[779].sum()) == 450 * tr.shape[1]
    assert [int(x) for x in tr[779, :8]] == g["frequency_first8"]


@pytest.mark.parametrize("kind", [I.KIND_G2, I.KIND_FQ])
def test_trace_golden(oracle, golden, kind):
    g = [t for t in golden["traces"] if t["kind"] == kind][0]
    inp, ts = I.make_inputs(kind, g["instances"], I.config_seed(g["config_id"]))
    tr, res = oracle.generate_trace(kind, inp, ts, want_results=True)
    assert hashlib.sha256(tr.tobytes()).hexdigest() == g["trace_sha256"]
    assert res.tolist() == g["result_limbs"]
    if kind == I.KIND_FQ:
        for k in range(g["instances"]):
            s, x = words_to_int(inp[k, 0:4]), words_to_int(inp[k, 4:8])
            assert limbs_to_int(res[k]) == pow(x, s, P)
            assert limbs_to_int(tr[16:32, 512 * k + 511]) == pow(x, s, P)
    else:
        for k in range(g["instances"]):
            assert (oracle.native_result(kind, inp[k]) == res[k]).all()


def test_infinity_is_rejected(oracle):
    """offset = -x: row 0 would add a point to its negative, unsupported by design (g1/add.rs:49-51)."""
    inp, ts = I.make_inputs(I.KIND_G1, 1, I.config_seed(98))
    y = words_to_int(inp[0, 8:12])
    inp[0, 12:16] = inp[0, 4:8]
    inp[0, 16:20] = [((P - y) >> (64 * i)) & I.MASK64 for i in range(4)]
    with pytest.raises(oracle.OracleError):
        oracle.generate_trace(I.KIND_G1, inp, ts)
