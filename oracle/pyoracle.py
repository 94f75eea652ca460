"""ctypes binding of the CPU oracle (oracle/build/libpb254_oracle.so).

ORACLE = test infrastructure. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "libpb254_oracle.so")
_lib = None

u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_last_error.restype = C.c_char_p
        _lib.orc_gl_mul.restype = C.c_uint64
        _lib.orc_gl_mul.argtypes = [C.c_uint64, C.c_uint64]
        _lib.orc_gl_inv.restype = C.c_uint64
        _lib.orc_gl_inv.argtypes = [C.c_uint64]
        _lib.orc_gl_root_of_unity.restype = C.c_uint64
        _lib.orc_gl_root_of_unity.argtypes = [C.c_uint]
        _lib.orc_trace_rows.restype = C.c_size_t
        _lib.orc_trace_rows.argtypes = [C.c_size_t, C.c_size_t]
        _lib.orc_proof_words.restype = C.c_size_t
        _lib.orc_proof_words.argtypes = [C.c_void_p]
        _lib.orc_proof_debug_words.restype = C.c_size_t
        _lib.orc_proof_debug_words.argtypes = [C.c_void_p, C.c_int]
        _lib.orc_proof_copy.argtypes = [C.c_void_p, C.c_void_p]
        _lib.orc_proof_debug_copy.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.orc_proof_free.argtypes = [C.c_void_p]
    return _lib


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


def _check(rc):
    if rc != 0:
        raise OracleError(rc, lib().orc_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int):
    lib().orc_set_num_threads(int(n))


def poseidon_round_constants():
    out = np.zeros(360, dtype=np.uint64)
    lib().orc_poseidon_round_constants(_p(out))
    return out


def poseidon_permute(state):
    s = _u64(state).copy()
    assert s.shape == (12,)
    lib().orc_poseidon_permute(_p(s))
    return s


def hash_no_pad(x):
    x = _u64(x)
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_hash_no_pad(_p(x), C.c_size_t(x.size), _p(out))
    return out


def two_to_one(l, r):
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_two_to_one(_p(_u64(l)), _p(_u64(r)), _p(out))
    return out


def fft(a):
    a = _u64(a).copy()
    lib().orc_fft(_p(a), C.c_size_t(a.size))
    return a


def ifft(a):
    a = _u64(a).copy()
    lib().orc_ifft(_p(a), C.c_size_t(a.size))
    return a


def lde_batch(values, rate_bits):
    """values: (cols, n) -> (coeffs (cols, n), lde (cols, n << rate_bits)) natural order."""
    values = _u64(values)
    cols, n = values.shape
    coeffs = np.zeros((cols, n), dtype=np.uint64)
    lde = np.zeros((cols, n << rate_bits), dtype=np.uint64)
    lib().orc_lde_batch(_p(values), C.c_size_t(cols), C.c_size_t(n), C.c_uint(rate_bits), _p(coeffs), _p(lde))
    return coeffs, lde


def commit(values, rate_bits, cap_height, from_coeffs=False, want_digests=False):
    """PolynomialBatch::from_values / from_coeffs -> cap (2^cap_height, 4) [, digest levels bottom-up]."""
    values = _u64(values)
    cols, n = values.shape
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    N = n << rate_bits
    dig = None
    if want_digests:
        total = 0
        m = N
        while m >= (1 << cap_height):
            total += m
            m //= 2
        dig = np.zeros((total, 4), dtype=np.uint64)
    _check(lib().orc_commit(_p(values), C.c_size_t(cols), C.c_size_t(n), C.c_uint(rate_bits), C.c_uint(cap_height),
                            C.c_int(1 if from_coeffs else 0), _p(cap), _p(dig) if dig is not None else None))
    return (cap, dig) if want_digests else cap


def commit_streamed(values, rate_bits, cap_height):
    """Cap of PolynomialBatch::from_values, computed in 8-column chunks (bounded host memory)."""
    values = _u64(values)
    cols, n = values.shape
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    _check(lib().orc_commit_streamed(_p(values), C.c_size_t(cols), C.c_size_t(n), C.c_uint(rate_bits),
                                     C.c_uint(cap_height), _p(cap)))
    return cap


def merkle_cap(leaves, cap_height):
    leaves = _u64(leaves)
    n, w = leaves.shape
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    _check(lib().orc_merkle(_p(leaves), C.c_size_t(n), C.c_size_t(w), C.c_uint(cap_height), _p(cap)))
    return cap


def width(kind):
    return lib().orc_width(kind)


def reg_len(kind):
    return lib().orc_reg_len(kind)


def num_aux(kind, num_challenges=2):
    return lib().orc_num_aux(kind, num_challenges)


def trace_rows(n_inputs, min_rows):
    return lib().orc_trace_rows(n_inputs, min_rows)


def gen_modulus_zero(input31):
    a = np.ascontiguousarray(input31, dtype=np.int64)
    assert a.shape == (31,)
    out = np.zeros(80, dtype=np.uint64)
    _check(lib().orc_gen_modulus_zero(_p(a), _p(out)))
    return out


def gen_is_modulus_zero(input16):
    a = np.ascontiguousarray(input16, dtype=np.int64)
    assert a.shape == (16,)
    out = np.zeros(96, dtype=np.uint64)
    z = C.c_int(0)
    _check(lib().orc_gen_is_modulus_zero(_p(a), _p(out), C.byref(z)))
    return z.value, out


def native_result(kind, input_row):
    row = _u64(input_row)
    out = np.zeros(reg_len(kind), dtype=np.uint64)
    _check(lib().orc_native_result(kind, _p(row), _p(out)))
    return out


def generate_trace(kind, inputs, timestamps, min_rows=1 << 16, want_results=False):
    inputs = _u64(inputs)
    timestamps = _u64(timestamps)
    k = inputs.shape[0]
    n = trace_rows(k, min_rows)
    cols = np.zeros((width(kind), n), dtype=np.uint64)
    res = np.zeros((k, reg_len(kind)), dtype=np.uint64) if want_results else None
    _check(lib().orc_generate_trace(kind, _p(inputs), _p(timestamps), C.c_size_t(k), C.c_size_t(min_rows), _p(cols),
                                    _p(res) if res is not None else None))
    return (cols, res) if want_results else cols


def _cfg(cfg):
    if cfg is None:
        return None
    return (C.c_uint * 7)(*[int(x) for x in cfg])


class Proof:
    """Owned handle to an oracle proof (+ optional debug artefacts)."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)

    def __del__(self):
        try:
            if self._h:
                lib().orc_proof_free(self._h)
                self._h = None
        except Exception:
            pass

    def words(self):
        n = lib().orc_proof_words(self._h)
        out = np.zeros(n, dtype=np.uint64)
        lib().orc_proof_copy(self._h, _p(out))
        return out

    def bytes(self):
        return self.words().tobytes()

    def debug(self, which):
        n = lib().orc_proof_debug_words(self._h, which)
        out = np.zeros(n, dtype=np.uint64)
        if n:
            lib().orc_proof_debug_copy(self._h, which, _p(out))
        return out


def prove(kind, trace_cols, cfg=None, keep_debug=False):
    trace_cols = _u64(trace_cols)
    assert trace_cols.shape[0] == width(kind)
    h = C.c_void_p()
    _check(lib().orc_prove(kind, _p(trace_cols), C.c_size_t(trace_cols.shape[1]), _cfg(cfg), C.c_int(int(keep_debug)),
                           C.byref(h)))
    return Proof(h.value)


def prove_inputs(kind, inputs, timestamps, min_rows=1 << 16, cfg=None, keep_debug=False):
    inputs = _u64(inputs)
    timestamps = _u64(timestamps)
    h = C.c_void_p()
    t_trace, t_prove = C.c_double(0), C.c_double(0)
    _check(lib().orc_prove_inputs(kind, _p(inputs), _p(timestamps), C.c_size_t(inputs.shape[0]), C.c_size_t(min_rows),
                                  _cfg(cfg), C.c_int(int(keep_debug)), C.byref(h), C.byref(t_trace), C.byref(t_prove)))
    return Proof(h.value), t_trace.value, t_prove.value


def verify(proof_words, inputs, timestamps):
    w = _u64(proof_words)
    inputs = _u64(inputs)
    timestamps = _u64(timestamps)
    _check(lib().orc_verify(_p(w), C.c_size_t(w.size), _p(inputs), _p(timestamps), C.c_size_t(inputs.shape[0])))
    return True


def eval_constraints_base(kind, local, nxt, aux_local, aux_next, betas, gammas, alphas, z_last, l_first, l_last):
    nch = len(betas)
    acc = np.zeros(nch, dtype=np.uint64)
    cnt = lib().orc_eval_constraints_base(kind, _p(_u64(local)), _p(_u64(nxt)), _p(_u64(aux_local)), _p(_u64(aux_next)),
                                          _p(_u64(betas)), _p(_u64(gammas)), _p(_u64(alphas)), C.c_uint(nch),
                                          C.c_uint64(int(z_last)), C.c_uint64(int(l_first)), C.c_uint64(int(l_last)),
                                          _p(acc))
    return acc, cnt
