"""Deterministic synthetic inputs for the three STARKs.

The reference's tests draw from ``rand::thread_rng()`` (src/starks/curves/g1/scalar_mul_stark.rs:553-566,
src/starks/common/utils.rs:21-25) and are therefore not reproducible; this module replaces that with a
SplitMix64 stream seeded by ``0x706232353400 + config_id``:

* scalar ``s``  = 32 PRNG bytes, little endian (uniform in [0, 2^256), like ``random_biguint``)
* points        = ``k * G`` with ``k`` a uniform 254-bit integer, ``G`` the G1 generator (1, 2) or the
                  standard G2 generator (always in the prime-order subgroup, like arkworks' ``rand``)
* ``fq_exp`` base = 256 PRNG bits mod p

Wire format (little-endian 4 x u64 per 256-bit value), one row per instance:
  G1: s, x.x, x.y, offset.x, offset.y                      (20 words)
  G2: s, x.x.c0, x.x.c1, x.y.c0, x.y.c1, offset...         (36 words)
  Fq: s, x                                                 (8 words)

Pure Python big-int arithmetic; this is host-side workload generation, not part of the prover.
"""
from __future__ import annotations

import numpy as np

BN254_P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
BN254_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
G1_GEN = (1, 2)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)
KIND_G1, KIND_G2, KIND_FQ = 0, 1, 2
KIND_NAMES = {KIND_G1: "g1", KIND_G2: "g2", KIND_FQ: "fq"}
IN_WORDS = {KIND_G1: 20, KIND_G2: 36, KIND_FQ: 8}
SEED_BASE = 0x706232353400
MASK64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & MASK64

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def bits256(self) -> int:
        return self.next() | (self.next() << 64) | (self.next() << 128) | (self.next() << 192)


# ---- field helpers: Fq as int, Fq2 as (c0, c1) with u^2 = -1 ---------------------------------
class _Fq:
    zero, one = 0, 1

    @staticmethod
    def add(a, b): return (a + b) % BN254_P
    @staticmethod
    def sub(a, b): return (a - b) % BN254_P
    @staticmethod
    def mul(a, b): return (a * b) % BN254_P
    @staticmethod
    def inv(a): return pow(a, -1, BN254_P)
    @staticmethod
    def is_zero(a): return a == 0
    @staticmethod
    def small(k): return k % BN254_P


class _Fq2:
    zero, one = (0, 0), (1, 0)

    @staticmethod
    def add(a, b): return ((a[0] + b[0]) % BN254_P, (a[1] + b[1]) % BN254_P)
    @staticmethod
    def sub(a, b): return ((a[0] - b[0]) % BN254_P, (a[1] - b[1]) % BN254_P)
    @staticmethod
    def mul(a, b):
        return ((a[0] * b[0] - a[1] * b[1]) % BN254_P, (a[0] * b[1] + a[1] * b[0]) % BN254_P)
    @staticmethod
    def inv(a):
        n = pow(a[0] * a[0] + a[1] * a[1], -1, BN254_P)
        return (a[0] * n % BN254_P, -a[1] * n % BN254_P)
    @staticmethod
    def is_zero(a): return a[0] == 0 and a[1] == 0
    @staticmethod
    def small(k): return (k % BN254_P, 0)


def _jac_double(F, P):
    X, Y, Z = P
    if F.is_zero(Z):
        return P
    A = F.mul(X, X); B = F.mul(Y, Y); C = F.mul(B, B)
    t = F.add(X, B)
    D = F.mul(F.small(2), F.sub(F.sub(F.mul(t, t), A), C))
    E = F.mul(F.small(3), A)
    Fv = F.mul(E, E)
    X3 = F.sub(Fv, F.mul(F.small(2), D))
    Y3 = F.sub(F.mul(E, F.sub(D, X3)), F.mul(F.small(8), C))
    Z3 = F.mul(F.small(2), F.mul(Y, Z))
    return (X3, Y3, Z3)


def _jac_add_affine(F, P, Q):
    """P Jacobian + Q affine (mixed addition), handling identity and doubling."""
    X1, Y1, Z1 = P
    if F.is_zero(Z1):
        return (Q[0], Q[1], F.one)
    Z1Z1 = F.mul(Z1, Z1)
    U2 = F.mul(Q[0], Z1Z1)
    S2 = F.mul(Q[1], F.mul(Z1, Z1Z1))
    H = F.sub(U2, X1)
    r = F.sub(S2, Y1)
    if F.is_zero(H):
        if F.is_zero(r):
            return _jac_double(F, P)
        return (F.one, F.one, F.zero)
    HH = F.mul(H, H); HHH = F.mul(H, HH); V = F.mul(X1, HH)
    X3 = F.sub(F.sub(F.mul(r, r), HHH), F.mul(F.small(2), V))
    Y3 = F.sub(F.mul(r, F.sub(V, X3)), F.mul(Y1, HHH))
    Z3 = F.mul(Z1, H)
    return (X3, Y3, Z3)


def _to_affine(F, P):
    X, Y, Z = P
    zi = F.inv(Z); zi2 = F.mul(zi, zi)
    return (F.mul(X, zi2), F.mul(Y, F.mul(zi, zi2)))


_TABLES: dict = {}


def _fixed_base_table(F, gen, key):
    """table[i][j] = (j * 2^(8 i)) * gen in affine, j = 1..255, i = 0..31."""
    if key in _TABLES:
        return _TABLES[key]
    table = []
    base = gen
    for _ in range(32):
        row = [None, base]
        acc = (base[0], base[1], F.one)
        jac = [acc]
        for _j in range(2, 256):
            acc = _jac_add_affine(F, acc, base)
            jac.append(acc)
        for pj in jac[1:]:
            row.append(_to_affine(F, pj))
        table.append(row)
        # base <- 256 * base
        nb = (base[0], base[1], F.one)
        for _ in range(8):
            nb = _jac_double(F, nb)
        base = _to_affine(F, nb)
    _TABLES[key] = table
    return table


def scalar_mul_gen(kind: int, k: int):
    """k * G (affine) for the G1 or G2 generator; k must not be a multiple of the group order."""
    F, gen = (_Fq, G1_GEN) if kind == KIND_G1 else (_Fq2, G2_GEN)
    table = _fixed_base_table(F, gen, kind)
    acc = (F.one, F.one, F.zero)
    for i in range(32):
        d = (k >> (8 * i)) & 0xFF
        if d:
            acc = _jac_add_affine(F, acc, table[i][d])
    return _to_affine(F, acc)


def fq_sqrt(a: int):
    """Square root in Fq (p = 3 mod 4), or None."""
    r = pow(a, (BN254_P + 1) // 4, BN254_P)
    return r if r * r % BN254_P == a % BN254_P else None


def fq2_sqrt(a):
    """Square root in Fq2 = Fq[u]/(u^2 + 1), or None."""
    a0, a1 = a
    if a1 == 0:
        r = fq_sqrt(a0)
        if r is not None:
            return (r, 0)
        r = fq_sqrt(-a0 % BN254_P)
        return None if r is None else (0, r)
    n = fq_sqrt((a0 * a0 + a1 * a1) % BN254_P)
    if n is None:
        return None
    half = pow(2, -1, BN254_P)
    for cand in ((a0 + n) * half % BN254_P, (a0 - n) * half % BN254_P):
        x0 = fq_sqrt(cand)
        if x0 is not None and x0 != 0:
            x1 = a1 * pow(2 * x0, -1, BN254_P) % BN254_P
            if _Fq2.mul((x0, x1), (x0, x1)) == (a0 % BN254_P, a1 % BN254_P):
                return (x0, x1)
    return None


G2_B = _Fq2.mul((3, 0), _Fq2.inv((9, 1)))  # b' = 3 / (9 + u)


def g2_point_with_x(x):
    """A point (x, y) of the G2 curve y^2 = x^3 + 3/(9+u) with the given x (any point of the curve, not necessarily
    of the prime-order subgroup), or None if x^3 + b' is not a square."""
    y = fq2_sqrt(_Fq2.add(_Fq2.mul(_Fq2.mul(x, x), x), G2_B))
    return None if y is None else (x, y)


def g1_point_with_x(x):
    y = fq_sqrt((x * x * x + 3) % BN254_P)
    return None if y is None else (x, y)


def _words(x: int):
    return [(x >> (64 * i)) & MASK64 for i in range(4)]


def make_inputs(kind: int, n_inputs: int, seed: int):
    """Returns (inputs[n, IN_WORDS] uint64, timestamps[n] uint64)."""
    rng = SplitMix64(seed)
    rows = []
    for _ in range(n_inputs):
        s = rng.bits256()
        row = _words(s)
        if kind == KIND_FQ:
            row += _words(rng.bits256() % BN254_P)
        else:
            for _pt in range(2):
                k = 0
                while k % BN254_R == 0:
                    k = rng.bits256() & ((1 << 254) - 1)
                pt = scalar_mul_gen(kind, k)
                if kind == KIND_G1:
                    row += _words(pt[0]) + _words(pt[1])
                else:
                    row += _words(pt[0][0]) + _words(pt[0][1]) + _words(pt[1][0]) + _words(pt[1][1])
        rows.append(row)
    inputs = np.array(rows, dtype=np.uint64).reshape(n_inputs, IN_WORDS[kind])
    timestamps = np.arange(n_inputs, dtype=np.uint64)
    return inputs, timestamps


def config_seed(config_id: int) -> int:
    return SEED_BASE + config_id
