"""the batch producers of hot-path work items (plonky2_bn254_b200/workloads.py) -
is_square Legendre batches (src/fields/fq.rs:283-295) and hash_to_g2 cofactor-clearing batches
(src/utils/hash_to_g2.rs:76-148, 195-208). CPU part: the host algorithms against independent big-integer checks and
the oracle's Poseidon; GPU part: the batches through the fq_exp / G2 STARK provers, native results against Python."""
import numpy as np
import pytest

from plonky2_bn254_b200 import inputs as I, workloads as Wk
from util import rand_field

P = I.BN254_P


def oracle_permute(oracle):
    return lambda st: np.stack([oracle.poseidon_permute(r) for r in np.asarray(st, dtype=np.uint64).reshape(-1, 12)])


def scalar_challenger(oracle, msg, count):
    """A literal, element-at-a-time restatement of plonky2's Challenger (observe_element / get_challenge)."""
    state, inp, out, res = np.zeros(12, dtype=np.uint64), [], [], []

    def duplex():
        nonlocal state, inp, out
        for i, v in enumerate(inp):
            state[i] = v
        inp = []
        state = oracle.poseidon_permute(state)
        out = list(state[:8])
    for e in msg:
        out = []
        inp.append(e)
        if len(inp) == 8:
            duplex()
    for _ in range(count):
        if inp or not out:
            duplex()
        res.append(int(out.pop()))
    return res


@pytest.mark.parametrize("ln", [0, 3, 8, 13, 16])
def test_hash_to_fq2_matches_scalar_challenger(oracle, ln):
    rng = np.random.default_rng(ln + 1)
    msgs = rand_field(rng, (3, ln))
    got = Wk.hash_to_fq2(msgs, oracle_permute(oracle))
    for m, (c0, c1) in zip(msgs, got):
        ch = scalar_challenger(oracle, list(m), 32)
        want = [sum((c & 0xFFFFFFFF) << (32 * i) for i, c in enumerate(ch[16 * h:16 * h + 16])) % P for h in range(2)]
        assert (c0, c1) == (want[0], want[1])


def test_map_to_curve_properties():
    rng = I.SplitMix64(99)
    for _ in range(12):
        u = (rng.bits256() % P, rng.bits256() % P)
        x, y = Wk.map_to_curve(u)
        assert I._Fq2.mul(y, y) == Wk._g(x)                      # on the curve (hash_to_g2.rs:146)
        assert Wk.fq2_sgn(y) == Wk.fq2_sgn(u)                    # sign rule (:143-145)
        cleared = Wk.g2_mul(Wk.G2_COFACTOR, (x, y))
        assert cleared is not None and Wk.g2_mul(I.BN254_R, cleared) is None   # cofactor clearing lands in the subgroup
    # SvdW constants (:114-118): tv4^2 = -3 g(Z) Z^2, tv6 = -4 g(Z) / (3 Z^2)
    assert I._Fq2.mul(Wk._TV4, Wk._TV4) == I._Fq2.mul(Wk._neg(Wk._GZ), (3, 0))
    assert I._Fq2.mul(Wk._TV6, (3, 0)) == I._Fq2.mul(Wk._neg((4, 0)), Wk._GZ)


def test_is_square_inputs_and_outputs():
    xs = [0, 1, 4, 3, P - 1, 5]
    inp, ts = Wk.is_square_inputs(xs)
    assert inp.shape == (6, 8) and (ts == np.arange(6)).all()
    assert Wk._int256(inp[2, :4]) == (P - 1) // 2 and Wk._int256(inp[3, 4:]) == 3
    sym = [pow(x, (P - 1) // 2, P) for x in xs]
    res = np.array([[(s >> (16 * i)) & 0xFFFF for i in range(16)] for s in sym], dtype=np.uint64)
    assert Wk.is_square_outputs(res).tolist() == [s == 1 for s in sym]


def test_hash_to_g2_inputs_roundtrip_python(oracle):
    """The offset bookkeeping: (cofactor * point + offset) - offset == cofactor * point, computed in Python."""
    rng = np.random.default_rng(5)
    msgs = rand_field(rng, (2, 8))
    inp, ts, offs = Wk.hash_to_g2_inputs(msgs, oracle_permute(oracle), seed=7)
    assert inp.shape == (2, 36) and Wk._int256(inp[0, :4]) == Wk.G2_COFACTOR
    res = []
    for row, off in zip(inp, offs):
        pt = Wk._g2_from_words(row[4:20])
        assert Wk._g2_from_words(row[20:36]) == off
        res.append(Wk._g2_limbs16(Wk.g2_add(Wk.g2_mul(Wk.G2_COFACTOR, pt), off)))
        # the limb layout of Proof.results() is the oracle's native_result layout
        assert (oracle.native_result(I.KIND_G2, row) == np.array(res[-1], dtype=np.uint64)).all()
    outs = Wk.hash_to_g2_outputs(np.array(res, dtype=np.uint64), offs)
    for row, o in zip(inp, outs):
        assert o == Wk.g2_mul(Wk.G2_COFACTOR, Wk._g2_from_words(row[4:20]))


@pytest.mark.gpu
def test_is_square_batch_through_fq_exp_stark(gpu_ctx, oracle):
    rng = I.SplitMix64(123)
    xs = [0, 1, 2, 3, P - 1] + [rng.bits256() % P for _ in range(27)]
    inp, ts = Wk.is_square_inputs(xs)
    pf = gpu_ctx.prove(I.KIND_FQ, inp, ts)
    got = Wk.is_square_outputs(pf.results())
    assert got.tolist() == [pow(x, (P - 1) // 2, P) == 1 for x in xs]
    assert got.tolist() == [x != 0 and I.fq_sqrt(x) is not None for x in xs]
    assert oracle.verify(pf.words(), inp, ts)


@pytest.mark.gpu
def test_hash_to_g2_batch_through_g2_stark(gpu_ctx, oracle):
    rng = np.random.default_rng(11)
    msgs = rand_field(rng, (3, 8))
    gpu_perm = lambda st: gpu_ctx.poseidon_permute(st)
    inp, ts, offs = Wk.hash_to_g2_inputs(msgs, gpu_perm, seed=21)
    inp2, _, _ = Wk.hash_to_g2_inputs(msgs, oracle_permute(oracle), seed=21)
    assert (inp == inp2).all()                                       # GPU Poseidon == oracle Poseidon in the challenger
    pf = gpu_ctx.prove(I.KIND_G2, inp, ts)
    outs = Wk.hash_to_g2_outputs(pf.results(), offs)
    for row, o in zip(inp, outs):
        want = Wk.g2_mul(Wk.G2_COFACTOR, Wk._g2_from_words(row[4:20]))
        assert o == want and Wk.g2_mul(I.BN254_R, o) is None
    assert oracle.verify(pf.words(), inp, ts)


def test_is_square_batch_hostsim(hostsim_ctx):
    """Same plumbing as the GPU test, on the host-simulation build (results come from the trace the STARK proves)."""
    xs = [0, 1, 2, 3, P - 1, 7]
    inp, ts = Wk.is_square_inputs(xs)
    pf = hostsim_ctx.prove(I.KIND_FQ, inp, ts)
    assert Wk.is_square_outputs(pf.results()).tolist() == [pow(x, (P - 1) // 2, P) == 1 for x in xs]
