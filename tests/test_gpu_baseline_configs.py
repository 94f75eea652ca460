"""GPU parity at the sizes BASELINE.json names, through the C ABI, against the CPU oracle:

  config 2  G1 x 1024, 2^19 rows, standard_fast_config   proof BYTES == oracle proof bytes
  config 3  G2 x 1024, 2^19 rows x 1295 columns          trace cells, proof BYTES == oracle proof bytes
  config 4  fq_exp x 4096, 2^21 rows, blow-up 8, 28 queries
            trace cells == oracle trace; trace cap == oracle cap (the oracle commits the 427 x 2^24 LDE in 8-column
            chunks, oracle.commit_streamed, so that host memory stays bounded); the oracle verifier accepts the proof;
            a full proof byte comparison with rate_bits = 3 at 2^16 and 2^18 rows
  edge cases on the device for G2 (offset == x, x.c0 equal only, x.c1 equal only: is_x_eq = is_c0_zero * is_c1_zero,
            src/starks/curves/g2/add.rs:59-130) and fq_exp (x in {0, 1, p-1}, s in {0, 2^256-1},
            src/starks/fields/exp_stark.rs:107-137): trace cells and proof bytes

These are the slow tests of the GPU tier (the oracle needs about a minute per 2^19-row proof on the GPU box's host
cores); everything is integer arithmetic, the bar is bit-exact."""
import numpy as np
import pytest

from plonky2_bn254_b200 import inputs as I

pytestmark = pytest.mark.gpu


def _same(a, b, what):
    assert a.shape == b.shape, what
    bad = np.flatnonzero(a.reshape(-1) != b.reshape(-1))
    assert bad.size == 0, (what, bad[:8].tolist())


def _use_all_cores(oracle):
    import os
    oracle.set_num_threads(os.cpu_count() or 1)


def test_config2_g1_1024_proof_bytes(gpu_ctx, oracle):
    """BASELINE configs[1]: 'batched G1 scalar-mul STARK, 2^10 scalar-muls in one trace, 1 B200, bit-exact vs CPU'."""
    _use_all_cores(oracle)
    inp, ts = I.make_inputs(I.KIND_G1, 1024, I.config_seed(2))
    pf = gpu_ctx.prove(I.KIND_G1, inp, ts)
    w = pf.words()
    assert int(w[2]) == 19
    ref, _, _ = oracle.prove_inputs(I.KIND_G1, inp, ts)
    _same(w, ref.words(), "config 2 proof bytes")
    res = pf.results().reshape(1024, -1)
    for i in (0, 511, 1023):
        assert (res[i] == oracle.native_result(I.KIND_G1, inp[i])).all()


def test_config3_g2_1024_trace_and_proof_bytes(gpu_ctx, oracle):
    """BASELINE configs[2]: G2 x 2^10 (Fq2 limbs), one proof per GPU."""
    _use_all_cores(oracle)
    inp, ts = I.make_inputs(I.KIND_G2, 1024, I.config_seed(3))
    tr = gpu_ctx.generate_trace(I.KIND_G2, inp, ts)
    ref_tr = oracle.generate_trace(I.KIND_G2, inp, ts)
    assert tr.shape == (1295, 1 << 19)
    _same(tr, ref_tr, "config 3 trace")
    del tr
    w = gpu_ctx.prove(I.KIND_G2, inp, ts).words()
    ref = oracle.prove(I.KIND_G2, ref_tr)
    _same(w, ref.words(), "config 3 proof bytes")


def test_config4_fq_4096_blowup8(gpu_ctx, oracle):
    """BASELINE configs[3]: fq_exp x 2^12, LDE blow-up 8 (rate_bits = 3, 28 query rounds)."""
    _use_all_cores(oracle)
    cfg = gpu_ctx.L.standard_fast_config()
    cfg.rate_bits, cfg.num_query_rounds = 3, 28
    inp, ts = I.make_inputs(I.KIND_FQ, 4096, I.config_seed(4))
    tr = gpu_ctx.generate_trace(I.KIND_FQ, inp, ts)
    ref_tr = oracle.generate_trace(I.KIND_FQ, inp, ts)
    assert tr.shape == (427, 1 << 21)
    _same(tr, ref_tr, "config 4 trace")
    del tr
    w = gpu_ctx.prove(I.KIND_FQ, inp, ts, config=cfg).words()
    assert int(w[2]) == 21 and int(w[3]) == 3
    cap = oracle.commit_streamed(ref_tr, 3, 4)
    _same(w[22:22 + 64], cap.reshape(-1), "config 4 trace cap")
    assert oracle.verify(w, inp, ts)
    assert gpu_ctx.L.verify(I.KIND_FQ, w, inp, ts, config=cfg)


@pytest.mark.parametrize("k", [3, 300])
def test_rate_bits_3_proof_bytes(gpu_ctx, oracle, k):
    """rate_bits = 3 against the oracle, byte for byte (2^16 and 2^18 rows)."""
    _use_all_cores(oracle)
    cfg = gpu_ctx.L.standard_fast_config()
    cfg.rate_bits, cfg.num_query_rounds = 3, 28
    inp, ts = I.make_inputs(I.KIND_FQ, k, I.config_seed(40 + k))
    pf = gpu_ctx.prove(I.KIND_FQ, inp, ts, config=cfg, keep_debug=True)
    ref, _, _ = oracle.prove_inputs(I.KIND_FQ, inp, ts, cfg=cfg.as_tuple(), keep_debug=True)
    for which, name in ((2, "challenges"), (0, "auxiliary columns"), (1, "quotient chunks"), (3, "query indices")):
        _same(pf.debug(which), ref.debug(which), name)
    _same(pf.words(), ref.words(), "rate_bits = 3 proof bytes")


def _int(words):
    return sum(int(w) << (64 * i) for i, w in enumerate(words))


def _g2_offset_sharing(inp_row, share_c0: bool):
    """An offset ON the G2 curve whose x shares exactly one Fq2 component with the row's x: the device chains
    re-associate additions, so the inputs must be curve points (as G2Affine is by construction)."""
    xc0, xc1 = _int(inp_row[4:8]), _int(inp_row[8:12])
    rng = I.SplitMix64(0xED6E + share_c0)
    while True:
        other = rng.bits256() % I.BN254_P
        x = (xc0, other) if share_c0 else (other, xc1)
        pt = I.g2_point_with_x(x)
        if pt is not None and other not in (xc0, xc1):
            (a, b), (c, d) = pt
            return I._words(a) + I._words(b) + I._words(c) + I._words(d)


def _g2_edge_inputs():
    inp, ts = I.make_inputs(I.KIND_G2, 6, I.config_seed(199))
    inp[0, 0:4] = 0                                   # s = 0
    inp[1, 0:4] = np.uint64(0xFFFFFFFFFFFFFFFF)       # s = 2^256 - 1
    inp[2, 20:36] = inp[2, 4:20]                      # offset == x: the doubling branch in an adding row
    inp[3, 20:36] = _g2_offset_sharing(inp[3], True)  # only x.c0 equal: is_c0_zero = 1, is_c1_zero = 0
    inp[4, 20:36] = _g2_offset_sharing(inp[4], False) # only x.c1 equal: is_c0_zero = 0, is_c1_zero = 1
    inp[5, 0:4] = [1, 0, 0, 0]                        # s = 1
    return inp, ts


def _fq_edge_inputs():
    inp, ts = I.make_inputs(I.KIND_FQ, 6, I.config_seed(198))
    inp[0, 4:8] = 0                                   # x = 0
    inp[1, 4:8] = [1, 0, 0, 0]                        # x = 1
    inp[2, 0:4] = 0                                   # s = 0
    inp[3, 0:4] = np.uint64(0xFFFFFFFFFFFFFFFF)       # s = 2^256 - 1
    inp[4, 0:4] = 0
    inp[4, 4:8] = 0                                   # 0^0 = 1
    inp[5, 4:8] = I._words(I.BN254_P - 1)             # x = -1
    return inp, ts


@pytest.mark.parametrize("kind,make", [(I.KIND_G2, _g2_edge_inputs), (I.KIND_FQ, _fq_edge_inputs)])
def test_device_edge_cases_g2_fq(gpu_ctx, oracle, kind, make):
    _use_all_cores(oracle)
    inp, ts = make()
    ref_tr = oracle.generate_trace(kind, inp, ts)
    _same(gpu_ctx.generate_trace(kind, inp, ts), ref_tr, "edge-case trace")
    pf = gpu_ctx.prove(kind, inp, ts)
    ref = oracle.prove(kind, ref_tr)
    _same(pf.words(), ref.words(), "edge-case proof bytes")
    res = pf.results().reshape(inp.shape[0], -1)
    for i in range(inp.shape[0]):
        assert (res[i] == oracle.native_result(kind, inp[i])).all()
    assert gpu_ctx.L.verify(kind, pf.words(), inp, ts)


@pytest.mark.parametrize("kind", [I.KIND_G1, I.KIND_G2])
def test_off_curve_point_is_an_error(gpu_ctx, kind):
    """G1Affine / G2Affine are curve points by construction in the reference; a coordinate pair that is not on the
    curve is refused (PB254_E_NOT_ON_CURVE) instead of producing a trace."""
    from plonky2_bn254_b200 import ffi
    inp, ts = I.make_inputs(kind, 2, I.config_seed(197))
    inp[1, 4] ^= np.uint64(1)
    with pytest.raises(ffi.Pb254Error) as e:
        gpu_ctx.generate_trace(kind, inp, ts)
    assert e.value.code == 8
