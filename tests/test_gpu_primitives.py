"""GPU parity of the commitment building blocks (K3 lde_batch, K4 leaf_hash, K5 merkle_levels and the
Poseidon permutation) against the CPU oracle, through the C ABI. Bit-exact (integer arithmetic)."""
import numpy as np
import pytest

from util import rand_field

pytestmark = pytest.mark.gpu


def test_poseidon_permute_matches_oracle_and_kats(gpu_ctx, oracle):
    rng = np.random.default_rng(7)
    st = rand_field(rng, (4096, 12))
    st[0] = 0
    st[1] = np.arange(12)
    st[2] = 0xFFFFFFFF00000000  # p - 1 everywhere
    out = gpu_ctx.poseidon_permute(st)
    assert out[0][0] == 0x3C18A9786CB0B359 and out[0][11] == 0x1792B1C4342109D7
    assert out[1][0] == 0xD64E1E3EFC5B8E9E and out[1][11] == 0x5C0A27FCB0E1459B
    for i in list(range(64)) + [4095]:
        assert (out[i] == oracle.poseidon_permute(st[i])).all(), i


@pytest.mark.parametrize("cols,log_n,rate_bits", [(3, 3, 1), (5, 8, 1), (4, 10, 1), (7, 12, 1), (2, 14, 2),
                                                   (9, 16, 1), (3, 12, 3), (2, 19, 1), (1, 21, 1), (1, 22, 1), (1, 21, 2)])
def test_lde_matches_oracle(gpu_ctx, oracle, cols, log_n, rate_bits):
    rng = np.random.default_rng(log_n * 10 + rate_bits)
    v = rand_field(rng, (cols, 1 << log_n))
    coeffs, lde = oracle.lde_batch(v, rate_bits)
    assert (gpu_ctx.lde_batch(v, rate_bits) == lde).all()
    assert (gpu_ctx.lde_batch(coeffs, rate_bits, from_coeffs=True) == lde).all()


@pytest.mark.parametrize("cols,log_n,rate_bits,cap", [(1, 4, 1, 4), (4, 6, 1, 4), (8, 8, 1, 4), (9, 10, 1, 4),
                                                       (17, 12, 1, 4), (130, 12, 1, 2), (3, 10, 3, 0)])
def test_commit_matches_oracle(gpu_ctx, oracle, cols, log_n, rate_bits, cap):
    rng = np.random.default_rng(cols * 100 + log_n)
    v = rand_field(rng, (cols, 1 << log_n))
    cap_o, dig_o = oracle.commit(v, rate_bits, cap, want_digests=True)
    cap_g, dig_g = gpu_ctx.commit(v, rate_bits, cap, want_digests=True)
    assert (dig_g == dig_o).all()
    assert (cap_g == cap_o).all()


def test_lde_linearity_large(gpu_ctx):
    """size-independent property at a production size: LDE(a + b) == LDE(a) + LDE(b) (mod p)."""
    rng = np.random.default_rng(3)
    n = 1 << 20
    a = rand_field(rng, (1, n))
    b = rand_field(rng, (1, n))
    P = np.uint64(0xFFFFFFFF00000001)

    def addmod(x, y):
        s = x + y
        wrapped = s < x
        s = np.where(wrapped, s + np.uint64(0xFFFFFFFF), s)
        return np.where(s >= P, s - P, s)

    la, lb, lab = (gpu_ctx.lde_batch(x, 1) for x in (a, b, addmod(a, b)))
    assert (addmod(la, lb) == lab).all()
    # the LDE on the even coset points restricted ... spot check: value 0 equals P(7) by Horner over few coeffs


@pytest.mark.parametrize("cols,rows,stride", [(21, 1000, 1024), (8, 129, 129), (5, 1, 7), (4, 300, 300)])
def test_leaf_hash_of_ragged_row_blocks(gpu_ctx, oracle, cols, rows, stride):
    """pb254_leaf_hash_rows_dev (the row-block form the sharded prover uses; on the device the tensor-core kernel, one
    CTA per 128 rows): row counts that are not a multiple of 128, a stride larger than the row count, a width that is
    not a multiple of the sponge rate, and the <= 4 column no-hash case - each digest against the oracle's hash."""
    import torch
    rng = np.random.default_rng(cols * 1000 + rows)
    m = rand_field(rng, (cols, stride))
    d_m = torch.from_numpy(m.view(np.int64)).to("cuda:0")   # the device of the gpu_ctx fixture
    d_out = torch.zeros((rows, 4), dtype=torch.int64, device="cuda:0")
    torch.cuda.synchronize()  # the context has its own (non-blocking) stream
    gpu_ctx.leaf_hash_rows_dev(d_m.data_ptr(), stride, cols, rows, d_out.data_ptr())
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(np.uint64)
    for i in sorted(set([0, 1, 127, 128, rows - 1, rows // 2]) & set(range(rows))):
        row = m[:, i]
        want = oracle.hash_no_pad(row) if cols > 4 else np.concatenate([row, np.zeros(4 - cols, dtype=np.uint64)])
        assert (got[i] == want).all(), i
