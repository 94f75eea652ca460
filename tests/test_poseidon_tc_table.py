"""The operand table of the tensor-core Poseidon (csrc/poseidon_constants_tcb.inc, read by csrc/poseidon_tc.cuh) checked
without a GPU: the 96 x 128 byte tile is un-swizzled and the kernel's data flow - state bytes and a one-hot of the
layer as the A row, u8 x u8 -> s32 products, the byte-weight fold - is replayed with numpy integers for whole
permutations, which must equal the oracle's Poseidon (plonky2 0.2.2 hash/poseidon.rs, pinned to upstream KATs in
tests/test_oracle_primitives.py). The GPU tier runs the real kernel against the same oracle."""
import os
import re

import numpy as np

from util import rand_field

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 0xFFFFFFFF00000001


def load_b_tile():
    txt = open(os.path.join(ROOT, "plonky2_bn254_b200", "csrc", "poseidon_constants_tcb.inc")).read()
    words = [int(w, 16) for w in re.findall(r"0x([0-9a-fA-F]{8})u", txt)]
    assert len(words) == 96 * 128 // 4
    img = np.array(words, dtype="<u4").view(np.uint8)
    tile = np.zeros((96, 128), dtype=np.int64)
    for n in range(96):                       # 16-byte chunk c of row n is stored at chunk position c ^ (n & 7)
        for c in range(8):
            pos = n * 128 + ((c ^ (n & 7)) << 4)
            tile[n, 16 * c:16 * c + 16] = img[pos:pos + 16]
    return tile


def test_b_tile_is_the_mds_matrix_and_the_round_constants(oracle):
    B = load_b_tile()
    circ = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
    rc = oracle.poseidon_round_constants()
    for r in range(12):
        for q in range(8):
            row = B[8 * r + q]
            for i in range(12):
                for qq in range(8):
                    want = (circ[(i - r) % 12] + (8 if r == 0 and i == 0 else 0)) if qq == q else 0
                    assert row[8 * i + qq] == want
            for L in range(29):
                assert row[96 + L] == (int(rc[(L + 1) * 12 + r]) >> (8 * q)) & 0xFF
            assert (row[96 + 29:] == 0).all()


def test_replayed_tensor_core_permutation_equals_the_oracle(oracle):
    B = load_b_tile()
    rc = [int(x) for x in oracle.poseidon_round_constants()]
    rng = np.random.default_rng(11)
    states = [list(map(int, s)) for s in rand_field(rng, (6, 12))]
    states[0] = [0] * 12
    states[1] = [P - 1] * 12
    for st in states:
        s = [(x + rc[i]) % P for i, x in enumerate(st)]          # round-0 constants are added before the first S-box
        for L in range(30):
            if L < 4 or L >= 26:
                s = [pow(x, 7, P) for x in s]
            else:
                s[0] = pow(s[0], 7, P)
            a = np.zeros(128, dtype=np.int64)                     # the A row of this state
            for i in range(12):
                for q in range(8):
                    a[8 * i + q] = (s[i] >> (8 * q)) & 0xFF
            if L < 29:
                a[96 + L] = 1
            acc = B @ a                                           # the 96 s32 accumulators of the state's TMEM lane
            assert acc.max() < 1 << 18
            s = [sum(int(acc[8 * r + q]) << (8 * q) for q in range(8)) % P for r in range(12)]   # the fold
        assert s == [int(x) for x in oracle.poseidon_permute(np.array(st, dtype=np.uint64))]
