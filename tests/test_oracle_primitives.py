"""Pins the CPU oracle's field / hash / NTT / Merkle layer (SURVEY.md 8c anchors): the upstream Poseidon
known-answer vectors, the Goldilocks constants, and structural properties of the NTT and commitment."""
import numpy as np

from util import GL_P, rand_field

KAT_ZERO = [0x3C18A9786CB0B359, 0xC4055E3364A246C3, 0x7953DB0AB48808F4, 0xC71603F33A1144CA, 0xD7709673896996DC,
            0x46A84E87642F44ED, 0xD032648251EE0B3C, 0x1C687363B207DF62, 0xDF8565563E8045FE, 0x40F5B37FF4254DAE,
            0xD070F637B431067C, 0x1792B1C4342109D7]
KAT_RANGE = [0xD64E1E3EFC5B8E9E, 0x53666633020AAA47, 0xD40285597C6A8825, 0x613A4F81E81231D2, 0x414754BFEBD051F0,
             0xCB1F8980294A023F, 0x6EB2A9E4D54A9D0F, 0x1902BC3AF467E056, 0xF045D5EAFDC6021F, 0xE4150F77CAAA3BE5,
             0xC9BFD01D39B50CCE, 0x5C0A27FCB0E1459B]


def test_poseidon_upstream_known_answers(oracle, golden):
    """plonky2 0.2.2 hash/poseidon_goldilocks.rs test_vectors (SURVEY.md C.2 KAT 1 / KAT 2)."""
    z = oracle.poseidon_permute(np.zeros(12, dtype=np.uint64))
    r = oracle.poseidon_permute(np.arange(12, dtype=np.uint64))
    assert [int(x) for x in z] == KAT_ZERO == golden["poseidon_kat"]["zeros"]
    assert [int(x) for x in r] == KAT_RANGE == golden["poseidon_kat"]["range12"]


def test_poseidon_round_constants_regenerate(oracle):
    """ChaCha8Rng::seed_from_u64(0) + gen_range(0..p): first four and last two constants."""
    rc = oracle.poseidon_round_constants()
    assert [int(x) for x in rc[:4]] == [0xB585F766F2144405, 0x7746A55F43921AD7, 0xB2FB0D31CEE799B4,
                                        0x0F6760A4803427D7]
    assert [int(x) for x in rc[-2:]] == [0xDFD1C4FEBCC81238, 0xBC8DFB627FE558FC]
    assert (rc < np.uint64(GL_P)).all()


def test_goldilocks_constants(oracle):
    assert pow(7, (GL_P - 1) >> 32, GL_P) == 1753635133440165772 == oracle.lib().orc_gl_root_of_unity(32)
    assert pow(7, (GL_P - 1) // 2, GL_P) == GL_P - 1
    for k in (1, 5, 16, 20):
        w = oracle.lib().orc_gl_root_of_unity(k)
        assert pow(w, 1 << k, GL_P) == 1 and pow(w, 1 << (k - 1), GL_P) == GL_P - 1
    rng = np.random.default_rng(0)
    for a, b in rand_field(rng, (50, 2)).tolist():
        assert oracle.lib().orc_gl_mul(a, b) == a * b % GL_P
        if a:
            assert oracle.lib().orc_gl_inv(a) * a % GL_P == 1
    # edge values of the reduction
    for a in (0, 1, GL_P - 1, 0xFFFFFFFF, 0xFFFFFFFF00000000):
        for b in (0, 1, GL_P - 1, 0xFFFFFFFF, 0xFFFFFFFF00000000):
            assert oracle.lib().orc_gl_mul(a, b) == a * b % GL_P


def test_fft_matches_naive_dft_and_inverts(oracle):
    rng = np.random.default_rng(1)
    n = 16
    a = rand_field(rng, (n,))
    w = oracle.lib().orc_gl_root_of_unity(4)
    want = [sum(int(a[j]) * pow(w, i * j, GL_P) for j in range(n)) % GL_P for i in range(n)]
    assert [int(x) for x in oracle.fft(a)] == want
    b = rand_field(rng, (1 << 12,))
    assert (oracle.ifft(oracle.fft(b)) == b).all()


def test_lde_is_the_coset_evaluation(oracle):
    """PolynomialBatch::from_values: lde[j] = P(7 w'^j), P interpolating the values on <w>."""
    rng = np.random.default_rng(2)
    v = rand_field(rng, (2, 32))
    coeffs, lde = oracle.lde_batch(v, 1)
    w2 = oracle.lib().orc_gl_root_of_unity(6)
    for c in range(2):
        for j in (0, 1, 17, 63):
            x = 7 * pow(w2, j, GL_P) % GL_P
            assert int(lde[c, j]) == sum(int(coeffs[c, i]) * pow(x, i, GL_P) for i in range(32)) % GL_P
        w1 = oracle.lib().orc_gl_root_of_unity(5)
        for i in (0, 3, 31):  # the coefficients interpolate the values
            assert int(v[c, i]) == sum(int(coeffs[c, k]) * pow(w1, i * k, GL_P) for k in range(32)) % GL_P


def test_hash_or_noop_and_merkle_structure(oracle):
    """hash_no_pad overwrites rate lanes chunk by chunk; leaves of <= 4 elements are not hashed
    (hash_or_noop); inner nodes are two_to_one."""
    x = np.arange(1, 12, dtype=np.uint64)
    st = np.zeros(12, dtype=np.uint64)
    st[:8] = x[:8]
    st = oracle.poseidon_permute(st)
    st[:3] = x[8:]
    st = oracle.poseidon_permute(st)
    assert (oracle.hash_no_pad(x) == st[:4]).all()
    rng = np.random.default_rng(3)
    leaves = rand_field(rng, (8, 3))  # narrow leaves: digest = zero-padded leaf
    pad = np.zeros((8, 4), dtype=np.uint64)
    pad[:, :3] = leaves
    lvl = [pad[i] for i in range(8)]
    while len(lvl) > 2:
        lvl = [oracle.two_to_one(lvl[2 * i], lvl[2 * i + 1]) for i in range(len(lvl) // 2)]
    assert (oracle.merkle_cap(leaves, 1) == np.array(lvl)).all()
    wide = rand_field(rng, (4, 9))
    cap = oracle.merkle_cap(wide, 2)
    assert all((cap[i] == oracle.hash_no_pad(wide[i])).all() for i in range(4))


def test_commit_golden(oracle, golden):
    import hashlib
    for g in golden["commit"]:
        rng = np.random.default_rng(g["seed"])
        break
    rng = np.random.default_rng(golden["commit"][0]["seed"])
    for g in golden["commit"]:
        hi = rng.integers(0, 1 << 32, size=(g["cols"], 1 << g["log_n"]), dtype=np.uint64)
        lo = rng.integers(0, 1 << 32, size=(g["cols"], 1 << g["log_n"]), dtype=np.uint64)
        v = (hi << np.uint64(32)) | lo
        v = np.where(v >= np.uint64(GL_P), v - np.uint64(GL_P), v)
        assert hashlib.sha256(v.tobytes()).hexdigest() == g["values_sha256"]
        _, lde = oracle.lde_batch(v, g["rate_bits"])
        assert hashlib.sha256(lde.tobytes()).hexdigest() == g["lde_sha256"]
        assert oracle.commit(v, g["rate_bits"], g["cap_height"]).tolist() == g["cap"]


def test_streamed_commit_equals_one_shot_commit(oracle):
    """oracle.commit_streamed (8-column chunks, bounded memory; used for the config-4 cap at 2^24 LDE rows) is the
    same function as oracle.commit."""
    from util import rand_field
    rng = np.random.default_rng(5)
    for cols in (5, 8, 13, 27):
        v = rand_field(rng, (cols, 128))
        for rate_bits, cap_height in ((1, 2), (3, 4)):
            assert (oracle.commit(v, rate_bits, cap_height) == oracle.commit_streamed(v, rate_bits, cap_height)).all()
