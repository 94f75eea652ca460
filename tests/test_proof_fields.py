"""pb254_proof_parse / ffi.ProofView: the fields of StarkProofWithMetadata rebuilt from the serialized proof
(what set_stark_proof_target walks, /root/reference/src/generators/g1/stark_proof.rs:173-178): parse -> serialize
round trip, every field where the verifier expects it (tampering through the parsed offsets is rejected with the
field's own check), Merkle paths of the parsed query rounds close on the parsed caps. Host code, runs on the CPU tier."""
import numpy as np
import pytest

from plonky2_bn254_b200 import ffi, inputs as I
from util import GL_P


@pytest.fixture(scope="module")
def hs_lib():
    from plonky2_bn254_b200 import build
    return ffi.Library(build.build_hostsim())


def test_layout_and_round_trip(hs_lib, fq_case):
    w = fq_case["words"]
    v = hs_lib.parse_proof(w)
    l = v.layout
    assert (v.kind, v.degree_bits) == (I.KIND_FQ, 16)
    assert tuple(getattr(l.config, f) for f, _ in l.config._fields_) == (1, 4, 2, 84, 16, 4, 5)
    assert (l.trace_width, l.aux_width, l.quotient_width, l.num_ctl_zs) == (427, 134, 4, 4)
    assert l.num_fri_layers == 3 and list(l.fri_arity_bits[:3]) == [4, 4, 4]  # 16 -> 12 -> 8 -> 4 <= final_poly_bits
    assert l.words == w.size and l.pow_witness == w.size - 1 and l.final_poly_words == 2 << 4
    assert v.trace_cap.shape == (16, 4) and v.openings["local_values"].shape == (427, 2)
    assert v.commit_phase_merkle_caps.shape == (3, 16, 4) and len(v.query_round_proofs) == 84
    assert v.query_round_proofs[0]["initial_trees_proof"][0][1].shape == (17 - 4, 4)
    assert (v.serialize() == w).all()


def test_fields_are_where_the_verifier_reads_them(hs_lib, fq_case):
    w0 = fq_case["words"]
    l = hs_lib.parse_proof(w0).layout
    q3 = int(l.query_round_proofs) + 3 * int(l.query_words)
    sites = {
        "init_challenger_state": (int(l.init_challenger_state) + 11, "init_challenger_state"),
        "auxiliary_polys_cap": (int(l.auxiliary_polys_cap) + 5, None),
        "quotient_polys_cap": (int(l.quotient_polys_cap) + 63, None),
        "next_values": (int(l.next_values) + 2 * 400, None),
        "auxiliary_polys_next": (int(l.auxiliary_polys_next) + 9, None),
        "ctl_zs_first": (int(l.ctl_zs_first) + 3, None),
        "quotient_polys": (int(l.quotient_polys) + 7, None),
        "commit_phase_merkle_caps": (int(l.commit_phase_merkle_caps) + 2 * int(l.cap_words) + 1, None),
        "trace leaf": (q3 + l.q_trace_leaf + 100, "Merkle path of the trace"),
        "aux path": (q3 + l.q_aux_path + 4, "Merkle path of the aux"),
        "quotient leaf": (q3 + l.q_quotient_leaf + 1, "Merkle path of the quotient"),
        "step 1 evals": (q3 + l.q_step_evals[1] + 3, None),
        "step 2 path": (q3 + l.q_step_path[2] + 2, None),
        "final_poly": (int(l.final_poly) + 1, None),
        "pow_witness": (int(l.pow_witness), None),
    }
    for name, (pos, msg) in sites.items():
        w = w0.copy()
        w[pos] = (int(w[pos]) + 1) % GL_P
        with pytest.raises(ffi.Pb254Error) as e:
            hs_lib.verify(I.KIND_FQ, w, fq_case["inputs"], fq_case["timestamps"])
        assert e.value.code == 7, name
        if msg:
            assert msg in str(e.value), (name, str(e.value))


def test_parsed_merkle_paths_close_on_the_parsed_caps(hs_lib, fq_case, oracle):
    """Independent of the product verifier: leaf hash + siblings of a parsed query round reach the parsed cap."""
    v = hs_lib.parse_proof(fq_case["words"])

    def hash_or_noop(vals):
        if len(vals) <= 4:
            return np.concatenate([vals, np.zeros(4 - len(vals), dtype=np.uint64)])
        s = np.zeros(12, dtype=np.uint64)
        for c in range(0, len(vals), 8):
            chunk = vals[c:c + 8]
            s[:len(chunk)] = chunk
            s = oracle.poseidon_permute(s)
        return s[:4].copy()

    def two_to_one(a, b):
        return oracle.poseidon_permute(np.concatenate([a, b, np.zeros(4, dtype=np.uint64)]))[:4].copy()

    indices = fq_case["proof"].debug(3)  # the query indices the oracle derived from the transcript (leaf positions)
    assert indices.size == 84
    for q in (0, 41, 83):
        for (leaf, path), cap in zip(v.query_round_proofs[q]["initial_trees_proof"],
                                     (v.trace_cap, v.auxiliary_polys_cap, v.quotient_polys_cap)):
            idx = int(indices[q])
            cur = hash_or_noop(leaf)
            for sib in path:
                cur = two_to_one(sib, cur) if idx & 1 else two_to_one(cur, sib)
                idx >>= 1
            assert (cur == cap[idx]).all()


def test_parse_rejects_foreign_or_truncated_words(hs_lib, fq_case):
    w = fq_case["words"]
    for bad in (w[:-1], np.concatenate([w, w[:1]]), np.zeros(64, dtype=np.uint64)):
        with pytest.raises(ffi.Pb254Error) as e:
            hs_lib.parse_proof(bad)
        assert e.value.code == 6
    w2 = w.copy()
    w2[1] = 7  # unknown kind
    with pytest.raises(ffi.Pb254Error):
        hs_lib.parse_proof(w2)


def test_golden_blob_fixture(hs_lib, golden, fq_case):
    """tests/golden/fq3_proof.bin (input of rust/pb254/tests/golden.rs) is the oracle's proof of golden case 42 and
    its pow-independent sections hash to the stored values."""
    import hashlib
    import os
    case = golden["proofs"][0]
    w = np.fromfile(os.path.join(os.path.dirname(__file__), "golden", case["blob_file"]), dtype="<u8")
    assert hashlib.sha256(w.tobytes()).hexdigest() == case["proof_sha256"]
    assert (w == fq_case["words"]).all()  # conftest's fq_case is the same batch (config_seed(42), 3 instances)
    inp = np.array(case["inputs"], dtype=np.uint64)
    assert (inp == fq_case["inputs"]).all()
    l = hs_lib.parse_proof(w).layout
    sec = w[int(l.local_values):int(l.commit_phase_merkle_caps)]
    assert hashlib.sha256(sec.tobytes()).hexdigest() == case["sections_sha256"]["openings"]
    assert hs_lib.verify(I.KIND_FQ, w, inp, np.array(case["timestamps"], dtype=np.uint64))
