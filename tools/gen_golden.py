"""Generates tests/golden/golden.json from the CPU oracle (run in the build container):

    python tools/gen_golden.py

The reference's own tests hold no golden vectors for this path (all 32 tests draw from thread_rng,
SURVEY.md 8c) and the Rust crate cannot be compiled offline, so these vectors pin the ORACLE against
regressions; they are not outputs of the Rust crate ("parity unpinned", see oracle/README.md). The two
Poseidon known-answer vectors ARE upstream's (plonky2 poseidon_goldilocks.rs test vectors).
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402
from plonky2_bn254_b200 import inputs as I  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()


def main():
    out = {"poseidon_kat": {}, "traces": [], "commit": [], "proofs": []}
    out["poseidon_kat"]["zeros"] = [int(x) for x in O.poseidon_permute(np.zeros(12, dtype=np.uint64))]
    out["poseidon_kat"]["range12"] = [int(x) for x in O.poseidon_permute(np.arange(12, dtype=np.uint64))]
    out["poseidon_round_constants_sha256"] = sha(O.poseidon_round_constants())
    for kind, k, cid in ((I.KIND_G1, 2, 40), (I.KIND_G2, 1, 41), (I.KIND_FQ, 3, 42)):
        inp, ts = I.make_inputs(kind, k, I.config_seed(cid))
        tr, res = O.generate_trace(kind, inp, ts, want_results=True)
        out["traces"].append({"kind": kind, "instances": k, "config_id": cid, "inputs_sha256": sha(inp),
                              "inputs": [[int(x) for x in r] for r in inp], "timestamps": [int(x) for x in ts],
                              "trace_sha256": sha(tr), "shape": list(tr.shape),
                              "result_limbs": [[int(x) for x in r] for r in res],
                              "frequency_first8": [int(x) for x in tr[tr.shape[0] - 2, :8]]})
    rng = np.random.default_rng(1234)
    for cols, log_n, rb, cap in ((5, 8, 1, 4), (9, 10, 1, 4), (3, 10, 3, 0)):
        hi = rng.integers(0, 1 << 32, size=(cols, 1 << log_n), dtype=np.uint64)
        lo = rng.integers(0, 1 << 32, size=(cols, 1 << log_n), dtype=np.uint64)
        v = (hi << np.uint64(32)) | lo
        v = np.where(v >= np.uint64(0xFFFFFFFF00000001), v - np.uint64(0xFFFFFFFF00000001), v)
        capv = O.commit(v, rb, cap)
        _, lde = O.lde_batch(v, rb)
        out["commit"].append({"cols": cols, "log_n": log_n, "rate_bits": rb, "cap_height": cap, "seed": 1234,
                              "values_sha256": sha(v), "lde_sha256": sha(lde), "cap": [[int(x) for x in r] for r in capv]})
    for kind, k, cid in ((I.KIND_FQ, 3, 42),):
        inp, ts = I.make_inputs(kind, k, I.config_seed(cid))
        pf, _, _ = O.prove_inputs(kind, inp, ts)
        w = pf.words()
        O.verify(w, inp, ts)
        # sections that do not depend on the proof-of-work witness (the reference's find_any returns an arbitrary valid
        # witness, so its query rounds may differ): what rust/pb254/tests/golden.rs compares with the Rust crate's proof
        from plonky2_bn254_b200 import build, ffi
        lay = ffi.Library(build.build_hostsim()).parse_proof(w).layout
        sec = {"init_challenger_state": (lay.init_challenger_state, 12),
               "caps": (lay.trace_cap, 3 * lay.cap_words),
               "openings": (lay.local_values, lay.commit_phase_merkle_caps - lay.local_values),
               "commit_phase_merkle_caps": (lay.commit_phase_merkle_caps, lay.num_fri_layers * lay.cap_words),
               "final_poly": (lay.final_poly, lay.final_poly_words)}
        w.astype("<u8").tofile(os.path.join(ROOT, "tests", "golden", "fq3_proof.bin"))
        out["proofs"].append({"kind": kind, "instances": k, "config_id": cid, "words": int(w.size),
                              "proof_sha256": sha(w), "trace_cap_first": [int(x) for x in w[22:26]],
                              "pow_witness": int(w[-1]), "blob_file": "fq3_proof.bin",
                              "inputs": [[int(x) for x in r] for r in inp], "timestamps": [int(x) for x in ts],
                              "sections_sha256": {k2: sha(w[int(a):int(a) + int(n)]) for k2, (a, n) in sec.items()}})
    path = os.path.join(ROOT, "tests", "golden", "golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
