//! One-command golden harness: closes "parity unpinned"  for whoever has cargo.
//!
//!   PB254_LIB_DIR=…/plonky2_bn254_b200 PB254_GOLDEN=…/tests/golden cargo test --release -- --nocapture
//!
//! UNCOMPILED here (no Rust toolchain in the build image). It runs the UNMODIFIED reference crate
//! (`plonky2_bn254`, added as a dev-dependency by path or git) on the inputs stored in tests/golden/golden.json
//! and compares with what the oracle / the B200 prover produced:
//!   1. `generate_trace` of the three STARKs: sha256 of the column-major little-endian u64 trace;
//!   2. `prove` of the fq_exp batch: every section of the proof that does not depend on the proof-of-work witness
//!      (init_challenger_state, the three caps, all openings, the commit-phase caps, final_poly) — the reference's
//!      rayon `find_any` returns an arbitrary valid witness, so its query rounds may legitimately differ;
//!   3. the reference's own `verify()` accepts the golden blob (the oracle's / GPU's proof) decoded by this crate.
use ark_bn254::{Fq, Fq2, G1Affine, G2Affine};
use ark_ff::{BigInt, PrimeField};
use num_bigint::BigUint;
use plonky2::field::extension::quadratic::QuadraticExtension;
use plonky2::field::extension::FieldExtension;
use plonky2::field::goldilocks_field::GoldilocksField;
use plonky2::field::types::PrimeField64;
use plonky2::plonk::config::PoseidonGoldilocksConfig;
use plonky2::util::timing::TimingTree;
use plonky2_bn254::starks::common::{prover::prove, verifier::verify};
use plonky2_bn254::starks::curves::g1::scalar_mul_stark::{G1ScalarMulInput, G1ScalarMulStark};
use plonky2_bn254::starks::curves::g2::scalar_mul_stark::{G2ScalarMulInput, G2ScalarMulStark};
use plonky2_bn254::starks::fields::exp_ctl::{fq_exp_ctl, fq_generate_ctl_values};
use plonky2_bn254::starks::fields::exp_stark::{FqExpInput, FqExpStark};
use serde_json::Value;
use sha2::{Digest, Sha256};
use starky::config::StarkConfig;

type F = GoldilocksField;
type C = PoseidonGoldilocksConfig;
type FE = QuadraticExtension<F>;
const D: usize = 2;

fn golden_dir() -> std::path::PathBuf {
    std::env::var("PB254_GOLDEN").map(Into::into).unwrap_or_else(|_| "../../tests/golden".into())
}
fn sha_words(w: impl Iterator<Item = u64>) -> String {
    let mut h = Sha256::new();
    for x in w {
        h.update(x.to_le_bytes());
    }
    hex::encode(h.finalize())
}
fn fq(w: &[u64]) -> Fq {
    Fq::from_bigint(BigInt([w[0], w[1], w[2], w[3]])).unwrap()
}
fn fq2(w: &[u64]) -> Fq2 {
    Fq2::new(fq(&w[..4]), fq(&w[4..8]))
}
fn big(w: &[u64]) -> BigUint {
    BigUint::from_bytes_le(&w.iter().flat_map(|x| x.to_le_bytes()).collect::<Vec<_>>())
}
fn rows(v: &Value) -> Vec<Vec<u64>> {
    v.as_array().unwrap().iter().map(|r| r.as_array().unwrap().iter().map(|x| x.as_u64().unwrap()).collect()).collect()
}
fn trace_sha(trace: &[plonky2::field::polynomial::PolynomialValues<F>]) -> String {
    sha_words(trace.iter().flat_map(|c| c.values.iter().map(|v| v.to_canonical_u64())))
}
fn fq_inputs(case: &Value) -> Vec<(FqExpInput, usize)> {
    let ts = rows(&Value::Array(vec![case["timestamps"].clone()]))[0].clone();
    rows(&case["inputs"]).iter().zip(ts).map(|(r, t)| (FqExpInput { s: big(&r[..4]), x: fq(&r[4..8]) }, t as usize)).collect()
}

#[test]
fn traces_match_the_golden_hashes() {
    let g: Value = serde_json::from_reader(std::fs::File::open(golden_dir().join("golden.json")).unwrap()).unwrap();
    for case in g["traces"].as_array().unwrap() {
        let ts = rows(&Value::Array(vec![case["timestamps"].clone()]))[0].clone();
        let inp = rows(&case["inputs"]);
        let got = match case["kind"].as_u64().unwrap() {
            0 => {
                let v: Vec<_> = inp.iter().zip(&ts).map(|(r, &t)| {
                    (G1ScalarMulInput { s: big(&r[..4]), x: G1Affine::new(fq(&r[4..8]), fq(&r[8..12])), offset: G1Affine::new(fq(&r[12..16]), fq(&r[16..20])) }, t as usize)
                }).collect();
                trace_sha(&G1ScalarMulStark::<F, D>::new().generate_trace(&v, 1 << 16))
            }
            1 => {
                let v: Vec<_> = inp.iter().zip(&ts).map(|(r, &t)| {
                    (G2ScalarMulInput { s: big(&r[..4]), x: G2Affine::new(fq2(&r[4..12]), fq2(&r[12..20])), offset: G2Affine::new(fq2(&r[20..28]), fq2(&r[28..36])) }, t as usize)
                }).collect();
                trace_sha(&G2ScalarMulStark::<F, D>::new().generate_trace(&v, 1 << 16))
            }
            _ => trace_sha(&FqExpStark::<F, D>::new().generate_trace(&fq_inputs(case), 1 << 16)),
        };
        assert_eq!(got, case["trace_sha256"].as_str().unwrap(), "trace of kind {}", case["kind"]);
    }
}

#[test]
fn fq_exp_proof_sections_match_and_reference_verifies_the_golden_blob() {
    let g: Value = serde_json::from_reader(std::fs::File::open(golden_dir().join("golden.json")).unwrap()).unwrap();
    let case = &g["proofs"][0];
    let inputs = fq_inputs(case);
    let stark = FqExpStark::<F, D>::new();
    let config = StarkConfig::standard_fast_config();
    let ctls = fq_exp_ctl::<F>();
    let extra = fq_generate_ctl_values::<F>(&inputs);

    // (3) the reference's verifier on the golden blob
    let bytes = std::fs::read(golden_dir().join(case["blob_file"].as_str().unwrap())).unwrap();
    let blob: Vec<u64> = bytes.chunks(8).map(|c| u64::from_le_bytes(c.try_into().unwrap())).collect();
    assert_eq!(sha_words(blob.iter().copied()), case["proof_sha256"].as_str().unwrap());
    let golden_proof = pb254::decode_proof(&blob).unwrap();
    verify(&stark, &config, &ctls, &golden_proof, &[], &extra).expect("reference verify() rejects the golden proof");

    // (2) the reference's prover, section by section
    let trace = stark.generate_trace(&inputs, 1 << 16);
    let p = prove::<F, C, _, D>(&stark, &config, &trace, &ctls, &[], &mut TimingTree::default()).unwrap();
    let l = pb254::layout_of(&blob).unwrap();
    let sec = |off: u64, n: u64| sha_words(blob[off as usize..(off + n) as usize].iter().copied());
    let ext = |v: &[FE]| v.iter().flat_map(|e| { let a: [F; 2] = e.to_basefield_array(); [a[0].to_canonical_u64(), a[1].to_canonical_u64()] }).collect::<Vec<u64>>();
    let cap = |c: &plonky2::hash::merkle_tree::MerkleCap<F, plonky2::hash::poseidon::PoseidonHash>| c.0.iter().flat_map(|h| h.elements.map(|e| e.to_canonical_u64())).collect::<Vec<u64>>();
    let want = &case["sections_sha256"];
    use plonky2::hash::hashing::PlonkyPermutation;
    let state: Vec<u64> = p.init_challenger_state.as_ref().iter().map(|e| e.to_canonical_u64()).collect();
    assert_eq!(sha_words(state.into_iter()), want["init_challenger_state"].as_str().unwrap());
    let mut caps = cap(&p.proof.trace_cap);
    caps.extend(cap(p.proof.auxiliary_polys_cap.as_ref().unwrap()));
    caps.extend(cap(p.proof.quotient_polys_cap.as_ref().unwrap()));
    assert_eq!(sha_words(caps.into_iter()), want["caps"].as_str().unwrap());
    let o = &p.proof.openings;
    let mut op = ext(&o.local_values);
    op.extend(ext(&o.next_values));
    op.extend(ext(o.auxiliary_polys.as_ref().unwrap()));
    op.extend(ext(o.auxiliary_polys_next.as_ref().unwrap()));
    op.extend(o.ctl_zs_first.as_ref().unwrap().iter().map(|e| e.to_canonical_u64()));
    op.extend(ext(o.quotient_polys.as_ref().unwrap()));
    assert_eq!(sha_words(op.into_iter()), want["openings"].as_str().unwrap());
    let fc: Vec<u64> = p.proof.opening_proof.commit_phase_merkle_caps.iter().flat_map(|c| cap(c)).collect();
    assert_eq!(sha_words(fc.into_iter()), want["commit_phase_merkle_caps"].as_str().unwrap());
    assert_eq!(sha_words(ext(&p.proof.opening_proof.final_poly.coeffs).into_iter()), want["final_poly"].as_str().unwrap());
    // sanity: the same sections of the golden blob itself
    assert_eq!(sec(l.final_poly, l.final_poly_words), want["final_poly"].as_str().unwrap());
}
