//! Safe wrapper over `pb254-sys` with the reference's types at the surface.
//!
//! UNCOMPILED (no Rust toolchain in the build image). It mirrors the call sites
//!   G1StarkProofGenerator::run_once   src/generators/g1/stark_proof.rs:136-179
//!   prove()                           src/starks/common/prover.rs:18-72
//!   verify()                          src/starks/common/verifier.rs:32-98
//! and rebuilds `StarkProofWithMetadata<F, C, D>` from the proof blob (layout: INTEGRATION.md §3).
use anyhow::{anyhow, Result};
use ark_bn254::{Fq, G1Affine};
use ark_ff::{BigInteger, PrimeField};
use num_bigint::BigUint;
use pb254_sys as sys;
use plonky2::field::extension::quadratic::QuadraticExtension;
use plonky2::field::goldilocks_field::GoldilocksField;
use plonky2::field::types::Field;
use plonky2::fri::proof::{FriInitialTreeProof, FriProof, FriQueryRound, FriQueryStep};
use plonky2::hash::hash_types::HashOut;
use plonky2::hash::merkle_proofs::MerkleProof;
use plonky2::hash::merkle_tree::MerkleCap;
use plonky2::hash::poseidon::PoseidonHash;
use plonky2::plonk::config::PoseidonGoldilocksConfig;
use starky::proof::{StarkOpeningSet, StarkProof, StarkProofWithMetadata};
use std::ffi::CStr;

type F = GoldilocksField;
type C = PoseidonGoldilocksConfig;
type FE = QuadraticExtension<F>;
const D: usize = 2;

/// One prover context per GPU (not `Sync`: one thread at a time, like the reference's generator loop).
pub struct Context(*mut sys::pb254_ctx);
unsafe impl Send for Context {}

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::pb254_last_error()).to_string_lossy().into_owned() }
}
fn check(rc: i32) -> Result<()> {
    if rc == sys::PB254_OK { Ok(()) } else { Err(anyhow!("pb254 error {rc}: {}", last_error())) }
}

fn push_biguint(out: &mut Vec<u64>, x: &BigUint) {
    let mut d = x.to_u64_digits();
    assert!(d.len() <= 4, "scalar wider than 256 bits (common/utils.rs:4)");
    d.resize(4, 0);
    out.extend_from_slice(&d);
}
fn push_fq(out: &mut Vec<u64>, x: &Fq) {
    out.extend_from_slice(&x.into_bigint().0); // canonical little-endian 4 x u64
}

/// Wire format of `G1ScalarMulInput {s, x, offset}` (src/starks/curves/g1/scalar_mul_stark.rs:37-41).
pub fn pack_g1(inputs: &[(BigUint, G1Affine, G1Affine)]) -> Vec<u64> {
    let mut w = Vec::with_capacity(inputs.len() * 20);
    for (s, x, off) in inputs {
        push_biguint(&mut w, s);
        push_fq(&mut w, &x.x);
        push_fq(&mut w, &x.y);
        push_fq(&mut w, &off.x);
        push_fq(&mut w, &off.y);
    }
    w
}

impl Context {
    pub fn new(device: i32) -> Result<Self> {
        let mut p = std::ptr::null_mut();
        check(unsafe { sys::pb254_ctx_create(device, std::ptr::null_mut(), &mut p) })?;
        Ok(Context(p))
    }

    /// `generate_trace` + `prove` (run_once lines 154-163) for `kind`; `words` as produced by `pack_*`.
    pub fn prove(&self, kind: i32, words: &[u64], timestamps: &[u64]) -> Result<StarkProofWithMetadata<F, C, D>> {
        let mut h = std::ptr::null_mut();
        check(unsafe {
            sys::pb254_prove(self.0, kind, words.as_ptr(), timestamps.as_ptr(), timestamps.len(), 1 << 16,
                             std::ptr::null(), 0, &mut h)
        })?;
        let blob = unsafe { std::slice::from_raw_parts(sys::pb254_proof_data(h), sys::pb254_proof_words(h)) }.to_vec();
        unsafe { sys::pb254_proof_free(h) };
        decode_proof(&blob)
    }
}
impl Drop for Context {
    fn drop(&mut self) {
        unsafe { sys::pb254_ctx_destroy(self.0) }
    }
}

/// `verify()` on the serialized proof (host side of libpb254).
pub fn verify(blob: &[u64], kind_words: &[u64], timestamps: &[u64]) -> Result<()> {
    check(unsafe { sys::pb254_verify(blob.as_ptr(), blob.len(), kind_words.as_ptr(), timestamps.as_ptr(), timestamps.len()) })
}

struct Reader<'a>(&'a [u64], usize);
impl<'a> Reader<'a> {
    fn f(&mut self) -> F { let v = F::from_canonical_u64(self.0[self.1]); self.1 += 1; v }
    fn ext(&mut self) -> FE { let a = self.f(); let b = self.f(); FE::from_basefield_array([a, b]) }
    fn fs(&mut self, n: usize) -> Vec<F> { (0..n).map(|_| self.f()).collect() }
    fn exts(&mut self, n: usize) -> Vec<FE> { (0..n).map(|_| self.ext()).collect() }
    fn hash(&mut self) -> HashOut<F> { HashOut { elements: [self.f(), self.f(), self.f(), self.f()] } }
    fn cap(&mut self, n: usize) -> MerkleCap<F, PoseidonHash> { MerkleCap((0..n).map(|_| self.hash()).collect()) }
    fn path(&mut self, n: usize) -> MerkleProof<F, PoseidonHash> { MerkleProof { siblings: (0..n).map(|_| self.hash()).collect() } }
}

/// Blob -> `StarkProofWithMetadata` (field order of SURVEY.md C.7 / INTEGRATION.md §3).
pub fn decode_proof(blob: &[u64]) -> Result<StarkProofWithMetadata<F, C, D>> {
    let (kind, l) = (blob[1] as usize, blob[2] as usize);
    let (rate_bits, cap_h, nch, nq, arity_bits, final_bits) =
        (blob[3] as usize, blob[4] as usize, blob[5] as usize, blob[6] as usize, blob[8] as usize, blob[9] as usize);
    let w = unsafe { sys::pb254_trace_width(kind as i32) } as usize;
    let a = unsafe { sys::pb254_num_aux(kind as i32, nch as u32) } as usize;
    let q = 2 * nch;
    let log_n = l + rate_bits;
    let mut arities = vec![];
    let mut db = l;
    while db > final_bits && db + rate_bits >= cap_h + arity_bits { arities.push(arity_bits); db -= arity_bits; }
    let mut r = Reader(blob, 10);
    let init_challenger_state = { let v = r.fs(12); <C as plonky2::plonk::config::GenericConfig<D>>::Hasher::Permutation::new(v) };
    let ncap = 1 << cap_h;
    let trace_cap = r.cap(ncap);
    let auxiliary_polys_cap = Some(r.cap(ncap));
    let quotient_polys_cap = Some(r.cap(ncap));
    let local_values = r.exts(w);
    let next_values = r.exts(w);
    let auxiliary_polys = Some(r.exts(a));
    let auxiliary_polys_next = Some(r.exts(a));
    let ctl_zs_first = Some(r.fs(2 * nch));
    let quotient_polys = Some(r.exts(q));
    let commit_phase_merkle_caps = arities.iter().map(|_| r.cap(ncap)).collect();
    let mut query_round_proofs = Vec::with_capacity(nq);
    for _ in 0..nq {
        let mut evals_proofs = vec![];
        for cols in [w, a, q] {
            let leaf = r.fs(cols);
            evals_proofs.push((leaf, r.path(log_n - cap_h)));
        }
        let mut steps = vec![];
        let mut ll = log_n;
        for ab in &arities {
            ll -= ab;
            let evals = r.exts(1 << ab);
            steps.push(FriQueryStep { evals, merkle_proof: r.path(ll - cap_h) });
        }
        query_round_proofs.push(FriQueryRound { initial_trees_proof: FriInitialTreeProof { evals_proofs }, steps });
    }
    let final_poly = plonky2::field::polynomial::PolynomialCoeffs::new(r.exts(1 << db));
    let pow_witness = r.f();
    let opening_proof = FriProof { commit_phase_merkle_caps, query_round_proofs, final_poly, pow_witness };
    let openings = StarkOpeningSet { local_values, next_values, auxiliary_polys, auxiliary_polys_next, ctl_zs_first, quotient_polys };
    let proof = StarkProof { trace_cap, auxiliary_polys_cap, quotient_polys_cap, openings, opening_proof };
    Ok(StarkProofWithMetadata { init_challenger_state, proof })
}
