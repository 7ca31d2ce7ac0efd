//! Bodies that replace the stock ones in a fork of halo2_proofs (tag v2023_04_20) when `C = bn256::G1Affine`.
//! NOT compiled in the backend's environment (no Rust toolchain).  File / function names are the fork's; INTEGRATION.md has
//! the Cargo `[patch]` stanza.  `Fr`, `Fq`, `G1Affine`, `G1` of halo2curves 0.3 are plain `[u64; 4]` aggregates in Montgomery
//! form, so slices are passed without copies.  Every call asserts `rc == 0`: the Rust originals are infallible apart from
//! their own `assert!`s, which the library reports as DE_ERR_ARG.  There is no CPU fallback.
use de_b200_sys::*;
use std::ffi::CStr;

thread_local! { static CTX: *mut de_ctx = unsafe { let mut c = std::ptr::null_mut(); ok(de_ctx_create(0, &mut c), std::ptr::null_mut()); c }; }
fn ctx() -> *mut de_ctx { CTX.with(|c| *c) }
// one context per visible GPU, largest power of two of them (DE_B200_GPUS caps it); element 0 is ctx()
thread_local! { static ALL: Vec<*mut de_ctx> = unsafe {
    let want: usize = std::env::var("DE_B200_GPUS").ok().and_then(|v| v.parse().ok()).unwrap_or(8);
    let mut v = vec![ctx()];
    for dev in 1..want { let mut c = std::ptr::null_mut(); if de_ctx_create(dev as i32, &mut c) != 0 { break; } v.push(c); }
    let mut n = 1; while n * 2 <= v.len() { n *= 2; } v.truncate(n); v
}; }
fn all_ctxs() -> Vec<*mut de_ctx> { ALL.with(|v| v.clone()) }
fn ok(rc: i32, c: *mut de_ctx) { if rc != 0 { panic!("de_b200: {}", unsafe { CStr::from_ptr(de_last_error(c)) }.to_string_lossy()); } }

// ---- src/arithmetic.rs -------------------------------------------------------------------------------------------------
pub fn best_multiexp(coeffs: &[Fr], bases: &[G1Affine]) -> G1 {
    assert_eq!(coeffs.len(), bases.len());
    let mut out = std::mem::MaybeUninit::<de_g1>::uninit();
    unsafe { ok(de_msm(ctx(), coeffs.as_ptr() as _, bases.as_ptr() as _, coeffs.len(), out.as_mut_ptr()), ctx()); std::mem::transmute(out.assume_init()) }
}
pub fn best_fft(a: &mut [Fr], omega: Fr, log_n: u32) {
    assert_eq!(a.len(), 1 << log_n);
    // long vectors go over every GPU of the box (all_ctxs(): one de_ctx per visible device, a power of two of them): block r travels
    // over GPU r's PCIe link and the two transposes of the four-step transform are NVLink peer stores inside the kernels
    let gpus = all_ctxs();
    if gpus.len() > 1 && log_n >= 22 {
        unsafe { ok(de_ntt_sharded(gpus.as_ptr(), gpus.len() as i32, a.as_mut_ptr() as _, &omega as *const Fr as _, log_n), ctx()) }
    } else {
        unsafe { ok(de_ntt(ctx(), a.as_mut_ptr() as _, &omega as *const Fr as _, log_n), ctx()) }
    }
}
pub fn eval_polynomial(poly: &[Fr], point: Fr) -> Fr {
    let mut out = Fr::zero();
    unsafe { ok(de_eval_polynomial(ctx(), poly.as_ptr() as _, poly.len(), &point as *const Fr as _, &mut out as *mut Fr as _), ctx()) };
    out
}
pub fn kate_division(a: &[Fr], b: Fr) -> Vec<Fr> {
    let mut q = vec![Fr::zero(); a.len() - 1];
    unsafe { ok(de_kate_division(ctx(), a.as_ptr() as _, a.len(), &b as *const Fr as _, q.as_mut_ptr() as _), ctx()) };
    q
}

// ---- src/poly/domain.rs: EvaluationDomain gains `dev: *mut de_domain` (de_domain_create in new(), de_domain_free in Drop) ----
impl EvaluationDomain<Fr> {
    pub fn coeff_to_extended(&self, p: Polynomial<Fr, Coeff>) -> Polynomial<Fr, ExtendedLagrangeCoeff> {
        let mut out = vec![Fr::zero(); self.extended_len()];
        unsafe { ok(de_coeff_to_extended(self.dev, p.values.as_ptr() as _, out.as_mut_ptr() as _), ctx()) };
        Polynomial { values: out, _marker: PhantomData }
    }
    pub fn extended_to_coeff(&self, mut p: Polynomial<Fr, ExtendedLagrangeCoeff>) -> Vec<Fr> {
        let mut len = 0usize;
        unsafe { ok(de_extended_to_coeff(self.dev, p.values.as_mut_ptr() as _, &mut len), ctx()) };
        p.values.truncate(len);
        p.values
    }
    pub fn lagrange_to_coeff(&self, mut p: Polynomial<Fr, LagrangeCoeff>) -> Polynomial<Fr, Coeff> {
        unsafe { ok(de_lagrange_to_coeff(self.dev, p.values.as_mut_ptr() as _), ctx()) };
        Polynomial { values: p.values, _marker: PhantomData }
    }
    pub fn coeff_to_lagrange(&self, mut p: Polynomial<Fr, Coeff>) -> Polynomial<Fr, LagrangeCoeff> {
        unsafe { ok(de_coeff_to_lagrange(self.dev, p.values.as_mut_ptr() as _), ctx()) };
        Polynomial { values: p.values, _marker: PhantomData }
    }
    pub fn divide_by_vanishing_poly(&self, mut p: Polynomial<Fr, ExtendedLagrangeCoeff>) -> Polynomial<Fr, ExtendedLagrangeCoeff> {
        unsafe { ok(de_divide_by_vanishing(self.dev, p.values.as_mut_ptr() as _), ctx()) };
        p
    }
}

// ---- src/poly/kzg/commitment.rs: ParamsKZG gains `dev: *mut de_params` (de_params_upload once in setup() / read()) ----------
impl ParamsKZG<Bn256> {
    pub fn commit(&self, poly: &Polynomial<Fr, Coeff>, _: Blind<Fr>) -> G1 { self.commit_basis(0, &poly.values) }
    pub fn commit_lagrange(&self, poly: &Polynomial<Fr, LagrangeCoeff>, _: Blind<Fr>) -> G1 { self.commit_basis(1, &poly.values) }
    fn commit_basis(&self, basis: i32, v: &[Fr]) -> G1 {
        let mut out = std::mem::MaybeUninit::<de_g1>::uninit();
        unsafe { ok(de_commit(self.dev, basis, v.as_ptr() as _, v.len(), out.as_mut_ptr()), ctx()); std::mem::transmute(out.assume_init()) }
    }
}

// ---- src/plonk/prover.rs: create_proof for KZGCommitmentScheme<Bn256>, ProverGWC, Blake2bWrite<_, _, Challenge255<_>> ---------
// pk.dev_prover is built once in keygen_pk: de_pk_upload (serialised pk.ev + fixed / sigma polynomials)
// and de_prover_create (cs.advice_queries, cs.fixed_queries, theta-compression graphs per lookup, vk.transcript_repr).
pub fn create_proof_b200<R: RngCore>(pk: &ProvingKey<G1Affine>, advice: &[Polynomial<Fr, LagrangeCoeff>], instances: &[&[Fr]], mut rng: R,
                                     transcript: &mut Blake2bWrite<Vec<u8>, G1Affine, Challenge255<G1Affine>>) {
    let need = unsafe { de_prover_random_count(pk.dev_prover) };
    let randoms: Vec<Fr> = (0..need).map(|_| Fr::random(&mut rng)).collect();   // drawn in create_proof's order
    let adv: Vec<*const de_fr> = advice.iter().map(|p| p.values.as_ptr() as *const de_fr).collect();
    let ins: Vec<*const de_fr> = instances.iter().map(|v| v.as_ptr() as *const de_fr).collect();
    let lens: Vec<usize> = instances.iter().map(|v| v.len()).collect();
    let mut proof = vec![0u8; unsafe { de_prover_proof_size(pk.dev_prover) }];
    let mut len = 0usize;
    unsafe { ok(de_create_proof(pk.dev_prover, adv.as_ptr(), ins.as_ptr(), lens.as_ptr(), randoms.as_ptr() as _, need, proof.as_mut_ptr(), proof.len(), &mut len), ctx()) };
    transcript.extend_proof_bytes(&proof[..len]);   // appends to the writer's inner Vec<u8>
}
