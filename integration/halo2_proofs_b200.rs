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

// A de_ctx (its stream and workspaces) is NOT thread-safe, and a de_params / de_domain / de_prover handle stays bound to the
// context it was created on.  halo2's ParamsKZG, EvaluationDomain and ProvingKey are Sync and routinely shared between rayon or
// prover threads, so every device-side object carries ITS OWN context behind a Mutex: whichever thread calls takes the lock,
// and errors are read from the object's context, not from the calling thread's.  (The free functions above and below use the
// calling thread's private context and need no lock.)  One context per object costs a stream and grow-only workspaces; a
// prover that wants N proofs in flight clones N (ParamsKZG, ProvingKey) pairs, as bench.py's workers do.
pub struct Dev<T> { inner: std::sync::Mutex<(*mut de_ctx, *mut T)> }
unsafe impl<T> Send for Dev<T> {}
unsafe impl<T> Sync for Dev<T> {}
impl<T> Dev<T> {
    /// `make` receives a fresh context on `device` and returns the handle created on it
    pub fn new(device: i32, make: impl FnOnce(*mut de_ctx) -> *mut T) -> Self {
        let mut c = std::ptr::null_mut();
        unsafe { ok(de_ctx_create(device, &mut c), std::ptr::null_mut()) };
        Dev { inner: std::sync::Mutex::new((c, make(c))) }
    }
    /// runs `f(handle)` under the lock and panics with the OBJECT's error string on a non-zero status
    pub fn call(&self, f: impl FnOnce(*mut T) -> i32) {
        let g = self.inner.lock().unwrap();
        ok(f(g.1), g.0);
    }
}

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

// ---- src/poly/domain.rs: EvaluationDomain gains `dev: Dev<de_domain>` (Dev::new(0, |c| de_domain_create(c, j, k, ..)) in new();
//      de_domain_free + de_ctx_destroy in Drop) ----
impl EvaluationDomain<Fr> {
    pub fn coeff_to_extended(&self, p: Polynomial<Fr, Coeff>) -> Polynomial<Fr, ExtendedLagrangeCoeff> {
        let mut out = vec![Fr::zero(); self.extended_len()];
        self.dev.call(|d| unsafe { de_coeff_to_extended(d, p.values.as_ptr() as _, out.as_mut_ptr() as _) });
        Polynomial { values: out, _marker: PhantomData }
    }
    pub fn extended_to_coeff(&self, mut p: Polynomial<Fr, ExtendedLagrangeCoeff>) -> Vec<Fr> {
        let mut len = 0usize;
        self.dev.call(|d| unsafe { de_extended_to_coeff(d, p.values.as_mut_ptr() as _, &mut len) });
        p.values.truncate(len);
        p.values
    }
    pub fn lagrange_to_coeff(&self, mut p: Polynomial<Fr, LagrangeCoeff>) -> Polynomial<Fr, Coeff> {
        self.dev.call(|d| unsafe { de_lagrange_to_coeff(d, p.values.as_mut_ptr() as _) });
        Polynomial { values: p.values, _marker: PhantomData }
    }
    pub fn coeff_to_lagrange(&self, mut p: Polynomial<Fr, Coeff>) -> Polynomial<Fr, LagrangeCoeff> {
        self.dev.call(|d| unsafe { de_coeff_to_lagrange(d, p.values.as_mut_ptr() as _) });
        Polynomial { values: p.values, _marker: PhantomData }
    }
    pub fn divide_by_vanishing_poly(&self, mut p: Polynomial<Fr, ExtendedLagrangeCoeff>) -> Polynomial<Fr, ExtendedLagrangeCoeff> {
        self.dev.call(|d| unsafe { de_divide_by_vanishing(d, p.values.as_mut_ptr() as _) });
        p
    }
}

// ---- src/poly/kzg/commitment.rs: ParamsKZG gains `dev: Dev<de_params>` (de_params_upload once in setup() / read()) ----------
impl ParamsKZG<Bn256> {
    pub fn commit(&self, poly: &Polynomial<Fr, Coeff>, _: Blind<Fr>) -> G1 { self.commit_basis(0, &poly.values) }
    pub fn commit_lagrange(&self, poly: &Polynomial<Fr, LagrangeCoeff>, _: Blind<Fr>) -> G1 { self.commit_basis(1, &poly.values) }
    fn commit_basis(&self, basis: i32, v: &[Fr]) -> G1 {
        let mut out = std::mem::MaybeUninit::<de_g1>::uninit();
        self.dev.call(|p| unsafe { de_commit(p, basis, v.as_ptr() as _, v.len(), out.as_mut_ptr()) });
        unsafe { std::mem::transmute(out.assume_init()) }
    }
}

// ---- src/plonk/prover.rs: create_proof for KZGCommitmentScheme<Bn256>, ProverGWC, Blake2bWrite<_, _, Challenge255<_>> ---------
// pk.dev_prover: Dev<de_prover> is built once in keygen_pk on ONE context (params, domain, pk and prover of a proving key share
// it: de_prover_create refuses handles of different contexts): de_pk_upload (serialised pk.ev + fixed / sigma polynomials)
// and de_prover_create (cs.advice_queries, cs.fixed_queries, theta-compression graphs per lookup, vk.transcript_repr).
pub fn create_proof_b200<R: RngCore>(pk: &ProvingKey<G1Affine>, advice: &[Polynomial<Fr, LagrangeCoeff>], instances: &[&[Fr]], mut rng: R,
                                     transcript: &mut Blake2bWrite<Vec<u8>, G1Affine, Challenge255<G1Affine>>) {
    let (mut need, mut size) = (0usize, 0usize);
    pk.dev_prover.call(|p| unsafe { need = de_prover_random_count(p); size = de_prover_proof_size(p); 0 });
    let randoms: Vec<Fr> = (0..need).map(|_| Fr::random(&mut rng)).collect();   // drawn in create_proof's order
    let adv: Vec<*const de_fr> = advice.iter().map(|p| p.values.as_ptr() as *const de_fr).collect();
    let ins: Vec<*const de_fr> = instances.iter().map(|v| v.as_ptr() as *const de_fr).collect();
    let lens: Vec<usize> = instances.iter().map(|v| v.len()).collect();
    let mut proof = vec![0u8; size];
    let mut len = 0usize;
    // the lock is held for the whole proof: two threads proving with the same ProvingKey serialise (clone the key per thread for
    // concurrency); the error string comes from the prover's own context
    pk.dev_prover.call(|p| unsafe { de_create_proof(p, adv.as_ptr(), ins.as_ptr(), lens.as_ptr(), randoms.as_ptr() as _, need, proof.as_mut_ptr(), proof.len(), &mut len) });
    transcript.extend_proof_bytes(&proof[..len]);   // appends to the writer's inner Vec<u8>
}
