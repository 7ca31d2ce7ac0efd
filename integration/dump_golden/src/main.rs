//! dump_golden — golden vectors from the REAL halo2_proofs (tag v2023_04_20) and the REAL reference circuit, in the formats
//! tests/golden/REF_DUMP_FORMAT.md describes.  Everything random comes from ChaCha20Rng seeds fixed below; the RNG handed to
//! create_proof is wrapped so that every `next_u64` it serves is recorded (Fr::random consumes eight of them per element).
//! NOT compiled in the backend's environment (no Rust toolchain there) - this is the staged half of the parity pin; the
//! consuming half (tests/test_ref_dump.py) is exercised there against a self-made dump of the same format.
use ff::{Field, PrimeField};
use halo2_delay_enc::{poseidon::Spec, DelayEncryptCircuit};
use halo2_proofs::{
    arithmetic::{best_fft, best_multiexp},
    dev::MockProver,
    halo2curves::bn256::{Bn256, Fr, G1Affine},
    halo2curves::group::{Curve, Group},
    plonk::{create_proof, keygen_pk, keygen_vk, verify_proof, Any, Expression},
    poly::{
        commitment::ParamsProver,
        kzg::{commitment::{KZGCommitmentScheme, ParamsKZG}, multiopen::{ProverGWC, VerifierGWC}, strategy::SingleStrategy},
        EvaluationDomain,
    },
    transcript::{Blake2bRead, Blake2bWrite, Challenge255, TranscriptReadBuffer, TranscriptWriterBuffer},
    SerdeFormat,
};
use num_bigint::{BigUint, RandomBits};
use rand_chacha::ChaCha20Rng;
use rand_core::{RngCore, SeedableRng};
use std::{fs::File, io::Write, path::Path};

struct Recording<R: RngCore> { inner: R, log: Vec<u64> }
impl<R: RngCore> RngCore for Recording<R> {
    fn next_u32(&mut self) -> u32 { self.next_u64() as u32 }
    fn next_u64(&mut self) -> u64 { let v = self.inner.next_u64(); self.log.push(v); v }
    fn fill_bytes(&mut self, dest: &mut [u8]) { for c in dest.chunks_mut(8) { let v = self.next_u64().to_le_bytes(); c.copy_from_slice(&v[..c.len()]); } }
    fn try_fill_bytes(&mut self, dest: &mut [u8]) -> Result<(), rand_core::Error> { self.fill_bytes(dest); Ok(()) }
}

fn raw(f: &Fr) -> [u8; 32] { unsafe { std::mem::transmute::<Fr, [u8; 32]>(*f) } }          // Montgomery limbs, as in memory
fn raw_pt(p: &G1Affine) -> [u8; 64] { unsafe { std::mem::transmute::<G1Affine, [u8; 64]>(*p) } }
fn put_frs(w: &mut impl Write, v: &[Fr]) { for f in v { w.write_all(&raw(f)).unwrap(); } }

fn expr_json(e: &Expression<Fr>) -> serde_json::Value {
    use serde_json::json;
    match e {
        Expression::Constant(c) => json!(["const", format!("{:?}", c)]),        // Debug of Fr = 0x + 64 hex digits, canonical
        Expression::Selector(_) => panic!("selectors are compressed away in pk.get_vk().cs()"),
        Expression::Fixed(q) => json!(["fixed", q.column_index(), q.rotation().0]),
        Expression::Advice(q) => json!(["advice", q.column_index(), q.rotation().0]),
        Expression::Instance(q) => json!(["instance", q.column_index(), q.rotation().0]),
        Expression::Challenge(c) => json!(["challenge", c.index()]),
        Expression::Negated(a) => json!(["neg", expr_json(a)]),
        Expression::Sum(a, b) => json!(["sum", expr_json(a), expr_json(b)]),
        Expression::Product(a, b) => json!(["prod", expr_json(a), expr_json(b)]),
        Expression::Scaled(a, c) => json!(["scaled", expr_json(a), format!("{:?}", c)]),
    }
}

fn main() {
    let out = std::env::args().nth(1).expect("usage: dump_golden <output dir>");
    let out = Path::new(&out);
    let mut rng = ChaCha20Rng::seed_from_u64(0xDE1A7E9C0DE);

    // ---- ref_msm.bin: u64 n | scalars | bases | best_multiexp(scalars, bases).to_affine()
    for log_n in [10u32, 16] {
        let n = 1usize << log_n;
        let scalars: Vec<Fr> = (0..n).map(|_| Fr::random(&mut rng)).collect();
        let g = G1Affine::generator();
        let bases: Vec<G1Affine> = { let mut acc = <G1Affine as halo2_proofs::halo2curves::CurveAffine>::CurveExt::identity();
            (0..n).map(|_| { acc = acc + g; acc.to_affine() }).collect() };
        let res = best_multiexp(&scalars, &bases).to_affine();
        let mut f = File::create(out.join(format!("ref_msm_{}.bin", log_n))).unwrap();
        f.write_all(&(n as u64).to_le_bytes()).unwrap();
        put_frs(&mut f, &scalars);
        for b in &bases { f.write_all(&raw_pt(b)).unwrap(); }
        f.write_all(&raw_pt(&res)).unwrap();
    }
    // ---- ref_fft.bin: u32 log_n | omega | input | best_fft(input, omega, log_n)
    for log_n in [10u32, 16] {
        let n = 1usize << log_n;
        let omega = Fr::ROOT_OF_UNITY.pow_vartime([1u64 << (Fr::S - log_n)]);
        let input: Vec<Fr> = (0..n).map(|_| Fr::random(&mut rng)).collect();
        let mut a = input.clone();
        best_fft(&mut a, omega, log_n);
        let mut f = File::create(out.join(format!("ref_fft_{}.bin", log_n))).unwrap();
        f.write_all(&log_n.to_le_bytes()).unwrap();
        f.write_all(&raw(&omega)).unwrap();
        put_frs(&mut f, &input);
        put_frs(&mut f, &a);
    }
    // ---- ref_ext.bin: u32 j | u32 k | u32 extended_k | coeffs | coeff_to_extended | extended_to_coeff(divide_by_vanishing_poly(ext))
    for (j, k) in [(3u32, 10u32), (5, 10)] {
        let dom = EvaluationDomain::<Fr>::new(j, k);
        let mut p = dom.empty_coeff();
        for c in p.iter_mut() { *c = Fr::random(&mut rng); }
        let coeffs: Vec<Fr> = p.to_vec();
        let ext = dom.coeff_to_extended(p);
        let ext_vals: Vec<Fr> = ext.to_vec();
        let back = dom.extended_to_coeff(dom.divide_by_vanishing_poly(ext));
        let mut f = File::create(out.join(format!("ref_ext_{}_{}.bin", j, k))).unwrap();
        for v in [j, k, dom.extended_k()] { f.write_all(&v.to_le_bytes()).unwrap(); }
        put_frs(&mut f, &coeffs);
        put_frs(&mut f, &ext_vals);
        f.write_all(&(back.len() as u64).to_le_bytes()).unwrap();
        put_frs(&mut f, &back);
    }
    // ---- ref_proof/: one seeded proof of the reference's DelayEncryptCircuit at its bench k (benches/delay_enc.rs)
    const K: u32 = 16;
    let dir = out.join("ref_proof");
    std::fs::create_dir_all(&dir).unwrap();
    let params = ParamsKZG::<Bn256>::setup(K, &mut rng);
    let mut n = BigUint::default();
    while n.bits() != 2048 { n = BigUint::from_bytes_le(&{ let mut b = vec![0u8; 256]; rng.fill_bytes(&mut b); b }); }
    let e = BigUint::from(rng.next_u64() & 31);
    let x = BigUint::from_bytes_le(&{ let mut b = vec![0u8; 256]; rng.fill_bytes(&mut b); b }) % &n;
    let circuit = DelayEncryptCircuit::<Fr, 5, 4> { n: n.clone(), e: e.clone(), x: x.clone(), spec: Spec::<Fr, 5, 4>::new(8, 57), num_input: 2,
                                                    message: vec![Fr::ZERO; 2] };
    let vk = keygen_vk(&params, &circuit).unwrap();
    let pk = keygen_pk(&params, vk, &circuit).unwrap();
    params.write_custom(&mut File::create(dir.join("params.bin")).unwrap(), SerdeFormat::RawBytes).unwrap();
    pk.write(&mut File::create(dir.join("pk.bin")).unwrap(), SerdeFormat::RawBytes).unwrap();
    pk.get_vk().write(&mut File::create(dir.join("vk.bin")).unwrap(), SerdeFormat::RawBytes).unwrap();
    // the constraint system after selector compression, as expression trees, and the query lists in their order
    let cs = pk.get_vk().cs();
    let kind = |a: &Any| match a { Any::Advice(_) => "advice", Any::Fixed => "fixed", Any::Instance => "instance" };
    let cs_json = serde_json::json!({
        "k": K, "n_fixed": cs.num_fixed_columns(), "n_advice": cs.num_advice_columns(), "n_instance": cs.num_instance_columns(),
        "degree": cs.degree(), "blinding_factors": cs.blinding_factors(),
        "gates": cs.gates().iter().flat_map(|g| g.polynomials().iter().map(expr_json)).collect::<Vec<_>>(),
        "lookups": cs.lookups().iter().map(|l| serde_json::json!([l.input_expressions().iter().map(expr_json).collect::<Vec<_>>(),
                                                                   l.table_expressions().iter().map(expr_json).collect::<Vec<_>>()])).collect::<Vec<_>>(),
        "permutation": cs.permutation().get_columns().iter().map(|c| serde_json::json!([kind(c.column_type()), c.index()])).collect::<Vec<_>>(),
        "advice_queries": cs.advice_queries().iter().map(|(c, r)| serde_json::json!([c.index(), r.0])).collect::<Vec<_>>(),
        "fixed_queries": cs.fixed_queries().iter().map(|(c, r)| serde_json::json!([c.index(), r.0])).collect::<Vec<_>>(),
        "instance_queries": cs.instance_queries().iter().map(|(c, r)| serde_json::json!([c.index(), r.0])).collect::<Vec<_>>(),
        "transcript_repr": format!("{:?}", pk.get_vk().transcript_repr()),
        "rsa": { "n": n.to_str_radix(16), "e": e.to_str_radix(16), "x": x.to_str_radix(16) },
    });
    File::create(dir.join("cs.json")).unwrap().write_all(serde_json::to_string_pretty(&cs_json).unwrap().as_bytes()).unwrap();
    // the witness as synthesize leaves it (MockProver runs the same synthesis): n_advice columns of 2^K values, unassigned = 0
    let mock = MockProver::run(K, &circuit, vec![vec![]]).unwrap();
    let mut f = File::create(dir.join("advice.bin")).unwrap();
    for col in mock.advice() { for cell in col { f.write_all(&raw(&match cell { halo2_proofs::dev::CellValue::Assigned(v) => *v, _ => Fr::ZERO })).unwrap(); } }
    // the proof, with every RNG word create_proof consumed
    let mut rec = Recording { inner: ChaCha20Rng::seed_from_u64(0x5EED), log: vec![] };
    let mut transcript = Blake2bWrite::<_, G1Affine, Challenge255<_>>::init(vec![]);
    create_proof::<KZGCommitmentScheme<Bn256>, ProverGWC<'_, Bn256>, _, _, _, _>(&params, &pk, &[circuit.clone()], &[&[&[]]], &mut rec, &mut transcript).unwrap();
    let proof = transcript.finalize();
    let mut rd = Blake2bRead::<_, G1Affine, Challenge255<_>>::init(&proof[..]);
    verify_proof::<KZGCommitmentScheme<Bn256>, VerifierGWC<'_, Bn256>, _, _, _>(&params.verifier_params(), pk.get_vk(), SingleStrategy::new(&params), &[&[&[]]], &mut rd).unwrap();
    File::create(dir.join("proof.bin")).unwrap().write_all(&proof).unwrap();
    let mut f = File::create(dir.join("rng_u64.bin")).unwrap();
    for v in &rec.log { f.write_all(&v.to_le_bytes()).unwrap(); }
    println!("wrote golden vectors to {}", out.display());
}
