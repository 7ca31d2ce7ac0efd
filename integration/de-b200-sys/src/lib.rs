//! de-b200-sys — `extern "C"` declarations for libde_b200.so, generated from include/de_b200.h (one entry per prototype).
//! NOT compiled in the backend's environment (no Rust toolchain there); the tested bindings of the same ABI are the Python
//! ctypes layer and the C++ mirror of the backend repository.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] #[derive(Clone, Copy)] pub struct de_fr { pub l: [u64; 4] }          // = halo2curves::bn256::Fr (Montgomery limbs)
#[repr(C)] #[derive(Clone, Copy)] pub struct de_fq { pub l: [u64; 4] }
#[repr(C)] #[derive(Clone, Copy)] pub struct de_g1_affine { pub x: de_fq, pub y: de_fq }        // = G1Affine, identity = zeros
#[repr(C)] #[derive(Clone, Copy)] pub struct de_g1 { pub x: de_fq, pub y: de_fq, pub z: de_fq }   // = G1 (Jacobian)
pub enum de_ctx {} pub enum de_params {} pub enum de_domain {} pub enum de_pk {} pub enum de_prover {}

#[repr(C)] #[derive(Clone, Copy)] pub struct de_value_source { pub kind: u32, pub index: u32, pub rotation: u32 }
#[repr(C)] #[derive(Clone, Copy)] pub struct de_calculation { pub op: u32, pub a: de_value_source, pub b: de_value_source,
                                                              pub horner_first: u32, pub horner_len: u32, pub target: u32 }
#[repr(C)] pub struct de_graph { pub constants: *const de_fr, pub n_constants: u32, pub rotations: *const i32, pub n_rotations: u32,
                                 pub calcs: *const de_calculation, pub n_calcs: u32, pub horner_parts: *const de_value_source,
                                 pub n_horner_parts: u32, pub n_intermediates: u32 }
#[repr(C)] pub struct de_pk_desc { pub n_fixed: u32, pub n_advice: u32, pub n_instance: u32, pub fixed_coeff: *const *const de_fr,
                                   pub n_perm_columns: u32, pub perm_column_kind: *const u32, pub perm_column_index: *const u32,
                                   pub sigma_coeff: *const *const de_fr, pub chunk_len: u32, pub blinding_factors: u32, pub delta: de_fr,
                                   pub gates: de_graph, pub n_lookups: u32, pub lookups: *const de_graph }
#[repr(C)] pub struct de_challenges { pub y: de_fr, pub beta: de_fr, pub gamma: de_fr, pub theta: de_fr,
                                      pub challenges: *const de_fr, pub n_challenges: u32 }
#[repr(C)] pub struct de_prover_desc { pub n_advice_queries: u32, pub advice_query_column: *const u32, pub advice_query_rotation: *const i32,
                                       pub n_fixed_queries: u32, pub fixed_query_column: *const u32, pub fixed_query_rotation: *const i32,
                                       pub lookup_input_graphs: *const de_graph, pub lookup_table_graphs: *const de_graph,
                                       pub transcript_repr: de_fr }

#[repr(C)] pub struct de_circuit_desc { pub kind: u32, pub k: u32, pub bits_len: u32, pub exp_bits: u32,
                                        pub n: *const u8, pub n_len: usize, pub e: *const u8, pub e_len: usize, pub x: *const u8, pub x_len: usize,
                                        pub message: *const de_fr, pub message_len: u32, pub key: [de_fr; 2], pub witness_only: u32, pub threads: u32, pub reuse_buffer: u32 }
#[repr(C)] #[derive(Clone, Copy)] pub struct de_assignment_info_t { pub k: u32, pub n_fixed: u32, pub n_advice: u32, pub n_outputs: u32,
                                                                    pub used_rows: u64, pub n_copies: u64, pub synthesis_ms: f64 }
pub enum de_assignment {}
pub const DE_CIRCUIT_MOD_POW: u32 = 0;
pub const DE_CIRCUIT_POSE_ENC: u32 = 1;
pub const DE_CIRCUIT_DELAY_ENC: u32 = 2;
pub const DE_CIRCUIT_RSA_PKCS1: u32 = 3;
pub const DE_CIRCUIT_BIGINT_SQUARE: u32 = 4; // the big-integer chip's square test (src/big_integer/chip.rs:2918-3030)
pub const DE_CIRCUIT_BIGINT_OPS: u32 = 5;    // its operator tests in one circuit (chip.rs:1479-2806)
pub const DE_CIRCUIT_POSEIDON_HASH: u32 = 6; // PoseidonHashCircuit (src/hash/chip.rs:113-236)

pub const DE_OK: c_int = 0;
pub const DE_ERR_ARG: c_int = -1;
pub const DE_ERR_CUDA: c_int = -2;
pub const DE_ERR_OOM: c_int = -3;
pub const DE_ERR_UNSUPPORTED: c_int = -4;
pub const DE_MODE_LATENCY: c_int = 0;
pub const DE_MODE_THROUGHPUT: c_int = 1;

extern "C" {
    pub fn de_ctx_create(device: c_int, out: *mut *mut de_ctx) -> c_int;
    pub fn de_ctx_destroy(ctx: *mut de_ctx) -> c_int;
    pub fn de_ctx_set_stream(ctx: *mut de_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn de_ctx_sync(ctx: *mut de_ctx) -> c_int;
    pub fn de_ctx_set_mode(ctx: *mut de_ctx, mode: c_int) -> c_int;
    pub fn de_last_error(ctx: *mut de_ctx) -> *const c_char;
    pub fn de_version() -> *const c_char;
    pub fn de_launch_count(ctx: *mut de_ctx) -> u64;
    pub fn de_timing_enable(ctx: *mut de_ctx, on: c_int) -> c_int;
    pub fn de_timing_reset(ctx: *mut de_ctx) -> c_int;
    pub fn de_timing_get(ctx: *mut de_ctx, kernel: *const c_char, total_ms: *mut f64, total_units: *mut f64, launches: *mut u64) -> c_int;
    pub fn de_fr_vec_op(ctx: *mut de_ctx, op: c_int, a: *const de_fr, b: *const de_fr, out: *mut de_fr, n: usize) -> c_int;
    pub fn de_fq_vec_op(ctx: *mut de_ctx, op: c_int, a: *const de_fq, b: *const de_fq, out: *mut de_fq, n: usize) -> c_int;
    pub fn de_msm(ctx: *mut de_ctx, scalars: *const de_fr, bases: *const de_g1_affine, n: usize, out: *mut de_g1) -> c_int;
    pub fn de_msm_dev(ctx: *mut de_ctx, d_scalars: *const de_fr, d_bases: *const de_g1_affine, n: usize, out: *mut de_g1) -> c_int;
    pub fn de_params_upload(ctx: *mut de_ctx, k: u32, g: *const de_g1_affine, g_lagrange: *const de_g1_affine, out: *mut *mut de_params) -> c_int;
    pub fn de_params_free(p: *mut de_params) -> c_int;
    pub fn de_commit(p: *mut de_params, basis: c_int, scalars: *const de_fr, n: usize, out: *mut de_g1) -> c_int;
    pub fn de_commit_batch(p: *mut de_params, basis: c_int, scalars: *const *const de_fr, n: usize, count: usize, out: *mut de_g1) -> c_int;
    pub fn de_commit_batch_dev(p: *mut de_params, basis: c_int, d_scalars: *const de_fr, stride: usize, n: usize, count: usize, out: *mut de_g1) -> c_int;
    pub fn de_ntt(ctx: *mut de_ctx, a: *mut de_fr, omega: *const de_fr, log_n: u32) -> c_int;
    pub fn de_ntt_dev(ctx: *mut de_ctx, d_a: *mut de_fr, omega: *const de_fr, log_n: u32, batch: usize, stride: usize) -> c_int;
    pub fn de_domain_create(ctx: *mut de_ctx, j: u32, k: u32, out: *mut *mut de_domain) -> c_int;
    pub fn de_domain_free(d: *mut de_domain) -> c_int;
    pub fn de_domain_info(d: *mut de_domain, extended_k: *mut u32, consts: *mut de_fr) -> c_int;
    pub fn de_coeff_to_extended(d: *mut de_domain, coeff_n: *const de_fr, ext_out: *mut de_fr) -> c_int;
    pub fn de_extended_to_coeff(d: *mut de_domain, ext_inout: *mut de_fr, out_len: *mut usize) -> c_int;
    pub fn de_lagrange_to_coeff(d: *mut de_domain, a: *mut de_fr) -> c_int;
    pub fn de_coeff_to_lagrange(d: *mut de_domain, a: *mut de_fr) -> c_int;
    pub fn de_divide_by_vanishing(d: *mut de_domain, ext_inout: *mut de_fr) -> c_int;
    pub fn de_coeff_to_extended_dev(d: *mut de_domain, d_coeff: *const de_fr, in_stride: usize, d_ext: *mut de_fr, out_stride: usize, batch: usize) -> c_int;
    pub fn de_extended_to_coeff_dev(d: *mut de_domain, d_ext: *mut de_fr, stride: usize, batch: usize, out_len: *mut usize) -> c_int;
    pub fn de_lagrange_to_coeff_dev(d: *mut de_domain, d_a: *mut de_fr, stride: usize, batch: usize) -> c_int;
    pub fn de_coeff_to_lagrange_dev(d: *mut de_domain, d_a: *mut de_fr, stride: usize, batch: usize) -> c_int;
    pub fn de_divide_by_vanishing_dev(d: *mut de_domain, d_ext: *mut de_fr, stride: usize, batch: usize) -> c_int;
    pub fn de_pk_upload(d: *mut de_domain, desc: *const de_pk_desc, out: *mut *mut de_pk) -> c_int;
    pub fn de_pk_free(pk: *mut de_pk) -> c_int;
    pub fn de_evaluate_h(pk: *mut de_pk, advice_coeff: *const *const de_fr, instance_coeff: *const *const de_fr, ch: *const de_challenges, perm_z_coeff: *const *const de_fr, lookup_coeff: *const *const de_fr, h_ext_out: *mut de_fr) -> c_int;
    pub fn de_evaluate_h_dev(pk: *mut de_pk, d_advice_coeff: *const de_fr, d_instance_coeff: *const de_fr, ch: *const de_challenges, d_perm_z_coeff: *const de_fr, d_lookup_coeff: *const de_fr, stride: usize, d_h_ext: *mut de_fr) -> c_int;
    pub fn de_pk_extend_dev(pk: *mut de_pk, d_advice_coeff: *const de_fr, d_instance_coeff: *const de_fr, d_perm_z_coeff: *const de_fr, d_lookup_coeff: *const de_fr, stride: usize) -> c_int;
    pub fn de_evaluate_h_rows_dev(pk: *mut de_pk, ch: *const de_challenges, d_h_ext: *mut de_fr) -> c_int;
    pub fn de_commit_batch_canonical_dev(p: *mut de_params, basis: c_int, d_scalars: *const de_fr, stride: usize, n: usize, count: usize, out_xy: *mut u8) -> c_int;
    pub fn de_eval_polynomial(ctx: *mut de_ctx, poly: *const de_fr, n: usize, point: *const de_fr, out: *mut de_fr) -> c_int;
    pub fn de_kate_division(ctx: *mut de_ctx, a: *const de_fr, n: usize, b: *const de_fr, q: *mut de_fr) -> c_int;
    pub fn de_prover_create(params: *mut de_params, pk: *mut de_pk, desc: *const de_prover_desc, out: *mut *mut de_prover) -> c_int;
    pub fn de_prover_free(p: *mut de_prover) -> c_int;
    pub fn de_prover_random_count(p: *mut de_prover) -> usize;
    pub fn de_prover_proof_size(p: *mut de_prover) -> usize;
    pub fn de_create_proof(p: *mut de_prover, advice: *const *const de_fr, instances: *const *const de_fr, instance_lens: *const usize, randoms: *const de_fr, n_randoms: usize, proof_out: *mut u8, proof_cap: usize, proof_len: *mut usize) -> c_int;
    pub fn de_create_proof_dev(p: *mut de_prover, d_advice: *const de_fr, advice_stride: usize, instances: *const *const de_fr, instance_lens: *const usize, d_randoms: *const de_fr, n_randoms: usize, proof_out: *mut u8, proof_cap: usize, proof_len: *mut usize) -> c_int;
    pub fn de_commit_range(p: *mut de_params, basis: c_int, scalars: *const de_fr, lo: usize, hi: usize, out_partial: *mut de_g1) -> c_int;
    pub fn de_g1_sum(ctx: *mut de_ctx, points: *const de_g1, count: usize, out: *mut de_g1) -> c_int;
    pub fn de_commit_sharded(shards: *mut *mut de_params, shard_lo: *const usize, shard_len: *const usize, n_shards: c_int, basis: c_int, scalars: *const de_fr, out: *mut de_g1) -> c_int;
    pub fn de_ntt_dist_stage1(ctx: *mut de_ctx, d_x: *const de_fr, omega: *const de_fr, log_n: u32, world: u32, rank: u32, d_z_peers: *const *mut de_fr) -> c_int;
    pub fn de_ntt_dist_stage2(ctx: *mut de_ctx, d_z: *const de_fr, omega: *const de_fr, log_n: u32, world: u32, rank: u32, d_out_peers: *const *mut de_fr) -> c_int;
    pub fn de_ntt_sharded_dev(ctxs: *const *mut de_ctx, n_gpus: c_int, d_x: *const *const de_fr, d_out: *const *mut de_fr, omega: *const de_fr, log_n: u32) -> c_int;
    pub fn de_ntt_sharded(ctxs: *const *mut de_ctx, n_gpus: c_int, a: *mut de_fr, omega: *const de_fr, log_n: u32) -> c_int;
    pub fn de_dev_alloc(ctx: *mut de_ctx, bytes: usize, d_ptr: *mut *mut c_void) -> c_int;
    pub fn de_dev_free(ctx: *mut de_ctx, d_ptr: *mut c_void) -> c_int;
    pub fn de_dev_copy(ctx: *mut de_ctx, d_dst: *mut c_void, d_src: *const c_void, bytes: usize) -> c_int;
    pub fn de_ipc_export(ctx: *mut de_ctx, d_ptr: *mut c_void, handle: *mut u8) -> c_int;
    pub fn de_ipc_import(ctx: *mut de_ctx, handle: *const u8, d_ptr: *mut *mut c_void) -> c_int;
    pub fn de_ipc_release(ctx: *mut de_ctx, d_ptr: *mut c_void) -> c_int;
    pub fn de_g1_mul_base_dev(ctx: *mut de_ctx, base: *const de_g1_affine, d_scalars: *const de_fr, n: usize, d_out: *mut de_g1_affine) -> c_int;
    pub fn de_g1_batch_normalize(ctx: *mut de_ctx, points: *const de_g1, count: usize, out: *mut de_g1_affine) -> c_int;
    pub fn de_int_peak(ctx: *mut de_ctx, gmul_per_s: *mut f64) -> c_int;
    pub fn de_int_peak_sqr(ctx: *mut de_ctx, gsqr_per_s: *mut f64) -> c_int;
    pub fn de_ntt_dist_run(ctx: *mut de_ctx, d_x: *const de_fr, omega: *const de_fr, log_n: u32, world: u32, rank: u32, d_z_peers: *const *mut de_fr, d_out_peers: *const *mut de_fr, d_flag_peers: *const *mut u32, epoch: u32, chunks: u32) -> c_int;
    pub fn de_ntt_dist_error(ctx: *mut de_ctx, timed_out: *mut c_int) -> c_int;
    pub fn de_ntt_dist_prepare(ctx: *mut de_ctx, omega: *const de_fr, log_n: u32, world: u32, rank: u32) -> c_int;
    // circuit front-end (host only)
    pub fn de_circuit_synthesize(desc: *const de_circuit_desc, out: *mut *mut de_assignment) -> c_int;
    pub fn de_circuit_witness(desc: *const de_circuit_desc, advice_out: *mut de_fr, info: *mut de_assignment_info_t) -> c_int;
    pub fn de_assignment_free(a: *mut de_assignment);
    pub fn de_assignment_info(a: *const de_assignment, info: *mut de_assignment_info_t) -> c_int;
    pub fn de_assignment_fixed(a: *const de_assignment, column: u32, out: *mut de_fr) -> c_int;
    pub fn de_assignment_advice(a: *const de_assignment, column: u32, out: *mut de_fr) -> c_int;
    pub fn de_assignment_copies(a: *const de_assignment, out: *mut u32) -> c_int;
    pub fn de_assignment_outputs(a: *const de_assignment, out: *mut de_fr) -> c_int;
    pub fn de_assignment_sigma(a: *const de_assignment, omega: *const de_fr, delta: *const de_fr, n_columns: u32, out: *mut de_fr) -> c_int;
    pub fn de_frontend_last_error() -> *const c_char;
    pub fn de_poseidon_permute(t: u32, r_f: u32, r_p: u32, state: *mut de_fr) -> c_int;
    pub fn de_poseidon_cipher(decrypt: c_int, key: *const de_fr, input: *const de_fr, n_in: u32, out: *mut de_fr) -> c_int;
}
