/*
 * de_b200.h — C ABI of libde_b200.so, the B200 (sm_100a) backend for the proving hot path under the halo2 circuits
 * of radiusxyz/delay-encryption-in-halo2.
 *
 * The reference has no FFI today: the path is reached through plain Rust generics of the un-vendored dependency
 * halo2_proofs (tag v2023_04_20, /root/reference/Cargo.toml:17) from create_proof / keygen_vk / keygen_pk
 * (/root/reference/benches/delay_enc.rs:86,103,123; benches/mod_pow.rs:163,180,201; benches/pose_enc.rs:89,106,127).
 * Each entry point below names the halo2_proofs function a patched crate would forward to it (INTEGRATION.md shows
 * the Rust `extern "C"` block and the [patch] wiring).
 *
 * Data conventions (zero-copy with halo2curves' in-memory layout):
 *   de_fr / de_fq   4 x u64 little-endian limbs, MONTGOMERY form (R = 2^256)       = halo2curves::bn256::{Fr,Fq}
 *   de_g1_affine    {x, y}, identity = all zero                                      = halo2curves::bn256::G1Affine
 *   de_g1           {x, y, z} Jacobian, identity z = 0; any valid representative     = halo2curves::bn256::G1
 * Every function returns 0 on success or a negative de_status; de_last_error() gives the message.  There is NO CPU
 * fallback: without a CUDA device every compute entry point fails with DE_ERR_CUDA.
 * Pointers named d_* are device pointers (e.g. torch tensors' data_ptr()); all others are host pointers that are
 * not retained past the call.  A context is bound to one device and one stream and is not thread-safe; use one
 * context per host thread / GPU (proof-batch sharding uses one per GPU).
 */
#ifndef DE_B200_H
#define DE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[4]; } de_fr;
typedef struct { uint64_t l[4]; } de_fq;
typedef struct { de_fq x, y; } de_g1_affine;
typedef struct { de_fq x, y, z; } de_g1;

typedef struct de_ctx de_ctx;
typedef struct de_params de_params;
typedef struct de_domain de_domain;
typedef struct de_pk de_pk;

enum de_status {
    DE_OK = 0,
    DE_ERR_ARG = -1,     /* bad argument (the Rust originals would panic on an assert) */
    DE_ERR_CUDA = -2,    /* CUDA runtime error, including "no device" */
    DE_ERR_OOM = -3,     /* device or pinned-host allocation failed */
    DE_ERR_UNSUPPORTED = -4
};

/* ---- context --------------------------------------------------------------------------------------------- */
int de_ctx_create(int device, de_ctx** out);
int de_ctx_destroy(de_ctx* ctx);
/* bind an existing CUDA stream (cudaStream_t as void*); NULL restores the context's own stream */
int de_ctx_set_stream(de_ctx* ctx, void* cuda_stream);
int de_ctx_sync(de_ctx* ctx);
/* Scheduling hint.  DE_MODE_LATENCY (default): one proof / one MSM at a time owns the GPU, so latency-bound tails are given
 * short dependency chains even at the price of idle lanes (tree-shaped bucket reduction), and a commitment's launch sequence
 * is replayed as a CUDA graph from the third call with the same buffers on.  DE_MODE_THROUGHPUT: several contexts run
 * concurrently on the same GPU (batches of proofs), so kernels keep every lane busy and leave the overlap to the other
 * streams (serial segmented sums, eager launches).  Results are identical; measured at k = 16 (round 2): 8.1 ms per proof
 * alone in latency mode, 162 proofs/s with 8 proofs in flight in throughput mode. */
enum de_mode { DE_MODE_LATENCY = 0, DE_MODE_THROUGHPUT = 1 };
int de_ctx_set_mode(de_ctx* ctx, int mode);
const char* de_last_error(de_ctx* ctx); /* ctx may be NULL: returns the last error of a failed de_ctx_create */
const char* de_version(void);
/* number of kernels this library launched on the context since creation (bench.py's gpu_launches); kernels inside a replayed
 * graph are counted at every replay */
uint64_t de_launch_count(de_ctx* ctx);
/* Roofline denominator, measured on the device the context is bound to ("the build must measure it on
 * the box"): Fr Montgomery multiplications per second of dependent-product chains at full occupancy (16 warps per SM, two
 * independent chains per thread) - the 136 IMAD.WIDE.U32 carry-chain multiplier every hot kernel here is made of.  Best of four
 * timed launches after two warm-up launches, ~10 ms in total; synchronises the context's stream. */
int de_int_peak(de_ctx* ctx, double* gmul_per_s);
/* The same measurement for the dedicated squaring (field.cuh sqr: 36 + 72 products): Fr squarings per second.  bench.py weighs
 * the 2 squarings of a mixed bucket addition by mul_peak / sqr_peak when it states executed work in multiplications. */
int de_int_peak_sqr(de_ctx* ctx, double* gsqr_per_s);

/* per-kernel device timing: CUDA events recorded around the named kernels on the context's stream.
 * Known names: "k_msm_accumulate", "k_msm_digit_sums", "k_ntt_pass", "k_eval_h".  units = points / elements / rows. */
int de_timing_enable(de_ctx* ctx, int on);
int de_timing_reset(de_ctx* ctx);
int de_timing_get(de_ctx* ctx, const char* kernel, double* total_ms, double* total_units, uint64_t* launches);

/* ---- a1: halo2curves Fr / Fq element-wise arithmetic (parity surface for the field kernels) ---------------- */
enum de_field_op { DE_OP_MUL = 0, DE_OP_ADD = 1, DE_OP_SUB = 2, DE_OP_FROM_MONT = 3, DE_OP_TO_MONT = 4,
                   DE_OP_INV = 5 /* ff::Field::invert; 0 -> 0 */, DE_OP_SQR = 6 /* ff::Field::square: the dedicated squaring of field.cuh */ };
int de_fr_vec_op(de_ctx* ctx, int op, const de_fr* a, const de_fr* b, de_fr* out, size_t n);
int de_fq_vec_op(de_ctx* ctx, int op, const de_fq* a, const de_fq* b, de_fq* out, size_t n);

/* ---- a3: halo2_proofs::arithmetic::best_multiexp(coeffs, bases) -> G1 ------------------------------------- */
int de_msm(de_ctx* ctx, const de_fr* scalars, const de_g1_affine* bases, size_t n, de_g1* out);
/* same with device-resident inputs; result written to host */
int de_msm_dev(de_ctx* ctx, const de_fr* d_scalars, const de_g1_affine* d_bases, size_t n, de_g1* out);

/* ---- a8: halo2_proofs::poly::kzg::commitment::ParamsKZG --------------------------------------------------- */
/* Stage g[0..2^k) and g_lagrange[0..2^k) in HBM once per ParamsKZG (either may be NULL).  Window tables
 * 2^(c*t) * base are precomputed on the device so every later commit is a single-bucket-set Pippenger. */
int de_params_upload(de_ctx* ctx, uint32_t k, const de_g1_affine* g, const de_g1_affine* g_lagrange, de_params** out);
int de_params_free(de_params* p);
/* ParamsKZG::commit (basis 0, uses g[..n]) / ParamsKZG::commit_lagrange (basis 1, uses g_lagrange[..n]); the Blind
 * argument of the Rust API is unused for KZG and has no counterpart here. */
int de_commit(de_params* p, int basis, const de_fr* scalars, size_t n, de_g1* out);
/* `count` polynomials of n scalars each in one launch sequence (the 5/10/5 same-round commitments of create_proof) */
int de_commit_batch(de_params* p, int basis, const de_fr* const* scalars, size_t n, size_t count, de_g1* out);
/* device-resident: d_scalars holds count polynomials, `stride` elements apart */
int de_commit_batch_dev(de_params* p, int basis, const de_fr* d_scalars, size_t stride, size_t n, size_t count, de_g1* out);

/* ---- a4: halo2_proofs::arithmetic::best_fft(a, omega, log_n) for Scalar = Fr ------------------------------ */
int de_ntt(de_ctx* ctx, de_fr* a, const de_fr* omega, uint32_t log_n);
/* device-resident, `batch` vectors `stride` elements apart, in place */
int de_ntt_dev(de_ctx* ctx, de_fr* d_a, const de_fr* omega, uint32_t log_n, size_t batch, size_t stride);

/* ---- a5-a7: halo2_proofs::poly::EvaluationDomain ---------------------------------------------------------- */
int de_domain_create(de_ctx* ctx, uint32_t j /* cs.degree() */, uint32_t k, de_domain** out); /* EvaluationDomain::new */
int de_domain_free(de_domain* d);
/* consts = {omega, omega_inv, extended_omega, extended_omega_inv} */
int de_domain_info(de_domain* d, uint32_t* extended_k, de_fr consts[4]);
int de_coeff_to_extended(de_domain* d, const de_fr* coeff_n, de_fr* ext_out);     /* coeff_to_extended */
int de_extended_to_coeff(de_domain* d, de_fr* ext_inout, size_t* out_len);         /* extended_to_coeff: first out_len entries */
int de_lagrange_to_coeff(de_domain* d, de_fr* a);                                  /* lagrange_to_coeff */
int de_coeff_to_lagrange(de_domain* d, de_fr* a);                                  /* coeff_to_lagrange */
int de_divide_by_vanishing(de_domain* d, de_fr* ext_inout);                        /* divide_by_vanishing_poly */
/* device-resident, batched variants (vectors `stride` elements apart) */
int de_coeff_to_extended_dev(de_domain* d, const de_fr* d_coeff, size_t in_stride, de_fr* d_ext, size_t out_stride, size_t batch);
int de_extended_to_coeff_dev(de_domain* d, de_fr* d_ext, size_t stride, size_t batch, size_t* out_len);
int de_lagrange_to_coeff_dev(de_domain* d, de_fr* d_a, size_t stride, size_t batch);
int de_coeff_to_lagrange_dev(de_domain* d, de_fr* d_a, size_t stride, size_t batch);
int de_divide_by_vanishing_dev(de_domain* d, de_fr* d_ext, size_t stride, size_t batch);

/* ---- a9: halo2_proofs::plonk::evaluation::Evaluator::evaluate_h ------------------------------------------- */
/* A compiled GraphEvaluator (constants / rotations / calculations) plus the permutation and lookup arguments'
 * shapes; mirrors the structures evaluate_h walks. */
enum de_calc_op { DE_CALC_ADD = 0, DE_CALC_SUB = 1, DE_CALC_MUL = 2, DE_CALC_SQUARE = 3, DE_CALC_DOUBLE = 4,
                  DE_CALC_NEGATE = 5, DE_CALC_HORNER = 6, DE_CALC_STORE = 7 };
enum de_value_kind { DE_VAL_CONSTANT = 0, DE_VAL_INTERMEDIATE = 1, DE_VAL_FIXED = 2, DE_VAL_ADVICE = 3,
                     DE_VAL_INSTANCE = 4, DE_VAL_CHALLENGE = 5, DE_VAL_BETA = 6, DE_VAL_GAMMA = 7, DE_VAL_THETA = 8,
                     DE_VAL_Y = 9, DE_VAL_PREVIOUS = 10 };
typedef struct { uint32_t kind; uint32_t index; uint32_t rotation; /* index into rotations[] */ } de_value_source;
typedef struct {
    uint32_t op;            /* de_calc_op */
    de_value_source a, b;   /* operands (b unused for unary ops; HORNER: a = start value, b = multiplier) */
    uint32_t horner_first;  /* HORNER: parts are horner_parts[horner_first .. horner_first + horner_len) */
    uint32_t horner_len;
    uint32_t target;        /* intermediate slot written */
} de_calculation;
typedef struct {
    const de_fr* constants; uint32_t n_constants;
    const int32_t* rotations; uint32_t n_rotations;
    const de_calculation* calcs; uint32_t n_calcs;
    const de_value_source* horner_parts; uint32_t n_horner_parts;
    uint32_t n_intermediates;  /* result of the graph = intermediate written by the last calculation (zero if none) */
} de_graph;
typedef struct {
    uint32_t n_fixed, n_advice, n_instance;
    const de_fr* const* fixed_coeff;   /* fixed column polynomials, coefficient form, n each */
    /* permutation argument */
    uint32_t n_perm_columns;           /* columns in the permutation, in order */
    const uint32_t* perm_column_kind;  /* de_value_kind (FIXED / ADVICE / INSTANCE) per column */
    const uint32_t* perm_column_index;
    const de_fr* const* sigma_coeff;   /* permutation polynomials, coefficient form */
    uint32_t chunk_len;                /* cs.degree() - 2 */
    uint32_t blinding_factors;         /* cs.blinding_factors() */
    de_fr delta;                       /* Fr::DELTA */
    /* custom gates: value = value * y + gate, in order */
    de_graph gates;
    /* lookups */
    /* lookups: one compiled graph each, whose result is (compressed_input + beta) * (compressed_table + gamma) as
     * Evaluator::new builds it */
    uint32_t n_lookups;
    const de_graph* lookups;
} de_pk_desc;
typedef struct { de_fr y, beta, gamma, theta; const de_fr* challenges; uint32_t n_challenges; } de_challenges;

/* ProvingKey staging: fixed / sigma / l0 / l_last / l_active cosets are computed on the device once and stay in HBM */
int de_pk_upload(de_domain* d, const de_pk_desc* desc, de_pk** out);
int de_pk_free(de_pk* pk);
/* advice / instance polynomials in coefficient form (n each); permutation z polys (ceil(n_perm_columns/chunk_len)),
 * lookup polys in block order {all product z | all permuted inputs a' | all permuted tables s'} (3 * n_lookups entries;
 * the order create_proof commits them in): all coefficient form, n each.
 * Output: h evaluations over the extended domain, BEFORE divide_by_vanishing_poly (as evaluate_h returns them). */
int de_evaluate_h(de_pk* pk, const de_fr* const* advice_coeff, const de_fr* const* instance_coeff,
                  const de_challenges* ch, const de_fr* const* perm_z_coeff, const de_fr* const* lookup_coeff,
                  de_fr* h_ext_out);
/* device-resident: each d_* block holds its polynomials `stride` elements apart (coefficient form, n valid entries);
 * the extended cosets are produced in the pk's HBM workspace by one batched coeff_to_extended, then one fused kernel
 * walks the extended domain. */
int de_evaluate_h_dev(de_pk* pk, const de_fr* d_advice_coeff, const de_fr* d_instance_coeff, const de_challenges* ch,
                      const de_fr* d_perm_z_coeff, const de_fr* d_lookup_coeff, size_t stride, de_fr* d_h_ext);

/* the two halves of de_evaluate_h_dev, so that a prover can run the coset transforms on a second stream while the same
 * round's commitments are computed (they do not depend on any challenge), and only the row kernel after y is known */
int de_pk_extend_dev(de_pk* pk, const de_fr* d_advice_coeff, const de_fr* d_instance_coeff, const de_fr* d_perm_z_coeff,
                     const de_fr* d_lookup_coeff, size_t stride);
int de_evaluate_h_rows_dev(de_pk* pk, const de_challenges* ch, de_fr* d_h_ext);

/* ---- a8 (transcript form): commitments as canonical affine coordinates ------------------------------------ */
/* like de_commit_batch_dev, but each result is written as 64 bytes: x || y, little-endian CANONICAL integers (what
 * Fq::to_repr returns and the Blake2b transcript hashes); the identity is 64 zero bytes */
int de_commit_batch_canonical_dev(de_params* p, int basis, const de_fr* d_scalars, size_t stride, size_t n, size_t count, uint8_t* out_xy);

/* ---- section 8f row 4: halo2_proofs::arithmetic::{eval_polynomial, kate_division} -------------------------- */
int de_eval_polynomial(de_ctx* ctx, const de_fr* poly, size_t n, const de_fr* point, de_fr* out);
/* q = a / (X - b), n - 1 coefficients written; n >= 2 */
int de_kate_division(de_ctx* ctx, const de_fr* a, size_t n, const de_fr* b, de_fr* q);

/* ---- section 8f rows 1-2: halo2_proofs::plonk::create_proof (KZG, ProverGWC, Blake2bWrite / Challenge255) --- */
typedef struct de_prover de_prover;
typedef struct {
    /* cs.advice_queries / cs.fixed_queries in the ConstraintSystem's order: (column index, rotation) */
    uint32_t n_advice_queries; const uint32_t* advice_query_column; const int32_t* advice_query_rotation;
    uint32_t n_fixed_queries; const uint32_t* fixed_query_column; const int32_t* fixed_query_rotation;
    /* per lookup of the pk: compiled graphs of the theta-compressed input / table expressions
     * (Horner over theta of the argument's expressions; evaluated on the n rows of the lagrange domain) */
    const de_graph* lookup_input_graphs;
    const de_graph* lookup_table_graphs;
    de_fr transcript_repr;   /* vk.transcript_repr */
} de_prover_desc;
/* params and pk must live on the same context.  Allocates every per-proof buffer once. */
int de_prover_create(de_params* params, de_pk* pk, const de_prover_desc* desc, de_prover** out);
int de_prover_free(de_prover* p);
/* number of Fr::random(rng) draws one proof consumes, and the proof size in bytes */
size_t de_prover_random_count(de_prover* p);
size_t de_prover_proof_size(de_prover* p);
/* advice: n_advice columns of n values as synthesize leaves them (lagrange form; the last blinding_factors + 1 rows are
 * overwritten with blinding values); instances: the public inputs per instance column; randoms: the draws of
 * Fr::random(rng) in create_proof's order - the library never touches an RNG.
 * Writes the proof (the bytes Blake2bWrite::finalize returns). */
int de_create_proof(de_prover* p, const de_fr* const* advice, const de_fr* const* instances, const size_t* instance_lens,
                    const de_fr* randoms, size_t n_randoms, uint8_t* proof_out, size_t proof_cap, size_t* proof_len);
/* the same with the advice columns (advice_stride elements apart) and the random draws already resident in HBM; the
 * public inputs stay host pointers (they are hashed into the transcript on the host) */
int de_create_proof_dev(de_prover* p, const de_fr* d_advice, size_t advice_stride, const de_fr* const* instances,
                        const size_t* instance_lens, const de_fr* d_randoms, size_t n_randoms, uint8_t* proof_out,
                        size_t proof_cap, size_t* proof_len);

/* ---- multi-GPU: MSM base-range sharding -------------------------------------------- */
/* shard s of n_shards: commits scalars[lo..hi) against the matching base range of p and returns the partial sum;
 * the host (or de_g1_sum) adds the n_shards partial points. */
int de_commit_range(de_params* p, int basis, const de_fr* scalars, size_t lo, size_t hi, de_g1* out_partial);
int de_g1_sum(de_ctx* ctx, const de_g1* points, size_t count, de_g1* out);
/* the same inside ONE process (a Rust prover driving several GPUs): shards[i] = ParamsKZG staged on its own context / GPU
 * from the base slice [shard_lo[i], shard_lo[i] + shard_len[i]) of the SRS (slice padded to a power of two with identity
 * points); one host thread per shard commits its slice of `scalars` (host, full length) concurrently and the partial
 * points are added on shards[0]'s GPU.  Multi-process sharding uses de_commit + an all-gather instead (de_b200/sharding.py). */
int de_commit_sharded(de_params* const* shards, const size_t* shard_lo, const size_t* shard_len, int n_shards, int basis,
                      const de_fr* scalars, de_g1* out);
/* ---- multi-GPU: best_fft of ONE vector spread over W = 1, 2, 4 or 8 GPUs (NTT row) ----------- */
/* Four-step transform N = W * M with both exchanges done as peer-memory stores (NVLink) from inside the kernels, no copy pass
 * and no library collective:
 *   input  (cyclic):  rank r holds x_r[t] = a[r + W t], t < M;
 *   output (blocks):  rank q holds A[q M .. (q + 1) M), natural order -  A = best_fft(a, omega, log_n).
 * stage 1 on rank r: local M-point transform whose last pass stores column j into row r of the exchange buffer z (M elements) of
 * rank j / (M / W);  stage 2 on rank q: row i1 of its z times omega^(i1 j), W-point transform down the columns, output j1 stored
 * into rank j1's output block.  d_z_peers / d_out_peers hold the W ranks' buffers as pointers valid on THIS device
 * (peer access inside one process, de_ipc_import across processes).  The caller orders the stages: every rank's stage 1 must
 * have completed before any stage 2 starts, and every stage 2 before the outputs are read or the next stage 1 begins (stream
 * events inside one process - de_ntt_sharded_dev does it -, a barrier on the stream across processes).  11 + log2 W <= log_n <= 28. */
int de_ntt_dist_stage1(de_ctx* ctx, const de_fr* d_x, const de_fr* omega, uint32_t log_n, uint32_t world, uint32_t rank,
                       de_fr* const* d_z_peers);
int de_ntt_dist_stage2(de_ctx* ctx, const de_fr* d_z, const de_fr* omega, uint32_t log_n, uint32_t world, uint32_t rank,
                       de_fr* const* d_out_peers);
/* One whole call on THIS rank with the exchange PIPELINED and the ranks ordered by flags in peer memory instead of stream
 * barriers (one process per GPU): the peer-store pass runs as up to `chunks` (1 .. 8) ranges of CTAs, each followed by a flag
 * store to every rank; the cross stage of range k starts on a second, high-priority stream of the context as soon as every
 * rank has signalled range k, i.e. it runs (NVLink-bound) under the pass of range k + 1 (multiply-bound).
 * ONE rank per device (a rank's stream parked on an event can hold the hardware queue another rank's kernels of the same device
 * sit in; the spinning wait would then give up) - several ranks on one device belong to de_ntt_sharded_dev, which orders them with
 * events.  d_flag_peers[r] = rank r's flag words (de_dev_alloc of DE_NTT_DIST_FLAG_BYTES, zeroed once, mapped on every rank);
 * epoch = 1, 2, 3, ... per call, identical on all ranks.  On return (stream order on the context's stream) this rank's
 * output block is complete and every rank has finished reading its exchange buffer.  A rank that never arrives makes the
 * waits give up after ~2 s instead of hanging the device; de_ntt_dist_error reports (and clears) that. */
#define DE_NTT_DIST_FLAG_BYTES 288
int de_ntt_dist_run(de_ctx* ctx, const de_fr* d_x, const de_fr* omega, uint32_t log_n, uint32_t world, uint32_t rank,
                    de_fr* const* d_z_peers, de_fr* const* d_out_peers, uint32_t* const* d_flag_peers, uint32_t epoch, uint32_t chunks);
int de_ntt_dist_error(de_ctx* ctx, int* timed_out);
/* everything of de_ntt_dist_run that allocates or synchronises, ahead of time (optional with one rank per device; REQUIRED for
 * all ranks before the first run when one thread drives several ranks of the same device) */
int de_ntt_dist_prepare(de_ctx* ctx, const de_fr* omega, uint32_t log_n, uint32_t world, uint32_t rank);
/* The same inside ONE process: ctxs[r] is rank r (normally one context per GPU; several contexts on one GPU also work), d_x[r] /
 * d_out[r] its input / output block (d_out[r] may equal d_x[r]); the exchange is pipelined in the same way, ordered by CUDA events
 * (cross stage of range k on every rank's second stream after all ranks' range k).  Asynchronous: the result is complete in stream order on every
 * context's stream; exchange buffers live in the contexts' workspaces.  Two exceptions to "asynchronous": the FIRST call for a
 * given (context, log_n, omega) builds the plan's twiddle tables (device allocations and one stream synchronisation); and when
 * a launch fails part-way the call waits for every participating stream before it returns the error, so that no peer is still
 * storing into a buffer the caller may now free. */
int de_ntt_sharded_dev(de_ctx* const* ctxs, int n_gpus, const de_fr* const* d_x, de_fr* const* d_out, const de_fr* omega,
                       uint32_t log_n);
/* best_fft(a, omega, log_n) for a HOST vector, natural order in and out, over the GPUs of ctxs (distinct contexts): block r
 * travels over GPU r's own PCIe link (one host thread per GPU), is dealt to the cyclic slices by peer stores, transformed as
 * above and read back.  Synchronous. */
int de_ntt_sharded(de_ctx* const* ctxs, int n_gpus, de_fr* a, const de_fr* omega, uint32_t log_n);
/* plain device allocations that another process can map (cudaMalloc + CUDA IPC; 64-byte handles) */
int de_dev_alloc(de_ctx* ctx, size_t bytes, void** d_ptr);
int de_dev_free(de_ctx* ctx, void* d_ptr);
/* device-to-device copy on the context's stream (fills / reads de_dev_alloc buffers from other device memory) */
int de_dev_copy(de_ctx* ctx, void* d_dst, const void* d_src, size_t bytes);
int de_ipc_export(de_ctx* ctx, void* d_ptr, uint8_t handle[64]);
int de_ipc_import(de_ctx* ctx, const uint8_t handle[64], void** d_ptr);
int de_ipc_release(de_ctx* ctx, void* d_ptr);

/* d_out[i] = [d_scalars[i]] base (affine, Montgomery), device-resident: the fixed-base multiplications of
 * ParamsKZG::setup (g[i] = [s^i] G) and the generator of synthetic SRS bases for the size sweeps */
int de_g1_mul_base_dev(de_ctx* ctx, const de_g1_affine* base, const de_fr* d_scalars, size_t n, de_g1_affine* d_out);
/* group::Curve::batch_normalize: Jacobian -> affine (identity -> all zero).  The Jacobian representative an MSM returns
 * depends on the (atomic) accumulation order; the affine point does not. */
int de_g1_batch_normalize(de_ctx* ctx, const de_g1* points, size_t count, de_g1_affine* out);

/* ---- section 8f row 3: circuit front-end (witness generation) ---------------------------------------------------- */
/* The reference's chips (/root/reference/src/big_integer, src/rsa, src/poseidon/chip.rs, src/hash, src/encryption) and its
 * three bench circuits as HOST witness generators for the MainGate + RangeChip constraint-system shape this library proves
 * (de_b200/plonk.py: main_gate_shape): Circuit::synthesize of
 *   DE_CIRCUIT_MOD_POW    benches/mod_pow.rs:36-140        RSACircuit (x^e mod n, variable exp_bits-bit e)
 *   DE_CIRCUIT_POSE_ENC   src/encryption/chip.rs:114-198   PoseidonEncCircuit (key[2], message)
 *   DE_CIRCUIT_DELAY_ENC  src/lib.rs:103-318               DelayEncryptCircuit (n, e, x, message)
 *   DE_CIRCUIT_RSA_PKCS1  src/rsa/chip.rs:119-212          verify_pkcs1v15_signature (n, e fixed, x = signature, message =
 *                                                          SHA-256 digest as four 64-bit limbs); output = the accept bit
 *   DE_CIRCUIT_BIGINT_SQUARE src/big_integer/chip.rs:2918-3030  the big-integer chip's square test (n = a, x = expected a * a as one
 *                                                          integer); outputs = the is_equal_muled bit, then the 2 n1 - 1 Muled limbs
 *   DE_CIRCUIT_BIGINT_OPS    src/big_integer/chip.rs:1479-2806  the chip's operator tests in one circuit (x = a, e = b, n = modulus,
 *                                                          a, b < n): outputs = (limb count, limbs...) records of add, sub, sub's
 *                                                          overflow bit, mul_mod, pow_mod and pow_mod_fixed_exp with the low
 *                                                          exp_bits bits of b, is_equal_fresh, is_less_than, is_less_than_or_equal,
 *                                                          is_less_than(a, n)
 *   DE_CIRCUIT_POSEIDON_HASH src/hash/chip.rs:113-236           PoseidonHashCircuit (message = the inputs, any count): HasherChip's
 *                                                          digest constrained equal to the native sponge's; outputs = the 5 state words
 * One pass yields the fixed columns, the advice columns and the copy constraints; keygen uses fixed + copies, create_proof
 * the advice columns.  No context and no GPU: these run wherever the library loads.  Errors: de_frontend_last_error(). */
enum de_circuit_kind { DE_CIRCUIT_MOD_POW = 0, DE_CIRCUIT_POSE_ENC = 1, DE_CIRCUIT_DELAY_ENC = 2, DE_CIRCUIT_RSA_PKCS1 = 3,
                       DE_CIRCUIT_BIGINT_SQUARE = 4, DE_CIRCUIT_BIGINT_OPS = 5, DE_CIRCUIT_POSEIDON_HASH = 6 };
typedef struct {
    uint32_t kind;       /* de_circuit_kind */
    uint32_t k;          /* 2^k rows */
    uint32_t bits_len;   /* BITS_LEN (2048) */
    uint32_t exp_bits;   /* EXP_LIMB_BITS (5) */
    const uint8_t* n; size_t n_len;   /* little-endian bytes */
    const uint8_t* e; size_t e_len;
    const uint8_t* x; size_t x_len;
    const de_fr* message; uint32_t message_len;   /* Montgomery field elements */
    de_fr key[2];                                  /* DE_CIRCUIT_POSE_ENC only */
    uint32_t witness_only;  /* 1: the pass create_proof makes - advice columns only, no fixed columns / copy constraints */
    uint32_t threads;       /* witness-only passes: host threads that emit the independent row ranges of the RSA region (the
                               mul_mod calls of pow_mod); 0 or 1 = the calling thread alone.  The rows do not depend on it. */
    uint32_t reuse_buffer;  /* de_circuit_witness only: 1 = advice_out already holds the result of an earlier witness pass for the
                               same proving key and is not zeroed first - a pass writes the same cells whatever the witness, so a
                               prover that cycles its staging buffers saves the 10 MB memset.  (The reference's layout depends on
                               one value: x^e mod n is assigned as a constant, limb by limb up to its top non-zero limb - a
                               different limb count is a different circuit, with its own keys.) */
} de_circuit_desc;
typedef struct {
    uint32_t k, n_fixed, n_advice, n_outputs;
    uint64_t used_rows, n_copies;
    double synthesis_ms;
} de_assignment_info_t;
typedef struct de_assignment de_assignment;
int de_circuit_synthesize(const de_circuit_desc* desc, de_assignment** out);
/* the pass create_proof makes, straight into the caller's buffer: advice_out receives the 5 advice columns, 2^k Montgomery
 * elements each, column after column (e.g. the pinned staging buffer de_create_proof uploads from); info may be NULL */
int de_circuit_witness(const de_circuit_desc* desc, de_fr* advice_out, de_assignment_info_t* info);
void de_assignment_free(de_assignment* a);
int de_assignment_info(const de_assignment* a, de_assignment_info_t* info);
/* column `column` as 2^k Montgomery field elements */
int de_assignment_fixed(const de_assignment* a, uint32_t column, de_fr* out);
int de_assignment_advice(const de_assignment* a, uint32_t column, de_fr* out);
/* n_copies x (left column, left row, right column, right row); columns are positions in the permutation: 0-4 advice, 5 instance */
int de_assignment_copies(const de_assignment* a, uint32_t* out);
/* the circuit's results: mod_pow / delay_enc: the limbs of x^e mod n (delay_enc: then the two key words and the three
 * ciphertext words); pose_enc: the ciphertext; rsa_pkcs1: the accept bit */
int de_assignment_outputs(const de_assignment* a, de_fr* out);
/* permutation::keygen::Assembly::build_pk's sigma columns from the copy constraints: n_columns x 2^k values, lagrange form */
int de_assignment_sigma(const de_assignment* a, const de_fr* omega, const de_fr* delta, uint32_t n_columns, de_fr* out);
const char* de_frontend_last_error(void);
/* native Poseidon (src/poseidon/permutation.rs Spec::permute with Spec::new(r_f, r_p), width t) and the duplex cipher
 * (src/encryption/poseidon_enc.rs PoseidonCipher::{encrypt, decrypt}; decrypt returns DE_ERR_UNSUPPORTED on a bad tag) */
int de_poseidon_permute(uint32_t t, uint32_t r_f, uint32_t r_p, de_fr* state);
int de_poseidon_cipher(int decrypt, const de_fr key[2], const de_fr* in, uint32_t n_in, de_fr* out);

#ifdef __cplusplus
}
#endif
#endif /* DE_B200_H */
