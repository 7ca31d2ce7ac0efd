// halo2_b200.hpp — header-only C++ host-side mirror of the halo2_proofs surface served by libde_b200.so.
// The reference's host language is Rust, which this environment cannot compile; this mirror keeps the reference's names,
// argument meaning and failure behaviour (the Rust functions panic on violated asserts; these throw std::runtime_error).
//   halo2_proofs::arithmetic::{best_multiexp, best_fft, eval_polynomial, kate_division}, poly::EvaluationDomain,
//   poly::kzg::commitment::ParamsKZG, plonk::ProvingKey (device staging), plonk::create_proof (Prover);
//   the reference's own circuits as witness generators: DelayEncryptCircuit, RSACircuit, PoseidonEncCircuit (Circuit::synthesize)
// Element types are the C ABI's (include/de_b200.h): Montgomery limbs, byte-identical to halo2curves.
#pragma once
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/de_b200.h"

namespace halo2_b200 {

class Context {
public:
    explicit Context(int device = 0) {
        int rc = de_ctx_create(device, &h_);
        if (rc != DE_OK) throw std::runtime_error(std::string("de_ctx_create: ") + de_last_error(nullptr));
    }
    ~Context() { de_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    de_ctx* handle() const { return h_; }
    void check(int rc, const char* what) const {
        if (rc != DE_OK) throw std::runtime_error(std::string(what) + ": " + de_last_error(h_));
    }
    std::vector<de_g1_affine> batch_normalize(const std::vector<de_g1>& pts) const {  // group::Curve::batch_normalize
        std::vector<de_g1_affine> out(pts.size());
        check(de_g1_batch_normalize(h_, pts.data(), pts.size(), out.data()), "batch_normalize");
        return out;
    }

private:
    de_ctx* h_ = nullptr;
};

// arithmetic::best_multiexp(coeffs, bases): assert_eq!(coeffs.len(), bases.len())
inline de_g1 best_multiexp(const Context& ctx, const std::vector<de_fr>& coeffs, const std::vector<de_g1_affine>& bases) {
    if (coeffs.size() != bases.size()) throw std::runtime_error("best_multiexp: coeffs.len() != bases.len()");
    de_g1 out;
    ctx.check(de_msm(ctx.handle(), coeffs.data(), bases.data(), coeffs.size(), &out), "best_multiexp");
    return out;
}
// arithmetic::best_fft(a, omega, log_n): assert_eq!(a.len(), 1 << log_n); in place
inline void best_fft(const Context& ctx, std::vector<de_fr>& a, const de_fr& omega, uint32_t log_n) {
    if (a.size() != (size_t(1) << log_n)) throw std::runtime_error("best_fft: a.len() != 1 << log_n");
    ctx.check(de_ntt(ctx.handle(), a.data(), &omega, log_n), "best_fft");
}

class EvaluationDomain {  // poly::EvaluationDomain::new(j, k)
public:
    EvaluationDomain(const Context& ctx, uint32_t j, uint32_t k) : ctx_(ctx), k_(k) {
        ctx.check(de_domain_create(ctx.handle(), j, k, &h_), "EvaluationDomain::new");
        de_fr c[4];
        ctx.check(de_domain_info(h_, &extended_k_, c), "EvaluationDomain::new");
        omega = c[0]; omega_inv = c[1]; extended_omega = c[2]; extended_omega_inv = c[3];
    }
    ~EvaluationDomain() { de_domain_free(h_); }
    EvaluationDomain(const EvaluationDomain&) = delete;
    uint32_t k() const { return k_; }
    uint32_t extended_k() const { return extended_k_; }
    size_t extended_len() const { return size_t(1) << extended_k_; }
    std::vector<de_fr> coeff_to_extended(const std::vector<de_fr>& a) const {
        need(a.size(), size_t(1) << k_, "coeff_to_extended");
        std::vector<de_fr> out(extended_len());
        ctx_.check(de_coeff_to_extended(h_, a.data(), out.data()), "coeff_to_extended");
        return out;
    }
    std::vector<de_fr> extended_to_coeff(std::vector<de_fr> a) const {
        need(a.size(), extended_len(), "extended_to_coeff");
        size_t len = 0;
        ctx_.check(de_extended_to_coeff(h_, a.data(), &len), "extended_to_coeff");
        a.resize(len);
        return a;
    }
    void lagrange_to_coeff(std::vector<de_fr>& a) const {
        need(a.size(), size_t(1) << k_, "lagrange_to_coeff");
        ctx_.check(de_lagrange_to_coeff(h_, a.data()), "lagrange_to_coeff");
    }
    void coeff_to_lagrange(std::vector<de_fr>& a) const {
        need(a.size(), size_t(1) << k_, "coeff_to_lagrange");
        ctx_.check(de_coeff_to_lagrange(h_, a.data()), "coeff_to_lagrange");
    }
    void divide_by_vanishing_poly(std::vector<de_fr>& a) const {
        need(a.size(), extended_len(), "divide_by_vanishing_poly");
        ctx_.check(de_divide_by_vanishing(h_, a.data()), "divide_by_vanishing_poly");
    }
    de_domain* handle() const { return h_; }
    de_fr omega, omega_inv, extended_omega, extended_omega_inv;

private:
    static void need(size_t got, size_t want, const char* what) {
        if (got != want) throw std::runtime_error(std::string(what) + ": wrong polynomial length");
    }
    const Context& ctx_;
    de_domain* h_ = nullptr;
    uint32_t k_, extended_k_ = 0;
};

class ParamsKZG {  // poly::kzg::commitment::ParamsKZG: g / g_lagrange staged in HBM once
public:
    ParamsKZG(const Context& ctx, uint32_t k, const de_g1_affine* g, const de_g1_affine* g_lagrange) : ctx_(ctx), k_(k) {
        ctx.check(de_params_upload(ctx.handle(), k, g, g_lagrange, &h_), "ParamsKZG");
    }
    ~ParamsKZG() { de_params_free(h_); }
    ParamsKZG(const ParamsKZG&) = delete;
    uint32_t k() const { return k_; }
    de_g1 commit(const de_fr* poly, size_t n) const { return commit_basis(0, poly, n); }           // Blind unused for KZG
    de_g1 commit_lagrange(const de_fr* poly, size_t n) const { return commit_basis(1, poly, n); }
    std::vector<de_g1> commit_lagrange_batch(const std::vector<const de_fr*>& polys, size_t n) const {
        std::vector<de_g1> out(polys.size());
        ctx_.check(de_commit_batch(h_, 1, polys.data(), n, polys.size(), out.data()), "commit_lagrange (batch)");
        return out;
    }

private:
public:
    de_params* handle() const { return h_; }

private:
    de_g1 commit_basis(int basis, const de_fr* poly, size_t n) const {
        de_g1 out;
        ctx_.check(de_commit(h_, basis, poly, n, &out), "commit");
        return out;
    }
    const Context& ctx_;
    de_params* h_ = nullptr;
    uint32_t k_;
};

// arithmetic::{eval_polynomial, kate_division}
inline de_fr eval_polynomial(const Context& ctx, const std::vector<de_fr>& poly, const de_fr& point) {
    de_fr out;
    ctx.check(de_eval_polynomial(ctx.handle(), poly.data(), poly.size(), &point, &out), "eval_polynomial");
    return out;
}
inline std::vector<de_fr> kate_division(const Context& ctx, const std::vector<de_fr>& a, const de_fr& b) {
    if (a.size() < 2) throw std::runtime_error("kate_division: polynomial needs at least 2 coefficients");
    std::vector<de_fr> q(a.size() - 1);
    ctx.check(de_kate_division(ctx.handle(), a.data(), a.size(), &b, q.data()), "kate_division");
    return q;
}

// ---- SerdeFormat::RawBytes key files (/root/reference/benches/delay_enc.rs:84-115: vk.write / VerifyingKey::read, pk.write /
// ProvingKey::read).  Layout (halo2_proofs v2023_04_20; the Python twin is de_b200/serde.py):
//   vk := k: u32 | n_fixed: u32 | fixed commitments (64 B raw affine each) | permutation commitments | selectors (n bits each)
//   pk := vk | l0 | l_last | l_active_row | fixed_values | fixed_polys | fixed_cosets | permutations | polys | cosets
//   polynomial := len: u32 | len x 32 B Montgomery limbs;  slice := count: u32 | polynomials;  u32 headers big-endian (a
//   little-endian file is recognised by its k).  n_perm / n_selectors come from the constraint system, as in halo2.
struct VerifyingKeyRaw {
    uint32_t k = 0;
    std::vector<de_g1_affine> fixed_commitments, permutation_commitments;
    std::vector<std::vector<bool>> selectors;
};
struct ProvingKeyRaw {
    VerifyingKeyRaw vk;
    std::vector<de_fr> l0, l_last, l_active_row;
    std::vector<std::vector<de_fr>> fixed_values, fixed_polys, fixed_cosets, permutations, polys, cosets;
};
namespace detail {
struct RawReader {
    const std::vector<uint8_t>& b;
    size_t pos = 0;
    bool big = true;
    explicit RawReader(const std::vector<uint8_t>& bytes) : b(bytes) {
        if (b.size() < 8) throw std::runtime_error("RawBytes key file is truncated");
        const uint32_t be = (uint32_t(b[0]) << 24) | (uint32_t(b[1]) << 16) | (uint32_t(b[2]) << 8) | b[3];
        const uint32_t le = (uint32_t(b[3]) << 24) | (uint32_t(b[2]) << 16) | (uint32_t(b[1]) << 8) | b[0];
        if (be >= 1 && be <= 28) big = true;
        else if (le >= 1 && le <= 28) big = false;
        else throw std::runtime_error("not a RawBytes key file: k out of range in either byte order");
    }
    const uint8_t* take(size_t n) {
        if (pos + n > b.size()) throw std::runtime_error("RawBytes key file is truncated");
        const uint8_t* p = b.data() + pos;
        pos += n;
        return p;
    }
    uint32_t u32() {
        const uint8_t* p = take(4);
        return big ? (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]
                   : (uint32_t(p[3]) << 24) | (uint32_t(p[2]) << 16) | (uint32_t(p[1]) << 8) | p[0];
    }
    std::vector<de_fr> polynomial() {
        const uint32_t m = u32();
        std::vector<de_fr> v(m);
        std::memcpy(v.data(), take(size_t(m) * sizeof(de_fr)), size_t(m) * sizeof(de_fr));
        return v;
    }
    std::vector<std::vector<de_fr>> slice() {
        const uint32_t count = u32();
        std::vector<std::vector<de_fr>> v;
        for (uint32_t i = 0; i < count; i++) v.push_back(polynomial());
        return v;
    }
    VerifyingKeyRaw vk(uint32_t n_perm, uint32_t n_selectors) {
        VerifyingKeyRaw v;
        v.k = u32();
        const uint32_t nf = u32();
        if (v.k < 1 || v.k > 28 || nf > 4096) throw std::runtime_error("not a RawBytes verifying key");
        v.fixed_commitments.resize(nf);
        std::memcpy(v.fixed_commitments.data(), take(size_t(nf) * 64), size_t(nf) * 64);
        v.permutation_commitments.resize(n_perm);
        std::memcpy(v.permutation_commitments.data(), take(size_t(n_perm) * 64), size_t(n_perm) * 64);
        const size_t n = size_t(1) << v.k;
        for (uint32_t s = 0; s < n_selectors; s++) {
            const uint8_t* p = take((n + 7) / 8);
            std::vector<bool> bits(n);
            for (size_t i = 0; i < n; i++) bits[i] = (p[i / 8] >> (i % 8)) & 1;
            v.selectors.push_back(bits);
        }
        return v;
    }
};
}  // namespace detail
// VerifyingKey::read(reader, SerdeFormat::RawBytes)
inline VerifyingKeyRaw read_verifying_key_raw(const std::vector<uint8_t>& bytes, uint32_t n_perm, uint32_t n_selectors) {
    detail::RawReader r(bytes);
    VerifyingKeyRaw v = r.vk(n_perm, n_selectors);
    if (r.pos != bytes.size()) throw std::runtime_error("trailing bytes after the verifying key");
    return v;
}
// ProvingKey::read(reader, SerdeFormat::RawBytes): fixed_polys / polys are what de_pk_upload takes (the cosets are recomputed on
// the device)
inline ProvingKeyRaw read_proving_key_raw(const std::vector<uint8_t>& bytes, uint32_t n_perm, uint32_t n_selectors) {
    detail::RawReader r(bytes);
    ProvingKeyRaw pk;
    pk.vk = r.vk(n_perm, n_selectors);
    pk.l0 = r.polynomial();
    pk.l_last = r.polynomial();
    pk.l_active_row = r.polynomial();
    pk.fixed_values = r.slice();
    pk.fixed_polys = r.slice();
    pk.fixed_cosets = r.slice();
    pk.permutations = r.slice();
    pk.polys = r.slice();
    pk.cosets = r.slice();
    if (r.pos != bytes.size()) throw std::runtime_error("trailing bytes after the proving key");
    const size_t n = size_t(1) << pk.vk.k;
    if (pk.polys.size() != n_perm || pk.fixed_polys.size() != pk.vk.fixed_commitments.size()) throw std::runtime_error("proving key does not match the constraint system");
    for (auto& p : pk.fixed_polys) if (p.size() != n) throw std::runtime_error("proving key: wrong polynomial length");
    for (auto& p : pk.polys) if (p.size() != n) throw std::runtime_error("proving key: wrong polynomial length");
    return pk;
}

class ProvingKey {  // the evaluator's / prover's view of plonk::ProvingKey, staged in HBM (de_pk_upload)
public:
    ProvingKey(const Context& ctx, const EvaluationDomain& domain, const de_pk_desc& desc) : ctx_(ctx) {
        ctx.check(de_pk_upload(domain.handle(), &desc, &h_), "ProvingKey");
    }
    ~ProvingKey() { de_pk_free(h_); }
    ProvingKey(const ProvingKey&) = delete;
    de_pk* handle() const { return h_; }

private:
    const Context& ctx_;
    de_pk* h_ = nullptr;
};

// plonk::create_proof for one circuit (KZG, ProverGWC, Blake2bWrite / Challenge255): every polynomial stays in HBM.
// `randoms` are the Fr::random(rng) draws in create_proof's order (random_count() of them); returns the proof bytes.
class Prover {
public:
    Prover(const Context& ctx, const ParamsKZG& params, const ProvingKey& pk, const de_prover_desc& desc) : ctx_(ctx) {
        ctx.check(de_prover_create(params.handle(), pk.handle(), &desc, &h_), "Prover");
    }
    ~Prover() { de_prover_free(h_); }
    Prover(const Prover&) = delete;
    size_t random_count() const { return de_prover_random_count(h_); }
    size_t proof_size() const { return de_prover_proof_size(h_); }
    std::vector<uint8_t> create_proof(const std::vector<const de_fr*>& advice, const std::vector<std::vector<de_fr>>& instances,
                                      const std::vector<de_fr>& randoms) const {
        std::vector<const de_fr*> ip;
        std::vector<size_t> il;
        for (const auto& v : instances) {
            ip.push_back(v.data());
            il.push_back(v.size());
        }
        std::vector<uint8_t> proof(proof_size());
        size_t len = 0;
        ctx_.check(de_create_proof(h_, advice.data(), ip.data(), il.data(), randoms.data(), randoms.size(), proof.data(), proof.size(), &len),
                   "create_proof");
        proof.resize(len);
        return proof;
    }

private:
    const Context& ctx_;
    de_prover* h_ = nullptr;
};

// MSM with the bases sharded by contiguous ranges over several GPUs of ONE process (de_commit_sharded)
inline de_g1 commit_sharded(const std::vector<const ParamsKZG*>& shards, const std::vector<size_t>& lo, const std::vector<size_t>& len, int basis,
                            const de_fr* scalars, const Context& ctx0) {
    std::vector<de_params*> h;
    for (auto* s : shards) h.push_back(s->handle());
    de_g1 out;
    ctx0.check(de_commit_sharded(h.data(), lo.data(), len.data(), (int)h.size(), basis, scalars, &out), "commit_sharded");
    return out;
}

// arithmetic::best_fft for a vector spread over several GPUs of ONE process (de_ntt_sharded): natural order in and out
inline void best_fft_sharded(const std::vector<const Context*>& ctxs, std::vector<de_fr>& a, const de_fr& omega, uint32_t log_n) {
    if (ctxs.empty()) throw std::runtime_error("best_fft_sharded: no contexts");
    if (a.size() != (size_t(1) << log_n)) throw std::runtime_error("best_fft: a.len() != 1 << log_n");
    std::vector<de_ctx*> h;
    for (auto* c : ctxs) h.push_back(c->handle());
    ctxs[0]->check(de_ntt_sharded(h.data(), (int)h.size(), a.data(), &omega, log_n), "best_fft_sharded");
}

// ---- the reference's circuits (Circuit::synthesize as a host witness generator; no Context, no GPU) ------------------------
// What halo2's two synthesis passes hand to keygen (fixed columns, copy constraints -> sigma) and to create_proof (advice).
class Assignment {
public:
    explicit Assignment(de_assignment* h) : h_(h) {
        if (de_assignment_info(h_, &info_) != DE_OK) throw std::runtime_error("de_assignment_info failed");
    }
    ~Assignment() { de_assignment_free(h_); }
    Assignment(const Assignment&) = delete;
    Assignment(Assignment&& o) noexcept : h_(o.h_), info_(o.info_) { o.h_ = nullptr; }
    uint32_t k() const { return info_.k; }
    size_t rows() const { return size_t(1) << info_.k; }
    uint32_t n_fixed() const { return info_.n_fixed; }
    uint32_t n_advice() const { return info_.n_advice; }
    uint64_t used_rows() const { return info_.used_rows; }
    double synthesis_ms() const { return info_.synthesis_ms; }
    std::vector<de_fr> fixed(uint32_t column) const { return column_of(de_assignment_fixed, column, "fixed"); }
    std::vector<de_fr> advice(uint32_t column) const { return column_of(de_assignment_advice, column, "advice"); }
    // (left column, left row, right column, right row) per copy constraint; columns are positions in the permutation
    std::vector<uint32_t> copies() const {
        std::vector<uint32_t> c(4 * info_.n_copies);
        if (info_.n_copies && de_assignment_copies(h_, c.data()) != DE_OK) throw std::runtime_error("de_assignment_copies failed");
        return c;
    }
    // the circuit's results (x^e mod n limbs, key words, ciphertext ...; de_b200.h: de_assignment_outputs)
    std::vector<de_fr> outputs() const {
        std::vector<de_fr> o(info_.n_outputs ? info_.n_outputs : 1);
        if (de_assignment_outputs(h_, o.data()) != DE_OK) throw std::runtime_error("de_assignment_outputs failed");
        o.resize(info_.n_outputs);
        return o;
    }
    // permutation::keygen::Assembly::build_pk: n_columns sigma columns in lagrange form
    std::vector<de_fr> sigma(const de_fr& omega, const de_fr& delta, uint32_t n_columns) const {
        std::vector<de_fr> s(size_t(n_columns) * rows());
        if (de_assignment_sigma(h_, &omega, &delta, n_columns, s.data()) != DE_OK) throw std::runtime_error(de_frontend_last_error());
        return s;
    }

private:
    template <typename Fn>
    std::vector<de_fr> column_of(Fn fn, uint32_t column, const char* what) const {
        std::vector<de_fr> v(rows());
        if (fn(h_, column, v.data()) != DE_OK) throw std::runtime_error(std::string("no such ") + what + " column");
        return v;
    }
    de_assignment* h_;
    de_assignment_info_t info_;
};

// Common part of the three bench circuits: synthesize() is keygen's pass, witness() the pass create_proof makes - advice columns
// only, written column after column into the caller's buffer (5 * 2^k elements, e.g. the pinned buffer the prover uploads from).
class Circuit {
public:
    Assignment synthesize(uint32_t k) const {
        de_circuit_desc d = desc(k);
        de_assignment* a = nullptr;
        if (de_circuit_synthesize(&d, &a) != DE_OK) throw std::runtime_error(de_frontend_last_error());
        return Assignment(a);
    }
    de_assignment_info_t witness(uint32_t k, de_fr* advice_out, uint32_t threads = 1, bool reuse_buffer = false) const {
        de_circuit_desc d = desc(k);
        d.threads = threads;
        d.reuse_buffer = reuse_buffer ? 1 : 0;
        de_assignment_info_t info;
        if (de_circuit_witness(&d, advice_out, &info) != DE_OK) throw std::runtime_error(de_frontend_last_error());
        return info;
    }
    virtual ~Circuit() {}

protected:
    virtual de_circuit_desc desc(uint32_t k) const = 0;
    static de_circuit_desc blank(uint32_t kind, uint32_t k) {
        de_circuit_desc d;
        std::memset(&d, 0, sizeof d);
        d.kind = kind;
        d.k = k;
        d.bits_len = 2048;  // BITS_LEN, src/lib.rs:38
        d.exp_bits = 5;     // EXP_LIMB_BITS, src/lib.rs:39
        return d;
    }
};

// benches/mod_pow.rs:36-140 RSACircuit: x^e mod n; integers as little-endian bytes
class RSACircuit : public Circuit {
public:
    RSACircuit(std::vector<uint8_t> n, std::vector<uint8_t> e, std::vector<uint8_t> x) : n_(std::move(n)), e_(std::move(e)), x_(std::move(x)) {}

protected:
    de_circuit_desc desc(uint32_t k) const override { return with_rsa(blank(DE_CIRCUIT_MOD_POW, k)); }
    de_circuit_desc with_rsa(de_circuit_desc d) const {
        d.n = n_.data(); d.n_len = n_.size();
        d.e = e_.data(); d.e_len = e_.size();
        d.x = x_.data(); d.x_len = x_.size();
        return d;
    }
    std::vector<uint8_t> n_, e_, x_;
};

// src/lib.rs:103-318 DelayEncryptCircuit: x^e mod n -> Poseidon hash -> key -> Poseidon encryption of `message`
class DelayEncryptCircuit : public RSACircuit {
public:
    DelayEncryptCircuit(std::vector<uint8_t> n, std::vector<uint8_t> e, std::vector<uint8_t> x, std::vector<de_fr> message)
        : RSACircuit(std::move(n), std::move(e), std::move(x)), message_(std::move(message)) {}

protected:
    de_circuit_desc desc(uint32_t k) const override {
        de_circuit_desc d = with_rsa(blank(DE_CIRCUIT_DELAY_ENC, k));
        d.message = message_.data();
        d.message_len = (uint32_t)message_.size();
        return d;
    }
    std::vector<de_fr> message_;
};

// src/encryption/chip.rs:114-198 PoseidonEncCircuit: duplex encryption of `message` under key (k0, k1)
class PoseidonEncCircuit : public Circuit {
public:
    PoseidonEncCircuit(const de_fr& k0, const de_fr& k1, std::vector<de_fr> message) : message_(std::move(message)) { key_[0] = k0; key_[1] = k1; }

protected:
    de_circuit_desc desc(uint32_t k) const override {
        de_circuit_desc d = blank(DE_CIRCUIT_POSE_ENC, k);
        d.message = message_.data();
        d.message_len = (uint32_t)message_.size();
        d.key[0] = key_[0];
        d.key[1] = key_[1];
        return d;
    }
    de_fr key_[2];
    std::vector<de_fr> message_;
};

}  // namespace halo2_b200
