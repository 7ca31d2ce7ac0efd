// poseidon.hpp — the reference's Poseidon over BN254 Fr: parameter generation (Grain LFSR, Cauchy MDS, optimised round
// constants, sparse matrices), the native permutation / sponge / duplex encryption used for expected values, and the
// in-circuit permutation on the row emitter of maingate.hpp.
//   native:  /root/reference/src/poseidon/{grain,matrix,spec,permutation,poseidon}.rs, src/encryption/poseidon_enc.rs
//   circuit: /root/reference/src/poseidon/chip.rs, src/hash/chip.rs, src/encryption/chip.rs
// The width T is a run-time value here (the reference fixes it with const generics: T = 5, RATE = 4 in every circuit; its
// known-answer tests also use T = 3), so matrices are vectors of rows.
#pragma once
#include <map>
#include <memory>
#include <mutex>
#include <tuple>

#include "maingate.hpp"

namespace de {
namespace fe {

typedef std::vector<F> Vec;
typedef std::vector<Vec> Mat;

// ---- matrix.rs ----------------------------------------------------------------------------------------------------------
inline Mat mat_identity(size_t t) {
    Mat m(t, Vec(t));
    for (size_t i = 0; i < t; i++) m[i][i] = F::one();
    return m;
}
inline Mat mat_transpose(const Mat& a) {
    Mat r(a.size(), Vec(a.size()));
    for (size_t i = 0; i < a.size(); i++)
        for (size_t j = 0; j < a.size(); j++) r[j][i] = a[i][j];
    return r;
}
inline Mat mat_mul(const Mat& a, const Mat& b) {
    const size_t t = a.size();
    Mat r(t, Vec(t));
    for (size_t i = 0; i < t; i++)
        for (size_t j = 0; j < t; j++)
            for (size_t k = 0; k < t; k++) r[i][j] = r[i][j] + a[i][k] * b[k][j];
    return r;
}
inline Vec mat_mul_vector(const Mat& a, const Vec& v) {
    Vec r(a.size());
    for (size_t i = 0; i < a.size(); i++)
        for (size_t j = 0; j < v.size(); j++) r[i] = r[i] + v[j] * a[i][j];
    return r;
}
// matrix.rs:86-125: Gauss-Jordan on [m | I] without pivoting, as the reference does it
inline Mat mat_invert(const Mat& a) {
    const size_t t = a.size();
    Mat m(t, Vec(2 * t));
    for (size_t i = 0; i < t; i++) {
        for (size_t j = 0; j < t; j++) m[i][j] = a[i][j];
        m[i][t + i] = F::one();
    }
    for (size_t i = 0; i < t; i++)
        for (size_t j = 0; j < t; j++)
            if (i != j) {
                const F r = m[j][i] * m[i][i].invert();
                for (size_t k = 0; k < 2 * t; k++) m[j][k] = m[j][k] - r * m[i][k];
            }
    Mat res(t, Vec(t));
    for (size_t i = 0; i < t; i++) {
        const F inv = m[i][i].invert();
        for (size_t j = 0; j < t; j++) res[i][j] = m[i][t + j] * inv;
    }
    return res;
}

// ---- grain.rs -----------------------------------------------------------------------------------------------------------
class Grain {
  public:
    Grain(uint32_t t, uint32_t r_f, uint32_t r_p) {
        auto append = [&](int nbits, uint64_t v) {
            for (int i = nbits - 1; i >= 0; i--) bits.push_back((v >> i) & 1);
        };
        append(2, 1);     // prime field
        append(4, 0);     // x^alpha s-box
        append(12, 254);  // Fr::NUM_BITS
        append(12, t);
        append(10, r_f);
        append(10, r_p);
        append(30, (1u << 30) - 1);
        for (int i = 0; i < 160; i++) new_bit();
    }
    F next_field_element() {  // rejection sampling of 254-bit big-endian draws
        for (;;) {
            uint8_t bytes[32];
            draw(bytes, 32);
            const BigUint v = BigUint::from_bytes_le(bytes, 32);
            if (v < modulus()) return F::from_big(v);
        }
    }
    F next_field_element_without_rejection() {  // the same draw reduced mod r (from_uniform_bytes of 64 bytes)
        uint8_t bytes[64];
        draw(bytes, 64);
        return F(host::fr_from_wide(bytes));
    }

  private:
    std::vector<uint8_t> bits;  // 80-bit state, bits[0] is the oldest
    static BigUint modulus() { return BigUint::from_limbs(host::FR_FIELD.mod, 4); }
    bool new_bit() {
        const uint8_t b = bits[0] ^ bits[62] ^ bits[51] ^ bits[38] ^ bits[23] ^ bits[13];
        bits.erase(bits.begin());
        bits.push_back(b);
        return b;
    }
    bool next() {  // Iterator::next: keep the bit after a 1, drop the bit after a 0
        while (!new_bit()) new_bit();
        return new_bit();
    }
    void draw(uint8_t* bytes, size_t len) {
        memset(bytes, 0, len);
        for (int i = 0; i < 254; i++) {
            const int pos = 254 - 1 - i;  // MSB first
            if (next()) bytes[pos / 8] |= (uint8_t)(1u << (pos % 8));
        }
    }
};

// ---- spec.rs ------------------------------------------------------------------------------------------------------------
struct SparseMDS {
    Vec row, col_hat;
};
struct Spec {
    uint32_t t = 0, r_f = 0, r_p = 0;
    Mat mds, pre_sparse_mds;
    std::vector<SparseMDS> sparse;
    std::vector<Vec> start, end;
    Vec partial;

    Spec(uint32_t t_, uint32_t r_f_, uint32_t r_p_) : t(t_), r_f(r_f_), r_p(r_p_) {
        Grain grain(t, r_f, r_p);
        std::vector<Vec> constants(r_f + r_p, Vec(t));
        for (auto& rc : constants)
            for (auto& c : rc) c = grain.next_field_element();
        Vec xs(t), ys(t);
        for (auto& x : xs) x = grain.next_field_element_without_rejection();
        for (auto& y : ys) y = grain.next_field_element_without_rejection();
        mds.assign(t, Vec(t));
        for (uint32_t i = 0; i < t; i++)
            for (uint32_t j = 0; j < t; j++) mds[i][j] = (xs[i] + ys[j]).invert();  // spec.rs:137-147 cauchy
        optimized_constants(constants);
        sparse_matrices();
    }

  private:
    // spec.rs:326-377 calculate_optimized_constants
    void optimized_constants(const std::vector<Vec>& constants) {
        const Mat inverse_mds = mat_invert(mds);
        const uint32_t half = r_f / 2;
        start.assign(half, Vec(t));
        start[0] = constants[0];
        for (uint32_t i = 1; i < half; i++) start[i] = mat_mul_vector(inverse_mds, constants[i]);
        Vec acc = constants[half + r_p];
        partial.assign(r_p, F::zero());
        for (uint32_t idx = 0; idx < r_p; idx++) {
            // partial rounds walked backwards: constants[half + r_p - 1 - idx]
            Vec tmp = mat_mul_vector(inverse_mds, acc);
            partial[r_p - 1 - idx] = tmp[0];
            tmp[0] = F::zero();
            const Vec& rc = constants[half + r_p - 1 - idx];
            for (uint32_t j = 0; j < t; j++) acc[j] = tmp[j] + rc[j];
        }
        start.push_back(mat_mul_vector(inverse_mds, acc));
        end.assign(half - 1, Vec(t));
        for (uint32_t i = 0; i + 1 < half; i++) end[i] = mat_mul_vector(inverse_mds, constants[half + r_p + 1 + i]);
    }
    // spec.rs:166-199 factorise: m = m' * m'' with m' = [[1, 0], [0, m_hat]] and m'' sparse
    static void factorise(const Mat& m, Mat* prime, SparseMDS* sparse_out) {
        const size_t t = m.size();
        Vec w(t - 1);
        Mat m_hat(t - 1, Vec(t - 1));
        for (size_t i = 1; i < t; i++) {
            w[i - 1] = m[i][0];
            for (size_t j = 1; j < t; j++) m_hat[i - 1][j - 1] = m[i][j];
        }
        const Vec w_hat = mat_mul_vector(mat_invert(m_hat), w);
        *prime = mat_identity(t);
        for (size_t i = 1; i < t; i++)
            for (size_t j = 1; j < t; j++) (*prime)[i][j] = m_hat[i - 1][j - 1];
        // m'' = [[m_00 | m_0i], [w_hat | I]], transposed, read as (first row, first column below the corner)
        Mat pp = mat_identity(t);
        pp[0] = m[0];
        for (size_t i = 1; i < t; i++) pp[i][0] = w_hat[i - 1];
        pp = mat_transpose(pp);
        sparse_out->row = pp[0];
        sparse_out->col_hat.assign(t - 1, F::zero());
        for (size_t i = 1; i < t; i++) sparse_out->col_hat[i - 1] = pp[i][0];
    }
    // spec.rs:379-396 calculate_sparse_matrices
    void sparse_matrices() {
        const Mat mt = mat_transpose(mds);
        Mat acc = mt;
        sparse.resize(r_p);
        for (uint32_t i = 0; i < r_p; i++) {
            Mat prime;
            factorise(acc, &prime, &sparse[r_p - 1 - i]);
            acc = mat_mul(mt, prime);
        }
        pre_sparse_mds = mat_transpose(acc);
    }

  public:
    static void sbox_full(Vec& s) {
        for (auto& e : s) e = e.pow5();
    }
    static void apply_sparse(const SparseMDS& m, Vec& s) {
        Vec w = s;
        F s0 = F::zero();
        for (size_t i = 0; i < w.size(); i++) s0 = s0 + m.row[i] * w[i];
        s[0] = s0;
        for (size_t i = 1; i < w.size(); i++) s[i] = m.col_hat[i - 1] * w[0] + w[i];
    }
    // permutation.rs:5-48 Spec::permute
    void permute(Vec& s) const {
        const uint32_t half = r_f / 2;
        for (uint32_t i = 0; i < t; i++) s[i] = s[i] + start[0][i];
        for (uint32_t r = 1; r < half; r++) {
            sbox_full(s);
            for (uint32_t i = 0; i < t; i++) s[i] = s[i] + start[r][i];
            s = mat_mul_vector(mds, s);
        }
        sbox_full(s);
        for (uint32_t i = 0; i < t; i++) s[i] = s[i] + start.back()[i];
        s = mat_mul_vector(pre_sparse_mds, s);
        for (uint32_t r = 0; r < r_p; r++) {
            s[0] = s[0].pow5() + partial[r];
            apply_sparse(sparse[r], s);
        }
        for (const Vec& rc : end) {
            sbox_full(s);
            for (uint32_t i = 0; i < t; i++) s[i] = s[i] + rc[i];
            s = mat_mul_vector(mds, s);
        }
        sbox_full(s);
        s = mat_mul_vector(mds, s);
    }
};

// Spec::new is parameter generation (Grain stream, 57 matrix factorisations: tens of milliseconds); the reference builds it once
// per bench outside create_proof (benches/delay_enc.rs:69) and hands it to the circuit.  Here: one shared instance per
// (t, r_f, r_p), built on first use.
inline const Spec& shared_spec(uint32_t t, uint32_t r_f, uint32_t r_p) {
    static std::mutex mu;
    static std::map<std::tuple<uint32_t, uint32_t, uint32_t>, std::unique_ptr<Spec>> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto& slot = cache[std::make_tuple(t, r_f, r_p)];
    if (!slot) slot.reset(new Spec(t, r_f, r_p));
    return *slot;
}

// poseidon.rs: the sponge the reference hashes and encrypts with
struct Poseidon {
    const Spec& spec;
    Vec state, absorbing;
    Poseidon(const Spec& s, const Vec& init) : spec(s), state(init) {}
    static Vec hash_state(uint32_t t) {  // State::default(): [2^64, 0, ...]
        Vec s(t);
        s[0] = F::from_big(BigUint::pow2(64));
        return s;
    }
    void update(const Vec& elements) {
        Vec input = absorbing;
        input.insert(input.end(), elements.begin(), elements.end());
        const size_t rate = spec.t - 1;
        for (size_t off = 0; off < input.size(); off += rate) {
            const size_t len = std::min(rate, input.size() - off);
            if (len < rate) {
                absorbing.assign(input.begin() + off, input.begin() + off + len);
            } else {
                for (size_t i = 0; i < len; i++) state[1 + i] = state[1 + i] + input[off + i];
                spec.permute(state);
                absorbing.clear();
            }
        }
    }
    Vec squeeze(int h_flag) {
        Vec last = absorbing;
        if (h_flag == 1) last.push_back(F::one());
        for (size_t i = 0; i < last.size(); i++) state[1 + i] = state[1 + i] + last[i];
        spec.permute(state);
        absorbing.clear();
        return state;
    }
};

enum { MESSAGE_CAPACITY = 2, CIPHER_SIZE = 3 };  // encryption/poseidon_enc.rs:10-11

// encryption/poseidon_enc.rs:98-131 PoseidonCipher::encrypt, statement by statement: the additions of the message to the
// sponge state act on a COPY of the state (State::words() returns one), so the duplex only ever absorbs what update()
// receives - whole RATE-sized chunks - and a shorter message reaches the ciphertext words but not the tag.
inline Vec poseidon_encrypt(const Spec& spec, const F& k0, const F& k1, const Vec& message) {
    Poseidon enc(spec, {F::zero(), F::zero(), k0, k1, F::one()});
    Vec cipher(CIPHER_SIZE);
    enc.update({});
    enc.squeeze(0);
    size_t i = 0;
    const size_t rate = spec.t - 1;
    for (size_t off = 0; off < message.size(); off += rate) {
        const size_t len = std::min(rate, message.size() - off);
        for (size_t j = 0; j < len; j++)
            if (i < MESSAGE_CAPACITY) cipher[i++] = enc.state[1 + j] + message[off + j];
        if (len == rate) enc.update(Vec(message.begin() + off, message.begin() + off + len));
        else enc.squeeze(0);
    }
    cipher[MESSAGE_CAPACITY] = enc.state[1];
    return cipher;
}
// encryption/poseidon_enc.rs:133-165 PoseidonCipher::decrypt; returns false when the tag does not match
inline bool poseidon_decrypt(const Spec& spec, const F& k0, const F& k1, const Vec& cipher, Vec* message) {
    Poseidon enc(spec, {F::zero(), F::zero(), k0, k1, F::one()});
    enc.update({});
    enc.squeeze(0);
    Vec state_2 = enc.state, msg(MESSAGE_CAPACITY);
    for (size_t i = 0; i < MESSAGE_CAPACITY; i++) {
        msg[i] = cipher[i] - state_2[(i + 1) % spec.t];
        state_2[(i + 1) % spec.t] = cipher[i];
    }
    enc.update(msg);
    enc.squeeze(0);
    if (cipher[MESSAGE_CAPACITY] != enc.state[1]) return false;
    *message = msg;
    return true;
}

// ---- poseidon/chip.rs ---------------------------------------------------------------------------------------------------
class PoseidonChip {
  public:
    PoseidonChip(MainGate& g, const Spec& s, const std::vector<Cell>& initial) : gate(g), spec(s), state(initial) {}
    MainGate& gate;
    const Spec& spec;
    std::vector<Cell> state, absorbing;

    static PoseidonChip new_hash(MainGate& g, const Spec& s) {  // chip.rs:128-150: constants [2^64, 0, 0, 0, 0]
        std::vector<Cell> init;
        for (const F& w : Poseidon::hash_state(s.t)) init.push_back(g.assign_constant(w));
        return PoseidonChip(g, s, init);
    }
    static PoseidonChip new_enc(MainGate& g, const Spec& s, const F& k0, const F& k1, bool as_witness) {
        // chip.rs:52-88 new_enc assigns the initial state [0, 0, k0, k1, 1] as constants, :90-126 new_enc_de as witness values
        std::vector<Cell> init;
        for (const F& w : Vec{F::zero(), F::zero(), k0, k1, F::one()}) init.push_back(as_witness ? g.assign_value(w) : g.assign_constant(w));
        return PoseidonChip(g, s, init);
    }
    void sbox_full(const Vec& constants) {  // chip.rs:192-200
        for (size_t i = 0; i < state.size(); i++) {
            Cell t = gate.mul(state[i], state[i]);
            t = gate.mul(t, t);
            state[i] = gate.mul_add_constant(t, state[i], constants[i]);
        }
    }
    void sbox_part(const F& constant) {  // chip.rs:202-211
        Cell t = gate.mul(state[0], state[0]);
        t = gate.mul(t, t);
        state[0] = gate.mul_add_constant(t, state[0], constant);
    }
    // chip.rs:214-268: state[0] += c0; state[1 + i] += input_i + c_(1+i); the rest += c (+ 1 on the first of them when hashing)
    void absorb_with_pre_constants(const std::vector<Cell>& inputs, const Vec& pre, bool h_flag) {
        if (inputs.size() >= spec.t) throw std::runtime_error("absorb: more inputs than the rate");
        const size_t offset = inputs.size() + 1;
        state[0] = gate.add_constant(state[0], pre[0]);
        for (size_t i = 0; i < inputs.size(); i++) state[1 + i] = gate.add_with_constant(state[1 + i], inputs[i], pre[1 + i]);
        for (size_t i = offset; i < spec.t; i++) state[i] = gate.add_constant(state[i], pre[i] + ((h_flag && i == offset) ? F::one() : F::zero()));
    }
    void apply_mds(const Mat& m) {  // chip.rs:270-296
        std::vector<Cell> next;
        for (const Vec& row : m) {
            std::vector<Term> terms;
            for (size_t i = 0; i < state.size(); i++) terms.push_back(Term::assigned(state[i], row[i]));
            next.push_back(gate.compose(terms, F::zero()));
        }
        state = next;
    }
    void apply_sparse_mds(const SparseMDS& m) {  // chip.rs:298-333
        std::vector<Term> terms;
        for (size_t i = 0; i < state.size(); i++) terms.push_back(Term::assigned(state[i], m.row[i]));
        std::vector<Cell> next = {gate.compose(terms, F::zero())};
        for (size_t i = 1; i < state.size(); i++)
            next.push_back(gate.compose({Term::assigned(state[0], m.col_hat[i - 1]), Term::assigned(state[i], F::one())}, F::zero()));
        state = next;
    }
    // chip.rs:335-378 permutation (h_flag false) / :380-419 perm_hash (h_flag true)
    void permutation(const std::vector<Cell>& inputs, bool h_flag) {
        const uint32_t half = spec.r_f / 2;
        absorb_with_pre_constants(inputs, spec.start[0], h_flag);
        for (uint32_t r = 1; r < half; r++) {
            sbox_full(spec.start[r]);
            apply_mds(spec.mds);
        }
        sbox_full(spec.start.back());
        apply_mds(spec.pre_sparse_mds);
        for (uint32_t r = 0; r < spec.r_p; r++) {
            sbox_part(spec.partial[r]);
            apply_sparse_mds(spec.sparse[r]);
        }
        for (const Vec& rc : spec.end) {
            sbox_full(rc);
            apply_mds(spec.mds);
        }
        sbox_full(Vec(spec.t));
        apply_mds(spec.mds);
    }
    // hash/chip.rs:60-86 HasherChip::hash
    std::vector<Cell> hash() {
        const std::vector<Cell> input = absorbing;
        absorbing.clear();
        const size_t rate = spec.t - 1;
        size_t padding_offset = 0;
        for (size_t off = 0; off < input.size(); off += rate) {
            const size_t len = std::min(rate, input.size() - off);
            padding_offset = rate - len;
            permutation(std::vector<Cell>(input.begin() + off, input.begin() + off + len), true);
        }
        if (padding_offset == 0) permutation({}, true);
        return state;
    }
    // encryption/chip.rs:72-112 PoseidonEncChip::absorb_and_relese: the message is added to the rate words (those sums are the
    // ciphertext) and then handed to permutation(), which adds it once more with the round constants
    std::vector<Cell> absorb_and_release() {
        std::vector<Cell> cipher_text;
        const std::vector<Cell> input = absorbing;
        absorbing.clear();
        const size_t rate = spec.t - 1;
        size_t i = 0;
        for (size_t off = 0; off < input.size(); off += rate) {
            const size_t len = std::min(rate, input.size() - off);
            for (size_t j = 0; j < len; j++) {
                state[1 + j] = gate.add(state[1 + j], input[off + j]);
                if (i < MESSAGE_CAPACITY) {
                    cipher_text.push_back(state[1 + j]);
                    i++;
                }
            }
            permutation(std::vector<Cell>(input.begin() + off, input.begin() + off + len), false);
        }
        cipher_text.push_back(state[1]);
        return cipher_text;
    }
};

}  // namespace fe
}  // namespace de
