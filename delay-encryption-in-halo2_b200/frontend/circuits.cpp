// circuits.cpp — the reference's three bench circuits as witness generators, and the C ABI over them (include/de_b200.h,
// section "circuit front-end"):
//   DE_CIRCUIT_MOD_POW    benches/mod_pow.rs:36-140       RSACircuit: x^e mod n, 2048-bit n, variable 5-bit e
//   DE_CIRCUIT_POSE_ENC   src/encryption/chip.rs:114-198  PoseidonEncCircuit: duplex encryption of MESSAGE_CAPACITY words
//   DE_CIRCUIT_DELAY_ENC  src/lib.rs:103-318              DelayEncryptCircuit: mod-pow -> Poseidon hash -> key -> encryption
//   DE_CIRCUIT_RSA_PKCS1  src/rsa/chip.rs:119-212         signature check with e = 65537 (the reference's known-answer triples)
//   DE_CIRCUIT_BIGINT_SQUARE src/big_integer/chip.rs:2918-3030  the chip's own square test (its 31 limb products are a known answer)
//   DE_CIRCUIT_BIGINT_OPS    src/big_integer/chip.rs:1479-2806  the chip's operator tests (add, sub, mul_mod, pow_mod, comparisons)
//   DE_CIRCUIT_POSEIDON_HASH src/hash/chip.rs:113-236           PoseidonHashCircuit: HasherChip against the native sponge
// A synthesis pass fills the fixed columns (selectors, constants, range tables), the advice columns (the witness) and the
// copy constraints of de_b200/plonk.py: main_gate_shape().  keygen consumes fixed + copies (de_assignment_sigma builds the
// permutation columns), create_proof consumes the advice columns.  Host code only: no GPU is needed or used.
#include <chrono>
#include <memory>

#include "../../include/de_b200.h"
#include "bigint_chip.hpp"
#include "poseidon.hpp"

using namespace de::fe;

struct de_assignment {
    Assignment as;
    double synth_ms = 0;
};

namespace {

thread_local std::string g_frontend_error;

F fr_in(const de_fr& v) {
    de::host::HFr h;
    memcpy(h.l, v.l, 32);
    return F(h);
}
void fr_out(const F& f, de_fr* out) { memcpy(out->l, f.v.l, 32); }

struct RsaInputs {
    BigUint n, e, x;
};

// the region "rsa modpow with 2048 bits" shared by RSACircuit and DelayEncryptCircuit (benches/mod_pow.rs:104-133,
// src/lib.rs:179-206); returns the constant-assigned result x^e mod n
// the one place where the reference's layout depends on a value: x^e mod n is assigned as a CONSTANT (src/lib.rs:200-203 through
// chip.rs:1255-1285 assign_constant), limb by limb up to its top non-zero limb, so its limb count decides the region's row count
uint32_t constant_limbs(const BigUint& v) { return (uint32_t)((v.bits() + RSAChip::LIMB_WIDTH - 1) / RSAChip::LIMB_WIDTH); }

AssignedInteger rsa_region(RSAChip& rsa, const RsaInputs& in) {
    BigIntChip& bigint = rsa.bigint;
    const uint32_t num_limbs = rsa.bits_len / RSAChip::LIMB_WIDTH;
    if (in.n.bits() > rsa.bits_len || in.x.bits() > rsa.bits_len) throw std::runtime_error("n or x wider than bits_len");
    if (in.e.bits() > rsa.exp_limb_bits) throw std::runtime_error("e wider than exp_limb_bits");
    const std::vector<BigUint> e_limbs = decompose_big(in.e, 1, rsa.exp_limb_bits);
    const AssignedInteger n = bigint.assign_integer(decompose_big(in.n, num_limbs, RSAChip::LIMB_WIDTH));  // assign_public_key
    const AssignedInteger e = bigint.assign_integer(e_limbs);
    const AssignedInteger x = bigint.assign_integer(decompose_big(in.x, num_limbs, RSAChip::LIMB_WIDTH));
    trace_lap("rsa: inputs assigned");
    const AssignedInteger powed = rsa.modpow_var(x, n, e);
    trace_lap("rsa: modpow");
    const BigUint valid = big_pow_mod(in.x, in.e, in.n);
    const AssignedInteger valid_assigned = bigint.assign_constant_fresh(valid);
    bigint.assert_equal_fresh(powed, valid_assigned);
    return valid_assigned;
}

void configure_range(RangeChip& range, uint32_t bits_len) {
    std::vector<uint32_t> comp, over;
    RSAChip::compute_range_lens(bits_len / RSAChip::LIMB_WIDTH, &comp, &over);
    range.configure(comp, over);
}

void synth_mod_pow(Assignment& as, const de_circuit_desc& d, const RsaInputs& in) {
    as.init(d.k, true);
    MainGate gate(as);
    RangeChip range(as, gate);
    configure_range(range, d.bits_len);
    range.load_table();
    RSAChip rsa(gate, range, d.bits_len, d.exp_bits);
    for (const Cell& c : rsa_region(rsa, in)) as.outputs.push_back(c.value);
}

// expected ciphertext cells, the chip's permutation of the initial state, the message, absorb, equality with the expected
void enc_region(MainGate& gate, const Spec& spec, const F& k0, const F& k1, const Vec& message, uint32_t num_input, bool key_as_witness,
                const Cell* key_cells, Assignment& as) {
    const Vec expected = poseidon_encrypt(spec, k0, k1, message);
    std::vector<Cell> expected_cells;
    for (const F& v : expected) expected_cells.push_back(gate.assign_value(v));
    PoseidonChip chip = PoseidonChip::new_enc(gate, spec, k0, k1, key_as_witness);
    if (key_cells) {
        gate.assert_equal(chip.state[2], key_cells[0]);
        gate.assert_equal(chip.state[3], key_cells[1]);
    }
    chip.permutation({}, false);
    for (uint32_t i = 0; i < num_input && i < message.size(); i++) chip.absorbing.push_back(gate.assign_value(message[i]));
    const std::vector<Cell> cipher_text = chip.absorb_and_release();
    for (size_t i = 0; i < cipher_text.size(); i++) {
        if (cipher_text[i].value != expected_cells[i].value)
            throw std::runtime_error("the circuit's ciphertext differs from PoseidonCipher::encrypt (the reference's circuit and native cipher "
                                     "agree only where adding the message twice changes nothing, e.g. the all-zero message of its tests and benches)");
        gate.assert_equal(cipher_text[i], expected_cells[i]);
        as.outputs.push_back(cipher_text[i].value);
    }
}

void synth_pose_enc(Assignment& as, const de_circuit_desc& d, const Vec& message) {
    as.init(d.k, false);
    MainGate gate(as);
    const Spec& spec = shared_spec(5, 8, 57);
    enc_region(gate, spec, fr_in(d.key[0]), fr_in(d.key[1]), message, d.message_len, false, nullptr, as);
}

// src/hash/chip.rs:113-193 PoseidonHashCircuit (the circuit of test_example_hash): the native sponge's digest words as witnesses,
// HasherChip over the inputs one update() each, hash(), the RATE words after the capacity word constrained equal to the expected
void synth_poseidon_hash(Assignment& as, const de_circuit_desc& d, const Vec& inputs) {
    as.init(d.k, false);
    MainGate gate(as);
    const Spec& spec = shared_spec(5, 8, 57);
    const uint32_t rate = spec.t - 1;
    Poseidon ref_hasher(spec, Poseidon::hash_state(spec.t));
    ref_hasher.update(inputs);
    const Vec expected = ref_hasher.squeeze(1);
    std::vector<Cell> expected_cells;
    for (uint32_t i = 0; i < rate; i++) expected_cells.push_back(gate.assign_value(expected[spec.t - rate + i]));
    PoseidonChip hasher = PoseidonChip::new_hash(gate, spec);
    for (const F& v : inputs) hasher.absorbing.push_back(gate.assign_value(v));
    const std::vector<Cell> out = hasher.hash();
    for (uint32_t i = 0; i < rate; i++) {
        if (out[spec.t - rate + i].value != expected_cells[i].value) throw std::runtime_error("poseidon_hash: HasherChip and the native sponge disagree");
        gate.assert_equal(out[spec.t - rate + i], expected_cells[i]);
    }
    for (const Cell& c : out) as.outputs.push_back(c.value);
}

// regions "hash mapping from 2048bit" and "poseidon region" of DelayEncryptCircuit (src/lib.rs:240-300)
void delay_enc_tail(MainGate& gate, Assignment& as, const AssignedInteger& rsa_output, const de_circuit_desc& d, const Vec& message) {
    // three limbs per field element in base 2^64, the last element from limbs 30 and 31
    const Spec& spec = shared_spec(5, 8, 57);
    PoseidonChip hasher = PoseidonChip::new_hash(gate, spec);
    const Cell base1 = gate.assign_constant(F::from_big(BigUint::pow2(RSAChip::LIMB_WIDTH)));
    const Cell base2 = gate.mul(base1, base1);
    for (size_t i = 0; i < rsa_output.size() / 3; i++) {
        Cell a_poly = rsa_output[3 * i];
        a_poly = gate.mul_add(rsa_output[3 * i + 1], base1, a_poly);
        a_poly = gate.mul_add(rsa_output[3 * i + 2], base2, a_poly);
        hasher.absorbing.push_back(a_poly);
    }
    hasher.absorbing.push_back(gate.mul_add(rsa_output[31], base1, rsa_output[30]));
    const std::vector<Cell> h = hasher.hash();
    const Cell h_out[2] = {h[1], h[2]};
    as.outputs.push_back(h_out[0].value);
    as.outputs.push_back(h_out[1].value);
    trace_lap("delay_enc: hash region");
    // the hash output is the encryption key
    enc_region(gate, spec, h_out[0].value, h_out[1].value, message, d.message_len, true, h_out, as);
    trace_lap("delay_enc: poseidon region");
}

void synth_delay_enc(Assignment& as, const de_circuit_desc& d, const RsaInputs& in, const Vec& message) {
    as.init(d.k, true);
    MainGate gate(as);
    RangeChip range(as, gate);
    configure_range(range, d.bits_len);
    RSAChip rsa(gate, range, d.bits_len, d.exp_bits);
    trace_lap("delay_enc: columns zeroed");
    // region lengths are remembered per layout: (bits_len, exp_bits, limbs of the constant result)
    const BigUint result = in.n.is_zero() ? BigUint() : big_pow_mod(in.x, in.e, in.n);
    const uint64_t key = ((uint64_t)d.bits_len << 32) | ((uint64_t)constant_limbs(result) << 16) | d.exp_bits;
    const size_t rsa_rows = as.witness_only && as.threads > 1 ? known_rows("delay_enc rsa region", key) : 0;
    if (rsa_rows) {
        // a witness-only pass that knows where the RSA region ends: the two Poseidon regions need its VALUE only (x^e mod n, a
        // plain integer computation), so they are emitted from row rsa_rows on by another thread while the RSA region is written
        if (d.bits_len != 2048) throw std::runtime_error("delay_enc packs exactly 32 limbs (src/lib.rs:248-250): bits_len must be 2048");
        if (in.n.is_zero()) throw std::runtime_error("mul_mod: zero modulus");
        if (rsa_rows > as.usable) throw std::runtime_error("not enough rows: the circuit needs more than 2^" + std::to_string(as.k) + " - 6 usable rows");
        AssignedInteger result_values(32);
        {
            const std::vector<BigUint> limbs = decompose_big(result, 32, RSAChip::LIMB_WIDTH);
            for (size_t i = 0; i < 32; i++) result_values[i].value = F::from_big(limbs[i]);
        }
        RangeTask tail;
        tail.run(as, rsa_rows, as.usable - rsa_rows, [&](Assignment& part) {
            MainGate g(part);
            delay_enc_tail(g, part, result_values, d, message);
        });
        AssignedInteger rsa_output;
        try {
            rsa_output = rsa_region(rsa, in);
        } catch (...) {
            if (tail.worker.joinable()) tail.worker.join();
            throw;
        }
        trace_lap("delay_enc: rsa region");
        tail.join(as, false);
        if (as.offset != rsa_rows) throw std::runtime_error("delay_enc: the RSA region did not end where the Poseidon regions began");
        for (size_t i = 0; i < 32; i++)
            if (rsa_output[i].value != result_values[i].value) throw std::runtime_error("delay_enc: x^e mod n differs between the circuit and the integers");
        as.offset = tail.part.offset;
        for (const Cell& c : rsa_output) as.outputs.push_back(c.value);
        as.outputs.insert(as.outputs.end(), tail.part.outputs.begin(), tail.part.outputs.end());
        return;
    }
    const AssignedInteger rsa_output = rsa_region(rsa, in);
    trace_lap("delay_enc: rsa region");
    if (as.witness_only) known_rows("delay_enc rsa region", key, as.offset);
    range.load_table();
    if (rsa_output.size() != 32) throw std::runtime_error("delay_enc packs exactly 32 limbs (src/lib.rs:248-250): bits_len must be 2048");
    for (const Cell& c : rsa_output) as.outputs.push_back(c.value);
    delay_enc_tail(gate, as, rsa_output, d, message);
}

void synth_rsa_pkcs1(Assignment& as, const de_circuit_desc& d, const RsaInputs& in) {
    // n = modulus, x = signature, e = the fixed public exponent, message[0..4) = the SHA-256 digest as four 64-bit limbs
    as.init(d.k, true);
    MainGate gate(as);
    RangeChip range(as, gate);
    configure_range(range, d.bits_len);
    range.load_table();
    RSAChip rsa(gate, range, d.bits_len, d.exp_bits);
    const uint32_t num_limbs = d.bits_len / RSAChip::LIMB_WIDTH;
    const AssignedInteger n = rsa.bigint.assign_integer(decompose_big(in.n, num_limbs, RSAChip::LIMB_WIDTH));
    const AssignedInteger sig = rsa.bigint.assign_integer(decompose_big(in.x, num_limbs, RSAChip::LIMB_WIDTH));
    if (d.message_len != 4) throw std::runtime_error("rsa_pkcs1: the digest is four 64-bit limbs");
    std::vector<BigUint> digest;
    for (uint32_t i = 0; i < 4; i++) digest.push_back(fr_in(d.message[i]).to_big());
    const AssignedInteger hashed = rsa.bigint.assign_integer(digest);
    as.outputs.push_back(rsa.verify_pkcs1v15_signature(n, in.e, hashed, sig).value);
}

// the reference's big-integer chip test circuits (src/big_integer/chip.rs:2918-3030 "test_square_circuit" and its siblings): a is
// a constant Fresh integer, the chip squares it (mul: one mul_add row per limb product), the expected product is a constant with
// n1 + n1 - 1 limbs and is_equal_muled compares the two through the carry chain.  n = a, x = the expected product.
// Outputs: the accept bit (the reference asserts it; returning it lets a test see a rejection), then the Muled limbs of a * a.
void synth_bigint_square(Assignment& as, const de_circuit_desc& d, const RsaInputs& in) {
    as.init(d.k, true);
    MainGate gate(as);
    RangeChip range(as, gate);
    configure_range(range, d.bits_len);
    range.load_table();
    BigIntChip bigint(gate, range, RSAChip::LIMB_WIDTH, d.bits_len);
    if (in.n.bits() > d.bits_len) throw std::runtime_error("bigint_square: a wider than bits_len");
    const AssignedInteger a = bigint.assign_constant_fresh(in.n);
    const uint32_t n1 = (uint32_t)a.size();
    const AssignedInteger aa = bigint.mul(a, a);                                  // chip.rs:434-440 square = mul(a, a)
    const AssignedInteger want = bigint.assign_constant(in.x, 2 * n1 - 1);        // chip.rs:122-131 assign_constant_muled
    as.outputs.push_back(bigint.is_equal_muled(aa, want, n1, n1).value);
    for (const Cell& c : aa) as.outputs.push_back(c.value);
}

// the remaining operator tests of the reference's big-integer chip in ONE circuit (src/big_integer/chip.rs:1479-2806: TestAdd,
// TestSub / TestOverflowSub, TestMulModEqual, TestPowMod, TestPowModFixedExp, TestFreshEqual, TestLessThan,
// TestLessThanOrEqual, TestInField): x = a, e = b, n = the modulus; a, b < n for the modular operators; the exponent of the two
// pow_mod forms is the low exp_bits bits of b.  Outputs: one record per operator, (limb count, limbs ...), in the order
// add, sub, sub's overflow bit, mul_mod, pow_mod, pow_mod_fixed_exp, is_equal_fresh, is_less_than, is_less_than_or_equal,
// is_less_than(a, n).
void synth_bigint_ops(Assignment& as, const de_circuit_desc& d, const RsaInputs& in) {
    as.init(d.k, true);
    MainGate gate(as);
    RangeChip range(as, gate);
    configure_range(range, d.bits_len);
    range.load_table();
    BigIntChip bigint(gate, range, RSAChip::LIMB_WIDTH, d.bits_len);
    const uint32_t L = bigint.num_limbs;
    if (in.n.bits() > d.bits_len || in.x.bits() > d.bits_len || in.e.bits() > d.bits_len) throw std::runtime_error("bigint_ops: an operand is wider than bits_len");
    if (d.exp_bits == 0 || d.exp_bits > 64) throw std::runtime_error("bigint_ops: exp_bits out of range");
    const AssignedInteger a = bigint.assign_integer(decompose_big(in.x, L, RSAChip::LIMB_WIDTH));
    const AssignedInteger b = bigint.assign_integer(decompose_big(in.e, L, RSAChip::LIMB_WIDTH));
    const AssignedInteger n = bigint.assign_integer(decompose_big(in.n, L, RSAChip::LIMB_WIDTH));
    auto record = [&](const AssignedInteger& v) {
        as.outputs.push_back(F::from_u64(v.size()));
        for (const Cell& c : v) as.outputs.push_back(c.value);
    };
    record(bigint.add(a, b));
    const std::pair<AssignedInteger, Cell> subbed = bigint.sub(a, b);
    record(subbed.first);
    record({subbed.second});
    record(bigint.mul_mod(a, b, n));
    const BigUint e_small = in.e.low_bits(d.exp_bits);
    const AssignedInteger e_limb = bigint.assign_integer({e_small});
    record(bigint.pow_mod(a, e_limb, n, d.exp_bits));
    record(bigint.pow_mod_fixed_exp(a, e_small, n));
    record({bigint.is_equal_fresh(a, b)});
    record({bigint.is_less_than(a, b)});
    record({bigint.is_less_than_or_equal(a, b)});
    record({bigint.is_less_than(a, n)});
}

}  // namespace

extern "C" {

const char* de_frontend_last_error(void) { return g_frontend_error.c_str(); }

static int synthesize_into(const de_circuit_desc* d, de_fr* advice_out, de_assignment** out) {
    if (!d || !out) return DE_ERR_ARG;
    *out = nullptr;
    try {
        if (d->k < 4 || d->k > 24) throw std::runtime_error("k out of range");
        std::unique_ptr<de_assignment> a(new de_assignment());
        a->as.borrowed_advice = (F*)advice_out;
        RsaInputs in;
        Vec message;
        for (uint32_t i = 0; i < d->message_len; i++) message.push_back(fr_in(d->message[i]));
        if (d->kind != DE_CIRCUIT_POSE_ENC && d->kind != DE_CIRCUIT_POSEIDON_HASH) {
            if (!d->n || !d->e || !d->x || d->bits_len == 0 || d->bits_len % 64) throw std::runtime_error("RSA inputs missing or bits_len not a multiple of 64");
            in.n = BigUint::from_bytes_le(d->n, d->n_len);
            in.e = BigUint::from_bytes_le(d->e, d->e_len);
            in.x = BigUint::from_bytes_le(d->x, d->x_len);
        }
        a->as.witness_only = d->witness_only != 0;
        a->as.reuse_borrowed = advice_out && d->witness_only && d->reuse_buffer;
        a->as.threads = d->threads > 64 ? 64 : (d->threads ? d->threads : 1);
        const auto t0 = std::chrono::steady_clock::now();
        trace_lap("pass begins", true);
        switch (d->kind) {
            case DE_CIRCUIT_MOD_POW: synth_mod_pow(a->as, *d, in); break;
            case DE_CIRCUIT_POSE_ENC: synth_pose_enc(a->as, *d, message); break;
            case DE_CIRCUIT_DELAY_ENC: synth_delay_enc(a->as, *d, in, message); break;
            case DE_CIRCUIT_RSA_PKCS1: synth_rsa_pkcs1(a->as, *d, in); break;
            case DE_CIRCUIT_BIGINT_SQUARE: synth_bigint_square(a->as, *d, in); break;
            case DE_CIRCUIT_BIGINT_OPS: synth_bigint_ops(a->as, *d, in); break;
            case DE_CIRCUIT_POSEIDON_HASH: synth_poseidon_hash(a->as, *d, message); break;
            default: throw std::runtime_error("unknown circuit kind");
        }
        trace_lap("circuit emitted");
        a->as.finalize();
        trace_lap("inverses");
        a->synth_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        *out = a.release();
        return DE_OK;
    } catch (const std::exception& e) {
        g_frontend_error = std::string("de_circuit_synthesize: ") + e.what();
        return DE_ERR_ARG;
    }
}

int de_circuit_synthesize(const de_circuit_desc* d, de_assignment** out) { return synthesize_into(d, nullptr, out); }

int de_circuit_witness(const de_circuit_desc* d, de_fr* advice_out, de_assignment_info_t* info) {
    if (!d || !advice_out) return DE_ERR_ARG;
    de_circuit_desc w = *d;
    w.witness_only = 1;
    de_assignment* a = nullptr;
    const int rc = synthesize_into(&w, advice_out, &a);
    if (rc != DE_OK) return rc;
    if (info) de_assignment_info(a, info);
    delete a;
    return DE_OK;
}

void de_assignment_free(de_assignment* a) { delete a; }

int de_assignment_info(const de_assignment* a, de_assignment_info_t* info) {
    if (!a || !info) return DE_ERR_ARG;
    info->k = a->as.k;
    info->n_fixed = a->as.n_fixed;
    info->n_advice = N_ADVICE;
    info->used_rows = (uint64_t)a->as.offset;
    info->n_copies = (uint64_t)a->as.copies.size();
    info->n_outputs = (uint32_t)a->as.outputs.size();
    info->synthesis_ms = a->synth_ms;
    return DE_OK;
}

int de_assignment_fixed(const de_assignment* a, uint32_t column, de_fr* out) {
    if (!a || !out || column >= a->as.n_fixed || a->as.witness_only) return DE_ERR_ARG;
    memcpy(out, a->as.fixed[column], sizeof(de_fr) * a->as.n);
    return DE_OK;
}
int de_assignment_advice(const de_assignment* a, uint32_t column, de_fr* out) {
    if (!a || !out || column >= N_ADVICE) return DE_ERR_ARG;
    memcpy(out, a->as.advice[column], sizeof(de_fr) * a->as.n);
    return DE_OK;
}
int de_assignment_copies(const de_assignment* a, uint32_t* out) {
    if (!a || !out) return DE_ERR_ARG;
    static_assert(sizeof(Copy) == 16, "Copy is four u32");
    memcpy(out, a->as.copies.data(), sizeof(Copy) * a->as.copies.size());
    return DE_OK;
}
int de_assignment_outputs(const de_assignment* a, de_fr* out) {
    if (!a || !out) return DE_ERR_ARG;
    for (size_t i = 0; i < a->as.outputs.size(); i++) fr_out(a->as.outputs[i], &out[i]);
    return DE_OK;
}

// permutation::keygen::Assembly + build_pk's sigma columns: every cell starts as its own cycle, each copy constraint merges
// two cycles (smaller into larger), and sigma[column][row] = delta^(mapped column) * omega^(mapped row).  Columns are the
// permutation's (5 advice + 1 instance); the output is n_columns * n field elements, column-major, lagrange form.
int de_assignment_sigma(const de_assignment* a, const de_fr* omega, const de_fr* delta, uint32_t n_columns, de_fr* out) {
    if (!a || !omega || !delta || !out || n_columns < N_ADVICE || n_columns > 16) return DE_ERR_ARG;
    const size_t n = a->as.n, cells = (size_t)n_columns * n;
    std::vector<uint32_t> mapping(cells), aux(cells), sizes(cells, 1);
    for (size_t i = 0; i < cells; i++) mapping[i] = aux[i] = (uint32_t)i;
    for (const Copy& c : a->as.copies) {
        if (c.lcol >= n_columns || c.rcol >= n_columns) return DE_ERR_ARG;
        const uint32_t l = (uint32_t)(c.lcol * n + c.lrow), r = (uint32_t)(c.rcol * n + c.rrow);
        if (aux[l] == aux[r]) continue;
        uint32_t left_cycle = aux[l], right_cycle = aux[r];
        if (sizes[left_cycle] < sizes[right_cycle]) std::swap(left_cycle, right_cycle);
        sizes[left_cycle] += sizes[right_cycle];
        uint32_t i = right_cycle;
        do {
            aux[i] = left_cycle;
            i = mapping[i];
        } while (i != right_cycle);
        std::swap(mapping[l], mapping[r]);
    }
    std::vector<F> omega_pows(n), delta_pows(n_columns);
    const F w = fr_in(*omega), dl = fr_in(*delta);
    omega_pows[0] = F::one();
    for (size_t i = 1; i < n; i++) omega_pows[i] = omega_pows[i - 1] * w;
    delta_pows[0] = F::one();
    for (uint32_t c = 1; c < n_columns; c++) delta_pows[c] = delta_pows[c - 1] * dl;
    for (size_t i = 0; i < cells; i++) {
        const uint32_t m = mapping[i];
        fr_out(delta_pows[m / n] * omega_pows[m % n], &out[i]);
    }
    return DE_OK;
}

// ---- native Poseidon (known-answer surface) ----
int de_poseidon_permute(uint32_t t, uint32_t r_f, uint32_t r_p, de_fr* state) {
    if (!state || t < 2 || t > 16 || r_f < 2 || (r_f & 1) || r_p < 1 || r_p > 256) return DE_ERR_ARG;
    try {
        const Spec& spec = shared_spec(t, r_f, r_p);
        Vec s;
        for (uint32_t i = 0; i < t; i++) s.push_back(fr_in(state[i]));
        spec.permute(s);
        for (uint32_t i = 0; i < t; i++) fr_out(s[i], &state[i]);
        return DE_OK;
    } catch (const std::exception& e) {
        g_frontend_error = std::string("de_poseidon_permute: ") + e.what();
        return DE_ERR_ARG;
    }
}
int de_poseidon_cipher(int decrypt, const de_fr key[2], const de_fr* in, uint32_t n_in, de_fr* out) {
    if (!key || !in || !out) return DE_ERR_ARG;
    try {
        const Spec& spec = shared_spec(5, 8, 57);
        Vec v;
        for (uint32_t i = 0; i < n_in; i++) v.push_back(fr_in(in[i]));
        if (!decrypt) {
            const Vec c = poseidon_encrypt(spec, fr_in(key[0]), fr_in(key[1]), v);
            for (size_t i = 0; i < c.size(); i++) fr_out(c[i], &out[i]);
            return DE_OK;
        }
        if (n_in != CIPHER_SIZE) return DE_ERR_ARG;
        Vec m;
        if (!poseidon_decrypt(spec, fr_in(key[0]), fr_in(key[1]), v, &m)) {
            g_frontend_error = "de_poseidon_cipher: authentication tag mismatch";
            return DE_ERR_UNSUPPORTED;
        }
        for (size_t i = 0; i < m.size(); i++) fr_out(m[i], &out[i]);
        return DE_OK;
    } catch (const std::exception& e) {
        g_frontend_error = std::string("de_poseidon_cipher: ") + e.what();
        return DE_ERR_ARG;
    }
}

}  // extern "C"
