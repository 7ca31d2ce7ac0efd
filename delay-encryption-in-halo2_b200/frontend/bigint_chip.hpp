// bigint_chip.hpp — witness generation of the reference's big-integer chip and RSA chip
// (/root/reference/src/big_integer/chip.rs, src/rsa/chip.rs) on the row emitter of maingate.hpp.  Integers are vectors of
// 64-bit limbs held in advice cells: "Fresh" limbs are range-checked below 2^64, "Muled" limbs are sums of limb products
// (below 32 * 2^128 + 2^64 for 2048-bit operands).  Every function names the reference function it follows; the sequence of
// MainGate / RangeChip instructions is the reference's, so the emitted rows carry the same witness values in the same order.
#pragma once
#include <atomic>
#include <exception>
#include <thread>

#include "maingate.hpp"

namespace de {
namespace fe {

typedef std::vector<Cell> AssignedInteger;  // limbs, least significant first

class BigIntChip {
  public:
    BigIntChip(MainGate& g, RangeChip& r, uint32_t limb_width, uint32_t bits_len)
        : gate(g), range(r), limb_width(limb_width), num_limbs(bits_len / limb_width) {
        if (bits_len % limb_width) throw std::runtime_error("BigIntChip: bits_len must be a multiple of limb_width");
    }
    MainGate& gate;
    RangeChip& range;
    uint32_t limb_width, num_limbs;

    static const uint32_t NUM_LOOKUP_LIMBS = 8;  // chip.rs:1167
    static uint32_t sublimb_bit_len(uint32_t bit_len_limb) {  // chip.rs:1351-1359
        const uint32_t v = bit_len_limb / NUM_LOOKUP_LIMBS;
        return v == 0 ? 1 : v;
    }
    static BigUint compute_mul_word_max(uint32_t limb_width, uint32_t min_n) {  // chip.rs:1361-1369
        const BigUint base_m1 = BigUint::pow2(limb_width) - BigUint(1);
        return BigUint(min_n) * base_m1 * base_m1 + base_m1;
    }
    // chip.rs:1215-1253: the table lengths RangeChip::configure needs for this chip
    static void compute_range_lens(uint32_t limb_width, uint32_t num_limbs, std::vector<uint32_t>* comp, std::vector<uint32_t>* over) {
        const uint32_t out_comp = limb_width / NUM_LOOKUP_LIMBS;
        const uint32_t out_over = limb_width % out_comp;
        const BigUint out_base = BigUint::pow2(limb_width);
        const uint32_t fresh_carry_bits = (uint32_t)(out_base + out_base).bits() - limb_width;
        const uint32_t fresh_comp = sublimb_bit_len(fresh_carry_bits);
        const BigUint mul_word_max = compute_mul_word_max(limb_width, num_limbs);
        const uint32_t mul_carry_bits = (uint32_t)(mul_word_max + mul_word_max).bits() - limb_width;
        const uint32_t mul_comp = sublimb_bit_len(mul_carry_bits);
        *comp = {out_comp, fresh_comp, mul_comp};
        *over = {out_over, fresh_carry_bits % fresh_comp, mul_carry_bits % mul_comp};
    }

    BigUint to_big_uint(const AssignedInteger& a) const {  // mod.rs: AssignedInteger::to_big_uint
        BigUint acc;
        for (size_t i = a.size(); i-- > 0;) acc = (acc << limb_width) + a[i].value.to_big();
        return acc;
    }
    Cell range_limb(const F& v) { return range.assign(v, sublimb_bit_len(limb_width), limb_width); }

    // chip.rs:65-85 assign_integer: every limb through the range chip
    AssignedInteger assign_integer(const std::vector<BigUint>& limbs) {
        AssignedInteger out;
        for (const BigUint& l : limbs) out.push_back(range_limb(F::from_big(l)));
        return out;
    }
    // chip.rs:1255-1285 assign_constant
    AssignedInteger assign_constant(const BigUint& integer, uint32_t max_num_limbs) {
        const size_t nb = integer.bits();
        const uint32_t n = (uint32_t)(nb % limb_width == 0 ? nb / limb_width : nb / limb_width + 1);
        if (n > max_num_limbs) throw std::runtime_error("assign_constant: the integer has too many limbs");
        AssignedInteger out;
        for (const BigUint& l : decompose_big(integer, n, limb_width)) out.push_back(gate.assign_constant(F::from_big(l)));
        const Cell zero = gate.assign_constant(F::zero());
        for (uint32_t i = n; i < max_num_limbs; i++) out.push_back(zero);
        return out;
    }
    AssignedInteger assign_constant_fresh(const BigUint& integer) { return assign_constant(integer, num_limbs); }
    // chip.rs:149-165 max_value
    AssignedInteger max_value(uint32_t n) {
        const F limb_max = F::from_big(BigUint::pow2(limb_width) - BigUint(1));
        AssignedInteger out;
        for (uint32_t i = 0; i < n; i++) out.push_back(gate.assign_constant(limb_max));
        return out;
    }
    // chip.rs:1325-1348 div_mod_main_gate: (a / n, a mod n) of field elements read as integers
    std::pair<Cell, Cell> div_mod_main_gate(const Cell& a, const Cell& n) {
        F qf, rf;
        if (n.value == F::pow2(limb_width)) {  // every call site divides by 2^limb_width: shifts instead of a division
            const U256 av = a.value.to_u256();
            qf = F::from_u256(av.shr(limb_width));
            rf = F::from_u64(av.extract(0, limb_width));
        } else {
            BigUint q, r;
            BigUint::divmod(a.value.to_big(), n.value.to_big(), &q, &r);
            qf = F::from_big(q);
            rf = F::from_big(r);
        }
        const Cell qc = gate.assign_value(qf);
        const Cell rc = gate.assign_value(rf);
        const Cell nq = gate.mul(n, qc);
        const Cell a_sub_nq = gate.sub(a, nq);
        gate.assert_equal(rc, a_sub_nq);
        return {qc, rc};
    }
    // chip.rs:260-311 add: limb-wise with range-checked sum and carry; max(n1, n2) + 1 limbs
    AssignedInteger add(AssignedInteger a, AssignedInteger b) {
        const size_t max_n = std::max(a.size(), b.size());
        const Cell zero_value = gate.assign_constant(F::zero());
        a.resize(max_n, zero_value);
        b.resize(max_n, zero_value);
        AssignedInteger c_vals;
        std::vector<Cell> carrys = {zero_value};
        const Cell limb_max_val = gate.assign_constant(F::pow2(limb_width));
        for (size_t i = 0; i < max_n; i++) {
            const Cell a_b = gate.add(a[i], b[i]);
            const Cell sum = gate.add(a_b, carrys[i]);
            const U256 sum_big = sum.value.to_u256();
            const Cell c = range_limb(F::from_u64(sum_big.extract(0, limb_width)));
            const Cell carry = range_limb(F::from_u256(sum_big.shr(limb_width)));
            const Cell c_add_carry = gate.mul_add(carry, limb_max_val, c);
            gate.assert_equal(sum, c_add_carry);
            c_vals.push_back(c);
            carrys.push_back(carry);
        }
        c_vals.push_back(carrys[max_n]);
        return c_vals;
    }
    // chip.rs:1290-1320 sub_unchecked (a >= b): witness c = a - b, assert a = b + c
    AssignedInteger sub_unchecked(const AssignedInteger& a, const AssignedInteger& b) {
        if (a.size() < b.size()) throw std::runtime_error("sub_unchecked: a has fewer limbs than b");
        const BigUint a_big = to_big_uint(a), b_big = to_big_uint(b);
        if (a_big < b_big) throw std::runtime_error("sub_unchecked: a < b");
        BigUint c_big = a_big - b_big;
        AssignedInteger c;
        for (size_t i = 0; i < a.size(); i++) {
            c.push_back(range_limb(F::from_big(c_big.low_bits(limb_width))));
            c_big = c_big >> limb_width;
        }
        const AssignedInteger added = add(b, c);
        assert_equal_fresh(a, added);
        return c;
    }
    // chip.rs:313-376 sub: (|a - b|, is_overflowed) through a + max - b
    std::pair<AssignedInteger, Cell> sub(const AssignedInteger& a, const AssignedInteger& b) {
        const size_t n2 = b.size();
        const AssignedInteger max_int = max_value((uint32_t)n2);
        const AssignedInteger inflated_a = add(a, max_int);
        const AssignedInteger inflated_subed = sub_unchecked(inflated_a, b);
        const Cell one = gate.assign_bit(F::one());
        const Cell is_not_overflowed = gate.is_equal(inflated_subed[n2], one);
        const Cell is_overflowed = gate.not_(is_not_overflowed);
        const size_t num_limbs_l = inflated_subed.size();
        const size_t num_limbs_r = a.size() > n2 ? a.size() : n2;
        const Cell zero_value = gate.assign_constant(F::zero());
        AssignedInteger sel_l, sel_r;
        for (size_t i = 0; i < num_limbs_l; i++)
            sel_l.push_back(i >= n2 ? gate.select(inflated_subed[i], zero_value, is_not_overflowed)
                                    : gate.select(inflated_subed[i], b[i], is_not_overflowed));
        for (size_t i = 0; i < num_limbs_r; i++) {
            if (i >= a.size()) sel_r.push_back(gate.select(max_int[i], zero_value, is_not_overflowed));
            else if (i >= n2) sel_r.push_back(gate.select(zero_value, a[i], is_not_overflowed));
            else sel_r.push_back(gate.select(max_int[i], a[i], is_not_overflowed));
        }
        return {sub_unchecked(sel_l, sel_r), is_overflowed};
    }
    // chip.rs:389-422 mul: limb convolution, one mul_add row per limb product; n1 + n2 - 1 Muled limbs
    AssignedInteger mul(const AssignedInteger& a, const AssignedInteger& b) {
        const size_t d0 = a.size(), d1 = b.size(), d = d0 + d1 - 1;
        AssignedInteger c;
        for (size_t i = 0; i < d; i++) {
            Cell acc = gate.assign_constant(F::zero());
            size_t j = d1 >= i + 1 ? 0 : i + 1 - d1;
            while (j < d0 && j <= i) {
                acc = gate.mul_add(a[j], b[i - j], acc);
                j++;
            }
            c.push_back(acc);
        }
        return c;
    }
    // chip.rs:545-632 mul_mod: witness q, r with a b = q n + r, checked limb-wise with carries
    AssignedInteger mul_mod(const AssignedInteger& a, const AssignedInteger& b, const AssignedInteger& n) {
        const size_t n1 = a.size(), n2 = b.size();
        if (n1 != n.size()) throw std::runtime_error("mul_mod: a and n differ in limb count");
        const BigUint n_big = to_big_uint(n);
        if (n_big.is_zero()) throw std::runtime_error("mul_mod: zero modulus");
        BigUint q_big, prod_big;
        BigUint::divmod(to_big_uint(a) * to_big_uint(b), n_big, &q_big, &prod_big);
        std::vector<BigUint> quotients = decompose_big(q_big, n2, limb_width), prods = decompose_big(prod_big, n1, limb_width);
        if (!(q_big >> (limb_width * n2)).is_zero() || !(prod_big >> (limb_width * n1)).is_zero())
            throw std::runtime_error("mul_mod: quotient or remainder does not fit its limbs");
        const AssignedInteger quotient_int = assign_integer(quotients);
        const AssignedInteger prod_int = assign_integer(prods);
        const AssignedInteger ab = mul(a, b);
        const AssignedInteger qn = mul(quotient_int, n);
        AssignedInteger eq_a, eq_b;
        for (size_t i = 0; i < n1 + n2 - 1; i++) {
            eq_a.push_back(ab[i]);
            eq_b.push_back(i < n1 ? gate.add(qn[i], prod_int[i]) : qn[i]);
        }
        assert_equal_muled(eq_a, eq_b, (uint32_t)n1, (uint32_t)n2);
        return prod_int;
    }
    AssignedInteger square_mod(const AssignedInteger& a, const AssignedInteger& n) { return mul_mod(a, a, n); }
    // chip.rs:667-699 pow_mod: variable exponent, bits taken from the limbs of e
    AssignedInteger pow_mod(const AssignedInteger& a, const AssignedInteger& e, const AssignedInteger& n, uint32_t exp_limb_bits) {
        std::vector<Cell> e_bits;
        for (const Cell& limb : e) {
            const std::vector<Cell> bits = gate.to_bits(limb, exp_limb_bits);
            e_bits.insert(e_bits.end(), bits.begin(), bits.end());
        }
        AssignedInteger acc = assign_constant_fresh(BigUint(1));
        if (gate.as.witness_only && gate.as.threads > 1 && !e_bits.empty() && a.size() == n.size())
            return pow_mod_parallel(a, e_bits, acc, n);
        AssignedInteger squared = a;
        for (const Cell& e_bit : e_bits) {
            const AssignedInteger muled = mul_mod(acc, squared, n);
            for (size_t j = 0; j < acc.size(); j++) acc[j] = gate.select(muled[j], acc[j], e_bit);
            squared = square_mod(squared, n);
        }
        return acc;
    }
    // The loop of pow_mod for a witness-only pass on several threads.  The VALUES of the chain (acc_i, squared_i) are ten
    // modular multiplications of plain integers; with them known, every mul_mod of the loop is an independent emitter of a row
    // range whose length R depends on the limb counts only, so range i can be written by any thread into its reserved rows.
    // The calling thread emits the select rows between the ranges.  The rows are the sequential pass's, cell for cell
    // (tests/test_frontend.py compares the two).
    AssignedInteger value_cells(const BigUint& v, size_t limbs) const {
        AssignedInteger out(limbs);
        const std::vector<BigUint> parts = decompose_big(v, (uint32_t)limbs, limb_width);
        for (size_t i = 0; i < limbs; i++) out[i].value = F::from_big(parts[i]);
        return out;
    }
    AssignedInteger pow_mod_parallel(const AssignedInteger& a, const std::vector<Cell>& e_bits, AssignedInteger acc, const AssignedInteger& n) {
        Assignment& as = gate.as;
        const size_t limbs = n.size();
        const BigUint n_big = to_big_uint(n);
        if (n_big.is_zero()) throw std::runtime_error("mul_mod: zero modulus");
        struct Job {
            AssignedInteger a, b;
            size_t start = 0, end = 0;
            std::vector<std::pair<uint32_t, uint32_t>> pending;
            std::exception_ptr error;
        };
        std::vector<Job> jobs(2 * e_bits.size());
        std::vector<BigUint> muled_values(e_bits.size());
        {
            BigUint acc_v = to_big_uint(acc), sq_v = to_big_uint(a), q;
            for (size_t i = 0; i < e_bits.size(); i++) {
                jobs[2 * i].a = value_cells(acc_v, limbs);
                jobs[2 * i].b = jobs[2 * i + 1].a = jobs[2 * i + 1].b = i == 0 ? a : value_cells(sq_v, limbs);
                BigUint::divmod(acc_v * sq_v, n_big, &q, &muled_values[i]);
                if (!(e_bits[i].value == F::zero())) acc_v = muled_values[i];
                BigUint sq_next;
                BigUint::divmod(sq_v * sq_v, n_big, &q, &sq_next);
                sq_v = sq_next;
            }
        }
        auto run_job = [&](Job& j, size_t rows) {
            try {
                Assignment part;
                part.view(as, j.start, rows);
                MainGate g(part);
                RangeChip r(part, g);
                BigIntChip chip(g, r, limb_width, limb_width * num_limbs);
                chip.mul_mod(j.a, j.b, n);
                j.end = part.offset;
                j.pending.swap(part.pending_inverse);
            } catch (...) {
                j.error = std::current_exception();
            }
            Assignment::fence();  // non-temporal stores of this range visible before the join
        };
        // R: rows of one mul_mod at these limb counts (a property of the layout, not of the values).  The first pass of a
        // process at a given width emits the first range here, in place, and measures it; later passes know it.
        trace_lap("pow_mod: values");
        const uint64_t shape_key = ((uint64_t)limb_width << 32) | limbs;
        const size_t first = as.offset;
        size_t R = known_rows("mul_mod", shape_key);
        const bool first_inline = R == 0;
        if (first_inline) {
            mul_mod(jobs[0].a, jobs[0].b, n);
            R = as.offset - first;
            known_rows("mul_mod", shape_key, R);
            jobs[0].end = as.offset;
        } else {
            as.need_rows(R);
            as.offset += R;
        }
        jobs[0].start = first;
        // the ranges go to the workers while this thread writes the selects between them
        std::atomic<size_t> next_job(first_inline ? 1 : 0);
        std::vector<std::thread> workers;
        const size_t n_workers = std::min<size_t>(as.threads - 1, jobs.size() - (first_inline ? 1 : 0));
        auto drain = [&] {
            for (;;) {
                const size_t j = next_job.fetch_add(1);
                if (j >= jobs.size()) return;
                run_job(jobs[j], R);
            }
        };
        // lay the ranges out first: mul_mod_i | selects_i | square_mod_i | mul_mod_(i+1) ...; the selects are one row per limb
        // in every layout of this front-end, measured on the first iteration like R
        std::exception_ptr main_error;
        size_t select_rows = 0;
        try {
            for (size_t i = 0; i < e_bits.size(); i++) {
                if (i == 1) jobs[2].start = as.offset;
                else if (i > 1 && jobs[2 * i].start != as.offset) throw std::runtime_error("pow_mod: row ranges out of step");
                if (i > 0) {
                    as.need_rows(R);
                    as.offset += R;
                }
                if (i == 1) {
                    // all starts are known from here on: R and the select rows are both measured
                    size_t at = as.offset;
                    for (size_t t = 1; t < e_bits.size(); t++) {
                        at += select_rows;            // selects of iteration t
                        jobs[2 * t + 1].start = at;   // square_mod of iteration t
                        at += R;
                        if (t + 1 < e_bits.size()) jobs[2 * (t + 1)].start = at, at += R;
                    }
                    if (at > as.usable) throw std::runtime_error("not enough rows: the circuit needs more than 2^" + std::to_string(as.k) + " - 6 usable rows");
                    for (size_t w = 0; w < n_workers; w++) workers.emplace_back(drain);
                }
                const size_t before = as.offset;
                const AssignedInteger muled = value_cells(muled_values[i], limbs);
                for (size_t j = 0; j < acc.size(); j++) acc[j] = gate.select(muled[j], acc[j], e_bits[i]);
                if (i == 0) select_rows = as.offset - before;
                else if (as.offset - before != select_rows) throw std::runtime_error("pow_mod: select rows differ between iterations");
                as.need_rows(R);
                if (i == 0) jobs[1].start = as.offset;
                else if (jobs[2 * i + 1].start != as.offset) throw std::runtime_error("pow_mod: row ranges out of step");
                as.offset += R;
            }
            if (e_bits.size() == 1)
                for (size_t w = 0; w < n_workers; w++) workers.emplace_back(drain);
        } catch (...) {
            main_error = std::current_exception();
            next_job.store(jobs.size());
        }
        trace_lap("pow_mod: selects");
        if (!main_error) drain();
        trace_lap("pow_mod: drained");
        for (std::thread& t : workers) t.join();
        trace_lap("pow_mod: joined");
        if (main_error) std::rethrow_exception(main_error);
        for (Job& j : jobs) {
            if (j.error) std::rethrow_exception(j.error);
            if (j.end != j.start + R) throw std::runtime_error("pow_mod: a mul_mod range did not fill its reserved rows");
            as.pending_inverse.insert(as.pending_inverse.end(), j.pending.begin(), j.pending.end());
        }
        return acc;
    }
    // chip.rs:715-747 pow_mod_fixed_exp
    AssignedInteger pow_mod_fixed_exp(const AssignedInteger& a, const BigUint& e, const AssignedInteger& n) {
        AssignedInteger acc = assign_constant(BigUint(1), (uint32_t)a.size());
        AssignedInteger squared = a;
        for (size_t i = 0; i < e.bits(); i++) {
            const AssignedInteger cur_sq = squared;
            squared = square_mod(cur_sq, n);
            if (!e.bit(i)) continue;
            acc = mul_mod(acc, cur_sq, n);
        }
        return acc;
    }
    // chip.rs:786-812 is_equal_fresh
    Cell is_equal_fresh(const AssignedInteger& a, const AssignedInteger& b) {
        const size_t n1 = a.size(), n2 = b.size();
        const bool is_a_larger = n1 > n2;
        const size_t max_n = is_a_larger ? n1 : n2;
        Cell eq_bit = gate.assign_bit(F::one());
        for (size_t i = 0; i < max_n; i++) {
            Cell flag;
            if (is_a_larger && i >= n2) flag = gate.is_zero(a[i]);
            else if (!is_a_larger && i >= n1) flag = gate.is_zero(b[i]);
            else flag = gate.is_equal(a[i], b[i]);
            eq_bit = gate.and_(eq_bit, flag);
        }
        return eq_bit;
    }
    // chip.rs:830-898 is_equal_muled: a - b + word_max carried limb by limb must reproduce the carries of word_max alone
    Cell is_equal_muled(const AssignedInteger& a, const AssignedInteger& b, uint32_t num_limbs_l, uint32_t num_limbs_r) {
        const uint32_t min_n = num_limbs_r >= num_limbs_l ? num_limbs_l : num_limbs_r;
        const BigUint word_max = compute_mul_word_max(limb_width, min_n);
        const uint32_t nl = num_limbs_l + num_limbs_r - 1;
        const uint32_t carry_bits = (uint32_t)(word_max + word_max).bits() - limb_width;
        const F word_max_f = F::from_big(word_max);
        const Cell limb_max = gate.assign_constant(F::pow2(limb_width));
        Cell accumulated_extra = gate.assign_constant(F::zero());
        std::vector<Cell> carry = {gate.assign_constant(F::zero())}, cs;
        Cell eq_bit = gate.assign_bit(F::one());
        for (uint32_t i = 0; i < nl; i++) {
            const Cell a_b = gate.sub(a[i], b[i]);
            const Cell sum = gate.add_with_constant(a_b, carry[i], word_max_f);
            const std::pair<Cell, Cell> qc = div_mod_main_gate(sum, limb_max);
            carry.push_back(qc.first);
            cs.push_back(qc.second);
            accumulated_extra = gate.add_constant(accumulated_extra, word_max_f);
            const std::pair<Cell, Cell> qa = div_mod_main_gate(accumulated_extra, limb_max);
            eq_bit = gate.and_(eq_bit, gate.is_equal(cs[i], qa.second));
            accumulated_extra = qa.first;
            if (i < nl - 1) {
                const Cell range_assigned = range.assign(carry[i + 1].value, sublimb_bit_len(carry_bits), carry_bits);
                eq_bit = gate.and_(eq_bit, gate.is_equal(carry[i + 1], range_assigned));
            } else {
                eq_bit = gate.and_(eq_bit, gate.is_equal(carry[i + 1], accumulated_extra));
            }
        }
        return eq_bit;
    }
    void assert_equal_fresh(const AssignedInteger& a, const AssignedInteger& b) { gate.assert_one(is_equal_fresh(a, b)); }  // chip.rs:1055-1063
    void assert_equal_muled(const AssignedInteger& a, const AssignedInteger& b, uint32_t n1, uint32_t n2) {            // chip.rs:1075-1085
        gate.assert_one(is_equal_muled(a, b, n1, n2));
    }
    // chip.rs:911-944: a < b  =  (a <= b) and not (a == b)
    Cell is_less_than_or_equal(const AssignedInteger& a, const AssignedInteger& b) { return sub(a, b).second; }
    Cell is_less_than(const AssignedInteger& a, const AssignedInteger& b) {
        const Cell is_overflowed = is_less_than_or_equal(a, b);
        const Cell is_eq = is_equal_fresh(a, b);
        return gate.and_(is_overflowed, gate.not_(is_eq));
    }
    void assert_in_field(const AssignedInteger& a, const AssignedInteger& n) { gate.assert_one(is_less_than(a, n)); }  // chip.rs:1153-1161
};

// /root/reference/src/rsa/chip.rs
class RSAChip {
  public:
    static const uint32_t LIMB_WIDTH = 64;  // rsa/chip.rs:217
    RSAChip(MainGate& g, RangeChip& r, uint32_t bits_len, uint32_t exp_limb_bits)
        : gate(g), range(r), bigint(g, r, LIMB_WIDTH, bits_len), bits_len(bits_len), exp_limb_bits(exp_limb_bits) {}
    MainGate& gate;
    RangeChip& range;
    BigIntChip bigint;
    uint32_t bits_len, exp_limb_bits;

    // rsa/chip.rs:252-257 compute_range_lens
    static void compute_range_lens(uint32_t num_limbs, std::vector<uint32_t>* comp, std::vector<uint32_t>* over) {
        BigIntChip::compute_range_lens(LIMB_WIDTH, num_limbs, comp, over);
        comp->push_back(32 / BigIntChip::NUM_LOOKUP_LIMBS);
    }
    // rsa/chip.rs:102-117 modpow_public_key
    AssignedInteger modpow_var(const AssignedInteger& x, const AssignedInteger& n, const AssignedInteger& e) {
        Assignment& as = gate.as;
        const uint64_t key = ((uint64_t)x.size() << 32) | n.size();
        const size_t check_rows = as.witness_only && as.threads > 1 ? known_rows("assert_in_field", key) : 0;
        if (check_rows) {
            // x < n on another thread, into its own rows, while this one (and its helpers) emits the exponentiation
            as.need_rows(check_rows);
            RangeTask check;
            const uint32_t lw = bigint.limb_width, bits = bigint.limb_width * bigint.num_limbs;
            check.run(as, as.offset, check_rows, [&x, &n, lw, bits](Assignment& part) {
                MainGate g(part);
                RangeChip r(part, g);
                BigIntChip(g, r, lw, bits).assert_in_field(x, n);
            });
            as.offset += check_rows;
            try {
                const AssignedInteger powed = bigint.pow_mod(x, e, n, exp_limb_bits);
                check.join(as, true);
                return powed;
            } catch (...) {
                if (check.worker.joinable()) check.worker.join();
                throw;
            }
        }
        const size_t before = as.offset;
        bigint.assert_in_field(x, n);
        if (as.witness_only) known_rows("assert_in_field", key, as.offset - before);
        return bigint.pow_mod(x, e, n, exp_limb_bits);
    }
    AssignedInteger modpow_fixed(const AssignedInteger& x, const AssignedInteger& n, const BigUint& e) {
        bigint.assert_in_field(x, n);
        return bigint.pow_mod_fixed_exp(x, e, n);
    }
    // rsa/chip.rs:119-212 verify_pkcs1v15_signature (SHA-256 prefix, EMSA-PKCS1-v1_5 layout in 64-bit limbs); returns the bit
    Cell verify_pkcs1v15_signature(const AssignedInteger& n, const BigUint& e, const AssignedInteger& hashed_msg, const AssignedInteger& sig) {
        Cell is_eq = gate.assign_constant(F::one());
        const AssignedInteger powed = modpow_fixed(sig, n, e);
        const uint32_t hash_len = 4;
        for (uint32_t i = 0; i < hash_len; i++) is_eq = gate.and_(is_eq, gate.is_equal(powed[i], hashed_msg[i]));
        const Cell prefix_64_1 = gate.assign_constant(F::from_u64(217300885422736416ull));
        const Cell prefix_64_2 = gate.assign_constant(F::from_u64(938447882527703397ull));
        const Cell e1 = gate.is_equal(powed[hash_len], prefix_64_1);
        const Cell e2 = gate.is_equal(powed[hash_len + 1], prefix_64_2);
        is_eq = gate.and_(is_eq, e1);
        is_eq = gate.and_(is_eq, e2);
        const BigUint rem = powed[hash_len + 2].value.to_big();
        const Cell remain_low = range.assign(F::from_big(rem.low_bits(32)), 4, 32);
        const Cell remain_high = range.assign(F::from_big(rem >> 32), 4, 32);
        const Cell u32_assign = gate.assign_constant(F::from_u64(1ull << 32));
        gate.assert_equal(powed[hash_len + 2], gate.mul_add(remain_high, u32_assign, remain_low));
        is_eq = gate.and_(is_eq, gate.is_equal(remain_low, gate.assign_constant(F::from_u64(3158320u))));
        const Cell ff_32 = gate.assign_constant(F::from_u64(4294967295u));
        is_eq = gate.and_(is_eq, gate.is_equal(remain_high, ff_32));
        const Cell ff_64 = gate.assign_constant(F::from_u64(18446744073709551615ull));
        const uint32_t nl = bits_len / LIMB_WIDTH;
        for (uint32_t i = hash_len + 3; i < nl - 1; i++) is_eq = gate.and_(is_eq, gate.is_equal(powed[i], ff_64));
        const Cell last_em = gate.assign_constant(F::from_u64(562949953421311ull));
        is_eq = gate.and_(is_eq, gate.is_equal(powed[nl - 1], last_em));
        return is_eq;
    }
};

}  // namespace fe
}  // namespace de
