// maingate.hpp — row emitter for the constraint-system shape of de_b200/plonk.py: main_gate_shape(): halo2wrong's MainGate
// (5 advice columns a..e, one degree-3 gate over 9 fixed columns) and the RangeChip's tagged lookup tables (6 more fixed
// columns).  The reference's chips (/root/reference/src/big_integer, src/rsa, src/poseidon/chip.rs, src/hash, src/encryption)
// are written against maingate::{MainGateInstructions, RangeInstructions, RegionCtx}; that crate is not vendored in the
// reference, so the instruction set below follows its published interface (one gate row per arithmetic instruction,
// `compose` / `decompose` in chunks of four terms with the running sum in column e) and emits rows for THIS repository's
// fixed-column layout:
//
//   gate      a sa + b sb + c sc + d sd + e se + a b s_mul_ab + c d s_mul_cd + e(next row) se_next + s_constant = 0
//   lookups   (s_comp tag_comp, s_comp x) in (t_tag, t_value) for x = a, b, c, d;   (s_over tag_over, s_over e) in (t_tag, t_value)
//   equality  every advice column and the instance column (copy constraints)
//
// Values are BN254 Fr elements in Montgomery form (host arithmetic of csrc/transcript.hpp).  Witness generation is
// sequential big-integer work on a few 10^4 rows: it runs on the host, once per proof, before the GPU pipeline starts.
#pragma once
#include <stdio.h>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif
#include <stdlib.h>

#include <chrono>

#include <exception>
#include <map>
#include <mutex>
#include <new>
#include <stdexcept>
#include <thread>
#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "../csrc/transcript.hpp"
#include "biguint.hpp"

namespace de {
namespace fe {

using host::HFr;

// canonical (non-Montgomery) 256-bit integer: what fe_to_big yields, without heap traffic on the hot paths
struct U256 {
    uint64_t l[4];
    size_t bits() const {
        for (int i = 3; i >= 0; i--)
            if (l[i]) return 64 * (size_t)i + (64 - (size_t)__builtin_clzll(l[i]));
        return 0;
    }
    // bits [lo, lo + n), n <= 64
    uint64_t extract(size_t lo, size_t n) const {
        if (lo >= 256 || n == 0) return 0;
        const size_t w = lo / 64, b = lo % 64;
        uint64_t v = l[w] >> b;
        if (b && w + 1 < 4) v |= l[w + 1] << (64 - b);
        return n >= 64 ? v : (v & ((1ull << n) - 1));
    }
    U256 shr(size_t s) const {
        U256 r = {{0, 0, 0, 0}};
        const size_t w = s / 64, b = s % 64;
        for (size_t i = 0; i + w < 4; i++) {
            r.l[i] = l[i + w] >> b;
            if (b && i + w + 1 < 4) r.l[i] |= l[i + w + 1] << (64 - b);
        }
        return r;
    }
};

// ---- Fr on the host ---------------------------------------------------------------------------------------------------
struct F {
    HFr v;
    F() { memset(v.l, 0, sizeof(v.l)); }
    explicit F(const HFr& x) : v(x) {}
    static F zero() { return F(); }
    static F one() { return F(host::fr_one()); }
    static F from_u64(uint64_t x) {
        HFr t = {{x, 0, 0, 0}};
        return F(host::fr_mul(t, host::FR_FIELD.r2));
    }
    // canonical integer < 2^256 -> field element (reduced mod r)
    static F from_big(const BigUint& b) {
        if (b.bits() > 256) throw std::runtime_error("F::from_big: value wider than 256 bits");
        HFr t = {{b.limb(0), b.limb(1), b.limb(2), b.limb(3)}};
        return F(host::fr_mul(t, host::FR_FIELD.r2));  // mont_mul accepts any 256-bit first operand
    }
    BigUint to_big() const {  // maingate::fe_to_big
        HFr c = host::fr_from_mont(v);
        return BigUint::from_limbs(c.l, 4);
    }
    U256 to_u256() const {
        HFr c = host::fr_from_mont(v);
        U256 r;
        memcpy(r.l, c.l, 32);
        return r;
    }
    static F from_u256(const U256& u) {
        HFr t;
        memcpy(t.l, u.l, 32);
        return F(host::fr_mul(t, host::FR_FIELD.r2));
    }
    // 2^bit and the integers below 2^8 as field elements, from tables built on first use
    static const F& pow2(size_t bit) {
        static const std::vector<F> table = [] {
            std::vector<F> t(256);
            t[0] = F::one();
            for (size_t i = 1; i < 256; i++) t[i] = t[i - 1] + t[i - 1];
            return t;
        }();
        return table.at(bit);
    }
    static F small(uint64_t x) {
        static const std::vector<F> table = [] {
            std::vector<F> t(256);
            for (size_t i = 1; i < 256; i++) t[i] = t[i - 1] + F::one();
            return t;
        }();
        return x < 256 ? table[x] : from_u64(x);
    }
    bool is_zero() const { return host::is_zero(v); }
    bool operator==(const F& o) const { return v == o.v; }
    bool operator!=(const F& o) const { return !(v == o.v); }
    F operator+(const F& o) const { return F(host::fr_add(v, o.v)); }
    F operator*(const F& o) const { return F(host::fr_mul(v, o.v)); }
    F operator-() const {
        if (is_zero()) return *this;
        HFr r;
        host::u128 borrow = 0;
        for (int i = 0; i < 4; i++) {
            host::u128 t = (host::u128)host::FR_FIELD.mod[i] - v.l[i] - borrow;
            r.l[i] = (uint64_t)t;
            borrow = (t >> 64) & 1;
        }
        return F(r);
    }
    F operator-(const F& o) const { return *this + (-o); }
    F invert() const { return F(host::mont_inv(host::FR_FIELD, v)); }  // this != 0
    F pow5() const {
        F t = *this * *this;
        return t * t * *this;
    }
};

// DE_FE_TRACE=1: where a synthesis pass spends its time, on stderr (milliseconds since the latest pass of the process began)
inline void trace_lap(const char* what, bool restart = false) {
    static const bool on = getenv("DE_FE_TRACE") != nullptr;
    if (!on) return;
    static std::chrono::steady_clock::time_point t0;
    if (restart) t0 = std::chrono::steady_clock::now();
    fprintf(stderr, "[de frontend] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
}

// ---- the assignment a synthesis pass produces ---------------------------------------------------------------------------
enum FixedColumn { SA = 0, SB, SC, SD, SE, SE_NEXT, S_MUL_AB, S_MUL_CD, S_CONSTANT, T_TAG, T_VALUE, TAG_COMP, TAG_OVER, S_COMP, S_OVER };
enum { N_ADVICE = 5, N_FIXED_MAIN = 9, N_FIXED_RANGE = 15, INSTANCE_POS = 5 /* position of the instance column in the permutation */ };

struct Cell {
    uint32_t col = 0;  // advice column (permutation position)
    uint32_t row = 0;
    F value;
};

struct Copy {
    uint32_t lcol, lrow, rcol, rrow;
};

struct Assignment {
    uint32_t k = 0, n_fixed = 0;
    size_t n = 0, usable = 0, offset = 0;  // offset: next free row
    // columns are flat arrays of n field elements each: owned (calloc) or, for the advice columns, borrowed from the caller
    // (de_circuit_witness writes straight into the buffer the prover uploads from)
    struct Columns {
        F* base = nullptr;
        size_t n = 0, count = 0;
        bool owned = false;
        F* operator[](size_t c) const { return base + c * n; }
        void alloc(size_t count_, size_t n_) {
            release();
            static_assert(sizeof(F) == 32, "F is four u64");
            base = (F*)calloc(count_ * n_, sizeof(F));  // zero bytes = zero field elements
            if (!base) throw std::bad_alloc();
            n = n_; count = count_; owned = true;
        }
        void borrow(F* p, size_t count_, size_t n_) {
            release();
            base = p; n = n_; count = count_; owned = false;
        }
        void release() {
            if (owned) free(base);
            base = nullptr; owned = false;
        }
        ~Columns() { release(); }
        Columns() {}
        Columns(const Columns&) = delete;
        Columns& operator=(const Columns&) = delete;
    } fixed, advice;
    F* borrowed_advice = nullptr;  // set before init(): 5 * 2^k elements, need not be zeroed
    bool reuse_borrowed = false;   // ... unless they hold an earlier pass of the same circuit: then nothing is zeroed at all
    // a borrowed destination is a pinned staging buffer the GPU reads next: cells are written with non-temporal stores, so the
    // 10 MB do not sit dirty in the caches of a dozen cores when the DMA engine comes for them (measured on the GPU box: a
    // host-to-device copy of 10 MB takes 0.2-0.3 ms from clean memory, 0.85-1.1 ms after 12 threads have just written it)
    bool streaming = false;
    void put(uint32_t column, uint32_t row, const F& v) {
        F* dst = advice[column] + row;
#if defined(__x86_64__)
        if (streaming) {
            const __m128i* src = (const __m128i*)&v;
            _mm_stream_si128((__m128i*)dst, _mm_loadu_si128(src));
            _mm_stream_si128((__m128i*)dst + 1, _mm_loadu_si128(src + 1));
            return;
        }
#endif
        *dst = v;
    }
    static void fence() {
#if defined(__x86_64__)
        _mm_sfence();
#endif
    }
    std::vector<Copy> copies;
    std::vector<F> outputs;  // circuit-level results (ciphertext, RSA result limbs) for the callers' known-answer checks
    // range tables: bit length -> tag
    std::vector<std::pair<uint32_t, uint32_t>> comp_tags, over_tags;
    // create_proof's second call of Circuit::synthesize only collects the advice values (WitnessCollection ignores fixed
    // assignments and copy constraints): witness_only skips both
    bool witness_only = false;
    // witness-only passes may emit row ranges whose length does not depend on the values (every halo2 layout is value
    // independent) from several threads: `threads` > 1 lets BigIntChip::pow_mod do so, each thread through a view()
    uint32_t threads = 1;
    // cells that hold the inverse of a value (is_zero's auxiliary witness): filled by finalize() with ONE batched inversion
    std::vector<std::pair<uint32_t, uint32_t>> pending_inverse;  // (advice column, row); the cell currently holds the value itself

    // 10 MB at k = 16: worth the helpers when the pass has them
    void zero_columns(void* p, size_t bytes) const {
        const size_t parts = threads > 1 && bytes >= ((size_t)4 << 20) ? std::min<size_t>(threads, 4) : 1;
        if (parts == 1) {
            memset(p, 0, bytes);
            return;
        }
        const size_t each = (bytes / parts + 4095) & ~(size_t)4095;
        std::vector<std::thread> helpers;
        for (size_t i = 1; i < parts; i++)
            if (i * each < bytes) helpers.emplace_back([=] { memset((char*)p + i * each, 0, std::min(each, bytes - i * each)); });
        memset(p, 0, std::min(each, bytes));
        for (std::thread& t : helpers) t.join();
    }
    void init(uint32_t k_, bool with_range) {
        k = k_;
        n = (size_t)1 << k;
        usable = n - 6;  // blinding_factors + 1 rows at the end are not usable (cs.blinding_factors() = 5)
        n_fixed = with_range ? N_FIXED_RANGE : N_FIXED_MAIN;
        if (!witness_only) fixed.alloc(n_fixed, n);
        if (borrowed_advice) {
            if (!reuse_borrowed) zero_columns(borrowed_advice, sizeof(F) * N_ADVICE * n);
            advice.borrow(borrowed_advice, N_ADVICE, n);
            streaming = witness_only && ((uintptr_t)borrowed_advice & 15) == 0;
        } else {
            advice.alloc(N_ADVICE, n);
        }
    }
    // a witness-only emitter over the rows [start, start + rows) of parent's advice columns (the columns stay parent's)
    void view(const Assignment& parent, size_t start, size_t rows) {
        k = parent.k; n = parent.n; n_fixed = parent.n_fixed;
        usable = start + rows;  // need_rows() stops a range that runs over its reservation
        offset = start;
        witness_only = true;
        comp_tags = parent.comp_tags;
        over_tags = parent.over_tags;
        advice.borrow(parent.advice.base, N_ADVICE, parent.n);
        streaming = parent.streaming;
    }
    void set_fixed(int column, uint32_t row, const F& v) {
        if (!witness_only) fixed[column][row] = v;
    }
    // Montgomery's trick over the deferred inverses (every pending value is non-zero)
    void finalize() {
        fence();
        const size_t m = pending_inverse.size();
        if (!m) return;
        std::vector<F> pre(m);
        F acc = F::one();
        for (size_t i = 0; i < m; i++) {
            pre[i] = acc;
            acc = acc * advice[pending_inverse[i].first][pending_inverse[i].second];
        }
        F inv = acc.invert();
        for (size_t i = m; i-- > 0;) {
            F& cell = advice[pending_inverse[i].first][pending_inverse[i].second];
            const F value = cell;
            cell = inv * pre[i];
            inv = inv * value;
        }
        pending_inverse.clear();
    }
    void need_rows(size_t rows) const {
        if (offset + rows > usable)
            throw std::runtime_error("not enough rows: the circuit needs more than 2^" + std::to_string(k) + " - 6 usable rows");
    }
    void copy(const Cell& a, const Cell& b) {
        if (a.col == b.col && a.row == b.row) return;
        if (a.value != b.value) throw std::runtime_error("copy constraint between unequal cells (row " + std::to_string(a.row) + " / " + std::to_string(b.row) + ")");
        if (!witness_only) copies.push_back({a.col, a.row, b.col, b.row});
    }
};

// ---- witness-only passes on several threads ----------------------------------------------------------------------------------
// The rows a region takes do not depend on the witness (fixed columns and selectors are laid out at keygen), so a pass that has
// seen a region once knows where everything after it starts.  known_rows() remembers such lengths per (region, shape key) for the
// process; RangeTask emits a region into its reserved rows of the parent's columns on another thread.
inline size_t known_rows(const char* region, uint64_t key, size_t set_to = 0) {
    static std::mutex m;
    static std::map<std::pair<std::string, uint64_t>, size_t> rows;
    std::lock_guard<std::mutex> lock(m);
    size_t& slot = rows[{region, key}];
    if (set_to) slot = set_to;
    return slot;
}
struct RangeTask {
    Assignment part;
    std::thread worker;
    std::exception_ptr error;
    size_t start = 0, reserved = 0;
    template <class Body>
    void run(const Assignment& parent, size_t start_row, size_t rows, Body body) {
        start = start_row;
        reserved = rows;
        part.view(parent, start_row, rows);
        worker = std::thread([this, body] {
            try {
                body(part);
            } catch (...) {
                error = std::current_exception();
            }
            Assignment::fence();  // this thread's non-temporal stores are globally visible before the join
        });
    }
    // after join: rethrows the task's error; exact = the region must have filled its reservation to the row
    void join(Assignment& parent, bool exact) {
        if (worker.joinable()) worker.join();
        if (error) std::rethrow_exception(error);
        if (exact && part.offset != start + reserved) throw std::runtime_error("a region emitted on another thread did not fill its reserved rows");
        parent.pending_inverse.insert(parent.pending_inverse.end(), part.pending_inverse.begin(), part.pending_inverse.end());
    }
    ~RangeTask() {
        if (worker.joinable()) worker.join();
    }
};

// maingate::Term
struct Term {
    enum Kind { ZERO, ASSIGNED, UNASSIGNED } kind = ZERO;
    uint32_t src_col = 0, src_row = 0;  // ASSIGNED: the source cell (copy-constrained to the row's cell)
    F value;                            // value placed in the row
    F base;                             // coefficient in the linear part
    static Term zero() { return Term(); }
    static Term assigned(const Cell& c, const F& base) {
        Term t;
        t.kind = ASSIGNED; t.src_col = c.col; t.src_row = c.row; t.value = c.value; t.base = base;
        return t;
    }
    static Term unassigned(const F& v, const F& base) {
        Term t;
        t.kind = UNASSIGNED; t.value = v; t.base = base;
        return t;
    }
};

// maingate::CombinationOptionCommon (the variants the reference's chips reach)
struct Combination {
    F mul_ab, mul_cd, next;  // s_mul_ab, s_mul_cd, se_next
    static Combination add() { return Combination(); }
    static Combination mul() {
        Combination c;
        c.mul_ab = F::one();
        return c;
    }
    static Combination add_to_next(const F& n) {
        Combination c;
        c.next = n;
        return c;
    }
};

class MainGate {
  public:
    explicit MainGate(Assignment& a) : as(a), one(F::one()), minus_one(-F::one()) {}
    Assignment& as;
    const F one, minus_one;

    // create_proof's pass (witness_only) needs the VALUES of a row and nothing else - no selectors, no copy constraints: the hot
    // instructions write their up to five cells directly (nv = cells used, the rest of the row stays zero) and return cell `res`
    Cell fast_row(const F* v, int nv, int res) {
        as.need_rows(1);
        const uint32_t row = (uint32_t)as.offset++;
        for (int c = 0; c < nv; c++) as.put((uint32_t)c, row, v[c]);
        Cell out;
        out.col = (uint32_t)res; out.row = row; out.value = v[res];
        return out;
    }

    // one gate row: the five terms go to a..e with their bases as sa..se.  Returns the five cells of the row.
    void apply(const Term (&t)[5], const F& constant, const Combination& opt, Cell (&out)[5]) {
        as.need_rows(1);
        const uint32_t row = (uint32_t)as.offset;
        for (uint32_t c = 0; c < 5; c++) {
            if (t[c].kind != Term::ZERO) as.put(c, row, t[c].value);
            out[c].col = c; out[c].row = row; out[c].value = t[c].value;
            if (t[c].kind == Term::ASSIGNED && !as.witness_only && !(t[c].src_col == c && t[c].src_row == row))
                as.copies.push_back({t[c].src_col, t[c].src_row, c, row});
        }
        if (!as.witness_only) {
            for (uint32_t c = 0; c < 5; c++)
                if (t[c].kind != Term::ZERO) as.fixed[SA + c][row] = t[c].base;
            as.fixed[S_MUL_AB][row] = opt.mul_ab;
            as.fixed[S_MUL_CD][row] = opt.mul_cd;
            as.fixed[SE_NEXT][row] = opt.next;
            as.fixed[S_CONSTANT][row] = constant;
        }
        as.offset++;
    }
    Cell apply1(const Term& a, const Term& b, const Term& c, const Term& d, const Term& e, const F& constant, const Combination& opt,
                int result_col) {
        Term t[5] = {a, b, c, d, e};
        Cell out[5];
        apply(t, constant, opt, out);
        return out[result_col];
    }

    // ---- assignment ----
    Cell assign_value(const F& v) {  // a witness cell, unconstrained
        if (as.witness_only) return fast_row(&v, 1, 0);
        return apply1(Term::unassigned(v, F::zero()), Term::zero(), Term::zero(), Term::zero(), Term::zero(), F::zero(), Combination::add(), 0);
    }
    Cell assign_constant(const F& c) {  // -a + c = 0
        return apply1(Term::unassigned(c, minus_one), Term::zero(), Term::zero(), Term::zero(), Term::zero(), c, Combination::add(), 0);
    }
    Cell assign_bit(const F& bit) {  // a b - c = 0 with a = b = c
        Term t[5] = {Term::unassigned(bit, F::zero()), Term::unassigned(bit, F::zero()), Term::unassigned(bit, minus_one), Term::zero(), Term::zero()};
        Cell out[5];
        apply(t, F::zero(), Combination::mul(), out);
        as.copy(out[0], out[1]);
        as.copy(out[0], out[2]);
        return out[0];
    }
    // ---- arithmetic: one row each ----
    Cell add(const Cell& a, const Cell& b) { return add_with_constant(a, b, F::zero()); }
    Cell add_with_constant(const Cell& a, const Cell& b, const F& k) {  // a + b + k - c = 0
        if (as.witness_only) {
            const F v[3] = {a.value, b.value, a.value + b.value + k};
            return fast_row(v, 3, 2);
        }
        return apply1(Term::assigned(a, one), Term::assigned(b, one), Term::unassigned(a.value + b.value + k, minus_one), Term::zero(),
                      Term::zero(), k, Combination::add(), 2);
    }
    Cell add_constant(const Cell& a, const F& k) {  // a + k - b = 0
        if (as.witness_only) {
            const F v[2] = {a.value, a.value + k};
            return fast_row(v, 2, 1);
        }
        return apply1(Term::assigned(a, one), Term::unassigned(a.value + k, minus_one), Term::zero(), Term::zero(), Term::zero(), k,
                      Combination::add(), 1);
    }
    Cell sub(const Cell& a, const Cell& b) {  // a - b - c = 0
        if (as.witness_only) {
            const F v[3] = {a.value, b.value, a.value - b.value};
            return fast_row(v, 3, 2);
        }
        return apply1(Term::assigned(a, one), Term::assigned(b, minus_one), Term::unassigned(a.value - b.value, minus_one), Term::zero(),
                      Term::zero(), F::zero(), Combination::add(), 2);
    }
    Cell mul(const Cell& a, const Cell& b) {  // a b - c = 0
        if (as.witness_only) {
            const F v[3] = {a.value, b.value, a.value * b.value};
            return fast_row(v, 3, 2);
        }
        return apply1(Term::assigned(a, F::zero()), Term::assigned(b, F::zero()), Term::unassigned(a.value * b.value, minus_one), Term::zero(),
                      Term::zero(), F::zero(), Combination::mul(), 2);
    }
    Cell mul_add(const Cell& a, const Cell& b, const Cell& to_add) {  // a b + c - d = 0
        if (as.witness_only) {
            const F v[4] = {a.value, b.value, to_add.value, a.value * b.value + to_add.value};
            return fast_row(v, 4, 3);
        }
        return apply1(Term::assigned(a, F::zero()), Term::assigned(b, F::zero()), Term::assigned(to_add, one),
                      Term::unassigned(a.value * b.value + to_add.value, minus_one), Term::zero(), F::zero(), Combination::mul(), 3);
    }
    Cell mul_add_constant(const Cell& a, const Cell& b, const F& k) {  // a b + k - c = 0
        if (as.witness_only) {
            const F v[3] = {a.value, b.value, a.value * b.value + k};
            return fast_row(v, 3, 2);
        }
        return apply1(Term::assigned(a, F::zero()), Term::assigned(b, F::zero()), Term::unassigned(a.value * b.value + k, minus_one), Term::zero(),
                      Term::zero(), k, Combination::mul(), 2);
    }
    // ---- booleans ----
    Cell and_(const Cell& a, const Cell& b) { return mul(a, b); }
    Cell not_(const Cell& a) {  // 1 - a - b = 0
        return apply1(Term::assigned(a, minus_one), Term::unassigned(one - a.value, minus_one), Term::zero(), Term::zero(), Term::zero(), one,
                      Combination::add(), 1);
    }
    // r = 1 if a == 0 else 0: r is a bit; a a' + r - 1 = 0; r a = 0
    Cell is_zero(const Cell& a) {
        const bool z = a.value.is_zero();
        Cell r = assign_bit(z ? one : F::zero());
        // a' = 1 / a (1 when a = 0) is needed by no later instruction: the cell takes the value now and its inverse in finalize()
        const Cell inv_cell = apply1(Term::assigned(a, F::zero()), Term::unassigned(z ? one : a.value, F::zero()), Term::assigned(r, one),
                                     Term::zero(), Term::zero(), minus_one, Combination::mul(), 1);
        if (!z) as.pending_inverse.push_back({inv_cell.col, inv_cell.row});
        apply1(Term::assigned(a, F::zero()), Term::assigned(r, F::zero()), Term::zero(), Term::zero(), Term::zero(), F::zero(), Combination::mul(), 0);
        return r;
    }
    Cell is_equal(const Cell& a, const Cell& b) { return is_zero(sub(a, b)); }
    // cond a + (1 - cond) b:  a cond - cond b + b - e = 0   (cells: a, cond, cond, b, result)
    Cell select(const Cell& a, const Cell& b, const Cell& cond) {
        const F res = cond.value * a.value + b.value - cond.value * b.value;
        if (as.witness_only) {
            const F v[5] = {a.value, cond.value, cond.value, b.value, res};
            return fast_row(v, 5, 4);
        }
        Combination opt;
        opt.mul_ab = one;
        opt.mul_cd = minus_one;
        return apply1(Term::assigned(a, F::zero()), Term::assigned(cond, F::zero()), Term::assigned(cond, F::zero()), Term::assigned(b, one),
                      Term::unassigned(res, minus_one), F::zero(), opt, 4);
    }
    // ---- assertions ----
    void assert_equal(const Cell& a, const Cell& b) { as.copy(a, b); }
    void assert_zero(const Cell& a) {
        if (!a.value.is_zero()) throw std::runtime_error("assert_zero on a non-zero cell (row " + std::to_string(a.row) + ")");
        apply1(Term::assigned(a, one), Term::zero(), Term::zero(), Term::zero(), Term::zero(), F::zero(), Combination::add(), 0);
    }
    void assert_one(const Cell& a) {
        if (a.value != one) throw std::runtime_error("assert_one on a cell that is not one (row " + std::to_string(a.row) + ")");
        apply1(Term::assigned(a, one), Term::zero(), Term::zero(), Term::zero(), Term::zero(), minus_one, Combination::add(), 0);
    }
    // ---- compose / decompose: chunks of four terms, running sum in e ----
    // result = constant + sum of value_i * base_i.  Row j holds terms 4j .. 4j+3 in a..d and the remaining sum R_j in e with
    // coefficient -1; every row but the last adds e of the next row: terms_j - R_j + R_(j+1) = 0.  R_0 (the first row's e) is
    // the result.  on_row(row index, is_last) lets the range chip tag the rows it emits.
    // `remaining` (optional): the callers that know the running sums R_j without field multiplications (the range chip: R_j is
    // the value with its low 4 j w bits cleared) pass them; otherwise they are accumulated from the terms.
    template <class OnRow>
    Cell decompose(const std::vector<Term>& terms, const F& constant, OnRow on_row, std::vector<Cell>* term_cells, const F* remaining = nullptr) {
        if (terms.empty()) throw std::runtime_error("decompose: no terms");
        const size_t chunks = (terms.size() + 3) / 4;
        as.need_rows(chunks);
        F sums[64];
        if (!remaining) {
            if (chunks > 64) throw std::runtime_error("decompose: too many terms");
            // suffix sums of the chunks: R_j = sum of chunks j .. last (+ the constant in R_0)
            F acc = F::zero();
            for (size_t j = chunks; j-- > 0;) {
                for (size_t i = 4 * j; i < 4 * j + 4 && i < terms.size(); i++) acc = acc + terms[i].value * terms[i].base;
                sums[j] = acc;
            }
            sums[0] = sums[0] + constant;
            remaining = sums;
        }
        Cell result;
        for (size_t j = 0; j < chunks; j++) {
            Term row[5];
            for (size_t i = 0; i < 4 && 4 * j + i < terms.size(); i++) row[i] = terms[4 * j + i];
            row[4] = Term::unassigned(remaining[j], minus_one);
            const bool last = j + 1 == chunks;
            on_row((uint32_t)as.offset, last);
            Cell out[5];
            apply(row, j == 0 ? constant : F::zero(), last ? Combination::add() : Combination::add_to_next(one), out);
            if (j == 0) result = out[4];
            if (term_cells)
                for (size_t i = 0; i < 4 && 4 * j + i < terms.size(); i++) term_cells->push_back(out[i]);
        }
        return result;
    }
    Cell compose(const std::vector<Term>& terms, const F& constant) {
        return decompose(terms, constant, [](uint32_t, bool) {}, nullptr);
    }
    // little-endian bits of a (each asserted to be a bit), recomposed and tied to a
    std::vector<Cell> to_bits(const Cell& a, size_t number_of_bits) {
        const U256 v = a.value.to_u256();
        if (v.bits() > number_of_bits) throw std::runtime_error("to_bits: the value does not fit");
        std::vector<Cell> bits;
        std::vector<Term> terms;
        for (size_t i = 0; i < number_of_bits; i++) {
            bits.push_back(assign_bit(v.extract(i, 1) ? one : F::zero()));
            terms.push_back(Term::assigned(bits.back(), F::pow2(i)));
        }
        assert_equal(compose(terms, F::zero()), a);
        return bits;
    }
};

// maingate::RangeChip: range checks by decomposition into table-checked limbs.
class RangeChip {
  public:
    RangeChip(Assignment& a, MainGate& g) : as(a), gate(g) {}
    Assignment& as;
    MainGate& gate;

    // RangeChip::configure(composition_bit_lens, overflow_bit_lens): distinct non-zero lengths, ascending, get tags 1, 2, ...
    void configure(std::vector<uint32_t> comp, std::vector<uint32_t> over) {
        auto uniq = [](std::vector<uint32_t>& v) {
            std::sort(v.begin(), v.end());
            v.erase(std::unique(v.begin(), v.end()), v.end());
            v.erase(std::remove(v.begin(), v.end(), 0u), v.end());
        };
        uniq(comp);
        uniq(over);
        uint32_t tag = 1;
        for (uint32_t b : comp) as.comp_tags.push_back({b, tag++});
        for (uint32_t b : over) as.over_tags.push_back({b, tag++});
    }
    // RangeChip::load_table: (0, 0) and, per tag, the values 0 .. 2^bits - 1 in the fixed columns t_tag / t_value
    void load_table() {
        size_t row = 0;
        auto put = [&](uint32_t tag, uint64_t v) {
            if (row >= as.usable) throw std::runtime_error("range table does not fit the usable rows");
            as.set_fixed(T_TAG, (uint32_t)row, F::small(tag));
            as.set_fixed(T_VALUE, (uint32_t)row, F::small(v));
            row++;
        };
        put(0, 0);
        for (auto& p : as.comp_tags)
            for (uint64_t v = 0; v < (1ull << p.first); v++) put(p.second, v);
        for (auto& p : as.over_tags)
            for (uint64_t v = 0; v < (1ull << p.first); v++) put(p.second, v);
    }
    static uint32_t tag_of(const std::vector<std::pair<uint32_t, uint32_t>>& tags, uint32_t bits, const char* what) {
        for (auto& p : tags)
            if (p.first == bits) return p.second;
        throw std::runtime_error(std::string("RangeChip: no ") + what + " table for " + std::to_string(bits) + " bits");
    }
    // RangeInstructions::assign(value, limb_bit_len, bit_len): value < 2^bit_len, proven by decomposing it into limbs of
    // limb_bit_len bits (a last, narrower limb of bit_len % limb_bit_len bits when that is non-zero).  Every row of the
    // decomposition is looked up under the composition tag; the narrow limb is copied into column e of one more row that is
    // looked up under the overflow tag (this layout's overflow lookup reads e).
    Cell assign(const F& value, uint32_t limb_bit_len, uint32_t bit_len) {
        const uint32_t overflow_len = bit_len % limb_bit_len;
        const uint32_t nlimbs = bit_len / limb_bit_len + (overflow_len ? 1 : 0);
        const U256 big = value.to_u256();
        if (big.bits() > bit_len || limb_bit_len > 16 || bit_len > 255) throw std::runtime_error("RangeChip::assign: value wider than " + std::to_string(bit_len) + " bits");
        std::vector<Term> terms;
        terms.reserve(nlimbs);
        for (uint32_t i = 0; i < nlimbs; i++) terms.push_back(Term::unassigned(F::small(big.extract((size_t)i * limb_bit_len, limb_bit_len)), F::pow2((size_t)i * limb_bit_len)));
        const F ctag = F::small(tag_of(as.comp_tags, limb_bit_len, "composition"));
        // running sums of the decomposition rows: the value with the bits of the earlier rows cleared
        F remaining[64];
        const uint32_t chunks = (nlimbs + 3) / 4;
        if (chunks > 64) throw std::runtime_error("RangeChip::assign: too many limbs");
        remaining[0] = value;
        for (uint32_t j = 1; j < chunks; j++) {
            const size_t cleared = (size_t)4 * j * limb_bit_len;
            U256 m = big.shr(cleared);
            // shift back up
            U256 up = {{0, 0, 0, 0}};
            const size_t w = cleared / 64, b = cleared % 64;
            for (size_t i = 0; i + w < 4; i++) {
                up.l[i + w] |= m.l[i] << b;
                if (b && i + w + 1 < 4) up.l[i + w + 1] |= m.l[i] >> (64 - b);
            }
            remaining[j] = F::from_u256(up);
        }
        std::vector<Cell> cells;
        Cell result = gate.decompose(terms, F::zero(),
                                     [&](uint32_t row, bool) {
                                         as.set_fixed(TAG_COMP, row, ctag);
                                         as.set_fixed(S_COMP, row, F::one());
                                     },
                                     &cells, remaining);
        if (overflow_len) {
            as.need_rows(1);
            const uint32_t row = (uint32_t)as.offset;
            Term t[5] = {Term::zero(), Term::zero(), Term::zero(), Term::zero(), Term::assigned(cells.back(), F::zero())};
            Cell out[5];
            as.set_fixed(TAG_OVER, row, F::small(tag_of(as.over_tags, overflow_len, "overflow")));
            as.set_fixed(S_OVER, row, F::one());
            gate.apply(t, F::zero(), Combination::add(), out);
        }
        return result;
    }
};

}  // namespace fe
}  // namespace de
