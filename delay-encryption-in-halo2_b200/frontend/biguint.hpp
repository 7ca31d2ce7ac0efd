// biguint.hpp — the little arbitrary-precision unsigned integer the circuit front-end needs where the reference uses
// num_bigint::BigUint (witness computation of the big-integer chip: products, quotient / remainder by the RSA modulus,
// limb decomposition; /root/reference/src/big_integer/chip.rs:563-590, src/big_integer/utils.rs:2-17).  64-bit limbs,
// little-endian, always normalised (no leading zero limbs).  Sizes here are <= 4096 bits: schoolbook multiplication and
// schoolbook (Knuth D) division.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

namespace de {
namespace fe {

typedef unsigned __int128 u128;

struct BigUint {
    std::vector<uint64_t> l;

    BigUint() {}
    BigUint(uint64_t v) {
        if (v) l.push_back(v);
    }
    static BigUint from_limbs(const uint64_t* p, size_t n) {
        BigUint r;
        r.l.assign(p, p + n);
        r.trim();
        return r;
    }
    static BigUint from_bytes_le(const uint8_t* p, size_t n) {
        BigUint r;
        r.l.assign((n + 7) / 8, 0);
        for (size_t i = 0; i < n; i++) r.l[i / 8] |= (uint64_t)p[i] << (8 * (i % 8));
        r.trim();
        return r;
    }
    static BigUint pow2(size_t bit) {
        BigUint r;
        r.l.assign(bit / 64 + 1, 0);
        r.l[bit / 64] = 1ull << (bit % 64);
        return r;
    }
    void trim() {
        while (!l.empty() && l.back() == 0) l.pop_back();
    }
    bool is_zero() const { return l.empty(); }
    size_t bits() const {
        if (l.empty()) return 0;
        return 64 * (l.size() - 1) + (64 - (size_t)__builtin_clzll(l.back()));
    }
    bool bit(size_t i) const { return i / 64 < l.size() && ((l[i / 64] >> (i % 64)) & 1); }
    uint64_t limb(size_t i) const { return i < l.size() ? l[i] : 0; }
    uint64_t low_u64() const { return limb(0); }

    static int cmp(const BigUint& a, const BigUint& b) {
        if (a.l.size() != b.l.size()) return a.l.size() < b.l.size() ? -1 : 1;
        for (size_t i = a.l.size(); i-- > 0;) {
            if (a.l[i] != b.l[i]) return a.l[i] < b.l[i] ? -1 : 1;
        }
        return 0;
    }
    bool operator==(const BigUint& o) const { return cmp(*this, o) == 0; }
    bool operator!=(const BigUint& o) const { return cmp(*this, o) != 0; }
    bool operator<(const BigUint& o) const { return cmp(*this, o) < 0; }
    bool operator>=(const BigUint& o) const { return cmp(*this, o) >= 0; }

    BigUint operator+(const BigUint& o) const {
        BigUint r;
        const size_t n = std::max(l.size(), o.l.size());
        r.l.assign(n + 1, 0);
        u128 carry = 0;
        for (size_t i = 0; i < n; i++) {
            u128 v = (u128)limb(i) + o.limb(i) + carry;
            r.l[i] = (uint64_t)v;
            carry = v >> 64;
        }
        r.l[n] = (uint64_t)carry;
        r.trim();
        return r;
    }
    // *this - o; requires *this >= o (the reference's BigUint subtraction panics otherwise: callers check first)
    BigUint operator-(const BigUint& o) const {
        BigUint r;
        r.l.assign(l.size(), 0);
        u128 borrow = 0;
        for (size_t i = 0; i < l.size(); i++) {
            u128 v = (u128)l[i] - o.limb(i) - borrow;
            r.l[i] = (uint64_t)v;
            borrow = (v >> 64) & 1;
        }
        r.trim();
        return r;
    }
    BigUint operator*(const BigUint& o) const {
        BigUint r;
        if (l.empty() || o.l.empty()) return r;
        r.l.assign(l.size() + o.l.size(), 0);
        for (size_t i = 0; i < l.size(); i++) {
            u128 carry = 0;
            for (size_t j = 0; j < o.l.size(); j++) {
                u128 v = (u128)l[i] * o.l[j] + r.l[i + j] + carry;
                r.l[i + j] = (uint64_t)v;
                carry = v >> 64;
            }
            r.l[i + o.l.size()] = (uint64_t)carry;
        }
        r.trim();
        return r;
    }
    BigUint operator<<(size_t s) const {
        BigUint r;
        if (l.empty()) return r;
        const size_t w = s / 64, b = s % 64;
        r.l.assign(l.size() + w + 1, 0);
        for (size_t i = 0; i < l.size(); i++) {
            r.l[i + w] |= l[i] << b;
            if (b) r.l[i + w + 1] |= l[i] >> (64 - b);
        }
        r.trim();
        return r;
    }
    BigUint operator>>(size_t s) const {
        BigUint r;
        const size_t w = s / 64, b = s % 64;
        if (w >= l.size()) return r;
        r.l.assign(l.size() - w, 0);
        for (size_t i = 0; i < r.l.size(); i++) {
            r.l[i] = l[i + w] >> b;
            if (b && i + w + 1 < l.size()) r.l[i] |= l[i + w + 1] << (64 - b);
        }
        r.trim();
        return r;
    }
    // the low `nbits` bits
    BigUint low_bits(size_t nbits) const {
        BigUint r;
        const size_t w = (nbits + 63) / 64;
        r.l.assign(l.begin(), l.begin() + std::min(w, l.size()));
        if (nbits % 64 && r.l.size() == w) r.l[w - 1] &= (1ull << (nbits % 64)) - 1;
        r.trim();
        return r;
    }
    // quotient and remainder; d != 0.  Schoolbook long division in base 2^64 (Knuth, TAOCP vol. 2, 4.3.1, algorithm D).
    static void divmod(const BigUint& a, const BigUint& d, BigUint* q, BigUint* r) {
        if (cmp(a, d) < 0) {
            if (q) *q = BigUint();
            if (r) *r = a;
            return;
        }
        const size_t nd = d.l.size(), na = a.l.size();
        BigUint quo;
        quo.l.assign(na - nd + 1, 0);
        if (nd == 1) {
            u128 rem = 0;
            for (size_t i = na; i-- > 0;) {
                const u128 cur = (rem << 64) | a.l[i];
                quo.l[i] = (uint64_t)(cur / d.l[0]);
                rem = cur % d.l[0];
            }
            quo.trim();
            if (q) *q = quo;
            if (r) *r = BigUint((uint64_t)rem);
            return;
        }
        // D1: normalise so that the divisor's top limb has its high bit set
        const int s = __builtin_clzll(d.l.back());
        std::vector<uint64_t> v(nd), u(na + 1);
        for (size_t i = nd; i-- > 0;) v[i] = (d.l[i] << s) | ((s && i) ? (d.l[i - 1] >> (64 - s)) : 0);
        u[na] = s ? (a.l[na - 1] >> (64 - s)) : 0;
        for (size_t i = na; i-- > 0;) u[i] = (a.l[i] << s) | ((s && i) ? (a.l[i - 1] >> (64 - s)) : 0);
        for (size_t j = na - nd + 1; j-- > 0;) {
            // D3: estimate the quotient digit from the top two limbs, correct it with the third
            const u128 num = ((u128)u[j + nd] << 64) | u[j + nd - 1];
            u128 qhat = num / v[nd - 1], rhat = num % v[nd - 1];
            while ((qhat >> 64) || (uint64_t)qhat * (u128)v[nd - 2] > ((rhat << 64) | u[j + nd - 2])) {
                qhat--;
                rhat += v[nd - 1];
                if (rhat >> 64) break;
            }
            // D4: multiply and subtract
            u128 borrow = 0, carry = 0;
            for (size_t i = 0; i < nd; i++) {
                const u128 p = (uint64_t)qhat * (u128)v[i] + carry;
                carry = p >> 64;
                const u128 t = (u128)u[i + j] - (uint64_t)p - borrow;
                u[i + j] = (uint64_t)t;
                borrow = (t >> 64) & 1;
            }
            const u128 t = (u128)u[j + nd] - (uint64_t)carry - borrow;
            u[j + nd] = (uint64_t)t;
            if ((t >> 64) & 1) {  // D6: the estimate was one too large: add the divisor back
                qhat--;
                u128 c = 0;
                for (size_t i = 0; i < nd; i++) {
                    const u128 w = (u128)u[i + j] + v[i] + c;
                    u[i + j] = (uint64_t)w;
                    c = w >> 64;
                }
                u[j + nd] += (uint64_t)c;
            }
            quo.l[j] = (uint64_t)qhat;
        }
        quo.trim();
        if (q) *q = quo;
        if (r) {
            BigUint rem;
            rem.l.assign(nd, 0);
            for (size_t i = 0; i < nd; i++) rem.l[i] = (u[i] >> s) | ((s && i + 1 <= nd) ? (u[i + 1] << (64 - s)) : 0);
            rem.trim();
            *r = rem;
        }
    }
    BigUint operator/(const BigUint& d) const {
        BigUint q;
        divmod(*this, d, &q, nullptr);
        return q;
    }
    BigUint operator%(const BigUint& d) const {
        BigUint r;
        divmod(*this, d, nullptr, &r);
        return r;
    }
    // little-endian bytes, exactly n of them (value must fit)
    void to_bytes_le(uint8_t* out, size_t n) const {
        memset(out, 0, n);
        for (size_t i = 0; i < n && i / 8 < l.size(); i++) out[i] = (uint8_t)(l[i / 8] >> (8 * (i % 8)));
    }
};

// big_integer/utils.rs:2-17 big_pow_mod (square-and-multiply instead of the reference's recursion; same value)
inline BigUint big_pow_mod(const BigUint& a, const BigUint& e, const BigUint& n) {
    BigUint acc(1), base = a % n;
    for (size_t i = 0; i < e.bits(); i++) {
        if (e.bit(i)) acc = (acc * base) % n;
        base = (base * base) % n;
    }
    return acc % n;
}

// maingate::decompose_big(e, number_of_limbs, bit_len): limbs of bit_len bits, least significant first
inline std::vector<BigUint> decompose_big(const BigUint& e, size_t number_of_limbs, size_t bit_len) {
    std::vector<BigUint> out;
    BigUint cur = e;
    for (size_t i = 0; i < number_of_limbs; i++) {
        out.push_back(cur.low_bits(bit_len));
        cur = cur >> bit_len;
    }
    return out;
}

}  // namespace fe
}  // namespace de
