// field.cuh — BN254 Fr / Fq arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form (R = 2^256).
//
// Replaces halo2curves::bn256::{Fr,Fq} mul/add/sub/square (reached from the reference
// everywhere, e.g. /root/reference/src/poseidon/spec.rs:30-35).  Memory layout is the halo2curves one: 4 x u64
// little-endian limbs = 8 x u32 little-endian limbs, so host slices are consumed without conversion.
//
// The multiply is a word-serial Montgomery (CIOS) written as two interleaved 32-bit carry chains ("even" and "odd"
// columns) so that every 32x32->64 product lands on a lo/hi register pair without a carry break:
//   per row: acc += a*b_i (2 chains of 8 mad), m = acc[0]*inv, acc += m*p (2 chains of 8 mad), acc >>= 32 (renaming).
// ptxas fuses each mad.lo.cc/madc.hi.cc pair on the same operands into one IMAD.WIDE.U32(.X).
//
// Every PTX instruction goes through a one-line wrapper that has a host emulation, so the exact limb schedule is
// unit-tested on the CPU (tests/test_field_host.py) before it ever runs on a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DE_HD __host__ __device__ __forceinline__
#define DE_D __device__ __forceinline__
#else
#define DE_HD inline
#define DE_D inline
#endif

namespace de {

// ------------------------------------------------------------------------------------------------------------
// PTX carry-chain wrappers (device) with bit-exact host emulation
// ------------------------------------------------------------------------------------------------------------
namespace ptx {
#if defined(__CUDA_ARCH__)
DE_D uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
DE_D uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
DE_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DE_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DE_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DE_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DE_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DE_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DE_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DE_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DE_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DE_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
// Host emulation: one carry/borrow flag, exactly like PTX CC.CF (sub sets CF = borrow).
static thread_local uint32_t CF = 0;
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; CF = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + CF; CF = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + CF; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; CF = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - CF; CF = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - CF; }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)mul_lo(a, b) + c; CF = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)mul_lo(a, b) + c + CF; CF = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)mul_hi(a, b) + c + CF; CF = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return mul_hi(a, b) + c + CF; }
#endif
}  // namespace ptx

// ------------------------------------------------------------------------------------------------------------
// Field parameters.  The limbs are returned from switch-free constexpr functions so that, after unrolling, the
// modulus reaches ptxas as immediates.
// ------------------------------------------------------------------------------------------------------------
struct FrParams {
    static constexpr uint32_t INV = 0xefffffffu;  // -r^-1 mod 2^32
    DE_HD static constexpr uint32_t p(int i) {
        constexpr uint32_t v[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return v[i];
    }
    DE_HD static constexpr uint32_t one(int i) {  // R mod r
        constexpr uint32_t v[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return v[i];
    }
    DE_HD static constexpr uint32_t r2(int i) {  // R^2 mod r
        constexpr uint32_t v[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return v[i];
    }
};
struct FqParams {
    static constexpr uint32_t INV = 0xe4866389u;
    DE_HD static constexpr uint32_t p(int i) {
        constexpr uint32_t v[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return v[i];
    }
    DE_HD static constexpr uint32_t one(int i) {
        constexpr uint32_t v[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return v[i];
    }
    DE_HD static constexpr uint32_t r2(int i) {
        constexpr uint32_t v[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return v[i];
    }
};

// ------------------------------------------------------------------------------------------------------------
// Field element
// ------------------------------------------------------------------------------------------------------------
template <class P>
struct alignas(16) Fp {
    uint32_t l[8];

    DE_HD static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = 0;
        return r;
    }
    DE_HD static Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = P::one(i);
        return r;
    }
    DE_HD static Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = P::r2(i);
        return r;
    }
    DE_HD bool is_zero() const {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) t |= l[i];
        return t == 0;
    }
    DE_HD bool operator==(const Fp& o) const {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) t |= l[i] ^ o.l[i];
        return t == 0;
    }
    DE_HD bool operator!=(const Fp& o) const { return !(*this == o); }
};

// r = (t >= p) ? t - p : t      (t < 2p, so one conditional subtraction gives the canonical representative)
template <class P>
DE_D void final_sub(uint32_t (&t)[8]) {
    uint32_t s[8];
    s[0] = ptx::sub_cc(t[0], P::p(0));
#pragma unroll
    for (int i = 1; i < 8; i++) s[i] = ptx::subc_cc(t[i], P::p(i));
    uint32_t borrow = ptx::subc(0, 0);  // 0 - 0 - CF: 0xffffffff when t < p
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = borrow ? t[i] : s[i];
}

template <class P>
DE_D Fp<P> add(const Fp<P>& a, const Fp<P>& b) {
    uint32_t t[8];
    t[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) t[i] = ptx::addc_cc(a.l[i], b.l[i]);
    t[7] = ptx::addc(a.l[7], b.l[7]);  // a + b < 2p < 2^255: no carry out
    final_sub<P>(t);
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = t[i];
    return r;
}

template <class P>
DE_D Fp<P> sub(const Fp<P>& a, const Fp<P>& b) {
    uint32_t t[8];
    t[0] = ptx::sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) t[i] = ptx::subc_cc(a.l[i], b.l[i]);
    uint32_t borrow = ptx::subc(0, 0);  // all ones when a < b
    Fp<P> r;
    r.l[0] = ptx::add_cc(t[0], borrow & P::p(0));
#pragma unroll
    for (int i = 1; i < 7; i++) r.l[i] = ptx::addc_cc(t[i], borrow & P::p(i));
    r.l[7] = ptx::addc(t[7], borrow & P::p(7));
    return r;
}

template <class P>
DE_D Fp<P> neg(const Fp<P>& a) {
    return sub(Fp<P>::zero(), a);
}
template <class P>
DE_D Fp<P> dbl(const Fp<P>& a) {
    return add(a, a);
}

// acc[j], acc[j+1] (j even) = lo/hi of a[j] * bi: four independent products, no carries.
DE_D void mul_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        acc[j] = ptx::mul_lo(a[j], bi);
        acc[j + 1] = ptx::mul_hi(a[j], bi);
    }
}
// acc[0..7] += a[0,2,4,6] * bi on lo/hi pairs: one carry chain; the carry out is left in CC.
DE_D void cmad_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
    acc[0] = ptx::mad_lo_cc(a[0], bi, acc[0]);
    acc[1] = ptx::madc_hi_cc(a[0], bi, acc[1]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
        acc[j] = ptx::madc_lo_cc(a[j], bi, acc[j]);
        acc[j + 1] = ptx::madc_hi_cc(a[j], bi, acc[j + 1]);
    }
}
// acc = (acc >> 64) + a[0,2,4,6] * bi, consuming the carry already in CC; the top pair starts from zero.
DE_D void madc_n_rshift(uint32_t* acc, const uint32_t* a, uint32_t bi) {
#pragma unroll
    for (int j = 0; j < 6; j += 2) {
        acc[j] = ptx::madc_lo_cc(a[j], bi, acc[j + 2]);
        acc[j + 1] = ptx::madc_hi_cc(a[j], bi, acc[j + 3]);
    }
    acc[6] = ptx::madc_lo_cc(a[6], bi, 0);
    acc[7] = ptx::madc_hi(a[6], bi, 0);
}

// One CIOS row.  `even` holds columns of weight 2^(32j), `odd` columns of weight 2^(32(j+1)).
// On exit the roles of the two arrays are swapped for the next row (the >>32 is pure renaming).
template <class P>
DE_D void mad_row(uint32_t* even, uint32_t* odd, const uint32_t* a, uint32_t bi, const uint32_t* mod, bool first) {
    if (first) {
        mul_n(odd, a + 1, bi);
        mul_n(even, a, bi);
    } else {
        even[0] = ptx::add_cc(even[0], odd[1]);
        madc_n_rshift(odd, a + 1, bi);
        cmad_n(even, a, bi);
        odd[7] = ptx::addc(odd[7], 0);
    }
    uint32_t mi = ptx::mul_lo(even[0], P::INV);
    cmad_n(odd, mod + 1, mi);
    cmad_n(even, mod, mi);
    odd[7] = ptx::addc(odd[7], 0);
}

template <class P>
DE_D Fp<P> mul(const Fp<P>& a, const Fp<P>& b) {
    uint32_t even[8], odd[8], mod[8];
#pragma unroll
    for (int i = 0; i < 8; i++) mod[i] = P::p(i);
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        mad_row<P>(even, odd, a.l, b.l[i], mod, i == 0);
        mad_row<P>(odd, even, a.l, b.l[i + 1], mod, false);
    }
    // after the last (odd-indexed) row the roles are swapped: T = odd + (even << 32) with odd[0] == 0; result = T >> 32
    uint32_t t[8];
    t[0] = ptx::add_cc(odd[1], even[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) t[i] = ptx::addc_cc(odd[i + 1], even[i]);
    t[7] = ptx::addc(even[7], 0);
    final_sub<P>(t);
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = t[i];
    return r;
}

// Dedicated squaring: 36 + 72 products instead of 64 + 72.  a^2 = sum_i a_i 2^(32i) * (a_i 2^(32i) + 2 * sum_{j>i} a_j 2^(32j)),
// so CIOS row i multiplies a_i with the limbs c_i = a_i, c_(i+1) = a_(i+1) << 1, c_j = (2a)_j for j >= i + 2 only (a < p < 2^254,
// so 2a still has 8 limbs) and the reduction rows stay as in mul().  Relative to the row's base the product a_i * c_j lands on
// columns (j, j + 1) exactly as a * b_i does in mad_row, so the even / odd chains are the same with their leading products
// replaced by carry ripples.  The running sum after row i is below 2^(32i + 288): it fits the 9-limb window like mul()'s.
template <class P, int I>
DE_D void sqr_row(uint32_t* even, uint32_t* odd, const uint32_t* c, uint32_t bi, const uint32_t* mod) {
    if (I == 0) {
        mul_n(odd, c + 1, bi);
        mul_n(even, c, bi);
    } else {
        even[0] = ptx::add_cc(even[0], odd[1]);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (2 * k + 1 >= I) {
                odd[2 * k] = ptx::madc_lo_cc(c[2 * k + 1], bi, odd[2 * k + 2]);
                odd[2 * k + 1] = ptx::madc_hi_cc(c[2 * k + 1], bi, odd[2 * k + 3]);
            } else {
                odd[2 * k] = ptx::addc_cc(odd[2 * k + 2], 0);
                odd[2 * k + 1] = ptx::addc_cc(odd[2 * k + 3], 0);
            }
        }
        odd[6] = ptx::madc_lo_cc(c[7], bi, 0);
        odd[7] = ptx::madc_hi(c[7], bi, 0);
        constexpr int J0 = (I + 1) & ~1;  // first even column with a product in this row
#pragma unroll
        for (int j = J0; j < 8; j += 2) {
            even[j] = j == J0 ? ptx::mad_lo_cc(c[j], bi, even[j]) : ptx::madc_lo_cc(c[j], bi, even[j]);
            even[j + 1] = ptx::madc_hi_cc(c[j], bi, even[j + 1]);
        }
        if (J0 < 8) odd[7] = ptx::addc(odd[7], 0);
    }
    uint32_t mi = ptx::mul_lo(even[0], P::INV);
    cmad_n(odd, mod + 1, mi);
    cmad_n(even, mod, mi);
    odd[7] = ptx::addc(odd[7], 0);
}
template <class P, int I>
DE_D void sqr_row_operands(uint32_t* c, const uint32_t* a, const uint32_t* d) {
#pragma unroll
    for (int j = 0; j < 8; j++) c[j] = j == I ? a[j] : j == I + 1 ? a[j] << 1 : d[j];
}

template <class P>
DE_D Fp<P> sqr(const Fp<P>& a) {
#if defined(DE_SQR_IS_MUL)
    return mul(a, a);
#else
    uint32_t even[8], odd[8], mod[8], d[8], c[8];
    d[0] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) mod[i] = P::p(i);
#pragma unroll
    for (int j = 1; j < 8; j++) d[j] = (a.l[j] << 1) | (a.l[j - 1] >> 31);
    sqr_row_operands<P, 0>(c, a.l, d); sqr_row<P, 0>(even, odd, c, a.l[0], mod);
    sqr_row_operands<P, 1>(c, a.l, d); sqr_row<P, 1>(odd, even, c, a.l[1], mod);
    sqr_row_operands<P, 2>(c, a.l, d); sqr_row<P, 2>(even, odd, c, a.l[2], mod);
    sqr_row_operands<P, 3>(c, a.l, d); sqr_row<P, 3>(odd, even, c, a.l[3], mod);
    sqr_row_operands<P, 4>(c, a.l, d); sqr_row<P, 4>(even, odd, c, a.l[4], mod);
    sqr_row_operands<P, 5>(c, a.l, d); sqr_row<P, 5>(odd, even, c, a.l[5], mod);
    sqr_row_operands<P, 6>(c, a.l, d); sqr_row<P, 6>(even, odd, c, a.l[6], mod);
    sqr_row_operands<P, 7>(c, a.l, d); sqr_row<P, 7>(odd, even, c, a.l[7], mod);
    uint32_t t[8];
    t[0] = ptx::add_cc(odd[1], even[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) t[i] = ptx::addc_cc(odd[i + 1], even[i]);
    t[7] = ptx::addc(even[7], 0);
    final_sub<P>(t);
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = t[i];
    return r;
#endif
}

// Montgomery form -> canonical integer (Fr::to_repr): multiply by 1.
template <class P>
DE_D Fp<P> from_mont(const Fp<P>& a) {
    Fp<P> o = Fp<P>::zero();
    o.l[0] = 1;
    return mul(a, o);
}
template <class P>
DE_D Fp<P> to_mont(const Fp<P>& a) {
    return mul(a, Fp<P>::r2());
}

// a^e for a small public exponent (square-and-multiply, MSB first)
template <class P>
DE_D Fp<P> pow_u64(const Fp<P>& a, uint64_t e) {
    Fp<P> acc = Fp<P>::one();
    for (int i = 63; i >= 0; i--) {
        acc = sqr(acc);
        if ((e >> i) & 1) acc = mul(acc, a);
    }
    return acc;
}

// Field inversion by the binary extended Euclidean algorithm (HAC 14.61 for an odd modulus): only shifts, additions and
// subtractions on 8 limbs, about a third of the issue slots and a quarter of the dependent latency of Fermat's a^(p-2) with
// this multiplier (381 dependent Montgomery products).  Not constant time, which is irrelevant for public proof data.
// Montgomery in, Montgomery out: binary_inv(a R) = a^-1 R^-1, and two more products by R^2 give a^-1 R.  inv(0) = 0
// (ff::Field::invert returns None there; the batch inversions of the prover skip zeros before calling).
namespace bgcd {
DE_D bool is_one(const uint32_t (&a)[8]) {
    uint32_t t = a[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < 8; i++) t |= a[i];
    return t == 0;
}
DE_D void shr1(uint32_t (&a)[8], uint32_t top) {  // a = (top:a) >> 1
#pragma unroll
    for (int i = 0; i < 7; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
    a[7] = (a[7] >> 1) | (top << 31);
}
// a >= b
DE_D bool geq(const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    ptx::sub_cc(a[0], b[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) ptx::subc_cc(a[i], b[i]);
    return ptx::subc(0, 0) == 0;  // no borrow
}
DE_D void sub_in_place(uint32_t (&a)[8], const uint32_t (&b)[8]) {  // a -= b, a >= b
    a[0] = ptx::sub_cc(a[0], b[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) a[i] = ptx::subc_cc(a[i], b[i]);
    a[7] = ptx::subc(a[7], b[7]);
}
// x = x / 2 mod p for x < p
template <class P>
DE_D void half_mod(uint32_t (&x)[8]) {
    uint32_t carry = 0;
    if (x[0] & 1u) {
        x[0] = ptx::add_cc(x[0], P::p(0));
#pragma unroll
        for (int i = 1; i < 8; i++) x[i] = ptx::addc_cc(x[i], P::p(i));
        carry = ptx::addc(0, 0);
    }
    shr1(x, carry);
}
// x = (x - y) mod p for x, y < p
template <class P>
DE_D void sub_mod(uint32_t (&x)[8], const uint32_t (&y)[8]) {
    x[0] = ptx::sub_cc(x[0], y[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) x[i] = ptx::subc_cc(x[i], y[i]);
    const uint32_t borrow = ptx::subc(0, 0);
    x[0] = ptx::add_cc(x[0], borrow & P::p(0));
#pragma unroll
    for (int i = 1; i < 7; i++) x[i] = ptx::addc_cc(x[i], borrow & P::p(i));
    x[7] = ptx::addc(x[7], borrow & P::p(7));
}
}  // namespace bgcd

template <class P>
DE_D Fp<P> inv(const Fp<P>& a) {
    if (a.is_zero()) return a;
    uint32_t u[8], v[8], x1[8], x2[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        u[i] = a.l[i];
        v[i] = P::p(i);
        x1[i] = 0;
        x2[i] = 0;
    }
    x1[0] = 1;
    // invariants: x1 * a = u, x2 * a = v (mod p); gcd(u, v) = 1 throughout, so the loop ends with u == 1 or v == 1
    while (!bgcd::is_one(u) && !bgcd::is_one(v)) {
        while (!(u[0] & 1u)) {
            bgcd::shr1(u, 0);
            bgcd::half_mod<P>(x1);
        }
        while (!(v[0] & 1u)) {
            bgcd::shr1(v, 0);
            bgcd::half_mod<P>(x2);
        }
        if (bgcd::geq(u, v)) {
            bgcd::sub_in_place(u, v);
            bgcd::sub_mod<P>(x1, x2);
        } else {
            bgcd::sub_in_place(v, u);
            bgcd::sub_mod<P>(x2, x1);
        }
    }
    Fp<P> r;
    const bool take1 = bgcd::is_one(u);
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = take1 ? x1[i] : x2[i];
    // r = (a R)^-1 = a^-1 R^-1 as a plain residue; two Montgomery products by R^2 lift it to a^-1 R
    const Fp<P> r2 = Fp<P>::r2();
    return mul(mul(r, r2), r2);
}

// 16-byte vectorised global/shared access
template <class P>
DE_D Fp<P> load(const Fp<P>* p) {
#if defined(__CUDA_ARCH__)
    Fp<P> r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
#else
    return *p;
#endif
}
template <class P>
DE_D void store(Fp<P>* p, const Fp<P>& v) {
#if defined(__CUDA_ARCH__)
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
#else
    *p = v;
#endif
}

#if defined(__CUDACC__)
// One 256-bit access per element (LDG/STG.E.256, sm_100+): a warp touches 1 KiB of whole 128-byte lines per instruction.  Used
// where stores cross NVLink (the multi-GPU transform's exchanges): two 16-byte stores per thread leave every line half written
// per instruction, and a peer cannot merge the halves the way the local L2 does.
template <class P>
__device__ __forceinline__ Fp<P> load256(const Fp<P>* p) {
    Fp<P> r;
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
                 : "l"(p));
    return r;
}
template <class P>
__device__ __forceinline__ void store256(Fp<P>* p, const Fp<P>& v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v.l[0]), "r"(v.l[1]), "r"(v.l[2]), "r"(v.l[3]),
                 "r"(v.l[4]), "r"(v.l[5]), "r"(v.l[6]), "r"(v.l[7])
                 : "memory");
}
#endif

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

}  // namespace de
