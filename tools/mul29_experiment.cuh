// mul29_experiment.cuh — EXPERIMENT, not part of libde_b200.so: a carry-less Montgomery multiplication on 9 x 29-bit limbs
// (64-bit column accumulators, 81 + 81 IMAD.WIDE without carry flags), measured by tools/int_peak against the product's
// word-serial carry-chain multiplier (field.cuh mul).  Result on B200 (profiles/r01_int_peak_v2.jsonl): 38-39 Gmul/s against
// 65.9: a 32x32->64 multiply-add is half-rate in every form (so 162 of them lose to 136), and ptxas splits every `mad.wide`
// with a 64-bit addend into IMAD.WIDE (addend RZ) + a 3-input IADD3 / IADD3.X pair, which loads the ALU pipe with ~290
// instructions per multiplication.  Kept for the record of what was tried.
#pragma once
#include "field.cuh"

namespace de {
namespace ptx {
#if defined(__CUDA_ARCH__)
DE_D uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c) { uint64_t r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }
#else
inline uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c) { return (uint64_t)a * b + c; }
#endif
}  // namespace ptx
struct Fr29 {
    static constexpr uint32_t INV29 = 0x0fffffffu;  // -r^-1 mod 2^29
    DE_HD static constexpr uint32_t p29(int i) {
        constexpr uint32_t v[9] = {0x10000001u, 0x1f0fac9fu, 0x0e5c2450u, 0x07d090f3u, 0x1585d283u, 0x02db40c0u, 0x00a6e141u, 0x0e5c2634u, 0x0030644eu};
        return v[i];
    }
};
// ------------------------------------------------------------------------------------------------------------
// Carry-less Montgomery multiplication on 9 x 29-bit limbs.
//
// Measured on B200 (tools/int_peak): IMAD.WIDE.U32 with a plain 64-bit accumulator issues at the full IMAD rate
// (17.9 T/s), while the carry-chain form the word-serial multiplier above compiles to (IMAD.WIDE.U32.X with predicate
// carries) issues at half of it (9.2 T/s).  With 29-bit limbs a column of the schoolbook product holds at most 18 products
// of 58 bits, which fits a 64-bit accumulator, so no multiply-add needs a carry: 81 + 81 IMAD.WIDE at full rate, and the
// carry / conversion work (shifts, masks, 64-bit adds) goes to the ALU pipe, which runs beside the multiplier pipe.
// Same contract as mul(): 8 x 32-bit Montgomery limbs (R = 2^256) in and out, fully reduced result.  Operand a enters
// shifted left by 5 bits, so that nine 29-bit reduction rounds (2^-261) leave a * b * 2^-256.
// ------------------------------------------------------------------------------------------------------------
template <class P>
DE_HD void to_limbs29(const uint32_t (&w)[8], uint32_t (&l)[9], int shift /* 0 or 5 */) {
    // l = (w << shift) in 29-bit limbs; bit position of limb i in w is 29 i - shift
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const int bit = 29 * i - shift;  // may be negative for i = 0
        uint64_t v;
        if (bit < 0) {
            v = (uint64_t)w[0] << (-bit);
        } else {
            const int word = bit >> 5, off = bit & 31;
            const uint64_t lo = w[word];
            const uint64_t hi = (word + 1 < 8) ? w[word + 1] : 0;
            v = (lo | (hi << 32)) >> off;
        }
        l[i] = (uint32_t)v & 0x1fffffffu;
    }
}

template <class P, bool SQUARE>
DE_D Fp<P> mul29_impl(const Fp<P>& a, const Fp<P>& b) {
    uint32_t al[9], bl[9];
    to_limbs29<P>(a.l, al, 5);
    if (SQUARE) {
        to_limbs29<P>(a.l, bl, 0);
    } else {
        to_limbs29<P>(b.l, bl, 0);
    }
    uint64_t t[18];
#pragma unroll
    for (int k = 0; k < 18; k++) t[k] = 0;
#pragma unroll
    for (int i = 0; i < 9; i++)
#pragma unroll
        for (int j = 0; j < 9; j++) t[i + j] = ptx::mad_wide(al[i], bl[j], t[i + j]);
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const uint32_t m = ((uint32_t)t[i] * Fr29::INV29) & 0x1fffffffu;
#pragma unroll
        for (int j = 0; j < 9; j++) t[i + j] = ptx::mad_wide(m, Fr29::p29(j), t[i + j]);
        t[i + 1] += t[i] >> 29;  // t[i] is now a multiple of 2^29
    }
    // t[9 .. 17] hold the result in unnormalised 29-bit columns: propagate, then repack into 8 x 32 bits
#pragma unroll
    for (int k = 9; k < 17; k++) {
        t[k + 1] += t[k] >> 29;
        t[k] &= 0x1fffffffu;
    }
    uint32_t r[8];
#pragma unroll
    for (int wd = 0; wd < 8; wd++) {
        const int bit = 32 * wd;           // bit position inside the result
        const int k = bit / 29, off = bit % 29;
        uint64_t v = t[9 + k] >> off;      // 29 - off bits
        v |= t[9 + k + 1] << (29 - off);   // next limb (the top one may carry a few extra bits: the value is < 2^255)
        if (29 - off + 29 < 32 && 9 + k + 2 < 18) v |= t[9 + k + 2] << (58 - off);
        r[wd] = (uint32_t)v;
    }
    final_sub<P>(r);
    Fp<P> out;
#pragma unroll
    for (int i = 0; i < 8; i++) out.l[i] = r[i];
    return out;
}
template <class P>
DE_D Fp<P> mul29(const Fp<P>& a, const Fp<P>& b) {
    return mul29_impl<P, false>(a, b);
}


}  // namespace de
