// transcript.hpp — host side of the proof transcript: Blake2b (RFC 7693) with halo2's personalisation, the
// Blake2bWrite / Challenge255 framing, and the handful of scalar Fr operations create_proof does on challenges
// (x^n, x * omega^rot, powers of delta).  Pure C++17, no CUDA: unit-tested on the CPU (tests/test_host_cpp.py).
//
// Restates halo2_proofs::transcript::{Blake2bWrite, Challenge255} (tag v2023_04_20; the reference instantiates it at
// /root/reference/benches/delay_enc.rs:120) and halo2curves' Fr::from_uniform_bytes / to_repr, G1Affine::to_bytes
// (SURVEY.md Appendix F).  This is bookkeeping on a few dozen scalars per proof, not a CPU path for the hot arithmetic.
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

namespace de {
namespace host {

typedef unsigned __int128 u128;

// ---- BN254 Fr / Fq, 4 x 64-bit Montgomery limbs (R = 2^256), the in-memory form of halo2curves::bn256::{Fr, Fq} -------
struct HFr {
    uint64_t l[4];
    bool operator==(const HFr& o) const { return memcmp(l, o.l, 32) == 0; }
};
struct HField {
    uint64_t mod[4];
    uint64_t inv;  // -p^-1 mod 2^64
    HFr r, r2;     // R mod p, R^2 mod p
};
static const HField FR_FIELD = {{0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
                                0xc2e1f593efffffffull,
                                {{0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full}},
                                {{0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull}}};
static const HField FQ_FIELD = {{0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
                                0x87d20782e4866389ull,
                                {{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}},
                                {{0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full}}};

inline bool geq_mod(const HField& F, const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > F.mod[i]) return true;
        if (a[i] < F.mod[i]) return false;
    }
    return true;
}
inline void sub_mod(const HField& F, uint64_t a[4]) {
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a[i] - F.mod[i] - borrow;
        a[i] = (uint64_t)t;
        borrow = (t >> 64) & 1;
    }
}
// Montgomery product a * b / R mod p.  Requires b < p; a may be any 256-bit value (used by from_wide).
inline HFr mont_mul(const HField& F, const HFr& a, const HFr& b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 carry = 0;
        for (int j = 0; j < 4; j++) {
            u128 v = (u128)a.l[j] * b.l[i] + t[j] + carry;
            t[j] = (uint64_t)v;
            carry = v >> 64;
        }
        u128 v = (u128)t[4] + carry;
        t[4] = (uint64_t)v;
        t[5] = (uint64_t)(v >> 64);
        uint64_t m = t[0] * F.inv;
        carry = ((u128)m * F.mod[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) {
            u128 w = (u128)m * F.mod[j] + t[j] + carry;
            t[j - 1] = (uint64_t)w;
            carry = w >> 64;
        }
        v = (u128)t[4] + carry;
        t[3] = (uint64_t)v;
        t[4] = t[5] + (uint64_t)(v >> 64);
        t[5] = 0;
    }
    HFr r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq_mod(F, r.l)) sub_mod(F, r.l);
    return r;
}
inline HFr mont_add(const HField& F, const HFr& a, const HFr& b) {
    HFr r;
    u128 carry = 0;
    for (int i = 0; i < 4; i++) {
        u128 v = (u128)a.l[i] + b.l[i] + carry;
        r.l[i] = (uint64_t)v;
        carry = v >> 64;
    }
    if (geq_mod(F, r.l)) sub_mod(F, r.l);  // a + b < 2p < 2^255: no carry out
    return r;
}
inline bool is_zero(const HFr& a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
// a^(p - 2) (Fermat); a != 0
inline HFr mont_inv(const HField& F, const HFr& a) {
    uint64_t e[4] = {F.mod[0] - 2, F.mod[1], F.mod[2], F.mod[3]};  // both moduli end in a limb >= 2
    HFr acc = F.r;
    for (int i = 255; i >= 0; i--) {
        acc = mont_mul(F, acc, acc);
        if ((e[i / 64] >> (i % 64)) & 1) acc = mont_mul(F, acc, a);
    }
    return acc;
}
inline HFr fr_mul(const HFr& a, const HFr& b) { return mont_mul(FR_FIELD, a, b); }
inline HFr fr_add(const HFr& a, const HFr& b) { return mont_add(FR_FIELD, a, b); }
inline HFr fr_one() { return FR_FIELD.r; }
inline HFr fr_pow(HFr base, uint64_t e) {
    HFr acc = fr_one();
    while (e) {
        if (e & 1) acc = fr_mul(acc, base);
        base = fr_mul(base, base);
        e >>= 1;
    }
    return acc;
}
inline HFr fr_from_mont(const HFr& a) {  // canonical integer (Fr::to_repr as limbs)
    HFr one = {{1, 0, 0, 0}};
    return fr_mul(a, one);
}
// Fr::from_uniform_bytes: the 512-bit little-endian integer reduced mod r, returned in Montgomery form
inline HFr fr_from_wide(const uint8_t b[64]) {
    HFr d0, d1;
    memcpy(d0.l, b, 32);
    memcpy(d1.l, b + 32, 32);
    const HFr r3 = fr_mul(FR_FIELD.r2, FR_FIELD.r2);  // R^3
    return fr_add(fr_mul(d0, FR_FIELD.r2), fr_mul(d1, r3));
}

// group::Curve::batch_normalize + Fq::to_repr on the host for the handful of commitments a proof hashes: `count` Jacobian
// points (x, y, z as 12 Montgomery u64 limbs each) -> 64 bytes x || y of canonical little-endian coordinates per point
// (identity -> zeros).  One shared inversion (Montgomery's trick).
inline void g1_jacobian_to_canonical(const uint64_t* jac, size_t count, uint8_t* out_xy) {
    const HField& F = FQ_FIELD;
    std::vector<HFr> pre(count);
    HFr acc = F.r;
    for (size_t i = 0; i < count; i++) {
        HFr z;
        memcpy(z.l, jac + 12 * i + 8, 32);
        pre[i] = acc;
        if (!is_zero(z)) acc = mont_mul(F, acc, z);
    }
    HFr inv = count ? mont_inv(F, acc) : F.r;
    const HFr one = {{1, 0, 0, 0}};
    for (size_t i = count; i-- > 0;) {
        HFr x, y, z;
        memcpy(x.l, jac + 12 * i, 32);
        memcpy(y.l, jac + 12 * i + 4, 32);
        memcpy(z.l, jac + 12 * i + 8, 32);
        if (is_zero(z)) {
            memset(out_xy + 64 * i, 0, 64);
            continue;
        }
        const HFr zi = mont_mul(F, inv, pre[i]);
        inv = mont_mul(F, inv, z);
        const HFr zi2 = mont_mul(F, zi, zi);
        const HFr ax = mont_mul(F, mont_mul(F, x, zi2), one);
        const HFr ay = mont_mul(F, mont_mul(F, mont_mul(F, y, zi2), zi), one);
        memcpy(out_xy + 64 * i, ax.l, 32);
        memcpy(out_xy + 64 * i + 32, ay.l, 32);
    }
}

// ---- Blake2b-512 with a 16-byte personalisation string, no key ---------------------------------------------------------
struct Blake2b {
    uint64_t h[8];
    uint64_t t0, t1;
    uint8_t buf[128];
    size_t buflen;

    static inline uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    static inline uint64_t load64(const uint8_t* p) {
        uint64_t v;
        memcpy(&v, p, 8);
        return v;
    }
    explicit Blake2b(const char personal[16]) {
        static const uint64_t IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                       0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
        for (int i = 0; i < 8; i++) h[i] = IV[i];
        h[0] ^= 0x01010000ull ^ 64ull;  // digest length 64, no key, fanout 1, depth 1
        h[6] ^= load64((const uint8_t*)personal);
        h[7] ^= load64((const uint8_t*)personal + 8);
        t0 = t1 = 0;
        buflen = 0;
        memset(buf, 0, sizeof(buf));
    }
    void compress(const uint8_t block[128], bool last) {
        static const uint64_t IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                       0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
        static const uint8_t SIGMA[12][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
        uint64_t m[16], v[16];
        for (int i = 0; i < 16; i++) m[i] = load64(block + 8 * i);
        for (int i = 0; i < 8; i++) {
            v[i] = h[i];
            v[i + 8] = IV[i];
        }
        v[12] ^= t0;
        v[13] ^= t1;
        if (last) v[14] = ~v[14];
#define DE_B2_G(a, b, c, d, x, y)      \
    v[a] = v[a] + v[b] + (x);          \
    v[d] = rotr(v[d] ^ v[a], 32);      \
    v[c] = v[c] + v[d];                \
    v[b] = rotr(v[b] ^ v[c], 24);      \
    v[a] = v[a] + v[b] + (y);          \
    v[d] = rotr(v[d] ^ v[a], 16);      \
    v[c] = v[c] + v[d];                \
    v[b] = rotr(v[b] ^ v[c], 63);
        for (int r = 0; r < 12; r++) {
            const uint8_t* s = SIGMA[r];
            DE_B2_G(0, 4, 8, 12, m[s[0]], m[s[1]])
            DE_B2_G(1, 5, 9, 13, m[s[2]], m[s[3]])
            DE_B2_G(2, 6, 10, 14, m[s[4]], m[s[5]])
            DE_B2_G(3, 7, 11, 15, m[s[6]], m[s[7]])
            DE_B2_G(0, 5, 10, 15, m[s[8]], m[s[9]])
            DE_B2_G(1, 6, 11, 12, m[s[10]], m[s[11]])
            DE_B2_G(2, 7, 8, 13, m[s[12]], m[s[13]])
            DE_B2_G(3, 4, 9, 14, m[s[14]], m[s[15]])
        }
#undef DE_B2_G
        for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
    }
    void update(const uint8_t* data, size_t len) {
        while (len) {
            if (buflen == 128) {  // the buffer is only flushed when more input follows (the last block is special)
                t0 += 128;
                if (t0 < 128) t1++;
                compress(buf, false);
                buflen = 0;
            }
            size_t take = 128 - buflen;
            if (take > len) take = len;
            memcpy(buf + buflen, data, take);
            buflen += take;
            data += take;
            len -= take;
        }
    }
    // digest of the data so far; the state itself is left untouched (halo2 finalises a clone on every challenge)
    void digest(uint8_t out[64]) const {
        Blake2b c = *this;
        c.t0 += c.buflen;
        if (c.t0 < c.buflen) c.t1++;
        memset(c.buf + c.buflen, 0, 128 - c.buflen);
        c.compress(c.buf, true);
        memcpy(out, c.h, 64);
    }
};

// ---- Blake2bWrite<_, G1Affine, Challenge255<_>> ----------------------------------------------------------------------------
struct TranscriptWriter {
    Blake2b state;
    std::vector<uint8_t> proof;
    TranscriptWriter() : state("Halo2-Transcript") {}

    // challenge as an Fr in Montgomery form
    HFr squeeze_challenge() {
        const uint8_t prefix = 0;
        state.update(&prefix, 1);
        uint8_t d[64];
        state.digest(d);
        return fr_from_wide(d);
    }
    // scalar given as canonical little-endian bytes (Fr::to_repr)
    void common_scalar(const uint8_t repr[32]) {
        const uint8_t prefix = 2;
        state.update(&prefix, 1);
        state.update(repr, 32);
    }
    void write_scalar(const uint8_t repr[32]) {
        common_scalar(repr);
        proof.insert(proof.end(), repr, repr + 32);
    }
    // affine point given as canonical little-endian x || y (64 bytes); false for the identity, which halo2 refuses to hash
    bool write_point(const uint8_t xy[64]) {
        bool zero = true;
        for (int i = 0; i < 64; i++) zero = zero && xy[i] == 0;
        if (zero) return false;
        const uint8_t prefix = 1;
        state.update(&prefix, 1);
        state.update(xy, 64);
        uint8_t c[32];
        memcpy(c, xy, 32);
        c[31] |= (uint8_t)((xy[32] & 1) << 7);  // G1Affine::to_bytes: sign of y in the top bit
        proof.insert(proof.end(), c, c + 32);
        return true;
    }
};

}  // namespace host
}  // namespace de
