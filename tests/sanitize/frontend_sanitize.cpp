// Sanitizer driver for the circuit front-end (host C++ only, no GPU): compiled together with frontend/circuits.cpp under
// -fsanitize=thread or -fsanitize=address,undefined by tests/test_frontend_sanitizers.py.  It runs the witness pass the way
// bench.py's throughput arm does - several provers' passes at once, each on several host threads, into dirty and reused
// buffers - and checks every cell against the one-thread pass, so a data race between the row-range workers, the region
// registry or the shared Poseidon parameters shows up as a sanitizer report (non-zero exit) or as a differing cell.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/de_b200.h"

static std::vector<uint8_t> pattern(size_t len, uint8_t seed, bool top_and_odd) {
    std::vector<uint8_t> v(len);
    uint32_t s = 0x9E3779B9u * (seed + 1);
    for (size_t i = 0; i < len; i++) {
        s = s * 1664525u + 1013904223u;
        v[i] = (uint8_t)(s >> 24);
    }
    if (top_and_odd) {
        v[0] |= 1;
        v[len - 1] |= 0x80;
    } else {
        v[len - 1] &= 0x3F;  // below any modulus with its top bit set
    }
    return v;
}

struct Inputs {
    std::vector<uint8_t> n, e, x;
    de_fr message[2];
};

static de_circuit_desc make_desc(const Inputs& in, uint32_t kind, uint32_t k, uint32_t threads, uint32_t reuse) {
    de_circuit_desc d;
    memset(&d, 0, sizeof d);
    d.kind = kind;
    d.k = k;
    d.bits_len = 2048;
    d.exp_bits = 5;
    d.n = in.n.data();
    d.n_len = in.n.size();
    d.e = in.e.data();
    d.e_len = in.e.size();
    d.x = in.x.data();
    d.x_len = in.x.size();
    d.message = in.message;
    d.message_len = 2;
    d.threads = threads;
    d.reuse_buffer = reuse;
    return d;
}

static int witness(const Inputs& in, uint32_t kind, uint32_t k, uint32_t threads, uint32_t reuse, std::vector<de_fr>& out) {
    const de_circuit_desc d = make_desc(in, kind, k, threads, reuse);
    de_assignment_info_t info;
    const int rc = de_circuit_witness(&d, out.data(), &info);
    if (rc != DE_OK) fprintf(stderr, "de_circuit_witness(kind %u, threads %u): %s\n", kind, threads, de_frontend_last_error());
    return rc;
}

static bool same(const std::vector<de_fr>& a, const std::vector<de_fr>& b) { return memcmp(a.data(), b.data(), a.size() * sizeof(de_fr)) == 0; }

int main() {
    Inputs in;
    in.n = pattern(256, 1, true);
    in.x = pattern(256, 2, false);
    in.e = {21};
    memset(in.message, 0, sizeof in.message);
    int bad = 0;
    for (uint32_t kind : {(uint32_t)DE_CIRCUIT_DELAY_ENC, (uint32_t)DE_CIRCUIT_MOD_POW}) {
        const uint32_t k = kind == DE_CIRCUIT_DELAY_ENC ? 16 : 17;
        const size_t cells = (size_t)5 << k;
        std::vector<de_fr> want(cells);
        if (witness(in, kind, k, 1, 0, want) != DE_OK) return 2;
        // several provers at once, each pass on several threads, dirty destination first and a reused one afterwards
        std::vector<std::vector<de_fr>> got(3, std::vector<de_fr>(cells));
        for (auto& g : got) memset(g.data(), 0x5A, cells * sizeof(de_fr));
        std::vector<int> rcs(got.size(), 0);
        for (uint32_t round = 0; round < 2; round++) {
            std::vector<std::thread> provers;
            for (size_t p = 0; p < got.size(); p++)
                provers.emplace_back([&, p] { rcs[p] = witness(in, kind, k, 2 + (uint32_t)p, round, got[p]); });
            for (auto& t : provers) t.join();
            for (size_t p = 0; p < got.size(); p++) {
                if (rcs[p] != DE_OK) return 2;
                if (!same(got[p], want)) {
                    fprintf(stderr, "kind %u round %u prover %zu: cells differ from the one-thread pass\n", kind, round, p);
                    bad++;
                }
            }
        }
    }
    // the keygen pass (fixed columns, copy constraints, sigma) beside a witness pass: allocation / free paths under the sanitizer
    {
        de_circuit_desc d = make_desc(in, DE_CIRCUIT_DELAY_ENC, 16, 1, 0);
        de_assignment* a = nullptr;
        if (de_circuit_synthesize(&d, &a) != DE_OK) {
            fprintf(stderr, "de_circuit_synthesize: %s\n", de_frontend_last_error());
            return 2;
        }
        de_assignment_info_t info;
        de_assignment_info(a, &info);
        std::vector<de_fr> col((size_t)1 << 16), want(5u << 16);
        if (witness(in, DE_CIRCUIT_DELAY_ENC, 16, 4, 0, want) != DE_OK) return 2;
        for (uint32_t c = 0; c < info.n_advice; c++) {
            de_assignment_advice(a, c, col.data());
            if (memcmp(col.data(), want.data() + ((size_t)c << 16), col.size() * sizeof(de_fr))) {
                fprintf(stderr, "advice column %u of the keygen pass differs from the witness pass\n", c);
                bad++;
            }
        }
        std::vector<uint32_t> copies(4 * info.n_copies);
        de_assignment_copies(a, copies.data());
        de_assignment_free(a);
        printf("delay_enc: %llu used rows, %llu copy constraints\n", (unsigned long long)info.used_rows, (unsigned long long)info.n_copies);
    }
    // error paths: a zero modulus and an exponent wider than exp_bits must come back as errors, not as crashes
    {
        Inputs z = in;
        z.n.assign(256, 0);
        std::vector<de_fr> out(5u << 16);
        const de_circuit_desc d = make_desc(z, DE_CIRCUIT_DELAY_ENC, 16, 3, 0);
        if (de_circuit_witness(&d, out.data(), nullptr) == DE_OK) bad++;
        Inputs w = in;
        w.e = {0xFF};
        const de_circuit_desc d2 = make_desc(w, DE_CIRCUIT_MOD_POW, 17, 3, 0);
        std::vector<de_fr> out2(5u << 17);
        if (de_circuit_witness(&d2, out2.data(), nullptr) == DE_OK) bad++;
    }
    printf(bad ? "FAILED (%d)\n" : "ok\n", bad);
    return bad ? 1 : 0;
}
