// Host build of csrc/transcript.hpp for the CPU tests (tests/test_transcript_host.py).  Test-only.
#include <cstddef>
#include "transcript.hpp"
using namespace de::host;
extern "C" {
void h_blake2b_personal(const char* personal16, const uint8_t* data, size_t len, size_t split, uint8_t* out64) {
    Blake2b b(personal16);
    b.update(data, split);  // two updates: exercises the buffering
    b.update(data + split, len - split);
    b.digest(out64);
}
void h_fr_mul4(const uint64_t* a, const uint64_t* b, uint64_t* o) { HFr x, y; memcpy(x.l, a, 32); memcpy(y.l, b, 32); HFr r = fr_mul(x, y); memcpy(o, r.l, 32); }
void h_fr_add4(const uint64_t* a, const uint64_t* b, uint64_t* o) { HFr x, y; memcpy(x.l, a, 32); memcpy(y.l, b, 32); HFr r = fr_add(x, y); memcpy(o, r.l, 32); }
void h_fr_pow4(const uint64_t* a, uint64_t e, uint64_t* o) { HFr x; memcpy(x.l, a, 32); HFr r = fr_pow(x, e); memcpy(o, r.l, 32); }
void h_fr_from_wide(const uint8_t* b64, uint64_t* o) { HFr r = fr_from_wide(b64); memcpy(o, r.l, 32); }
void h_g1_jacobian_to_canonical(const uint64_t* jac, size_t count, uint8_t* out) { g1_jacobian_to_canonical(jac, count, out); }
// script: sequence of ops over a TranscriptWriter: 'p' + 64 bytes (write_point), 's' + 32 bytes (write_scalar),
// 'c' + 32 bytes (common_scalar), 'q' (squeeze: appends the challenge, canonical 32 bytes, to `challenges`)
size_t h_transcript_script(const uint8_t* script, size_t len, uint8_t* proof_out, uint8_t* challenges_out, size_t* n_challenges) {
    TranscriptWriter t;
    size_t pos = 0, nc = 0;
    while (pos < len) {
        uint8_t op = script[pos++];
        if (op == 'p') { t.write_point(script + pos); pos += 64; }
        else if (op == 's') { t.write_scalar(script + pos); pos += 32; }
        else if (op == 'c') { t.common_scalar(script + pos); pos += 32; }
        else if (op == 'q') { HFr c = fr_from_mont(t.squeeze_challenge()); memcpy(challenges_out + 32 * nc, c.l, 32); nc++; }
    }
    memcpy(proof_out, t.proof.data(), t.proof.size());
    *n_challenges = nc;
    return t.proof.size();
}
}
