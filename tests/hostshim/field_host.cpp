// Host build of csrc/field.cuh (PTX wrappers emulated) so the limb schedule of the device Montgomery
// arithmetic can be checked against the oracle without a GPU.  Test-only.
#include <cstddef>
#include "field.cuh"
using namespace de;
template <class F, class Op> static void run(const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n, Op op) {
    for (size_t i = 0; i < n; i++) {
        F x, y;
        for (int k = 0; k < 8; k++) { x.l[k] = a[8 * i + k]; y.l[k] = b[8 * i + k]; }
        F r = op(x, y);
        for (int k = 0; k < 8; k++) o[8 * i + k] = r.l[k];
    }
}
extern "C" {
void h_fr_mul(const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n) { run<Fr>(a, b, o, n, [](Fr x, Fr y) { return mul(x, y); }); }
void h_fr_add(const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n) { run<Fr>(a, b, o, n, [](Fr x, Fr y) { return add(x, y); }); }
void h_fr_sub(const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n) { run<Fr>(a, b, o, n, [](Fr x, Fr y) { return sub(x, y); }); }
void h_fq_mul(const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n) { run<Fq>(a, b, o, n, [](Fq x, Fq y) { return mul(x, y); }); }
void h_fq_add(const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n) { run<Fq>(a, b, o, n, [](Fq x, Fq y) { return add(x, y); }); }
void h_fq_sub(const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n) { run<Fq>(a, b, o, n, [](Fq x, Fq y) { return sub(x, y); }); }
void h_fr_sqr(const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n) { run<Fr>(a, b, o, n, [](Fr x, Fr) { return sqr(x); }); }
void h_fq_sqr(const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n) { run<Fq>(a, b, o, n, [](Fq x, Fq) { return sqr(x); }); }
void h_fr_inv(const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n) { run<Fr>(a, b, o, n, [](Fr x, Fr) { return inv(x); }); }
void h_fq_inv(const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n) { run<Fq>(a, b, o, n, [](Fq x, Fq) { return inv(x); }); }
void h_fr_from_mont(const uint32_t* a, const uint32_t* b, uint32_t* o, size_t n) { run<Fr>(a, b, o, n, [](Fr x, Fr) { return from_mont(x); }); }
}
