// Exercises host/halo2_b200.hpp against input / expected files written by tests/test_host_cpp.py (which computes the
// expected values with the oracle).  Usage: host_mirror_test <dir>.  Exit code 0 = all comparisons bit-exact.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include "halo2_b200.hpp"

template <class T> static std::vector<T> rd(const std::string& p) {
    std::ifstream f(p, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error("cannot open " + p);
    size_t bytes = f.tellg();
    f.seekg(0);
    std::vector<T> v(bytes / sizeof(T));
    f.read(reinterpret_cast<char*>(v.data()), bytes);
    return v;
}
template <class T> static bool same(const std::vector<T>& a, const std::vector<T>& b) {
    return a.size() == b.size() && std::memcmp(a.data(), b.data(), a.size() * sizeof(T)) == 0;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::string d = argv[1];
    try {
        halo2_b200::Context ctx(0);
        auto scalars = rd<de_fr>(d + "/scalars.bin");
        auto bases = rd<de_g1_affine>(d + "/bases.bin");
        auto omega = rd<de_fr>(d + "/omega.bin");
        auto want_fft = rd<de_fr>(d + "/fft.bin");
        auto want_ext = rd<de_fr>(d + "/ext.bin");
        auto want_msm = rd<de_g1_affine>(d + "/msm_affine.bin");
        uint32_t k = 0;
        while ((size_t(1) << k) < scalars.size()) k++;
        int fails = 0;
        // best_multiexp + batch_normalize
        de_g1 p = halo2_b200::best_multiexp(ctx, scalars, bases);
        if (!same(ctx.batch_normalize({p}), want_msm)) { std::cerr << "best_multiexp mismatch\n"; fails++; }
        // ParamsKZG::commit_lagrange
        halo2_b200::ParamsKZG params(ctx, k, nullptr, bases.data());
        de_g1 c = params.commit_lagrange(scalars.data(), scalars.size());
        if (!same(ctx.batch_normalize({c}), want_msm)) { std::cerr << "commit_lagrange mismatch\n"; fails++; }
        // best_fft
        auto a = scalars;
        halo2_b200::best_fft(ctx, a, omega[0], k);
        if (!same(a, want_fft)) { std::cerr << "best_fft mismatch\n"; fails++; }
        // EvaluationDomain
        halo2_b200::EvaluationDomain dom(ctx, 5, k);
        if (!same(dom.coeff_to_extended(scalars), want_ext)) { std::cerr << "coeff_to_extended mismatch\n"; fails++; }
        auto back = dom.extended_to_coeff(want_ext);
        back.resize(scalars.size());
        if (!same(back, scalars)) { std::cerr << "extended_to_coeff round trip mismatch\n"; fails++; }
        // the reference's asserts
        bool threw = false;
        try { bases.pop_back(); halo2_b200::best_multiexp(ctx, scalars, bases); } catch (const std::runtime_error&) { threw = true; }
        if (!threw) { std::cerr << "length mismatch not rejected\n"; fails++; }
        std::cout << (fails ? "FAIL" : "OK") << std::endl;
        return fails ? 1 : 0;
    } catch (const std::exception& e) {
        std::cerr << "exception: " << e.what() << std::endl;
        return 3;
    }
}
