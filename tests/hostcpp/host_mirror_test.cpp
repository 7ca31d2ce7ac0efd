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
        // best_fft over two contexts (ranks of the multi-GPU transform; both on device 0 here)
        {
            halo2_b200::Context ctx1(0);
            auto v = rd<de_fr>(d + "/sharded_in.bin");
            auto w13 = rd<de_fr>(d + "/sharded_omega.bin");
            halo2_b200::best_fft_sharded({&ctx, &ctx1}, v, w13[0], 13);
            if (!same(v, rd<de_fr>(d + "/sharded_fft.bin"))) { std::cerr << "best_fft_sharded mismatch\n"; fails++; }
            bool threw2 = false;
            try { v.pop_back(); halo2_b200::best_fft_sharded({&ctx, &ctx1}, v, w13[0], 13); } catch (const std::runtime_error&) { threw2 = true; }
            if (!threw2) { std::cerr << "best_fft_sharded length mismatch not rejected\n"; fails++; }
        }
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
        // arithmetic::eval_polynomial / kate_division
        {
            auto point = rd<de_fr>(d + "/point.bin");
            auto want_eval = rd<de_fr>(d + "/eval.bin");
            auto want_kate = rd<de_fr>(d + "/kate.bin");
            std::vector<de_fr> got_eval = {halo2_b200::eval_polynomial(ctx, scalars, point[0])};
            if (!same(got_eval, want_eval)) { std::cerr << "eval_polynomial mismatch\n"; fails++; }
            if (!same(halo2_b200::kate_division(ctx, scalars, point[0]), want_kate)) { std::cerr << "kate_division mismatch\n"; fails++; }
        }
        // plonk::create_proof through the C++ mirror, from a SERIALISED proving key (the compiled
        // GraphEvaluator, the permutation columns, the query lists and the polynomials as flat files)
        {
            const std::string p = d + "/proof/";
            auto meta = rd<uint32_t>(p + "meta.bin");  // k, n_fixed, n_advice, n_instance, n_perm, chunk_len, blinding, n_intermediates, degree
            const uint32_t pk_k = meta[0], n_fixed = meta[1], n_advice = meta[2], n_perm = meta[4];
            const size_t pn = size_t(1) << pk_k;
            auto fixed = rd<de_fr>(p + "fixed_coeff.bin"), sigma = rd<de_fr>(p + "sigma_coeff.bin"), advice = rd<de_fr>(p + "advice.bin");
            auto randoms = rd<de_fr>(p + "randoms.bin"), delta = rd<de_fr>(p + "delta.bin"), trepr = rd<de_fr>(p + "transcript_repr.bin");
            auto g = rd<de_g1_affine>(p + "g.bin"), gl = rd<de_g1_affine>(p + "g_lagrange.bin");
            auto perm_kind = rd<uint32_t>(p + "perm_kind.bin"), perm_index = rd<uint32_t>(p + "perm_index.bin");
            auto consts = rd<de_fr>(p + "g_consts.bin");
            auto rots = rd<int32_t>(p + "g_rots.bin");
            auto calcs = rd<de_calculation>(p + "g_calcs.bin");
            auto parts = rd<de_value_source>(p + "g_parts.bin");
            auto aq_col = rd<uint32_t>(p + "aq_col.bin"), fq_col = rd<uint32_t>(p + "fq_col.bin");
            auto aq_rot = rd<int32_t>(p + "aq_rot.bin"), fq_rot = rd<int32_t>(p + "fq_rot.bin");
            auto want = rd<uint8_t>(p + "want_proof.bin");
            // the same polynomials through the RawBytes proving-key FILE (ProvingKey::read): the prover below is staged from it
            auto pk_bytes = rd<uint8_t>(p + "pk.bin");
            auto pkraw = halo2_b200::read_proving_key_raw(pk_bytes, n_perm, /* selectors of the MainGate-only shape */ 0);
            if (pkraw.vk.k != pk_k || pkraw.fixed_polys.size() != n_fixed) { std::cerr << "pk.bin header mismatch\n"; fails++; }
            for (uint32_t i = 0; i < n_fixed; i++)
                if (std::memcmp(pkraw.fixed_polys[i].data(), fixed.data() + i * pn, pn * sizeof(de_fr)) != 0) { std::cerr << "pk.bin fixed poly mismatch\n"; fails++; break; }
            std::vector<const de_fr*> fptr, sptr, aptr;
            for (uint32_t i = 0; i < n_fixed; i++) fptr.push_back(pkraw.fixed_polys[i].data());
            for (uint32_t i = 0; i < n_perm; i++) sptr.push_back(pkraw.polys[i].data());
            for (uint32_t i = 0; i < n_advice; i++) aptr.push_back(advice.data() + i * pn);
            de_pk_desc desc;
            std::memset(&desc, 0, sizeof(desc));
            desc.n_fixed = n_fixed; desc.n_advice = n_advice; desc.n_instance = meta[3];
            desc.fixed_coeff = fptr.data();
            desc.n_perm_columns = n_perm; desc.perm_column_kind = perm_kind.data(); desc.perm_column_index = perm_index.data();
            desc.sigma_coeff = sptr.data();
            desc.chunk_len = meta[5]; desc.blinding_factors = meta[6]; desc.delta = delta[0];
            desc.gates.constants = consts.data(); desc.gates.n_constants = (uint32_t)consts.size();
            desc.gates.rotations = rots.data(); desc.gates.n_rotations = (uint32_t)rots.size();
            desc.gates.calcs = calcs.data(); desc.gates.n_calcs = (uint32_t)calcs.size();
            desc.gates.horner_parts = parts.data(); desc.gates.n_horner_parts = (uint32_t)parts.size();
            desc.gates.n_intermediates = meta[7];
            halo2_b200::ParamsKZG pparams(ctx, pk_k, g.data(), gl.data());
            halo2_b200::EvaluationDomain pdom(ctx, meta[8], pk_k);
            halo2_b200::ProvingKey ppk(ctx, pdom, desc);
            de_prover_desc pd;
            std::memset(&pd, 0, sizeof(pd));
            pd.n_advice_queries = (uint32_t)aq_col.size(); pd.advice_query_column = aq_col.data(); pd.advice_query_rotation = aq_rot.data();
            pd.n_fixed_queries = (uint32_t)fq_col.size(); pd.fixed_query_column = fq_col.data(); pd.fixed_query_rotation = fq_rot.data();
            pd.transcript_repr = trepr[0];
            halo2_b200::Prover prover(ctx, pparams, ppk, pd);
            if (prover.random_count() != randoms.size()) { std::cerr << "random_count mismatch\n"; fails++; }
            auto proof = prover.create_proof(aptr, {std::vector<de_fr>()}, randoms);
            if (!same(proof, want)) { std::cerr << "create_proof bytes mismatch\n"; fails++; }
        }
        std::cout << (fails ? "FAIL" : "OK") << std::endl;
        return fails ? 1 : 0;
    } catch (const std::exception& e) {
        std::cerr << "exception: " << e.what() << std::endl;
        return 3;
    }
}
