// msm_reduce.cu — host side of the bucket reduction (msm_reduce.cuh): which of the three reduction layouts runs for a window
// width and context mode, and their launch sequences.  Split from msm.cu so that the two halves of the MSM compile in parallel.
// Reference semantics: the running-sum fold of halo2_proofs::arithmetic::best_multiexp, regrouped.
#include <stdlib.h>

#include "common.cuh"
#include "msm_reduce.cuh"

namespace de {

// buckets: count * nsets sets of NB XYZZ points; red: (ndigits * 32 + 1) * nsets_total points; red2: scratch of the two-digit
// layouts (sized by msm_enqueue); leaves `count` Jacobian points at d_out
int msm_reduce_enqueue(de_ctx* ctx, cudaStream_t st, unsigned int c, unsigned int nsets, unsigned int NB, size_t count, size_t n, XYZZ* buckets,
                       XYZZ* red, XYZZ* red2, Jac* d_out) {
    const unsigned int ndigits = (c - 1 + 4) / 5;
    const unsigned int nsets_total = (unsigned int)(count * nsets);
    XYZZ* dsums = red;
    XYZZ* set_out = red + (size_t)ndigits * 32 * nsets_total;
    static const bool rowcol_always = getenv("DE_REDUCE_ROWCOL") != nullptr;  // A/B switch for measurements
    if (c >= 11 && c <= 16 && (ctx->mode == DE_MODE_LATENCY || rowcol_always)) {
        // latency-oriented two-digit reduction (4c): row / column sums in one launch, bit sums, one warp per set for the fold
        const unsigned int cm1 = c - 1, w0 = (cm1 + 1) / 2, w1 = cm1 - w0;
        const unsigned int V0 = 1u << w0, V1 = 1u << w1;
        XYZZ* D0 = red2;
        XYZZ* D1 = D0 + (size_t)nsets_total * V0;
        XYZZ* S = D1 + (size_t)nsets_total * V1;
        TimedLaunch tl = timing_begin(ctx, "k_msm_digit_sums", (double)n * count);
        k_bucket_rowcol<<<dim3(V1 + V0 / DE_RC_COLS, nsets_total), DE_RC_THREADS, 0, st>>>(buckets, NB, w0, w1, D0, D1);
        DE_CHECK_LAUNCH(ctx);
        k_bucket_bitsums<<<dim3(w0 + w1 + 1, nsets_total), DE_RC_THREADS, 0, st>>>(D0, D1, w0, w1, S);
        DE_CHECK_LAUNCH(ctx);
        k_bucket_bits_final<<<nsets_total, 64, 0, st>>>(S, w0, w1, set_out);
        DE_CHECK_LAUNCH(ctx);
        timing_end(ctx, tl);
    } else if (c >= 11) {
        // two-digit reduction: row / column plain sums by segmented additions, then the short weighted sums
        const unsigned int cm1 = c - 1, w0 = (cm1 + 1) / 2, w1 = cm1 - w0;
        const unsigned long long V0 = 1ull << w0, V1 = 1ull << w1;
        const unsigned int nd0 = (w0 + 4) / 5, nd1 = (w1 + 4) / 5;
        XYZZ* bufA = red2;
        XYZZ* bufB = red2 + (size_t)nsets_total * NB / 8 * 2;
        XYZZ* D0 = bufB + (size_t)nsets_total * NB / 64 * 2 + 64;
        XYZZ* D1 = D0 + nsets_total * V0;
        XYZZ* ds0 = D1 + nsets_total * V1;
        XYZZ* ds1 = ds0 + (size_t)nsets_total * nd0 * 32;
        auto segsum = [&](const XYZZ* in, XYZZ* out, unsigned long long n_out, unsigned int seg) -> int {
            k_xyzz_segsum<<<(unsigned int)((n_out + 127) / 128), 128, 0, st>>>(in, out, n_out, seg);
            DE_CHECK_LAUNCH(ctx);
            return DE_OK;
        };
        // reduce `len` contiguous terms per sum down to 1, 8 (then whatever is left) at a time
        auto reduce_rows = [&](const XYZZ* in, unsigned long long n_sums, unsigned long long len, XYZZ* final_out, XYZZ* t0, XYZZ* t1) -> int {
            const XYZZ* cur = in;
            XYZZ* tmp[2] = {t0, t1};
            int flip = 0;
            while (len > 1) {
                unsigned int seg = len >= 8 ? 8 : (unsigned int)len;
                unsigned long long nlen = len / seg;
                XYZZ* dst = nlen == 1 ? final_out : tmp[flip];
                DE_TRY(segsum(cur, dst, n_sums * nlen, seg));
                cur = dst;
                flip ^= 1;
                len = nlen;
            }
            return DE_OK;
        };
        TimedLaunch tl = timing_begin(ctx, "k_msm_digit_sums", (double)n * count);
        // D1[u] = sum over the V0 contiguous buckets of row u
        DE_TRY(reduce_rows(buckets, nsets_total * V1, V0, D1, bufA, bufB));
        // D0[v]: first level strided over the rows (8 at a time), then contiguous
        const unsigned int seg = V1 >= 8 ? 8 : (unsigned int)V1;
        const unsigned int Q = (unsigned int)(V1 / seg);
        const unsigned long long n_out = nsets_total * V0 * Q;
        XYZZ* first = Q == 1 ? D0 : bufA;
        k_xyzz_colsum<<<(unsigned int)((n_out + 127) / 128), 128, 0, st>>>(buckets, first, NB, w0, Q, seg, n_out);
        DE_CHECK_LAUNCH(ctx);
        if (Q > 1) DE_TRY(reduce_rows(bufA, nsets_total * V0, Q, D0, bufB, bufA + n_out));
        timing_end(ctx, tl);
        k_msm_digit_sums<<<dim3(nd0 * 32, nsets_total), 128, 0, st>>>(D0, (unsigned int)V0, w0, ds0);
        DE_CHECK_LAUNCH(ctx);
        k_msm_digit_sums<<<dim3(nd1 * 32, nsets_total), 128, 0, st>>>(D1, (unsigned int)V1, w1, ds1);
        DE_CHECK_LAUNCH(ctx);
        k_msm_digit_final2<<<nsets_total, 128, 0, st>>>(ds0, nd0, ds1, nd1, w0, set_out);
        DE_CHECK_LAUNCH(ctx);
    } else {
        DE_TIMED(ctx, "k_msm_digit_sums", (double)n * count,
                 (k_msm_digit_sums<<<dim3(ndigits * 32, nsets_total), 128, 0, st>>>(buckets, NB, c - 1, dsums)));
        DE_CHECK_LAUNCH(ctx);
        k_msm_digit_final<<<nsets_total, 128, 0, st>>>(dsums, ndigits, set_out);
        DE_CHECK_LAUNCH(ctx);
    }
    k_msm_combine<<<(unsigned int)count, 32, 0, st>>>(set_out, nsets, c, d_out);
    DE_CHECK_LAUNCH(ctx);
    return DE_OK;
}

}  // namespace de
