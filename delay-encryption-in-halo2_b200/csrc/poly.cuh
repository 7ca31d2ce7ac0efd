// poly.cuh — device-side entry points of poly.cu used by the prover (prover.cu).
#pragma once
#include "common.cuh"

namespace de {

// out[e] = polys[e](points[point_index[e]]) for e < count (point_index == nullptr: points[e]); n coefficients each.
// d_out (Montgomery) and d_out_canonical (Fr::to_repr limbs) may each be nullptr.
int eval_polynomials_dev(de_ctx* ctx, const Fr* const* d_polys, size_t n, const unsigned int* d_point_index, const Fr* d_points, size_t count,
                         Fr* d_out, Fr* d_out_canonical);
size_t kate_scratch_elems(size_t n, size_t count);
int kate_division_dev(de_ctx* ctx, const Fr* const* d_as, size_t n, const de_fr* host_bs, size_t count, Fr* d_q, size_t q_stride, Fr* d_scratch);

}  // namespace de
