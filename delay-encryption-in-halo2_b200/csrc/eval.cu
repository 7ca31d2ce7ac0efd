// eval.cu — quotient-polynomial evaluator (halo2_proofs::plonk::evaluation::Evaluator::evaluate_h), SURVEY.md row a9.
#include "common.cuh"

using namespace de;

extern "C" {
int de_pk_upload(de_domain* d, const de_pk_desc* desc, de_pk** out) {
    (void)d; (void)desc;
    if (out) *out = nullptr;
    return DE_ERR_UNSUPPORTED;
}
int de_pk_free(de_pk* pk) { (void)pk; return DE_OK; }
int de_evaluate_h(de_pk* pk, const de_fr* const* advice_coeff, const de_fr* const* instance_coeff, const de_challenges* ch,
                  const de_fr* const* perm_z_coeff, const de_fr* const* lookup_coeff, de_fr* h_ext_out) {
    (void)pk; (void)advice_coeff; (void)instance_coeff; (void)ch; (void)perm_z_coeff; (void)lookup_coeff; (void)h_ext_out;
    return DE_ERR_UNSUPPORTED;
}
}
