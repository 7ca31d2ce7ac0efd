"""ctypes binding of libde_b200.so (include/de_b200.h).  Fails loudly when the CUDA library is missing: there is no
CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# DE_B200_LIB: another build of the same library (A/B measurements of compile-time variants, tools/ab.sh)
LIB_PATH = os.environ.get("DE_B200_LIB") or os.path.join(_PKG, "libde_b200.so")

DE_OK, DE_ERR_ARG, DE_ERR_CUDA, DE_ERR_OOM, DE_ERR_UNSUPPORTED = 0, -1, -2, -3, -4
OP_MUL, OP_ADD, OP_SUB, OP_FROM_MONT, OP_TO_MONT, OP_INV, OP_SQR = 0, 1, 2, 3, 4, 5, 6

# every symbol include/de_b200.h declares (tests/test_abi.py checks the header against this list and the .so)
SYMBOLS = [
    "de_ctx_create", "de_ctx_destroy", "de_ctx_set_stream", "de_ctx_sync", "de_last_error", "de_version", "de_launch_count",
    "de_timing_enable", "de_timing_reset", "de_timing_get",
    "de_fr_vec_op", "de_fq_vec_op", "de_msm", "de_msm_dev", "de_params_upload", "de_params_free", "de_commit",
    "de_commit_batch", "de_commit_batch_dev", "de_ntt", "de_ntt_dev", "de_domain_create", "de_domain_free", "de_domain_info",
    "de_coeff_to_extended", "de_extended_to_coeff", "de_lagrange_to_coeff", "de_coeff_to_lagrange", "de_divide_by_vanishing",
    "de_coeff_to_extended_dev", "de_extended_to_coeff_dev", "de_lagrange_to_coeff_dev", "de_coeff_to_lagrange_dev",
    "de_divide_by_vanishing_dev", "de_pk_upload", "de_pk_free", "de_evaluate_h", "de_evaluate_h_dev", "de_pk_extend_dev", "de_evaluate_h_rows_dev", "de_commit_range", "de_g1_sum", "de_g1_batch_normalize",
    "de_commit_batch_canonical_dev", "de_eval_polynomial", "de_kate_division", "de_prover_create", "de_prover_free",
    "de_prover_random_count", "de_prover_proof_size", "de_create_proof", "de_create_proof_dev", "de_g1_mul_base_dev", "de_ctx_set_mode", "de_commit_sharded",
    "de_ntt_dist_stage1", "de_ntt_dist_stage2", "de_ntt_sharded_dev", "de_ntt_sharded", "de_dev_alloc", "de_dev_free", "de_dev_copy", "de_ipc_export", "de_ipc_import",
    "de_ipc_release", "de_int_peak", "de_int_peak_sqr", "de_ntt_dist_run", "de_ntt_dist_error", "de_ntt_dist_prepare",
    "de_circuit_synthesize", "de_circuit_witness", "de_assignment_free", "de_assignment_info", "de_assignment_fixed", "de_assignment_advice",
    "de_assignment_copies", "de_assignment_outputs", "de_assignment_sigma", "de_frontend_last_error", "de_poseidon_permute",
    "de_poseidon_cipher",
]


class DeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"de_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the CUDA extension is required; there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    P, SZ, U32, I = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int
    L.de_ctx_create.argtypes = [I, C.POINTER(P)]
    L.de_ctx_destroy.argtypes = [P]
    L.de_ctx_set_stream.argtypes = [P, P]
    L.de_ctx_sync.argtypes = [P]
    L.de_ctx_set_mode.argtypes = [P, I]
    L.de_last_error.argtypes = [P]
    L.de_last_error.restype = C.c_char_p
    L.de_version.restype = C.c_char_p
    L.de_launch_count.argtypes = [P]
    L.de_launch_count.restype = C.c_uint64
    L.de_timing_enable.argtypes = [P, I]
    L.de_timing_reset.argtypes = [P]
    L.de_timing_get.argtypes = [P, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.de_fr_vec_op.argtypes = [P, I, P, P, P, SZ]
    L.de_fq_vec_op.argtypes = [P, I, P, P, P, SZ]
    L.de_msm.argtypes = [P, P, P, SZ, P]
    L.de_msm_dev.argtypes = [P, P, P, SZ, P]
    L.de_params_upload.argtypes = [P, U32, P, P, C.POINTER(P)]
    L.de_params_free.argtypes = [P]
    L.de_commit.argtypes = [P, I, P, SZ, P]
    L.de_commit_batch.argtypes = [P, I, C.POINTER(P), SZ, SZ, P]
    L.de_commit_batch_dev.argtypes = [P, I, P, SZ, SZ, SZ, P]
    L.de_ntt.argtypes = [P, P, P, U32]
    L.de_ntt_dev.argtypes = [P, P, P, U32, SZ, SZ]
    L.de_domain_create.argtypes = [P, U32, U32, C.POINTER(P)]
    L.de_domain_free.argtypes = [P]
    L.de_domain_info.argtypes = [P, C.POINTER(U32), P]
    L.de_coeff_to_extended.argtypes = [P, P, P]
    L.de_extended_to_coeff.argtypes = [P, P, C.POINTER(SZ)]
    L.de_lagrange_to_coeff.argtypes = [P, P]
    L.de_coeff_to_lagrange.argtypes = [P, P]
    L.de_divide_by_vanishing.argtypes = [P, P]
    L.de_coeff_to_extended_dev.argtypes = [P, P, SZ, P, SZ, SZ]
    L.de_extended_to_coeff_dev.argtypes = [P, P, SZ, SZ, C.POINTER(SZ)]
    L.de_lagrange_to_coeff_dev.argtypes = [P, P, SZ, SZ]
    L.de_coeff_to_lagrange_dev.argtypes = [P, P, SZ, SZ]
    L.de_divide_by_vanishing_dev.argtypes = [P, P, SZ, SZ]
    L.de_pk_upload.argtypes = [P, P, C.POINTER(P)]
    L.de_pk_free.argtypes = [P]
    L.de_evaluate_h.argtypes = [P, P, P, P, P, P, P]
    L.de_evaluate_h_dev.argtypes = [P, P, P, P, P, P, SZ, P]
    L.de_pk_extend_dev.argtypes = [P, P, P, P, P, SZ]
    L.de_evaluate_h_rows_dev.argtypes = [P, P, P]
    L.de_commit_range.argtypes = [P, I, P, SZ, SZ, P]
    L.de_g1_sum.argtypes = [P, P, SZ, P]
    L.de_g1_batch_normalize.argtypes = [P, P, SZ, P]
    L.de_commit_batch_canonical_dev.argtypes = [P, I, P, SZ, SZ, SZ, P]
    L.de_eval_polynomial.argtypes = [P, P, SZ, P, P]
    L.de_kate_division.argtypes = [P, P, SZ, P, P]
    L.de_prover_create.argtypes = [P, P, P, C.POINTER(P)]
    L.de_prover_free.argtypes = [P]
    L.de_prover_random_count.argtypes = [P]
    L.de_prover_random_count.restype = SZ
    L.de_prover_proof_size.argtypes = [P]
    L.de_prover_proof_size.restype = SZ
    L.de_create_proof.argtypes = [P, C.POINTER(P), C.POINTER(P), C.POINTER(SZ), P, SZ, P, SZ, C.POINTER(SZ)]
    L.de_commit_sharded.argtypes = [C.POINTER(P), C.POINTER(SZ), C.POINTER(SZ), I, I, P, P]
    L.de_g1_mul_base_dev.argtypes = [P, P, P, SZ, P]
    L.de_create_proof_dev.argtypes = [P, P, SZ, C.POINTER(P), C.POINTER(SZ), P, SZ, P, SZ, C.POINTER(SZ)]
    L.de_ntt_dist_stage1.argtypes = [P, P, P, U32, U32, U32, C.POINTER(P)]
    L.de_ntt_dist_stage2.argtypes = [P, P, P, U32, U32, U32, C.POINTER(P)]
    L.de_ntt_sharded_dev.argtypes = [C.POINTER(P), I, C.POINTER(P), C.POINTER(P), P, U32]
    L.de_ntt_sharded.argtypes = [C.POINTER(P), I, P, P, U32]
    L.de_dev_alloc.argtypes = [P, SZ, C.POINTER(P)]
    L.de_dev_free.argtypes = [P, P]
    L.de_dev_copy.argtypes = [P, P, P, SZ]
    L.de_ipc_export.argtypes = [P, P, P]
    L.de_ipc_import.argtypes = [P, P, C.POINTER(P)]
    L.de_ipc_release.argtypes = [P, P]
    L.de_int_peak.argtypes = [P, C.POINTER(C.c_double)]
    L.de_int_peak_sqr.argtypes = [P, C.POINTER(C.c_double)]
    L.de_ntt_dist_run.argtypes = [P, P, P, U32, U32, U32, C.POINTER(P), C.POINTER(P), C.POINTER(P), U32, U32]
    L.de_ntt_dist_error.argtypes = [P, C.POINTER(I)]
    L.de_ntt_dist_prepare.argtypes = [P, P, U32, U32, U32]
    L.de_circuit_synthesize.argtypes = [P, C.POINTER(P)]
    L.de_circuit_witness.argtypes = [P, P, P]
    L.de_assignment_free.argtypes = [P]
    L.de_assignment_free.restype = None
    L.de_assignment_info.argtypes = [P, P]
    L.de_assignment_fixed.argtypes = [P, U32, P]
    L.de_assignment_advice.argtypes = [P, U32, P]
    L.de_assignment_copies.argtypes = [P, P]
    L.de_assignment_outputs.argtypes = [P, P]
    L.de_assignment_sigma.argtypes = [P, P, P, U32, P]
    L.de_frontend_last_error.restype = C.c_char_p
    L.de_poseidon_permute.argtypes = [U32, U32, U32, P]
    L.de_poseidon_cipher.argtypes = [I, P, P, U32, P]
    for s in SYMBOLS:
        fn = getattr(L, s)
        if s not in ("de_last_error", "de_version", "de_launch_count", "de_prover_random_count", "de_prover_proof_size",
                     "de_assignment_free", "de_frontend_last_error"):
            fn.restype = C.c_int
    _lib = L
    return L
