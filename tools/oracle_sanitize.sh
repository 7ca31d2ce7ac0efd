#!/bin/bash
# The CPU oracle (oracle/oracle.c, test infrastructure) under AddressSanitizer + UBSan and under ThreadSanitizer: a copy of oracle/
# is built with the sanitizer in a scratch directory (the repository's liboracle.so is not touched) and driven from Python with the
# runtime preloaded: MSM (1 and 4 threads), FFT (1 and 3 threads), the domain operations, eval_polynomial / kate_division, batch
# inversion, then three whole restated proofs (k = 5, 6, 9; with and without lookups, public inputs) checked by the restated verifier.
# Usage: tools/oracle_sanitize.sh      (exit 0 = no report)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d /tmp/orc_san.XXXXXX)
cp "$ROOT"/oracle/*.py "$ROOT"/oracle/oracle.c "$ROOT"/oracle/Makefile "$W"/
cat > "$W/run.py" <<P
import sys
sys.path.insert(0, "$ROOT/delay-encryption-in-halo2_b200")
sys.path.insert(0, "$W")
import orc, pyoracle as po, pyprover as pp
assert orc.__file__.startswith("$W")
from de_b200 import circuits, plonk
n = 1 << 11
s = orc.uniform_fr(1, n); b = orc.gen_bases(n)
assert (orc.g1_to_affine(orc.best_multiexp(s, b, threads=4)) == orc.g1_to_affine(orc.best_multiexp(s, b, threads=1))).all()
assert (orc.g1_to_affine(orc.msm_naive(s[:64], b[:64])) == orc.g1_to_affine(orc.best_multiexp(s[:64], b[:64]))).all()
d = orc.Domain(5, 11)
assert (d.extended_to_coeff(d.coeff_to_extended(s).copy())[:n] == s).all()
assert (orc.best_fft(s, d.omega, 11, threads=3) == orc.best_fft(s, d.omega, 11, threads=1)).all()
x = orc.uniform_fr(2, 1)
orc.eval_polynomial(s, x[0]); orc.kate_division(s, x[0]); orc.fr_batch_invert(s, threads=3)
for with_lookups, k, used, n_public in [(False, 5, 20, 0), (True, 6, 40, 2), (True, 9, 300, 3)]:
    asg = circuits.satisfied_assignment(with_lookups, k, 0xFA57 + k, used, n_public=n_public)
    params = pp.setup(k, 0x1234567)
    pk = pp.keygen(params, asg.shape, pp.Queries(*plonk.collect_queries(asg.shape)), asg.fixed, asg.copies, 0xABCDEF)
    rng = po.Xoshiro(99)
    draws = [rng.uniform_fr() for _ in range(pp.random_count(asg.shape, 1 << k))]
    proof = pp.create_proof_fast(params, pk, asg.advice, asg.instances, pp.to_mont(draws))
    assert pp.verify_proof(params, pk.vk, asg.instances, proof)
print("ok")
P
for mode in "address,undefined -fno-sanitize-recover=undefined" "thread"; do
  rm -f "$W/liboracle.so"
  make -s -C "$W" CC=gcc CFLAGS="-O1 -g -march=x86-64-v3 -fPIC -std=gnu11 -fsanitize=$mode" liboracle.so
  touch "$W/liboracle.so"
  if [ "$mode" = thread ]; then pre=$(gcc -print-file-name=libtsan.so); else pre=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so); fi
  echo "== -fsanitize=$mode"
  LD_PRELOAD=$pre ASAN_OPTIONS=detect_leaks=0 TSAN_OPTIONS=exitcode=66 python "$W/run.py"
done
rm -rf "$W"
