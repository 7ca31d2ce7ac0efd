#!/bin/bash
# compute-sanitizer over one small real proof and the small kernels' tests: memcheck (out-of-bounds / misaligned), racecheck
# (shared-memory hazards in the reduction trees, scans and NTT tiles), initcheck (reads of device memory nobody wrote),
# synccheck (barriers under divergence).  Run on a B200: tools/sanitize.sh [outdir]; one log per tool + a summary line each.
out=${1:-gpurun_out}
mkdir -p "$out"
target=(python -c "import __graft_entry__ as g; g.smoke()")
for tool in memcheck racecheck initcheck synccheck; do
    log="$out/r02_sanitizer_${tool}.log"
    timeout 1200 compute-sanitizer --tool "$tool" --print-limit 20 --error-exitcode 9 "${target[@]}" > "$log" 2>&1
    rc=$?
    echo "$tool rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$log" | tail -1)"
done
