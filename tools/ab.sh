#!/bin/bash
# A/B of environment switches on one GPU: every argument is one "NAME=VALUE[,NAME=VALUE...]" setting (or "base"); each runs the
# short default bench (no CPU leg, no MSM / other-config blocks) and prints proofs/s, single-proof latency and the proof digest.
O=gpurun_out
mkdir -p $O
for s in "$@"; do
  tag=$(echo "$s" | tr ',=/' '___')
  envs=""
  if [ "$s" != "base" ]; then envs=$(echo "$s" | tr ',' ' '); fi
  env $envs timeout 400 python bench.py --no-cpu-baseline --no-msm --no-other-configs --steps ${AB_STEPS:-12} ${AB_ARGS} > $O/ab_${tag}.json 2> $O/ab_${tag}.err
  python - "$s" $O/ab_${tag}.json <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.2f e2e %.2f latency_ms %.3f e2e_latency_ms %.3f sha %s acc_ms %.4f" % (d["value"], d["e2e"]["value"], d["latency"]["create_proof_s"] * 1e3,
          d["latency"]["e2e_create_proof_s"] * 1e3, d["config"]["proof_sha256"][:12], d["roofline"]["avg_launch_ms"]), d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
done
