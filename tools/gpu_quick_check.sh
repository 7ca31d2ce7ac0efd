#!/bin/bash
# One-GPU quick check of a build: the GPU parity suite, smoke(), and the default bench without its CPU-baseline leg.
TAG=${1:-quick}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; tail -2 $O/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 500 python bench.py --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; python - <<P
import json
d = json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "ntt_gb_s", "gpu_launches")}, d["e2e"]["value"], d["latency"]["create_proof_s"], d["clocks"])
P
tail -2 $O/${TAG}_bench.err
