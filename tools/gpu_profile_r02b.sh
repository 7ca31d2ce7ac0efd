#!/bin/bash
# One-GPU profile pass of the final round-2 build: GPU parity suite, launch list of the latency path (one proof in flight), ncu
# --set full of the bucket fill / evaluator / NTT passes.  Outputs under gpurun_out/<tag>_*; run through gpurun.
TAG=${1:-r02b}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; tail -2 $O/${TAG}_pytest.log
CMD="python bench.py --steps 2 --warmup 3 --inflight 1 --mode latency --no-cpu-baseline --no-msm --no-other-configs"
timeout 200 $CMD > $O/${TAG}_plain.json 2> $O/${TAG}_plain.err && \
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${TAG}_launches_latency.csv $CMD > $O/${TAG}_ncu1.log 2>&1
python tools/launch_summary.py $O/${TAG}_launches_latency.csv > $O/${TAG}_launches_latency_summary.txt 2>&1; head -12 $O/${TAG}_launches_latency_summary.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_msm_accumulate|k_eval_h|k_ntt_pass|k_msm_digits|k_msm_scatter|k_msm_merge_large|k_bucket_rowcol" -s 250 -c 40 -f -o $O/${TAG}_full $CMD > $O/${TAG}_ncu2.log 2>&1
ncu -i $O/${TAG}_full.ncu-rep --page raw --csv > $O/${TAG}_full_raw.csv 2>/dev/null; wc -c $O/${TAG}_full_raw.csv; rm -f $O/${TAG}_full.ncu-rep
