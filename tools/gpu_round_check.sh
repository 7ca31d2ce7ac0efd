#!/bin/bash
# One-GPU end-of-session check: parity suite, smoke, default bench + reference arm, launch list and ncu --set full of the multi-GPU
# NTT kernels (two ranks on one device), single-GPU MSM at 2^26.  Outputs under gpurun_out/<tag>_*; run through gpurun.
TAG=${1:-r01s3}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ntt_dist.py -x -q > $O/${TAG}_ntt_dist_pytest.log 2>&1; tail -2 $O/${TAG}_ntt_dist_pytest.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; tail -2 $O/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 700 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 300 $O/${TAG}_bench.json; tail -2 $O/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_reference_arm.err; cut -c1-200 $O/${TAG}_bench_reference_arm.json
CMD="python tools/ntt_dist_sweep.py --gpus 2 --log-n 24 --reps 1"
timeout 120 $CMD > $O/${TAG}_ntt_virtual_plain.log 2>&1 && \
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_ntt_dist_launches.csv $CMD > $O/${TAG}_ncu1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_ntt_cross|k_ntt_pass" -s 16 -c 8 -f -o $O/${TAG}_ntt_dist_full $CMD > $O/${TAG}_ncu2.log 2>&1
ncu -i $O/${TAG}_ntt_dist_full.ncu-rep --page raw --csv > $O/${TAG}_ntt_dist_full_raw.csv 2>/dev/null; wc -c $O/${TAG}_ntt_dist_full_raw.csv
DE_MSM_TABLE_MB=80000 timeout 500 python tools/sweep.py --msm-from 26 --msm-to 26 --ntt-to 15 > $O/${TAG}_msm_2p26.jsonl 2> $O/${TAG}_msm_2p26.err; cat $O/${TAG}_msm_2p26.jsonl; tail -2 $O/${TAG}_msm_2p26.err
timeout 200 python bench.py --config pose_enc --inflight 32 --no-cpu-baseline > $O/${TAG}_pose_enc_inflight32.json 2> $O/${TAG}_pose_enc_inflight32.err; cut -c1-160 $O/${TAG}_pose_enc_inflight32.json
