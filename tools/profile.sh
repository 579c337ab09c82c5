#!/bin/bash
# One ncu --set full capture of the evaluation kernel + the launch list of a short bench run (run under gpurun, ONE GPU).
#   tools/profile.sh <tag> [objects]
# Outputs (gpurun_out/): prof_<tag>.ncu-rep, launches_<tag>.csv, ncu_run_<tag>.log
set -u
TAG=${1:-dev}
N=${2:-500000}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name gk_eval_kernel --launch-skip 3 --launch-count 1 \
    -o gpurun_out/prof_${TAG} -f python tools/kernel_probe.py ${N} 4 > gpurun_out/ncu_run_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --objects ${N} --e2e-steps 1 --cpu-sample 20 > gpurun_out/launches_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_run_${TAG}.log
