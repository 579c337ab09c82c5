#!/bin/bash
# One ncu --set full capture of the evaluation kernel + the launch list of a short bench run (run under gpurun, ONE GPU).
#   tools/profile.sh <tag> [objects] [kernel]
# kernel: gk_spec_kernel (default: the kernel generated for the constraint set -- what a page of >= 8192 objects runs on) or
#         gk_eval_kernel (the netlist interpreter; the probe then runs with GK_SPEC=0)
# Outputs (gpurun_out/): prof_<tag>.ncu-rep, launches_<tag>.csv, ncu_run_<tag>.log; summarise with profiles/ncu_summary.py / ncu_lines.py.
set -u
TAG=${1:-dev}
N=${2:-1000000}
KERNEL=${3:-gk_spec_kernel}
mkdir -p gpurun_out
if [ "$KERNEL" = "gk_eval_kernel" ]; then export GK_SPEC=0; fi
ncu --set full --clock-control none --import-source on --kernel-name ${KERNEL} --launch-skip 3 --launch-count 1 \
    -o gpurun_out/prof_${TAG} -f python tools/kernel_probe.py ${N} 4 > gpurun_out/ncu_run_${TAG}.log 2>&1
unset GK_SPEC
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --objects ${N} --e2e-steps 1 --cpu-sample 20 > gpurun_out/launches_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_run_${TAG}.log
