#!/bin/bash
# ncu evidence for the consumers of the records (QP solve, line search, SQP loop) — run under gpurun from the repo root:
#   gpurun --timeout 1200 -- 'bash profiles/run_sqp_profile.sh r01'
# then, here:  python profiles/summarize_sqp.py r01
TAG=${1:-r01}
mkdir -p gpurun_out
for m in quadruped quadrotor rc_car; do python profiles/sqp_timing.py $m; done > gpurun_out/sqp_timing_${TAG}.log 2>&1
cat gpurun_out/sqp_timing_${TAG}.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/sqp_launches_${TAG}.csv \
    python profiles/sqp_timing.py quadruped > gpurun_out/ncu_sqp_launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"qp_twisted|qp_schur" -s 2 -c 1 -f -o gpurun_out/sqp_${TAG} \
    python profiles/sqp_timing.py quadruped > gpurun_out/ncu_sqp_full_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"line_search" -s 2 -c 1 -f -o gpurun_out/ls_${TAG} \
    python profiles/sqp_timing.py quadruped > gpurun_out/ncu_ls_full_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"qp_riccati" -s 1 -c 1 -f -o gpurun_out/riccati_${TAG} \
    python profiles/sqp_timing.py quadrotor > gpurun_out/ncu_riccati_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
