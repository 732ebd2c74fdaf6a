#!/bin/bash
# Reproduces the committed profiling artefacts on a B200 box (run under gpurun from the repo root):
#   gpurun --timeout 1200 -- 'bash profiles/run_profile.sh r01'                                  # headline (quadruped fp64)
#   gpurun --timeout 1200 -- 'bash profiles/run_profile.sh r01_quadrotor --model quadrotor --horizon 30 --batch 4096 --dtype f32'
# 1. bench line (not under a profiler)            -> gpurun_out/bench_<tag>.json
# 2. ncu launch list of the same command          -> gpurun_out/launches_<tag>.csv
# 3. one `--set full` capture of the sweep kernel -> gpurun_out/sweep_<tag>.ncu-rep
# then, here:  python profiles/summarize.py <tag>   (writes the tracked summaries under profiles/)
TAG=${1:-r01}
shift
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 10 "$@" > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 3000 gpurun_out/bench_${TAG}.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sweep|structured|team" -s 4 -c 2 -f -o gpurun_out/sweep_${TAG} \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -5
