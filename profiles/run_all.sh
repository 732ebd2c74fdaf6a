#!/bin/bash
# One gpurun call that refreshes every tracked artefact of a round (the raw .ncu-rep files together exceed what gpurun copies back, so
# the summaries are made on the box and only they — plus the QP report — return):
#   gpurun --timeout 1800 -- 'bash profiles/run_all.sh r02'    then here:   cp gpurun_out/summaries_r02/* profiles/
TAG=${1:-r02}
bash profiles/run_profile.sh ${TAG} > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"quadruped_compact" -s 2 -c 1 -f -o gpurun_out/compact_${TAG} \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_compact_${TAG}.log 2>&1
bash profiles/run_sqp_profile.sh ${TAG} > /dev/null 2>&1
python profiles/summarize.py ${TAG}
python profiles/summarize_sqp.py ${TAG}
python profiles/ncu_summary.py gpurun_out/compact_${TAG}.ncu-rep > profiles/${TAG}_ncu_compact_sweep.json
python profiles/ncu_functions.py gpurun_out/sqp_${TAG}.ncu-rep > profiles/${TAG}_ncu_qp_functions.txt
python profiles/rbd_timing.py > profiles/${TAG}_rbd_timing.txt 2>&1
python profiles/tape_timing.py > profiles/${TAG}_tape_timing.txt 2>&1
mkdir -p gpurun_out/summaries_${TAG}
cp profiles/${TAG}_* profiles/traffic.json gpurun_out/summaries_${TAG}/
rm -f gpurun_out/sweep_${TAG}.ncu-rep gpurun_out/ls_${TAG}.ncu-rep gpurun_out/riccati_${TAG}.ncu-rep gpurun_out/compact_${TAG}.ncu-rep
cat profiles/${TAG}_sqp_timing.txt profiles/${TAG}_rbd_timing.txt
