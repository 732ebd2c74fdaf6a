#!/bin/bash
# 8-GPU box: the H2D probe, then the bench at N = 8, 4, 2 (weak scaling) — gpurun --gpus 8 --timeout 900 -- 'bash profiles/run_scaling.sh r02b'
TAG=${1:-r02b}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29601 profiles/h2d_probe.py > gpurun_out/h2d_probe_${TAG}.txt 2>&1
$TR --nproc-per-node 4 --master-port 29602 profiles/h2d_probe.py > gpurun_out/h2d_probe4_${TAG}.txt 2>&1
for n in 8 4 2; do
  $TR --nproc-per-node $n --master-port 2961$n bench.py --gpus $n --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_${n}gpu.json 2> gpurun_out/${TAG}_bench_${n}gpu.err
done
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-other-configs > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
cat gpurun_out/h2d_probe_${TAG}.txt | grep -v "^\s*$" | head -30; head -4 gpurun_out/h2d_probe4_${TAG}.txt
for n in 1 2 4 8; do python - <<EOF
import json
d=json.loads([l for l in open("gpurun_out/${TAG}_bench_${n}gpu.json") if l.startswith("{")][-1])
print($n, "value %.0fM e2e %.0fM compact %.0fM sqp %.2f ms sqp_e2e %.1f ms parity %s"%(d["value"]/1e6, d["e2e"]["value"]/1e6, d["compact"]["value"]/1e6, d["sqp_loop"]["ms_per_solve"], d["sqp_loop"]["e2e"]["ms_per_solve"], (d.get("parity") or {}).get("ok")))
EOF
done
