for cfg in "quadrotor 30 4096 f32" "rc_car 60 8192 f32" "quadrotor 30 4096 f64" "rc_car 60 8192 f64"; do set -- $cfg
 for env in "UNGAR_B200_NOOP=1" "UNGAR_B200_FORCE_GENERIC=1" "UNGAR_B200_FORCE_TPN=1 UNGAR_B200_TPN_STRUCT=0"; do
  env $env python bench.py --model $1 --horizon $2 --batch $3 --dtype $4 --steps 200 --warmup 10 --cpu-seconds 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d[\"roofline\"]; print(\"$cfg [$env]\", \"kernel_ms %.4f\"%r[\"kernel_ms\"], \"frac %.3f\"%r[\"frac\"], \"ms/step %.4f\"%d[\"ms_per_step\"], \"parity %.1e\"%d[\"parity\"][\"max_rel_err\"])"
 done
done
