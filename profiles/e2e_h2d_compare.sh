for cfg in "--model quadruped" "--model rc_car --horizon 60 --batch 8192 --dtype f32" "--model quadrotor --horizon 30 --batch 4096 --dtype f32"; do
 for env in "UNGAR_B200_H2D_PITCHED=1" "UNGAR_B200_H2D_PITCHED=0"; do
  env $env python bench.py $cfg --steps 100 --warmup 5 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$cfg [$env] value %.4g e2e %.4g e2e_full_xp %.4g'%(d['value'], d['e2e']['value'], (d.get('e2e_full_xp') or {}).get('value',0)))"
 done
done
