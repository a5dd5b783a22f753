# usage: bash tools/knob_sweep_c0ip.sh -- q-block variants on the C0IP problems at 4 Mi cells (BASELINE config 3)
for V in c0ip_paper c0ip_minus; do
for cfg in "CM_Q_SWEEPS_WIDE=3" "CM_Q_SWEEPS_WIDE=6" "CM_Q_MG=1"; do
  echo "== $V $cfg"
  env $cfg timeout 300 python bench.py --nx 2048 --variant $V --steps 2 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
t=sys.stdin.read().strip()
try:
    d=json.loads(t)
    print('ms', round(d['ms_per_step'],2), 'its', d['linear_iterations_per_step'], {k:round(v,2) for k,v in d['phases_ms'].items()}, 'res', d['residual_norm_after_step'])
except Exception:
    print('FAILED:', t[-200:])"
done
done
