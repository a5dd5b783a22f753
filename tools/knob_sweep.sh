# usage: bash tools/knob_sweep.sh  -- Newton-step time for multigrid / q-sweep variants (experiment tool)
for cfg in "CM_MG_PRE=1 CM_MG_POST=2 CM_Q_SWEEPS=1" "CM_MG_PRE=2 CM_MG_POST=2 CM_Q_SWEEPS=2" "CM_MG_PRE=2 CM_MG_POST=2 CM_Q_SWEEPS=1" "CM_MG_PRE=1 CM_MG_POST=3 CM_Q_SWEEPS=1" "CM_MG_PRE=2 CM_MG_POST=1 CM_Q_SWEEPS=1" "CM_MG_PRE=2 CM_MG_POST=3 CM_Q_SWEEPS=1" "CM_MG_PRE=3 CM_MG_POST=3 CM_Q_SWEEPS=2"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms', round(d['ms_per_step'],2), 'its', d['linear_iterations_per_step'], 'krylov', round(d['phases_ms']['krylov_solve'],2), 'res', d['residual_norm_after_step'])"
done
