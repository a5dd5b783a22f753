# usage: bash tools/dist_knobs.sh N  -- strip-partitioned Newton step on N GPUs for a few multigrid variants
N=${1:-8}
for cfg in "CM_DIST_LD=3 CM_MG_POST=2" "CM_DIST_LD=3 CM_MG_POST=1" "CM_DIST_LD=2 CM_MG_POST=2" "CM_DIST_LD=2 CM_MG_POST=1" "CM_DIST_LD=4 CM_MG_POST=1"; do
  echo "== N=$N $cfg"
  env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
lines=[l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')]
d=json.loads(lines[-1])
print('ms', round(d['ms_per_step'],2), 'its', d.get('linear_iterations_per_step'), {k:round(v,3) for k,v in d.get('phases_ms',{}).items()})"
done
