# A/B of library variants on the N=2 weak-scaling bench line: tools/n2_ab.sh <variant>...  ("main" = in-tree)
for v in "$@"; do
  if [ "$v" = main ]; then lib=""; else lib="$PWD/wgsparkl_b200/_variants/lib_$v.so"; fi
  B200MPM_LIB=$lib timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-extras 2>&1 | grep '^{' | tail -1 > gpurun_out/n2_$v.json
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/n2_$v.json').read())
    print('$v', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], [round(x,1) for x in d['config']['timed_regions_ms']])
except Exception as e:
    print('$v', 'failed', e)
PY
done
