# N=8 and N=4 weak-scaling bench lines (no extras) in one 8-GPU call
for n in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 --no-extras 2>&1 | grep '^{' | tail -1 > gpurun_out/r02_bench_n$n.json
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n$n.json').read())
    print($n, d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['timed_regions_ms'])
except Exception as e:
    print($n, 'failed', e)
PY
done
