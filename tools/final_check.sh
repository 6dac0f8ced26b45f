# End-of-round check on N GPUs (default 8): smoke, N-GPU and N/2-GPU weak-scaling lines (no extras)
n=${1:-8}
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for k in $n $((n/2)); do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 2952$k bench.py --gpus $k --steps 5 --warmup 3 --no-extras 2>&1 | grep '^{' | tail -1 > gpurun_out/r02_bench_n$k.json
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n$k.json').read())
    print($k, '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], [round(x,1) for x in d['config']['timed_regions_ms']], d['config']['grid_overflow'])
except Exception as e:
    print($k, 'failed', e)
PY
done
