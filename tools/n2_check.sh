timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | grep '^{' | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | grep '^{' | tail -1
