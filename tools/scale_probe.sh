#!/bin/bash
# e2e at N=2: default, no host_all readback, no gather
N=2
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline --no-sustained --no-parity 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['e2e']['ms_per_step'])"; }
run default
ME_BENCH_HOSTALL=0 run no_host_all
ME_BENCH_GATHER=0 run no_gather
