#!/bin/bash
# multi-GPU bench lines (torchrun, one rank per GPU): $1 = N, $2 = tag
N=$1; tag=${2:-r2}
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err
run --config fusion --steps 30 --warmup 3 > gpurun_out/bench_fusion_${tag}_n$N.json 2> gpurun_out/bench_fusion_${tag}_n$N.err
run --config train3 --steps 20 --warmup 3 > gpurun_out/bench_train3_${tag}_n$N.json 2> gpurun_out/bench_train3_${tag}_n$N.err
for f in bench_${tag}_n$N bench_fusion_${tag}_n$N bench_train3_${tag}_n$N; do cut -c1-260 gpurun_out/$f.json; tail -2 gpurun_out/$f.err; done
