#!/bin/bash
# Round-2 evidence run on one B200 (under gpurun).  $1 = tag (e.g. r2z).
#  1 bench line (driver contract) + the reference arm            -> bench_$tag.json, bench_${tag}_reference.json
#  2 ncu launch list, durations only, then with DRAM bytes       -> launches_$tag.csv, launches_${tag}_dram.csv
#  3 timeout 400 ncu --set full of the chain kernel (first chain launch)     -> prof_chain_$tag.ncu-rep
#  4 in-kernel trace of the chains + per-launch event times      -> chain_trace_$tag.log
#  5 BASELINE configs 3 and 4                                     -> bench_fusion_$tag.json, bench_train3_$tag.json
tag=${1:-r2z}
mkdir -p gpurun_out
timeout 300 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sustained --no-parity > gpurun_out/b_ncu.log 2>&1
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_${tag}_dram.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sustained --no-parity > gpurun_out/b_ncu2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_chain -s 6 -c 3 -o gpurun_out/prof_chain_$tag \
    python tools/chain_smoke.py 32 416 > gpurun_out/prof_chain_$tag.log 2>&1
timeout 200 python tools/chain_trace.py > gpurun_out/chain_trace_$tag.log 2>&1
timeout 200 python bench.py --config fusion --steps 30 --warmup 3 > gpurun_out/bench_fusion_$tag.json 2> gpurun_out/bench_fusion_$tag.err
timeout 200 python bench.py --config train3 --steps 20 --warmup 3 > gpurun_out/bench_train3_$tag.json 2> gpurun_out/bench_train3_$tag.err
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_info_$tag.csv
cut -c1-300 gpurun_out/bench_$tag.json
cut -c1-300 gpurun_out/bench_${tag}_reference.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")" > gpurun_out/smoke_$tag.log 2>&1; tail -2 gpurun_out/smoke_$tag.log
