#!/bin/bash
# Round-end evidence run on one B200 (under gpurun): bench line, ncu launch list (durations only), DRAM bytes of the conv
# launches (second pass), in-kernel phase trace, one `ncu --set full` capture of the main kernels.  $1 = tag (e.g. r1j).
tag=${1:-final}
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_${tag}_dram.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
python tools/conv_trace.py > gpurun_out/trace_$tag.jsonl 2>&1
ONLY_BLOCKS=3,7,14,39,64 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 74 -c 10 \
    -o gpurun_out/prof_$tag python tools/conv_trace.py > gpurun_out/prof_$tag.log 2>&1
python tools/bench_fusion.py 30 > gpurun_out/fusion_$tag.log 2>&1
cut -c1-400 gpurun_out/bench_$tag.json
tail -1 gpurun_out/fusion_$tag.log
