#!/bin/bash
# ncu launch list (durations + DRAM bytes) of the fusion bench (BASELINE config 3)
tag=${1:-r2y}
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_fusion_$tag.csv python bench.py --config fusion --steps 2 --warmup 1 > gpurun_out/b_ncu3.log 2>&1
tail -3 gpurun_out/b_ncu3.log | cut -c1-300
