#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_train3.csv \
    python bench.py --config train3 --steps 2 --warmup 1 > gpurun_out/b_ncu4.log 2>&1
tail -2 gpurun_out/b_ncu4.log | cut -c1-200
