#!/bin/bash
# bench parity block for a few synthetic head statistics + the fusion pipeline test and bench
python -m pytest tests/test_gpu_models.py -q -k "pipeline" 2>&1 | tail -5
for w in "-1.75 0.5" "-2.0 1.0" "-2.3 1.5" "-2.1 1.5"; do
  set -- $w
  python bench.py --steps 5 --warmup 3 --no-sustained --no-cpu-baseline --obj-bias $1 --head-gain $2 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', d['value'], json.dumps(d['parity']))"
done
python bench.py --config fusion --steps 30 --warmup 3 > gpurun_out/bench_fusion_r2y.json 2> gpurun_out/bench_fusion_r2y.err
cut -c1-1800 gpurun_out/bench_fusion_r2y.json; tail -3 gpurun_out/bench_fusion_r2y.err
