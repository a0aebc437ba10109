#!/bin/bash
# Round-2 check of the conv chain kernel on one B200 (under gpurun): chain-vs-per-layer tests, the Darknet model tests,
# then the bench line with chains off and on.  $1 = tag.
tag=${1:-r2a}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q > gpurun_out/chain_tests_$tag.log 2>&1
echo "chain tests rc=$?" | tee -a gpurun_out/chain_tests_$tag.log
tail -15 gpurun_out/chain_tests_$tag.log
timeout 900 python -m pytest tests/test_gpu_models.py -x -q -k "darknet" > gpurun_out/model_tests_$tag.log 2>&1
echo "model tests rc=$?" | tee -a gpurun_out/model_tests_$tag.log
tail -15 gpurun_out/model_tests_$tag.log
ME_CONV_CHAIN=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_nochain.json 2> gpurun_out/bench_${tag}_nochain.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_chain.json 2> gpurun_out/bench_${tag}_chain.err
for f in nochain chain; do python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/bench_${tag}_$f.json").read().strip().splitlines()[-1])
    print("$f", "fps", round(l["value"]), "e2e", round(l["e2e"]["value"]), "conv_ms", round(l["roofline"]["conv_ms_per_step"], 3), "TF", round(l["roofline"]["achieved"], 1), "launches", l["roofline"]["launches"])
except Exception as e:
    print("$f failed:", e)
    print(open("gpurun_out/bench_${tag}_$f.err").read()[-2000:])
PY
done
