# Which operand stream starves the pair kernel?  ME_CONV_DBG 32 = no weight loads, 64 = no activation loads (results are
# garbage, timing only).  noref-style timing through the probe's "spot" cases.
for dbg in 0 32 64; do echo "DBG $dbg"; for c in d53_52_3x3_128_256_res d53_26_3x3_256_512_res; do ME_PAIR_SPLIT=0 ME_CONV_DBG=$dbg python tools/gpu_probe_conv.py $c 2>&1 | grep PROBE | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l.split(' ',1)[1]); print('  %-28s ms=%.4f TF=%.0f' % (r['case'], r.get('ms',0), r.get('tflops',0)))
"; done; done
