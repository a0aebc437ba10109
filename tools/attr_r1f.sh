for dbg in 19; do echo "DBG $dbg"; for c in e53_208_3x3_32_64_res d53_104_3x3_64_128_res d53_26_1x1_512_256 d53_13_1x1_1024_512 d53_52_1x1_256_128; do ME_CONV_MODE=single ME_CONV_DBG=$dbg python tools/gpu_probe_conv.py $c 2>&1 | grep PROBE | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l.split(' ',1)[1]); print('  %-28s ms=%.4f TF=%.0f' % (r['case'], r.get('ms',0), r.get('tflops',0)))
"; done; done
