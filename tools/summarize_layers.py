"""Groups tools/layer_profile.py output by layer shape: python tools/summarize_layers.py file.jsonl"""
import json
import sys
from collections import OrderedDict

rows = [json.loads(l.split(' ', 1)[1]) for l in open(sys.argv[1]) if l.startswith('LAYER')]
tot = sum(r['ms'] for r in rows)
print('total ms %.3f' % tot)
g = OrderedDict()
for r in rows:
    if r['type'] != 'convolutional':
        g[(r['type'], r['block'])] = [r['ms'], 1, 0, 0, 0]
        continue
    key = (r['hw'], r['k'], r['s'], r['cin'], r['cout'], r['res'])
    e = g.setdefault(key, [0, 0, 0, 0, 0])
    e[0] += r['ms']; e[1] += 1; e[2] = r['tflops']; e[3] = r['gbs']; e[4] += r['gflop']
for k, v in g.items():
    print(k, 'ms=%.3f n=%d tflops=%s gbs=%s share=%.1f%% gflop=%.0f' % (v[0], v[1], v[2], v[3], 100 * v[0] / tot, v[4]))
