"""Maps the conv-only graph replay at the end of an ncu launch list (bench.py) onto Darknet-53 layers.
usage: python tools/ncu_layers.py launches.csv [cfg]"""
import csv, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from millieye_b200 import configs
from millieye_b200.engine import describe_blocks
from millieye_b200.parse_config import parse_model_config
from collections import OrderedDict
cfg = sys.argv[2] if len(sys.argv) > 2 else 'yolov3'
N, S = 32, 416
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = [r for r in csv.DictReader(lines) if r['Metric Name'] == 'gpu__time_duration.sum']
names = [x['Kernel Name'] for x in rows]; us = [float(x['Metric Value']) / 1000 for x in rows]
_, blocks = describe_blocks(parse_model_config(configs.cfg_path(cfg)))
convs = [b for b in blocks if b['type'] == 'convolutional']
idx = [i for i, n in enumerate(names) if 'conv_first' in n]
start = idx[-1]
# spatial sizes
hw = []; s = S
for i, b in enumerate(blocks):
    t = b['type']
    if t == 'convolutional': s = (s + 2 * ((b['size'] - 1) // 2) - b['size']) // b['stride'] + 1
    elif t == 'maxpool': s = s // 2 if b['stride'] == 2 else s
    elif t == 'upsample': s *= 2
    elif t == 'route': s = hw[b['layers'][0]]
    elif t == 'shortcut': s = hw[i - 1]
    hw.append(s)
ci = [i for i, b in enumerate(blocks) if b['type'] == 'convolutional']
g = OrderedDict(); tot = 0; tot_ideal = 0
for j, bi in enumerate(ci):
    b = blocks[bi]; t = us[start + j]; tot += t
    res = bi + 1 < len(blocks) and blocks[bi + 1]['type'] == 'shortcut'
    o = hw[bi]; i_ = o * b['stride']
    gf = 2.0 * N * o * o * b['filters'] * b['cin'] * b['size'] ** 2 / 1e9
    gb = (2.0 * N * (i_ * i_ * b['cin'] + o * o * b['filters'] * (2 if res else 1)) + 2.0 * b['filters'] * b['cin'] * b['size'] ** 2) / 1e9
    if bi == 0: gb = (4.0 * N * i_ * i_ * 3 + 2.0 * N * o * o * b['filters']) / 1e9
    ideal = max(gf / 1404.7 * 1e3, gb / 6535.7 * 1e6)   # MEASURED_PEAKS.json: sustained bf16 TF, copy GB/s
    tot_ideal += ideal
    key = (o, b['size'], b['stride'], b['cin'], b['filters'], res, names[start + j].split('<')[0].split('::')[-1] + '<' + names[start + j].split('<')[1].split('>')[0] + '>')
    e = g.setdefault(key, [0, 0, 0, 0]); e[0] += t; e[1] += 1; e[2] += ideal; e[3] += gf
print('conv total us %.1f ideal %.1f' % (tot, tot_ideal))
for k, v in g.items():
    print('%-70s us=%7.1f n=%2d ideal=%6.1f eff=%.2f lost=%6.1f TF=%5.0f' % (k, v[0], v[1], v[2], v[2] / v[0], v[0] - v[2], v[3] / v[0] * 1e3))
