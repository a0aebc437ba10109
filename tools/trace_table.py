"""Tabulates tools/conv_trace.py output (one row per distinct layer shape): python tools/trace_table.py trace.jsonl"""
import json
import sys

rows = [json.loads(l.split(' ', 1)[1]) for l in open(sys.argv[1]) if l.startswith('TRACE')]
seen = {}
keys = ['ctas', 'tiles', 'total', 'total_max', 'setup', 'fill', 'mma', 'w_full', 'w_acc', 'tail', 'p_empty', 'e_tfull', 'e_other', 'e_work']
print('%-30s' % 'layer (hw,k,s,cin,cout,res) xN' + ' '.join('%8s' % k for k in keys))
for r in rows:
    key = (r['hw'], r['k'], r['s'], r['cin'], r['cout'], r['res'])
    seen.setdefault(key, []).append(r)
tot = 0
for key, rs in seen.items():
    r = rs[0]
    tot += sum(x['total_max'] for x in rs)
    print('%-30s' % (str(key).replace(' ', '') + ' x%d' % len(rs)) + ' '.join('%8s' % r[k] for k in keys))
print('sum of total_max over all conv GEMM launches: %d cycles' % tot)
