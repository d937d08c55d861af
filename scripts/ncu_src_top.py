"""Top stall locations of an `ncu --page source --csv` export, per kernel.  usage: ncu_src_top.py src.csv [ntop]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'rows': []}
        blocks.append(cur)
    elif cur is not None and r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] and len(r) >= len(cur['hdr']) - 2:
        cur['rows'].append(r)
seen = set()
for b in blocks:
    if b['name'] in seen:
        continue
    seen.add(b['name'])
    h = b['hdr']
    isrc, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
    stall_cols = [i for i, k in enumerate(h) if k.startswith('stall_') and 'Not' not in k]
    tot = sum(int(r[isamp] or 0) for r in b['rows'])
    totex = sum(int(r[iex] or 0) for r in b['rows'])
    print('=====', b['name'], 'samples', tot, 'warp-instructions', totex)
    agg = {}
    for r in b['rows']:
        for i in stall_cols:
            agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
    print('   stall totals:', sorted(((v, k[6:]) for k, v in agg.items() if v), reverse=True)[:8])
    order = sorted(range(len(b['rows'])), key=lambda n: -int(b['rows'][n][isamp] or 0))[:ntop]
    for n in order:
        r = b['rows'][n]
        st = sorted(((int(r[i] or 0), h[i][6:]) for i in stall_cols), reverse=True)[:2]
        print('%5d %6s %5.1f%% ex=%9s  %-72s %s' % (n, r[isamp], 100.0 * int(r[isamp]) / max(tot, 1), r[iex], r[isrc][:72], st))
