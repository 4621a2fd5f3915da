"""Profiling aid: list the SASS instructions with the most warp-stall samples per kernel from `ncu --page source --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 45
ks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {'name': r[1], 'rows': []}; ks.append(cur); continue
    if r and r[0] == "Address": cur['hdr'] = r; continue
    if cur is not None and r: cur['rows'].append(r)
for ki in want:
    k = ks[ki]
    h = k['hdr']; si = h.index('# Samples'); src = h.index('Source')
    tot = sum(int(r[si]) for r in k['rows'])
    print(k['name'][:70], 'total samples', tot, 'ninstr', len(k['rows']))
    idx = sorted(range(len(k['rows'])), key=lambda i: -int(k['rows'][i][si]))[:topn]
    for i in sorted(idx):
        r = k['rows'][i]
        stalls = {h[j]: int(r[j]) for j in range(h.index('stall_barrier'), h.index('stall_wait') + 1) if r[j] not in ('', '0')}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
        print(i, r[si], r[src].strip()[:80], top)
