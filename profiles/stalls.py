"""Per-opcode stall summary from an ncu report's source page. Usage: python profiles/stalls.py <rep> [kernel-index ...]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
which = [int(a) for a in sys.argv[2:]] or None
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
blocks = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'rows': []}; blocks.append(cur); continue
    if cur is None: continue
    if r and r[0] == 'Address': cur['hdr'] = r; continue
    cur['rows'].append(r)
keys = ['stall_wait','stall_math','stall_short_sb','stall_selected','stall_not_selected','stall_barrier','stall_lg','stall_long_sb','stall_dispatch','stall_mio','stall_branch_resolving']
for bi, b in enumerate(blocks):
    if which and bi not in which: continue
    h = b['hdr']; idx = {k: i for i, k in enumerate(h)}
    agg = collections.defaultdict(collections.Counter)
    for r in b['rows']:
        if len(r) < len(h): continue
        src = r[idx['Source']].strip().split()
        op = (src[1] if src and src[0].startswith('@') else (src[0] if src else '?')).split('.')[0]
        agg[op]['samples'] += int(r[idx['# Samples']] or 0)
        agg[op]['inst'] += int(r[idx['Instructions Executed']] or 0)
        for k in keys: agg[op][k] += int(r[idx[k]] or 0)
    ts = sum(v['samples'] for v in agg.values()) or 1; ti = sum(v['inst'] for v in agg.values()) or 1
    print(f"===== [{bi}] {b['name'][:80]}  samples={ts} inst={ti}")
    print(f"{'op':8} {'smp%':>6} {'inst%':>6} " + ' '.join(f"{k[6:11]:>5}" for k in keys))
    tot = collections.Counter()
    for op, v in sorted(agg.items(), key=lambda kv: -kv[1]['samples'])[:14]:
        print(f"{op:8} {100*v['samples']/ts:6.2f} {100*v['inst']/ti:6.2f} " + ' '.join(f"{100*v[k]/ts:5.1f}" for k in keys))
    for v in agg.values():
        for k in keys: tot[k] += v[k]
    print(f"{'TOTAL':8} {'':6} {'':6} " + ' '.join(f"{100*tot[k]/ts:5.1f}" for k in keys))
