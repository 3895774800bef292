"""Dev helper: split the ncu source-page CSV of one kernel at its barriers and print, per segment, the executed
instruction mix, the share of warp-stall samples and the stall reasons. usage: sass_segments.py src.csv [kernel_index]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
heads = [i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
h = rows[heads[k]]
body = rows[heads[k] + 1:(heads[k + 1] - 1 if k + 1 < len(heads) else len(rows))]
ix = {n: i for i, n in enumerate(h)}
def f(r, n):
    try: return float(r[ix[n]])
    except Exception: return 0.0
seg, cur = [], []
for r in body:
    cur.append(r)
    s = r[ix['Source']]
    if 'BAR.SYNC' in s or 'SYNCS.PHASECHK' in s or 'SYNCS.ARRIVE' in s:
        seg.append(cur); cur = []
seg.append(cur)
tot = sum(f(r, '# Samples') for r in body); toti = sum(f(r, 'Instructions Executed') for r in body)
print('total samples', tot, 'inst', toti)
stalls = ['stall_barrier', 'stall_dispatch', 'stall_long_sb', 'stall_math', 'stall_mio', 'stall_not_selected', 'stall_selected',
          'stall_short_sb', 'stall_wait', 'stall_branch_resolving', 'stall_no_inst', 'stall_lg']
for j, s in enumerate(seg):
    smp = sum(f(r, '# Samples') for r in s); ins = sum(f(r, 'Instructions Executed') for r in s)
    if ins < 1e6 and smp < 100: continue
    ops = collections.Counter()
    for r in s:
        p = r[ix['Source']].split()
        if not p: continue
        op = p[1] if p[0].startswith('@') and len(p) > 1 else p[0]
        ops[op.split('.')[0]] += f(r, 'Instructions Executed')
    st = {n[6:]: int(sum(f(r, n) for r in s)) for n in stalls}
    print(f"seg{j}: n_sass={len(s)} inst={ins/1e6:.1f}M ({100*ins/toti:.1f}%) samples={smp:.0f} ({100*smp/tot:.1f}%) end='{s[-1][ix['Source']][:44]}'")
    print('   ', {k: round(v / 1e6, 1) for k, v in ops.most_common(10)})
    print('   ', {k: v for k, v in st.items() if v})
