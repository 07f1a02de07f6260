import csv, collections, sys
src, title = sys.argv[1], sys.argv[2]
print("# " + title)
with open(src) as f:
    lines=[l for l in f if not l.startswith('==')]
agg=collections.defaultdict(list)
for row in csv.DictReader(lines):
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    v = v/1000 if u=='ns' else (v*1000 if u=='ms' else v)
    agg[row['Kernel Name'].split('(')[0]].append(v)
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print(f"{k[:40]:40s} n={len(v):4d} avg_us={sum(v)/len(v):9.1f} max_us={max(v):9.1f} share={sum(v)/tot:6.3f}")
