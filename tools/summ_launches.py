"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/summ_launches.py <csv> [top]"""
import csv, collections, sys
path=sys.argv[1]
lines=[l for l in open(path) if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.defaultdict(lambda:[0,0.0]); tot=0; n=0
for row in r:
    name=row['Kernel Name']; v=float(row['Metric Value'].replace(',','')); unit=row['Metric Unit']
    if unit=='ns': v/=1e3
    elif unit=='ms': v*=1e3
    agg[name][0]+=1; agg[name][1]+=v; tot+=v; n+=1
print("launches %d, total %.1f us"%(n,tot))
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 25]:
    print("%6d %10.1f us %5.1f%%  avg %7.1f us  %s"%(c,t,100*t/tot,t/c,k[:100]))
