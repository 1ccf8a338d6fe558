#!/bin/bash
# per-kernel durations of one bench step (ncu launch list): gpu_launches.sh tag [nx] [lib]
TAG=${1:-launches}; NX=${2:-256}
OUT=gpurun_out/$TAG; mkdir -p $OUT
[ -n "$3" ] && export NDSPMHD_B200_LIB=$PWD/$3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file $OUT/launches.csv python bench.py --nx $NX --steps 1 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("$OUT/launches.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0].replace('void ','').replace('ndk::','').replace('<unnamed>::','')
    t=float(r[-1])/1e6
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=t
tot=sum(a[1] for a in agg.values())
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{t:8.3f} ms {100*t/tot:5.1f}%  x{n:3d}  {k}")
print(f"{tot:8.3f} ms total over {len(rows)} launches")
PY
