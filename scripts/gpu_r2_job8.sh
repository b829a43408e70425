#!/bin/bash
out=gpurun_out; mkdir -p $out; tag=r2j8
for dt in f32 f64; do
timeout 300 python scripts/graded_launches.py $dt 4 > $out/graded_plain_${dt}_$tag.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/graded_launches_${dt}_$tag.csv python scripts/graded_launches.py $dt 4 > /dev/null 2>&1
echo "ncu_rc=$?"
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('$out/graded_launches_${dt}_$tag.csv')) if len(r)>14 and r[0].isdigit()]
tot=collections.Counter(); cnt=collections.Counter()
for r in rows:
    n=r[4].split('(')[0][:70]; tot[n]+=float(r[14])/1e3; cnt[n]+=1
print('$dt', 'total us', round(sum(tot.values())))
for n,v in tot.most_common(22): print(f'  {v:9.1f} us  x{cnt[n]:3d}  {n}')
PY
done
