#!/bin/bash
OUT=gpurun_out/r2s; mkdir -p $OUT
MODES=1 REPS=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv3_wgrad_rs --launch-skip 0 -c 1 -o $OUT/rs2 -f python scripts/bench_wgrad.py > $OUT/ncu2.log 2>&1
ncu -i $OUT/rs2.ncu-rep --page source --csv --print-source=sass > $OUT/rs2.source.csv 2>/dev/null
rm -f $OUT/rs2.ncu-rep
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/r2s/rs2.source.csv')))[2:]
tot=sum(int(r[2] or 0) for r in rows if len(r)>5)
print("total samples",tot)
for i,r in enumerate(rows):
    if len(r)<6: continue
    s=r[1]
    if ('TRYWAIT' in s or 'UTCHMMA' in s or 'UTCBAR' in s or 'UBLKCP' in s or 'UTMALDG' in s or 'LDGSTS' in s or 'ARRIVE' in s) and int(r[5] or 0)>0:
        print(i,r[2],r[5],s.strip()[:80])
P
