for f in 0 1 2 4 7; do
SC_TOWER_DEBUG=$f ncu --kernel-name regex:clip_tower --launch-skip 64 --launch-count 9 --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_dbg$f.csv python scripts/profile_clip_tower.py 64 fp16 > /dev/null 2>&1
echo "flags=$f"; python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02f_dbg$f.csv')) if len(r)>10]
idx={h:i for i,h in enumerate(rows[0])}
print([round(float(r[idx['Metric Value']].replace(',',''))/1e3,1) for r in rows[1:]])
PY
done
