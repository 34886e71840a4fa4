#!/bin/bash
mkdir -p gpurun_out
for c in j2_plate tet10 heat_tet4 hex8; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 60 --csv --log-file gpurun_out/r2g_launches_$c.csv python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2g_$c.log 2>&1
done
python - <<'PY'
import csv, glob
for f in sorted(glob.glob("gpurun_out/r2g_launches_*.csv")):
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    if not rows: print(f, "empty"); continue
    hdr = rows[0]; ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    print(f)
    agg = {}
    for r in rows[1:]:
        try: v = float(r[iv].replace(",", ""))
        except Exception: continue
        k = r[ik][:90]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    iu = hdr.index("Metric Unit"); unit = rows[1][iu]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"   {n:3d} x {t/n:12.1f} {unit}  {k}")
PY
