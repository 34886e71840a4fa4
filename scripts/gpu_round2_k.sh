#!/bin/bash
# DRAM traffic of the dominant kernel of every configuration at full size (ncu, two metrics), ncu --set full of the hex8 kernel
mkdir -p gpurun_out
for c in hex8 heat_tet4 j2_plate tet10; do
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -c 12 --csv --log-file gpurun_out/r2k_traffic_$c.csv python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2k_$c.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_iso -s 3 -c 1 -o gpurun_out/r2k_ncu_iso python bench.py --edge 100 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_heat_tet4 -s 2 -c 1 -o gpurun_out/r2k_ncu_heat python bench.py --config heat_tet4 --scale 0.1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_ncu_heat.log 2>&1
python - <<'PY'
import csv, glob
for f in sorted(glob.glob("gpurun_out/r2k_traffic_*.csv")):
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    if not rows: print(f, "empty"); continue
    hdr = rows[0]; ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    print(f)
    for r in rows[1:]:
        print("   ", r[ik][:70], r[im], r[iv], r[iu])
PY
