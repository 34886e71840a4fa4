#!/bin/bash
# final kernels of configs [2]-[4]: ncu --set full at reduced size (one launch each) and the DRAM traffic of one launch at full size
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_heat_tet4 -s 2 -c 1 -o gpurun_out/r2ab_ncu_heat python bench.py --config heat_tet4 --scale 0.1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2ab_ncu_heat.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_assemble<" -s 2 -c 1 -o gpurun_out/r2ab_ncu_tet10 python bench.py --config tet10 --scale 0.1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2ab_ncu_tet10.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_iso -s 2 -c 1 -o gpurun_out/r2ab_ncu_r1 python bench.py --config j2_plate --scale 0.1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2ab_ncu_r1.log 2>&1
for c in heat_tet4 j2_plate tet10; do
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -c 14 --csv --log-file gpurun_out/r2ab_traffic_$c.csv python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2ab_$c.log 2>&1
done
ls -la gpurun_out/r2ab_*
