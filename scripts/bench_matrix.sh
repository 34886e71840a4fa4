for n in 100 200; do for mma in 0 1; do for s in 1 0; do
FDK_MMA=$mma FDK_SMALL_CTA=$s python bench.py --n $n --steps 5 --no-cpu-baseline > gpurun_out/mx.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/mx.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("n=$n mma=$mma small=$s  ms=%.3f  Melem/s=%.1f frac=%.4f e2e=%.1f" % (d["ms_per_step"], d["value"], d["roofline"]["frac"], d["e2e"]["value"]))
else:
    print("n=$n mma=$mma small=$s FAILED"); print(open("gpurun_out/mx.log").read()[-600:])
PY
done; done; done
