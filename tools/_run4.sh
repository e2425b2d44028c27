for nt in 250 1000 4000; do
for cfg in 32561 31282; do
  MBG_MMA_CFG=$cfg python bench.py --workload trimers64 --frames 2 --steps 3 --warmup 3 --no-cpu-baseline --n-train $nt > gpurun_out/b_tmp.txt 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/b_tmp.txt") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("ntrain $nt cfg $cfg", d["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"])
else:
    print(open("gpurun_out/b_tmp.txt").read()[-2000:])
PY
done
done
