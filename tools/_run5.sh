for fr in 1 2 8; do
for cfg in 32561; do
  MBG_MMA_CFG=$cfg python bench.py --workload trimers64 --frames $fr --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_tmp.txt 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/b_tmp.txt") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("frames $fr cfg $cfg", d["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["peak"], d["clocks"])
else:
    print(open("gpurun_out/b_tmp.txt").read()[-2000:])
PY
done
done
