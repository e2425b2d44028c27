python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in 0 23841 25121 32561 31282 21283 41282 42561 21922; do
  MBG_MMA_CFG=$cfg python bench.py --workload trimers64 --frames 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_mma_$cfg.txt 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/b_mma_$cfg.txt") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("cfg $cfg", d["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"])
else:
    print(open("gpurun_out/b_mma_$cfg.txt").read()[-2000:])
PY
done
