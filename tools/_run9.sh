python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline > gpurun_out/b3_box.json 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/b3_box.json').read().strip().splitlines()[-1]); print('box', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['contract_share_of_step'], d['e2e'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/l3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; grep -E "k_cull|k_compact|k_scan" gpurun_out/l3.csv | awk -F'","' '{print $5, $NF}' | head -12
