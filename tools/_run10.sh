python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
python bench.py --workload trimers64 --no-cpu-baseline --frames 2 > gpurun_out/b4_tri.json 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/b4_tri.json').read().strip().splitlines()[-1]); print('tri', d['value'], d['roofline']['frac'], d['roofline']['frac_of_fma_pipe_peak'], d['roofline']['avg_launch_ms'])"
done
python bench.py --no-cpu-baseline > gpurun_out/b4_box.json 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/b4_box.json').read().strip().splitlines()[-1]); print('box', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_of_fma_pipe_peak'], d['roofline']['avg_launch_ms'], d['roofline']['contract_share_of_step'])"
