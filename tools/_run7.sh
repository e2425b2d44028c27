python bench.py > gpurun_out/r1_bench_box140.json 2> gpurun_out/r1_bench_box140.err; tail -c 1500 gpurun_out/r1_bench_box140.json
python bench.py --workload trimers64 --no-cpu-baseline --frames 2 > gpurun_out/r1_bench_trimers64.json 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/r1_bench_trimers64.json').read().strip().splitlines()[-1]); print('tri', d['value'], d['roofline']['frac'], d['roofline']['frac_of_fma_pipe_peak'], d['roofline']['avg_launch_ms'])"
python bench.py --workload trimers64 --no-cpu-baseline --frames 2 --contraction fma > gpurun_out/r1_bench_trimers64_fma.json 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/r1_bench_trimers64_fma.json').read().strip().splitlines()[-1]); print('tri fma', d['value'], d['roofline']['frac'], d['roofline']['frac_of_fma_pipe_peak'], d['roofline']['avg_launch_ms'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_box140.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_matern_mma.*9 -c 1 -o gpurun_out/r1_prof_mma9_box140 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_box.log 2>&1
ls -la gpurun_out | tail -5
