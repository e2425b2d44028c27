python -m mbgdml_b200.build >/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_matern_mma -c 1 -o gpurun_out/prof_mma_r python bench.py --workload trimers64 --frames 1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mma_r.log 2>&1
