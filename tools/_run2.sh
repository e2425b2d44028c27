set -x
MBG_MMA_CFG=2512 ncu --set full --clock-control none --import-source on -k regex:k_matern_mma -c 1 -o gpurun_out/prof_mma_2512 python bench.py --workload trimers64 --frames 1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mma_2512.log 2>&1
MBG_MMA_CFG=0 ncu --set full --clock-control none --import-source on -k regex:k_matern_mma -c 1 -o gpurun_out/prof_mma_2256 python bench.py --workload trimers64 --frames 1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mma_2256.log 2>&1
ls -la gpurun_out
