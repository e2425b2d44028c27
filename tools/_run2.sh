for cfg in 32561; do
MBG_MMA_CFG=$cfg ncu --set full --clock-control none --import-source on -k regex:k_matern_mma -c 1 -o gpurun_out/prof_mma_q_$cfg python bench.py --workload trimers64 --frames 1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mma_q_$cfg.log 2>&1
done
