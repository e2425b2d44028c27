ncu --set full --clock-control none --import-source on --kernel-id ::regex:k_matern_mma:12 -o gpurun_out/r1_prof_mma9_box140 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_box.log 2>&1
tail -3 gpurun_out/ncu_full_box.log | cut -c1-200
