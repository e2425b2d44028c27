python bench.py > gpurun_out/r1_final_box140.json 2> gpurun_out/r1_final_box140.err; tail -c 600 gpurun_out/r1_final_box140.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_final_box140_reference.json 2>&1; tail -c 400 gpurun_out/r1_final_box140_reference.json
python bench.py --workload trimers64 --no-cpu-baseline --frames 1 > gpurun_out/r1_final_trimers64.json 2>&1
python bench.py --workload monomers --no-cpu-baseline > gpurun_out/r1_final_monomers.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_final_launches_box140.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-id ::regex:k_matern_mma:12 -o gpurun_out/r1_final_prof_mma9_box140 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_box.log 2>&1
python __graft_entry__.py --smoke 2>&1 | tail -2
