python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --workload trimers64 --frames 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b2_tri.txt 2>&1; tail -1 gpurun_out/b2_tri.txt | cut -c1-1500
python bench.py --workload monomers --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b2_mono.txt 2>&1; tail -1 gpurun_out/b2_mono.txt | cut -c1-2500
( time python bench.py ) > gpurun_out/b2_box.txt 2>&1; tail -5 gpurun_out/b2_box.txt
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/b2_ref.txt 2>&1; tail -5 gpurun_out/b2_ref.txt
