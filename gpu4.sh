mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.txt 2> gpurun_out/bench.err
ncu --set full --clock-control none --import-source on -k regex:sample_kernel -s 1 -c 1 -o gpurun_out/prof_sample python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 1 -c 1 -o gpurun_out/prof_score2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu4.log 2>&1
tail -12 gpurun_out/pytest_gpu.txt; cat gpurun_out/bench.txt; tail -5 gpurun_out/bench.err
