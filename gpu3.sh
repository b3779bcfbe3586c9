mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.txt 2> gpurun_out/bench.err
tail -12 gpurun_out/pytest_gpu.txt; cat gpurun_out/bench.txt; tail -5 gpurun_out/bench.err
