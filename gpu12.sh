mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt
timeout 300 python scripts_refine_groups.py > gpurun_out/refine_groups.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.txt 2> gpurun_out/bench.err
tail -4 gpurun_out/pytest_gpu.txt; cat gpurun_out/refine_groups.txt; cat gpurun_out/bench.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms_last_step'], d['score_launch'])"; tail -5 gpurun_out/bench.err
