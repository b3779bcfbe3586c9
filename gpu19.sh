mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1
timeout 300 python bench.py > gpurun_out/bench.txt 2> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.txt 2> gpurun_out/bench_ref.err
tail -5 gpurun_out/pytest_gpu.txt; tail -2 gpurun_out/smoke.txt | cut -c1-400; cat gpurun_out/bench.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['steps'], d['e2e']['value'], d['clocks'], d['cpu_baseline'], d['batched_e2e']['value'])"; cut -c1-300 gpurun_out/bench_ref.txt; tail -3 gpurun_out/bench.err
