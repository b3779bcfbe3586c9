mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt
timeout 300 python scripts_backward_timing.py > gpurun_out/backward_timing.txt 2>&1
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_forward.py tests/test_gpu_backward.py -m gpu -x -q -k "known_answer or test_backward_accumulates" > gpurun_out/sanitizer_racecheck.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.txt 2> gpurun_out/bench.err
tail -4 gpurun_out/pytest_gpu.txt; cat gpurun_out/backward_timing.txt; tail -6 gpurun_out/sanitizer_racecheck.txt; cat gpurun_out/bench.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms_last_step'])"
