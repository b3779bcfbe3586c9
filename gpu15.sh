mkdir -p gpurun_out
timeout 300 python scripts_backward_timing.py > gpurun_out/backward_timing.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_forward.py tests/test_gpu_backward.py -m gpu -x -q -k "known_answer or backward_matches_oracle or stride0" > gpurun_out/sanitizer_memcheck.txt 2>&1
tail -15 gpurun_out/sanitizer_memcheck.txt > gpurun_out/sanitizer_memcheck_tail.txt
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_forward.py -m gpu -x -q -k "known_answer" > gpurun_out/sanitizer_racecheck.txt 2>&1
cat gpurun_out/backward_timing.txt; tail -8 gpurun_out/sanitizer_memcheck.txt; tail -8 gpurun_out/sanitizer_racecheck.txt
