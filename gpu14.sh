mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.txt 2> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:prefilter_kernel -s 9 -c 1 -o gpurun_out/prof_prefilter3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu5.log 2>&1
tail -4 gpurun_out/pytest_gpu.txt; cat gpurun_out/bench.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms_last_step'], d['score_launch'], d['roofline']['frac'])"; tail -5 gpurun_out/bench.err
