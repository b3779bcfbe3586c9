mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_multi.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench2.txt 2> gpurun_out/bench2.err
cat gpurun_out/pytest_multi.txt; cat gpurun_out/bench2.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"; tail -3 gpurun_out/bench2.err
