mkdir -p gpurun_out
timeout 300 python scripts_score_variants.py > gpurun_out/score_variants.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench2.txt 2> gpurun_out/bench2.err
cat gpurun_out/score_variants.txt; cat gpurun_out/bench2.txt | cut -c1-600; tail -5 gpurun_out/bench2.err
