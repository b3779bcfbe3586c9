mkdir -p gpurun_out
timeout 120 python scripts_dbg_ws.py > gpurun_out/dbg_ws.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:prefilter_kernel -s 6 -c 1 -o gpurun_out/prof_prefilter python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:refine_kernel -s 1 -c 1 -o gpurun_out/prof_refine python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu6.log 2>&1
cat gpurun_out/dbg_ws.txt
