set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r01_pytest.log
python bench.py --steps 30 --warmup 5 > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
python bench.py --steps 30 --warmup 5 --layout nchw --no-cpu-baseline > gpurun_out/r01_bench_nchw.json 2>> gpurun_out/r01_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mghs_pool_nhwc -s 3 -c 2 -o gpurun_out/r01_pool_fwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mghs_pool_bwd -s 3 -c 1 -o gpurun_out/r01_pool_bwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline >> gpurun_out/r01_ncu_full.log 2>&1
cat gpurun_out/r01_pytest.log gpurun_out/r01_bench.json
