set -x
python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | tail -15
python -m pytest tests -m gpu -q --deselect tests/test_conv_gpu.py 2>&1 | tail -8
python scripts/bench_conv.py bf16 fp32 2>&1 | tee gpurun_out/r01_bench_conv_v2b.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_full2_bf16.json 2> gpurun_out/r01_bench_full2.err; tail -3 gpurun_out/r01_bench_full2.err; cat gpurun_out/r01_bench_full2_bf16.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_full2.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r01_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mghs_pool_nhwc -s 3 -c 1 -o gpurun_out/r01_pool_fwd_v3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r01_ncu_full.log 2>&1
