set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r01_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python scripts/bench_conv.py bf16 fp32 2>&1 | tee gpurun_out/r01_bench_conv_v2.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r01_bench_full.json 2> gpurun_out/r01_bench_full.err; tail -3 gpurun_out/r01_bench_full.err; cat gpurun_out/r01_bench_full.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_full.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r01_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mghs_pool_nhwc -s 3 -c 1 -o gpurun_out/r01_pool_fwd_v3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r01_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_igemm2 -s 20 -c 2 -o gpurun_out/r01_conv2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r01_ncu_conv2.log 2>&1
ls -la gpurun_out
