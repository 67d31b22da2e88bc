set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r01_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_full.json 2> gpurun_out/r01_bench_full.err; tail -3 gpurun_out/r01_bench_full.err; cat gpurun_out/r01_bench_full.json
ncu --set full --clock-control none --import-source on -k regex:mghs_pool_stream -s 3 -c 1 -o gpurun_out/r01_pool_fwd_v4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r01_ncu_full.log 2>&1
ls -la gpurun_out
