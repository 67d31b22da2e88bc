set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 > gpurun_out/r01_bench_full_bf16.json 2> gpurun_out/r01_bench_full.err; tail -3 gpurun_out/r01_bench_full.err; cat gpurun_out/r01_bench_full_bf16.json
python bench.py --steps 10 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/r01_bench_full_fp32.json 2>> gpurun_out/r01_bench_full.err; cat gpurun_out/r01_bench_full_fp32.json
python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r01_bench_full_nograph.json 2>> gpurun_out/r01_bench_full.err; cat gpurun_out/r01_bench_full_nograph.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_full.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r01_ncu_bench.log 2>&1
