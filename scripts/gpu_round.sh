set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r01_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/r01_bench_full.json 2> gpurun_out/r01_bench_full.err; tail -3 gpurun_out/r01_bench_full.err; head -c 600 gpurun_out/r01_bench_full.json
