set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r01_pytest_gpu.log
python scripts/bench_conv.py bf16 fp32 2>&1 | tee gpurun_out/r01_bench_conv_v3.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_full.json 2> gpurun_out/r01_bench_full.err; tail -3 gpurun_out/r01_bench_full.err; cat gpurun_out/r01_bench_full.json
