set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_launches_infer.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-train > gpurun_out/r01_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 4 -c 1 -o gpurun_out/r01_wgrad python scripts/bench_conv.py wgrad > gpurun_out/r01_ncu_wgrad.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_igemm2 -s 16 -c 1 -o gpurun_out/r01_conv3x3 python scripts/bench_conv.py bf16 > gpurun_out/r01_ncu_conv.log 2>&1
ls -la gpurun_out | tail -8
