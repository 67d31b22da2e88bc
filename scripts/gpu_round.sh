set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 95 -c 1 -o gpurun_out/r01_wgrad_sfa python scripts/bench_conv.py wgrad > gpurun_out/r01_ncu_wgrad.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_igemm2 -s 95 -c 1 -o gpurun_out/r01_conv3x3_sfa python scripts/bench_conv.py bf16 > gpurun_out/r01_ncu_conv.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r01_launches_infer_final.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-train --no-encoders > gpurun_out/r01_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mghs_pool_stream -s 3 -c 1 -o gpurun_out/r01_pool_fwd_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-train --no-encoders > gpurun_out/r01_ncu_full.log 2>&1
ls -la gpurun_out | tail -6
