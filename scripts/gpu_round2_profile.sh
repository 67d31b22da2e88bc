# round-2 profiling pass (one gpurun call)
mkdir -p gpurun_out
NCU="ncu --profile-from-start off --clock-control none"
timeout 600 python -m pytest tests/test_tail_gpu.py tests/test_pool_gpu.py tests/test_hotpath_gpu.py tests/test_dense_gpu.py -q -x > gpurun_out/r02_tests_b.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_tests_b.log
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_infer_b.csv python scripts/bench_infer.py 3 bf16 > gpurun_out/r02_ncu_infer.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:predictor_tail -o gpurun_out/r02_tail_full_b -f python scripts/bench_infer.py 3 bf16 > gpurun_out/r02_ncu_tail.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:conv_igemm2 --launch-skip 20 --launch-count 6 -o gpurun_out/r02_conv_sfa_full -f python scripts/bench_infer.py 3 bf16 > gpurun_out/r02_ncu_conv.log 2>&1
timeout 600 python scripts/bench_pool.py NCH=32768 NCH=16384 NCH=65536 NCH=32768,CELLCOST=16 NCH=65536,CELLCOST=16 NCH=32768,CELLCOST=12 PROBE=1 NCH=32768 2>&1 | grep -v Warning | tee gpurun_out/r02_pool_sweep2.txt
