# round-2 profiling pass (one gpurun call): launch lists, the fused tail's --set full capture, pool sweep + write ceiling
mkdir -p gpurun_out
NCU="ncu --profile-from-start off --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_train.csv python scripts/bench_train.py 3 batch > gpurun_out/r02_ncu_train.log 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_infer.csv python scripts/bench_infer.py 3 bf16 > gpurun_out/r02_ncu_infer.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:predictor_tail -o gpurun_out/r02_tail_full -f python scripts/bench_infer.py 3 bf16 > gpurun_out/r02_ncu_tail.log 2>&1
timeout 600 python scripts/bench_pool.py V=4 V=4,PROBE=1 V=4 V=4,PROBE=1 V=4,ZT=8192 V=4,ZT=16384 V=4,NCH=8192 V=4,NCH=32768 V=4,THREADS=64,PERSM=8 V=4,MINB=5,PERSM=5 V=4,CELLCOST=4 V=4,CELLCOST=16 2>&1 | grep -v Warning | tee gpurun_out/r02_pool_sweep1.txt
