timeout 600 python scripts/bench_pool.py V=4,PERSM=4 V=4,PERSM=5 V=4,PERSM=4 V=4,PERSM=5 V=4,PERSM=4,PREFETCH=0 V=4,PERSM=3 V=2 2>&1 | grep -v Warning | tee gpurun_out/r01_pool_sweep14.txt
