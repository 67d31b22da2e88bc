# Round-end evidence run (one GPU): launch list of the timed bench steps, ncu --set full of the stereo kernel.
set -x
mkdir -p gpurun_out
DHD_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_infer_final.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-train --no-encoders --no-dhdl > gpurun_out/r01_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stereo_cost -s 13 -c 1 -o gpurun_out/r01_stereo_final -f python scripts/bench_stereo.py --no-ref --bn 4 > gpurun_out/r01_ncu_stereo.log 2>&1
tail -2 gpurun_out/r01_ncu_stereo.log
wc -l gpurun_out/r01_launches_infer_final.csv
