# launch list of one training step (BatchNorm on batch statistics) + the graph-timed quick bench (one gpurun call)
mkdir -p gpurun_out
TAG=${1:-x}
NCU="ncu --profile-from-start off --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_train_$TAG.csv python scripts/bench_train.py 3 batch > gpurun_out/ncu_train_$TAG.log 2>&1
timeout 900 python bench.py --quick --no-dhdl > gpurun_out/bench_quick_$TAG.json 2> gpurun_out/bench_quick_$TAG.err
python - <<PY
import json
for l in open('gpurun_out/bench_quick_$TAG.json'):
    if l.startswith('{'):
        d = json.loads(l)
        inf = d['extras']['inference']
        print('train ms', d['ms_per_step'], 'train+enc', d['extras']['train_step_with_encoders']['ms_per_step'], 'infer ms', inf['ms_per_step'], inf.get('stage_ms'), 'enc', d['extras']['inference_with_encoders']['ms_per_step'])
PY
