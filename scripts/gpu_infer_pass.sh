# launch list of one inference step + the graph-timed bench (one gpurun call)
mkdir -p gpurun_out
TAG=${1:-x}
NCU="ncu --profile-from-start off --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_infer_$TAG.csv python scripts/bench_infer.py 3 bf16 > gpurun_out/ncu_infer_$TAG.log 2>&1
timeout 900 python bench.py --quick > gpurun_out/bench_quick_$TAG.json 2> gpurun_out/bench_quick_$TAG.err
python - <<PY
import json
for l in open('gpurun_out/bench_quick_$TAG.json'):
    if l.startswith('{'):
        d = json.loads(l)
        inf = d['extras']['inference']
        print('train ms', d['ms_per_step'], 'infer ms', inf['ms_per_step'], inf.get('stage_ms'), 'enc', d['extras']['inference_with_encoders']['ms_per_step'])
PY
