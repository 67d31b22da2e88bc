# final single-GPU evidence of the round (one gpurun call): bench line, reference arm, other configs, launch lists
mkdir -p gpurun_out
NCU="ncu --profile-from-start off --clock-control none"
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; echo "ref rc=$?"
timeout 900 python scripts/bench_configs.py > gpurun_out/r02_configs_1gpu.json 2> gpurun_out/r02_configs_1gpu.err; echo "configs rc=$?"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_infer.csv python scripts/bench_infer.py 3 bf16 > /dev/null 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_train.csv python scripts/bench_train.py 3 batch > /dev/null 2>&1
python - <<PY
import json
for l in open('gpurun_out/r02_bench_1gpu.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'pool ms', d['roofline']['kernel_ms_avg'])
        ex = d['extras']
        print('train+enc', ex['train_step_with_encoders']['ms_per_step'], 'infer', ex['inference']['ms_per_step'], ex['inference'].get('stage_ms'), 'e2e', ex['inference']['e2e']['value'])
        print('fp32', ex['inference_fp32']['ms_per_step'], 'enc', ex['inference_with_encoders']['ms_per_step'], 'dhdl', ex['dhd_l_view_transformer']['ms_per_step'], 'ref', ex['reference_cuda_path'].get('speedup'))
        print('cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
PY
