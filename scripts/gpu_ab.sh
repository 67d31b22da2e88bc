# A/B of an environment switch on the quick bench: bash scripts/gpu_ab.sh VAR [value_a value_b]
VAR=$1; A=${2:-0}; B=${3:-1}
for v in $A $B $A $B; do
  env $VAR=$v timeout 600 python bench.py --quick --no-dhdl --no-encoders --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('$VAR=$v', 'train ms %.4f' % d['ms_per_step'], 'infer ms %.4f' % d['extras']['inference']['ms_per_step'], 'clk', d['clocks']['sm_mhz'])
"
done
