"""Inference step of the hot path alone (eager launches, for ncu launch lists / --set full captures).  Usage:
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
    python scripts/bench_infer.py [steps] [precision] [encoders] [images]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dhd_b200.pipeline import HotPathStep  # noqa: E402
from dhd_b200 import synth as O  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
precision = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
encoders = len(sys.argv) > 3 and sys.argv[3] not in ('0', 'no')
images = len(sys.argv) > 4 and sys.argv[4] not in ('0', 'no')
cfg, B = O.DHD_S, 4
step = HotPathStep(cfg, B, precision=precision, use_graph=False, encoders=encoders, images=images)
host = step.make_host_inputs(O.synthetic_rig(B, cfg['ncams'], cfg['input_size'], seed=100), seed=100)
step.alloc_static(host)
step.upload(host)
for _ in range(max(steps - 1, 0)):
    step.run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()                      # ncu --profile-from-start off: the last step only
step.run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('occ classes', int(step.occ.max()) + 1)
