"""Training step of the hot path alone (eager launches, for ncu launch lists).  Usage:
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/bench_train.py [steps] [frozen|batch] [encoders]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dhd_b200.pipeline import TrainStep  # noqa: E402
from dhd_b200 import synth as O  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
bn = sys.argv[2] if len(sys.argv) > 2 else 'frozen'          # 'batch': BatchNorm2d in training mode + Dropout
encoders = len(sys.argv) > 3 and sys.argv[3] not in ('0', 'no')
cfg, B = O.DHD_S, 4
ts = TrainStep(cfg, B, bn=bn, encoders=encoders)
host = ts.make_host_inputs(O.synthetic_rig(B, cfg['ncams'], cfg['input_size'], seed=100), seed=100)
ts.alloc_static(host)
ts.upload(host)
for _ in range(max(steps - 1, 0)):
    ts.train_step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()                      # ncu --profile-from-start off: the last step only
ts.train_step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('loss', float(ts.loss[0]), float(ts.loss_height[0]))
