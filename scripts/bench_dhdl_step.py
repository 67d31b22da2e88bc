"""One DHD-L detector step (BASELINE configs[4], dhd_b200.detector_step.DetectorStep) under cudaProfilerStart/Stop, for ncu
launch lists.  Usage: ncu --profile-from-start off ... python scripts/bench_dhdl_step.py [infer|train]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dhd_b200 import synth  # noqa: E402
from dhd_b200.detector_step import DetectorStep  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'infer'
step = DetectorStep(synth.dhd_l_model_cfg('bf16'), 2, seed=0)
img_inputs, kw = step.make_inputs(200)
fn = (lambda: step.infer_step(img_inputs)) if mode == 'infer' else (lambda: step.train_step(img_inputs, kw))
for _ in range(2):
    fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')
