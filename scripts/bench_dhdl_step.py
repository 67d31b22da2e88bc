"""One DHD-L detector step (BASELINE configs[4], dhd_b200.detector_step.DetectorStep) under cudaProfilerStart/Stop, for ncu
launch lists.  Usage: ncu --profile-from-start off ... python scripts/bench_dhdl_step.py [infer|train] [dhd_l|dhd_s_images]
(dhd_s_images: the DHD-S detector from camera images, image backbone + neck trained)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dhd_b200 import synth  # noqa: E402
from dhd_b200.detector_step import DetectorStep  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'infer'
which = sys.argv[2] if len(sys.argv) > 2 else 'dhd_l'
if which == 'dhd_s_images':
    step = DetectorStep(synth.dhd_s_model_cfg('bf16', images=True), 4, seed=0)
else:
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
