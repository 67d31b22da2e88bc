"""DHD-L inference step through the detector (DHD_stereo.simple_test, B=2), activation fast path on / off.
Usage: python scripts/bench_dhdl_infer_ab.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dhd_b200 import synth  # noqa: E402
from dhd_b200.detector_step import DetectorStep  # noqa: E402

step = DetectorStep(synth.dhd_l_model_cfg('bf16'), 2, seed=0)
img_inputs, _ = step.make_inputs(200)
out = {}
for name, flag in (('act_path', True), ('tensor_path', False), ('act_path_again', True)):
    step.model.act_path = flag
    for _ in range(3):
        step.infer_step(img_inputs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 20
    for _ in range(n):
        step.infer_step(img_inputs)
    e1.record()
    torch.cuda.synchronize()
    out[name] = {'ms_per_step': e0.elapsed_time(e1) / n, 'samples_per_s': 2 * n / (e0.elapsed_time(e1) * 1e-3)}
ok = step.capture_infer(img_inputs)
if ok:
    step.model.act_path = True
    ref = step.infer_step(img_inputs)
    got = step.infer_step_graphed(img_inputs)
    import numpy as np
    same = all(np.array_equal(a, b) for a, b in zip(ref, got))
    for _ in range(3):
        step.infer_step_graphed()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step.infer_step_graphed()
    e1.record()
    torch.cuda.synchronize()
    out['act_path_cuda_graph'] = {'ms_per_step': e0.elapsed_time(e1) / 20, 'samples_per_s': 40 / (e0.elapsed_time(e1) * 1e-3),
                                  'equals_eager': bool(same)}
else:
    out['act_path_cuda_graph'] = {'capture_error': step.capture_error}
print(json.dumps({'config': 'BASELINE configs[4] DHD-L inference through DHD_stereo.simple_test, B=2, one B200 (incl. D2H of the class maps)',
                  **out}))
