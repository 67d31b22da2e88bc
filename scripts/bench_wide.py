"""The widened inference step (real encoders) alone, eager launches, for ncu launch lists."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dhd_b200 import synth as O  # noqa: E402
from dhd_b200.pipeline import HotPathStep  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg, B = O.DHD_S, 4
st = HotPathStep(cfg, B, precision='bf16', use_graph=False, encoders=True)
host = st.make_host_inputs(O.synthetic_rig(B, cfg['ncams'], cfg['input_size'], seed=100), seed=100)
st.alloc_static(host)
st.upload(host)
for _ in range(steps):
    st._front(); st._pool(); st._back()
torch.cuda.synchronize()
