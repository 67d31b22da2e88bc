"""TEST INFRASTRUCTURE -- generates tests/golden/dense_*.npz by running the UNMODIFIED
reference dense modules (HeightNet, MGHS.depth_net, SFA, predictor imported from
/root/reference through oracle/ref_loader.py) on seeded weights and inputs.  Build container only:

    python -m oracle.make_golden_dense

Weights and inputs are NOT stored: tests regenerate them with
oracle.dense_oracle.seeded_state_dict / seeded_tensor and verify the SHA first.
"""
import hashlib
import os

import numpy as np
import torch

from . import dense_oracle as DO
from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')

# (BN images, fH, fW) of the HeightNet / depth head case; (B, Dy, Dx) of the SFA / predictor case
HN_SHAPE = (2, 16, 44)
BEV_SHAPE = (1, 24, 40)


def sha_sd(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def build_reference_modules():
    ns = ref_loader.load_reference()
    hn = ns.HeightNet(256, 256, 65).eval()
    sfa = ns.SFA(512, 256).eval()
    head = ns.predictor(in_dim=256, out_dim=256, Dz=16, num_classes=18, use_predicter=True,
                        class_balance=False, loss_occ=None).eval()
    depth_net = torch.nn.Conv2d(256, 44 + 64, 1)
    return hn, sfa, head, depth_net


def inputs():
    BN, fH, fW = HN_SHAPE
    B, Dy, Dx = BEV_SHAPE
    x = DO.seeded_tensor((BN, 256, fH, fW), 11)
    mlp = DO.seeded_tensor((1, BN, 27), 12, scale=4.0)
    bev = DO.seeded_tensor((B, 512, Dy, Dx), 13)
    return x, mlp, bev


def main():
    hn, sfa, head, depth_net = build_reference_modules()
    sds = {}
    for name, m, seed in (('heightnet', hn, 21), ('sfa', sfa, 22), ('predictor', head, 23), ('depth_net', depth_net, 24)):
        sd = DO.seeded_state_dict(m, seed)
        m.load_state_dict(sd)
        sds[name] = sd
    x, mlp, bev = inputs()
    with torch.no_grad():
        height = hn(x, mlp)
        y = depth_net(x)
        fused = sfa(bev)
        occ = head(fused)
    np.savez_compressed(
        os.path.join(OUT, 'dense_modules.npz'),
        height=height.numpy(), depth_net=y.numpy(), sfa=fused.numpy(), occ=occ.numpy(),
        **{'sha_' + k: sha_sd(v) for k, v in sds.items()},
        input_sha=hashlib.sha256(b''.join(t.numpy().tobytes() for t in (x, mlp, bev))).hexdigest())
    print('wrote dense_modules.npz', height.shape, y.shape, fused.shape, occ.shape)


if __name__ == '__main__':
    main()
