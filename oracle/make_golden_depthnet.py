"""TEST INFRASTRUCTURE -- generates tests/golden/depthnet.npz by running the UNMODIFIED reference
DepthNet (models/model_utils/depthnet.py:172-415, imported from /root/reference through
oracle/ref_loader.py) on seeded weights and inputs, in the two forms the configs use:

  'mono'    MGHS_Depth of DHD-M: stereo=False, ASPP + DCN                       (DHD-M.py depthnet_cfg)
  'stereo'  MGHS_Stereo of DHD-L: stereo=True, bias=5, use_dcn=False, aspp_mid_channels=96 (DHD-L.py:114-117),
            with the plane-sweep cost volume (calculate_cost_volumn, 310-361) from a synthetic two-frame rig
  'stereo_first' the same net when there is no previous frame (cv_feat_list[0] is None, 389-396)

Build container only:   python -m oracle.make_golden_depthnet
Weights and inputs are NOT stored: tests regenerate them (seeded, exactly representable) and check the SHA.
"""
import hashlib
import math
import os

import numpy as np
import torch

from . import dense_oracle as DO
from . import ref_loader
from .make_golden_dense import sha_sd

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')

# small stand-in of DHD-L: 2 cameras of one sample, 64x176 input -> 1/16 map 4x11, 1/4 stereo map 16x44
B, NCAM = 1, 2
INPUT = (64, 176)
DEPTH_CFG = (1.0, 45.0, 2.0)          # D = 22 hypotheses
C_IN, C_CTX, C_STEREO = 64, 32, 16
C_MID_MONO = 256                      # the DCN of the mono form runs 4 groups of 64 channels, as in DHD-M
BIAS = 5.0


def n_depth():
    lo, hi, st = DEPTH_CFG
    return int(round((hi - lo) / st))


def frustum(downsample):
    """create_frustum (lss_heightmap.py:105-134, sid=False) at `downsample`: (D, fH, fW, 3) = (u, v, d)."""
    H, W = INPUT
    fH, fW = H // downsample, W // downsample
    lo, hi, st = DEPTH_CFG
    d = torch.arange(lo, hi, st, dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
    D = d.shape[0]
    u = torch.linspace(0, W - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
    v = torch.linspace(0, H - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
    return torch.stack((u, v, d), -1)


def rig():
    """Two cameras (front, front-right) of a nuScenes-like rig scaled to the 64x176 crop, and the
    current-camera -> previous-camera transform of a vehicle that drove 1.1 m forward while yawing 2 degrees."""
    H, W = INPUT
    s = W / 1600.0
    intr = torch.tensor([[1266.0, 0.0, 816.0], [0.0, 1266.0, 491.0], [0.0, 0.0, 1.0]])
    intrins = intr.expand(B, NCAM, 3, 3).clone()
    post_rots = torch.zeros(B, NCAM, 3, 3)
    post_trans = torch.zeros(B, NCAM, 3)
    k2s = torch.zeros(B, NCAM, 4, 4)
    base = torch.tensor([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])     # camera axes -> ego axes
    for n, yaw_deg in enumerate((0.0, -55.0)):
        sc = s * (1.0 + 0.01 * n)
        post_rots[0, n] = torch.diag(torch.tensor([sc, sc, 1.0]))
        post_trans[0, n] = torch.tensor([0.0, -140.0 * s * (1.0 + 0.02 * n), 0.0])
        yaw = math.radians(yaw_deg)
        Rz = torch.tensor([[math.cos(yaw), -math.sin(yaw), 0.0], [math.sin(yaw), math.cos(yaw), 0.0], [0.0, 0.0, 1.0]])
        cam2ego = torch.eye(4)
        cam2ego[:3, :3] = Rz @ base
        cam2ego[:3, 3] = torch.tensor([1.5 * math.cos(yaw), 1.5 * math.sin(yaw), 1.5])
        dyaw = math.radians(2.0)
        ego_k2p = torch.eye(4)                                                      # key ego -> previous ego
        ego_k2p[:3, :3] = torch.tensor([[math.cos(dyaw), -math.sin(dyaw), 0.0], [math.sin(dyaw), math.cos(dyaw), 0.0],
                                        [0.0, 0.0, 1.0]])
        ego_k2p[:3, 3] = torch.tensor([1.1, 0.05, 0.0])
        k2s[0, n] = torch.inverse(cam2ego) @ ego_k2p @ cam2ego
    return dict(k2s_sensor=k2s, intrins=intrins, post_rots=post_rots, post_trans=post_trans)


def smooth_feature(shape, seed):
    """Seeded feature map with spatial correlation (a 3x3 box blur of white noise, twice): real stereo features are
    smooth, and on white noise the bilinear sample amplifies fp32 rounding of the sampling coordinate."""
    x = DO.seeded_tensor(shape, seed)
    k = torch.ones(1, 1, 3, 3) / 8.0                 # power-of-two scale keeps the values exactly representable
    for _ in range(2):
        x = torch.nn.functional.conv2d(x.flatten(0, 1)[:, None], k, padding=1).view(shape)
    x = x.clone()
    x[:, :, :2, :5] = 0.0                            # an exactly-zero patch: exercises the `== 0` test of the bias
    return x


def inputs():
    H, W = INPUT
    BN = B * NCAM
    x = DO.seeded_tensor((BN, C_IN, H // 16, W // 16), 31)
    mlp = DO.seeded_tensor((B, NCAM, 27), 32, scale=4.0)
    prev = smooth_feature((BN, C_STEREO, H // 4, W // 4), 33)
    curr = smooth_feature((BN, C_STEREO, H // 4, W // 4), 34)
    return x, mlp, prev, curr


def stereo_metas(prev, curr):
    m = rig()
    m.update(frustum=frustum(4), cv_downsample=4, downsample=16, grid_config=dict(depth=list(DEPTH_CFG)),
             cv_feat_list=[prev, curr])
    return m


def build(ns, stereo):
    if stereo:
        return ns.DepthNet(C_IN, C_IN, C_CTX, n_depth(), use_dcn=False, aspp_mid_channels=32, stereo=True,
                           bias=BIAS).eval()
    return ns.DepthNet(C_IN, C_MID_MONO, C_CTX, n_depth(), use_dcn=True, use_aspp=True).eval()


def main():
    ns = ref_loader.load_reference()
    x, mlp, prev, curr = inputs()
    out = {}
    mono = build(ns, False)
    sd_m = DO.seeded_state_dict(mono, 41)
    mono.load_state_dict(sd_m)
    st = build(ns, True)
    sd_s = DO.seeded_state_dict(st, 42)
    st.load_state_dict(sd_s)
    metas = stereo_metas(prev, curr)
    with torch.no_grad():
        out['mono'] = mono(x, mlp).numpy()
        out['stereo'] = st(x, mlp, metas).numpy()
        out['cost_volume'] = st.calculate_cost_volumn(metas).numpy()
        H, W = INPUT
        out['grid'] = st.gen_grid(metas, B, NCAM, n_depth(), H // 4, W // 4, H, W).numpy()
        first = dict(metas, cv_feat_list=[None, curr])
        out['stereo_first'] = st(x, mlp, first).numpy()
    sha_in = hashlib.sha256(b''.join(t.numpy().tobytes() for t in (x, mlp, prev, curr))).hexdigest()
    np.savez_compressed(os.path.join(OUT, 'depthnet.npz'), sha_mono=sha_sd(sd_m), sha_stereo=sha_sd(sd_s),
                        input_sha=sha_in, **out)
    print('wrote depthnet.npz', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
