"""Shared helpers for the parity tests (test infrastructure, imports the oracle)."""
import hashlib
import os

import numpy as np
import torch

from oracle import mghs_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = {
    'cfg1_b1': (O.CFG1, 1, 0, False),
    'mini_mghs_b2': (O.MINI, 2, 3, True),
    'dhds_b1': (O.DHD_S, 1, 0, False),
    'dhds_b2_flip': (O.DHD_S, 2, 7, True),
}


def sha(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes())
    return h.hexdigest()


RIG_KEYS = ('sensor2ego', 'ego2global', 'cam2imgs', 'post_rots', 'post_trans', 'bda')
MAT_KEYS = ('inv_post_rot', 'post_tran', 'combine', 'trans', 'bda')


def load_case(name):
    """Seeded inputs (exact-arithmetic generator, checked by SHA) + the camera rig stored in
    the fixture + the reference outputs."""
    cfg, B, seed, flip = CASES[name]
    gold = np.load(os.path.join(GOLDEN, name + '.npz'))
    rig = tuple(torch.from_numpy(gold['rig_' + k]) for k in RIG_KEYS)
    inputs, depth, feat, height = O.synthetic_inputs(cfg, B, seed=seed, flip_bda=flip, rig=rig)
    ts = [depth, feat] + ([height] if height is not None else [])
    assert sha(*ts) == str(gold['input_sha']), 'seeded synthetic inputs differ from the ones the fixture was made with'
    return cfg, B, inputs, depth, feat, height, gold


def fixture_mats(gold):
    """Per-camera 3x3s the reference derived from the rig (computed where the fixture was made)."""
    return [torch.from_numpy(gold['mat_' + k]) for k in MAT_KEYS]


def grids_of(cfg):
    return [cfg['bev_grid']] + list(cfg['mask_grids'])


def oracle_ranks(coor, grid):
    """Per-point voxel rank (-1 = not kept), oracle quantiser."""
    lower, interval, size = O.grid_infos(grid['x'], grid['y'], grid['z'])
    idx, kept = O.quantise(coor, lower, interval, size)
    B = coor.shape[0]
    n = coor.numel() // 3
    idx = idx.view(n, 3)
    batch = torch.arange(B).view(B, 1).expand(B, n // B).reshape(n)
    dz, dy, dx = int(size[2]), int(size[1]), int(size[0])
    r = batch * (dz * dy * dx) + idx[:, 2] * (dy * dx) + idx[:, 1] * dx + idx[:, 0]
    r = torch.where(kept.view(n), r, torch.full_like(r, -1))
    return r.int()
