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


# ---------------------------------------------------------------------------------------------- teacher-forced references
class StoredRound(torch.autograd.Function):
    """An activation the CUDA path keeps in bf16 -- and whose gradient it keeps in bf16 too: rounded in both directions."""

    @staticmethod
    def forward(ctx, t):
        return t.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class Forced(torch.autograd.Function):
    """The reference continues from the value the CUDA path stored (v); straight-through gradient, rounded to bf16."""

    @staticmethod
    def forward(ctx, t, v):
        return v.clone()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float(), None


class ForcedF:
    """torch.nn.functional stand-in for oracle.dense_oracle: the outputs of the functions named in `forced` are replaced,
    call by call, by the CUDA path's stored activations (lists consumed in call order); the functions in `rounded` get
    the bf16 storage rounding only.  `drift[name]` collects the relative distance between each forced value and the
    reference's own value: a per-layer forward check that does not accumulate.  With identical stored activations both
    backward passes see the same ReLU masks / pooling winners / BatchNorm statistics, so a gradient comparison
    measures the backward kernels instead of the amplified forward rounding."""

    def __init__(self, forced, rounded=()):
        self.forced, self.rounded = {k: list(v) for k, v in forced.items()}, set(rounded)
        self.drift = {k: [] for k in forced}

    def __getattr__(self, name):
        import torch.nn.functional as F
        fn = getattr(F, name)
        if name in self.forced:
            def call(*a, **k):
                t = fn(*a, **k)
                assert self.forced[name], 'more %s calls than stored activations' % name
                v = self.forced[name].pop(0)
                assert v.shape == t.shape, (name, tuple(v.shape), tuple(t.shape))
                self.drift[name].append(float((t.detach() - v).norm() / v.norm().clamp_min(1e-20)))
                return Forced.apply(t, v)
            return call
        if name in self.rounded:
            return lambda *a, **k: StoredRound.apply(fn(*a, **k))
        return fn

    def exhausted(self):
        return all(not v for v in self.forced.values())
