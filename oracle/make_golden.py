"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the UNMODIFIED
reference code (imported from /root/reference through oracle/ref_loader.py) on seeded
synthetic inputs.  Run in the build container only:

    python -m oracle.make_golden

Inputs are NOT stored: tests regenerate them from the same seeds with
oracle.mghs_oracle.synthetic_inputs and verify `input_sha` first.
"""
import hashlib
import os
import warnings

import numpy as np
import torch

from . import mghs_oracle as O
from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')

MINI = O.MINI


def sha(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def input_sha(inputs, depth, feat, height):
    ts = [depth, feat] + ([height] if height is not None else [])
    return sha(*ts)


def ref_geometry_object(ns, cfg):
    """Geometry-only reference MGHS (SURVEY.md appendix B recipe)."""
    m = ns.MGHS.__new__(ns.MGHS)
    torch.nn.Module.__init__(m)
    m.sid = False
    m.collapse_z = True
    m.out_channels = cfg['C']
    m.create_grid_infos(**cfg['bev_grid'])
    m.frustum = m.create_frustum(cfg['depth'], cfg['input_size'], cfg['downsample'])
    m.height_range = cfg['height_range']
    m.mask_range = cfg['mask_range']
    if cfg['mask_grids']:
        m.mask_1_grid, m.mask_2_grid, m.mask_3_grid = [dict(g, depth=cfg['depth']) for g in cfg['mask_grids']]
    return m


def ref_ranks(m, coor, grid):
    """Per-point voxel rank (-1 = not kept) from the reference's own prepare."""
    m.create_grid_infos(**grid)
    rb, rd, rf, st, ln = m.voxel_pooling_prepare_v2(coor)
    ranks = torch.full((coor.numel() // 3,), -1, dtype=torch.int32)
    ranks[rd.long()] = rb
    return ranks, st.numel(), rb.numel()


def gen_case(ns, name, cfg, B, seed, full_outputs, flip_bda=False, n_sample=2048):
    inputs, depth, feat, height = O.synthetic_inputs(cfg, B, seed=seed, flip_bda=flip_bda)
    m = ref_geometry_object(ns, cfg)
    D = m.D
    m.D = D
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter('ignore')
        coor = m.get_ego_coor(*inputs[1:7])
        grids = [cfg['bev_grid']] + list(cfg['mask_grids'])
        data = dict(input_sha=np.array(input_sha(inputs, depth, feat, height)),
                    coor_sha=np.array(sha(coor)))
        # the camera rig and the per-camera 3x3s the reference derives from it with torch
        # (host libm / LAPACK results are stored, not regenerated, so they are the same everywhere)
        for k, v in zip(('sensor2ego', 'ego2global', 'cam2imgs', 'post_rots', 'post_trans', 'bda'), inputs[1:]):
            data['rig_' + k] = v.numpy()
        for k, v in zip(('inv_post_rot', 'post_tran', 'combine', 'trans', 'bda'),
                        O.camera_matrices(inputs[1], inputs[3], inputs[4], inputs[5], inputs[6])):
            data['mat_' + k] = v.numpy()
        for p, g in enumerate(grids):
            r, n_int, n_kept = ref_ranks(m, coor, g)
            data['n_intervals_%d' % p] = np.array(n_int)
            data['n_kept_%d' % p] = np.array(n_kept)
            if full_outputs:
                data['ranks_%d' % p] = r.numpy()
            else:
                data['ranks_sha_%d' % p] = np.array(sha(r))
        if cfg['mask_grids']:
            outs = m.view_transform(list(inputs), depth, feat, height)
            outs = (outs[0], outs[3], outs[4], outs[5])
            mid, _ = O.height_masks(height, cfg['height_range'], cfg['mask_range'])
            k = torch.argmax(height, dim=1)
            hv = m.height_feature_to_height_map(height, m.height_range)
            m1, m2, m3 = m.create_mask_3(hv, *cfg['mask_range'])
            ref_mid = m1.to(torch.int8) + 2 * m2.to(torch.int8) + 3 * m3.to(torch.int8)
            data['mask_id_sha'] = np.array(sha(ref_mid))
            if full_outputs:
                data['mask_id'] = ref_mid.numpy()
        else:
            m.create_grid_infos(**cfg['bev_grid'])
            N = cfg['ncams']
            fH, fW = depth.shape[-2:]
            o = m.voxel_pooling_v2(coor, depth.view(B, N, D, fH, fW), feat.view(B, N, cfg['C'], fH, fW))
            outs = (o,)
        g = torch.Generator().manual_seed(1234)
        for p, o in enumerate(outs):
            data['out_sum_%d' % p] = np.array(o.double().sum().item())
            data['out_abs_sum_%d' % p] = np.array(o.double().abs().sum().item())
            data['out_nnz_%d' % p] = np.array(int((o != 0).sum()))
            if full_outputs:
                data['out_%d' % p] = o.numpy()
            else:
                flat = o.flatten()
                nz = torch.nonzero(flat).flatten()
                pick = nz[torch.randperm(nz.numel(), generator=g)[:n_sample]]
                zero = torch.nonzero(flat == 0).flatten()
                pick = torch.cat([pick, zero[torch.randperm(zero.numel(), generator=g)[:n_sample // 4]]])
                data['sample_idx_%d' % p] = pick.numpy()
                data['sample_val_%d' % p] = flat[pick].numpy()
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **data)
    print(name, os.path.getsize(path) // 1024, 'KiB',
          {k: int(v) for k, v in data.items() if k.startswith('n_')})


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_loader.load_reference()
    gen_case(ns, 'cfg1_b1', O.CFG1, 1, seed=0, full_outputs=True)
    gen_case(ns, 'mini_mghs_b2', MINI, 2, seed=3, full_outputs=True, flip_bda=True)
    gen_case(ns, 'dhds_b1', O.DHD_S, 1, seed=0, full_outputs=False)
    gen_case(ns, 'dhds_b2_flip', O.DHD_S, 2, seed=7, full_outputs=False, flip_bda=True)


if __name__ == '__main__':
    main()
