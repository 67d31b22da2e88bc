"""TEST INFRASTRUCTURE -- CPU (torch) restatement of the reference's dense hot-path modules,
written as pure functions of a state_dict so the same seeded weights can be pushed through the
reference classes (in the build container), this oracle (anywhere) and the CUDA engines.

  heightnet_forward   models/model_utils/depthnet.py:605-652 (trunk 418-487, ASPP 88-108,
                      Mlp 136-147, SELayer 158-169)
  depth_head_forward  models/necks/lss_heightmap.py:482-485
  depthnet_forward    models/model_utils/depthnet.py:362-415 (ctor 172-243), stereo_sampling_grid 245-308,
                      stereo_cost_volume 310-361
  sfa_forward         models/necks/mix.py:37-59, 87-90
  predictor_forward   models/dense_heads/occ_head.py:84-100

Third-party layers not under /root/reference, restated from their published
definitions -- PARITY UNPINNED at these three boundaries (no reference test vector exists and
the packages cannot be installed here):
  mmcv-full 1.5.3  DCN == DeformConv2dPack (depthnet.py:225-236, 466-477): 3x3 offset conv ->
                   deform_groups*2*kh*kw channels in (dy, dx) order per tap -> deformable conv, no bias
                   (restated with torchvision.ops.deform_conv2d, same offset layout)
  mmdet 2.25.1     BasicBlock: conv3x3-BN-ReLU-conv3x3-BN + identity -> ReLU
  mmcv-full 1.5.3  ConvModule: conv -> [norm] -> ReLU by default (occ_head.py:52-60)
Everything that IS under /root/reference is pinned against the real classes by
tests/test_dense_oracle.py.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class DeformConv2dPack(nn.Module):
    """Stand-in handed to the reference's build_conv_layer(type='DCN') by oracle/ref_loader.py."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 dilation=1, groups=1, deform_groups=1, bias=False, **kw):
        super().__init__()
        k = kernel_size
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.groups, self.deform_groups = groups, deform_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, k, k))
        nn.init.kaiming_uniform_(self.weight, nonlinearity='relu')
        self.conv_offset = nn.Conv2d(in_channels, deform_groups * 2 * k * k, k, stride=stride,
                                     padding=padding, dilation=dilation, bias=True)
        nn.init.zeros_(self.conv_offset.weight)
        nn.init.zeros_(self.conv_offset.bias)

    def forward(self, x):
        from torchvision.ops import deform_conv2d
        off = self.conv_offset(x)
        return deform_conv2d(x, off, self.weight, None, stride=self.stride, padding=self.padding,
                             dilation=self.dilation)


# ------------------------------------------------------------------ seeded, exactly representable weights
def seeded_state_dict(module, seed):
    """Fill every parameter / buffer with values that are integers scaled by powers of two
    (bit-identical on every host; no libm in the generator), at magnitudes that keep activations
    O(1): weights ~ U(-a, a), a = 2^round(log2(sqrt(3 / fan_in))); BN gamma in [0.75, 1.25],
    beta / running_mean in [-0.25, 0.25], running_var in [0.5, 1.5]."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    import math
    for name, t in module.state_dict().items():
        if name.endswith('num_batches_tracked'):
            sd[name] = torch.zeros_like(t)
            continue
        r = torch.randint(-2048, 2048, t.shape, generator=g).double() / 2048.0      # U(-1, 1), 12 bits
        if name.endswith('running_var'):
            v = 1.0 + 0.5 * r
        elif name.endswith('running_mean') or (name.endswith('.bias') and t.dim() == 1):
            v = 0.25 * r
        elif t.dim() == 1:                                                           # BN weight
            v = 1.0 + 0.25 * r
        else:
            fan_in = t[0].numel()
            a = 2.0 ** round(math.log2(math.sqrt(3.0 / fan_in)))
            if 'conv_offset' in name:
                a *= 0.5                                                             # offsets of ~1 px
            v = a * r
        sd[name] = v.to(t.dtype)
    return sd


def seeded_tensor(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(-2048, 2048, shape, generator=g).float() * (scale / 1024.0)


# ------------------------------------------------------------------ functional restatements
BN_TRAIN = False     # True: BatchNorm2d in training mode (batch statistics), for the training-path parity tests


def _bn(sd, p, x, eps=1e-5):
    if BN_TRAIN and x.dim() == 4 and x.shape[2] * x.shape[3] > 1:
        return F.batch_norm(x, None, None, sd[p + '.weight'], sd[p + '.bias'], True, 0.0, eps)
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'], sd[p + '.weight'],
                        sd[p + '.bias'], False, 0.0, eps)


def _basic_block(sd, p, x):
    out = F.relu(_bn(sd, p + '.bn1', F.conv2d(x, sd[p + '.conv1.weight'], padding=1)))
    out = _bn(sd, p + '.bn2', F.conv2d(out, sd[p + '.conv2.weight'], padding=1))
    return F.relu(out + x)


def _aspp(sd, p, x):
    outs = [F.relu(_bn(sd, p + '.aspp1.bn', F.conv2d(x, sd[p + '.aspp1.atrous_conv.weight'])))]
    for k, d in (('aspp2', 6), ('aspp3', 12), ('aspp4', 18)):
        outs.append(F.relu(_bn(sd, '%s.%s.bn' % (p, k),
                               F.conv2d(x, sd['%s.%s.atrous_conv.weight' % (p, k)], padding=d, dilation=d))))
    g = F.adaptive_avg_pool2d(x, 1)
    g = F.relu(_bn(sd, p + '.global_avg_pool.2', F.conv2d(g, sd[p + '.global_avg_pool.1.weight'])))
    outs.append(F.interpolate(g, size=x.shape[2:], mode='bilinear', align_corners=True))
    y = F.conv2d(torch.cat(outs, 1), sd[p + '.conv1.weight'])
    return F.relu(_bn(sd, p + '.bn1', y))          # Dropout(0.5) is the identity in eval


def _dcn(sd, p, x, groups=4):
    from torchvision.ops import deform_conv2d
    off = F.conv2d(x, sd[p + '.conv_offset.weight'], sd[p + '.conv_offset.bias'], padding=1)
    return deform_conv2d(x, off, sd[p + '.weight'], None, stride=1, padding=1, dilation=1)


def heightnet_forward(sd, x, mlp_input, prefix=''):
    """Eval-mode HeightNet (non-stereo).  sd: state_dict with the reference's names."""
    p = prefix
    m = F.batch_norm(mlp_input.reshape(-1, mlp_input.shape[-1]), sd[p + 'bn.running_mean'],
                     sd[p + 'bn.running_var'], sd[p + 'bn.weight'], sd[p + 'bn.bias'], False, 0.0, 1e-5)
    x = F.conv2d(x, sd[p + 'reduce_conv.0.weight'], sd[p + 'reduce_conv.0.bias'], padding=1)
    x = F.relu(_bn(sd, p + 'reduce_conv.1', x))
    se = F.linear(F.relu(F.linear(m, sd[p + 'depth_mlp.fc1.weight'], sd[p + 'depth_mlp.fc1.bias'])),
                  sd[p + 'depth_mlp.fc2.weight'], sd[p + 'depth_mlp.fc2.bias'])[..., None, None]
    se = F.relu(F.conv2d(se, sd[p + 'depth_se.conv_reduce.weight'], sd[p + 'depth_se.conv_reduce.bias']))
    se = F.conv2d(se, sd[p + 'depth_se.conv_expand.weight'], sd[p + 'depth_se.conv_expand.bias'])
    x = x * torch.sigmoid(se)
    i = 0
    while (p + 'depth_conv.%d.bn2.weight' % i) in sd:      # BasicBlocks (ASPP has conv1 + bn1 but no bn2)
        x = _basic_block(sd, p + 'depth_conv.%d' % i, x)
        i += 1
    if (p + 'depth_conv.%d.aspp1.atrous_conv.weight' % i) in sd:
        x = _aspp(sd, p + 'depth_conv.%d' % i, x)
        i += 1
    if (p + 'depth_conv.%d.conv_offset.weight' % i) in sd:
        x = _dcn(sd, p + 'depth_conv.%d' % i, x)
        i += 1
    return F.conv2d(x, sd[p + 'depth_conv.%d.weight' % i], sd[p + 'depth_conv.%d.bias' % i])


# ------------------------------------------------------------------ camera-aware DepthNet + plane-sweep cost volume
def stereo_sampling_grid(frustum, k2s_sensor, intrins, post_rots, post_trans, hi, wi):
    """models/model_utils/depthnet.py:245-308 (gen_grid): every point of the 1/4-resolution frustum template
    (u, v, d) of the CURRENT image -> un-augment -> current camera -> previous camera (k2s_sensor) -> previous
    pixel -> re-augment -> normalised [-1, 1] sampling coordinate; points behind the previous camera (z < 1e-3)
    are sent to -2 (outside).  Returns (B*N, D*H, W, 2) fp32, the layout F.grid_sample takes."""
    B, N = post_trans.shape[:2]
    D, H, W, _ = frustum.shape
    bc = lambda t, *tail: t.reshape(B, N, 1, 1, 1, *tail)
    p = (frustum - bc(post_trans, 3)).unsqueeze(-1)
    p = bc(torch.inverse(post_rots), 3, 3).matmul(p)
    p = torch.cat([p[..., :2, :] * p[..., 2:3, :], p[..., 2:3, :]], dim=5)
    to_prev = k2s_sensor[:, :, :3, :3].contiguous().matmul(torch.inverse(intrins))
    p = bc(to_prev, 3, 3).matmul(p) + bc(k2s_sensor[:, :, :3, 3].contiguous(), 3, 1)
    behind = p[..., 2, 0] < 1e-3
    p = bc(intrins, 3, 3).matmul(p)
    uv = (p[..., :2, :] / p[..., 2:3, :])
    uv = bc(post_rots[..., :2, :2], 2, 2).matmul(uv).squeeze(-1) + bc(post_trans[..., :2], 2)
    gx = uv[..., 0] / (wi - 1.0) * 2.0 - 1.0
    gy = uv[..., 1] / (hi - 1.0) * 2.0 - 1.0
    gx[behind] = -2
    gy[behind] = -2
    return torch.stack([gx, gy], dim=-1).view(B * N, D * H, W, 2)


def stereo_cost_volume(prev, curr, grid, D, bias=0.0, group_size=4):
    """depthnet.py:310-361: warp the previous frame's stereo feature to every depth hypothesis of the current
    pixel (bilinear, zeros outside, align_corners=True), L1 distance to the current feature summed over the
    channels (the reference accumulates it 4 channels at a time), `bias` added where the warped channel
    C - group_size is exactly 0 (its test for "sample fell outside"), softmax over depth of the negated cost."""
    BN, C, H, W = curr.shape
    cost = torch.zeros(BN, D, H, W, dtype=curr.dtype, device=curr.device)
    last = None
    for c0 in range(0, C, group_size):
        last = F.grid_sample(prev[:, c0:c0 + group_size], grid, align_corners=True, padding_mode='zeros')
        last = last.view(BN, -1, D, H, W)
        cost += (curr[:, c0:c0 + group_size, None] - last).abs().sum(dim=1)
    if bias != 0:
        cost = torch.where(last[:, 0] == 0, cost + bias, cost)
    return (-cost).softmax(dim=1)


def _camera_gate(sd, p, m, mlp, se, x):
    """Mlp (depthnet.py:136-147, fc1-ReLU-fc2; Dropout(0) twice) -> SELayer (158-169) gate on x."""
    g = F.linear(F.relu(F.linear(m, sd[p + mlp + '.fc1.weight'], sd[p + mlp + '.fc1.bias'])),
                 sd[p + mlp + '.fc2.weight'], sd[p + mlp + '.fc2.bias'])[..., None, None]
    g = F.relu(F.conv2d(g, sd[p + se + '.conv_reduce.weight'], sd[p + se + '.conv_reduce.bias']))
    g = F.conv2d(g, sd[p + se + '.conv_expand.weight'], sd[p + se + '.conv_expand.bias'])
    return x * torch.sigmoid(g)


def depthnet_forward(sd, x, mlp_input, cost_volume=None, prefix=''):
    """Eval-mode DepthNet of MGHS_Depth / MGHS_Stereo (depthnet.py:362-415).  cost_volume: None (stereo=False) or
    the (B*N, D, 4fH, 4fW) matching probabilities (zeros when there is no previous frame, 389-396); it goes through
    cost_volumn_net (two stride-2 conv3x3 + BN, 207-213), is concatenated to the gated depth feature, and the first
    BasicBlock then carries the plain 1x1 `downsample` convolution on its identity path (205-206, 217-218).
    Returns (B*N, D + C_context, fH, fW): raw depth logits then the context feature."""
    p = prefix
    m = F.batch_norm(mlp_input.reshape(-1, mlp_input.shape[-1]), sd[p + 'bn.running_mean'],
                     sd[p + 'bn.running_var'], sd[p + 'bn.weight'], sd[p + 'bn.bias'], False, 0.0, 1e-5)
    x = F.conv2d(x, sd[p + 'reduce_conv.0.weight'], sd[p + 'reduce_conv.0.bias'], padding=1)
    x = F.relu(_bn(sd, p + 'reduce_conv.1', x))
    context = _camera_gate(sd, p, m, 'context_mlp', 'context_se', x)
    context = F.conv2d(context, sd[p + 'context_conv.weight'], sd[p + 'context_conv.bias'])
    y = _camera_gate(sd, p, m, 'depth_mlp', 'depth_se', x)
    if cost_volume is not None:
        cv = cost_volume
        for k in (0, 2):
            q = p + 'cost_volumn_net.%d' % k
            cv = _bn(sd, p + 'cost_volumn_net.%d' % (k + 1),
                     F.conv2d(cv, sd[q + '.weight'], sd[q + '.bias'], stride=2, padding=1))
        y = torch.cat([y, cv], dim=1)
    i = 0
    while (p + 'depth_conv.%d.bn2.weight' % i) in sd:
        q = p + 'depth_conv.%d' % i
        if (q + '.downsample.weight') in sd:
            out = F.relu(_bn(sd, q + '.bn1', F.conv2d(y, sd[q + '.conv1.weight'], padding=1)))
            out = _bn(sd, q + '.bn2', F.conv2d(out, sd[q + '.conv2.weight'], padding=1))
            y = F.relu(out + F.conv2d(y, sd[q + '.downsample.weight'], sd[q + '.downsample.bias']))
        else:
            y = _basic_block(sd, q, y)
        i += 1
    if (p + 'depth_conv.%d.aspp1.atrous_conv.weight' % i) in sd:
        y = _aspp(sd, p + 'depth_conv.%d' % i, y)
        i += 1
    if (p + 'depth_conv.%d.conv_offset.weight' % i) in sd:
        y = _dcn(sd, p + 'depth_conv.%d' % i, y)
        i += 1
    y = F.conv2d(y, sd[p + 'depth_conv.%d.weight' % i], sd[p + 'depth_conv.%d.bias' % i])
    return torch.cat([y, context], dim=1)


def depth_head_forward(sd, x, n_depth, prefix='depth_net.'):
    y = F.conv2d(x, sd[prefix + 'weight'], sd[prefix + 'bias'])
    return y[:, :n_depth].softmax(dim=1), y[:, n_depth:]


def sfa_forward(sd, x, prefix=''):
    p = prefix
    C = x.shape[1] // 2
    bev, vox = x[:, :C], x[:, C:]
    s = x.mean(-1).mean(-1)
    a1 = torch.sigmoid(F.linear(F.relu(F.linear(s, sd[p + 'mysk_7.fc.0.weight'], sd[p + 'mysk_7.fc.0.bias'])),
                                sd[p + 'mysk_7.fc.2.weight'], sd[p + 'mysk_7.fc.2.bias']))[..., None, None]
    b1, v1 = a1 * bev, (1 - a1) * vox
    q = p + 'mysk_7.spacial_leanring'
    t = F.relu(_bn(sd, q + '.1', F.conv2d(b1 + v1, sd[q + '.0.weight'], sd[q + '.0.bias'])))
    a2 = torch.sigmoid(_bn(sd, q + '.4', F.conv2d(t, sd[q + '.3.weight'], sd[q + '.3.bias'])))
    fuse = a2 * b1 + (1 - a2) * v1
    r = F.relu(_bn(sd, p + 'mix_residual.1', F.conv2d(fuse, sd[p + 'mix_residual.0.weight'], padding=1)))
    r = _bn(sd, p + 'mix_residual.4', F.conv2d(r, sd[p + 'mix_residual.3.weight'], padding=1))
    sc = _bn(sd, p + 'mix_shortcut.1', F.conv2d(x, sd[p + 'mix_shortcut.0.weight']))
    return F.relu(r + sc)


def predictor_forward(sd, x, Dz=16, num_classes=18, prefix=''):
    p = prefix
    y = F.relu(F.conv2d(x, sd[p + 'final_conv.conv.weight'], sd[p + 'final_conv.conv.bias'], padding=1))
    y = y.permute(0, 3, 2, 1)
    y = F.linear(F.softplus(F.linear(y, sd[p + 'predicter.0.weight'], sd[p + 'predicter.0.bias'])),
                 sd[p + 'predicter.2.weight'], sd[p + 'predicter.2.bias'])
    return y.view(y.shape[0], y.shape[1], y.shape[2], Dz, num_classes)


# ------------------------------------------------------------------ BEV / voxel encoders
def _double_conv(sd, p, x):
    """backbones/unet.py:45-61"""
    x = F.relu(_bn(sd, p + '.1', F.conv2d(x, sd[p + '.0.weight'], padding=1)))
    return F.relu(_bn(sd, p + '.4', F.conv2d(x, sd[p + '.3.weight'], padding=1)))


def unet_forward(sd, x, prefix=''):
    """backbones/unet.py:25-40 (bilinear=False): inc, 4x (MaxPool2d(2) + DoubleConv), 4x (ConvTranspose2d(2, 2),
    pad to the skip's size, cat([skip, up]), DoubleConv), 1x1 outc."""
    p = prefix
    xs = [_double_conv(sd, p + 'inc.double_conv', x)]
    for k in range(1, 5):
        xs.append(_double_conv(sd, p + 'down%d.maxpool_conv.1.double_conv' % k, F.max_pool2d(xs[-1], 2)))
    y = xs[4]
    for k in range(1, 5):
        skip = xs[4 - k]
        y = F.conv_transpose2d(y, sd[p + 'up%d.up.weight' % k], sd[p + 'up%d.up.bias' % k], stride=2)
        dy, dx = skip.shape[2] - y.shape[2], skip.shape[3] - y.shape[3]
        y = F.pad(y, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
        y = _double_conv(sd, p + 'up%d.conv.double_conv' % k, torch.cat([skip, y], dim=1))
    return F.conv2d(y, sd[p + 'outc.conv.weight'], sd[p + 'outc.conv.bias'])


def custom_resnet_forward(sd, x, prefix='', strides=(2, 2, 2), num_layer=(2, 2, 2)):
    """backbones/resnet.py:10-80 (block_type='Basic'): returns the list of stage outputs."""
    feats = []
    for i, (st, nl) in enumerate(zip(strides, num_layer)):
        for j in range(nl):
            p = '%slayers.%d.%d' % (prefix, i, j)
            s = st if j == 0 else 1
            out = F.relu(_bn(sd, p + '.bn1', F.conv2d(x, sd[p + '.conv1.weight'], stride=s, padding=1)))
            out = _bn(sd, p + '.bn2', F.conv2d(out, sd[p + '.conv2.weight'], padding=1))
            idn = x
            if (p + '.downsample.weight') in sd:
                idn = F.conv2d(x, sd[p + '.downsample.weight'], sd[p + '.downsample.bias'], stride=s, padding=1)
            x = F.relu(out + idn)
        feats.append(x)
    return feats


def fpn_lss_forward(sd, feats, prefix='', index=(0, 2), scale=4, scale2=2):
    """necks/lss_fpn.py:62-74 (lateral=None, extra_upsample=2)."""
    p = prefix
    x2, x1 = feats[index[0]], feats[index[1]]
    x1 = F.interpolate(x1, scale_factor=scale, mode='bilinear', align_corners=True)
    x = torch.cat([x2, x1], dim=1)
    x = F.relu(_bn(sd, p + 'conv.1', F.conv2d(x, sd[p + 'conv.0.weight'], padding=1)))
    x = F.relu(_bn(sd, p + 'conv.4', F.conv2d(x, sd[p + 'conv.3.weight'], padding=1)))
    x = F.interpolate(x, scale_factor=scale2, mode='bilinear', align_corners=True)
    x = F.relu(_bn(sd, p + 'up2.2', F.conv2d(x, sd[p + 'up2.1.weight'], padding=1)))
    return F.conv2d(x, sd[p + 'up2.4.weight'], sd[p + 'up2.4.bias'])


# ------------------------------------------------------------------ image backbone + neck (the 8(f)-4 widening)
def image_resnet_forward(sd, img, depth=50, out_indices=(2, 3), prefix='', num_stages=4):
    """mmdet 2.25.1 `ResNet(depth, style='pytorch')` in eval mode = the torchvision ResNet the reference's configs load
    (`pretrained='torchvision://resnet50'`, projects/configs/DHD/DHD-S.py:44-55): conv1 7x7/2 + bn + relu, maxpool 3x3/2,
    Bottleneck stages (1x1, 3x3 with the stage stride, 1x1; a 1x1 stride-s conv + bn on the identity of every stage's
    first block).  mmdet is not part of the reference tree: pinned against torchvision.models.resnet50 (tests).  With
    BN_TRAIN the BatchNorms use batch statistics (mmdet's train() with norm_eval=False, DHD-S.py:50-51)."""
    blocks = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}[depth][:num_stages]
    x = F.relu(_bn(sd, prefix + 'bn1', F.conv2d(img, sd[prefix + 'conv1.weight'], stride=2, padding=3)))
    x = F.max_pool2d(x, 3, stride=2, padding=1)
    outs = []
    for li, nb in enumerate(blocks):
        for bi in range(nb):
            p = '%slayer%d.%d.' % (prefix, li + 1, bi)
            stride = 2 if (li > 0 and bi == 0) else 1
            idn = x
            y = F.relu(_bn(sd, p + 'bn1', F.conv2d(x, sd[p + 'conv1.weight'])))
            y = F.relu(_bn(sd, p + 'bn2', F.conv2d(y, sd[p + 'conv2.weight'], stride=stride, padding=1)))
            y = _bn(sd, p + 'bn3', F.conv2d(y, sd[p + 'conv3.weight']))
            if p + 'downsample.0.weight' in sd:
                idn = _bn(sd, p + 'downsample.1', F.conv2d(x, sd[p + 'downsample.0.weight'], stride=stride))
            x = F.relu(y + idn)
        if li in out_indices:
            outs.append(x)
    return outs


def custom_fpn_forward(sd, inputs, out_ids=(0,), start_level=0, prefix=''):
    """projects/mmdet3d_plugin/models/necks/fpn.py:153-176 for norm_cfg=None / act_cfg=None / no extra convs (the DHD
    configs): lateral 1x1 convs, laterals[i-1] += nearest-interpolated laterals[i], 3x3 conv on the out_ids levels."""
    n = len(inputs) - start_level
    lats = [F.conv2d(inputs[i + start_level], sd['%slateral_convs.%d.conv.weight' % (prefix, i)],
                     sd['%slateral_convs.%d.conv.bias' % (prefix, i)]) for i in range(n)]
    for i in range(n - 1, 0, -1):
        lats[i - 1] = lats[i - 1] + F.interpolate(lats[i], size=lats[i - 1].shape[2:], mode='nearest')
    return [F.conv2d(lats[i], sd['%sfpn_convs.%d.conv.weight' % (prefix, j)], sd['%sfpn_convs.%d.conv.bias' % (prefix, j)],
                     padding=1) for j, i in enumerate(out_ids)]
