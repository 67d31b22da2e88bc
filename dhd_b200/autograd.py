"""torch.autograd bridges of the plugin modules: `module.forward` under autograd runs the training engines of
dhd_b200.train (forward with saved activations on the tcgen05 kernels) and `.backward()` calls their hand-written
backward passes, so the reference's training code drives the B200 path unchanged:

    losses = model.forward_train(img_inputs=..., voxel_semantics=..., mask_camera=..., gt_depth=..., gt_height=...)
    sum(losses.values()).backward(); optimizer.step()

(reference: ordinary differentiable nn.Modules -- MGHS.forward lss_heightmap.py:461-490, SFA.forward mix.py:87-90,
predictor.forward occ_head.py:84-100, UNet / CustomResNet / FPN_LSS -- under mmcv's runner, tools/train.py:276.)

Contract of every bridge
  * activations cross module boundaries as fp32 torch tensors (logical NCHW, channels_last memory); inside a module
    they are bf16 NHWC (mixed precision: bf16 operands, fp32 accumulation, fp32 master weights / gradients);
  * parameter gradients are ACCUMULATED INTO `param.grad` by the backward kernels (what autograd's AccumulateGrad
    would do); the parameters are passed to the Function only so that autograd schedules the backward even when the
    input needs no gradient.  Consequence: per-parameter autograd hooks (torch DDP's reducer) do not fire -- use
    dhd_b200.shard.GradBucket (one flat all-reduce, the reference's DDP bucket) for data parallelism;
  * BatchNorm follows `module.training`: batch statistics + running-stat updates in train(), frozen statistics in
    eval() with autograd on (fine-tuning with frozen BN); Dropout (HeightNet's ASPP) is on in train() only;
  * one forward may be in flight per module (the engine owns ONE set of saved activations), which is how the
    reference's detectors call them (DHD_stereo runs its second frame under no_grad -> the inference engine).
"""
import torch

from . import dense as D
from . import train as T


def wants_grad(module, *tensors):
    """True when this call must build a backward: autograd is on and either an input requires grad or the module is
    in train() mode with trainable parameters.  A module in eval() mode fed with plain tensors runs the inference
    engine (in its own precision mode) even outside torch.no_grad()."""
    if not torch.is_grad_enabled():
        return False
    if any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        return True
    return bool(module.training) and any(p.requires_grad for p in module.parameters())


def trainer_of(module, factory, device, batch_bn=None):
    """The module's training engine, built once per (device, train/eval mode) and re-packed (`refresh`) whenever a
    parameter changed since (optimizer.step bumps the tensors' version counters).  batch_bn overrides the BatchNorm
    mode implied by module.training (mmdet's ResNet keeps its BatchNorms in eval mode under train() when norm_eval)."""
    batch = bool(module.training) if batch_bn is None else bool(batch_bn)
    key = (str(device), batch)
    versions = tuple((id(p), p._version) for p in module.parameters())
    slot = module.__dict__.get('_trainer')
    if slot is None or slot[0] != key:
        T.set_bn_mode('batch' if batch else 'frozen')
        try:
            tr = factory()
        finally:
            T.set_bn_mode('frozen')
        module.__dict__['_trainer'] = slot = [key, tr, versions]
    elif slot[2] != versions:
        with T.batched_repack():                     # the bf16 re-pack of all the module's layers in a few launches
            slot[1].refresh()
        slot[2] = versions
    return slot[1]


def _params(module):
    return [p for p in module.parameters() if p.requires_grad]


def to_act(x):
    """fp32 (N, C, H, W) tensor in either memory format -> bf16 NHWC activation."""
    if isinstance(x, D.Act):
        return x
    if not x.is_cuda:
        raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
    return D.pack_any(x.detach().float(), 1)


def from_act(a, C=None):
    """bf16 NHWC activation -> fp32 logical (N, C, H, W) tensor in channels_last memory."""
    C = a.C if C is None else C
    v = a.data.view(a.N, a.H, a.W, a.ld)[..., a.coff:a.coff + C]
    return v.float().permute(0, 3, 1, 2)


def grad_to_act(g, C_pad=None):
    """fp32 gradient (N, C, H, W) -> bf16 NHWC Act, channels zero-padded to C_pad."""
    N, C, H, W = g.shape
    Cp = C if C_pad is None else C_pad
    a = D.Act(torch.zeros(N, H, W, Cp, dtype=torch.bfloat16, device=g.device), Cp, 1) if Cp != C else \
        D.Act.empty(N, H, W, C, 1, g.device)
    a.data[..., :C].copy_(g.permute(0, 2, 3, 1))
    return a


# ------------------------------------------------------------------------------------------------ SFA
class _SFAFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, trainer, *params):
        out = trainer.forward(to_act(x))
        T.flush_bn_counters()
        ctx.trainer = trainer
        return from_act(out)

    @staticmethod
    def backward(ctx, g):
        dx = ctx.trainer.backward(grad_to_act(g), want_dx=ctx.needs_input_grad[0])
        return (from_act(dx) if dx is not None else None, None) + (None,) * (len(ctx.needs_input_grad) - 2)


def sfa_forward(module, x):
    tr = trainer_of(module, lambda: T.SFATrainer(module, x.device), x.device)
    return _SFAFn.apply(x, tr, *_params(module))


# ------------------------------------------------------------------------------------------------ predictor
class _PredictorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, trainer, *params):
        out = trainer.forward(to_act(x))                     # (B, Dx, Dy, Dz, n_cls) fp32, a fresh tensor per call
        T.flush_bn_counters()
        ctx.trainer = trainer
        return out

    @staticmethod
    def backward(ctx, g):
        tr = ctx.trainer
        x = tr.saved[0]
        N, H, W = x.N, x.H, x.W
        # gradient of the (B, Dx, Dy, Dz*n_cls) logits -> the (B, Dy, Dx, nout_pad) bf16 activation the GEMMs read
        dlog = tr._act('dlog', N, H, W, tr.nout_pad, zero=True)
        dlog.data[..., :tr.nout].copy_(g.reshape(N, W, H, tr.nout).permute(0, 2, 1, 3))
        tr.dlog = dlog
        dx = tr.backward(want_dx=ctx.needs_input_grad[0])
        return (from_act(dx) if dx is not None else None, None) + (None,) * (len(ctx.needs_input_grad) - 2)


def predictor_forward(module, x):
    tr = trainer_of(module, lambda: T.PredictorTrainer(module, x.device), x.device)
    return _PredictorFn.apply(x, tr, *_params(module))


# ------------------------------------------------------------------------------------------------ encoders
class _UNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, trainer, *params):
        out = trainer.forward(to_act(x))
        T.flush_bn_counters()
        ctx.trainer = trainer
        return from_act(out, trainer.n_classes)

    @staticmethod
    def backward(ctx, g):
        tr = ctx.trainer
        dx = tr.backward(grad_to_act(g, tr.outc.cout_pad))
        return (from_act(dx) if ctx.needs_input_grad[0] else None, None) + (None,) * (len(ctx.needs_input_grad) - 2)


def unet_forward(module, x):
    tr = trainer_of(module, lambda: T.UNetTrainer(module, x.device), x.device)
    return _UNetFn.apply(x, tr, *_params(module))


class _ResNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, trainer, *params):
        feats = trainer.forward(to_act(x))
        T.flush_bn_counters()
        ctx.trainer = trainer
        return tuple(from_act(f) for f in feats)

    @staticmethod
    def backward(ctx, *gs):
        tr = ctx.trainer
        dfeats = {si: grad_to_act(g) for si, g in zip(tr.output_ids, gs) if g is not None}
        last = len(tr.stages) - 1
        if last not in dfeats:                      # the deepest stage always seeds the backward walk
            f = tr.saved[-1][2]
            dfeats[last] = D.Act(torch.zeros_like(f.data), f.C, 1)
        dx = tr.backward(dfeats)
        return (from_act(dx) if ctx.needs_input_grad[0] else None, None) + (None,) * (len(ctx.needs_input_grad) - 2)


def resnet_forward(module, x):
    tr = trainer_of(module, lambda: T.CustomResNetTrainer(module, x.device), x.device)
    return list(_ResNetFn.apply(x, tr, *_params(module)))


class _FPNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, trainer, n_feats, *args):
        feats = [to_act(f) for f in args[:n_feats]]
        out = trainer.forward(feats)
        T.flush_bn_counters()
        ctx.trainer, ctx.n_feats = trainer, n_feats
        return from_act(out)

    @staticmethod
    def backward(ctx, g):
        tr = ctx.trainer
        d = tr.backward(grad_to_act(g))
        grads = [from_act(d[i]) if (i in d and ctx.needs_input_grad[2 + i]) else None for i in range(ctx.n_feats)]
        return (None, None) + tuple(grads) + (None,) * (len(ctx.needs_input_grad) - 2 - ctx.n_feats)


def fpn_forward(module, feats):
    dev = feats[0].device
    tr = trainer_of(module, lambda: T.FPNLSSTrainer(module, dev), dev)
    return _FPNFn.apply(tr, len(feats), *feats, *_params(module))


# ------------------------------------------------------------------------------------------------ image backbone + neck
class _ImageResNetFn(torch.autograd.Function):
    """mmdet ResNet (projects/configs/DHD/DHD-S.py:44-55) under autograd: the images carry no gradient, the parameters
    do (dhd_b200.train_backbone.ImageResNetTrainer)."""

    @staticmethod
    def forward(ctx, img, trainer, *params):
        if not img.is_cuda:
            raise RuntimeError('dhd_b200: expected CUDA tensors (the hot path has no CPU fallback)')
        feats = trainer.forward(img.detach())
        T.flush_bn_counters()
        ctx.trainer = trainer
        return tuple(from_act(f) for f in feats)

    @staticmethod
    def backward(ctx, *gs):
        tr = ctx.trainer
        tr.backward({li: grad_to_act(g) for li, g in zip(tr.out_indices, gs) if g is not None})
        return (None,) * len(ctx.needs_input_grad)


def image_resnet_forward(module, img):
    from .train_backbone import ImageResNetTrainer
    if module.frozen_stages >= 0:
        raise NotImplementedError('ResNet(frozen_stages >= 0) under autograd: the DHD configs train every stage '
                                  '(frozen_stages=-1); freeze the whole backbone instead')
    dev = img.device
    tr = trainer_of(module, lambda: ImageResNetTrainer(module, dev), dev,
                    batch_bn=module.training and not module.norm_eval)
    return _ImageResNetFn.apply(img, tr, *_params(module))


class _CustomFPNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, trainer, n_feats, *args):
        outs = trainer.forward([to_act(f) for f in args[:n_feats]])
        T.flush_bn_counters()
        ctx.trainer, ctx.n_feats = trainer, n_feats
        return tuple(from_act(o) for o in outs)

    @staticmethod
    def backward(ctx, *gs):
        tr = ctx.trainer
        d = tr.backward([None if g is None else grad_to_act(g) for g in gs])
        grads = [from_act(d[i]) if (i in d and ctx.needs_input_grad[2 + i]) else None for i in range(ctx.n_feats)]
        return (None, None) + tuple(grads) + (None,) * (len(ctx.needs_input_grad) - 2 - ctx.n_feats)


def custom_fpn_forward(module, feats):
    from .train_backbone import CustomFPNTrainer
    dev = feats[0].device
    tr = trainer_of(module, lambda: CustomFPNTrainer(module, dev), dev)
    return list(_CustomFPNFn.apply(tr, len(feats), *feats, *_params(module)))


# ------------------------------------------------------------------------------------------------ MGHS (DHD-S)
class _MGHSFn(torch.autograd.Function):
    """MGHS.forward (lss_heightmap.py:461-490) with a backward: pool backward -> depth_net backward, and the height
    distribution's gradient (get_height_loss, LH:595-622) -> HeightNet backward.  The height -> mask argmax is not
    differentiable (as in the reference): the masks carry no gradient."""

    @staticmethod
    def forward(ctx, x, vt, t_depth, t_height, plan, pixmask_fn, mlp_input, layout, *params):
        B, N, C, H, W = x.shape
        xa = D.pack_input(x.detach().reshape(B * N, C, H, W).float(), 1)
        depth, feat = t_depth.forward(xa)
        height = t_height.forward(xa, mlp_input)
        T.flush_bn_counters()
        pixmask = pixmask_fn(height)
        outs = plan.alloc_outputs(layout, x.device)
        plan.raw_forward(depth, feat, pixmask, outs, layout)
        ctx.state = (vt, t_depth, t_height, plan, pixmask, depth, feat, layout, plan.workspace, (B, N, C, H, W))
        if layout == 'nhwc':
            outs = [o.permute(0, 3, 1, 2) for o in outs]
        return (outs[0], depth.clone(), height.clone()) + tuple(outs[1:])

    @staticmethod
    def backward(ctx, g_bev, g_depth, g_height, *g_masked):
        vt, t_depth, t_height, plan, pixmask, depth, feat, layout, ws, (B, N, C, H, W) = ctx.state
        gouts = [g_bev] + list(g_masked)
        if layout == 'nhwc':
            gouts = [None if g is None else g.permute(0, 2, 3, 1) for g in gouts]
        dx = None
        have_pool = any(g is not None for g in gouts)
        if have_pool or g_depth is not None:
            if have_pool:
                dgrad, fgrad = plan._backward(depth, feat.view(B, N, H, W, -1), pixmask, gouts, layout, ws)
                dgrad = dgrad.view(B * N, -1, H, W)
            else:
                dgrad, fgrad = torch.zeros_like(depth), torch.zeros_like(feat)
            if g_depth is not None:                  # a loss on the depth distribution itself joins the pool's gradient
                dgrad = dgrad + g_depth
            dx = t_depth.backward(dgrad, fgrad.view(B * N, H, W, -1), want_dx=ctx.needs_input_grad[0])
        if g_height is not None:
            # gradient at the softmax output -> at the logits: p * (g - sum_k p_k g_k), then the trainer's bf16 buffer
            p = t_height.saved['height']
            dlog = p * (g_height - (p * g_height).sum(dim=1, keepdim=True))
            dz = t_height._act('dz', B * N, H, W, t_height.head.cout_pad)
            dz.data.zero_()
            dz.data[..., :dlog.shape[1]].copy_(dlog.permute(0, 2, 3, 1))
            t_height.dz = dz
            dxh = t_height.backward(want_dx=ctx.needs_input_grad[0])
            if dxh is not None:
                dx = dxh if dx is None else D.Act(dx.data + dxh.data, dx.C, 1)
        gx = from_act(dx, C).reshape(B, N, C, H, W) if (dx is not None and ctx.needs_input_grad[0]) else None
        return (gx,) + (None,) * (len(ctx.needs_input_grad) - 1)


def mghs_forward(vt, input, plan, pixmask_fn):
    x, mlp_input = input[0], input[7]
    dev = x.device
    t_depth = trainer_of(vt.depth_net, lambda: T.DepthHeadTrainer(vt.depth_net, vt.D, dev), dev)
    drop = 0.5 if vt.height_net.training else 0.0        # nn.Dropout(0.5) behind the ASPP, depthnet.py:81
    t_height = trainer_of(vt.height_net, lambda: T.HeightNetTrainer(vt.height_net, dev, loss_weight=vt.loss_height_weight,
                                                                    dropout=drop), dev)
    layout = vt.out_layout if vt.collapse_z else 'ncdhw'
    params = _params(vt.depth_net) + _params(vt.height_net)
    return _MGHSFn.apply(x, vt, t_depth, t_height, plan, pixmask_fn, mlp_input, layout, *params)


# ------------------------------------------------------------------------------------------------ MGHS_Depth / MGHS_Stereo
class _MGHSDepthFn(torch.autograd.Function):
    """MGHS_Depth.forward (lss_heightmap.py:751-856) with a backward: pool backward -> camera-aware DepthNet (both
    gated branches, cost_volumn_net when stereo), the depth / height distributions' own gradients
    (get_depth_and_height_loss, LH:859-897) -> DepthNet / HeightNet.  The plane-sweep cost volume is a constant of the
    step (computed under no_grad, as the reference does, depthnet.py:405-407)."""

    @staticmethod
    def forward(ctx, x, vt, t_depth, t_height, plan, pixmask_fn, mlp_input, cost_volume, layout, *params):
        B, N, C, H, W = x.shape
        xa = D.pack_input(x.detach().reshape(B * N, C, H, W).float(), 1)
        depth, feat = t_depth.forward(xa, mlp_input, cost_volume)
        height = t_height.forward(xa, mlp_input)
        T.flush_bn_counters()
        pixmask = pixmask_fn(height)
        outs = plan.alloc_outputs(layout, x.device)
        plan.raw_forward(depth, feat, pixmask, outs, layout)
        ctx.state = (t_depth, t_height, plan, pixmask, depth, feat, layout, plan.workspace, (B, N, C, H, W))
        return tuple(outs) + (depth.clone(), height.clone())

    @staticmethod
    def backward(ctx, *gs):
        t_depth, t_height, plan, pixmask, depth, feat, layout, ws, (B, N, C, H, W) = ctx.state
        gouts, g_depth, g_height = list(gs[:-2]), gs[-2], gs[-1]
        want_dx = ctx.needs_input_grad[0]
        dgrad = fgrad = None
        if any(g is not None for g in gouts):
            dgrad, fgrad = plan._backward(depth, feat.view(B, N, H, W, -1), pixmask, gouts, layout, ws)
            dgrad = dgrad.view(B * N, -1, H, W)
            fgrad = fgrad.view(B * N, H, W, -1)
        if g_depth is not None:
            dgrad = g_depth if dgrad is None else dgrad + g_depth
        dx = None
        if dgrad is not None or fgrad is not None:
            dx = t_depth.backward(depth_grad=dgrad, feat_grad=fgrad, want_dx=want_dx)
        if g_height is not None:
            p = t_height.saved['height']
            dlog = p * (g_height - (p * g_height).sum(dim=1, keepdim=True))
            dz = t_height._act('dz', B * N, H, W, t_height.head.cout_pad)
            dz.data.zero_()
            dz.data[..., :dlog.shape[1]].copy_(dlog.permute(0, 2, 3, 1))
            t_height.dz = dz
            dxh = t_height.backward(want_dx=want_dx)
            if dxh is not None:
                dx = dxh if dx is None else D.Act(dx.data + dxh.data, dx.C, 1)
        gx = from_act(dx, C).reshape(B, N, C, H, W) if (dx is not None and want_dx) else None
        return (gx,) + (None,) * (len(ctx.needs_input_grad) - 1)


def mghs_depth_forward(vt, input, stereo_metas, plan, pixmask_fn):
    """-> (bev, bev_w_z, depth, height) of MGHS_Depth / MGHS_Stereo under autograd."""
    x, mlp_input = input[0], input[7]
    dev = x.device
    dn, hn = vt.depth_net, vt.height_net
    drop = 0.5 if vt.training else 0.0
    t_depth = trainer_of(dn, lambda: T.DepthNetTrainer(dn, dev, loss_weight=vt.loss_depth_weight, dropout=drop), dev)
    t_height = trainer_of(hn, lambda: T.HeightNetTrainer(hn, dev, loss_weight=vt.loss_height_weight, dropout=drop), dev)
    cv = None
    if dn.stereo:
        if stereo_metas is None:
            raise RuntimeError('DepthNet(stereo=True) called without stereo_metas')
        B, N, _, H, W = x.shape
        scale = float(stereo_metas['downsample']) / stereo_metas['cv_downsample']
        Hs, Ws = int(H * scale), int(W * scale)
        cv = D.Act(torch.zeros(B * N, Hs, Ws, t_depth.Dcv_pad, dtype=torch.bfloat16, device=dev), t_depth.Dcv_pad, 1)
        if stereo_metas['cv_feat_list'][0] is not None:            # zeros when there is no previous frame (depthnet.py:389-396)
            with torch.no_grad():
                dn.calculate_cost_volumn(stereo_metas, out_act=cv)
    elif stereo_metas is not None:
        raise RuntimeError('DepthNet(stereo=False) called with stereo_metas')
    if vt.collapse_z:
        raise NotImplementedError('MGHS_Depth under autograd: collapse_z=False (every DHD-M / DHD-L config)')
    params = _params(dn) + _params(hn)
    outs = _MGHSDepthFn.apply(x, vt, t_depth, t_height, plan, pixmask_fn, mlp_input, cv, 'ncdhw_cat', *params)
    return outs[0], outs[1], outs[2], outs[3]
