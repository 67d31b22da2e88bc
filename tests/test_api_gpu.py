"""GPU tests of the module / detector boundary (SURVEY 8(b) "module API"): the plugin modules are differentiable
nn.Modules like the reference's (MGHS.forward lss_heightmap.py:461-490, SFA.forward mix.py:87-90, predictor.forward
occ_head.py:84-100), and the DHD detector has the reference's methods (DHD_model.py:84-241) so the reference's runner
can drive it:  forward_train -> loss dict -> backward -> the gradients TrainStep's hand-wired pipeline produces."""
import numpy as np
import pytest
import torch

from oracle import dense_oracle as DO

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20))


def _detector_on(step):
    """A DHD detector whose children ARE the modules of a TrainStep (same parameters, same .grad views)."""
    import projects.mmdet3d_plugin  # noqa: F401
    from projects.mmdet3d_plugin.models.detectors.DHD_model import DHD
    from projects.mmdet3d_plugin.models.necks import Identity
    det = DHD.__new__(DHD)
    torch.nn.Module.__init__(det)
    det.img_backbone = det.img_neck = None
    det.img_view_transformer = step.vt
    det.img_bev_encoder_backbone, det.img_bev_encoder_neck = step.bev_backbone, step.bev_neck
    for i, u in enumerate(step.voxel):
        setattr(det, 'img_voxel_encoder%d' % i, u)
        setattr(det, 'img_voxel_neck%d' % i, Identity())
    det.mix, det.occ_head = step.sfa, step.head
    det.upsample, det.train_cfg, det.test_cfg = False, None, None
    return det


def test_forward_train_backward_matches_train_step_gradients(cuda_lib):
    """DHD.forward_train(img_inputs=..., voxel_semantics, mask_camera, gt_depth, gt_height) -> the reference's loss
    dict (DM:135-186) -> sum(losses).backward() through dhd_b200.autograd == the gradients of the hand-wired
    TrainStep (pipeline.py) on the same modules, inputs and labels (frozen-BatchNorm fine-tuning form; the two paths
    differ only in where the loss gradient is rounded to bf16: torch's fp32 losses vs the fused loss kernels)."""
    from dhd_b200 import synth
    from dhd_b200.pipeline import TrainStep
    cfg, B = synth.DHD_S, 1
    step = TrainStep(cfg, B, encoders=True, bn='frozen')
    host = step.make_host_inputs(synth.synthetic_rig(B, cfg['ncams'], cfg['input_size'], seed=3), seed=3)
    step.alloc_static(host)
    step.upload(host)
    step._fwd_bwd()
    torch.cuda.synchronize()
    want = step.bucket.flat.clone()
    loss_occ_want = step.loss.clone()            # [loss_occ, avg_factor, sem_scal, geo_scal]
    loss_h_want = float(step.loss_height)
    assert float(want.abs().max()) > 0

    step.bucket.zero()
    det = _detector_on(step).eval()              # eval(): frozen BatchNorm, no Dropout -- as TrainStep(bn='frozen')
    s = step.static
    x = s['x'].clone().requires_grad_(True)      # image features in the `imgs` slot (backbone outside the path)
    img_inputs = [x, s['sensor2ego'], s['ego2global'], s['cam2imgs'], s['post_rots'], s['post_trans'], s['bda']]
    losses = det.forward_train(img_inputs=img_inputs, img_metas=[{}] * B, voxel_semantics=step.labels,
                               mask_camera=step.mask_camera, gt_depth=step.gt_depth, gt_height=step.gt_height)
    assert set(losses) == {'loss_height', 'loss_occ', 'loss_voxel_sem_scal', 'loss_voxel_geo_scal'}
    assert abs(float(losses['loss_height']) - loss_h_want) <= 2e-3 * max(1.0, abs(loss_h_want))
    assert abs(float(losses['loss_occ']) - float(loss_occ_want[0])) <= 5e-3 * abs(float(loss_occ_want[0]))
    assert abs(float(losses['loss_voxel_sem_scal']) - float(loss_occ_want[2])) <= 5e-3 * abs(float(loss_occ_want[2]))
    assert abs(float(losses['loss_voxel_geo_scal']) - float(loss_occ_want[3])) <= 5e-3 * abs(float(loss_occ_want[3]))
    sum(losses.values()).backward()
    torch.cuda.synchronize()
    got = step.bucket.flat
    assert torch.isfinite(got).all() and x.grad is not None and torch.isfinite(x.grad).all()
    # per parameter group (the groups see very different gradient scales)
    o, worst = 0, {}
    names = {id(p): n for n, p in det.named_parameters()}
    for p in step.bucket.params:
        n = p.numel()
        g, w = got[o:o + n], want[o:o + n]
        o += n
        if float(w.norm()) == 0.0:
            assert float(g.norm()) == 0.0, names.get(id(p))
            continue
        top = names.get(id(p), '?').split('.')[0]
        worst[top] = max(worst.get(top, 0.0), _rel(g, w))
    assert worst and max(worst.values()) <= 3e-2, worst
    assert _rel(got, want) <= 1e-2, _rel(got, want)


def test_modules_are_differentiable_like_the_reference(cuda_lib):
    """SFA and predictor in train() mode (BatchNorm on batch statistics): forward + backward through the module
    call itself, gradients land in .grad, a second step after an in-place weight update sees the new weights."""
    import projects.mmdet3d_plugin  # noqa: F401
    from projects.mmdet3d_plugin.models.dense_heads.occ_head import predictor
    from projects.mmdet3d_plugin.models.necks.mix import SFA
    torch.manual_seed(0)
    sfa = SFA(512, 256).cuda().train()
    head = predictor(in_dim=256, out_dim=256, Dz=16, num_classes=18, use_predicter=True, class_balance=True,
                     loss_occ=dict(type='CrossEntropyLoss', use_sigmoid=False, ignore_index=255, loss_weight=1.0)).cuda().train()
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn(1, 512, 24, 40, device='cuda', generator=g)
    lab = torch.randint(0, 18, (1, 40, 24, 16), device='cuda', generator=g)
    mask = torch.rand(1, 40, 24, 16, device='cuda', generator=g) < 0.5
    opt = torch.optim.SGD(list(sfa.parameters()) + list(head.parameters()), lr=0.05)
    vals = []
    for _ in range(3):
        opt.zero_grad()
        occ = head(sfa(x))
        assert occ.shape == (1, 40, 24, 16, 18) and occ.requires_grad
        losses = head.loss(occ, lab, mask)
        total = sum(losses.values())
        total.backward()
        for n, p in list(sfa.named_parameters()) + list(head.named_parameters()):
            assert p.grad is not None and torch.isfinite(p.grad).all(), n
        vals.append(float(total))
        opt.step()
    assert vals[2] < vals[0], 'three SGD steps on one batch did not lower the loss: %s' % vals
    # BatchNorm really ran on batch statistics: the running means moved off their initial zeros, and the step counter
    # follows torch's (one bump per training-mode forward)
    assert float(sfa.mix_residual[1].running_mean.abs().max()) > 0
    assert int(sfa.mix_residual[1].num_batches_tracked) == 3


def test_engine_follows_in_place_weight_updates(cuda_lib):
    """the inference engines snapshot folded-BN bf16 weights; an in-place parameter update (optimizer.step,
    `with no_grad(): p.mul_()`), .to(device) and load_state_dict must all be seen by the next forward."""
    import projects.mmdet3d_plugin  # noqa: F401
    from projects.mmdet3d_plugin.models.dense_heads.occ_head import predictor
    from projects.mmdet3d_plugin.models.necks.mix import SFA
    sfa = SFA(512, 256).eval()
    head = predictor(in_dim=256, out_dim=256, Dz=16, num_classes=18, loss_occ=None).eval()
    sfa.load_state_dict(DO.seeded_state_dict(sfa, 1))
    head.load_state_dict(DO.seeded_state_dict(head, 2))
    sfa, head = sfa.cuda(), head.cuda()
    x = DO.seeded_tensor((1, 512, 8, 16), 3).cuda()
    a = head(sfa(x)).clone()
    assert torch.equal(head(sfa(x)), a)                          # cached engine, same result
    with torch.no_grad():
        head.predicter[2].bias.add_(1.0)                         # what an optimizer step does
    b = head(sfa(x))
    assert torch.allclose(b, a + 1.0, atol=1e-5), 'the head kept evaluating the old bias'
    with torch.no_grad():
        sfa.mix_shortcut[1].weight.mul_(0.5)                     # a BatchNorm affine parameter (folded into the conv)
    c = head(sfa(x))
    assert not torch.allclose(c, b, atol=1e-4), 'the SFA engine kept the old folded BatchNorm'
    sd = {k: v.clone() for k, v in sfa.state_dict().items()}
    sd['mix_shortcut.1.weight'] *= 2.0
    sfa.load_state_dict(sd)
    assert torch.allclose(head(sfa(x)), b, atol=1e-4)           # back to the previous weights
    # a runner that calls model.eval() before every test batch keeps the compiled engines; a real train() / eval() switch
    # (what follows a `.data` writer such as mmcv's EMA hook) drops them
    eng = sfa._engine
    sfa.eval()
    head(sfa(x))
    assert sfa._engine is eng, 'eval() on a module already in eval mode rebuilt the engine'
    sfa.train()
    sfa.eval()
    head(sfa(x))
    assert sfa._engine is not eng


def test_detector_reference_api_inference(cuda_lib):
    """DHD built from the DHD-S model config: simple_test / forward(return_loss=False) return the reference's list of
    (Dx, Dy, Dz) uint8 maps == softmax(-1).argmax(-1) of the logits forward_hot_path gives; train_step returns the
    runner's dict; an unregistered image backbone is a loud placeholder, not a silent None."""
    import projects.mmdet3d_plugin  # noqa: F401
    from dhd_b200 import compat as C
    from dhd_b200 import synth
    from tests.test_encoders_gpu import _dhd_s_model_cfg
    cfg = _dhd_s_model_cfg()
    cfg['img_backbone'] = dict(type='SwinTransformer', embed_dims=128)     # DHD-L's backbone: not part of this build
    model = C.DETECTORS.build(cfg).eval()
    model.load_state_dict(DO.seeded_state_dict(model, 77))
    model = model.cuda()
    B, N = 1, 6
    rig = [t.cuda() for t in synth.synthetic_rig(B, N, (256, 704), seed=5)]
    x = DO.seeded_tensor((B, N, 256, 16, 44), 78).cuda()
    if not C.HAVE_MMDET3D:
        with pytest.raises(NotImplementedError, match='not registered'):
            model.image_encoder(torch.zeros(B, N, 3, 256, 704, device='cuda'))
    img_inputs = [x] + rig
    with torch.no_grad():
        occ_list = model(return_loss=False, img_inputs=[img_inputs], img_metas=[[{}] * B])
        logits, _, _ = model.forward_hot_path(x, rig)
    assert isinstance(occ_list, list) and len(occ_list) == B
    assert occ_list[0].shape == (200, 200, 16) and occ_list[0].dtype == np.uint8
    want = logits.softmax(-1).argmax(-1).to(torch.uint8).cpu().numpy()
    assert np.array_equal(np.stack(occ_list), want)
    assert np.array_equal(np.stack(model.simple_test_occ(logits)), want)     # logits in, class maps out
    # the runner's entry point
    g = torch.Generator(device='cuda').manual_seed(2)
    data = dict(img_inputs=[x.clone().requires_grad_(True)] + rig, img_metas=[{}] * B,
                voxel_semantics=torch.randint(0, 18, (B, 200, 200, 16), device='cuda', generator=g),
                mask_camera=torch.rand(B, 200, 200, 16, device='cuda', generator=g) < 0.5,
                gt_depth=torch.where(torch.rand(B, N, 256, 704, device='cuda', generator=g) < 0.02,
                                     1.0 + 40.0 * torch.rand(B, N, 256, 704, device='cuda', generator=g), torch.zeros((), device='cuda')),
                gt_height=torch.where(torch.rand(B, N, 256, 704, device='cuda', generator=g) < 0.02,
                                      -1.0 + 6.0 * torch.rand(B, N, 256, 704, device='cuda', generator=g), torch.zeros((), device='cuda')))
    out = model.train_step(data, None)
    assert set(out) == {'loss', 'log_vars', 'num_samples'} and out['num_samples'] == B
    assert {'loss_height', 'loss_occ', 'loss_voxel_sem_scal', 'loss_voxel_geo_scal', 'loss'} <= set(out['log_vars'])
    out['loss'].backward()
    assert model.occ_head.predicter[2].weight.grad is not None
    assert model.img_view_transformer.depth_net.weight.grad is not None


def test_dhd_stereo_forward_train_two_frames(cuda_lib):
    """DHD_stereo (DHD-M wiring at reduced image size: two temporal frames + the stereo reference frame) through the
    reference's entry point: forward_train -> dict(loss_depth, loss_height, loss_occ, sem_scal, geo_scal)
    (DHD_model.py:577-614) -> backward reaches the camera-aware DepthNet (incl. cost_volumn_net), HeightNet, both
    pre-process nets, the encoders, SFA and the head; the previous frame runs under no_grad on the inference engines."""
    import projects.mmdet3d_plugin  # noqa: F401
    from dhd_b200 import compat as C
    from dhd_b200 import synth
    from tests.test_encoders_gpu import _dhd_s_model_cfg
    c, D_, size = 64, 88, (128, 352)
    cfg = _dhd_s_model_cfg()
    grid = dict(cfg['img_view_transformer']['grid_config'], depth=[1.0, 45.0, 0.5])
    vt = dict(cfg['img_view_transformer'], type='MGHS_Stereo', grid_config=grid, input_size=size, collapse_z=False,
              loss_depth_weight=0.05, depthnet_cfg=dict(use_dcn=False, aspp_mid_channels=96, stereo=True, bias=5.0),
              heightnet_cfg=dict(use_dcn=False, aspp_mid_channels=96))
    for k in ('mask_1_grid', 'mask_2_grid', 'mask_3_grid'):
        vt[k] = dict(vt[k], depth=[1.0, 45.0, 0.5])
    cfg.update(type='DHD_stereo', img_view_transformer=vt, num_adj=1, align_after_view_transfromation=False,
               pre_process=dict(type='CustomResNet', numC_input=c, num_layer=[1], num_channels=[c], stride=[1], backbone_output_ids=[0]),
               pre_process_net_3d=dict(type='CustomResNet', numC_input=c * 16, num_layer=[1], num_channels=[c * 16], stride=[1],
                                       backbone_output_ids=[0]),
               img_bev_encoder_backbone=dict(type='CustomResNet', numC_input=c * 2, num_channels=[c * 2, c * 4, c * 8]),
               img_voxel_encoder0_backbone=dict(type='UNet', n_channels=c * 8, n_classes=64),
               img_voxel_encoder1_backbone=dict(type='UNet', n_channels=c * 8, n_classes=128),
               img_voxel_encoder2_backbone=dict(type='UNet', n_channels=c * 16, n_classes=64))
    model = C.DETECTORS.build(cfg)
    model.load_state_dict(DO.seeded_state_dict(model, 91))
    model = model.cuda().train()
    B, N, nf = 1, 6, model.num_frame                       # 3 = key + previous + stereo reference
    fH, fW = size[0] // 16, size[1] // 16
    rig = synth.synthetic_rig(B, N, size, seed=7)
    rep = lambda t: torch.cat([t] * nf, dim=1).cuda()      # the same rig for every frame (ego at rest)
    s2e, e2g, K, pr, pt, bda = rig
    feats = DO.seeded_tensor((B, N * nf, 256, fH, fW), 92).cuda().requires_grad_(True)
    stereo = DO.seeded_tensor((B, N * nf, 64, 4 * fH, 4 * fW), 93).cuda()
    img_inputs = [(feats, stereo), rep(s2e), rep(e2g), rep(K), rep(pr), rep(pt), bda.cuda()]
    g = torch.Generator(device='cuda').manual_seed(4)
    hit = torch.rand(B, N, *size, device='cuda', generator=g) < 0.03
    kw = dict(voxel_semantics=torch.randint(0, 18, (B, 200, 200, 16), device='cuda', generator=g),
              mask_camera=torch.rand(B, 200, 200, 16, device='cuda', generator=g) < 0.5,
              gt_depth=torch.where(hit, 1.0 + 40.0 * torch.rand(B, N, *size, device='cuda', generator=g), torch.zeros((), device='cuda')),
              gt_height=torch.where(hit, -1.0 + 6.0 * torch.rand(B, N, *size, device='cuda', generator=g), torch.zeros((), device='cuda')))
    losses = model.forward_train(img_inputs=img_inputs, img_metas=[{}] * B, **kw)
    assert set(losses) == {'loss_depth', 'loss_height', 'loss_occ', 'loss_voxel_sem_scal', 'loss_voxel_geo_scal'}
    total = sum(losses.values())
    assert torch.isfinite(total)
    total.backward()
    vtm = model.img_view_transformer
    for name, p in (('depth_net.context_conv', vtm.depth_net.context_conv.weight),
                    ('depth_net.cost_volumn_net', vtm.depth_net.cost_volumn_net[0].weight),
                    ('depth_net.depth_conv.0.downsample', vtm.depth_net.depth_conv[0].downsample.weight),
                    ('height_net head', list(vtm.height_net.depth_conv)[-1].weight),
                    ('pre_process_net', model.pre_process_net.layers[0][0].conv1.weight),
                    ('pre_process_net_3d', model.pre_process_net_3d.layers[0][0].conv1.weight),
                    ('voxel encoder 2', model.img_voxel_encoder2.inc.double_conv[0].weight),
                    ('occ head', model.occ_head.predicter[2].weight)):
        assert p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().max()) > 0, name
    assert feats.grad is not None and float(feats.grad[:, :N].abs().max()) >= 0.0
    # inference through the same detector (eval, no_grad): the reference's list of uint8 maps
    model.eval()
    with torch.no_grad():
        occ = model.simple_test(None, [{}] * B, img=img_inputs)
    assert len(occ) == B and occ[0].shape == (200, 200, 16) and occ[0].dtype == np.uint8


def test_dhd_stereo_inference_act_path_equals_tensor_path(cuda_lib):
    """DHD_stereo.simple_test in the bf16 speed mode (the DHD-L wiring of dhd_b200.synth at a reduced image size): the
    inference fast path -- pool kernel writing bf16 NHWC activations with z collapsed into channels, pre-process nets,
    frame concatenation, encoders writing channel slices of the SFA input -- against the module-by-module tensor path
    (fp32 (B, C, Dz, Dy, Dx) tensors, `torch.cat(x.unbind(dim=2), 1)`, DHD_model.py:313-374, 517-541).  Both round the
    pooled sums and every module output to bf16 at the same points, so the class maps agree."""
    import projects.mmdet3d_plugin  # noqa: F401
    from dhd_b200 import synth
    from dhd_b200.detector_step import DetectorStep
    cfg = synth.dhd_l_model_cfg('bf16')
    cfg['img_view_transformer'] = dict(cfg['img_view_transformer'], input_size=(128, 352))
    step = DetectorStep(cfg, 1, seed=3)
    step.model.load_state_dict(DO.seeded_state_dict(step.model, 17))
    img_inputs, _ = step.make_inputs(5)
    model = step.model
    assert model.eval()._act_path_ok()
    fast = step.infer_step(img_inputs)
    model.act_path = False
    assert not model._act_path_ok()
    slow = step.infer_step(img_inputs)
    model.act_path = True
    assert len(fast) == len(slow) == 1 and fast[0].shape == (200, 200, 16) and fast[0].dtype == np.uint8
    agree = float((np.asarray(fast[0]) == np.asarray(slow[0])).mean())
    assert len(np.unique(np.asarray(slow[0]))) >= 2                       # a non-trivial class map
    assert agree >= 0.9995, agree


def test_dhd_inference_act_path_and_cuda_graph_equal_the_tensor_path(cuda_lib):
    """DHD.simple_test (DHD-S, bf16 speed mode): (1) the activation fast path (pool kernel -> bf16 NHWC activations ->
    encoders writing channel slices of the SFA input) against the reference-shaped tensor path of extract_img_feat
    (DM:84-114); (2) the whole step captured as ONE CUDA graph (DetectorStep.capture_infer: no host synchronisation left
    in prepare_inputs / the view transformer) replays to exactly the eager result, also for new inputs."""
    import projects.mmdet3d_plugin  # noqa: F401
    from dhd_b200 import synth
    from dhd_b200.detector_step import DetectorStep
    step = DetectorStep(synth.dhd_s_model_cfg('bf16', images=False), 1, seed=3)
    step.model.load_state_dict(DO.seeded_state_dict(step.model, 19))
    a, _ = step.make_inputs(5)
    b, _ = step.make_inputs(6)
    model = step.model
    assert model.eval()._dhd_act_path_ok()
    fast = step.infer_step(a)
    model.act_path = False
    slow = step.infer_step(a)
    model.act_path = True
    agree = float((np.asarray(fast[0]) == np.asarray(slow[0])).mean())
    assert fast[0].shape == (200, 200, 16) and len(np.unique(np.asarray(slow[0]))) >= 2 and agree >= 0.9995, agree
    assert step.capture_infer(a), getattr(step, 'capture_error', '')
    assert np.array_equal(step.infer_step_graphed()[0], fast[0])
    want_b = step.infer_step(b)
    assert not np.array_equal(want_b[0], fast[0])
    assert np.array_equal(step.infer_step_graphed(b)[0], want_b[0])
