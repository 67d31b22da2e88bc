"""GPU parity tests of the fused occupancy-head tail (dhd_predictor_tail: Linear 256->512 + Softplus + Linear 512->288
+ per-z argmax in one back-to-back tcgen05 GEMM kernel; reference models/dense_heads/occ_head.py:63-67, 84-100, 141-153).

The kernel computes in bf16 operands / fp32 accumulation with the hidden layer rounded to bf16 (exactly what the
layer-by-layer bf16 engine does), so the checker is a float64 torch restatement that rounds to bf16 AT THE SAME
POINTS: what is left is the summation order and the ex2 / lg2 softplus (<= 1e-6) -- tolerance atol 1e-3 + rtol 1e-3
of logits of scale ~1 (a hidden value sitting on a bf16 rounding boundary may round the other way: one bf16 ulp of
one of 512 terms).  Class maps: bit-exact against softmax(-1).argmax(-1) of the kernel's own logits."""
import os

import pytest
import torch

from oracle import dense_oracle as DO

pytestmark = pytest.mark.gpu


def _head(seed):
    import projects.mmdet3d_plugin  # noqa: F401
    from projects.mmdet3d_plugin.models.dense_heads.occ_head import predictor
    head = predictor(in_dim=256, out_dim=256, Dz=16, num_classes=18, use_predicter=True, class_balance=False,
                     loss_occ=None, precision='bf16').eval()
    head.load_state_dict(DO.seeded_state_dict(head, seed))
    return head


def _reference(head, t_bf16):
    """t: (N, H, W, 256) bf16-valued activation = ReLU(final_conv(x)) -> logits (N, W, H, 16, 18) float64, rounding the
    weights and the hidden layer to bf16 where the kernel does."""
    r = lambda w: w.detach().to(torch.bfloat16).double()
    w1, b1 = r(head.predicter[0].weight), head.predicter[0].bias.detach().double()
    w2, b2 = r(head.predicter[2].weight), head.predicter[2].bias.detach().double()
    h = torch.nn.functional.softplus(t_bf16.double() @ w1.t() + b1)
    h = h.float().to(torch.bfloat16).double()
    y = h @ w2.t() + b2
    N, H, W, _ = y.shape
    return y.permute(0, 2, 1, 3).reshape(N, W, H, 16, 18)            # occ_pred.permute(0, 3, 2, 1) of the NCHW map


def _run_tail(head, t_bf16, want_logits=True, want_occ=True):
    from dhd_b200 import dense as D
    from dhd_b200.modules import PredictorEngine
    eng = PredictorEngine(head, 'bf16', 'cuda')
    assert eng.fused_tail_ok()
    N, H, W, C = t_bf16.shape
    t = D.Act(t_bf16.cuda().contiguous(), C, 1)
    logits = torch.full((N, W, H, 288), float('nan'), device='cuda') if want_logits else None
    occ = torch.full((N, W, H, 16), 255, dtype=torch.uint8, device='cuda') if want_occ else None
    D.predictor_tail(t, eng.fc0.w, eng.fc0.bias, eng.fc2.w, eng.fc2.bias, 16, 18, logits=logits, occ=occ)
    torch.cuda.synchronize()
    return (logits.view(N, W, H, 16, 18) if want_logits else None), occ


@pytest.mark.parametrize('N,H,W', [(1, 13, 21),        # 273 pixels: 3 tiles, the last one ragged
                                   (1, 8, 16),         # exactly one tile
                                   (2, 100, 100),      # 157 tiles on 148 CTAs: some CTAs run two tiles (barrier phases)
                                   (3, 37, 200)])      # 22 200 pixels, W != H: the (x, y) transposition
def test_fused_tail_matches_bf16_rounding_reference(cuda_lib, N, H, W):
    head = _head(61)
    g = torch.Generator().manual_seed(N * 1000 + H)
    t = torch.relu(torch.randn(N, H, W, 256, generator=g)).to(torch.bfloat16)
    want = _reference(head, t)
    logits, occ = _run_tail(head.cuda(), t)
    got = logits.double().cpu()
    assert not torch.isnan(got).any(), 'a pixel was not written'
    err = (got - want).abs()
    bad = err > 1e-3 + 1e-3 * want.abs()
    assert not bad.any(), '%d / %d logits off, max abs err %.3g (scale %.3g)' % (int(bad.sum()), bad.numel(), err.max(), want.abs().max())
    assert torch.equal(occ, logits.softmax(-1).argmax(-1).to(torch.uint8))      # torch's own ops on the GPU
    # the class map alone (the inference mode: logits never written) is the same map
    _, occ_only = _run_tail(head, t, want_logits=False)
    assert torch.equal(occ_only, occ)
    # and the logits alone
    logits_only, _ = _run_tail(head, t, want_occ=False)
    assert torch.equal(logits_only, logits)


def test_fused_tail_exact_ties_take_the_first_class(cuda_lib):
    """Zero weights in the last layer: every class of a z plane gets the same logit (the bias) -> argmax 0; a bias
    bump on two classes -> the lower index."""
    head = _head(62)
    with torch.no_grad():
        head.predicter[2].weight.zero_()
        b = torch.zeros(16, 18)
        b[3, 7] = b[3, 11] = 1.5
        b[5, 17] = 0.25
        b[9, :] = -2.0
        head.predicter[2].bias.copy_(b.view(-1))
    t = torch.relu(torch.randn(1, 16, 24, 256, generator=torch.Generator().manual_seed(1))).to(torch.bfloat16)
    logits, occ = _run_tail(head.cuda(), t)
    want = torch.zeros(16, dtype=torch.uint8)
    want[3], want[5] = 7, 17
    assert torch.equal(occ.cpu(), want.view(1, 1, 1, 16).expand(1, 24, 16, 16))
    assert torch.equal(logits.cpu(), b.view(1, 1, 1, 16, 18).expand(1, 24, 16, 16, 18))
    # near-ties 1..3 ulps apart in every z plane, the larger logit in the HIGHER class: softmax may merge them and
    # torch then answers the lower class -- the epilogue's exact path must agree with torch's own ops on the GPU
    with torch.no_grad():
        nb = torch.randn(16, 18, generator=torch.Generator().manual_seed(4)) * 0.05
        for z in range(16):
            lo_c, hi_c = z % 9, 9 + (z * 5) % 9
            top = nb[z].max() + 0.05 + 0.02 * z              # magnitude < 1: a few ulps are a gap below 6e-8
            nb[z, lo_c] = top
            nb[z, hi_c] = (top.view(torch.int32) + (z % 4)).view(torch.float32)
        head.predicter[2].bias.copy_(nb.view(-1).to(head.predicter[2].bias.device))
    logits, occ = _run_tail(head, t)
    want = logits.softmax(-1).argmax(-1).to(torch.uint8)
    assert torch.equal(occ, want)
    assert bool((want != logits.argmax(-1).to(torch.uint8)).any()), 'no plane exercised the softmax-rounding case'


def test_fused_tail_matches_layer_by_layer_engine_at_full_size(cuda_lib):
    """DHD-S B=4 (160 000 pixels, 1250 tiles, 8-9 tiles per CTA): the fused head against the layer-by-layer bf16
    engine (three dhd_conv2d_fwd GEMMs + dhd_occ_argmax) on the same input -- same roundings, other summation order."""
    from dhd_b200 import dense as D
    from dhd_b200.modules import PredictorEngine
    head = _head(63).cuda()
    eng = PredictorEngine(head, 'bf16', 'cuda')
    g = torch.Generator(device='cuda').manual_seed(5)
    x = D.Act(torch.randn(4, 200, 200, 256, device='cuda', generator=g).to(torch.bfloat16), 256, 1)
    occ_f = torch.empty(4, 200, 200, 16, dtype=torch.uint8, device='cuda')
    occ_l = torch.empty_like(occ_f)
    fused = eng(x, occ=occ_f)
    os.environ['DHD_TAIL_FUSED'] = '0'
    try:
        assert not eng.fused_tail_ok()
        layered = eng(x, occ=occ_l)
    finally:
        del os.environ['DHD_TAIL_FUSED']
    torch.cuda.synchronize()
    err = (fused - layered).abs()
    assert float(err.max()) <= 1e-3 + 1e-3 * float(layered.abs().max()), 'max abs diff %.3g' % float(err.max())
    assert torch.equal(occ_f, fused.softmax(-1).argmax(-1).to(torch.uint8))
    assert torch.equal(occ_l, layered.softmax(-1).argmax(-1).to(torch.uint8))
    agree = float((occ_f == occ_l).float().mean())
    assert agree > 0.995, 'fused / layered class maps agree on %.5f' % agree
    # inference mode of the engine: class map only, logits never written
    occ_only = torch.empty_like(occ_f)
    assert eng(x, occ=occ_only, want_logits=False) is None
    torch.cuda.synchronize()
    assert torch.equal(occ_only, occ_f)


def test_fused_tail_hidden_output_for_training(cuda_lib):
    """dhd_predictor_tail with `hidden`: the Softplus output it writes for the backward == the hidden layer of the
    layer-by-layer path (bf16 rounding of the same fp32 values up to the MUFU softplus: 1 bf16 ulp), and the logits are
    unchanged by asking for it."""
    from dhd_b200 import dense as D
    g = torch.Generator().manual_seed(21)
    N, H, W = 2, 37, 29
    x = D.pack_input(torch.randn(N, 256, H, W, generator=g).cuda(), 1)
    w1 = (torch.randn(512, 256, generator=g) / 16).cuda()
    w2 = (torch.randn(288, 512, generator=g) / 22).cuda()
    b1, b2 = torch.randn(512, generator=g).cuda(), torch.randn(288, generator=g).cuda()
    p1, p2 = D.pack_weight(w1[:, :, None, None], 1), D.pack_weight(w2[:, :, None, None], 1)
    lo_a = torch.empty(N, W, H, 288, device='cuda')
    lo_b = torch.empty(N, W, H, 288, device='cuda')
    hid = D.Act.empty(N, H, W, 512, 1, 'cuda')
    hid.data.fill_(float('nan'))
    D.predictor_tail(x, p1, b1, p2, b2, 16, 18, logits=lo_a)
    D.predictor_tail(x, p1, b1, p2, b2, 16, 18, logits=lo_b, hidden=hid)
    torch.cuda.synchronize()
    assert torch.equal(lo_a, lo_b)
    xin = x.data.float().view(N * H * W, 256)
    want = torch.nn.functional.softplus(xin @ w1.bfloat16().float().t() + b1)
    got = hid.data.float().view(N * H * W, 512)
    assert torch.isfinite(got).all()
    assert float((got - want).abs().max()) <= 2 ** -7 * float(want.abs().max())
