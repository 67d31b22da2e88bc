"""The hot path as one re-runnable inference step (bench.py and the multi-GPU harness).

Per step and per GPU (DHD-S, B samples = 6B camera images), everything through the C-ABI library:

  front   pack (fp32 NCHW image features -> NHWC split-bf16)            dhd_pack_nchw_to_nhwc
          depth_net 1x1 + softmax / context split                       dhd_conv2d_fwd        (a3)
          HeightNet (19 tcgen05 convs, SE gate, ASPP, DCN) + softmax    dhd_conv2d_fwd, ...   (a4)
  pool    height -> mask id                                             dhd_height_to_mask    (a5)
          geometry + binning of the four grids                          dhd_mghs_prepare      (a7, a8)
          fused masked lift-splat, four BEV tensors, single write       dhd_mghs_pool_fwd     (a6, a9, a11)
  back    [BEV / voxel encoders: outside the path -> resident synthetic (B,512,Dy,Dx) features]
          SFA (squeeze, 5 convs + gates), predictor (3 GEMMs)            dhd_conv2d_fwd, ...   (a14, a15)
          [bf16 mode: Linear + Softplus + Linear + per-z argmax fused    dhd_predictor_tail]
          class map (argmax over 18 classes) uint8                      dhd_occ_argmax (split-bf16 modes)

The front and the back of the step are captured into two CUDA graphs (launch-bound otherwise:
~75 kernels); the pool kernel between them is an eager launch so it can be timed in place.
The pool backward (a10) is timed separately by bench.py (`extras`); the training step (forward with
saved activations, losses, backward, gradient all-reduce, AdamW) is `TrainStep` below.
"""
import ctypes
import json
import os

import torch

from . import _lib
from . import dense as D
from .modules import DepthHeadEngine, HeightNetEngine, PredictorEngine, SFAEngine
from .pool import MghsPool, height_to_mask

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def algorithmic_bytes(cfg, B):
    """SURVEY.md 8(d) / BASELINE.md section 3: every output element written exactly once,
    every input element read once; index tables, bins and workspaces are not counted."""
    N = cfg['ncams']
    h, w = cfg['input_size']
    fH, fW = h // cfg['downsample'], w // cfg['downsample']
    D_ = len(range(int(cfg['depth'][0]), int(cfg['depth'][1]), int(cfg['depth'][2])))
    C = cfg['C']
    P = B * N * fH * fW
    grids = [cfg['bev_grid']] + list(cfg['mask_grids'])
    out_elems = 0
    for g in grids:
        dx = int(round((g['x'][1] - g['x'][0]) / g['x'][2]))
        dy = int(round((g['y'][1] - g['y'][0]) / g['y'][2]))
        dz = int(round((g['z'][1] - g['z'][0]) / g['z'][2]))
        out_elems += B * dz * dy * dx * C
    reads = P * D_ * 4 + P * C * 4 + P
    return {'pool_fwd_bytes': reads + out_elems * 4,
            'pool_bwd_bytes': out_elems * 4 + reads + P * D_ * 4 + P * C * 4,
            'pool_out_bytes': out_elems * 4}


def dense_flops(cfg, B, Dy=200, Dx=200):
    """Algorithmic forward FLOPs (2*M*N*K per conv) of the dense stages, for the report."""
    N = cfg['ncams']
    h, w = cfg['input_size']
    M = B * N * (h // cfg['downsample']) * (w // cfg['downsample'])
    c = 256
    hn = M * c * c * 2 * (9 + 6 * 9 + 1 + 3 * 9 + 4 + 9 / 4) + M * c * 2 * (18 * 9 + 65)
    dn = M * c * 108 * 2
    Mb = B * Dy * Dx
    sfa = Mb * 2 * (256 * 256 * 2 + 512 * 256 + 2 * 9 * 256 * 256)
    head = Mb * 2 * (9 * 256 * 256 + 256 * 512 + 512 * 288)
    return {'heightnet': hn, 'depth_net': dn, 'sfa': sfa, 'predictor': head}


def encoder_flops(B, Dy=200, Dx=200, c=64):
    """Algorithmic forward FLOPs of the BEV encoder (CustomResNet + FPN_LSS) and the three UNets at DHD-S sizes."""
    def conv(hw, cin, cout, k=3):
        return 2.0 * B * hw * cin * cout * k * k
    hw = [Dy * Dx // (4 ** i) for i in range(5)]
    hw[4] = (Dy // 16) * (Dx // 16)
    tot = 0.0
    cin = c
    for i, ch in enumerate((2 * c, 4 * c, 8 * c)):                      # CustomResNet
        tot += conv(hw[i + 1], cin, ch) * 2 + conv(hw[i + 1], ch, ch) * 3
        cin = ch
    tot += conv(hw[1], 10 * c, 512) + conv(hw[1], 512, 512) + conv(hw[0], 512, 256) + conv(hw[0], 256, 256, 1)
    for nin, ncls in ((4 * c, 64), (4 * c, 128), (8 * c, 64)):          # UNets
        chans = [64, 128, 256, 512, 1024]
        tot += conv(hw[0], nin, 64) + conv(hw[0], 64, 64)
        for k in range(1, 5):
            tot += conv(hw[k], chans[k - 1], chans[k]) + conv(hw[k], chans[k], chans[k])
        for k in range(3, -1, -1):
            tot += conv(hw[k + 1], chans[k + 1], chans[k], 2) + conv(hw[k], 2 * chans[k], chans[k]) + conv(hw[k], chans[k], chans[k])
        tot += conv(hw[0], 64, ncls, 1)
    return tot


def pool_traffic_bytes():
    """dram read+write bytes per launch of the pool kernel from the committed ncu capture."""
    p = os.path.join(_ROOT, 'profiles', 'pool_fwd_traffic.json')
    if os.path.exists(p):
        try:
            return json.load(open(p)).get('dram_bytes_per_launch')
        except Exception:  # noqa: BLE001
            return None
    return None


class HotPathStep:
    def __init__(self, cfg, B, precision='bf16', deterministic=True, device='cuda', seed=0,
                 use_graph=True, encoders=False, keep_logits=False, images=False, backbone_depth=50):
        from projects.mmdet3d_plugin.models.dense_heads.occ_head import predictor
        from projects.mmdet3d_plugin.models.necks.lss_heightmap import MGHS
        from projects.mmdet3d_plugin.models.necks.mix import SFA
        self.cfg, self.B, self.precision, self.deterministic = cfg, B, precision, deterministic
        self.parts = D.PRECISIONS[precision][0]
        self.N = cfg['ncams']
        h, w = cfg['input_size']
        self.fH, self.fW = h // cfg['downsample'], w // cfg['downsample']
        self.C, self.Cin = cfg['C'], cfg['C_in']
        self.device = torch.device(device)
        g = cfg['mask_grids']
        torch.manual_seed(seed)               # random-init weights of the DHD-S architecture
        self.vt = MGHS(grid_config=dict(cfg['bev_grid'], depth=cfg['depth']), input_size=cfg['input_size'],
                       in_channels=self.Cin, out_channels=self.C, height_range=cfg['height_range'],
                       height_interval=0.1, mask_range=cfg['mask_range'],
                       mask_1_grid=dict(g[0], depth=cfg['depth']), mask_2_grid=dict(g[1], depth=cfg['depth']),
                       mask_3_grid=dict(g[2], depth=cfg['depth']), downsample=cfg['downsample'],
                       loss_height_weight=0.1,           # DHD-S.py:100
                       precision=precision).eval().to(self.device)
        self.sfa = SFA(512, 256, precision=precision).eval().to(self.device)
        self.head = predictor(in_dim=256, out_dim=256, Dz=16, num_classes=18, use_predicter=True,
                              class_balance=False, loss_occ=None, precision=precision).eval().to(self.device)
        self.encoders = encoders
        self.images = images                  # True: the step starts from the camera IMAGES (ResNet + CustomFPN, 8(f)-4)
        if images:
            # img_backbone / img_neck of DHD-S.py:44-62 (ResNet-50, out_indices (2, 3) -> CustomFPN 256 ch at 1/16)
            from projects.mmdet3d_plugin.models.backbones.image_resnet import ResNet
            from projects.mmdet3d_plugin.models.necks.fpn import CustomFPN
            from .backbone import CustomFPNEngine, ResNetEngine
            self.img_backbone = ResNet(depth=backbone_depth, out_indices=(2, 3), precision=precision).eval().to(self.device)
            self.img_neck = CustomFPN(in_channels=[1024, 2048], out_channels=self.Cin, num_outs=1, start_level=0, out_ids=[0],
                                      precision=precision).eval().to(self.device)
            self.e_img_backbone = ResNetEngine(self.img_backbone, precision, self.device)
            self.e_img_neck = CustomFPNEngine(self.img_neck, precision, self.device)
        self.keep_logits = keep_logits        # True: also write the fp32 logits (184 MB at B=4) -- parity tests
        if encoders:
            # the widened path: the real BEV encoder and the three voxel encoders of
            # DHD-S.py:106-131 between the pool and the SFA instead of resident stand-in features
            from projects.mmdet3d_plugin.models.backbones import CustomResNet, UNet
            from projects.mmdet3d_plugin.models.necks import FPN_LSS
            from .encoders import CustomResNetEngine, FPNLSSEngine, UNetEngine
            c = self.C
            self.bev_backbone = CustomResNet(c, num_channels=[2 * c, 4 * c, 8 * c], precision=precision).eval().to(self.device)
            self.bev_neck = FPN_LSS(8 * c + 2 * c, 256, precision=precision).eval().to(self.device)
            self.voxel = [UNet(4 * c, 64, precision=precision), UNet(4 * c, 128, precision=precision),
                          UNet(8 * c, 64, precision=precision)]
            self.voxel = [u.eval().to(self.device) for u in self.voxel]
            self.e_backbone = CustomResNetEngine(self.bev_backbone, precision, self.device)
            self.e_neck = FPNLSSEngine(self.bev_neck, precision, self.device)
            self.e_voxel = [UNetEngine(u, precision, self.device) for u in self.voxel]
        self.D = self.vt.D
        self.depth_engine = DepthHeadEngine(self.vt.depth_net, self.D, precision, self.device)
        self.height_engine = HeightNetEngine(self.vt.height_net, precision, self.device)
        self.sfa_engine = SFAEngine(self.sfa, precision, self.device)
        self.head_engine = PredictorEngine(self.head, precision, self.device)
        grids = [cfg['bev_grid']] + list(g)
        self.plan = MghsPool(B, self.N, self.D, self.fH, self.fW, self.C, grids[0]['x'], grids[0]['y'],
                             [(gr['z'], m) for m, gr in enumerate(grids)])
        self.Dy, self.Dx = self.plan.Dy, self.plan.Dx
        self.workspace = torch.empty(self.plan.ws_bytes, dtype=torch.uint8, device=self.device)
        # with the real encoders behind it (bf16 speed mode) the pool writes bf16 NHWC directly: the four outputs ARE
        # the encoders' input activations (half the bytes, no conversion pass); everywhere else fp32, as the reference
        self.pool_layout = 'nhwc_bf16' if (encoders and self.parts == 1 and type(self).__name__ == 'HotPathStep') else 'nhwc'
        self.outs = self.plan.alloc_outputs(self.pool_layout, self.device)
        self.frustum = self.vt.frustum.to(self.device)
        gen = torch.Generator(device=self.device).manual_seed(7)
        # stand-in for the BEV / voxel encoders' output (outside the SURVEY 8 path): channels-last fp32 values, held
        # the way dhd_b200.encoders hands them to the SFA -- one bf16 NHWC activation
        self.encoded = torch.randn(B, self.Dy, self.Dx, 512, device=self.device, generator=gen)
        self.encoded_act = D.pack_nhwc(self.encoded, self.parts)
        self.occ = torch.empty(B, self.Dx, self.Dy, 16, dtype=torch.uint8, device=self.device)
        self.host_occ = torch.empty(self.occ.shape, dtype=torch.uint8, pin_memory=True)
        self.result, self.host_result = self.occ, self.host_occ      # what a step hands back to the host
        # pool backward leg (timed separately)
        self.gouts = [torch.randn(o.shape, device=self.device, generator=gen) for o in self.outs]
        self.depth_grad = torch.empty(B * self.N, self.D, self.fH, self.fW, device=self.device)
        self.feat_grad = torch.empty(B, self.N, self.fH, self.fW, self.C, device=self.device)
        self.h2d_bytes = 0
        self.d2h_bytes = self.host_occ.numel()
        self.use_graph = use_graph
        self.graph = None
        self.static = None
        self.launches_per_step = None
        self._last = {}

    def stage_names(self):
        mid = ['split(pool outputs -> bf16)', 'CustomResNet + FPN_LSS (BEV encoder)', '3x UNet (voxel encoders)'] \
            if self.encoders else ['[encoder stand-in: resident bf16 NHWC features]']
        first = ['ResNet-%d image backbone (stem im2col + tcgen05 GEMM, MaxPool, Bottleneck stages)' % self.img_backbone.depth,
                 'CustomFPN'] if self.images else ['pack']
        return first + ['depth_net(1x1+softmax)', 'HeightNet(+softmax)', 'height_to_mask',
                'mghs_prepare(geometry+binning, 4 grids)', 'mghs_pool_fwd(nhwc, fused 4-pass)'] + mid + \
            ['SFA', 'predictor 3x3 conv', 'fused head tail: Linear+Softplus+Linear+per-z argmax (dhd_predictor_tail)'
             if self.head_engine.fused_tail_ok() and not self.keep_logits else 'predictor MLP (2 GEMMs) + occ_argmax']

    # ---- inputs -----------------------------------------------------------------------------
    def make_host_inputs(self, rig, seed):
        """Pinned host buffers of one step: image features + camera geometry."""
        s2e, e2g, K, pr, pt, bda = rig
        g = torch.Generator().manual_seed(seed)
        if self.images:
            ih, iw = self.cfg['input_size']
            x = torch.randn(self.B, self.N, 3, ih, iw, generator=g)       # normalised camera images
        else:
            x = torch.randn(self.B, self.N, self.Cin, self.fH, self.fW, generator=g)
        h = {'x': x, 'sensor2ego': s2e.contiguous(), 'ego2global': e2g.contiguous(), 'cam2imgs': K.contiguous(),
             'post_rots': pr.contiguous(), 'post_trans': pt.contiguous(), 'bda': bda.contiguous()}
        h = {k: v.pin_memory() for k, v in h.items()}
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in h.values())
        return h

    def alloc_static(self, host):
        self.static = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host.items()}
        return self.static

    def upload(self, host):
        for k, v in host.items():
            self.static[k].copy_(v, non_blocking=True)

    # ---- one step ---------------------------------------------------------------------------
    def _front(self):
        s = self.static
        B, N = self.B, self.N
        # geometry + binning depend on the camera tensors only: a side stream (a parallel branch of the graph)
        # runs them under the HeightNet convolutions, which hold one CTA per SM and leave its other slots idle
        main = torch.cuda.current_stream()
        if not hasattr(self, '_prep_stream'):
            self._prep_stream = torch.cuda.Stream()

        # the per-camera 3x3s (the reference's own torch.inverse / matmul calls: ~25 tiny library launches) start at once
        self._prep_stream.wait_stream(main)
        with torch.cuda.stream(self._prep_stream):
            # ... and so does the camera-aware SE gate of HeightNet (get_mlp_input + Mlp + SELayer: camera tensors only)
            mlp = self.vt.get_mlp_input(s['sensor2ego'], s['ego2global'], s['cam2imgs'], s['post_rots'],
                                        s['post_trans'], s['bda'])
            gate = self.height_engine.gate(mlp)
            gate.record_stream(main)
            gate_ready = torch.cuda.Event()
            gate_ready.record(self._prep_stream)
            cam_mats = self.plan.camera_matrices(s['sensor2ego'], s['cam2imgs'], s['post_rots'], s['post_trans'], s['bda'])

        def fork_prepare():        # late in HeightNet: the bins are still in L2 when the pool starts
            self._prep_stream.wait_stream(main)
            with torch.cuda.stream(self._prep_stream):
                self.plan.prepare(frustum=self.frustum, cam_mats=cam_mats,
                                  deterministic=self.deterministic, workspace=self.workspace)
        if self.images:
            feats = self.e_img_backbone(s['x'].view(B * N, 3, *self.cfg['input_size']))
            xa = self.e_img_neck(feats)[0]
        else:
            xa = D.pack_input(s['x'].view(B * N, self.Cin, self.fH, self.fW), self.parts)
        # depth_net's 1x1 (132 tiles) rides in the launch of HeightNet's first BasicBlock convolution
        depth, feat, depth_conv = self.depth_engine(xa, defer=True)
        main.wait_event(gate_ready)
        height = self.height_engine(xa, mlp, softmax=True, hook=fork_prepare, gate=gate, batch_with=[depth_conv])
        pixmask = height_to_mask(height, self.cfg['height_range'], self.cfg['mask_range'])
        main.wait_stream(self._prep_stream)
        self._last = {'depth': depth, 'feat': feat, 'height': height, 'pixmask': pixmask}

    def _pool(self):
        L = self._last
        self.plan.raw_forward(L['depth'], L['feat'], L['pixmask'], self.outs, self.pool_layout, workspace=self.workspace)

    def _encode(self, parallel=True):
        """pool outputs -> (B, 512, Dy, Dx) = cat(bev encoder, three voxel encoders), DM:103-114: every encoder
        writes its channel slice of one NHWC buffer."""
        B, H, W = self.B, self.Dy, self.Dx
        if not hasattr(self, '_enc_act'):
            self._enc_act = D.Act.empty(B, H, W, 512, self.parts, self.device)
        enc = self._enc_act
        if self.pool_layout == 'nhwc_bf16':
            ins = [D.Act(o, o.shape[-1], 1) for o in self.outs]        # the pool wrote the activations themselves
        else:
            ins = [D.pack_nhwc(o, self.parts) for o in self.outs]      # fp32 pool outputs -> split-bf16 activations
        # the four encoders are independent: the three UNets run on side streams (parallel branches of the CUDA
        # graph), so their small deep levels (12x12 .. 50x50 maps: 5-80 tiles for 148 SMs) share the GPU with the
        # others' full-resolution layers instead of each leaving most SMs idle
        main = torch.cuda.current_stream()
        if not hasattr(self, '_enc_streams'):
            self._enc_streams = [torch.cuda.Stream() for _ in self.e_voxel]
        lo = 256
        for st, e, x in zip(self._enc_streams, self.e_voxel, ins[1:]):
            st = st if parallel else main
            st.wait_stream(main)
            with torch.cuda.stream(st):
                e(x, out=enc.slice(lo, lo + e.n_classes))
            lo += e.n_classes
        self.e_neck(self.e_backbone(ins[0]), out=enc.slice(0, 256))
        for st in self._enc_streams:
            main.wait_stream(st)
        return enc

    def _back(self):
        enc = self._encode() if self.encoders else self.encoded_act
        fused = self.sfa_engine(enc)
        # inference tail (occ_head.py:141-153): the class map is what leaves the step.  bf16 speed mode: ONE fused kernel
        # (Linear + Softplus + Linear + per-z argmax, dhd_predictor_tail) -- the logits never reach HBM; the split-bf16
        # precision modes keep the layer-by-layer head + dhd_occ_argmax and expose the logits (parity tests)
        fused_tail = self.head_engine.fused_tail_ok() and not self.keep_logits
        self._logits = self.head_engine(fused, occ=self.occ, want_logits=not fused_tail)

    def capture(self):
        """Warm up eagerly, count this library's launches for one step, then capture the front and
        the back of the step into two CUDA graphs; the pool kernel between them stays an eager
        launch so bench.py can bracket it with CUDA events inside the timed region."""
        lib = _lib.load()
        for _ in range(2):
            n0 = lib.dhd_launch_count()
            self._front(); self._pool(); self._back()
            torch.cuda.synchronize()
            self.launches_per_step = lib.dhd_launch_count() - n0
        if not self.use_graph:
            return False
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._front(); self._pool(); self._back()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                self._front()
            with torch.cuda.graph(g2):
                self._back()
            self.graph = (g1, g2)
            return True
        except Exception as e:  # noqa: BLE001  (capture unsupported by some torch op: run eagerly)
            self.graph = None
            self.graph_error = repr(e)[:300]
            torch.cuda.synchronize()
            return False

    def run(self, pool_events=None):
        st = torch.cuda.current_stream()
        if self.graph is not None:
            self.graph[0].replay()
        else:
            self._front()
        if pool_events is not None:
            pool_events[0].record(st)
        self._pool()
        if pool_events is not None:
            pool_events[1].record(st)
        if self.graph is not None:
            self.graph[1].replay()
        else:
            self._back()

    def run_e2e(self, host):
        """Host buffers in, host buffer out: H2D of the step inputs, the step, D2H of the class map
        (all on the compute stream, nothing overlapped)."""
        self.upload(host)
        self.run()
        self.host_result.copy_(self.result, non_blocking=True)

    # ---- streamed end-to-end: the copies of neighbouring steps overlap the compute of this one ----
    def e2e_open(self, host):
        """Start a stream of steps: queue the H2D copy of the first step's inputs."""
        if not hasattr(self, '_h2d'):
            self._h2d, self._d2h = torch.cuda.Stream(), torch.cuda.Stream()
            self._stage = [{k: torch.empty_like(v) for k, v in self.static.items()} for _ in range(2)]
            self._h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_free = [torch.cuda.Event(), torch.cuda.Event()]
            self._d2h_done, self._step_done = torch.cuda.Event(), torch.cuda.Event()
        main = torch.cuda.current_stream()
        self._i = 0
        for e in self._stage_free:
            e.record(main)
        self._d2h_done.record(main)
        self._queue_h2d(host, 0)

    def _queue_h2d(self, host, slot):
        self._h2d.wait_event(self._stage_free[slot])
        with torch.cuda.stream(self._h2d):
            for k, v in host.items():
                self._stage[slot][k].copy_(v, non_blocking=True)
            self._h2d_done[slot].record(self._h2d)

    def run_e2e_streamed(self, host, next_host=None):
        """One step of the stream: inputs arrive from pinned host memory (copied on a side stream
        while the previous step computed), the class map leaves to pinned host memory on another;
        `next_host` = the following step's inputs (None for the last step).  Every step still
        moves h2d_bytes in and d2h_bytes out."""
        main = torch.cuda.current_stream()
        slot = self._i & 1
        main.wait_event(self._h2d_done[slot])
        for k, v in self._stage[slot].items():          # device-to-device into the graph's input buffers
            self.static[k].copy_(v, non_blocking=True)
        self._stage_free[slot].record(main)
        if next_host is not None:
            self._queue_h2d(next_host, slot ^ 1)
        main.wait_event(self._d2h_done)                 # the previous result has left the device buffer
        self.run()
        self._step_done.record(main)
        self._d2h.wait_event(self._step_done)
        with torch.cuda.stream(self._d2h):
            self.host_result.copy_(self.result, non_blocking=True)
            self._d2h_done.record(self._d2h)
        self._i += 1

    def e2e_close(self):
        """Join the side streams into the compute stream (before the caller's closing event)."""
        torch.cuda.current_stream().wait_event(self._d2h_done)

    def run_pool_bwd(self):
        st = torch.cuda.current_stream()
        arr = (ctypes.c_void_p * len(self.gouts))(*[g.data_ptr() for g in self.gouts])
        L = self._last
        _lib.check(self.plan._lib.dhd_mghs_pool_bwd(
            ctypes.byref(self.plan.cfg), ctypes.c_void_p(L['depth'].data_ptr()),
            ctypes.c_void_p(L['feat'].data_ptr()), ctypes.c_void_p(L['pixmask'].data_ptr()),
            ctypes.c_void_p(self.workspace.data_ptr()), arr, 0,
            ctypes.c_void_p(self.depth_grad.data_ptr()), ctypes.c_void_p(self.feat_grad.data_ptr()),
            ctypes.c_void_p(st.cuda_stream)), 'mghs_pool_bwd')

    def ncu_traffic_bytes(self):
        return pool_traffic_bytes()


class TrainStep(HotPathStep):
    """One data-parallel TRAINING step of the hot path on this rank's B samples (bf16 operands, fp32
    accumulation / master weights / gradients):

      forward   pack -> depth_net -> HeightNet (its mask gates the pool; argmax is not differentiable, so its
                gradient comes from the height loss, lss_heightmap.py:595-622) -> prepare -> fused pool forward
                [BEV / voxel encoders: outside the path -> resident synthetic features]
                SFA (frozen BN) -> predictor -> class-weighted masked cross-entropy (occ_head.py:102-131)
      backward  predictor -> SFA (gradient w.r.t. the encoder features is produced and dropped at the
                boundary) ; [encoders' backward: outside the path -> resident synthetic gradients of the
                four pool outputs] -> fused pool backward -> depth_net backward (dL/d image features)
                height loss -> HeightNet backward (DCN, ASPP, BasicBlocks, SE gate, reduce conv)
      exchange  ONE all-reduce of the flat fp32 gradient bucket (NCCL; launched asynchronously as soon
                as the last gradient is written, joined before the optimizer)
      update    AdamW (DHD-S.py:262: lr 2e-4, weight decay 1e-2) on the fp32 master weights, then the
                bf16 forward / data-gradient weights are re-packed.
    """

    def __init__(self, cfg, B, device='cuda', seed=0, encoders=False, bn='frozen', dropout=None):
        """encoders=True: the real BEV / voxel encoders sit between the pool and the SFA in forward AND backward, so
        the occupancy loss reaches the pool and depth_net through them (no stand-in tensors anywhere)."""
        super().__init__(cfg, B, precision='bf16', device=device, seed=seed, use_graph=False, encoders=encoders)
        from projects.mmdet3d_plugin.models.dense_heads.occ_head import predictor
        from . import shard
        from . import train as T
        from .train import DepthHeadTrainer, HeightNetTrainer, PredictorTrainer, SFATrainer
        # bn='batch': every conv -> BatchNorm2d pair runs torch's training mode (batch statistics, trainable gamma /
        # beta, running statistics updated); the two BatchNorms that see one value per image (HeightNet's BatchNorm1d
        # on the camera vector and the ASPP global-pool branch) stay frozen either way
        self.bn_mode = bn
        if getattr(self.vt, 'sid', False):
            raise NotImplementedError('TrainStep bins gt_depth linearly (dhd_gt_downsample); sid=True is not used by any DHD config')
        T.set_bn_mode(bn)
        torch.manual_seed(seed + 1)
        self.head = predictor(in_dim=256, out_dim=256, Dz=16, num_classes=18, use_predicter=True, class_balance=True,
                              loss_occ=dict(type='CrossEntropyLoss', use_sigmoid=False, ignore_index=255,
                                            loss_weight=1.0)).to(self.device)
        hn = self.vt.height_net
        always_frozen = [hn.bn] + [m for m in hn.modules() if type(m).__name__ == 'ASPP' for m in [m.global_avg_pool[2]]]
        for m in list(self.sfa.modules()) + list(hn.modules()):
            if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)) and (bn == 'frozen' or any(m is f for f in always_frozen)):
                for p in m.parameters():
                    p.requires_grad_(False)
        self.t_depth = DepthHeadTrainer(self.vt.depth_net, self.D, self.device)
        self.t_sfa = SFATrainer(self.sfa, self.device)
        self.t_head = PredictorTrainer(self.head, self.device)
        # nn.Dropout(0.5) behind HeightNet's ASPP (depthnet.py:81): on in the reference's training mode, i.e. together
        # with batch-statistics BatchNorm; off with the frozen (eval-like) BatchNorm unless asked for
        self.dropout = (0.5 if bn == 'batch' else 0.0) if dropout is None else float(dropout)
        self.t_height = HeightNetTrainer(self.vt.height_net, self.device, loss_weight=self.vt.loss_height_weight, dropout=self.dropout,
                                         seed=seed + 17)
        enc_params = []
        if encoders:
            from .train import CustomResNetTrainer, FPNLSSTrainer, UNetTrainer
            enc_mods = [self.bev_backbone, self.bev_neck] + self.voxel
            for m in enc_mods:
                for sub in m.modules():
                    if isinstance(sub, torch.nn.BatchNorm2d) and bn == 'frozen':
                        for p in sub.parameters():
                            p.requires_grad_(False)
                enc_params += [p for p in m.parameters() if p.requires_grad]
            self.t_backbone = CustomResNetTrainer(self.bev_backbone, self.device)
            self.t_neck = FPNLSSTrainer(self.bev_neck, self.device)
            self.t_voxel = [UNetTrainer(u, self.device) for u in self.voxel]
        params = list(self.vt.depth_net.parameters()) + [p for p in self.vt.height_net.parameters() if p.requires_grad] + \
            enc_params + [p for p in self.sfa.parameters() if p.requires_grad] + list(self.head.parameters())
        self.bucket = shard.GradBucket(params)
        # flat-buffer spans of the three backward segments (parameter order above: depth_net, HeightNet | encoders | SFA, head)
        self._spans = {'front': self.bucket.span(list(self.vt.depth_net.parameters()) + list(self.vt.height_net.parameters())),
                       'enc': self.bucket.span(enc_params),
                       'tail': self.bucket.span(list(self.sfa.parameters()) + list(self.head.parameters()))}
        self.grad_bf16 = os.environ.get('DHD_GRAD_BF16', '0') != '0'      # gradients travel as bf16 (half the NCCL bytes)
        self.opt = shard.FlatAdamW(self.bucket, lr=2e-4, weight_decay=1e-2)      # DHD-S.py:262
        self._refresh()                              # the parameters moved into the flat buffer: re-bind every cached view
        gen = torch.Generator(device=self.device).manual_seed(11)
        for g in self.gouts:
            g.mul_(1e-3)
        self._label_seed = seed
        npix = B * self.N * self.fH * self.fW
        self.height_label = torch.empty(npix, dtype=torch.int32, device=self.device)
        self.depth_label = torch.empty(npix, dtype=torch.int32, device=self.device)
        self.height_fg = torch.empty(npix, dtype=torch.uint8, device=self.device)
        self.n_params = self.bucket.flat.numel()
        self.loss = None
        # what a training step hands back to the host: [loss_occ, avg_factor, sem_scal, geo_scal, loss_height]
        self.result = torch.zeros(5, device=self.device)
        self.host_result = torch.zeros(5, pin_memory=True)
        self.d2h_bytes = self.host_result.numel() * 4
        self.grad_clip = 5.0                         # DHD-S.py:263 optimizer_config grad_clip max_norm=5, norm_type=2
        self.lr_schedule = None                      # optional callable(step_index) -> lr (warm-up / step policy hooks)
        self._step_index = 0
        T.set_bn_mode('frozen')

    # ---- inputs: image features + camera geometry + the supervision of one step ------------------
    def make_host_inputs(self, rig, seed):
        """Pinned host buffers of one TRAINING step: what HotPathStep takes plus voxel_semantics / mask_camera
        (B, Dx, Dy, Dz) and the sparse LiDAR maps gt_depth / gt_height (B, N, H_in, W_in) -- the keys the reference's
        Collect3D hands to forward_train (DHD-S.py train_pipeline; ~2 % of the pixels carry a return)."""
        h = super().make_host_inputs(rig, seed)
        g = torch.Generator().manual_seed(1000 + seed)
        B = self.B
        h_in, w_in = self.cfg['input_size']
        hit = torch.rand(B, self.N, h_in, w_in, generator=g) < 0.02
        extra = {
            'voxel_semantics': torch.randint(0, 18, (B, self.Dx, self.Dy, 16), generator=g).to(torch.uint8),
            'mask_camera': (torch.rand(B, self.Dx, self.Dy, 16, generator=g) < 0.5).to(torch.uint8),
            'gt_depth': torch.where(hit, 1.0 + 59.0 * torch.rand(B, self.N, h_in, w_in, generator=g), torch.zeros(())),
            'gt_height': torch.where(hit, -2.0 + 8.0 * torch.rand(B, self.N, h_in, w_in, generator=g), torch.zeros(())),
        }
        h.update({k: v.contiguous().pin_memory() for k, v in extra.items()})
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in h.values())
        return h

    @property
    def labels(self):
        return self.static['voxel_semantics']

    @property
    def mask_camera(self):
        return self.static['mask_camera']

    @property
    def gt_depth(self):
        return self.static['gt_depth']

    @property
    def gt_height(self):
        return self.static['gt_height']

    def run(self, pool_events=None):
        self.train_step(pool_events)

    def stage_names(self):
        mid = ['split(pool outputs -> bf16)', 'CustomResNet + FPN_LSS', '3x UNet'] if self.encoders else \
            ['[encoder stand-in: resident features / resident pool-output gradients]']
        return ['pack', 'depth_net(1x1+softmax)', 'HeightNet(+softmax, BatchNorm batch statistics, Dropout)', 'height_to_mask',
                'mghs_prepare', 'mghs_pool_fwd', 'gt_downsample x2 + height loss'] + mid + \
            ['SFA fwd', 'predictor fwd', 'occupancy loss (CE + sem_scal + geo_scal)', 'predictor bwd', 'SFA bwd'] + \
            (['encoders bwd'] if self.encoders else []) + \
            ['mghs_pool_bwd', 'depth_net bwd', 'HeightNet bwd', 'gradient all-reduce (NCCL)', 'grad clip', 'AdamW', 'weight re-pack']

    def capture_train(self):
        """Capture the step into CUDA graphs: forward up to the binned frustum (graph 1), [the pool kernel stays an
        eager launch so bench.py can bracket it with CUDA events inside the timed region], the rest of the forward +
        losses + backward (graph 2), and the weight re-pack (graph 3).  The step is ~360 library launches plus the
        small torch ops of the gate MLPs, i.e. launch-bound when issued one by one.  The all-reduce, the gradient
        clip and the optimizer stay eager."""
        for _ in range(2):
            self.train_step()
        torch.cuda.synchronize()
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._fwd_bwd()
                self._refresh()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g1, g2, g2b, g2c, g3 = (torch.cuda.CUDAGraph() for _ in range(5))
            with torch.cuda.graph(g1):
                self._fwd_front()
            with torch.cuda.graph(g2):
                self._fwd_bwd_rest()
            if self.encoders:
                with torch.cuda.graph(g2b):
                    self._bwd_encoders()
            else:
                g2b = None
            with torch.cuda.graph(g2c):
                self._bwd_front()
            with torch.cuda.graph(g3):
                self._refresh()
            self.train_graph = (g1, g2, g2b, g2c, g3)
        except Exception as e:  # noqa: BLE001
            self.train_graph = None
            self.train_graph_error = repr(e)[:300]
            torch.cuda.synchronize()
        return self.train_graph is not None

    def _refresh(self):
        ts = [self.t_depth, self.t_height, self.t_sfa, self.t_head]
        if self.encoders:
            ts += [self.t_backbone, self.t_neck] + self.t_voxel
        from . import train as T
        with T.batched_repack():                  # every layer's bf16 operands in a few launches instead of one each
            for t in ts:
                t.refresh()

    def clip_grad_norm(self):
        """optimizer_config = dict(grad_clip=dict(max_norm=5, norm_type=2)) (DHD-S.py:263): one L2 norm of the flat
        gradient bucket, scaled in place when it exceeds max_norm (torch.nn.utils.clip_grad_norm_'s rule, no host sync)."""
        if not self.grad_clip:
            return None
        norm = torch.linalg.vector_norm(self.bucket.flat)
        # the coefficient stays a device scalar and is applied inside the AdamW kernel (no scaling pass over the bucket)
        self._clip_coef = torch.clamp(self.grad_clip / (norm + 1e-6), max=1.0).float()
        return norm

    def train_step(self, pool_events=None):
        g = getattr(self, 'train_graph', None)
        st = torch.cuda.current_stream()
        if g is not None:
            g[0].replay()
        else:
            self._fwd_front()
        if pool_events is not None:
            pool_events[0].record(st)
        self._pool()
        if pool_events is not None:
            pool_events[1].record(st)
        # the gradient exchange is launched per backward segment, as soon as that segment's gradients are final, and
        # runs on NCCL's stream under the segments that follow (torch DDP's bucketed overlap, tools/train.py:195):
        # [head + SFA] -> all-reduce | [encoders] -> all-reduce | [pool, depth_net, HeightNet] -> all-reduce -> join
        if g is not None:
            g[1].replay()
        else:
            self._fwd_bwd_rest()
        lo_tail, hi_tail = self._spans['tail']
        self.bucket.all_reduce_async(lo_tail, hi_tail, bf16=self.grad_bf16)
        if self.encoders:
            if g is not None:
                g[2].replay()
            else:
                self._bwd_encoders()
            self.bucket.all_reduce_async(*self._spans['enc'], bf16=self.grad_bf16)
        if g is not None:
            g[3].replay()
        else:
            self._bwd_front()
        self.bucket.all_reduce_async(*self._spans['front'], bf16=self.grad_bf16)
        self.bucket.wait()
        self.grad_norm = self.clip_grad_norm()
        if self.lr_schedule is not None:
            for pg in self.opt.param_groups:
                pg['lr'] = self.lr_schedule(self._step_index)
        self.opt.step(grad_scale=self._clip_coef if self.grad_clip else None)
        self._step_index += 1
        if g is not None:
            g[4].replay()
        else:
            self._refresh()

    def weight_hash(self):
        """Order-sensitive fingerprint of every trainable parameter (fp64 sum of value x position weight): replicas that
        stepped in lock-step have bit-identical hashes (bench.py all-gathers and compares them)."""
        acc = torch.zeros((), dtype=torch.float64, device=self.device)
        for i, p in enumerate(self.bucket.params):
            v = p.detach().double().flatten()
            w = torch.arange(1, v.numel() + 1, device=self.device, dtype=torch.float64).remainder_(8191.0).add_(1.0)
            acc = acc + (v * w).sum() * (1.0 + 1e-3 * i)
        return acc

    def _fwd_bwd(self):
        self._fwd_front()
        self._pool()
        self._fwd_bwd_rest()
        self._bwd_encoders()
        self._bwd_front()

    def _fwd_front(self):
        s = self.static
        B, N = self.B, self.N
        self.bucket.zero()
        # geometry + binning depend on the camera tensors only: a parallel branch of the captured graph (the ~25 tiny
        # library launches of the reference's torch.inverse / matmul calls and the binning kernels run under the
        # dense front instead of after it)
        main = torch.cuda.current_stream()
        if not hasattr(self, '_prep_stream'):
            self._prep_stream = torch.cuda.Stream()
        self._prep_stream.wait_stream(main)
        with torch.cuda.stream(self._prep_stream):     # the ~25 tiny library launches of the per-camera 3x3s: at once
            cam_mats = self.plan.camera_matrices(s['sensor2ego'], s['cam2imgs'], s['post_rots'], s['post_trans'], s['bda'])

        def fork_prepare():        # the binning kernels late in HeightNet: the bins are still in L2 when the pool starts
            self._prep_stream.wait_stream(main)
            with torch.cuda.stream(self._prep_stream):
                self.plan.prepare(frustum=self.frustum, cam_mats=cam_mats, deterministic=self.deterministic,
                                  workspace=self.workspace)
        self.t_height.trunk_hook = fork_prepare
        # ---- forward
        xa = D.pack_input(s['x'].view(B * N, self.Cin, self.fH, self.fW), 1)
        depth, feat = self.t_depth.forward(xa)
        mlp = self.vt.get_mlp_input(s['sensor2ego'], s['ego2global'], s['cam2imgs'], s['post_rots'],
                                    s['post_trans'], s['bda'])
        height = self.t_height.forward(xa, mlp)
        if self.t_height.trunk_hook is not None:     # a trunk without ASPP never fired it
            self.t_height.trunk_hook = None
            fork_prepare()
        pixmask = height_to_mask(height, self.cfg['height_range'], self.cfg['mask_range'])
        main.wait_stream(self._prep_stream)
        self._last = {'depth': depth, 'feat': feat, 'height': height, 'pixmask': pixmask}

    def _fwd_bwd_rest(self):
        B, N = self.B, self.N
        # get_height_loss (lss_heightmap.py:595-622): labels from the height map, foreground = pixels whose depth
        # return falls into a depth bin (binned with the depth config the module holds at loss time, the LH:455 quirk)
        from . import train as T
        dc = self.vt.mask_3_grid['depth']
        T.gt_downsample(self.gt_depth, self.vt.downsample, dc[0] - dc[2], dc[2], self.vt.D, label=self.depth_label,
                        valid=self.height_fg)
        T.gt_downsample(self.gt_height, self.vt.downsample, self.vt.height_range[0], self.vt.height_interval, self.vt.H,
                        label=self.height_label)
        self.loss_height = self.t_height.loss(self.height_label, self.height_fg)
        if self.encoders:
            if not hasattr(self, '_enc_act'):
                self._enc_act = D.Act.empty(B, self.Dy, self.Dx, 512, 1, self.device)
            enc = self._enc_act
            ins = [D.pack_nhwc(o, 1) for o in self.outs]
            main = torch.cuda.current_stream()
            if not hasattr(self, '_enc_streams'):
                self._enc_streams = [torch.cuda.Stream() for _ in self.t_voxel]
            lo, self._enc_slices = 256, []
            for st, t, x in zip(self._enc_streams, self.t_voxel, ins[1:]):      # independent encoders: parallel branches
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    t.forward(x, out=enc.slice(lo, lo + t.n_classes))
                self._enc_slices.append((lo, lo + t.n_classes))
                lo += t.n_classes
            self.t_neck.forward(self.t_backbone.forward(ins[0]), out=enc.slice(0, 256))
            for st in self._enc_streams:
                main.wait_stream(st)
        else:
            enc = self.encoded_act
        fused = self.t_sfa.forward(enc)
        self.t_head.forward(fused)
        from . import train as T
        T.flush_bn_counters()                        # num_batches_tracked of every batch-statistics BatchNorm of the step
        self.loss = self.t_head.loss(self.labels, self.mask_camera)
        # ---- backward
        dfused = self.t_head.backward()
        self._denc = self.t_sfa.backward(dfused)

    def _bwd_encoders(self):
        """Middle segment of the backward (only with the real encoders): occupancy loss -> encoders -> pool outputs."""
        denc = self._denc
        if self.encoders:
            # occupancy loss -> encoders -> the four pool outputs (fp32 NHWC gradients the pool backward reads)
            main = torch.cuda.current_stream()
            for st, t, (a, b), g in zip(self._enc_streams, self.t_voxel, self._enc_slices, self.gouts[1:]):
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    t.backward(denc.slice(a, b), dx_f32=(g, D.nhwc_strides(g.shape[-1], self.Dy, self.Dx)))
            d0 = self.t_backbone.backward(self.t_neck.backward(denc.slice(0, 256)))
            self.gouts[0].copy_(d0.data.view(self.gouts[0].shape))
            for st in self._enc_streams:
                main.wait_stream(st)

    def _bwd_front(self):
        """Last segment of the backward: pool backward -> depth_net, and the height loss gradient -> HeightNet."""
        self.run_pool_bwd()
        self.t_depth.backward(self.depth_grad, self.feat_grad)
        self.t_height.backward(want_dx=True)
        self.result[:4].copy_(self.loss)
        self.result[4:].copy_(self.loss_height)
