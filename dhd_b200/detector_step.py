"""End-to-end step through the plugin DETECTOR (the reference's entry points, DHD_model.py): what the reference's
runner executes per iteration -- `forward_train` -> loss dict -> backward -> gradient all-reduce -> clip -> AdamW -- and
`simple_test` for inference, on synthetic inputs of a DHD config.  Used for BASELINE configs[4] (DHD-L: 6 x 512x1408,
two temporal frames + the stereo reference frame, key frame with gradients, previous frame under no_grad).  Eager
launches through dhd_b200.autograd (no CUDA graph): the module-level path, measured as such."""
import torch

from . import compat as C
from . import shard
from . import synth


class DetectorStep:
    def __init__(self, model_cfg, B, device='cuda', seed=0, lr=2e-4, weight_decay=1e-2, grad_clip=5.0, stereo_channels=None):
        import projects.mmdet3d_plugin  # noqa: F401
        torch.manual_seed(seed)
        self.model = C.DETECTORS.build(model_cfg).to(device)
        self.device, self.B = torch.device(device), B
        vt = self.model.img_view_transformer
        self.N = 6
        self.frames = getattr(self.model, 'num_frame', 1)
        self.stereo = hasattr(self.model, 'extra_ref_frames')
        self.size = tuple(vt.input_size)
        self.fH, self.fW = self.size[0] // vt.downsample, self.size[1] // vt.downsample
        self.Cin = vt.in_channels
        bb = getattr(self.model, 'img_backbone', None)
        self.images = bb is not None and type(bb).__name__ != 'MissingModule' and not self.stereo
        self.Cs = stereo_channels or synth.DHD_L_STEREO_CHANNELS
        self.grad_clip = grad_clip
        self.bucket = shard.GradBucket([p for p in self.model.parameters() if p.requires_grad])
        self.opt = torch.optim.AdamW(self.bucket.params, lr=lr, weight_decay=weight_decay, fused=True)
        self.n_params = self.bucket.flat.numel()

    def make_inputs(self, seed):
        """Device-resident synthetic step inputs: per-frame image features (+ 1/4-resolution stereo features), the camera
        tensors of every frame, voxel labels / camera mask, sparse LiDAR depth / height maps."""
        B, N, nf, dev = self.B, self.N, self.frames, self.device
        g = torch.Generator(device=dev).manual_seed(seed)
        rig = [t.to(dev) for t in synth.synthetic_rig(B, N, self.size, seed=seed)]
        s2e, e2g, K, pr, pt, bda = rig
        rep = lambda t: torch.cat([t] * nf, dim=1)
        if self.images:                                # the detector owns its image backbone: camera images in
            feats = torch.randn(B, N * nf, 3, *self.size, device=dev, generator=g)
        else:
            feats = torch.randn(B, N * nf, self.Cin, self.fH, self.fW, device=dev, generator=g)
        first = feats
        if self.stereo:
            st = torch.randn(B, N * nf, self.Cs, 4 * self.fH, 4 * self.fW, device=dev, generator=g)
            first = (feats, st)
        img_inputs = [first, rep(s2e), rep(e2g), rep(K), rep(pr), rep(pt), bda] if nf > 1 else [first] + rig
        hit = torch.rand(B, N, *self.size, device=dev, generator=g) < 0.02
        zero = torch.zeros((), device=dev)
        kw = dict(voxel_semantics=torch.randint(0, 18, (B, 200, 200, 16), device=dev, generator=g),
                  mask_camera=torch.rand(B, 200, 200, 16, device=dev, generator=g) < 0.5,
                  gt_depth=torch.where(hit, 1.0 + 44.0 * torch.rand(B, N, *self.size, device=dev, generator=g), zero),
                  gt_height=torch.where(hit, -1.0 + 6.4 * torch.rand(B, N, *self.size, device=dev, generator=g), zero))
        return img_inputs, kw

    def train_step(self, img_inputs, kw):
        self.model.train()
        self.bucket.zero()
        losses = self.model.forward_train(img_inputs=img_inputs, img_metas=[{}] * self.B, **kw)
        total = sum(losses.values())
        total.backward()
        self.bucket.all_reduce_async()
        self.bucket.wait()
        if self.grad_clip:
            flat = self.bucket.flat
            flat.mul_(torch.clamp(self.grad_clip / (torch.linalg.vector_norm(flat) + 1e-6), max=1.0))
        self.opt.step()
        return {k: v.detach() for k, v in losses.items()}

    @torch.no_grad()
    def infer_step(self, img_inputs):
        self.model.eval()
        return self.model.simple_test(None, [{}] * self.B, img=img_inputs)

    # ---- the inference step as ONE CUDA graph (the eager detector step is ~500 launches from Python: host-bound in part,
    # more so with 8 processes per host)
    @staticmethod
    def _clone(t):
        return t.clone() if torch.is_tensor(t) else type(t)(DetectorStep._clone(u) for u in t)

    @staticmethod
    def _copy(dst, src):
        if torch.is_tensor(dst):
            dst.copy_(src, non_blocking=True)
        else:
            for d, s in zip(dst, src):
                DetectorStep._copy(d, s)

    @torch.no_grad()
    def capture_infer(self, img_inputs):
        """Capture `simple_test(..., to_host=False)` on static copies of the inputs.  Returns True when the capture
        succeeded (an op that cannot be captured leaves the eager path in place)."""
        self.model.eval()
        self._static_in = [self._clone(t) for t in img_inputs]
        metas = [{}] * self.B
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):                       # engines, plans and workspaces exist before the capture
                self.model.simple_test(None, metas, img=self._static_in, to_host=False)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(graph):
                self._static_occ = self.model.simple_test(None, metas, img=self._static_in, to_host=False)
        except Exception as e:  # noqa: BLE001 -- report and keep the eager path
            self._infer_graph, self.capture_error = None, '%s: %s' % (type(e).__name__, e)
            torch.cuda.synchronize()
            return False
        self._infer_graph = graph
        return True

    @torch.no_grad()
    def infer_step_graphed(self, img_inputs=None):
        """Replay the captured step (new inputs are copied into the static buffers first) -> list of uint8 class maps."""
        if img_inputs is not None:
            self._copy(self._static_in, img_inputs)
        self._infer_graph.replay()
        return list(self._static_occ.cpu().numpy())
