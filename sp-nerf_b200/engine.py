"""Device-side state of one SPNeRF module: packed tensor-core operands, step tables and the
workspace buffers the C ABI needs, plus thin wrappers over every kernel entry point.
PyTorch is used here only for device memory and streams."""
import ctypes
import math

import torch

from . import _cabi

SLAB_BYTES = 16384
TILE = 128


def _p(t):
    return t.data_ptr() if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t, what):
    if t.device.type != "cuda":
        raise _cabi.SpnerfError(f"{what} must live on a CUDA device: spnerf_b200 has no CPU path")
    if t.device.index is not None and t.device.index != torch.cuda.current_device():
        # kernels are launched on the current device's stream: a tensor of another GPU would be dereferenced there
        raise _cabi.SpnerfError(f"{what} lives on {t.device} but the current CUDA device is "
                                f"cuda:{torch.cuda.current_device()}: call torch.cuda.set_device first "
                                "(one process per GPU, or a torch.cuda.device(...) block around the call)")


_tables = {}


def sampler_tables(n, device):
    """torch.linspace(0,1,n) and the Gaussian bin weights of modules/rendering.py:60,68-70, computed by
    torch on the host exactly as the reference does, so the device kernels reproduce its bits."""
    key = (n, str(device))
    if key not in _tables:
        t = torch.linspace(0, 1, n)
        x = torch.linspace(-3., 3., steps=(n - 1))
        g = (1. / math.sqrt(2 * math.pi) * torch.exp(-0.5 * x.pow(2)))
        _tables[key] = (t.to(device).contiguous(), g.to(device).contiguous())
    return _tables[key]


# ------------------------------------------------------------------------------------------------
# model-independent kernels
# ------------------------------------------------------------------------------------------------
def sample_coarse(rays, uniforms, n):
    """modules/rendering.py:128-144."""
    _require_cuda(rays, "rays")
    t_tab, _ = sampler_tables(n, rays.device)
    z = torch.empty(rays.shape[0], n, dtype=torch.float32, device=rays.device)
    _cabi.check(_cabi.lib().spnerf_sample_coarse(_p(rays), _p(t_tab), _p(uniforms), rays.shape[0], n, _p(z), _stream()),
                "spnerf_sample_coarse")
    return z


_rng_state = {}
_rng_seed = None


def manual_seed(seed):
    """Restart the sampler's device-side random stream (all devices) from `seed`: the same seed replays the same
    sample depths.  Without it the stream is seeded from torch's default generator on first use."""
    global _rng_seed
    _rng_seed = int(seed)
    for st in _rng_state.values():      # in place: a captured CUDA graph keeps reading the same words
        st.copy_(torch.tensor([_rng_seed, 0, 0], dtype=torch.int64))


def rng_state(device):
    """Device-resident {seed, step, 0} of the sampler's Philox stream (one per device)."""
    key = str(device)
    if key not in _rng_state:
        seed = _rng_seed if _rng_seed is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
        _rng_state[key] = torch.tensor([seed, 0, 0], dtype=torch.int64, device=device)
    return _rng_state[key]


def sample_coarse_rng(rays, n, state=None):
    """modules/rendering.py:128-144 with the uniforms of :143 drawn inside the kernel (no torch.rand launch)."""
    _require_cuda(rays, "rays")
    t_tab, _ = sampler_tables(n, rays.device)
    state = rng_state(rays.device) if state is None else state
    z = torch.empty(rays.shape[0], n, dtype=torch.float32, device=rays.device)
    _cabi.check(_cabi.lib().spnerf_sample_coarse_rng(_p(rays), _p(t_tab), _p(state), rays.shape[0], n, _p(z), _stream()),
                "spnerf_sample_coarse_rng")
    return z


def sample_guided(rays, z, weights, depth, u_pred, valid_depth=None, target_depths=None, target_std=None, u_gt=None,
                  want_indices=False):
    """modules/rendering.py:92-116 + :165-167.  Returns (z_unsort (B,2N), z_sorted (B,2N)[, indices])."""
    b, n = z.shape
    t_tab, g_tab = sampler_tables(n, rays.device)
    z_unsort = torch.empty(b, 2 * n, dtype=torch.float32, device=rays.device)
    z_sorted = torch.empty(b, 2 * n, dtype=torch.float32, device=rays.device)
    inds = torch.empty(b, n, dtype=torch.int32, device=rays.device) if want_indices else None
    a = _cabi.Guided()
    a.rays, a.z, a.weights, a.depth = _p(rays), _p(z), _p(weights), _p(depth)
    a.valid_depth = _p(valid_depth)
    if valid_depth is not None:
        if target_depths.dim() == 2:
            if target_depths.stride(1) != 1:
                target_depths = target_depths.contiguous()
            a.target_depth, a.target_depth_stride = _p(target_depths), target_depths.stride(0)
        else:
            a.target_depth, a.target_depth_stride = _p(target_depths), target_depths.stride(0)
        a.target_std, a.u_gt = _p(target_std), _p(u_gt)
    a.u_pred, a.t_table, a.gauss_table = _p(u_pred), _p(t_tab), _p(g_tab)
    a.n_rays, a.n_samples = b, n
    a.z_unsort, a.z_sorted, a.searchsorted_out = _p(z_unsort), _p(z_sorted), _p(inds)
    _cabi.check(_cabi.lib().spnerf_sample_guided(ctypes.byref(a), _stream()), "spnerf_sample_guided")
    return (z_unsort, z_sorted, inds) if want_indices else (z_unsort, z_sorted)


def composite_fwd(out, z, n_out, col_sem, n_sem, noise=None, noise_std=0.0, want_raw=True, want_samples=True,
                  want_aux=False, col_beta=-1):
    """models/spnerf.py:109-157.  want_samples=False skips the (B,N) weights / transparency outputs and
    want_aux=True adds the per-ray sums of eval.py:75-101 and the class argmax (image export)."""
    b, n = z.shape
    dev = z.device
    f32 = dict(dtype=torch.float32, device=dev)
    weights = torch.empty(b, n, **f32) if want_samples else None
    trans = torch.empty(b, n, **f32) if want_samples else None
    rgb, depth = torch.empty(b, 3, **f32), torch.empty(b, **f32)
    rgb_raw = torch.empty(b, 3, **f32) if want_raw else None
    sem = torch.empty(b, n_sem, **f32) if n_sem > 0 else None
    aux = torch.empty(b, 8, **f32) if want_aux else None
    amax = torch.empty(b, dtype=torch.int32, device=dev) if (want_aux and n_sem > 0) else None
    a = _cabi.CompositeFwd()
    a.out, a.z, a.noise = _p(out), _p(z), _p(noise)
    a.n_rays, a.n_samples, a.n_out, a.col_sem, a.n_sem, a.noise_std = b, n, n_out, max(col_sem, 0), n_sem, noise_std
    a.weights, a.transparency, a.rgb, a.rgb_raw, a.depth, a.sem_logits = \
        _p(weights), _p(trans), _p(rgb), _p(rgb_raw), _p(depth), _p(sem)
    a.ray_aux, a.sem_argmax, a.col_beta = _p(aux), _p(amax), col_beta
    _cabi.check(_cabi.lib().spnerf_composite_fwd(ctypes.byref(a), _stream()), "spnerf_composite_fwd")
    if want_aux:
        return weights, trans, rgb, rgb_raw, depth, sem, aux, amax
    return weights, trans, rgb, rgb_raw, depth, sem


def composite_bwd(out, z, weights, trans, rgb_raw, n_out, col_sem, n_sem, g_rgb=None, g_depth=None, g_sem=None,
                  g_w=None, g_t=None, g_out_ext=None, noise=None, noise_std=0.0, absmax=None):
    """Adjoint of composite_fwd.  Returns (g_out, g_sky_ray, absmax scalar).
    `absmax`: a zeroed (1,) accumulator to use instead of a fresh one (NetEngine.absmax is reset by the backward)."""
    b, n = z.shape
    dev = z.device
    g_out = torch.empty(b * n, n_out, dtype=torch.float32, device=dev)
    g_sky = torch.empty(b, 3, dtype=torch.float32, device=dev)
    if absmax is None:
        absmax = torch.zeros(1, dtype=torch.float32, device=dev)
    a = _cabi.CompositeBwd()
    a.out, a.z, a.noise, a.weights, a.transparency, a.rgb_raw = _p(out), _p(z), _p(noise), _p(weights), _p(trans), \
        _p(rgb_raw)
    a.g_rgb, a.g_depth, a.g_sem_logits, a.g_weights, a.g_transparency, a.g_out_ext = \
        _p(g_rgb), _p(g_depth), _p(g_sem), _p(g_w), _p(g_t), _p(g_out_ext)
    a.n_rays, a.n_samples, a.n_out, a.col_sem, a.n_sem, a.noise_std = b, n, n_out, max(col_sem, 0), n_sem, noise_std
    a.g_out, a.g_sky_ray, a.g_absmax = _p(g_out), _p(g_sky), _p(absmax)
    _cabi.check(_cabi.lib().spnerf_composite_bwd(ctypes.byref(a), _stream()), "spnerf_composite_bwd")
    return g_out, g_sky, absmax


_loss_ws = {}


def losses(n_rays, rgb=None, rgb_target=None, depth=None, z=None, weights=None, target_depth=None,
           target_weight=None, target_std=None, valid_depth=None, lambda_ds=0.0, use_all_depth=False,
           sem_logits=None, labels=None, lambda_ss=0.0, gnll=False):
    """Fused loss reductions + gradients (modules/metrics.py:27-45, 68-159, 162-183).
    Returns (scalars (8,), g_rgb, g_depth, g_sem_logits, g_weights); g_weights only for the GNLL depth variant."""
    ref = rgb if rgb is not None else (depth if depth is not None else sem_logits)
    dev = ref.device
    _require_cuda(ref, "loss inputs")
    f32 = dict(dtype=torch.float32, device=dev)
    out = torch.empty(8, **f32)
    g_rgb = torch.empty_like(rgb) if rgb is not None else None
    g_depth = torch.empty_like(depth) if depth is not None else None
    g_sem = torch.empty_like(sem_logits) if sem_logits is not None else None
    a = _cabi.Losses()
    a.n_rays = n_rays
    a.n_samples = z.shape[1] if z is not None else 0
    a.n_sem = sem_logits.shape[1] if sem_logits is not None else 0
    a.rgb, a.rgb_target, a.g_rgb = _p(rgb), _p(rgb_target), _p(g_rgb)
    a.depth, a.z, a.weights = _p(depth), _p(z), _p(weights)
    stride = 1
    if target_depth is not None and target_depth.dim() == 1 and target_depth.stride(0) != 1:
        if target_weight.stride(0) == target_depth.stride(0) and target_depth.stride(0) > 0:
            stride = target_depth.stride(0)          # the two columns of the reference's (B,2) `depths`, read in place
        else:
            target_depth, target_weight = target_depth.contiguous(), target_weight.contiguous()
    elif target_weight is not None and target_weight.dim() == 1 and target_weight.stride(0) != 1:
        target_weight = target_weight.contiguous()
    a.target_stride = stride
    a.target_depth, a.target_weight, a.target_std, a.valid_depth = \
        _p(target_depth), _p(target_weight), _p(target_std), _p(valid_depth)
    a.lambda_ds, a.use_all_depth, a.g_depth = lambda_ds, 1 if use_all_depth else 0, _p(g_depth)
    a.sem_logits, a.labels, a.lambda_ss, a.g_sem_logits = _p(sem_logits), _p(labels), lambda_ss, _p(g_sem)
    a.losses, a.workspace = _p(out), _p(_loss_workspace(dev))
    g_w = torch.empty_like(weights) if (gnll and depth is not None) else None
    a.gnll, a.g_weights = 1 if gnll else 0, _p(g_w)
    _cabi.check(_cabi.lib().spnerf_losses(ctypes.byref(a), _stream()), "spnerf_losses")
    return out, g_rgb, g_depth, g_sem, g_w


def _loss_workspace(dev):
    if dev not in _loss_ws:
        _loss_ws[dev] = torch.zeros(int(_cabi.lib().spnerf_losses_workspace_bytes()), dtype=torch.uint8, device=dev)
    return _loss_ws[dev]


def _row_strided(t, n_rays, n_samples):
    """(pointer tensor, element stride) of a (B,N[,1]) tensor whose element (r,i) sits at base + (r N + i) stride:
    a column view of the network's output rows is read in place, anything else is made contiguous."""
    t = t.detach()
    if t.dim() == 3:
        t = t[..., 0]
    if t.dtype == torch.float32 and t.shape == (n_rays, n_samples) and t.stride(1) >= 1 and (
            n_rays == 1 or t.stride(0) == n_samples * t.stride(1)):
        return t, t.stride(1)
    return t.float().contiguous(), 1


def loss_solar(trans_sc, weights_sc, sun_sc, lambda_sc, upstream=None, backward=False):
    """modules/metrics.py:17-24.  Forward: (2,) [sc_term2, sc_term3]; backward: d/d sun_sc (B,N) for the
    upstream gradients of the two scalars ((2,) device tensor, None = ones)."""
    b, n = trans_sc.shape
    _require_cuda(trans_sc, "transparency_sc")
    dev = trans_sc.device
    sun, stride = _row_strided(sun_sc, b, n)
    a = _cabi.LossSolar()
    a.n_rays, a.n_samples, a.lambda_sc = b, n, float(lambda_sc)
    a.transparency_sc, a.weights_sc, a.sun_sc, a.sun_stride = _p(trans_sc), _p(weights_sc), _p(sun), stride
    a.upstream = _p(upstream)
    if backward:
        res = torch.empty(b, n, dtype=torch.float32, device=dev)
        a.g_sun = _p(res)
    else:
        res = torch.empty(2, dtype=torch.float32, device=dev)
        a.losses, a.workspace = _p(res), _p(_loss_workspace(dev))
    _cabi.check(_cabi.lib().spnerf_loss_solar(ctypes.byref(a), 1 if backward else 0, _stream()), "spnerf_loss_solar")
    return res


def loss_uncertainty(rgb, rgb_target, weights, beta, beta_min=0.05, beta_ray=None, upstream=None, backward=False):
    """modules/metrics.py:10-14.  Forward: ((2,) [color, logbeta], beta_ray (B,)); backward: (g_rgb, g_weights,
    g_beta) for the upstream gradients of the two scalars."""
    b, n = weights.shape
    _require_cuda(rgb, "rgb")
    dev = rgb.device
    bt, stride = _row_strided(beta, b, n)
    a = _cabi.LossUncertainty()
    a.n_rays, a.n_samples, a.beta_min = b, n, float(beta_min)
    a.rgb, a.rgb_target, a.weights, a.beta, a.beta_stride = _p(rgb), _p(rgb_target), _p(weights), _p(bt), stride
    a.upstream = _p(upstream)
    f32 = dict(dtype=torch.float32, device=dev)
    if backward:
        g_rgb, g_w, g_b = torch.empty(b, 3, **f32), torch.empty(b, n, **f32), torch.empty(b, n, **f32)
        a.beta_ray, a.g_rgb, a.g_weights, a.g_beta = _p(beta_ray), _p(g_rgb), _p(g_w), _p(g_b)
        res = (g_rgb, g_w, g_b)
    else:
        vals, beta_ray = torch.empty(2, **f32), torch.empty(b, **f32)
        a.beta_ray, a.losses, a.workspace = _p(beta_ray), _p(vals), _p(_loss_workspace(dev))
        res = (vals, beta_ray)
    _cabi.check(_cabi.lib().spnerf_loss_uncertainty(ctypes.byref(a), 1 if backward else 0, _stream()),
                "spnerf_loss_uncertainty")
    return res


import os as _os
_ENV_DEBUG = int(_os.environ.get("SPNERF_DEBUG_FLAGS", "0"))     # kernel timing experiments only (include/spnerf_b200.h)


def adam_step(flat_params, flat_grads, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam(weight_decay=0) on flat fp32 buffers in one launch (main.py:96-97)."""
    _require_cuda(flat_params, "parameters")
    n = flat_params.numel()
    for t in (flat_grads, exp_avg, exp_avg_sq):
        if t.numel() != n or t.dtype != torch.float32 or not t.is_contiguous():
            raise _cabi.SpnerfError("adam_step: buffers must be contiguous fp32 of the parameters' size")
    _cabi.check(_cabi.lib().spnerf_adam_step(_p(flat_params), _p(flat_grads), _p(exp_avg), _p(exp_avg_sq), n, int(step),
                                             float(lr), float(betas[0]), float(betas[1]), float(eps), _stream()),
                "spnerf_adam_step")


# ------------------------------------------------------------------------------------------------
# per-model state
# ------------------------------------------------------------------------------------------------
class NetEngine:
    """Packed operands and workspaces for the fused point-network kernels (include/spnerf_b200.h)."""

    def __init__(self, module):
        self.module = module
        self.cfg = _cabi.NetConfig(
            feat=module.feat, layers=module.layers, skip_layer=module.skips[0] if len(module.skips) == 1 else -1,
            mapping=1 if module.uses_mapping else 0, sem=1 if module.sem else 0,
            num_sem_classes=module.num_sem_classes, emb_dim=module.semantic_size if module.sem else 0,
            beta=1 if module.beta else 0, t_dim=module.t_embedding_dims, relu=0 if getattr(module, "siren", True) else 1)
        self.sizes = _cabi.NetSizes()
        _cabi.check(_cabi.lib().spnerf_net_sizes(ctypes.byref(self.cfg), ctypes.byref(self.sizes)),
                    "spnerf_net_sizes (built: fc_units 512 or 256, fc_layers 8, skip 4, <= 8 classes, encoded input up to 64 "
                    "columns plus what fits the free aux columns: --mapping with <= 7 classes, <= 6 with --beta)")
        self.n_out = self.sizes.n_out
        self.col_sem = 8 + (1 if module.beta else 0) if module.sem else -1
        self.n_sem = module.num_sem_classes if module.sem else 0
        self.col_beta = 8 if module.beta else -1
        self.device = None
        self._packed_key = None
        self._grad_key = None
        self.names = [n for n, _ in module.named_parameters()]

    # -- buffers ------------------------------------------------------------------------------
    def _alloc(self, device):
        s = self.sizes
        self.device = device
        u8 = dict(dtype=torch.uint8, device=device)
        self.fwd_blob = torch.empty(max(int(s.fwd_blob_bytes), 16), **u8)
        self.bwd_blob = torch.empty(max(int(s.bwd_blob_bytes), 16), **u8)
        self.small = torch.empty(int(s.small_floats), dtype=torch.float32, device=device)
        self.fwd_steps = torch.empty(int(s.steps_bytes), **u8)
        self.bwd_steps = torch.empty(int(s.steps_bytes), **u8)
        self.wgrad_ws = torch.empty(int(_cabi.lib().spnerf_mlp_wgrad_workspace_bytes(ctypes.byref(self.cfg))), **u8)
        self.pack_ws = torch.empty(int(_cabi.lib().spnerf_net_pack_workspace_bytes(ctypes.byref(self.cfg))), **u8)
        # self-cleaning scratch: slots summed with atomics (flushed and cleared by the weight-gradient reduce kernel)
        # and the running max |dL/d out| of the compositing adjoint (reset by the same kernel)
        self.accum = torch.zeros(_cabi.ACCUM_FLOATS, dtype=torch.float32, device=device)
        self.absmax = torch.zeros(1, dtype=torch.float32, device=device)
        self._packed_key = None
        self._prepared_key = None
        self._grad_key = None
        self.n_packs = 0

    def ensure_packed(self, force=False):
        """(Re)pack the fp32 parameters if any of them changed since the last pack."""
        params = dict(self.module.named_parameters())
        dev = next(iter(params.values())).device
        if dev.type != "cuda":
            raise _cabi.SpnerfError("spnerf_b200 runs on CUDA devices only (no CPU path); move the model to cuda")
        if self.device != dev:
            self._alloc(dev)
        key = tuple((p.data_ptr(), p._version) for p in params.values())
        if key == self._packed_key and not force:
            return
        ptrs = tuple(p.data_ptr() for p in params.values())
        if ptrs != self._prepared_key:          # tables embed the parameter addresses
            table = (ctypes.c_void_p * _cabi.NUM_PARAMS)()
            for name, p in params.items():
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise _cabi.SpnerfError(f"parameter {name} must be contiguous fp32")
                table[_cabi.PARAM_SLOTS[name]] = p.data_ptr()
            _cabi.check(_cabi.lib().spnerf_net_prepare(
                ctypes.byref(self.cfg), table, _p(self.pack_ws), self.pack_ws.numel(), _p(self.fwd_blob),
                _p(self.bwd_blob), _p(self.small), _p(self.fwd_steps), _p(self.bwd_steps), _stream()),
                "spnerf_net_prepare")
            self._prepared_key = ptrs
        _cabi.check(_cabi.lib().spnerf_net_pack(ctypes.byref(self.cfg), _p(self.pack_ws), _p(self.fwd_blob),
                                                _p(self.bwd_blob), _p(self.small), _stream()), "spnerf_net_pack")
        self._packed_key = key
        self.n_packs += 1

    def mark_dirty(self):
        """The parameters were changed behind autograd's back (fused optimiser step): repack on next use."""
        self._packed_key = None

    def save_bytes(self, n_points):
        return ((n_points + TILE - 1) // TILE) * self.sizes.save_slabs_per_tile * SLAB_BYTES

    def grad_save_bytes(self, n_points):
        return ((n_points + TILE - 1) // TILE) * self.sizes.grad_slabs_per_tile * SLAB_BYTES

    # -- forward ------------------------------------------------------------------------------
    def sky(self, rays):
        n = rays.shape[0]
        sky = torch.empty(n, 3, dtype=torch.float32, device=rays.device)
        hidden = torch.empty(n, self.cfg.feat // 2, dtype=torch.float32, device=rays.device)
        _cabi.check(_cabi.lib().spnerf_sky_fwd(_p(self.small), ctypes.byref(self.cfg), _p(rays), n, _p(sky),
                                               _p(hidden), _stream()), "spnerf_sky_fwd")
        return sky, hidden

    def forward(self, rays, n_samples, z=None, xyz=None, dir_override=None, labels=None, t_emb=None, sky=None,
                save=False, debug_flags=0):
        """Network output rows (n_rays*n_samples, n_out) fp32 in the reference's column order."""
        self.ensure_packed()
        n_rays = rays.shape[0]
        n_points = n_rays * n_samples
        if sky is None:
            sky, _ = self.sky(rays)
        out = torch.empty(n_points, self.n_out, dtype=torch.float32, device=rays.device)
        saves = torch.empty(self.save_bytes(n_points), dtype=torch.uint8, device=rays.device) if save else None
        a = _cabi.MlpFwd()
        a.cfg = self.cfg
        a.rays, a.z, a.xyz, a.dir_override = _p(rays), _p(z), _p(xyz), _p(dir_override)
        a.labels, a.t_emb, a.sky = _p(labels), _p(t_emb), _p(sky)
        a.n_rays, a.n_samples, a.n_steps = n_rays, n_samples, self.sizes.fwd_steps
        a.blob, a.steps, a.small = _p(self.fwd_blob), _p(self.fwd_steps), _p(self.small)
        a.out, a.saves, a.debug_flags = _p(out), _p(saves), debug_flags | _ENV_DEBUG
        _cabi.check(_cabi.lib().spnerf_mlp_fwd(ctypes.byref(a), _stream()), "spnerf_mlp_fwd")
        return out, saves

    # -- backward -----------------------------------------------------------------------------
    def _views(self, flat):
        views, off = [], 0
        for p in self.module.parameters():
            views.append(flat[off:off + p.numel()].view(p.shape))
            off += p.numel()
        return views

    def _grad_buffer(self):
        """Persistent flat fp32 buffer shaped like the parameters: the kernels write into it (its
        address is baked into the weight-gradient scatter tables); callers get a copy."""
        n = sum(p.numel() for p in self.module.parameters())
        if getattr(self, "_gflat", None) is None or self._gflat.device != self.device or self._gflat.numel() != n:
            self._gflat = torch.empty(n, dtype=torch.float32, device=self.device)
            self._grad_key = None
        return self._gflat

    def backward(self, g_out, out, rays, n_samples, saves, absmax, labels=None, t_emb=None, g_sky_ray=None,
                 sky=None, sky_hidden=None, debug_flags=0, timer=None, copy=True):
        """All parameter gradients (+ d t_emb) from dL/d out.  Returns (flat, views, g_t_emb).
        Launches: backward-data, sky backward, weight-gradient GEMMs, reduce (+ flush of the atomically summed slots).
        copy=True returns a private copy (autograd hands the views to the caller for good); copy=False returns the
        engine's persistent buffer."""
        n_rays = rays.shape[0]
        n_points = n_rays * n_samples
        dev = rays.device
        work = self._grad_buffer()            # every slot is overwritten below: no memset
        by_name = dict(zip(self.names, self._views(work)))
        gsaves = torch.empty(self.grad_save_bytes(n_points), dtype=torch.uint8, device=dev)
        scale = torch.empty(1, dtype=torch.float32, device=dev)
        acc = self.accum.data_ptr()
        g_t = torch.zeros(n_rays, self.cfg.t_dim, dtype=torch.float32, device=dev) if t_emb is not None else None
        a = _cabi.MlpBwd()
        a.cfg = self.cfg
        a.g_out, a.out, a.rays, a.labels, a.t_emb = _p(g_out), _p(out), _p(rays), _p(labels), _p(t_emb)
        a.n_rays, a.n_samples, a.n_steps = n_rays, n_samples, self.sizes.bwd_steps
        a.blob, a.steps, a.small = _p(self.bwd_blob), _p(self.bwd_steps), _p(self.small)
        a.saves, a.grad_saves, a.g_absmax, a.scale_out = _p(saves), _p(gsaves), _p(absmax), _p(scale)
        a.g_emb = acc + 4 * _cabi.ACC_EMB if self.module.sem else None
        a.g_small_bias, a.g_t_emb, a.debug_flags = acc + 4 * _cabi.ACC_SMALL_BIAS, _p(g_t), debug_flags | _ENV_DEBUG
        _cabi.check(_cabi.lib().spnerf_mlp_bwd_data(ctypes.byref(a), _stream()), "spnerf_mlp_bwd_data")
        if timer is not None:
            timer.mark("mlp_bwd_data")
        if g_sky_ray is not None:             # before the weight-gradient launch: its reduce kernel flushes the scratch
            _cabi.check(_cabi.lib().spnerf_sky_bwd(
                _p(self.small), ctypes.byref(self.cfg), _p(rays), _p(sky), _p(sky_hidden), _p(g_sky_ray), n_rays,
                acc + 4 * _cabi.ACC_SKY_W0, acc + 4 * _cabi.ACC_SKY_B0, acc + 4 * _cabi.ACC_SKY_W2,
                acc + 4 * _cabi.ACC_SKY_B2, _stream()), "spnerf_sky_bwd")

        w = _cabi.MlpWgrad()
        w.cfg = self.cfg
        w.n_points, w.saves, w.grad_saves, w.scale = n_points, _p(saves), _p(gsaves), _p(scale)
        table = (ctypes.c_void_p * _cabi.NUM_PARAMS)()
        for name, v in by_name.items():
            table[_cabi.PARAM_SLOTS[name]] = v.data_ptr()
        w.grads_host = table
        w.workspace, w.workspace_bytes = _p(self.wgrad_ws), self.wgrad_ws.numel()
        w.accum, w.absmax_reset = acc, _p(self.absmax)
        # the scatter tables embed the gradient pointers: uploaded once per gradient buffer
        key = (work.data_ptr(), self.wgrad_ws.data_ptr(), acc)
        if key != self._grad_key:
            _cabi.check(_cabi.lib().spnerf_mlp_wgrad_prepare(ctypes.byref(w), _stream()), "spnerf_mlp_wgrad_prepare")
            self._grad_key = key
        _cabi.check(_cabi.lib().spnerf_mlp_bwd_weights(ctypes.byref(w), _stream()), "spnerf_mlp_bwd_weights")
        if timer is not None:
            timer.mark("mlp_bwd_weights")
        if not copy:                          # fused step: the persistent buffer itself (valid until the next backward)
            return work, self._views(work), g_t
        flat = work.clone()
        if timer is not None:
            timer.mark("grad_tail")
        return flat, self._views(flat), g_t
