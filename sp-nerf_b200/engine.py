"""Device-side state of one SPNeRF module: packed tensor-core operands, step tables and the
workspace buffers the C ABI needs.  PyTorch is used here only for device memory and streams."""
import ctypes

import torch

from . import _cabi

SLAB_BYTES = 16384
TILE = 128


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class NetEngine:
    """Packed operands for the fused point-network kernels (include/spnerf_b200.h)."""

    def __init__(self, module):
        self.module = module
        self.cfg = _cabi.NetConfig(
            feat=module.feat, layers=module.layers, skip_layer=module.skips[0] if len(module.skips) == 1 else -1,
            mapping=1 if module.uses_mapping else 0, sem=1 if module.sem else 0,
            num_sem_classes=module.num_sem_classes, emb_dim=module.semantic_size if module.sem else 0,
            beta=1 if module.beta else 0, t_dim=module.t_embedding_dims)
        self.sizes = _cabi.NetSizes()
        _cabi.check(_cabi.lib().spnerf_net_sizes(ctypes.byref(self.cfg), ctypes.byref(self.sizes)),
                    "spnerf_net_sizes (only fc_units=512, fc_layers=8, skip 4, encoded input <= 64 are built)")
        self.n_out = self.sizes.n_out
        self.device = None
        self._packed_key = None

    # -- buffers ------------------------------------------------------------------------------
    def _alloc(self, device):
        s = self.sizes
        self.device = device
        self.fwd_blob = torch.empty(max(int(s.fwd_blob_bytes), 16), dtype=torch.uint8, device=device)
        self.bwd_blob = torch.empty(max(int(s.bwd_blob_bytes), 16), dtype=torch.uint8, device=device)
        self.small = torch.empty(int(s.small_floats), dtype=torch.float32, device=device)
        self.fwd_steps = torch.empty(int(s.steps_bytes), dtype=torch.uint8, device=device)
        self.bwd_steps = torch.empty(int(s.steps_bytes), dtype=torch.uint8, device=device)
        self._packed_key = None

    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.module.parameters())

    def ensure_packed(self):
        """(Re)pack the fp32 parameters if any of them changed since the last pack."""
        params = dict(self.module.named_parameters())
        dev = next(iter(params.values())).device
        if dev.type != "cuda":
            raise _cabi.SpnerfError("spnerf_b200 runs on CUDA devices only (no CPU path); move the model to cuda")
        if self.device != dev:
            self._alloc(dev)
        key = self._param_key()
        if key == self._packed_key:
            return
        table = (ctypes.c_void_p * _cabi.NUM_PARAMS)()
        for name, p in params.items():
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise _cabi.SpnerfError(f"parameter {name} must be contiguous fp32")
            table[_cabi.PARAM_SLOTS[name]] = p.data_ptr()
        has_bwd = self.sizes.bwd_blob_bytes > 0
        _cabi.check(_cabi.lib().spnerf_net_pack(
            ctypes.byref(self.cfg), table, _ptr(self.fwd_blob), _ptr(self.bwd_blob) if has_bwd else None,
            _ptr(self.small), _ptr(self.fwd_steps), _ptr(self.bwd_steps) if has_bwd else None, _stream()),
            "spnerf_net_pack")
        self._packed_key = key

    # -- kernels ------------------------------------------------------------------------------
    def save_bytes(self, n_points):
        return ((n_points + TILE - 1) // TILE) * self.sizes.save_slabs_per_tile * SLAB_BYTES

    def sky(self, rays):
        n = rays.shape[0]
        sky = torch.empty(n, 3, dtype=torch.float32, device=rays.device)
        hidden = torch.empty(n, 256, dtype=torch.float32, device=rays.device)
        _cabi.check(_cabi.lib().spnerf_sky_fwd(_ptr(self.small), ctypes.byref(self.cfg), _ptr(rays), n, _ptr(sky),
                                               _ptr(hidden), _stream()), "spnerf_sky_fwd")
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
        a.rays, a.z, a.xyz = rays.data_ptr(), (z.data_ptr() if z is not None else None), \
            (xyz.data_ptr() if xyz is not None else None)
        a.dir_override = dir_override.data_ptr() if dir_override is not None else None
        a.labels = labels.data_ptr() if labels is not None else None
        a.t_emb = t_emb.data_ptr() if t_emb is not None else None
        a.sky = sky.data_ptr()
        a.n_rays, a.n_samples, a.n_steps = n_rays, n_samples, self.sizes.fwd_steps
        a.blob, a.steps, a.small = self.fwd_blob.data_ptr(), self.fwd_steps.data_ptr(), self.small.data_ptr()
        a.out = out.data_ptr()
        a.saves = saves.data_ptr() if saves is not None else None
        a.debug_flags = debug_flags
        _cabi.check(_cabi.lib().spnerf_mlp_fwd(ctypes.byref(a), _stream()), "spnerf_mlp_fwd")
        return out, saves
