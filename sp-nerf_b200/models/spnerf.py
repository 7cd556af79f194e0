"""SPNeRF point network and volume integrator: host-side mirror of models/spnerf.py.

The ``nn.Module`` keeps the reference's constructor, attributes, parameter names / shapes /
registration order and initialisation stream (so reference checkpoints load and a seeded
construction yields the same weights), but owns no PyTorch compute: ``forward`` and
``inference`` run the hand-written sm_100a kernels behind include/spnerf_b200.h.
"""
import math

import torch
from torch import nn

from .. import render_pass
from ..engine import NetEngine


class PositionalEncoding(nn.Module):
    """Descriptor of the frequency encoding (models/spnerf.py:5-37): 2**k, k < n_freqs, output
    [sin(f x), cos(f x)] per frequency, raw x not included.  Evaluated inside the CUDA kernel."""

    def __init__(self, n_freqs, in_channels):
        super().__init__()
        self.N_freqs = n_freqs
        self.in_channels = in_channels
        self.freq_bands = 2 ** torch.linspace(0, n_freqs - 1, n_freqs)
        self.out_channels = 2 * n_freqs * in_channels


class Sine(nn.Module):
    """Activation marker sin(w0 x) (models/spnerf.py:40-46); fused into the GEMM epilogues."""

    def __init__(self, w0=1.0):
        super().__init__()
        self.w0 = w0

    def extra_repr(self):
        return f"w0={self.w0}"


def _uniform_(linear, bound):
    with torch.no_grad():
        linear.weight.uniform_(-bound, bound)


def _mlp(widths, acts):
    mods = []
    for k in range(len(widths) - 1):
        mods.append(nn.Linear(widths[k], widths[k + 1]))
        if acts[k] is not None:
            mods.append(acts[k])
    return nn.Sequential(*mods)


class SPNeRF(nn.Module):
    """Same signature / attributes as models/spnerf.py:162-271."""

    def __init__(self, num_sem_classes=3, s_embedding_factor=1, layers=8, feat=256, mapping=False,
                 mapping_sizes=[10, 4], skips=[4], siren=True, t_embedding_dims=16, beta=False, sem=False):
        super().__init__()
        self.siren = bool(siren)
        self.layers, self.skips, self.feat = layers, list(skips), feat
        self.t_embedding_dims = t_embedding_dims
        self.input_sizes = [3, 0]
        self.rgb_padding = 0.001
        self.beta, self.sem = beta, sem
        self.num_sem_classes, self.s_embedding_factor = num_sem_classes, s_embedding_factor
        self.semantic_size = num_sem_classes * s_embedding_factor if sem else 0
        self.uses_mapping = bool(mapping)
        if mapping:
            self.mapping = [PositionalEncoding(n, c) for n, c in zip(mapping_sizes, self.input_sizes)]
            xyz_width = self.mapping[0].out_channels
        else:
            self.mapping = [nn.Identity(), nn.Identity()]
            xyz_width = 3
        # creation order below is the reference's, so a seeded construction draws the same stream
        if sem:
            self.semantic_embedding = nn.Embedding(num_sem_classes + 1, self.semantic_size,
                                                   padding_idx=num_sem_classes)
        self.input_size = xyz_width + self.semantic_size
        half = feat // 2
        # `nl` of models/spnerf.py:178: every hidden activation is a sine (w0 = 1; 30 in the first trunk layer) or,
        # with siren=False, a ReLU.  Marker modules only: the kernels apply the activation.
        def act(w0=1.0):
            return Sine(w0) if self.siren else nn.ReLU()
        trunk = []
        for i in range(layers):
            fan_in = self.input_size if i == 0 else feat + (self.input_size if i in self.skips else 0)
            trunk += [nn.Linear(fan_in, feat), act(30.0 if i == 0 else 1.0)]
        self.fc_net = nn.Sequential(*trunk)
        self.sigma_from_xyz = nn.Sequential(nn.Linear(feat, 1), nn.Softplus())
        self.feats_from_xyz = nn.Linear(feat, feat)
        if sem:
            self.logit_from_label = _mlp([feat, half, num_sem_classes], [act(), None])
        self.rgb_from_xyzdir = _mlp([feat, half, 3], [act(), nn.Sigmoid()])
        self.sun_v_net = _mlp([feat + 3, half, half, half, 1], [act(), act(), act(), nn.Sigmoid()])
        self.sky_color = _mlp([3, half, 3], [nn.ReLU(), nn.Sigmoid()])
        # SIREN ranges (models/spnerf.py:49-60, 251-255): all trunk / sun layers U(+-sqrt(6/fan_in)),
        # then the first layer of each U(+-1/fan_in)
        for net in (self.fc_net, self.sun_v_net) if self.siren else ():     # models/spnerf.py:251-255: only with siren
            lins = [m for m in net if isinstance(m, nn.Linear)]
            for m in lins:
                _uniform_(m, math.sqrt(6 / m.weight.size(-1)))
            _uniform_(lins[0], 1 / lins[0].weight.size(-1))
        if beta:
            self.beta_from_xyz = _mlp([t_embedding_dims + feat, half, 1], [act(), nn.Softplus()])
        self.number_of_outputs = 8 + (1 if beta else 0) + (num_sem_classes if sem else 0)
        self._engine = None

    # ------------------------------------------------------------------------------------------
    @property
    def engine(self):
        if self._engine is None:
            self._engine = NetEngine(self)
        return self._engine

    def forward(self, input_xyz, input_dir=None, input_sun_dir=None, input_t=None, input_s=None, sigma_only=False):
        """(P,3) points (+ per-point sun direction, transient embedding, label) -> (P, number_of_outputs)
        with columns [albedo(3), sigma, sun, sky(3), (beta), (logits)]  (models/spnerf.py:273-369).
        Differentiable w.r.t. the parameters (and input_t)."""
        if input_sun_dir is None:
            raise ValueError("input_sun_dir is required (models/spnerf.py:351)")
        out = render_pass.point_rows(self, input_xyz, input_sun_dir, input_t, input_s)
        return out[:, 3:4] if sigma_only else out


def inference(model, args, rays_xyz, z_vals, rays_d=None, sun_d=None, rays_t=None, semantics=None,
              z_vals_unsort=None):
    """Volume integration of one pass (models/spnerf.py:63-159): network on every sample, alpha
    compositing, shadow-aware shading, depth, mean semantic logits.  Same keys as the reference."""
    return render_pass.integrate(model, args, z_vals, xyz=rays_xyz, sun_d=sun_d, rays_t=rays_t, semantics=semantics,
                                 z_vals_unsort=z_vals_unsort)
