"""Host-side mirror of the reference interface: parameter names / shapes / init stream, argument
errors, the shared-memory slab model, synthetic batches.  CPU only."""
import types

import numpy as np
import pytest
import torch

import spnerf_b200
from oracle import spnerf_oracle as O
from parity_common import build_model, load_case, state_hash
from spnerf_b200 import _cabi, slab, synthetic
from spnerf_b200.models import SPNeRF, load_model
from spnerf_b200.modules import metrics, rendering


@pytest.mark.parametrize("kw", [dict(sem=True, mapping=True), dict(sem=False), dict(sem=True, beta=True, mapping=True)])
def test_state_dict_matches_reference_layout(kw):
    cfg = O.make_cfg(**kw)
    m = load_model(types.SimpleNamespace(**vars(cfg)))
    want = O.parameter_shapes(cfg)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert list(got) == list(want)
    assert got == want
    assert m.number_of_outputs == O.n_outputs(cfg)
    assert m.input_size == O.input_width(cfg)


@pytest.mark.parametrize("name", ["c1_test_sem", "c3_train_guided_mapping_sc", "guided_test_nosem"])
def test_seeded_init_reproduces_reference_weights(name):
    g, meta = load_case(name)
    model, _, _ = build_model(meta, "cpu")
    assert state_hash(model.state_dict()) == meta["state_sha256"]


def test_class_default_width_constructs_but_is_not_built():
    m = SPNeRF()                      # feat=256 (models/spnerf.py:163)
    assert m.fc_net[0].weight.shape == (256, 3)
    with pytest.raises(_cabi.SpnerfError):
        m.engine


def test_no_cpu_fallback():
    args = types.SimpleNamespace(**vars(O.make_cfg(sem=True)))
    model = load_model(args)
    batch = synthetic.make_batch(8)
    with pytest.raises(_cabi.SpnerfError):
        rendering.render_rays({"coarse": model}, args, batch["rays"], None, semantics=batch["sems"])
    with pytest.raises(_cabi.SpnerfError):
        metrics.SemanticLoss()({"sem_logits_coarse": torch.zeros(4, 3)}, torch.zeros(4, dtype=torch.long))


def test_render_rays_argument_errors():
    args = types.SimpleNamespace(**vars(O.make_cfg()))
    args.model = "nerf"
    with pytest.raises(ValueError):                      # modules/rendering.py:179
        rendering.render_rays({}, args, torch.zeros(2, 11), None)
    with pytest.raises(ValueError):                      # models/__init__.py:14-15
        load_model(args)
    with pytest.raises(ValueError):                      # metrics.py:192-193
        metrics.load_loss(args)
    args.model, args.n_importance = "sp-nerf", 64
    with pytest.raises(NotImplementedError):
        rendering.render_rays({"coarse": None}, args, torch.zeros(2, 11), None)


def test_loss_factories():
    a = types.SimpleNamespace(model="sp-nerf", beta=False, sc_lambda=0.1)
    assert isinstance(metrics.load_loss(a), metrics.SNerfLoss)
    a.beta = True
    assert isinstance(metrics.load_loss(a), metrics.SatNerfLoss)
    assert metrics.DepthLoss(lambda_ds=3.0, usealldepth=False).lambda_ds == 1.0       # metrics.py:71
    gn = metrics.DepthLoss(lambda_ds=3.0, GNLL=True, usealldepth=True)                 # constructible, like the reference's
    assert gn.GNLL and gn.lambda_ds == 1.0
    with pytest.raises(TypeError):                       # metrics.py:140: GaussianNLLLoss called without a variance
        gn({"depth_coarse": torch.zeros(4)}, torch.zeros(4))


def test_slab_layout_model():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((128, 64)).astype(np.float16)
    img = slab.pack_slab(x)
    assert np.array_equal(slab.unpack_slab(img, 128), x)
    r, c = 77, 45
    off = (r // 8) * 1024 + (r % 8) * 128 + (((c // 8) ^ (r % 8)) * 16) + (c % 8) * 2
    assert img[off:off + 2].view(np.float16)[0] == x[r, c]
    assert slab.pack_matrix(rng.standard_normal((256, 128)).astype(np.float16)).size == 256 * 128 * 2
    assert slab.idesc_f16(128, 256) == (1 << 4) | (32 << 17) | (8 << 24)


def test_synthetic_batch_shape_and_statistics():
    b = synthetic.make_batch(4096, seed=3)
    assert b["rays"].shape == (4096, 11) and b["rays"].dtype == torch.float32
    assert torch.allclose(b["rays"][:, 3:6].norm(dim=1), torch.ones(4096), atol=1e-5)
    assert float(b["rays"][:, 6].abs().max()) == 0.0 and 0.19 < float(b["rays"][:, 7].mean()) < 0.22
    assert set(b["sems"].unique().tolist()) <= {-100, 0, 1, 2}
    assert 0.6 < float((b["valid_depth"] > 0).float().mean()) < 0.76
    assert float(b["rays"][:, 0:3].abs().max()) <= 1.0
    assert torch.equal(b["depth_std"] > 0, b["valid_depth"] > 0)
