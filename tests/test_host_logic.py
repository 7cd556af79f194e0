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


def test_class_default_width_is_built_and_odd_widths_are_refused():
    m = SPNeRF()                      # feat=256 (models/spnerf.py:163)
    assert m.fc_net[0].weight.shape == (256, 3)
    assert m.engine.sizes.n_out == 8 and m.engine.cfg.feat == 256
    with pytest.raises(_cabi.SpnerfError):
        SPNeRF(feat=384).engine


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


def test_trainer_checkpoint_round_trip_with_the_reference_layout(tmp_path):
    """A Lightning checkpoint of the reference's NeRF_pl (state_dict keys 'nerf_coarse.<param>' / 'embedding_t.weight',
    main.py:48-57) loads into the trainer's flat-buffer views; the trainer's own checkpoints use the same layout."""
    import torch
    from spnerf_b200 import config
    from spnerf_b200.models import load_model
    from spnerf_b200.trainer import Trainer
    args = config.make_args(sem=True, num_sem_classes=3, fc_units=512, beta=True, t_embbeding_tau=4, lr=5e-4)
    torch.manual_seed(1)
    tr = Trainer(args, "cpu")
    torch.manual_seed(2)                                  # a "reference run" with other weights
    other = load_model(args)
    t_other = torch.nn.Embedding(30, 4)
    ref_ckpt = {"state_dict": {**{"nerf_coarse." + k: v.clone() for k, v in other.state_dict().items()},
                               "embedding_t.weight": t_other.weight.detach().clone()},
                "global_step": 1234, "pytorch-lightning_version": "1.3.7"}
    path = tmp_path / "epoch=27.ckpt"
    torch.save(ref_ckpt, path)
    flat_ptr = tr.flat.data_ptr()
    tr.load_checkpoint(str(path))
    for (k, v), p in zip(other.state_dict().items(), tr.models["coarse"].parameters()):
        assert torch.equal(p.detach(), v), k
        assert flat_ptr <= p.data_ptr() < flat_ptr + tr.flat.numel() * 4      # still a view of the flat buffer
    assert torch.equal(tr.models["t"].weight.detach(), t_other.weight.detach()) and tr.train_steps == 1234
    # our own checkpoint: same keys as the reference's, optimiser state restored
    tr.exp_avg.fill_(0.25)
    tr.opt_steps = 7
    mine = tmp_path / "mine.ckpt"
    tr.save_checkpoint(str(mine))
    saved = torch.load(mine, weights_only=False)
    assert sorted(saved["state_dict"]) == sorted(ref_ckpt["state_dict"])
    torch.manual_seed(3)
    tr2 = Trainer(args, "cpu")
    tr2.load_checkpoint(str(mine))
    assert torch.equal(tr2.flat, tr.flat) and float(tr2.exp_avg[0]) == 0.25 and tr2.opt_steps == 7
    bad = {"state_dict": {k: v for k, v in ref_ckpt["state_dict"].items() if "sigma" not in k}}
    import pytest
    with pytest.raises(RuntimeError):
        tr2.load_checkpoint(bad)


def test_epoch_batches_shuffle_covers_every_ray_once():
    import torch
    from spnerf_b200 import config, synthetic
    from spnerf_b200.trainer import Trainer
    args = config.make_args(sem=True, num_sem_classes=3, fc_units=512, batch_size=100)
    tr = Trainer(args, "cpu")
    pool = synthetic.make_batch(1050, seed=3)
    pool["index"] = torch.arange(1050)
    seen = torch.cat([b["index"] for b in tr.epoch_batches(pool, generator=torch.Generator().manual_seed(0))])
    assert seen.numel() == 1050 and torch.equal(torch.sort(seen).values, torch.arange(1050))
    assert not torch.equal(seen, torch.arange(1050))
    b0 = next(iter(tr.epoch_batches(pool, generator=torch.Generator().manual_seed(0))))
    assert b0["rays"].shape == (100, 11) and torch.equal(b0["rays"], pool["rays"][b0["index"]])
    assert sum(1 for _ in tr.epoch_batches(pool, drop_last=True)) == 10
