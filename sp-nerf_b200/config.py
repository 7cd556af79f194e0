"""The `args` fields the hot path reads, with the reference's defaults (modules/opt.py:31-107), for callers
that have no argparse namespace (bench.py, tools/, the trainer tests)."""
from types import SimpleNamespace


def make_args(**kw):
    d = dict(n_samples=64, n_importance=0, model="sp-nerf", beta=False, guidedsample=False, sc_lambda=0.0,
             margin=1e-4, stdscale=1.0, chunk=5120, noise_std=0.0, num_sem_classes=3, s_embedding_factor=1,
             fc_layers=8, fc_units=512, mapping=False, t_embbeding_tau=4, t_embbeding_vocab=30, sem=False,
             mapping_freqs=10, skips=(4,), lr=5e-4, batch_size=1024, max_train_steps=500000, depth=False,
             ds_lambda=0.0, ds_drop=0.25, GNLL=False, usealldepth=False, ss_lambda=4e-2, ss_drop=1.0)
    d.update(kw)
    return SimpleNamespace(**d)
