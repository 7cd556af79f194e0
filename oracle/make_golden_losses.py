"""Golden vectors for the depth-loss variants no rendering case exercises (test infrastructure, run in the
build container only: imports the UNMODIFIED reference from /root/reference):
    python oracle/make_golden_losses.py   ->  tests/golden/depth_loss_variants.npz
Cases: GNLL subset (metrics.py:129-130, std passed as var), MSE all-depth (:140,:154-156), MSE subset, each on the
same synthetic (z, weights, depth, targets); stored: inputs, loss, d loss / d depth, d loss / d weights."""
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import import_reference  # noqa: E402


def main():
    warnings.filterwarnings("ignore")
    _, _, ref_metrics = import_reference()
    g = torch.Generator().manual_seed(123)
    b, n = 96, 64
    z = torch.sort(torch.rand(b, n, generator=g) * 0.2, -1).values
    w = torch.softmax(torch.randn(b, n, generator=g) * 2, -1) * torch.rand(b, 1, generator=g)
    depth = (w * z).sum(-1)
    valid = (torch.rand(b, generator=g) < 0.7).long()
    td = 0.02 + 0.17 * torch.rand(b, generator=g)
    tw = torch.rand(b, generator=g)
    tstd = (1 - tw + 1e-4) * 0.05
    store = {"in_z": z.numpy(), "in_weights": w.numpy(), "in_depth": depth.numpy(), "in_valid": valid.numpy(),
             "in_target_depth": td.numpy(), "in_target_weight": tw.numpy(), "in_target_std": tstd.numpy()}
    cases = {"gnll_subset": dict(GNLL=True, usealldepth=False), "mse_all": dict(GNLL=False, usealldepth=True),
             "mse_subset": dict(GNLL=False, usealldepth=False)}
    for name, kw in cases.items():
        d_ = depth.clone().requires_grad_(True)
        w_ = w.clone().requires_grad_(True)
        res = {"z_vals_coarse": z, "depth_coarse": d_, "weights_coarse": w_}
        loss_fn = ref_metrics.DepthLoss(lambda_ds=1.5, margin=1e-4, stdscale=1.0, **kw)
        loss, ld = loss_fn(res, td, tw, target_valid_depth=valid, target_std=tstd)
        gd, gw = torch.autograd.grad(loss, [d_, w_], allow_unused=True)
        store[f"{name}_loss"] = np.array([float(loss)], dtype=np.float64)
        store[f"{name}_g_depth"] = (gd if gd is not None else torch.zeros_like(depth)).numpy()
        store[f"{name}_g_weights"] = (gw if gw is not None else torch.zeros_like(w)).numpy()
        print(name, float(loss), sorted(ld))
    store["meta"] = np.frombuffer(json.dumps({"lambda_ds": 1.5, "torch": torch.__version__}).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "depth_loss_variants.npz"), **store)


if __name__ == "__main__":
    main()
