"""Small driver for ncu: two BASELINE config 3 training steps (guided sampling + mapping, 16384 rays) through the public API."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import spnerf_b200  # noqa: F401
from spnerf_b200 import synthetic
from spnerf_b200.modules import metrics
from spnerf_b200.modules.rendering import render_rays

dev = torch.device("cuda:0")
args3 = bench.make_args(mapping=True, guidedsample=True, chunk=16384)
model3 = bench.build_model(args3, dev)
b3 = {k: v.to(dev) for k, v in synthetic.make_batch(16384, seed=300).items()}
loss_fn, dl, sl = metrics.SNerfLoss(0.0), metrics.DepthLoss(1.0, usealldepth=False), metrics.SemanticLoss(1.0)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    res = render_rays({"coarse": model3}, args3, b3["rays"], None, semantics=b3["sems"], mode="train",
                      valid_depth=b3["valid_depth"], target_depths=b3["depths"], target_std=b3["depth_std"])
    loss = loss_fn(res, b3["rgbs"])[0] + dl(res, b3["depths"][:, 0], b3["depths"][:, 1], target_valid_depth=b3["valid_depth"],
                                            target_std=b3["depth_std"])[0] + sl(res, b3["sems"])[0]
    for p_ in model3.parameters():
        p_.grad = None
    loss.backward()
torch.cuda.synchronize()
print("profile_c3 done")
