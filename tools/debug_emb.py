import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import spnerf_b200
from spnerf_b200 import config, synthetic, train_step, engine as E
from spnerf_b200.models import load_model
DEV = "cuda:0"
args = config.make_args(sem=True, num_sem_classes=3, fc_units=512)
torch.manual_seed(0)
model = load_model(args).to(DEV)
batch = {k: v.to(DEV) for k, v in synthetic.make_batch(512, seed=31).items()}
eng = model.engine
for it in range(2):
    flat, views, scalars, launches = train_step.fused_step(model, args, batch, repack=True)
    torch.cuda.synchronize()
    d = dict(zip(eng.names, views))
    print("iter", it, "emb grad", d["semantic_embedding.weight"].flatten().tolist())
    print(" sky b2", d["sky_color.2.bias"].tolist(), "rgb2 bias", d["rgb_from_xyzdir.2.bias"].tolist(), "sem2 bias", d["logit_from_label.2.bias"].tolist())
    print(" accum nonzero", int((eng.accum != 0).sum()), "absmax", float(eng.absmax), "scalars", scalars[:5].tolist())
