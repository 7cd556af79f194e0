"""Synthetic JAX_269-shaped ray batches (no dataset is available offline).

Statistics follow the bundled scene: origins on the top plane of the
256 m AOI in scene-normalised ECEF coordinates (centre and range from
Dataset/DFC2019_269/JSON/scene.loc), one near-constant viewing direction and far bound per image,
near = 0, sun direction (0,1,0) as in the four bundled JSONs, labels constant on 8x8 pixel blocks
(--dense_ss, datasets/satellite_scene.py:337-349) with ~2 % ignore labels (-100), depth priors on
~68 % of rays with the std model of datasets/satellite_scene.py:262,295.
The row layout is the reference's (B,11) [origin, direction, near, far, sun_dir]
(datasets/satellite_scene.py:577-592).
"""
import math

import torch

SCENE_CENTRE = (801532.84375, -5452266.5, 3200266.875)
SCENE_RANGE = 141.21875
VIEW_DIRS = ((-0.2876, 0.6954, -0.6586), (0.0024, 0.9402, -0.3406),
             (-0.1847, 0.8803, -0.4369), (-0.2125, 0.9578, -0.1934))
FAR = (0.2060, 0.2034, 0.1992, 0.2104)
IMG_W = 800


def _tangent_basis():
    up = torch.tensor(SCENE_CENTRE, dtype=torch.float64)
    up = up / up.norm()
    east = torch.linalg.cross(torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64), up)
    east = east / east.norm()
    north = torch.linalg.cross(up, east)
    return up.float(), east.float(), north.float()


def make_batch(n_rays, seed=269, n_images=2, sun_dir=(0.0, 1.0, 0.0), shuffled=True, device="cpu"):
    """A training-style batch.  Returns a dict with the keys of the reference's batch["color"]
    (datasets/satellite_scene.py:577-592): rays (B,11) f32, rgbs (B,3), ts (B,) long, sems (B,) long,
    valid_depth (B,) long, depths (B,2) [depth, correlation weight], depth_std (B,)."""
    g = torch.Generator().manual_seed(seed)
    up, east, north = _tangent_basis()
    if shuffled:
        img = torch.randint(0, n_images, (n_rays,), generator=g)
        px = torch.randint(0, IMG_W, (n_rays,), generator=g)
        py = torch.randint(0, IMG_W, (n_rays,), generator=g)
    else:  # consecutive pixels of image 0 (full-image inference order)
        idx = torch.arange(n_rays)
        img = torch.zeros(n_rays, dtype=torch.long)
        side = max(1, int(math.ceil(math.sqrt(n_rays))))
        px, py = idx % side, idx // side
    half = 128.0 / SCENE_RANGE                                   # 256 m AOI in normalised units
    span = float(max(int(px.max()) + 1, int(py.max()) + 1, IMG_W if shuffled else 1))
    u = (px.float() + 0.5) / span * 2 - 1
    v = (py.float() + 0.5) / span * 2 - 1
    origin = (u * half).unsqueeze(1) * east + (v * half).unsqueeze(1) * north + (-2.0 / SCENE_RANGE) * up
    dirs = torch.tensor(VIEW_DIRS)[img % 4] + 1e-4 * torch.randn(n_rays, 3, generator=g)
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    far = torch.tensor(FAR)[img % 4] + 1e-5 * (torch.rand(n_rays, generator=g) * 2 - 1)
    sun = torch.tensor(sun_dir).expand(n_rays, 3)
    rays = torch.cat([origin, dirs, torch.zeros(n_rays, 1), far.unsqueeze(1), sun], 1).float().contiguous()

    block = (px // 8) * 131 + (py // 8) * 17 + img * 7
    sems = (block * 2654435761 % 4294967296 // 65536) % 3
    ignore = torch.rand(n_rays, generator=g) < 0.02
    sems = torch.where(ignore, torch.full_like(sems, -100), sems).long()

    valid = (torch.rand(n_rays, generator=g) < 0.68).long()
    depth = 0.02 + 0.17 * torch.rand(n_rays, generator=g)
    corr = torch.rand(n_rays, generator=g)
    std = (1.0 * (1 - corr) + 1e-4) * 0.17
    depth = torch.where(valid > 0, depth, torch.zeros_like(depth))
    corr = torch.where(valid > 0, corr, torch.zeros_like(corr))
    std = torch.where(valid > 0, std, torch.zeros_like(std))
    batch = {
        "rays": rays,
        "rgbs": torch.rand(n_rays, 3, generator=g),
        "ts": (img % 2).long(),
        "sems": sems,
        "valid_depth": valid,
        "depths": torch.stack([depth, corr], 1).contiguous(),
        "depth_std": std,
    }
    return {k: t.to(device) for k, t in batch.items()}
