"""Golden vectors for ray generation / DSM point clouds, produced by the reference's OWN functions
imported unmodified from /root/reference (build container only):
    python oracle/make_golden_geometry.py   ->  tests/golden/geometry.npz
`rasterio`, `rpcm`, `cv2` and `kornia` are imported by the reference's modules but not used by these functions; they are
stubbed the way kornia is stubbed for the renderer.  The RPC camera is replaced by a smooth synthetic localisation
around the JAX_269 footprint (rpcm is absent from the image; only its call signature matters here).  The oracle
restatement (oracle/geometry_oracle.py) is asserted bit-equal to the reference before anything is written."""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle import geometry_oracle as G  # noqa: E402


def import_reference_geometry():
    for name in ("rasterio", "rpcm", "kornia", "kornia.losses"):
        sys.modules.setdefault(name, types.ModuleType(name))
    try:
        import cv2  # noqa: F401
    except ImportError:
        cv = types.ModuleType("cv2")
        cv.COLORMAP_JET = 2
        sys.modules["cv2"] = cv
    sys.modules["kornia.losses"].ssim = lambda *a, **k: None
    sys.modules["kornia"].losses = sys.modules["kornia.losses"]
    for name in [m for m in sys.modules if m.split(".")[0] in ("models", "modules", "datasets")]:
        del sys.modules[name]
    sys.path.insert(0, REF)
    try:
        from modules import utils as ref_utils
        from datasets import satellite_scene as ref_scene
    finally:
        sys.path.remove(REF)
    return ref_utils, ref_scene


class SyntheticCamera:
    """Stands in for rpcm.RPCModel: a pushbroom-like view 20 degrees off nadir over Jacksonville."""

    def localization(self, cols, rows, alts):
        cols, rows, alts = (np.asarray(v, np.float64) for v in (cols, rows, alts))
        lon = -81.6634 + 3.1e-6 * (cols - 400) + 2.0e-7 * (rows - 400) + 4.1e-6 * (alts / 100.0) + 1e-12 * cols * rows
        lat = 30.3168 - 2.7e-6 * (rows - 400) + 1.5e-7 * (cols - 400) - 2.9e-6 * (alts / 100.0) + 1e-12 * cols * cols
        return lon, lat


def main():
    ref_utils, ref_scene = import_reference_geometry()
    rng = np.random.default_rng(2024)
    cols, rows = np.meshgrid(np.arange(0, 800, 11), np.arange(0, 800, 13))
    cols, rows = cols.flatten(), rows.flatten()
    cam, min_alt, max_alt = SyntheticCamera(), -29.0, 73.0
    store = {"cols": cols, "rows": rows, "alts": np.array([min_alt, max_alt])}
    store["lon_near"], store["lat_near"] = cam.localization(cols, rows, max_alt * np.ones(cols.shape))
    store["lon_far"], store["lat_far"] = cam.localization(cols, rows, min_alt * np.ones(cols.shape))

    rays = ref_scene.get_rays(cols, rows, cam, min_alt, max_alt)                        # satellite_scene.py:21-68
    store["rays"] = rays.numpy()
    near, far = rays[:, :3], rays[:, :3] + rays[:, 7:8] * rays[:, 3:6]                  # :404-411
    pts = torch.cat([near, far], 0)
    sc = [ref_utils.rpc_scaling_params(pts[:, k]) for k in range(3)]
    ds = object.__new__(ref_scene.SatelliteSceneDataset)                                # methods only need center / range
    ds.center = torch.tensor([float(s[1]) for s in sc])                                 # :123
    ds.range = torch.max(torch.tensor([float(s[0]) for s in sc]))                       # :124
    store["center"], store["range"] = ds.center.numpy(), np.array([float(ds.range)], np.float32)
    normed = ds.normalize_rays(rays.clone())                                            # :415-425
    store["rays_normalized"] = normed.numpy()
    sun = ds.get_sun_dirs(61.3, 152.4, 1)                                               # :449-473
    store["sun_dir"] = sun.numpy()[0]
    rays11 = torch.cat([normed, sun.repeat(normed.shape[0], 1)], 1)
    depth = torch.from_numpy(rng.uniform(0.0, 1.0, normed.shape[0]).astype(np.float32)) * normed[:, 7]
    store["depth"] = depth.numpy()
    lats, lons, alts = ds.get_latlonalt_from_nerf_prediction(rays11, depth.view(-1, 1))  # :475-505
    store["lat"], store["lon"], store["alt"] = lats, lons, alts
    # modules/utils.py functions on their own
    la, lo, al = rng.uniform(-80, 80, 2000), rng.uniform(-180, 180, 2000), rng.uniform(-100, 9000, 2000)
    x, y, z = ref_utils.geodetic_to_ecef(la, lo, al)
    store.update(u_lat=la, u_lon=lo, u_alt=al, u_x=x, u_y=y, u_z=z)
    back = ref_utils.ecef_to_latlon_custom(x, y, z)
    store.update(u_back_lat=back[0], u_back_lon=back[1], u_back_alt=back[2])

    # ---- pin the oracle restatement: bit-equal ----
    o_rays = G.rays_from_localization(store["lon_near"], store["lat_near"], store["lon_far"], store["lat_far"], min_alt, max_alt)
    assert np.array_equal(o_rays, store["rays"])
    c, r = G.scene_scaling(o_rays)
    assert np.array_equal(c, store["center"]) and np.float32(r) == store["range"][0]
    assert np.array_equal(G.normalize_rays(o_rays, c, r), store["rays_normalized"])
    o_lat, o_lon, o_alt = G.points_to_geodetic(store["rays_normalized"], store["depth"], c, r)
    assert np.array_equal(o_lat, lats) and np.array_equal(o_lon, lons) and np.array_equal(o_alt, alts)
    assert all(np.array_equal(a, b) for a, b in zip(G.geodetic_to_ecef(la, lo, al), (x, y, z)))
    assert all(np.array_equal(a, b) for a, b in zip(G.ecef_to_geodetic(x, y, z), back))
    print("oracle restatement bit-equal to the reference on", len(cols), "rays; altitude round trip max err",
          float(np.abs(back[2] - al).max()), "m")
    store["meta"] = np.frombuffer(json.dumps({"numpy": np.__version__, "torch": torch.__version__, "n": int(len(cols))}).encode(),
                                  dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "geometry.npz"), **store)


if __name__ == "__main__":
    main()
