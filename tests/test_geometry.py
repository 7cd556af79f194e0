"""Ray generation and DSM point clouds: the numpy oracle against the golden vectors produced by
the reference's own functions (CPU), and the CUDA kernels against both (GPU, through the C ABI).
Bars: rays float32 bit-exact (fp64 transcendentals of the device may differ from glibc's in the last place, which can
move a float32 rounding tie: at most 1 float32 ulp on at most 0.1 % of the entries); latitude / longitude to 1e-11
degrees, altitude to 1e-8 m (fp64 against fp64)."""
import os

import numpy as np
import pytest
import torch

import spnerf_b200  # noqa: F401
from oracle import geometry_oracle as G
from parity_common import GOLDEN


def _g():
    return np.load(os.path.join(GOLDEN, "geometry.npz"))


def test_geometry_oracle_matches_reference_golden():
    g = _g()
    rays = G.rays_from_localization(g["lon_near"], g["lat_near"], g["lon_far"], g["lat_far"], g["alts"][0], g["alts"][1])
    assert rays.dtype == np.float32 and np.array_equal(rays, g["rays"])
    c, r = G.scene_scaling(rays)
    assert np.array_equal(c, g["center"]) and np.float32(r) == g["range"][0]
    assert np.array_equal(G.normalize_rays(rays, c, r), g["rays_normalized"])
    lat, lon, alt = G.points_to_geodetic(g["rays_normalized"], g["depth"], c, r)
    # the same numpy on the same ISA reproduces the bits; leave room for another host's libm
    assert np.allclose(lat, g["lat"], rtol=0, atol=1e-12) and np.allclose(lon, g["lon"], rtol=0, atol=1e-12)
    assert np.allclose(alt, g["alt"], rtol=0, atol=1e-8)
    x, y, z = G.geodetic_to_ecef(g["u_lat"], g["u_lon"], g["u_alt"])
    assert np.allclose(x, g["u_x"], rtol=0, atol=1e-8) and np.allclose(z, g["u_z"], rtol=0, atol=1e-8)
    back = G.ecef_to_geodetic(x, y, z)
    assert np.abs(back[2] - g["u_alt"]).max() < 1e-5        # the reference's closed form is accurate to ~1e-6 m
    # direction vectors are unit length, near = 0, the far point sits at the minimum altitude
    assert np.abs(np.linalg.norm(rays[:, 3:6].astype(np.float64), axis=1) - 1).max() < 1e-6
    assert np.all(rays[:, 6] == 0)


def _ulp_diff(a, b):
    ai, bi = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    return np.abs(ai - bi)


@pytest.mark.gpu
def test_rays_from_localization_on_the_device():
    from spnerf_b200.datasets import rays_from_localization
    g = _g()
    args = (g["lon_near"], g["lat_near"], g["lon_far"], g["lat_far"], g["alts"][0], g["alts"][1], "cuda:0")
    raw = rays_from_localization(*args).cpu().numpy()
    for got, want in ((raw, g["rays"]),
                      (rays_from_localization(*args, center=g["center"], scene_range=g["range"][0]).cpu().numpy(),
                       g["rays_normalized"])):
        d = _ulp_diff(got, want)
        assert d.max() <= 1 and (d > 0).mean() <= 1e-3, (int(d.max()), float((d > 0).mean()))
    full = rays_from_localization(*args, center=g["center"], scene_range=g["range"][0], sun_dir=g["sun_dir"])
    assert full.shape == (raw.shape[0], 11)
    assert torch.equal(full[:, 8:11].cpu(), torch.from_numpy(np.tile(g["sun_dir"], (raw.shape[0], 1))))
    assert np.array_equal(full[:, :8].cpu().numpy(), rays_from_localization(
        *args, center=g["center"], scene_range=g["range"][0]).cpu().numpy())


@pytest.mark.gpu
def test_get_rays_with_an_rpc_like_camera():
    """The reference's call: get_rays(cols, rows, rpc, min_alt, max_alt) with any object that has rpcm's localization()."""
    from oracle.make_golden_geometry import SyntheticCamera
    from spnerf_b200.datasets import get_rays
    g = _g()
    rays = get_rays(g["cols"], g["rows"], SyntheticCamera(), g["alts"][0], g["alts"][1], device="cuda:0")
    d = _ulp_diff(rays.cpu().numpy(), g["rays"])
    assert d.max() <= 1 and (d > 0).mean() <= 1e-3


@pytest.mark.gpu
def test_dsm_points_on_the_device():
    from spnerf_b200.datasets import get_latlonalt_from_nerf_prediction
    g = _g()
    rays = torch.from_numpy(g["rays_normalized"]).to("cuda:0")
    depth = torch.from_numpy(g["depth"]).to("cuda:0")
    lat, lon, alt = get_latlonalt_from_nerf_prediction(rays, depth.view(-1, 1), g["center"], g["range"][0])
    assert lat.dtype == torch.float64
    assert float((lat.cpu() - torch.from_numpy(g["lat"])).abs().max()) <= 1e-11
    assert float((lon.cpu() - torch.from_numpy(g["lon"])).abs().max()) <= 1e-11
    assert float((alt.cpu() - torch.from_numpy(g["alt"])).abs().max()) <= 1e-8
    # full-size property: depth 0 gives the maximum altitude, depth = far the minimum (4 M rays in a few ms)
    big = rays.repeat(1000, 1)[: 2048 * 2048]
    _, _, a0 = get_latlonalt_from_nerf_prediction(big, torch.zeros(big.shape[0], device="cuda:0"), g["center"], g["range"][0])
    _, _, a1 = get_latlonalt_from_nerf_prediction(big, big[:, 7].contiguous(), g["center"], g["range"][0])
    assert float((a0 - g["alts"][1]).abs().max()) < 0.6 and float((a1 - g["alts"][0]).abs().max()) < 0.6   # float32 ECEF: 0.5 m ulp
