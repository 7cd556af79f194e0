"""Ray generation and DSM point clouds on the device: mirror of the ray-facing functions of
datasets/satellite_scene.py (get_rays :21-68, normalize_rays :415-425, get_sun_dirs :463-473,
get_latlonalt_from_nerf_prediction :475-505) over the geodesy kernels of include/spnerf_b200.h.

What stays on the host, and why: `rpc.localization` (the rpcm package: iterative inversion of the rational
polynomial camera) before get_rays, and the UTM projection + rasterisation (pyproj, plyflatten) after the point
cloud -- neither library is available in this image, so no oracle could be pinned for them.  The caller passes an
object with the rpcm interface (`localization(cols, rows, alts) -> lons, lats`), exactly as the reference does.
"""
import ctypes

import numpy as np
import torch

from .. import _cabi
from ..engine import _p, _require_cuda, _stream


def _dev_f64(a, device):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(device)


def rays_from_localization(lons_near, lats_near, lons_far, lats_far, min_alt, max_alt, device, center=None,
                           scene_range=None, sun_dir=None):
    """Pixels localised at the maximum (near) and minimum (far) altitude -> (n, 8) float32 rays
    [origin, direction, near = 0, far]; with `center` / `scene_range` the rows are normalised like normalize_rays,
    with `sun_dir` (3 floats) they get the three sun-direction columns render_rays expects ((n, 11))."""
    device = torch.device(device)
    n = int(np.asarray(lons_near).size)
    a = _cabi.RaysFromGeodetic()
    bufs = [_dev_f64(v, device) for v in (lons_near, lats_near, lons_far, lats_far)]
    _require_cuda(bufs[0], "localised pixels")
    a.lon_near, a.lat_near, a.lon_far, a.lat_far = (_p(b) for b in bufs)
    a.alt_near, a.alt_far = float(max_alt), float(min_alt)
    a.normalize = 1 if center is not None else 0
    if center is not None:
        for k in range(3):
            a.center[k] = float(center[k])
        a.range = float(scene_range)
    a.has_sun = 1 if sun_dir is not None else 0
    if sun_dir is not None:
        for k in range(3):
            a.sun_dir[k] = float(sun_dir[k])
    width = 11 if sun_dir is not None else 8
    rays = torch.empty(n, width, dtype=torch.float32, device=device)
    a.row_stride, a.n_rays, a.rays = width, n, _p(rays)
    _cabi.check(_cabi.lib().spnerf_rays_from_geodetic(ctypes.byref(a), _stream()), "spnerf_rays_from_geodetic")
    return rays


def get_rays(cols, rows, rpc, min_alt, max_alt, device="cuda"):
    """datasets/satellite_scene.py:21-68 with the same arguments (+ the device): the two RPC localisations run on the
    host through the caller's `rpc` object, everything after them on the GPU."""
    cols, rows = np.asarray(cols), np.asarray(rows)
    lons_n, lats_n = rpc.localization(cols, rows, float(max_alt) * np.ones(cols.shape))
    lons_f, lats_f = rpc.localization(cols, rows, float(min_alt) * np.ones(cols.shape))
    return rays_from_localization(lons_n, lats_n, lons_f, lats_f, min_alt, max_alt, device)


def normalize_rays(rays, center, scene_range):
    """datasets/satellite_scene.py:415-425 (in place, float32), for rays produced elsewhere."""
    c = torch.as_tensor(center, dtype=torch.float32, device=rays.device)
    r = torch.as_tensor(scene_range, dtype=torch.float32, device=rays.device)
    rays[:, 0:3] -= c
    rays[:, 0:3] /= r
    rays[:, 6:8] /= r
    return rays


def get_sun_dirs(sun_elevation_deg, sun_azimuth_deg, n_rays, device="cuda"):
    """datasets/satellite_scene.py:446-473: the sun direction of an image, repeated per ray."""
    el, az = np.radians(float(sun_elevation_deg)), np.radians(float(sun_azimuth_deg))
    sun_d = np.array([np.sin(az) * np.cos(el), np.cos(az) * np.cos(el), np.sin(el)])
    return torch.from_numpy(np.tile(sun_d, (n_rays, 1))).type(torch.FloatTensor).to(device)


def get_latlonalt_from_nerf_prediction(rays, depth, center, scene_range):
    """datasets/satellite_scene.py:475-505: (n, >= 6) normalised rays + (n) predicted depths -> latitudes, longitudes
    (degrees) and altitudes (metres) as float64 device tensors (the reference returns numpy arrays)."""
    rays = rays.detach().float().contiguous()
    _require_cuda(rays, "rays")
    depth = depth.detach().float().reshape(-1).contiguous()
    n = rays.shape[0]
    out = [torch.empty(n, dtype=torch.float64, device=rays.device) for _ in range(3)]
    a = _cabi.PointsToGeodetic()
    a.rays, a.row_stride, a.depth, a.n_rays = _p(rays), rays.shape[1], _p(depth), n
    for k in range(3):
        a.center[k] = float(center[k])
    a.range = float(scene_range)
    a.lat, a.lon, a.alt = (_p(t) for t in out)
    _cabi.check(_cabi.lib().spnerf_points_to_geodetic(ctypes.byref(a), _stream()), "spnerf_points_to_geodetic")
    return tuple(out)
