"""CPU restatement (numpy, fp64) of the reference's ray generation and DSM point-cloud geometry.  TEST INFRASTRUCTURE ONLY
(imported by tests/ and nothing else); pinned against outputs of the reference's own functions by
oracle/make_golden_geometry.py -> tests/golden/geometry.npz.

Follows: modules/utils.py:80-100 (geodetic_to_ecef), :103-120 (ecef_to_latlon_custom), :49-56 (rpc_scaling_params);
datasets/satellite_scene.py:38-68 (get_rays after the two rpc.localization calls), :415-425 (normalize_rays),
:488-503 (get_latlonalt_from_nerf_prediction)."""
import numpy as np

WGS84_A = 6378137.0
WGS84_B = 6356752.314245


def geodetic_to_ecef(lat, lon, alt):                      # modules/utils.py:80-100
    e2 = 1 - (WGS84_B ** 2 / WGS84_A ** 2)
    la, lo = np.radians(lat), np.radians(lon)
    n = WGS84_A / np.sqrt(1 - e2 * np.sin(la) ** 2)
    return ((n + alt) * np.cos(la) * np.cos(lo), (n + alt) * np.cos(la) * np.sin(lo),
            ((WGS84_B ** 2 / WGS84_A ** 2) * n + alt) * np.sin(la))


def ecef_to_geodetic(x, y, z):                            # modules/utils.py:103-120
    a, e = 6378137.0, 8.1819190842622e-2
    asq, esq = a ** 2, e ** 2
    b = np.sqrt(asq * (1 - esq))
    bsq = b ** 2
    ep = np.sqrt((asq - bsq) / bsq)
    p = np.sqrt(x ** 2 + y ** 2)
    th = np.arctan2(a * z, b * p)
    lon = np.arctan2(y, x)
    lat = np.arctan2(z + (ep ** 2) * b * (np.sin(th) ** 3), p - esq * a * (np.cos(th) ** 3))
    n = a / np.sqrt(1 - esq * (np.sin(lat) ** 2))
    alt = p / np.cos(lat) - n
    return lat * 180 / np.pi, lon * 180 / np.pi, alt


def rays_from_localization(lons_near, lats_near, lons_far, lats_far, min_alt, max_alt):
    """datasets/satellite_scene.py:38-68 -> (n, 8) float32 [origin, unit direction, 0, |far - near|]."""
    near = np.vstack(geodetic_to_ecef(lats_near, lons_near, float(max_alt) * np.ones(lons_near.shape))).T
    far = np.vstack(geodetic_to_ecef(lats_far, lons_far, float(min_alt) * np.ones(lons_far.shape))).T
    d = far - near
    norm = np.linalg.norm(d, axis=1)
    rays = np.hstack([near, d / norm[:, None], 0.0 * norm[:, None], norm[:, None]])
    return rays.astype(np.float32)


def scene_scaling(rays):
    """datasets/satellite_scene.py:404-411 + :122-124: centre (3 float32) and range (float32) of a set of rays."""
    r = rays.astype(np.float32)
    pts = np.concatenate([r[:, :3], r[:, :3] + r[:, 7:8] * r[:, 3:6]], 0)
    scale, offset = [], []
    for k in range(3):                                     # modules/utils.py:49-56
        v = pts[:, k]
        s = (v.max() - v.min()) / 2
        scale.append(s)
        offset.append(v.min() + s)
    return np.array(offset, np.float32), np.float32(max(float(s) for s in scale))


def normalize_rays(rays, center, scene_range):            # datasets/satellite_scene.py:415-425 (float32 in place)
    r = rays.astype(np.float32).copy()
    c, s = center.astype(np.float32), np.float32(scene_range)
    for k in range(3):
        r[:, k] = (r[:, k] - c[k]) / s
    r[:, 6] = r[:, 6] / s
    r[:, 7] = r[:, 7] / s
    return r


def points_to_geodetic(rays, depth, center, scene_range):  # datasets/satellite_scene.py:488-503
    r = rays.astype(np.float64)
    xyz = (r[:, 0:3] + r[:, 3:6] * depth.astype(np.float64).reshape(-1, 1)) * float(np.float32(scene_range))
    xyz = xyz + center.astype(np.float32).astype(np.float64)[None, :]
    return ecef_to_geodetic(xyz[:, 0], xyz[:, 1], xyz[:, 2])
