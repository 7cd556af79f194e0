"""Geometry either side of the renderer (mirror of the parts of the reference's ``datasets`` package that touch rays)."""
from .satellite_scene import (get_rays, rays_from_localization, normalize_rays, get_sun_dirs,  # noqa: F401
                              get_latlonalt_from_nerf_prediction)
