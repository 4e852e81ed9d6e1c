"""Mirror of brdf/renderer.py:184-219 (`gen_light_xyz`), init-time host code."""
import numpy as np


def gen_light_xyz(envmap_h, envmap_w, envmap_radius=1e2):
    """Lat-long light positions (poles excluded) and per-pixel solid angles."""
    lat_step_size = np.pi / (envmap_h + 2)
    lng_step_size = 2 * np.pi / (envmap_w + 2)
    lats = np.linspace(
        np.pi / 2 - lat_step_size, -np.pi / 2 + lat_step_size, envmap_h)
    lngs = np.linspace(
        np.pi - lng_step_size, -np.pi + lng_step_size, envmap_w)
    lngs, lats = np.meshgrid(lngs, lats)
    r = envmap_radius * np.ones_like(lats)
    # xiuminglib sph2cart, 'lat-lng' convention (geometry/sph.py:184-193)
    z = r * np.sin(lats)
    x = r * np.cos(lats) * np.cos(lngs)
    y = r * np.cos(lats) * np.sin(lngs)
    xyz = np.stack((x, y, z), axis=-1)
    sin_colat = np.sin(np.pi / 2 - lats)
    areas = 4 * np.pi * sin_colat / np.sum(sin_colat)
    assert 0 not in areas, \
        "There shouldn't be light pixel that doesn't contribute"
    return xyz, areas
