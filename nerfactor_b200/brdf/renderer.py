"""Mirror of brdf/renderer.py:184-219 (`gen_light_xyz`), init-time host code: where the lights of
an h x w lat-long environment map sit (at `envmap_radius`) and which share of the sphere each
one stands for.  Bit-for-bit the reference's fp64 numbers (tests/golden/ref_pinned.npz)."""
import numpy as np


def gen_light_xyz(envmap_h, envmap_w, envmap_radius=1e2):
    """-> (xyz [h, w, 3], areas [h, w]).  Row 0 is the top of the map (latitude just below
    +pi/2), column 0 its left edge (longitude just below +pi); the two polar rows and the seam
    column a plain linspace would produce are left out by shrinking the range by one step of an
    (h + 2) x (w + 2) grid.  areas = 4 pi sin(colatitude) / sum sin(colatitude)."""
    h, w = int(envmap_h), int(envmap_w)
    dlat, dlng = np.pi / (h + 2), 2 * np.pi / (w + 2)
    lat_1d = np.linspace(np.pi / 2 - dlat, -np.pi / 2 + dlat, h)
    lng_1d = np.linspace(np.pi - dlng, -np.pi + dlng, w)
    lat = np.tile(lat_1d[:, None], (1, w))            # [h, w], contiguous like np.meshgrid's copies
    lng = np.tile(lng_1d[None, :], (h, 1))
    radius = envmap_radius * np.ones_like(lat)
    # latitude / longitude -> Cartesian, z up (xiuminglib geometry/sph.py:184-193 'lat-lng')
    xyz = np.empty((h, w, 3))
    xyz[..., 2] = radius * np.sin(lat)
    xyz[..., 0] = radius * np.cos(lat) * np.cos(lng)
    xyz[..., 1] = radius * np.cos(lat) * np.sin(lng)
    weight = np.sin(np.pi / 2 - lat)
    areas = 4 * np.pi * weight / np.sum(weight)
    if not np.all(areas != 0):
        raise AssertionError("every light pixel must carry a non-zero solid angle")
    return xyz, areas
