"""Stand-ins for trackpy.masks (third-party, absent here; restated from trackpy's published
algorithm, v0.3 series, masks.py: PARITY UNPINNED, see preprocessing.py in this directory)."""
import numpy as np

from .utils import validate_tuple


def gaussian_kernel(sigma, truncate=4.0):
    """1D discretised, normalised gaussian on [-lw, lw], lw = int(truncate*sigma + .5)."""
    lw = int(truncate * sigma + 0.5)
    x = np.arange(-lw, lw + 1)
    result = np.exp(x ** 2 / (-2 * sigma ** 2))
    return result / np.sum(result)


def _coords(radius, ndim):
    radius = validate_tuple(radius, ndim)
    points = [np.arange(-rad, rad + 1) for rad in radius]
    if len(radius) > 1:
        coords = np.array(np.meshgrid(*points, indexing="ij"))
    else:
        coords = np.array([points[0]])
    r = [(coord / rad) ** 2 for (coord, rad) in zip(coords, radius)]
    return coords, sum(r)


def binary_mask(radius, ndim):
    "Elliptical mask in a rectangular array"
    return _coords(radius, ndim)[1] <= 1


def r_squared_mask(radius, ndim):
    "Mask with values r^2 inside the ellipse, 0 outside"
    coords, r = _coords(radius, ndim)
    r2 = np.sum(coords ** 2, 0).astype(int)
    r2[r > 1] = 0
    return r2


def x_squared_masks(radius, ndim):
    "Per-axis masks with values x_k^2 inside the ellipse, 0 outside"
    coords, r = _coords(radius, ndim)
    masks = np.asarray(coords ** 2, dtype=int)
    masks[:, r > 1] = 0
    return masks
