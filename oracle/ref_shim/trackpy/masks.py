import numpy as np


def gaussian_kernel(sigma, truncate=4.0):
    """1D discretised, normalised gaussian on [-lw, lw], lw = int(truncate*sigma + .5)."""
    lw = int(truncate * sigma + 0.5)
    x = np.arange(-lw, lw + 1)
    result = np.exp(x ** 2 / (-2 * sigma ** 2))
    return result / np.sum(result)


def r_squared_mask(*args, **kwargs):
    raise NotImplementedError


def x_squared_masks(*args, **kwargs):
    raise NotImplementedError
