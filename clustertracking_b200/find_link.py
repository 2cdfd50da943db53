"""The FIND half of the reference's ``find_link`` (clustertracking/find_link.py:386-485, 914-990):
``preprocess`` -> ``grey_dilation`` -> ``characterize`` -> ``minmass`` filter, frame by frame, i.e.
everything ``_find_link_iter`` does to a frame before it hands the coordinates to the linker.

The LINKING itself (``Linker`` / ``FindLinker`` / ``Subnets``, find_link.py:82-912: a sequential,
frame-to-frame assignment with relocation) is host-side control logic outside the data-parallel
hot path and is not part of this package: ``find_features`` returns the per-frame feature table
that ``refine_leastsq`` (and a linker of the caller's choice) takes.
"""
import numpy as np
import pandas as pd

from . import find as _find
from . import preprocessing as _pre
from .utils import is_isotropic, validate_tuple


def find_features(frames, separation, diameter=None, minmass=0, noise_size=1, smoothing_size=None,
                  threshold=None, percentile=64, first_frame=0):
    """Feature table of a video: for every frame the steps of find_link.py:954-969 --
    ``preprocess`` (bandpass + rescaling), ``grey_dilation`` of the processed image (margin rule of
    find_link.py:935-936, ``precise=True``), ``characterize`` on the RAW image and the ``minmass``
    filter.  ``frames``: a sequence of equally shaped integer arrays or one stacked array.  The
    dilation / maxima search of all frames runs on the GPU in one batch (``ctk_find_maxima``).

    -> DataFrame with the position columns, ``mass``, ``signal``, ``size`` (or ``size_<axis>``) and
    ``frame`` (= first_frame + position in ``frames``)."""
    if not isinstance(frames, np.ndarray):
        frames = list(frames)
    if len(frames) == 0:
        raise ValueError("no frames")
    ndim = np.asarray(frames[0]).ndim
    shape = np.asarray(frames[0]).shape
    separation = validate_tuple(separation, ndim)
    if smoothing_size is None:                                              # find_link.py:440-441
        smoothing_size = separation
    smoothing_size = validate_tuple(smoothing_size, ndim)
    diameter = separation if diameter is None else validate_tuple(diameter, ndim)
    isotropic = is_isotropic(diameter)
    radius = tuple(int(d // 2) for d in diameter)
    margin = tuple(int(max(d // 2, s // 2 - 1)) for d, s in zip(diameter, separation))
    if any(s <= 2 * m for s, m in zip(shape, margin)):                      # find_link.py:939-949
        raise ValueError('The feature finding margins are larger than the image shape. Please use '
                         'smaller radius, separation or smoothing_size.')
    # the per-frame host steps (scipy.ndimage / numpy release the GIL) run on a thread pool
    from concurrent.futures import ThreadPoolExecutor
    from .utils import host_threads
    with ThreadPoolExecutor(host_threads(32)) as pool:
        processed = list(pool.map(lambda fr: np.asarray(_pre.preprocess(np.asarray(fr), noise_size,
                                                                        smoothing_size, threshold)),
                                  frames))
        found = _find.grey_dilation_batch(processed, separation, percentile, margin, precise=True)
        del processed
        pos_columns = ['z', 'y', 'x'][-ndim:]

        def describe(k):
            coords = found[k]
            if len(coords) == 0:
                return None
            extra = _pre.characterize(coords, np.asarray(frames[k]), radius, isotropic)   # find_link.py:964
            keep = extra['mass'] >= minmass
            table = pd.DataFrame(np.asarray(coords, dtype=np.float64)[keep], columns=pos_columns)
            for key, values in extra.items():
                table[key] = values[keep]
            table['frame'] = first_frame + k
            return table

        rows = [t for t in pool.map(describe, range(len(found))) if t is not None]
    if not rows:
        return pd.DataFrame(columns=pos_columns + ['mass', 'signal', 'frame'])
    return pd.concat(rows, ignore_index=True)
