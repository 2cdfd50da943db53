"""Helpers to read the fixtures written by ``oracle/make_golden.py``."""
import glob
import json
import os

import numpy as np
import pandas as pd

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def frame(data, prefix):
    cols = [str(c) for c in data[prefix + "columns"]]
    df = pd.DataFrame({c: data[prefix + "col_" + c] for c in cols}, index=data[prefix + "index"])
    return df[cols]


def meta(data):
    return json.loads(str(data["meta"]))


class Video(object):
    """Pre-rendered stack exposing what ``refine_leastsq`` needs from a pims reader."""

    def __init__(self, stack, first_frame=0):
        self.stack = stack
        self.first_frame = first_frame
        self.frame_shape = stack.shape[1:]

    def __getitem__(self, i):
        return self.stack[int(i) - self.first_frame]

    def __len__(self):
        return len(self.stack)


def refine_inputs(data, constraints_module):
    """-> (f0, reader, diameter, kwargs) for a ``refine_*`` fixture."""
    m = meta(data)
    kwargs = dict(m["kwargs"])
    diameter = m["diameter"]
    if isinstance(diameter, list):
        diameter = tuple(diameter)
    if m.get("constraint"):
        kind, dist = m["constraint"]
        ndim = 3 if "in_col_z" in data else 2
        kwargs["constraints"] = getattr(constraints_module, kind)(dist, ndim)
    image = data["image"]
    f0 = frame(data, "in_")
    ndim = 3 if "z" in f0.columns else 2
    reader = Video(image, int(m.get("first_frame", 0))) if image.ndim == ndim + 1 else image
    return f0, reader, diameter, kwargs
