"""Import the UNMODIFIED reference (``/root/reference/clustertracking``) in the build container.

TEST INFRASTRUCTURE ONLY -- used by ``oracle/make_golden.py`` to produce the golden
vectors in ``tests/golden/``.  ``/root/reference`` does not exist on the GPU box, so
nothing that runs there may import this module.

The reference does not import as shipped under numpy 2 / pandas 3 without trackpy and
pims.  This loader therefore

1. puts the stub ``trackpy`` / ``pims`` packages of ``oracle/ref_shim`` on ``sys.path``;
2. restores three numpy aliases that numpy >= 1.24 removed (``np.int``, ``np.float``, ``np.Inf``);
3. copies the reference package to a scratch directory under ``/tmp`` (the mount is read-only and
   reference sources must never be copied into this repository) and rewrites the five
   list-of-slices indexing expressions that numpy >= 1.23 rejects
   (``masks.py:23,25,68``; ``artificial.py:139,141``).  No arithmetic is touched.
"""
import importlib
import os
import shutil
import sys
import tempfile

import numpy as np

REFERENCE_ROOT = os.environ.get("CTK_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shim")

_PATCHES = {
    "masks.py": [
        ("cropped = image[[slice(c, c+s) for (c, s) in zip(padded_corner, shape)]]",
         "cropped = image[tuple([slice(c, c+s) for (c, s) in zip(padded_corner, shape)])]"),
        ("cropped = image[[slice(c, c+s) for (c, s) in zip(corner, shape)]]",
         "cropped = image[tuple([slice(c, c+s) for (c, s) in zip(corner, shape)])]"),
        ("return image[slices], origin", "return image[tuple(slices)], origin"),
    ],
    "artificial.py": [
        ("r = np.sqrt(np.sum(np.array(coords)**2, axis=0))",
         "r = np.sqrt(sum(c**2 for c in coords))"),
        ("image[rect] += spot.astype(image.dtype)",
         "image[tuple(rect)] += spot.astype(image.dtype)"),
    ],
}


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "clustertracking"))


def load():
    """Return the imported reference package (module object named ``clustertracking``)."""
    if "clustertracking" in sys.modules:
        return sys.modules["clustertracking"]
    if not available():
        raise ImportError("reference not present at %s" % REFERENCE_ROOT)
    for name, value in (("int", int), ("float", float), ("Inf", np.inf)):
        if not hasattr(np, name):
            setattr(np, name, value)
    scratch = tempfile.mkdtemp(prefix="ctk_ref_")
    dst = os.path.join(scratch, "clustertracking")
    shutil.copytree(os.path.join(REFERENCE_ROOT, "clustertracking"), dst,
                    ignore=shutil.ignore_patterns("tests", "__pycache__"))
    for fname, subs in _PATCHES.items():
        path = os.path.join(dst, fname)
        with open(path) as fh:
            text = fh.read()
        for old, new in subs:
            if old not in text:
                raise RuntimeError("patch target not found in %s: %r" % (fname, old))
            text = text.replace(old, new)
        with open(path, "w") as fh:
            fh.write(text)
    sys.path.insert(0, _SHIM)
    sys.path.insert(0, scratch)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("clustertracking")
