"""clustertracking_b200 -- B200-native implementation of clustertracking's ``refine_leastsq`` path.

Public names mirror the reference package (clustertracking/__init__.py:10-18) for the path this
project covers: ``refine_leastsq`` (cluster level and global level), ``find_clusters``,
``FitFunctions``, ``constraints``; and the callers either side of it: ``preprocess`` /
``characterize`` / ``find.grey_dilation`` / ``find_features`` (the find half of ``find_link``) and
``parallel.refine_leastsq_sharded`` (frames over the GPUs of one box).
"""
from .find import find_clusters                     # noqa: F401
from .fitfunc import FitFunctions                   # noqa: F401
from .refine import refine_leastsq                  # noqa: F401
from . import constraints                           # noqa: F401
from . import find                                  # noqa: F401
from . import preprocessing                         # noqa: F401
from .preprocessing import preprocess, characterize # noqa: F401
from .find_link import find_features                # noqa: F401
from .utils import RefineException                  # noqa: F401

__all__ = ["refine_leastsq", "find_clusters", "FitFunctions", "constraints", "find", "preprocessing",
           "preprocess", "characterize", "find_features", "RefineException"]
