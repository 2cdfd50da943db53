"""clustertracking_b200 -- B200-native implementation of clustertracking's ``refine_leastsq`` path.

Public names mirror the reference package (clustertracking/__init__.py:10-18) for the path this
project covers: ``refine_leastsq``, ``find_clusters``, ``FitFunctions``, ``constraints``.
"""
from .find import find_clusters                     # noqa: F401
from .fitfunc import FitFunctions                   # noqa: F401
from .refine import refine_leastsq                  # noqa: F401
from . import constraints                           # noqa: F401
from . import find                                  # noqa: F401
from .utils import RefineException                  # noqa: F401

__all__ = ["refine_leastsq", "find_clusters", "FitFunctions", "constraints", "find",
           "RefineException"]
