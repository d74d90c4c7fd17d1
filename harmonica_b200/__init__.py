"""
harmonica_b200: B200-native drop-in for harmonica's pairwise forward models.

The O(N_observers x N_sources) hot path behind ``prism_gravity``,
``prism_magnetic``, ``point_gravity``, ``Dataset.prism_layer.gravity()`` and
``EquivalentSources.predict()`` runs as hand-written CUDA (sm_100a) in
``libharmonica_b200.so``; this package is the host side: the reference's
function signatures, checks, messages, units and sign conventions over a thin
ctypes binding. There is no CPU fallback.
"""

from ._dipole import dipole_magnetic
from ._eqs import (
    EquivalentSources,
    EquivalentSourcesGB,
    EquivalentSourcesSph,
    eqs_fit,
    eqs_fit_gradient_boosted,
    eqs_jacobian,
    eqs_jacobian_spherical,
    eqs_predict,
    predict_numba_parallel,
    predict_numba_serial,
)
from ._lib import HarmonicaB200Error, init
from ._point import point_gravity
from ._prism_gravity import prism_gravity
from ._prism_layer import DatasetAccessorPrismLayer, PrismLayer, prism_layer, prism_layer_gravity
from ._prism_magnetic import prism_magnetic
from ._tesseroid import tesseroid_gravity
from ._tesseroid_layer import DatasetAccessorTesseroidLayer, TesseroidLayer, tesseroid_layer

__version__ = "0.1.0"

__all__ = [
    "DatasetAccessorPrismLayer",
    "EquivalentSources",
    "EquivalentSourcesGB",
    "EquivalentSourcesSph",
    "dipole_magnetic",
    "HarmonicaB200Error",
    "PrismLayer",
    "eqs_fit",
    "eqs_fit_gradient_boosted",
    "eqs_jacobian",
    "eqs_jacobian_spherical",
    "eqs_predict",
    "init",
    "point_gravity",
    "predict_numba_parallel",
    "predict_numba_serial",
    "prism_gravity",
    "prism_layer",
    "prism_layer_gravity",
    "prism_magnetic",
    "tesseroid_gravity",
    "tesseroid_layer",
    "TesseroidLayer",
    "DatasetAccessorTesseroidLayer",
]
