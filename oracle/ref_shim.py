"""
oracle/ref_shim.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE. Build container only.

Executes the reference's UNMODIFIED hot-path modules straight from
``/root/reference/src/harmonica`` (read-only; nothing is copied) with the
restated choclo kernels of ``oracle/choclo_numba.py`` injected as a fake
``choclo`` package, and with inert stand-ins for the heavyweight imports the
hot path never touches (verde, xarray, the pyvista helper).

Used to (1) check oracle/choclo_port.c against the reference's own wrappers and
jitted loops and (2) generate ``tests/golden/*.npz`` (oracle/make_golden.py).
``/root/reference`` does not exist on the GPU box, so nothing that runs there
imports this module; ``available()`` says whether it can be used.

Modules loaded (reference file):
  harmonica._forward.utils            src/harmonica/_forward/utils.py
  harmonica._forward.prisms.utils     src/harmonica/_forward/prisms/utils.py
  harmonica._forward.prisms.gravity   src/harmonica/_forward/prisms/gravity.py
  harmonica._forward.prisms.magnetic  src/harmonica/_forward/prisms/magnetic.py
  harmonica._forward.prisms.layer     src/harmonica/_forward/prisms/layer.py
  harmonica._forward.point            src/harmonica/_forward/point.py
  harmonica._equivalent_sources.utils src/harmonica/_equivalent_sources/utils.py
  harmonica._forward.dipole           src/harmonica/_forward/dipole.py
  harmonica._forward._tesseroid_utils src/harmonica/_forward/_tesseroid_utils.py
  harmonica._forward.tesseroid_gravity src/harmonica/_forward/tesseroid_gravity.py
      (which also pulls in _tesseroid_variable_density.py)
  load_ellipsoids(): harmonica._forward.ellipsoids.{ellipsoids,magnetic,utils}
      src/harmonica/_forward/ellipsoids/*.py -- NOT on the hot path and free of choclo (numpy +
      scipy, Clark 1986 / Takahashi 2018): the external field of a uniformly magnetised SPHERE is
      exactly a dipole's, which makes the reference's own ellipsoid code an independent check of
      the absolute values, units and signs of dipole_magnetic (oracle/make_golden_ellipsoid.py)
"""

import importlib
import os
import sys
import types

REFERENCE_SRC = "/root/reference/src/harmonica"


def available():
    return os.path.isdir(REFERENCE_SRC)


def _pkg(name, path):
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    mod.__package__ = name
    sys.modules[name] = mod
    return mod


_loaded = {}


def load():
    """Return a namespace with the reference's unmodified hot-path modules."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"{REFERENCE_SRC} is not mounted (GPU box?)")
    if "harmonica" in sys.modules and not getattr(sys.modules["harmonica"], "__hb200_shim__", False):
        raise RuntimeError("a real 'harmonica' is already imported; the shim is not needed")
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import choclo_numba

    choclo_numba.install_fake_choclo()

    hm = _pkg("harmonica", REFERENCE_SRC)
    hm.__hb200_shim__ = True
    _pkg("harmonica._forward", os.path.join(REFERENCE_SRC, "_forward"))
    _pkg("harmonica._forward.prisms", os.path.join(REFERENCE_SRC, "_forward", "prisms"))
    _pkg("harmonica._equivalent_sources", os.path.join(REFERENCE_SRC, "_equivalent_sources"))
    # inert stand-ins for imports that layer.py performs at module load
    # (layer.py:15-18); none of them is reached by _forward_gravity_prism_layer.
    vis = types.ModuleType("harmonica.visualization")
    vis.prism_to_pyvista = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    sys.modules["harmonica.visualization"] = vis
    if "verde" not in sys.modules:
        try:
            importlib.import_module("verde")
        except ImportError:
            sys.modules["verde"] = types.ModuleType("verde")
    if "xarray" not in sys.modules:
        try:
            importlib.import_module("xarray")
        except ImportError:
            xr = types.ModuleType("xarray")
            xr.register_dataset_accessor = lambda name: (lambda cls: cls)
            sys.modules["xarray"] = xr

    _loaded["utils"] = importlib.import_module("harmonica._forward.utils")
    _loaded["prism_utils"] = importlib.import_module("harmonica._forward.prisms.utils")
    _loaded["gravity"] = importlib.import_module("harmonica._forward.prisms.gravity")
    _loaded["magnetic"] = importlib.import_module("harmonica._forward.prisms.magnetic")
    _loaded["layer"] = importlib.import_module("harmonica._forward.prisms.layer")
    _loaded["point"] = importlib.import_module("harmonica._forward.point")
    _loaded["eqs_utils"] = importlib.import_module("harmonica._equivalent_sources.utils")
    _loaded["dipole"] = importlib.import_module("harmonica._forward.dipole")
    _loaded["tesseroid_utils"] = importlib.import_module("harmonica._forward._tesseroid_utils")
    _loaded["tesseroid"] = importlib.import_module("harmonica._forward.tesseroid_gravity")
    _loaded["tesseroid_variable_density"] = importlib.import_module(
        "harmonica._forward._tesseroid_variable_density"
    )
    return types.SimpleNamespace(**_loaded)


def load_ellipsoids():
    """The reference's unmodified ellipsoid modules (ellipsoids.py, magnetic.py, utils.py)."""
    load()
    if "harmonica._forward.ellipsoids" not in sys.modules:
        _pkg("harmonica._forward.ellipsoids", os.path.join(REFERENCE_SRC, "_forward", "ellipsoids"))
    return types.SimpleNamespace(
        ellipsoids=importlib.import_module("harmonica._forward.ellipsoids.ellipsoids"),
        magnetic=importlib.import_module("harmonica._forward.ellipsoids.magnetic"),
        gravity=importlib.import_module("harmonica._forward.ellipsoids.gravity"),
    )


def greens_func_cartesian():
    """The 3-line Green's function of cartesian.py:634-644, re-typed (that file
    imports verde/bordado at module load and cannot be executed here); it calls
    the reference's own ``distance_cartesian``."""
    from numba import jit

    ref = load()
    distance_cartesian = ref.utils.distance_cartesian

    @jit(nopython=True)
    def greens(east, north, upward, point_east, point_north, point_upward):
        distance = distance_cartesian((east, north, upward), (point_east, point_north, point_upward))
        return 1 / distance

    return greens


def greens_func_spherical():
    """The Green's function of spherical.py:412-424, re-typed for the same reason; it calls the
    reference's own ``distance_spherical``."""
    from numba import jit

    ref = load()
    distance_spherical = ref.utils.distance_spherical

    @jit(nopython=True)
    def greens(longitude, latitude, radius, point_longitude, point_latitude, point_radius):
        distance = distance_spherical(
            (longitude, latitude, radius), (point_longitude, point_latitude, point_radius)
        )
        return 1 / distance

    return greens
