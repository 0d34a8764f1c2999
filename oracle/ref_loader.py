"""Import the reference (amspector100/pyblip) for checking purposes only.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this.

``load()`` registers a synthetic ``pyblip`` package in ``sys.modules`` whose
sub-package search paths are

  1. ``oracle/_ref/pyblip/...``   the reference's Cython kernels compiled by
     ``oracle/build_ref.py`` (binaries only; present here and on the GPU box)
  2. ``/root/reference/pyblip/...``  the reference's pure-Python modules
     (``create_groups.py``, ``linear/linear.py`` ...; present in the build
     container only, never on the GPU box)

The reference's own ``pyblip/__init__.py`` is *not* executed because it imports
``blip.py`` -> cvxpy, which this image does not have.
A stub ``cvxpy`` module is installed only if some caller later imports
``pyblip.blip`` explicitly.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref", "pyblip")
REF_SRC = os.path.join(os.environ.get("PYBLIP_REFERENCE", "/root/reference"), "pyblip")

_SUBPACKAGES = ["cython_utils", "linear", "nprior", "probit"]


def have_compiled():
    return os.path.isdir(os.path.join(REF_BIN, "linear")) and any(
        f.startswith("_linear.") and f.endswith(".so")
        for f in os.listdir(os.path.join(REF_BIN, "linear")))


def have_python_sources():
    return os.path.isfile(os.path.join(REF_SRC, "create_groups.py"))


def _mk_pkg(name, paths):
    mod = types.ModuleType(name)
    mod.__path__ = [p for p in paths if os.path.isdir(p)]
    mod.__package__ = name
    sys.modules[name] = mod
    return mod


def _install_cvxpy_stub():
    if "cvxpy" in sys.modules:
        return
    try:
        import cvxpy  # noqa: F401
        return
    except Exception:
        pass
    cp = types.ModuleType("cvxpy")
    red = types.ModuleType("cvxpy.reductions")
    sol = types.ModuleType("cvxpy.reductions.solvers")
    defs = types.ModuleType("cvxpy.reductions.solvers.defines")
    defs.INSTALLED_SOLVERS = []
    sol.defines = defs
    red.solvers = sol
    cp.reductions = red
    for n, m in [("cvxpy", cp), ("cvxpy.reductions", red),
                 ("cvxpy.reductions.solvers", sol),
                 ("cvxpy.reductions.solvers.defines", defs)]:
        sys.modules[n] = m


def load():
    """Return the synthetic ``pyblip`` package (idempotent)."""
    existing = sys.modules.get("pyblip")
    if existing is not None and getattr(existing, "_b200_oracle_loader", False):
        return existing
    if not have_compiled():
        raise ImportError(
            "oracle/_ref is not built: run `python oracle/build_ref.py` in the "
            "build container (needs /root/reference)")
    root = _mk_pkg("pyblip", [REF_BIN, REF_SRC])
    root._b200_oracle_loader = True
    for sub in _SUBPACKAGES:
        m = _mk_pkg("pyblip." + sub,
                    [os.path.join(REF_BIN, sub), os.path.join(REF_SRC, sub)])
        setattr(root, sub, m)
    # compiled kernels (always available once built)
    for name in ["pyblip.cython_utils._truncnorm",
                 "pyblip.cython_utils._update_hparams",
                 "pyblip.linear._linear", "pyblip.linear._linear_multi",
                 "pyblip.nprior._nprior"]:
        importlib.import_module(name)
    return root


def load_python(modname):
    """Import one of the reference's pure-Python modules, e.g. 'create_groups'.
    Only works where /root/reference exists (fixture generation)."""
    load()
    if not have_python_sources():
        raise ImportError("reference Python sources are not present on this box")
    if modname in ("blip",):
        _install_cvxpy_stub()
    return importlib.import_module("pyblip." + modname)
