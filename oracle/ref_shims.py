"""ORACLE (test infrastructure).  Runs the reference's OWN source over stand-ins for its absent dependencies.

The reference's in-tree code (cellregmap/_cellregmap.py, _math.py, _simulate.py) is pure Python, but it imports six
third-party packages that are neither vendored under /root/reference nor installed in this image nor installable
(no network): glimix_core, numpy_sugar, chiscore (-> chi2comb), and -- inside compute_maf only -- dask and xarray.
`install()` registers stand-in modules for exactly the names the reference imports, backed by the restatements of this
package (lmm_port, sugar_port, chiscore_port); `load_reference()` then imports the unmodified reference package:

  * from `oracle/_ref/cellregmap/*.bin` (byte code), compiled by `oracle/build_ref.py` from the sources where they lie under
    /root/reference (compiled outputs only -- no reference source is copied; oracle/_ref is git-ignored and travels to
    the GPU box like any other built checker), or
  * from /root/reference itself when that directory exists (this container).

What this pins: every line of the reference's own logic on the path -- constructor and rho1 grid (:63-131), the SNP
loop, strict-'>' selection, permutation hooks, QSCov/PMat/ScoreStatistic, `run_association`'s positional quirk,
`predict_interaction`'s BLUP and output shapes, `lrt_pvalues`, `get_L_values`, `compute_maf`, and the simulator that
generates configs[0].  What it does NOT pin: the internals of the third-party packages (LMM likelihood and Brent
iterate path, economic_qs_linear/economic_svd, davies_pvalue/qfc, liu_sf) -- those are the stand-ins themselves.
`liu_sf` is pinned separately by the reference's known answers (tests/test_oracle_goldens.py).
"""
import importlib
import os
import sys
import types

import numpy as np

from . import chiscore_port, lmm_port, sugar_port

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_BUILD = os.path.join(_HERE, "_ref")
REF_SOURCE = "/root/reference"
SHIMMED = ("glimix_core", "glimix_core.lmm", "numpy_sugar", "numpy_sugar.linalg", "chiscore", "dask", "dask.array", "xarray")


def _module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__oracle_shim__ = True
    return mod


class _Epsilon:
    """numpy_sugar.epsilon"""
    small = sugar_port.EPS_SMALL
    tiny = sugar_port.EPS_TINY
    super_tiny = sugar_port.EPS_SUPER_TINY


def install(qs_method="svd"):
    """Register the stand-in modules (idempotent).  A real installation of any of the packages is never shadowed:
    if `glimix_core` imports, the reference runs on the real thing and the caller can tell from `is_shimmed()`.

    `qs_method="gram"` swaps economic_qs_linear's thin SVD for the eigendecomposition of the m x m Gram (same Q0, S0
    up to round-off, see crm_port._qs_via_gram); bench.py uses it to bound the set-up time of the CPU arm at n = 1e5."""
    def economic_qs_linear(G, return_q1=True):
        if qs_method == "gram" and not return_q1 and G.shape[0] > G.shape[1]:
            from .crm_port import _qs_via_gram
            return _qs_via_gram(np.asarray(G, float))
        return sugar_port.economic_qs_linear(G, return_q1=return_q1)

    try:
        import glimix_core  # noqa: F401
        if not getattr(glimix_core, "__oracle_shim__", False):
            return False
    except ImportError:
        pass
    linalg = _module("numpy_sugar.linalg", economic_qs_linear=economic_qs_linear, economic_svd=sugar_port.economic_svd,
                     rsolve=sugar_port.rsolve, economic_qs=sugar_port.economic_qs)
    sugar = _module("numpy_sugar", ddot=sugar_port.ddot, epsilon=_Epsilon, linalg=linalg)
    lmm = _module("glimix_core.lmm", LMM=lmm_port.LMM, FastScanner=lmm_port.FastScanner)
    glimix = _module("glimix_core", lmm=lmm)
    glimix.__path__ = []
    sugar.__path__ = []
    chiscore = _module("chiscore", davies_pvalue=chiscore_port.davies_pvalue, liu_sf=chiscore_port.liu_sf)
    # compute_maf imports these unconditionally (:613-614) and only uses them in isinstance checks
    dask_array = _module("dask.array", Array=type("Array", (), {}))
    dask = _module("dask", array=dask_array)
    dask.__path__ = []
    xarray = _module("xarray", DataArray=type("DataArray", (), {}))
    mods = {"numpy_sugar": sugar, "numpy_sugar.linalg": linalg, "glimix_core": glimix, "glimix_core.lmm": lmm, "chiscore": chiscore}
    for name, mod in (("dask", dask), ("dask.array", dask_array), ("xarray", xarray)):
        try:
            importlib.import_module(name)
        except ImportError:
            mods[name] = mod
    sys.modules.update(mods)
    return True


def is_shimmed():
    return bool(getattr(sys.modules.get("glimix_core"), "__oracle_shim__", False))


def available():
    return os.path.isdir(os.path.join(REF_SOURCE, "cellregmap")) or os.path.exists(os.path.join(REF_BUILD, "cellregmap", "__init__.bin"))


def _load_compiled(root):
    """Import the byte-compiled package oracle/_ref/cellregmap (files *.bin = .pyc content) without any source."""
    import importlib.machinery
    import importlib.util
    pkg_dir = os.path.join(root, "cellregmap")

    def load(name, fname, is_pkg=False):
        loader = importlib.machinery.SourcelessFileLoader(name, os.path.join(pkg_dir, fname))
        spec = importlib.util.spec_from_loader(name, loader, is_package=is_pkg)
        mod = importlib.util.module_from_spec(spec)
        if is_pkg:
            mod.__path__ = [pkg_dir]
        sys.modules[name] = mod
        try:
            loader.exec_module(mod)
        except BaseException:
            sys.modules.pop(name, None)
            raise
        return mod

    # submodules first (the package's __init__ imports from them), in dependency order
    pkg = types.ModuleType("cellregmap")
    pkg.__path__ = [pkg_dir]
    pkg.__package__ = "cellregmap"
    sys.modules["cellregmap"] = pkg
    try:
        for sub in ("_types", "_math", "_cellregmap", "_simulate"):
            setattr(pkg, sub, load("cellregmap." + sub, sub + ".bin"))
        init = importlib.machinery.SourcelessFileLoader("cellregmap", os.path.join(pkg_dir, "__init__.bin"))
        exec(init.get_code("cellregmap"), pkg.__dict__)
    except BaseException:
        for name in [n for n in sys.modules if n == "cellregmap" or n.startswith("cellregmap.")]:
            sys.modules.pop(name, None)
        raise
    return pkg


def load_reference(qs_method="svd", prefer_source=True):
    """The reference package (module `cellregmap`, plus `cellregmap._simulate`), imported unmodified over the stand-ins.
    Returns None when neither /root/reference nor oracle/_ref is present."""
    install(qs_method=qs_method)
    cached = sys.modules.get("cellregmap")
    if cached is not None and getattr(cached, "__oracle_loaded__", False):
        return cached
    if prefer_source and os.path.isdir(os.path.join(REF_SOURCE, "cellregmap")):
        root = REF_SOURCE
    elif os.path.exists(os.path.join(REF_BUILD, "cellregmap", "__init__.bin")):
        root = REF_BUILD
    else:
        return None
    if root == REF_BUILD:
        mod = _load_compiled(root)
    else:
        sys.path.insert(0, root)
        try:
            sys.dont_write_bytecode, saved = True, sys.dont_write_bytecode     # /root/reference is read-only
            mod = importlib.import_module("cellregmap")
            importlib.import_module("cellregmap._simulate")
        finally:
            sys.dont_write_bytecode = saved
            sys.path.remove(root)
    mod.__oracle_loaded__ = True
    mod.__oracle_root__ = root
    return mod
