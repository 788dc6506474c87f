"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Imports the *unmodified* reference modules from /root/reference so that golden
vectors can be generated and the numpy restatement (oracle/mof_oracle.py) can be
pinned against the real thing.  /root/reference exists only in the build
container (not on the GPU box), so nothing that runs under ``-m gpu``,
``smoke()`` or ``bench.py`` may import this file.

pyvista / matplotlib are imported at module top by the reference
(utils/compute_optical_flow.py:18, utils/find_singularity_point.py:14-15) but
are used only by loaders, plotting and the Jacobian classification, none of
which are on the hot path; they are absent offline and are stubbed here
(recipe of SURVEY.md section 8c).
"""
import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MOF_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "compute_optical_flow.py"))


def load():
    """-> (compute_optical_flow, find_singularity_point) reference modules."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in ("pyvista", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    # the reference package is called ``utils``; load it under a private alias so
    # it cannot collide with anything else on sys.path
    import importlib.util
    mods = []
    for short in ("compute_optical_flow", "find_singularity_point"):
        alias = f"_mof_reference_{short}"
        if alias in sys.modules:
            mods.append(sys.modules[alias])
            continue
        spec = importlib.util.spec_from_file_location(
            alias, os.path.join(REFERENCE_ROOT, "utils", short + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[alias] = mod
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


@contextlib.contextmanager
def quiet():
    """The reference prints per frame (compute_optical_flow.py:102); silence it."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
