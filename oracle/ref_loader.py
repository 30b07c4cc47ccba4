"""Load the UNMODIFIED reference (jjrreett/fea) from /root/reference under stubbed plot modules.

TEST INFRASTRUCTURE ONLY.  Nothing in the product (`fea_b200/`) may import this file.
It only works in the build container: /root/reference does not exist on the GPU box, so
nothing in the `-m gpu` tests, `smoke()` or `bench.py` may call it.  Its single purpose is
to pin `oracle/fea_oracle.py` to the real reference (tests/test_oracle_vs_reference.py) and
to generate the committed fixtures in tests/golden/ (oracle/make_golden.py).

Method: the reference modules import matplotlib / pyvista / turtle at
the top and run a solve plus a blocking plot at import time (no ``__main__`` guard:
cubebeam.py:60-66,111-124,233-245).  We pre-seed ``sys.modules`` with MagicMock plot
modules, put the reference directory on ``sys.path``, then either ``import utils`` (pure
functions) or ``runpy.run_path`` a script and read its globals.  ``truss.py`` never
terminates (truss.py:97 ``while True``) so only its prefix up to truss.py:95 is executed.
"""
from __future__ import annotations

import os
import runpy
import sys
from contextlib import contextmanager
from unittest.mock import MagicMock

REFERENCE_DIR = os.environ.get("FEA_REFERENCE_DIR", "/root/reference")

_STUBS = (
    "pyvista",
    "matplotlib",
    "matplotlib.pyplot",
    "mpl_toolkits",
    "mpl_toolkits.mplot3d",
    "mpl_toolkits.mplot3d.art3d",
    "turtle",
    "pythreejs",
)


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "utils.py"))


@contextmanager
def _reference_env():
    """sys.modules / sys.path patched so the reference imports; restored afterwards."""
    saved = {name: sys.modules.get(name) for name in _STUBS + ("utils",)}
    for name in _STUBS:
        sys.modules[name] = MagicMock()
    sys.modules.pop("utils", None)
    sys.path.insert(0, REFERENCE_DIR)
    try:
        yield
    finally:
        sys.path.remove(REFERENCE_DIR)
        for name, mod in saved.items():
            if mod is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = mod


def load_utils():
    """The reference's ``utils`` module object (utils.py:127 hexahedral_stiffness_matrix,
    utils.py:356 stack_faces_2d, utils.py:379/390 face helpers)."""
    if not available():
        raise FileNotFoundError(f"reference not found at {REFERENCE_DIR}")
    with _reference_env():
        import importlib

        mod = importlib.import_module("utils")
        sys.modules.pop("utils", None)
        return mod


def run_script(name: str) -> dict:
    """Run ``cubebeam.py`` / ``fea.py`` / ``euler_bernoulli.py`` verbatim, return its globals."""
    if not available():
        raise FileNotFoundError(f"reference not found at {REFERENCE_DIR}")
    with _reference_env():
        import io
        from contextlib import redirect_stdout

        with redirect_stdout(io.StringIO()):
            return runpy.run_path(os.path.join(REFERENCE_DIR, name))


def load_truss_prefix() -> dict:
    """Execute truss.py up to (not including) its endless loop (truss.py:97)."""
    if not available():
        raise FileNotFoundError(f"reference not found at {REFERENCE_DIR}")
    with open(os.path.join(REFERENCE_DIR, "truss.py")) as fh:
        src = fh.read()
    cut = src.index("while True:")
    glb: dict = {"__name__": "reference_truss"}
    with _reference_env():
        exec(compile(src[:cut], "truss.py", "exec"), glb)
    return glb
