"""TEST INFRASTRUCTURE ONLY -- loader for the real, unmodified reference package.

Only usable where ``/root/reference`` exists (the build container).  The GPU box
has no reference checkout, so nothing under ``tests -m gpu``, ``smoke()`` or
``bench.py`` may call :func:`load_reference`; those use the committed golden
fixtures (``tests/golden``) and the restatement in ``oracle/ilqr_oracle.py``.

The reference's pure-Python modules are imported from where they lie; its
Cython extension is taken from ``oracle/_ref`` (built by ``oracle/build_ref.sh``
from the reference's own sources).  ``matplotlib`` is absent in this image and
is only needed by the reference's plotting module, so it is stubbed
(SURVEY.md section 8c).
"""

import importlib.util
import os
import sys
from unittest import mock

REFERENCE_ROOT = os.environ.get("DPILQR_REFERENCE", "/root/reference")
_REF_NATIVE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "dpilqr", "__init__.py")) and any(
        f.startswith("bbdynamicswrap") and f.endswith(".so")
        for f in (os.listdir(_REF_NATIVE) if os.path.isdir(_REF_NATIVE) else [])
    )


def load_reference(name="dpilqr_reference"):
    """Import the reference as a package called ``name`` and return it."""
    if name in sys.modules:
        return sys.modules[name]
    if not reference_available():
        raise RuntimeError("reference checkout or oracle/_ref build not available")
    for mod in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation"):
        if mod not in sys.modules:
            try:
                importlib.import_module(mod)
            except Exception:
                sys.modules[mod] = mock.MagicMock()
    pkg_dir = os.path.join(REFERENCE_ROOT, "dpilqr")
    spec = importlib.util.spec_from_file_location(
        name,
        os.path.join(pkg_dir, "__init__.py"),
        submodule_search_locations=[pkg_dir, _REF_NATIVE],
    )
    module = importlib.util.module_from_spec(spec)
    sys.modules[name] = module
    # graphics.py does an absolute ``from dpilqr.util import ...``: alias the
    # package as ``dpilqr`` while it initialises, then put back whatever was
    # registered under that name (e.g. this repo's drop-in shim).
    saved = {k: v for k, v in sys.modules.items() if k == "dpilqr" or k.startswith("dpilqr.")}
    for k in saved:
        del sys.modules[k]
    sys.modules["dpilqr"] = module
    try:
        spec.loader.exec_module(module)
    except Exception:
        del sys.modules[name]
        raise
    finally:
        for k in [k for k in sys.modules if k == "dpilqr" or k.startswith("dpilqr.")]:
            if name != "dpilqr":
                sub = sys.modules.pop(k)
                if k != "dpilqr":
                    sys.modules.setdefault(name + k[len("dpilqr"):], sub)
        if name != "dpilqr":
            sys.modules.update(saved)
    return module
