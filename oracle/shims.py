"""Import the *real* reference modules in the build container (TEST INFRASTRUCTURE).

The reference needs ``torchdiffeq`` and ``munch``, neither of which is installed.  This helper
registers ``oracle.torchdiffeq_oracle`` under the name ``torchdiffeq`` and a minimal attribute
dict under ``munch`` and puts ``/root/reference`` on ``sys.path`` so that
``models/blackbox_ode.py``, ``models/decoders.py`` and ``data/cvs/cvs_data.py`` import UNCHANGED.
``/root/reference`` only exists in the build container; callers must check
``reference_available()`` first (the GPU box never has it).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SLODE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "blackbox_ode.py"))


class _Munch(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def install():
    from . import torchdiffeq_oracle

    sys.modules.setdefault("torchdiffeq", torchdiffeq_oracle)
    if "munch" not in sys.modules:
        m = types.ModuleType("munch")
        m.Munch = _Munch
        m.munchify = lambda d: _Munch(d)
        sys.modules["munch"] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def import_reference_blackbox():
    """Returns the reference's (blackbox_ode, decoders) modules, imported unchanged."""
    if not reference_available():
        raise FileNotFoundError(REFERENCE_ROOT)
    install()
    with contextlib.redirect_stdout(io.StringIO()):
        import models.blackbox_ode as bb  # type: ignore
        import models.decoders as dec  # type: ignore
    return bb, dec


def import_reference_cvs():
    if not reference_available():
        raise FileNotFoundError(REFERENCE_ROOT)
    install()
    import data.cvs.cvs_data as cvs  # type: ignore
    return cvs
