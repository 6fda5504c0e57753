"""Import the reference's REAL classes from ``baseline/_ref`` (TEST INFRASTRUCTURE; see make_ref.py).

``import_real(torchdiffeq_module)`` returns fresh copies of the reference's ``models.blackbox_ode`` and
``models.decoders`` modules bound to the given ``torchdiffeq`` implementation -- the product
(``structured_latent_odes_b200.torchdiffeq_api``, i.e. what ``install_as_torchdiffeq()`` registers) or the CPU
oracle -- so that one test can hold both side by side.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import types

from . import make_ref


class Munch(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _munch_module():
    m = types.ModuleType("munch")
    m.Munch = Munch
    m.munchify = lambda d: Munch(d)
    return m


def import_real(torchdiffeq_module):
    """(blackbox_ode, decoders) of the unmodified reference, with ``import torchdiffeq`` resolving to the argument."""
    if not make_ref.available():
        raise FileNotFoundError(make_ref.DEST)
    saved = {k: sys.modules.get(k) for k in ("torchdiffeq", "munch", "models", "models.blackbox_ode", "models.decoders")}
    sys.modules["torchdiffeq"] = torchdiffeq_module
    sys.modules.setdefault("munch", _munch_module())
    for k in ("models", "models.blackbox_ode", "models.decoders"):
        sys.modules.pop(k, None)
    sys.path.insert(0, make_ref.DEST)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            bb = importlib.import_module("models.blackbox_ode")
            dec = importlib.import_module("models.decoders")
    finally:
        sys.path.remove(make_ref.DEST)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return bb, dec
