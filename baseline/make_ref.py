"""Recipe for ``baseline/_ref/``: an UNMODIFIED copy of the reference's hot-path sources (TEST INFRASTRUCTURE).

``/root/reference`` exists only in the build container.  The GPU box receives the working tree, so the three
source files the hot path consists of are copied -- byte for byte, never edited -- into ``baseline/_ref/``,
which is git-ignored (the reference's sources never enter this repository's history) but travels with the
``gpurun`` snapshot like the built ``.so``.  ``tests/test_gpu_real_reference.py`` imports the reference's REAL
``OdeModel`` / ``Decoder`` classes from there after ``install_as_torchdiffeq()`` and drives them on the B200.

    models/blackbox_ode.py   OdeModel / OdeFunc / Dynamics  (the hot path)
    models/decoders.py       Decoder / GaussianDecoder      (boundary consumer, a9)
    data/cvs/cvs_data.py     dx_dt + generator              (a10)

``pip install /root/reference`` is impossible (a directory of scripts: no setup.py / pyproject.toml), and its
imports need ``torchdiffeq`` / ``munch`` (absent, no network): ``baseline/ref_shims.py`` supplies a ``munch``
attribute-dict; ``torchdiffeq`` is EITHER this package (the product under test) OR the CPU oracle.

Run by ``__graft_entry__.build()`` whenever ``/root/reference`` is present; a SHA-256 manifest is written next to
the copies so that a test can prove they are unmodified.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("SLODE_REFERENCE_ROOT", "/root/reference")
DEST = os.path.join(HERE, "_ref")
FILES = ["models/__init__.py", "models/blackbox_ode.py", "models/decoders.py",
         "data/__init__.py", "data/cvs/__init__.py", "data/cvs/cvs_data.py", "utils/__init__.py", "utils/utils.py"]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def make(verbose=False):
    """Copies the files if the reference checkout is present; returns DEST or None."""
    if not os.path.isfile(os.path.join(REF_ROOT, "models", "blackbox_ode.py")):
        return DEST if os.path.isfile(os.path.join(DEST, "models", "blackbox_ode.py")) else None
    manifest = {}
    for rel in FILES:
        src = os.path.join(REF_ROOT, rel)
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.isfile(src):
            shutil.copyfile(src, dst)
            manifest[rel] = _sha(dst)
        elif rel.endswith("__init__.py"):
            open(dst, "a").close()  # namespace marker only (the reference has none there)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_ROOT, "sha256": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"baseline/_ref: {len(manifest)} reference files copied unmodified from {REF_ROOT}")
    return DEST


def available() -> bool:
    return os.path.isfile(os.path.join(DEST, "models", "blackbox_ode.py"))


if __name__ == "__main__":
    print(make(verbose=True))
