"""Compiles the UNMODIFIED reference into ``oracle/_ref/`` (git-ignored byte code, no sources).  TEST INFRASTRUCTURE ONLY.

The reference (Suyashagarw/StreamOptima) is pure Python: its "build" is ``py_compile`` of the three modules of the path,
from the sources where they lie under ``/root/reference`` -- nothing is copied into the repository, only byte-code outputs are
written, and only here.  The directory travels to the GPU box like any built artefact (same image, same interpreter), where
``bench.py``'s CPU arm times the reference's own ``Y_Video_codec.encode()`` on the box's host cores
(``cpu_baseline.kind = "reference"``).  Tests never depend on it.  ``/root/reference`` only exists in the build container, so
this is a no-op anywhere else.
"""
from __future__ import annotations

import hashlib
import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
REFERENCE_DIR = "/root/reference"
MODULES = ("Encoder", "decoder", "video_manager")
EXT = ".pyref"          # CPython byte code (what py_compile writes); not named .pyc, which snapshot tools tend to skip


def build_ref(force: bool = False):
    """-> path of oracle/_ref, or None when the reference sources are not available here."""
    if not os.path.isfile(os.path.join(REFERENCE_DIR, "Encoder.py")):
        return OUT_DIR if os.path.isfile(os.path.join(OUT_DIR, "Encoder" + EXT)) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    manifest = {"python": sys.version.split()[0], "magic": py_compile.importlib.util.MAGIC_NUMBER.hex(), "modules": {}}
    for m in MODULES:
        src, dst = os.path.join(REFERENCE_DIR, m + ".py"), os.path.join(OUT_DIR, m + EXT)
        if force or not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst):
            py_compile.compile(src, cfile=dst, dfile=m + ".py", doraise=True, optimize=0,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        manifest["modules"][m] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(OUT_DIR, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return OUT_DIR


def load_compiled_reference():
    """``(Encoder_module, decoder_module)`` of the unmodified reference from ``oracle/_ref`` byte code, or ``None`` when it is
    absent or was compiled by another interpreter.  The stubs for the two absent plotting / metrics packages are the ones of
    ``oracle/reference_harness.py`` (no arithmetic of the path goes through them)."""
    import importlib.machinery
    import importlib.util
    if not os.path.isfile(os.path.join(OUT_DIR, "Encoder" + EXT)):
        return None
    try:
        man = json.load(open(os.path.join(OUT_DIR, "MANIFEST.json")))
        if man.get("magic") != importlib.util.MAGIC_NUMBER.hex():
            return None
    except Exception:
        return None
    from oracle import reference_harness as rh
    rh._install_stubs()
    mods = {}
    for m in ("decoder", "video_manager", "Encoder"):
        path = os.path.join(OUT_DIR, m + EXT)
        loader = importlib.machinery.SourcelessFileLoader(m, path)
        spec = importlib.util.spec_from_loader(m, loader)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[m] = mod                    # Encoder imports `decoder` by name (Encoder.py:15)
        loader.exec_module(mod)
        mods[m] = mod
    return mods["Encoder"], mods["decoder"]


if __name__ == "__main__":
    print(build_ref(True))
