"""Helpers to load the committed reference fixtures (tests/golden/*.npz, made by oracle/gen_golden.py)."""
import hashlib
import os
import zlib

import numpy as np

from oracle.golden_cases import CASES
from streamoptima_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return [n for n in CASES if os.path.exists(os.path.join(GOLDEN_DIR, n + ".npz"))]


def load_case(name):
    """-> (frames u8 [F,H,W], encoder kwargs, golden dict)."""
    kind, gkw = CASES[name]["gen"]
    frames = synth.make(kind, **gkw)
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    assert hashlib.sha256(frames.tobytes()).hexdigest() == str(g["sha_input"]), "synthetic generator drifted"
    g["mv_text"] = zlib.decompress(g["mv_text_z"].tobytes()).decode()
    g["res_text"] = zlib.decompress(g["res_text_z"].tobytes()).decode()
    return frames, dict(CASES[name]["enc"]), g
