"""Loader for the committed golden vectors (made by tests/golden/make_golden.py
from the reference itself)."""
import json
import os

import numpy as np

import synth

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(_DIR, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def block_input(m):
    x = synth.clip(m["clip"], m["n"], m["ci"], m["fs"])
    return np.ascontiguousarray(np.roll(x, m["roll"], axis=0))


def syn_input(m):
    x = synth.clip(m["clip"], m["n"], 2, m["fs"])
    return np.ascontiguousarray(np.roll(x, m["roll"], axis=0))
