"""Loader for the outputs of the Go parity dumper (baseline/go/README.md): tests/golden/from_go/<case>/*.npy plus the
inputs and manifest written by tests/golden/make_go_inputs.py."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FROM_GO = os.path.join(GOLDEN, "from_go")


def manifest():
    path = os.path.join(FROM_GO, "inputs", "manifest.json")
    if not os.path.exists(path):
        return []
    with open(path) as f:
        return json.load(f)


def have_go_outputs():
    return any(os.path.isdir(os.path.join(FROM_GO, c["name"])) and os.listdir(os.path.join(FROM_GO, c["name"]))
               for c in manifest())


def read_input(name):
    return np.fromfile(os.path.join(FROM_GO, "inputs", name), dtype="<f8")


def load_case(name):
    d = os.path.join(FROM_GO, name)
    return {f[:-4]: np.load(os.path.join(d, f)) for f in sorted(os.listdir(d)) if f.endswith(".npy")}
