"""Loader for tests/golden fixtures (written by oracle/make_golden.py from the unmodified reference)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fixture_names():
    return sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLDEN_DIR, "g_*.json")))


_cache = {}


def load(name):
    if name not in _cache:
        with open(os.path.join(GOLDEN_DIR, name + ".json"), "r", encoding="utf-8") as f:
            fx = json.load(f)
        arr = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        fx["index_rows"] = arr["index_rows"]
        fx["query_vecs"] = arr["query_vecs"]
        _cache[name] = fx
    return _cache[name]


def known_answers():
    with open(os.path.join(GOLDEN_DIR, "known_answers.json"), "r", encoding="utf-8") as f:
        return json.load(f)


def canon(x):
    return json.dumps(x, sort_keys=True)
