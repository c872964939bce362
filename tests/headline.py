"""Shared helpers for the headline-shape parity checks (tests/test_headline.py, tests/check_sharded.py, bench.py's parity
block): the committed goldens under tests/golden/headline_*.json were produced by the C oracle on the bench's own synthetic
input (tests/golden/make_headline_golden.py).  Nothing here touches oracle/ — the fixtures are plain JSON."""
import hashlib
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).reshape(-1).data).hexdigest()


def load_all():
    out = {}
    for name in sorted(os.listdir(GOLDEN_DIR)):
        if name.startswith("headline_") and name.endswith(".json"):
            out[name[len("headline_"):-5]] = json.load(open(os.path.join(GOLDEN_DIR, name)))
    return out


def find(log_n, n_cols, rate_bits, cap_height, seed=1):
    """the golden for this shape and input seed, or None"""
    for name, g in load_all().items():
        s = g["shape"]
        if (s["log_n"], s["n_cols"], s["rate_bits"], s["cap_height"], g["seed"]) == (log_n, n_cols, rate_bits, cap_height, seed):
            return name, g
    return None, None


def parity_block(cap, log_n, n_cols, rate_bits, cap_height, seed=1):
    """bench.py's `parity` object: sha256 of the cap this run produced and whether it equals the oracle's golden cap."""
    cap = np.ascontiguousarray(cap, dtype=np.uint64).reshape(-1, 4)
    name, g = find(log_n, n_cols, rate_bits, cap_height, seed)
    out = {"cap_sha256": sha(cap), "golden": None, "match": None}
    if g is not None:
        out["golden"] = f"tests/golden/headline_{name}.json (oracle/gl_oracle.c on the same SplitMix64 input)"
        out["match"] = bool(out["cap_sha256"] == g["sha256_cap"] and np.array_equal(cap, np.array(g["cap"], dtype=np.uint64)))
    return out
