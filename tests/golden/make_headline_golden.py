#!/usr/bin/env python3
"""Golden fixtures for the HEADLINE shapes (BASELINE.json configs[2] = 2^20 x 135 r=3 h=4, configs[4] = 2^22 x 256 r=1 h=4, and
configs[1] = 2^16 x 135), produced by the C oracle (oracle/gl_oracle.c) on the bench's own synthetic input
(SplitMix64, seed 1 — tests/oracle_c.py · splitmix_columns == bench.py · synth_columns).

Parity status: the oracle is a restatement (no Rust toolchain here), so these vectors are *regression anchors produced by the
restatement*, pinned to the reference only through the Poseidon KATs and the field constants ("parity unpinned" otherwise —
DESIGN.md §5).  What they buy: the 2^20 / 2^22 GPU plans (two 10-stage passes with the inter-pass twiddle; the 8+7+7 three-pass
plan) are compared word-for-word with an independent radix-2 CPU implementation at the full headline size, and every bench line
can carry a cap check.

    python tests/golden/make_headline_golden.py cfg2 cfg3            # ~2 min on 8 cores, ~21 GB RAM
    python tests/golden/make_headline_golden.py cfg5r1               # streamed per column, ~10 min, ~20 GB RAM

Stored per shape: the Merkle cap, sha256 of the cap / of the coefficients (column-major [C][N] little-endian u64) / of the
digests vector (plonky2 layout), 64 sampled leaf rows, and sha256 of each cap subtree's leaf rows (what one rank of the sharded
commit owns).
"""
import hashlib
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from oracle_c import OracleC, splitmix_columns  # noqa: E402

SHAPES = {"cfg2": (16, 135, 3, 4), "cfg3": (20, 135, 3, 4), "cfg5r1": (22, 256, 1, 4), "cfg2n": (16, 20, 3, 4)}
SEED = 1


def sample_rows(R):
    """64 leaf indices: the corners plus a fixed LCG walk (same list on the GPU side: tests/headline.py)."""
    idx = [0, 1, R - 1, R // 2, R // 2 - 1, R // 16, R // 16 - 1, 3]
    x = 0x9E3779B97F4A7C15
    while len(idx) < 64:
        x = (x * 6364136223846793005 + 1442695040888963407) & (2**64 - 1)
        idx.append((x >> 20) % R)
    return idx


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).reshape(-1).data).hexdigest()


def bitrev_perm(bits):
    i = np.arange(1 << bits, dtype=np.uint64)
    r = np.zeros_like(i)
    for b in range(bits):
        r |= ((i >> np.uint64(b)) & np.uint64(1)) << np.uint64(bits - 1 - b)
    return r.astype(np.int64)


def make(name, oc):
    log_n, n_cols, r, h = SHAPES[name]
    n, R = 1 << log_n, 1 << (log_n + r)
    t0 = time.time()
    if name != "cfg5r1":
        x = splitmix_columns(SEED, n_cols, n)
        res = oc.commit(x, r, h)
        coeffs, leaves, digests, cap = res["coeffs"], res["leaves"], res["digests"], res["cap"]
    else:
        # streamed: one column at a time through the oracle's ifft / coset_fft (zero-padded size-R transform, which is what
        # upstream's lde + coset_fft_with_options computes), leaves assembled row-major, then the oracle's MerkleTree::new
        x = splitmix_columns(SEED, n_cols, n)
        leaves = np.zeros((R, n_cols), dtype=np.uint64)
        coeffs = np.zeros((n_cols, n), dtype=np.uint64)
        perm = bitrev_perm(log_n + r)

        def one(c):
            co = oc.ifft(x[c])
            coeffs[c] = co
            pad = np.zeros(R, dtype=np.uint64)
            pad[:n] = co
            lde = oc.coset_fft(pad, 7)
            return c, lde[perm]

        with ThreadPoolExecutor(8) as ex:
            for c, col in ex.map(one, range(n_cols)):
                leaves[:, c] = col
        del x
        digests, cap = oc.merkle_new(leaves, h)
    rows = sample_rows(R)
    sub = R >> h
    out = {"shape": {"log_n": log_n, "n_cols": n_cols, "rate_bits": r, "cap_height": h}, "seed": SEED,
           "input": "tests/oracle_c.py · splitmix_columns(seed, n_cols, 2^log_n) (canonical)",
           "generator": "tests/golden/make_headline_golden.py (oracle/gl_oracle.c; parity unpinned beyond Poseidon KATs + field constants)",
           "cap": [[int(v) for v in row] for row in cap],
           "sha256_cap": sha(cap), "sha256_coeffs": sha(coeffs), "sha256_digests": sha(digests),
           "sha256_leaves": sha(leaves),
           "sha256_leaves_per_subtree": [sha(leaves[t * sub:(t + 1) * sub]) for t in range(1 << h)],
           "sample_rows": rows, "sample_leaves": [[int(v) for v in leaves[i]] for i in rows],
           "oracle_seconds": round(time.time() - t0, 1)}
    path = os.path.join(HERE, f"headline_{name}.json")
    json.dump(out, open(path, "w"), indent=0, separators=(",", ":"))
    print(name, "->", path, f"{time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    oc = OracleC()
    for name in sys.argv[1:] or ["cfg2", "cfg3"]:
        make(name, oc)
