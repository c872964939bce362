"""Parity at the HEADLINE shapes (BASELINE.json configs[1], configs[2], configs[4] r=1): the GPU commit, called through the C
ABI with host columns exactly as the Rust shim would, against the goldens the C oracle produced at the full size on the same
SplitMix64 input (tests/golden/headline_*.json, made by tests/golden/make_headline_golden.py).

What this adds over tests/test_gpu_parity.py: 2^20 runs as two 10-stage passes of ntt_pass_kernel<4,10> WITH the inter-pass
twiddle, 2^22 as the three-pass 8+7+7 plan — plans no smaller shape exercises — and the numbers bench.py quotes are measured
on exactly these commits.  Compared: the Merkle cap, sha256 of the coefficient matrix, sha256 of the whole digests vector
(plonky2 layout; its layer 0 is the hash of EVERY leaf row, so it pins all R x C LDE values through the KAT-pinned
permutation), 64 leaf rows word for word with their Merkle paths checked by the verifier's rule, and — memory permitting —
sha256 of the full leaf matrix."""
import os

import numpy as np
import pytest

import headline
from oracle_c import splitmix_columns

pytestmark = pytest.mark.gpu
GOLD = headline.load_all()


def _free_host_bytes():
    try:
        import psutil
        return psutil.virtual_memory().available
    except Exception:
        return 0


@pytest.mark.parametrize("name", sorted(GOLD))
def test_headline_commit_matches_oracle_golden(ctx, oc, name):
    import plonky25_b200 as g
    gold = GOLD[name]
    s = gold["shape"]
    log_n, n_cols, r, h = s["log_n"], s["n_cols"], s["rate_bits"], s["cap_height"]
    n, R = 1 << log_n, 1 << (log_n + r)
    if name.startswith("cfg5") and _free_host_bytes() < 3 * n * n_cols * 8:
        pytest.skip("not enough host memory for the 2^22 x 256 input")
    cols = splitmix_columns(gold["seed"], n_cols, n)
    pb = g.PolynomialBatch.from_values(list(cols), r, False, h, ctx=ctx)        # gl_commit, host columns, batch stays in HBM
    t = pb.merkle_tree
    try:
        assert headline.sha(t.cap.hashes) == gold["sha256_cap"]
        assert np.array_equal(t.cap.hashes, np.array(gold["cap"], dtype=np.uint64))
        del cols
        assert headline.sha(pb.polynomials) == gold["sha256_coeffs"], "coefficients differ from the oracle's"
        pb._polys = None
        assert headline.sha(t.digests) == gold["sha256_digests"], "digests differ from the oracle's"
        t._digests = None
        idx = gold["sample_rows"]
        rows, sib = t.open_batch(idx)
        assert np.array_equal(rows, np.array(gold["sample_leaves"], dtype=np.uint64)), "sampled LDE leaf rows differ"
        for q in range(0, len(idx), 8):                                         # verifier's rule (merkle_proofs.rs), oracle hashing
            assert oc.verify_path(rows[q], idx[q], sib[q], t.cap.hashes)
        leaf_bytes = R * n_cols * 8
        if leaf_bytes <= (2 << 30) or (os.environ.get("GL_TEST_FULL_LEAVES", "1") == "1" and _free_host_bytes() > 2.5 * leaf_bytes):
            assert headline.sha(t.leaves) == gold["sha256_leaves"], "leaf matrix differs from the oracle's"
            t._leaves = None
    finally:
        t.free()


@pytest.mark.parametrize("name", [k for k in sorted(GOLD) if GOLD[k]["shape"]["log_n"] <= 20])
def test_headline_device_path_matches_golden(ctx, name):
    """gl_dev_commit (the call bench.py times as `value`): columns already in HBM, all column groups in one LDE sweep — a
    different launch plan from gl_commit's chunked host path.  No torch here: the device buffer comes from gl_dev_alloc and is
    filled with gl_dev_upload."""
    import ctypes
    import plonky25_b200 as g
    gold = GOLD[name]
    s = gold["shape"]
    log_n, n_cols, r, h = s["log_n"], s["n_cols"], s["rate_bits"], s["cap_height"]
    n = 1 << log_n
    cols = splitmix_columns(gold["seed"], n_cols, n)
    lib = ctx.lib
    d = ctypes.c_void_p()
    assert lib.gl_dev_alloc(ctx.handle, n_cols * n, ctypes.byref(d)) == 0
    try:
        assert lib.gl_dev_upload(ctx.handle, cols.ctypes.data, d, n_cols * n) == 0
        cap = np.zeros((1 << h, 4), dtype=np.uint64)
        hd = ctypes.c_uint64()
        rc = lib.gl_dev_commit(ctx.handle, d, n, n_cols, log_n, r, h, 0, cap.ctypes.data, ctypes.byref(hd))
        assert rc == 0, lib.gl_ctx_last_error(ctx.handle).decode()
        lib.gl_tree_free(ctx.handle, hd.value)
        assert np.array_equal(cap, np.array(gold["cap"], dtype=np.uint64))
    finally:
        lib.gl_dev_free(ctx.handle, d)


def test_large_degree_narrow_batch_matches_oracle(ctx, oc):
    """the largest degree the suite runs: 2^24 x 5 columns, rate_bits 1 — the three-pass 8+8+8 NTT plan, 2^25 leaf rows, and 2^24 x 2
    with leaves of <= 4 elements (hash_or_noop copies them, no permutation) — against the oracle run live (narrow, so it takes seconds)"""
    import plonky25_b200 as g
    if _free_host_bytes() < 24 << 30:
        pytest.skip("not enough host memory")
    for (log_n, n_cols, r, h) in [(24, 5, 1, 4), (24, 2, 1, 0)]:
        n = 1 << log_n
        cols = splitmix_columns(900 + n_cols, n_cols, n)
        ref = oc.commit(cols, r, h, want=("digests",))
        pb = g.PolynomialBatch.from_values(list(cols), r, False, h, ctx=ctx)
        t = pb.merkle_tree
        try:
            assert np.array_equal(t.cap.hashes, ref["cap"])
            assert headline.sha(t.digests) == headline.sha(ref["digests"])
            t._digests = None
            rows, sib = t.open_batch([0, 1, (n << r) - 1, 12345678])
            for q, i in enumerate([0, 1, (n << r) - 1, 12345678]):
                assert oc.verify_path(rows[q], i, sib[q], ref["cap"])
        finally:
            t.free()
