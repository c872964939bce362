"""GPU tests of the streamed coset-sharded commit (include/gl_commit.h · gl_commit_coset_stream) that fit ONE GPU: the whole wave
machinery with a single rank (copy -> iNTT -> one-launch own-coset NTTs from per-group blocks -> leaf sponge wave by wave -> tree), against
the oracle; and the ticket wait's time limit.  The multi-rank form (tickets + NVLink pulls between processes) is covered by
tests/test_sharded_gpu.py (torchrun, >= 2 GPUs) and, in one process, by test_commit_multi_single_process on a multi-GPU box."""
import ctypes
import os

import numpy as np
import pytest

from oracle_c import splitmix_columns

pytestmark = pytest.mark.gpu


def _alloc(ctx, words):
    p = ctypes.c_void_p()
    assert ctx.lib.gl_dev_alloc(ctx.handle, max(int(words), 1), ctypes.byref(p)) == 0
    return p


def _download(ctx, ptr, words):
    out = np.zeros(int(words), dtype=np.uint64)
    assert ctx.lib.gl_dev_download(ctx.handle, ptr, out.ctypes.data, out.size) == 0
    return out


def _sizes(lib, plan):
    ew, sw, nw, no = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint32(), ctypes.c_uint32()
    assert lib.gl_stream_plan_sizes(ctypes.byref(plan), ctypes.byref(ew), ctypes.byref(sw), ctypes.byref(nw), ctypes.byref(no)) == 0
    return ew.value, sw.value, nw.value, no.value


@pytest.mark.parametrize("shape", [(12, 135, 3, 4, 0), (10, 19, 1, 3, 0), (9, 64, 2, 0, 0), (13, 20, 0, 2, 1), (21, 9, 1, 4, 0), (2, 23, 3, 1, 0)])
def test_stream_single_rank_matches_oracle(ctx, oc, shape):
    """n_peers = 1: waves of 8 columns; leaves, digests and cap equal the oracle's commit (and so the one-launch multi-source NTT, the 3-pass
    plan at 2^21, the tiny-transform path at 2^2, coefficient inputs and three calls on the same buffers / epochs)"""
    from plonky25_b200._lib import StreamPlan
    log_n, n_cols, r, h, is_coeffs = shape
    lib = ctx.lib
    n, rows, pitch = 1 << log_n, 1 << (log_n + r), (n_cols + 7) // 8 * 8
    cols = splitmix_columns(900 + log_n, n_cols, n)
    ref = oc.commit(cols, r, h, is_coeffs=bool(is_coeffs)) if log_n <= 16 else None
    plan = StreamPlan(n_cols, log_n, r, h, 1, 0, 8, pitch, 0)
    ew, sw, nw, no = _sizes(lib, plan)
    assert (nw, no) == ((n_cols + 7) // 8, n_cols)
    own, stage, leaves = _alloc(ctx, ew), _alloc(ctx, sw), _alloc(ctx, rows * pitch)
    n_dig = 2 * (rows - (1 << h))
    dig = _alloc(ctx, n_dig * 4)
    zero = np.zeros(ew, dtype=np.uint64)
    assert lib.gl_dev_upload(ctx.handle, zero.ctypes.data, own, ew) == 0      # ticket words start at 0 (gl_dev_ipc_alloc does this itself)
    try:
        host = np.ascontiguousarray(cols)
        ptrs = (ctypes.c_void_p * n_cols)(*[host[j].ctypes.data for j in range(n_cols)])
        peers = (ctypes.c_void_p * 1)(own.value)
        cap = np.zeros(4 << h, dtype=np.uint64)
        for epoch in range(3):
            plan.epoch = epoch
            rc = lib.gl_commit_coset_stream(ctx.handle, ctypes.byref(plan), ptrs, is_coeffs, peers, stage, leaves, dig, cap.ctypes.data)
            assert rc == 0, lib.gl_ctx_last_error(ctx.handle).decode()
        # the product's own single-GPU commit on the same input (bit-exact against the oracle elsewhere; here also at sizes the oracle skips)
        import plonky25_b200 as g
        make = g.PolynomialBatch.from_coeffs if is_coeffs else g.PolynomialBatch.from_values
        pb = make(list(cols), r, False, h, ctx=ctx)
        assert np.array_equal(cap.reshape(-1, 4), pb.merkle_tree.cap.hashes)
        got_dig = _download(ctx, dig, n_dig * 4).reshape(-1, 4)
        assert np.array_equal(got_dig, pb.merkle_tree.digests)
        if log_n <= 16:
            got_leaves = _download(ctx, leaves, rows * pitch).reshape(rows, pitch)[:, :n_cols]
            assert np.array_equal(got_leaves, pb.merkle_tree.leaves)
        if ref is not None:
            assert np.array_equal(cap.reshape(-1, 4), ref["cap"]) and np.array_equal(got_dig, ref["digests"])
            assert np.array_equal(got_leaves, ref["leaves"])
    finally:
        for p in (own, stage, leaves, dig):
            lib.gl_dev_free(ctx.handle, p)


def test_stream_rejects_bad_plans(ctx):
    from plonky25_b200._lib import StreamPlan
    lib = ctx.lib
    for plan in (StreamPlan(135, 10, 1, 1, 4, 0, 4, 136, 0),       # 4 ranks > 2 cosets
                 StreamPlan(135, 10, 3, 1, 1, 0, 4, 136, 0),       # a wave of 4 columns is not a multiple of the sponge rate
                 StreamPlan(135, 10, 3, 1, 8, 8, 4, 136, 0),       # rank out of range
                 StreamPlan(135, 10, 3, 1, 8, 0, 4, 128, 0),       # leaf pitch too small
                 StreamPlan(4, 10, 3, 1, 2, 0, 4, 8, 0),           # <= 4 columns: leaves are not hashed (hash_or_noop)
                 StreamPlan(135, 10, 3, 12, 8, 0, 4, 136, 0)):     # cap height above the rank's leaf range
        assert lib.gl_stream_plan_sizes(ctypes.byref(plan), None, None, None, None) != 0
        assert lib.gl_commit_coset_stream(ctx.handle, ctypes.byref(plan), None, 0, None, None, None, None, None) != 0


def test_stream_peer_timeout_is_an_error_not_a_hang(ctx):
    """rank 0 of 2 with a 'peer' buffer nobody ever signals: the ticket waits give up after GL_PEER_TIMEOUT_MS and the call reports it; the
    context stays usable"""
    from plonky25_b200._lib import StreamPlan
    lib = ctx.lib
    log_n, n_cols, r, h = 8, 19, 1, 0
    n, rows, pitch = 1 << log_n, 1 << log_n, 24                      # rank 0 owns one of the two cosets
    plan = StreamPlan(n_cols, log_n, r, h, 2, 0, 4, pitch, 0)
    ew, sw, nw, no = _sizes(lib, plan)
    bufs = [_alloc(ctx, ew), _alloc(ctx, ew), _alloc(ctx, sw), _alloc(ctx, rows * pitch), _alloc(ctx, 2 * (rows - 1) * 4)]
    zero = np.zeros(ew, dtype=np.uint64)
    for b in bufs[:2]:
        assert lib.gl_dev_upload(ctx.handle, zero.ctypes.data, b, ew) == 0
    old = os.environ.get("GL_PEER_TIMEOUT_MS")
    os.environ["GL_PEER_TIMEOUT_MS"] = "100"
    try:
        cols = np.ascontiguousarray(splitmix_columns(5, no, n))
        ptrs = (ctypes.c_void_p * no)(*[cols[j].ctypes.data for j in range(no)])
        peers = (ctypes.c_void_p * 2)(bufs[0].value, bufs[1].value)
        cap = np.zeros(4, dtype=np.uint64)
        rc = lib.gl_commit_coset_stream(ctx.handle, ctypes.byref(plan), ptrs, 0, peers, bufs[2], bufs[3], bufs[4], cap.ctypes.data)
        assert rc == -2 and b"did not publish" in lib.gl_ctx_last_error(ctx.handle)
        assert ctx.poseidon_permute(np.zeros((1, 12), dtype=np.uint64)).shape == (1, 12)      # still alive
    finally:
        if old is None:
            del os.environ["GL_PEER_TIMEOUT_MS"]
        else:
            os.environ["GL_PEER_TIMEOUT_MS"] = old
        for b in bufs:
            lib.gl_dev_free(ctx.handle, b)
