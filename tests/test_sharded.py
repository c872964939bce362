"""Host logic of the multi-GPU commit (plonky2.5_b200/sharded.py) on CPU: the shard plan, and the column->row exchange run
for real over torch.distributed with the gloo backend at world_size 2 (the N>1 path's collective, without GPUs)."""
import os
import socket

import numpy as np
import pytest

import plonky25_b200 as g
from plonky25_b200.sharded import ShardPlan, coset_blocks_of_rank, exchange_reference


def test_plan_partitions_cover_everything():
    for (cols, log_n, r, h, world) in [(135, 20, 3, 4, 8), (135, 16, 3, 4, 4), (256, 22, 1, 4, 2), (20, 10, 3, 4, 16), (17, 8, 2, 1, 2)]:
        p = ShardPlan(cols, log_n, r, h, world)
        assert sum(p.col_counts) == cols and p.col_offsets[0] == 0
        assert all(p.col_range(k)[1] == p.col_range(k + 1)[0] for k in range(world - 1))
        padded = [(c + 3) // 4 * 4 for c in p.col_counts]
        assert min(p.col_counts) >= 1 and max(padded) - min(padded) <= 4          # balanced in units of the NTT's 4-column groups
        if (cols + 3) // 4 >= world:
            assert all(o % 4 == 0 for o in p.col_offsets)                         # sector-aligned column offsets
        assert p.rows_per_rank * world == 1 << (log_n + r)
        assert (1 << p.local_cap_height) * world == 1 << h
        assert all(pt % 4 == 0 and pt >= c and pt - c < 4 for pt, c in zip(p.pitches, p.col_counts))
        # plonky2 digest layout: the global digests vector is the concatenation of the ranks' local vectors
        assert p.digests_per_rank() * world == 2 * ((1 << (log_n + r)) - (1 << h))
        for k in range(world):
            assert sum(p.recv_splits(k)) == p.rows_per_rank * sum(p.pitches)
            assert p.send_splits(k)[0] == p.recv_splits((k + 1) % world)[k]


def test_plan_rejects_bad_shapes():
    with pytest.raises(ValueError):
        ShardPlan(135, 10, 3, 4, 3)
    with pytest.raises(ValueError):
        ShardPlan(135, 10, 3, 2, 8)      # 4 subtrees cannot be split over 8 ranks
    with pytest.raises(ValueError):
        ShardPlan(3, 10, 3, 4, 4)


def _shard(plan, rank):
    """synthetic 'LDE output' of a rank: value encodes (row, global column); padding columns hold a poison value"""
    R, pt, c0 = plan.n_rows, plan.pitches[rank], plan.col_offsets[rank]
    a = np.full((R, pt), -1, dtype=np.int64)
    rows = np.arange(R, dtype=np.int64)[:, None]
    a[:, :plan.col_counts[rank]] = rows * 1000 + c0 + np.arange(plan.col_counts[rank], dtype=np.int64)[None, :]
    return a


def _worker(rank, world, port, cols, log_n, r, h, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = ShardPlan(cols, log_n, r, h, world)
        send = torch.from_numpy(_shard(plan, rank).reshape(-1))
        recv = torch.empty(sum(plan.recv_splits(rank)), dtype=torch.int64)
        dist.all_to_all_single(recv, send, plan.recv_splits(rank), plan.send_splits(rank))
        leaves = np.zeros((plan.rows_per_rank, cols), dtype=np.int64)
        off = 0
        for src in range(world):                       # what gl_dev_repack does on the device
            blk = recv[off:off + plan.rows_per_rank * plan.pitches[src]].numpy().reshape(plan.rows_per_rank, plan.pitches[src])
            leaves[:, plan.col_offsets[src]:plan.col_offsets[src] + plan.col_counts[src]] = blk[:, :plan.col_counts[src]]
            off += blk.size
        r0, _ = plan.row_range(rank)
        want = (np.arange(r0, r0 + plan.rows_per_rank, dtype=np.int64)[:, None] * 1000 + np.arange(cols, dtype=np.int64)[None, :])
        ok = bool(np.array_equal(leaves, want))
        # subtree-root gather
        cap_local = torch.full((4 << plan.local_cap_height,), rank, dtype=torch.int64)
        cap_all = [torch.empty_like(cap_local) for _ in range(world)]
        dist.all_gather(cap_all, cap_local)
        ok = ok and [int(c[0]) for c in cap_all] == list(range(world))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(135, 6, 3, 4), (19, 5, 1, 1)])
def test_exchange_over_gloo_world2(shape):
    import torch.multiprocessing as mp
    cols, log_n, r, h = shape
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_worker, args=(k, 2, port, cols, log_n, r, h, q)) for k in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_exchange_reference_model():
    plan = ShardPlan(37, 4, 2, 2, 4)
    shards = [_shard(plan, k) for k in range(4)]
    leaves = exchange_reference(plan, shards)
    for k in range(4):
        r0, _ = plan.row_range(k)
        want = (np.arange(r0, r0 + plan.rows_per_rank, dtype=np.int64)[:, None] * 1000 + np.arange(37, dtype=np.int64)[None, :])
        assert np.array_equal(leaves[k], want)


# ---- coset-sharded plan (the default multi-GPU plan): host logic on CPU, the exchange over gloo at world_size 2 -------------------------
def _coset_leaves(oc, coeff_block, log_n, rate_bits, s):
    """what gl_dev_lde_own_cosets computes for one coefficient block and one coset, with the oracle: rows m = in-place-DIF order of the
    size-N NTT of coeffs * g_s^j, g_s = 7 * w_R^s"""
    n = 1 << log_n
    P = 0xFFFF_FFFF_0000_0001
    w_R = pow(1753635133440165772, 1 << (32 - log_n - rate_bits), P)
    g_s = 7 * pow(w_R, s, P) % P
    rev = [int(format(i, "0%db" % log_n)[::-1], 2) if log_n else 0 for i in range(n)]
    cols = []
    for c in coeff_block:                                    # [n_cols][N]
        vals = oc.coset_fft(c, g_s)                          # natural order
        cols.append(vals[rev])
    return np.stack(cols, axis=1)                            # [N][n_cols]


def _coset_worker(rank, world, port, cols_shape, q):
    import torch
    import torch.distributed as dist
    from oracle_c import OracleC, splitmix_columns
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_cols, log_n, r, h = cols_shape
        oc = OracleC()
        oc.set_threads(1)
        plan = ShardPlan(n_cols, log_n, r, h, world)
        x = splitmix_columns(321, n_cols, 1 << log_n)
        c0, c1 = plan.col_range(rank)
        my_coeffs = np.stack([oc.ifft(x[j]) for j in range(c0, c1)])                 # gl_dev_intt on this rank's column shard
        # the exchange: every rank ends up with every coefficient block (the product pulls them over NVLink; here an all-gather of padded blocks)
        width = max(plan.col_counts)
        mine = torch.zeros((width, 1 << log_n), dtype=torch.int64)
        mine[:c1 - c0] = torch.from_numpy(my_coeffs.view(np.int64))
        allb = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allb, mine)
        blocks = [b.numpy().view(np.uint64)[:plan.col_counts[g]] for g, b in enumerate(allb)]
        # own cosets of every block -> own leaf range
        rows = []
        for (_, s) in coset_blocks_of_rank(plan, rank):
            rows.append(np.concatenate([_coset_leaves(oc, blocks[g], log_n, r, s) for g in range(world)], axis=1))
        leaves = np.concatenate(rows, axis=0)
        ref = oc.commit(x, r, h)
        r0, r1 = plan.row_range(rank)
        ok = bool(np.array_equal(leaves, ref["leaves"][r0:r1]))
        dig, cap = oc.merkle_new(leaves, plan.local_cap_height)
        per = (1 << h) // world
        ok = ok and bool(np.array_equal(cap, ref["cap"][rank * per:(rank + 1) * per]))
        ok = ok and bool(np.array_equal(dig, ref["digests"][rank * plan.digests_per_rank():(rank + 1) * plan.digests_per_rank()]))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_coset_blocks_partition_the_leaf_rows():
    for (cols, log_n, r, h, world) in [(135, 20, 3, 4, 8), (135, 16, 3, 4, 2), (20, 6, 2, 2, 4), (9, 5, 1, 1, 2), (7, 4, 0, 0, 1)]:
        p = ShardPlan(cols, log_n, r, h, world)
        seen = []
        for k in range(world):
            blocks = coset_blocks_of_rank(p, k)
            assert [b for b, _ in blocks] == list(range(k * len(blocks), (k + 1) * len(blocks)))      # contiguous leaf range
            assert len(blocks) * (1 << log_n) == p.rows_per_rank
            seen += [s for _, s in blocks]
        assert sorted(seen) == list(range(1 << r))                                                  # every coset exactly once
    with pytest.raises(ValueError):
        coset_blocks_of_rank(ShardPlan(64, 6, 1, 3, 4), 0)                                          # 4 ranks > 2 cosets


@pytest.mark.parametrize("shape", [(19, 5, 2, 2), (135, 4, 1, 1)])
def test_coset_plan_over_gloo_world2(shape):
    """two ranks, gloo: iNTT of the column shard -> exchange of coefficient blocks -> own cosets of all columns -> own subtrees: each
    rank's leaf range, digest slice and cap slice equal the single commit's (the oracle stands in for the kernels)"""
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_coset_worker, args=(k, 2, port, shape, q)) for k in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


# ---- streamed coset plan (gl_commit_coset_stream): the cyclic column deal, host logic on CPU + gloo at world_size 2 ----------------------
def test_stream_columns_partition_and_match_the_library():
    """every column is dealt to exactly one rank, wave w covers columns [w*G*gw, (w+1)*G*gw) left to right (what the sponge needs), and the
    C library derives the same wave count / per-rank column count from the same plan (no GPU involved)"""
    import ctypes
    from plonky25_b200 import _lib as L
    lib = L.load()
    for (cols, log_n, r, h, world) in [(135, 20, 3, 4, 8), (135, 20, 3, 4, 4), (135, 16, 3, 4, 2), (256, 22, 1, 4, 2), (19, 5, 2, 2, 4), (9, 5, 1, 1, 2),
                                       (9, 4, 3, 3, 8)]:
        p = ShardPlan(cols, log_n, r, h, world)
        gw, W = p.stream_group_width(), p.stream_waves()
        assert gw in (4, 8)
        dealt = [p.stream_columns(k) for k in range(world)]
        assert sorted(sum(dealt, [])) == list(range(cols))
        for w in range(W):
            wave = sorted(c for k in range(world) for c in dealt[k] if w * world * gw <= c < (w + 1) * world * gw)
            assert wave == list(range(w * world * gw, min(cols, (w + 1) * world * gw)))
        for k in range(world):
            assert dealt[k] == sorted(dealt[k]) and all(c // gw % world == k for c in dealt[k])
            sp = L.StreamPlan(cols, log_n, r, p.local_cap_height, world, k, gw, p.leaf_pitch, 0)
            ew, sw, nw, no = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint32(), ctypes.c_uint32()
            assert lib.gl_stream_plan_sizes(ctypes.byref(sp), ctypes.byref(ew), ctypes.byref(sw), ctypes.byref(nw), ctypes.byref(no)) == 0
            assert (nw.value, no.value) == (W, len(dealt[k]))
            assert ew.value == W * (gw << log_n) + world * W and sw.value == world * W * (gw << log_n)
    rng = np.random.default_rng(7)                            # the same arithmetic on 300 random plans
    for _ in range(300):
        world = 1 << int(rng.integers(0, 4))
        r = int(rng.integers(world.bit_length() - 1, 5))
        log_n = int(rng.integers(1, 24))
        cols = int(rng.integers(max(world, 5), 600))
        h = int(rng.integers(world.bit_length() - 1, min(log_n + r, 6) + 1))
        if h > log_n + r:
            continue
        p = ShardPlan(cols, log_n, r, h, world)
        gw, W = p.stream_group_width(), p.stream_waves()
        dealt = [p.stream_columns(k) for k in range(world)]
        assert sorted(sum(dealt, [])) == list(range(cols)) and (world * gw) % 8 == 0
        for k in range(world):
            sp = L.StreamPlan(cols, log_n, r, p.local_cap_height, world, k, gw, p.leaf_pitch, 0)
            nw, no = ctypes.c_uint32(), ctypes.c_uint32()
            assert lib.gl_stream_plan_sizes(ctypes.byref(sp), None, None, ctypes.byref(nw), ctypes.byref(no)) == 0, (cols, log_n, r, h, world)
            assert (nw.value, no.value) == (W, len(dealt[k]))
    bad = L.StreamPlan(135, 20, 1, 1, 4, 0, 4, 136, 0)       # 4 ranks > 2 cosets
    assert lib.gl_stream_plan_sizes(ctypes.byref(bad), None, None, None, None) != 0
    bad = L.StreamPlan(135, 20, 3, 1, 8, 0, 6, 136, 0)       # group width must be 4 or 8
    assert lib.gl_stream_plan_sizes(ctypes.byref(bad), None, None, None, None) != 0


def _stream_worker(rank, world, port, cols_shape, q):
    import torch
    import torch.distributed as dist
    from oracle_c import OracleC, splitmix_columns
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_cols, log_n, r, h = cols_shape
        oc = OracleC()
        oc.set_threads(1)
        plan = ShardPlan(n_cols, log_n, r, h, world)
        gw, W = plan.stream_group_width(), plan.stream_waves()
        x = splitmix_columns(654, n_cols, 1 << log_n)
        mine = plan.stream_columns(rank)
        host = x[mine]                                              # what this rank hands to commit_host, in order
        blocks = coset_blocks_of_rank(plan, rank)
        leaves = np.zeros((plan.rows_per_rank, plan.leaf_pitch), dtype=np.uint64)
        absorbed = 0
        for w in range(W):
            own = [j for j, c in enumerate(mine) if c // (gw * world) == w]
            grp = torch.zeros((gw, 1 << log_n), dtype=torch.int64)     # the wave's own group after the iNTT, padded to gw columns
            for k, j in enumerate(own):
                grp[k] = torch.from_numpy(oc.ifft(host[j]).view(np.int64))
            allg = [torch.zeros_like(grp) for _ in range(world)]
            dist.all_gather(allg, grp)                              # the product pulls the groups over NVLink behind the tickets
            for peer, gq in enumerate(allg):
                c0 = gw * (w * world + peer)
                nc = max(0, min(gw, n_cols - c0))
                if nc == 0:
                    continue
                for b, (_, s) in enumerate(blocks):
                    leaves[b << log_n:(b + 1) << log_n, c0:c0 + nc] = _coset_leaves(oc, gq.numpy().view(np.uint64)[:nc], log_n, r, s)
            c1 = min(n_cols, (w + 1) * world * gw)
            assert absorbed == w * world * gw and c1 > absorbed      # the sponge sees the leaf strictly left to right, 8-column aligned
            assert absorbed % 8 == 0
            absorbed = c1
        ref = oc.commit(x, r, h)
        r0, r1 = plan.row_range(rank)
        ok = bool(np.array_equal(leaves[:, :n_cols], ref["leaves"][r0:r1]))
        dig, cap = oc.merkle_new(np.ascontiguousarray(leaves[:, :n_cols]), plan.local_cap_height)
        per = (1 << h) // world
        ok = ok and bool(np.array_equal(cap, ref["cap"][rank * per:(rank + 1) * per]))
        ok = ok and bool(np.array_equal(dig, ref["digests"][rank * plan.digests_per_rank():(rank + 1) * plan.digests_per_rank()]))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(19, 5, 2, 2), (135, 4, 1, 1)])
def test_stream_plan_over_gloo_world2(shape):
    """two ranks, gloo: wave by wave — iNTT of the own column group, exchange of the wave's groups, own cosets into the own leaf range at the
    group's column offset; the finished leaf range, digest slice and cap slice equal the single commit's (oracle kernels)"""
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_stream_worker, args=(k, 2, port, shape, q)) for k in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
