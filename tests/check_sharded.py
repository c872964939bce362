#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun on N GPUs): the sharded commit's cap and every rank's digests slice
must equal the single-GPU commit and the CPU oracle on the same seeded columns.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/check_sharded.py
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import plonky25_b200 as g  # noqa: E402
from oracle_c import OracleC, splitmix_columns  # noqa: E402
from plonky25_b200.sharded import ShardedCommit, ShardPlan  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = g.Context(local)
    oc = OracleC()
    ok = True
    import headline
    shapes = [(8, 135, 3, 4, 0), (10, 19, 1, 3, 0), (12, 135, 3, 4, 0), (9, 256, 2, 4, 0)]
    # headline shapes with committed oracle goldens (seed 1 = bench.py's input): 2^16 x 135 always, 2^20 x 135 on request
    shapes.append((16, 135, 3, 4, 1))
    if os.environ.get("GL_CHECK_SHARDED_CFG3") == "1":
        shapes.append((20, 135, 3, 4, 1))
    for (log_n, cols, r, h, gold_seed) in shapes:
        if (1 << h) < world:
            continue
        gold = None
        if gold_seed:
            _, gold = headline.find(log_n, cols, r, h, gold_seed)
            if gold is None:
                continue
            x = splitmix_columns(gold_seed, cols, 1 << log_n)
            ref = {"cap": np.array(gold["cap"], dtype=np.uint64), "digests": None}
        else:
            x = splitmix_columns(77 + log_n, cols, 1 << log_n)
            ref = oc.commit(x, r, h, want=("digests",))
        plan = ShardPlan(cols, log_n, r, h, world)
        c0, c1 = plan.col_range(rank)
        d = torch.from_numpy(x[c0:c1].view(np.int64).copy()).to(dev)
        for mode in os.environ.get("GL_CHECK_MODES", "coset,stream,p2p,nccl,auto").split(","):
            if mode in ("coset", "stream") and world > (1 << r):
                continue
            sc = ShardedCommit(ctx, plan, rank, dist, torch, exchange=mode)
            if mode != "auto":
                assert sc.exchange == mode, (sc.exchange, mode)
            idx = sc.host_columns()          # the streamed plan deals the host columns cyclically (ShardPlan.stream_columns)
            host = torch.from_numpy(x[idx].view(np.int64).copy()).pin_memory()
            if mode == "stream":             # host columns only; run it three times on the same buffers (epochs / tickets)
                cap = sc.commit_host(host).reshape(-1, 4).copy()
                sc.commit_host(host)
            else:
                cap = sc.commit(d).reshape(-1, 4)
            cap2 = sc.commit_host(host).reshape(-1, 4)   # host-column path; buffers reused
            if mode == "auto":
                if rank == 0:
                    print(f"auto: device plan {sc.exchange}, host plan {sc._host().exchange}", flush=True)
                sc.digests = sc._host().digests
            dig = sc.digests.cpu().numpy().view(np.uint64)[:plan.digests_per_rank() * 4].reshape(-1, 4)
            good = np.array_equal(cap, ref["cap"]) and np.array_equal(cap2, ref["cap"])
            if gold is None:
                want_dig = ref["digests"][rank * plan.digests_per_rank():(rank + 1) * plan.digests_per_rank()]
                good = good and np.array_equal(dig, want_dig)
            else:
                # the global digests vector is the concatenation of the ranks' local vectors: sha256 over the gathered slices
                # must equal the oracle's sha256 of MerkleTree::digests
                parts = [None] * world
                dist.all_gather_object(parts, dig.tobytes())
                import hashlib
                good = good and hashlib.sha256(b"".join(parts)).hexdigest() == gold["sha256_digests"]
                # this rank's leaf rows = whole cap subtrees: sha256 per subtree against the oracle's leaves
                if mode != "auto" and (sc._peer_ptrs is not None or hasattr(sc, "leaves")):
                    per = (1 << h) // world
                    sub_rows = plan.rows_per_rank // per
                    lv = np.zeros(plan.rows_per_rank * plan.leaf_pitch, dtype=np.uint64)
                    assert ctx.lib.gl_dev_download(ctx.handle, sc.leaves_ptr, lv.ctypes.data, lv.size) == 0
                    lv = lv.reshape(plan.rows_per_rank, plan.leaf_pitch)[:, :cols]
                    for k in range(per):
                        good = good and headline.sha(lv[k * sub_rows:(k + 1) * sub_rows]) == gold["sha256_leaves_per_subtree"][rank * per + k]
            # single-GPU product path on rank 0 agrees too
            if rank == 0 and log_n <= 16:
                pb = g.PolynomialBatch.from_values(list(x), r, False, h, ctx=ctx)
                good = good and np.array_equal(pb.merkle_tree.cap.hashes, cap)
            print(f"rank {rank} shape {(log_n, cols, r, h)} world {world} {mode}: {'ok' if good else 'MISMATCH'}", flush=True)
            ok = ok and good
            sc.close()
    t = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(t)
    dist.destroy_process_group()
    sys.exit(int(t.item() != 0))


if __name__ == "__main__":
    main()
