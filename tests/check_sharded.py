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
    for (log_n, cols, r, h) in [(8, 135, 3, 4), (10, 19, 1, 3), (12, 135, 3, 4), (9, 256, 2, 4)]:
        if (1 << h) < world:
            continue
        x = splitmix_columns(77 + log_n, cols, 1 << log_n)
        ref = oc.commit(x, r, h, want=("digests",))
        plan = ShardPlan(cols, log_n, r, h, world)
        c0, c1 = plan.col_range(rank)
        d = torch.from_numpy(x[c0:c1].view(np.int64).copy()).to(dev)
        for mode in ("p2p", "nccl"):
            sc = ShardedCommit(ctx, plan, rank, dist, torch, exchange=mode)
            cap = sc.commit(d).reshape(-1, 4)
            cap2 = sc.commit_host(torch.from_numpy(x[c0:c1].view(np.int64).copy()).pin_memory()).reshape(-1, 4)   # host-column path; buffers reused
            dig = sc.digests.cpu().numpy().view(np.uint64)[:plan.digests_per_rank() * 4].reshape(-1, 4)
            want_dig = ref["digests"][rank * plan.digests_per_rank():(rank + 1) * plan.digests_per_rank()]
            good = np.array_equal(cap, ref["cap"]) and np.array_equal(cap2, ref["cap"]) and np.array_equal(dig, want_dig)
            # single-GPU product path on rank 0 agrees too
            if rank == 0:
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
