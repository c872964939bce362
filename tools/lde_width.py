#!/usr/bin/env python3
"""LDE time per column as a function of the shard width (one GPU, no exchange): gl_dev_lde of 2^20 x C, rate_bits 3."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import plonky25_b200 as g

ctx = g.Context(0)
lib = ctx.lib
dev = torch.device("cuda", 0)
log_n = 20
for cols in (16, 20, 32, 36, 64, 68, 136, 256):
    pitch = (cols + 7) // 8 * 8 if ((cols + 7) // 8 * 8 - cols) < 4 else (cols + 3) // 4 * 4
    colsd = torch.randint(0, 2**62, (cols, 1 << log_n), dtype=torch.int64, device=dev)
    rows = torch.empty(((1 << log_n) * 8, pitch), dtype=torch.int64, device=dev)
    best = {}
    for it in range(3):
        assert lib.gl_dev_lde(ctx.handle, colsd.data_ptr(), 1 << log_n, cols, log_n, 3, 0, rows.data_ptr(), pitch, None) == 0
        ms, _ = ctx.stage_times()
        for k in ("intt", "lde"):
            best[k] = min(best.get(k, 1e9), ms[k])
    print(f"cols {cols:4d} pitch {pitch:4d}  intt {best['intt']:7.3f} ms  lde {best['lde']:7.3f} ms  lde/col {best['lde'] / cols:6.4f} ms", flush=True)
    del colsd, rows
