#!/usr/bin/env python3
"""Small driver for ncu captures: `leaf` = one leaf-hash + tree build over 2^LOG leaves x 135; `lde` = one iNTT + coset LDE
of 2^20 x COLS (development aid, run under ncu on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import plonky25_b200 as g

what = sys.argv[1]
ctx = g.Context(0)
lib = ctx.lib
dev = torch.device("cuda", 0)
if what == "leaf":
    log_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    cols, pitch = 135, 136
    leaves = torch.randint(0, 2**62, (1 << log_rows, pitch), dtype=torch.int64, device=dev)
    dig = torch.empty((2 * ((1 << log_rows) - 16), 4), dtype=torch.int64, device=dev)
    cap = np.zeros(64, dtype=np.uint64)
    for _ in range(2):
        assert lib.gl_dev_merkle(ctx.handle, leaves.data_ptr(), 1 << log_rows, cols, pitch, 4, dig.data_ptr(), cap.ctypes.data) == 0
    print(ctx.stage_times()[0])
else:
    cols = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    log_n, pitch = 20, (cols + 7) // 8 * 8
    colsd = torch.randint(0, 2**62, (cols, 1 << log_n), dtype=torch.int64, device=dev)
    rows = torch.empty(((1 << log_n) * 8, pitch), dtype=torch.int64, device=dev)
    for _ in range(2):
        assert lib.gl_dev_lde(ctx.handle, colsd.data_ptr(), 1 << log_n, cols, log_n, 3, 0, rows.data_ptr(), pitch, None) == 0
    print(ctx.stage_times()[0])
