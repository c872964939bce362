#!/usr/bin/env python3
"""GPU probe (run under gpurun): integer-pipe micro-benchmarks (K7) + stage timings of device-resident commits."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import plonky25_b200 as g  # noqa: E402


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    ctx = g.Context(0)
    names = {0: "imad_wide_u32", 1: "lop3", 2: "mixed_imad_lop3", 3: "gl_modmul", 4: "poseidon_perm"}
    out["microbench_ops_per_s"] = {names[w]: ctx.microbench(w, 400 if w == 4 else 4000) for w in names}
    shapes = [(16, 135, 3, 4), (20, 135, 3, 4)]
    if "--big" in sys.argv:
        shapes.append((22, 256, 1, 4))
    out["commits"] = []
    lib = ctx.lib
    import ctypes
    for (log_n, n_cols, r, h) in shapes:
        n = 1 << log_n
        gen = torch.Generator(device="cuda").manual_seed(1)
        d = torch.randint(0, 2**63 - 1, (n_cols, n), dtype=torch.int64, device="cuda", generator=gen)
        cap = np.zeros(4 << h, dtype=np.uint64)
        best = None
        for it in range(4):
            hd = ctypes.c_uint64()
            t0 = time.time()
            rc = lib.gl_dev_commit(ctx.handle, d.data_ptr(), n, n_cols, log_n, r, h, 0, cap.ctypes.data, ctypes.byref(hd))
            dt = time.time() - t0
            assert rc == 0, lib.gl_ctx_last_error(ctx.handle)
            ms, launches = ctx.stage_times()
            lib.gl_tree_free(ctx.handle, hd.value)
            if best is None or dt < best["wall_s"]:
                best = {"wall_s": dt, "stage_ms": ms, "launches": launches}
        best.update({"log_n": log_n, "n_cols": n_cols, "rate_bits": r, "melem_per_s": n * n_cols / best["wall_s"] / 1e6})
        out["commits"].append(best)
        del d
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
