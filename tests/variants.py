#!/usr/bin/env python3
"""Build and time compile-time variants of libgl_commit.so (development aid; the product library is the default build).

    python tests/variants.py build            # here (no GPU): nvcc each variant into plonky2.5_b200/variants/
    python tests/variants.py run [names...]   # on the GPU box: KAT check + leaf-hash / LDE timing of each variant
Each variant runs in its own process (one CUDA library per process).

tests/cpp/variant_bench.cpp is the torch-free twin for short GPU calls (a C++ binary starts in milliseconds where `import torch`
on a fresh box takes a minute):  tests/cpp/build/variant_bench plonky2.5_b200/libgl_commit.so plonky2.5_b200/variants/*.so
checks every library against the C oracle and prints stage times for the 2^20 x 135 commit (one JSON line per library)."""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "plonky2.5_b200", "variants")
VARIANTS = {
    "default":        [],
    "lb64":           ["-DLEAF_BLOCK=64", "-DLEAF_MIN_BLOCKS=12"],
    "lb256":          ["-DLEAF_BLOCK=256", "-DLEAF_MIN_BLOCKS=3"],
    "mb5":            ["-DLEAF_MIN_BLOCKS=5"],
    "mb7":            ["-DLEAF_MIN_BLOCKS=7"],
    "mb8":            ["-DLEAF_MIN_BLOCKS=8"],
}


def build(names):
    sys.path.insert(0, os.path.join(ROOT, "plonky2.5_b200"))
    import build as b
    os.makedirs(VDIR, exist_ok=True)
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    procs = []
    for n in names:
        out = os.path.join(VDIR, f"libgl_commit_{n}.so")
        cmd = [b._nvcc(), *b.NVCC_FLAGS, *VARIANTS[n], "-o", out, os.path.join(b.CSRC, "gl_commit.cu")]
        procs.append((n, subprocess.Popen(cmd, env=env)))
    for n, p in procs:
        assert p.wait() == 0, n
        print("built", n)


def child(name):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes
    import numpy as np
    import torch
    import plonky25_b200 as g
    g._lib.LIB_PATH = os.path.join(VDIR, f"libgl_commit_{name}.so")
    ctx = g.Context(0)
    lib = ctx.lib
    from oracle_c import OracleC
    oc = OracleC()
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "poseidon_kat.json")))["vectors"]
    ok = True
    for v in kat:
        out = ctx.poseidon_permute(np.array(v["input"], dtype=np.uint64))[0]
        ok &= [int(x) for x in out] == [int(x) for x in v["output"]]
    rng = np.random.default_rng(11)
    st = rng.integers(0, 2**64, size=(2048, 12), dtype=np.uint64)       # includes non-canonical words
    st[:64] |= np.uint64(0xFFFFFFFF00000000)
    got = ctx.poseidon_permute(st)
    want = np.stack([oc.poseidon(r) for r in st])
    rand_ok = bool(np.array_equal(got, want))
    # random states against a second variant-independent path is done by the test-suite; here: timing
    dev = torch.device("cuda", 0)
    res = {"variant": name, "kat_ok": bool(ok), "random_ok": rand_ok}
    log_rows, cols = 22, 135
    pitch = 136
    leaves = torch.randint(0, 2**62, (1 << log_rows, pitch), dtype=torch.int64, device=dev)
    dig = torch.empty((2 * ((1 << log_rows) - 16), 4), dtype=torch.int64, device=dev)
    cap = np.zeros(64, dtype=np.uint64)
    best = {}
    for it in range(4):
        rc = lib.gl_dev_merkle(ctx.handle, leaves.data_ptr(), 1 << log_rows, cols, pitch, 4, dig.data_ptr(), cap.ctypes.data)
        assert rc == 0, lib.gl_ctx_last_error(ctx.handle)
        ms, _ = ctx.stage_times()
        for k in ("leaf_hash", "tree"):
            best[k] = min(best.get(k, 1e9), ms[k])
    res["leaf_hash_ms_2p22x135"] = round(best["leaf_hash"], 3)
    res["tree_ms_2p22"] = round(best["tree"], 3)
    res["perm_per_s"] = round((1 << log_rows) * 17 / best["leaf_hash"] * 1e3 / 1e9, 4)
    res["cap0"] = int(cap[0])
    del leaves, dig
    log_n = 20
    colsd = torch.randint(0, 2**62, (cols, 1 << log_n), dtype=torch.int64, device=dev)
    rows = torch.empty(((1 << log_n) * 8, pitch), dtype=torch.int64, device=dev)
    for it in range(3):
        rc = lib.gl_dev_lde(ctx.handle, colsd.data_ptr(), 1 << log_n, cols, log_n, 3, 0, rows.data_ptr(), pitch, None)
        assert rc == 0
        ms, _ = ctx.stage_times()
        for k in ("intt", "lde"):
            best[k] = min(best.get(k, 1e9), ms[k])
    res["intt_ms"] = round(best["intt"], 3); res["lde_ms"] = round(best["lde"], 3)
    res["lde_check"] = int(rows[12345, 77].item()) & 0xFFFFFFFF
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    cmd = sys.argv[1]
    names = sys.argv[2:] or list(VARIANTS)
    if cmd == "build":
        build(names)
    elif cmd == "child":
        child(sys.argv[2])
    else:
        for n in names:
            subprocess.run([sys.executable, os.path.abspath(__file__), "child", n])
