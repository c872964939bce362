#!/usr/bin/env python3
"""Do the NTT passes (integer ALU pipe, no fp64) and the Poseidon leaf hash (alu + fp64 + IMAD.WIDE) share the SM better than they
serialise?  Two contexts on one GPU, one host thread each: K_l coset-LDE calls against K_h leaf-hash calls, alone and concurrently
(development aid; result recorded in DESIGN.md §8)."""
import ctypes, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import plonky25_b200 as g
from oracle_c import splitmix_columns

log_n, cols, r, pitch = 20, 135, 3, 136
n = 1 << log_n
a, b = g.Context(0), g.Context(0)
lib = a.lib


def dalloc(ctx, words):
    p = ctypes.c_void_p()
    assert lib.gl_dev_alloc(ctx.handle, words, ctypes.byref(p)) == 0
    return p


d_cols = dalloc(a, cols * n)
x = splitmix_columns(1, cols, n)
assert lib.gl_dev_upload(a.handle, x.ctypes.data, d_cols, cols * n) == 0
d_rows = dalloc(a, (n << r) * pitch)
d_rows2 = dalloc(b, (n << r) * pitch)      # leaf matrix for the hashing context (contents irrelevant: timing only)
d_dig = dalloc(b, 2 * ((n << r) - 16) * 4)
cap = np.zeros(64, dtype=np.uint64)


def lde(k):
    for _ in range(k):
        assert lib.gl_dev_lde(a.handle, d_cols, n, cols, log_n, r, 0, d_rows, pitch, None) == 0


def hsh(k):
    for _ in range(k):
        assert lib.gl_dev_merkle(b.handle, d_rows2, n << r, cols, pitch, 4, d_dig, cap.ctypes.data) == 0


lde(2); hsh(1)
KL, KH = 20, 4
t = time.perf_counter(); lde(KL); t_l = time.perf_counter() - t
t = time.perf_counter(); hsh(KH); t_h = time.perf_counter() - t
ths = [threading.Thread(target=lde, args=(KL,)), threading.Thread(target=hsh, args=(KH,))]
t = time.perf_counter()
for th in ths: th.start()
for th in ths: th.join()
t_c = time.perf_counter() - t
print({"lde_alone_ms": round(t_l * 1e3 / KL, 2), "hash_alone_ms": round(t_h * 1e3 / KH, 2), "serial_total_ms": round((t_l + t_h) * 1e3, 1),
       "concurrent_total_ms": round(t_c * 1e3, 1), "gain": round(1 - t_c / (t_l + t_h), 3)})
