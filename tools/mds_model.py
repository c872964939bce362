#!/usr/bin/env python3
"""Exact-integer model of the fp64-pipe MDS layer used by csrc/poseidon.cuh (development aid + table generator).

The Poseidon MDS  out[r] = sum_i s[(i+r)%12]*CIRC[i] + (r==0)*8*s[0]  is a cyclic convolution of length 12 with
c'[k] = CIRC[(-k)%12].  Splitting x^12-1 = (x^6-1)(x^6+1) and x^6-1 = (x^3-1)(x^3+1) turns it into
   cyclic-3   with [16,32,16]          (6 fp64 ops)
   negacyclic-3 with [-1,-8,2]         (9)
   negacyclic-6 with [2,-4,16,1,-1,-1] (36)
plus 30 add/sub butterflies — all constants stay integers because plonky2 chose CIRC for exactly this.
mds_limb() mirrors the CUDA code operation by operation on Python ints (Fractions for the folded constants)."""
from fractions import Fraction

CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
DIAG0 = 8
P = 0xFFFFFFFF00000001
F6 = [2, -4, 16, 1, -1, -1]


def mds_direct(s):
    return [sum(s[(i + r) % 12] * CIRC[i] for i in range(12)) + (DIAG0 * s[0] if r == 0 else 0) for r in range(12)]


def fold_constants(c):
    """c[12]: per-lane additive constants of ONE limb (multiples of 4).  Returns k[12] = uu0..2, uv0..2, v0..5."""
    A = [Fraction(c[j] + c[j + 6], 2) for j in range(3)]
    B = [Fraction(c[j + 3] + c[j + 9], 2) for j in range(3)]
    uu = [(A[j] + B[j]) / 2 for j in range(3)]
    uv = [(A[j] - B[j]) / 2 for j in range(3)]
    v = [Fraction(c[j] - c[j + 6], 2) for j in range(6)]
    return uu + uv + v


def mds_limb(s, k):
    sp = [s[i] + s[i + 6] for i in range(6)]
    sm = [s[i] - s[i + 6] for i in range(6)]
    a = [sp[i] + sp[i + 3] for i in range(3)]
    b = [sp[i] - sp[i + 3] for i in range(3)]
    S = a[0] + a[1] + a[2]
    UU = [a[(j + 2) % 3] * 16 + (S * 16 + k[j]) for j in range(3)]
    UV = [b[2] * 8 + (b[1] * -2 + (b[0] * -1 + k[3])),
          b[2] * -2 + (b[1] * -1 + (b[0] * -8 + k[4])),
          b[2] * -1 + (b[1] * -8 + (b[0] * 2 + k[5]))]
    U = [UU[j] + UV[j] for j in range(3)] + [UU[j] - UV[j] for j in range(3)]
    V = []
    for n in range(6):
        acc = k[6 + n]
        for j in range(6):
            acc += sm[j] * (F6[n - j] if j <= n else -F6[6 + n - j])
        V.append(acc)
    U[0] += s[0] * (DIAG0 // 2)   # the diagonal term must reach out[0] = U0 + V0 but not out[6] = U0 - V0
    V[0] += s[0] * (DIAG0 // 2)
    return [U[n] + V[n] for n in range(6)] + [U[n] - V[n] for n in range(6)]


def rc_limbs(rc):
    """Split a round constant into two limbs that are multiples of 4: 4*lo32(rc/4), 4*hi32(rc/4) (mod p)."""
    q = rc * pow(4, P - 2, P) % P
    return 4 * (q & 0xFFFFFFFF), 4 * (q >> 32)


if __name__ == "__main__":
    import random
    rnd = random.Random(1)
    for _ in range(200):
        s = [rnd.randrange(-(1 << 34), 1 << 34) for _ in range(12)]
        c = [4 * rnd.randrange(1 << 32) for _ in range(12)]
        k = fold_constants(c)
        assert all(x.denominator == 1 for x in k), k
        got = mds_limb(s, [int(x) for x in k])
        want = [m + cc for m, cc in zip(mds_direct(s), c)]
        assert got == want, (got, want)
    for _ in range(100):
        rc = rnd.randrange(P)
        lo, hi = rc_limbs(rc)
        assert (lo + (hi << 32)) % P == rc and lo % 4 == 0 and hi % 4 == 0
    print("mds model ok")
