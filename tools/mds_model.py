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


# ======================================================================================================================
# v3: lanes 1..11 stay in fp64 limbs through the 22 partial rounds
# ======================================================================================================================
N_ROUNDS, FULL_HALF, N_PARTIAL = 30, 4, 22
BIAS = 1 << 52
# positivity offsets for limbs that leave the fp64 pipe after having been signed:  OFF_LO + OFF_HI*2^32 = 2^17 * p
OFF_LO = (1 << 49) + (1 << 17)
OFF_HI = (1 << 49) - (1 << 18)
assert OFF_LO + (OFF_HI << 32) == (1 << 17) * P and OFF_LO % 4 == 0 and OFF_HI % 4 == 0


def mds_field(v):
    return [x % P for x in mds_direct(v)]


def partial_constants(rc):
    """Push the lane 1..11 constants of partial rounds 5..25 forward through the MDS (sbox acts on lane 0 only, so
    sigma0(s + e + d) = sigma0(s + e) + d for d with zero lane 0).  Returns (lane0[r] for r = 5..26, tail[1..11] added
    before round 26's S-boxes)."""
    cur = [rc[12 * 5 + i] for i in range(12)]
    lane0 = {}
    for r in range(5, 26):
        lane0[r] = cur[0]
        d = [0] + cur[1:]
        nxt = mds_field(d)
        cur = [(rc[12 * (r + 1) + i] + nxt[i]) % P for i in range(12)]
    lane0[26] = cur[0]
    return lane0, cur[1:]


def check53(*vals):
    for v in vals:
        assert abs(v) < (1 << 53), "fp64 exactness bound violated: %d" % v


def mds_limb_checked(s, k):
    """mds_limb with the |x| < 2^53 check on every intermediate (all values are integers, so that is exactness)."""
    sp = [s[i] + s[i + 6] for i in range(6)]; sm = [s[i] - s[i + 6] for i in range(6)]
    a = [sp[i] + sp[i + 3] for i in range(3)]; b = [sp[i] - sp[i + 3] for i in range(3)]
    S = a[0] + a[1] + a[2]
    check53(*sp, *sm, *a, *b, S)
    UU = []
    for j in range(3):
        t = S * 16 + k[j]; check53(t); u = a[(j + 2) % 3] * 16 + t; check53(u); UU.append(u)
    UV = []
    for (c0, c1, c2, kk) in ((-1, -2, 8, k[3]), (-8, -1, -2, k[4]), (2, -8, -1, k[5])):
        t = kk
        for coef, x in ((c0, b[0]), (c1, b[1]), (c2, b[2])):
            t = t + coef * x; check53(t)
        UV.append(t)
    U = [UU[j] + UV[j] for j in range(3)] + [UU[j] - UV[j] for j in range(3)]
    V = []
    for n in range(6):
        acc = k[6 + n]
        for j in range(6):
            acc += sm[j] * (F6[n - j] if j <= n else -F6[6 + n - j]); check53(acc)
        V.append(acc)
    U[0] += s[0] * 4; V[0] += s[0] * 4
    o = [U[n] + V[n] for n in range(6)] + [U[n] - V[n] for n in range(6)]
    check53(*U, *V, *o)
    return o


def rint_div(x, sh):
    """round-to-nearest-even of x / 2^sh (what (x*2^-sh + 1.5*2^52) - 1.5*2^52 computes in fp64)"""
    q, r = divmod(x, 1 << sh)
    half = 1 << (sh - 1)
    if r > half or (r == half and (q & 1)):
        q += 1
    return q


def normalize(L, H):
    cL = rint_div(L, 32); L1 = L - (cL << 32); H1 = H + cL
    cH = rint_div(H1, 32); H2 = H1 - (cH << 32)
    return L1 - cH, H2 + cH


def sbox(x):
    return pow(x, 7, P)


def sbox_limbs(x, rnd=None):
    """x^7 as the two signed limbs the CUDA S-box hands to the fp64 MDS: the 128-bit product x^3 * x^4 = (z3:z2:z1:z0) is
    not reduced to 64 bits; with 2^64 = 2^32 - 1 and 2^96 = -1 its value is (z0 - z2 - z3) + (z1 + z2)*2^32."""
    x2 = x * x % P; x4 = x2 * x2 % P; x3 = x * x2 % P
    z = x3 * x4
    z0, z1, z2, z3 = (z >> 0) & 0xFFFFFFFF, (z >> 32) & 0xFFFFFFFF, (z >> 64) & 0xFFFFFFFF, z >> 96
    lo, hi = z0 - z2 - z3, z1 + z2
    assert (lo + (hi << 32)) % P == pow(x, 7, P)
    return lo, hi


def permute_v3(state, rc, stats=None):
    """Mirror of csrc/poseidon.cuh · permute (v3) on exact integers."""
    lane0_c, tail_c = partial_constants(rc)
    s = [(state[i] + rc[i]) % P for i in range(12)]

    def full_layer(lo, hi, r):
        lanes = [rc[12 * (r + 1) + i] if r < 29 else 0 for i in range(12)]
        limbs = [rc_limbs(c) for c in lanes]
        out = []
        for limb, ins, off in ((0, lo, OFF_LO), (1, hi, OFF_HI)):
            k = [int(x) for x in fold_constants([l[limb] + BIAS + off for l in limbs])]
            o = mds_limb_checked(ins, k)
            assert all(BIAS <= x < 2 * BIAS for x in o)
            out.append([x - BIAS for x in o])
        return [(out[0][i] + (out[1][i] << 32)) % P for i in range(12)]

    def full_round(s, r):
        lo, hi = zip(*[sbox_limbs(x) for x in s])
        return full_layer(list(lo), list(hi), r)

    for r in range(0, 4):
        s = full_round(s, r)
    x0 = s[0]
    L = [v & 0xFFFFFFFF for v in s]; H = [v >> 32 for v in s]          # lanes 1..11 resident (index 0 unused)
    for r in range(4, 26):
        L[0], H[0] = sbox_limbs(x0)
        c0 = lane0_c[r + 1]
        cl, ch = rc_limbs(c0)
        outs = []
        for limb, ins, cc, off in ((0, L, cl, OFF_LO), (1, H, ch, OFF_HI)):
            q = (cc + BIAS + off) // 4
            assert (cc + BIAS + off) % 4 == 0
            k = [q, 0, 0, q, 0, 0, 2 * q, 0, 0, 0, 0, 0]
            o = mds_limb_checked(ins, k)
            assert BIAS <= o[0] < 2 * BIAS, "lane 0 read-out out of the mantissa window"
            outs.append(o)
        x0 = ((outs[0][0] - BIAS) + ((outs[1][0] - BIAS) << 32)) % P       # recombine (offsets are = 0 mod p)
        L, H = outs[0], outs[1]
        if stats is not None:
            stats["max_unnorm"] = max(stats.get("max_unnorm", 0), max(abs(v) for v in L[1:] + H[1:]))
        if r & 1:
            for i in range(1, 12):
                L[i], H[i] = normalize(L[i], H[i])
                assert abs(L[i]) <= (1 << 31) + (1 << 21) and abs(H[i]) <= (1 << 31) + (1 << 21)
    # leave the fp64 domain: lanes 1..11 get tail constant + offset + bias, then recombine
    s = [x0]
    for i in range(1, 12):
        cl, ch = rc_limbs(tail_c[i - 1])
        al = L[i] + cl + OFF_LO + BIAS; ah = H[i] + ch + OFF_HI + BIAS
        check53(al, ah)
        assert BIAS <= al < 2 * BIAS and BIAS <= ah < 2 * BIAS
        s.append(((al - BIAS) + ((ah - BIAS) << 32)) % P)
    for r in range(26, 30):
        s = full_round(s, r)
    return s


def permute_naive(state, rc):
    s = list(state)
    for r in range(30):
        s = [(s[i] + rc[12 * r + i]) % P for i in range(12)]
        if r < 4 or r >= 26:
            s = [sbox(x) for x in s]
        else:
            s[0] = sbox(s[0])
        s = mds_field(s)
    return s


def selftest_v3():
    import os, sys, random
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from gen_poseidon_constants import constants
    rc = constants()
    rnd = random.Random(5)
    stats = {}
    tests = [[0] * 12, list(range(12)), [P - 1] * 12] + [[rnd.randrange(P) for _ in range(12)] for _ in range(40)]
    for st in tests:
        assert permute_v3(st, rc, stats) == permute_naive(st, rc)
    print("permute_v3 model ok; max |unnormalised resident limb| = 2^%.2f" % (stats["max_unnorm"].bit_length()))


if __name__ == "__main__":
    selftest_v3()


# ======================================================================================================================
# v4: partial rounds in the CRT ("frequency") domain of the circulant
# ======================================================================================================================
# With a_i, b_i, m_i as in mds_limb (a = cyclic-3 part, b = negacyclic-3 part, m = negacyclic-6 part of one limb vector),
# the MDS layer acts blockwise:  a' = 4*UU(a), b' = 4*UV(b), m' = 2*V(m)  — the 30 add/sub butterflies of the time-domain
# layer disappear when the 11 resident lanes stay in this domain for all 22 partial rounds.  Lane 0 is read out as
# (a0 + b0 + 2*m0)/4 (an integer identity) and written back by adding (y - lane0) to a0, b0, m0; the diagonal 8*s0 reaches
# a0', b0', m0'.  Re-normalisation carries are multiples of 4 so that the division by 4 stays exact on each limb.
OFF4_LO = (1 << 51) + (1 << 19)          # OFF4_LO + OFF4_HI*2^32 = 2^19 * p
OFF4_HI = (1 << 51) - (1 << 20)
assert OFF4_LO + (OFF4_HI << 32) == (1 << 19) * P


class BV:
    """exact integer value + proven bound on |value| (bounds are propagated through every operation: the asserts are a
    proof that no intermediate reaches 2^53, not a test on sample inputs)."""
    __slots__ = ("v", "m")

    def __init__(self, v, m):
        assert abs(v) <= m, (v, m)
        assert m < (1 << 53), "fp64 exactness bound violated: 2^%.2f" % (m.bit_length())
        self.v, self.m = v, m

    def __add__(self, o): return BV(self.v + o.v, self.m + o.m)
    def __sub__(self, o): return BV(self.v - o.v, self.m + o.m)
    def scale(self, c): return BV(self.v * c, self.m * abs(c))
    def fma(self, c, addend): return BV(self.v * c + addend.v, self.m * abs(c) + addend.m)   # self*c + addend


def bv_const(c):
    return BV(c, abs(c))


def freq_forward(s):
    sp = [s[i] + s[i + 6] for i in range(6)]; m = [s[i] - s[i + 6] for i in range(6)]
    a = [sp[i] + sp[i + 3] for i in range(3)]; b = [sp[i] - sp[i + 3] for i in range(3)]
    return a, b, m


NEG6 = [[(F6[n - j] if j <= n else -F6[6 + n - j]) for j in range(6)] for n in range(6)]   # V[n] = sum_j m[j]*NEG6[n][j]


def freq_round(a, b, m, y, e):
    """One partial round on one limb: lane 0 (raw value e) is replaced by y, then the MDS layer.  Returns a', b', m'.
    Mirrors csrc/poseidon.cuh · freq_round operation by operation (the component that carries y is last in every chain,
    so everything else can be issued while the S-box is still running)."""
    d = y - e
    a0 = a[0] + d; b0 = b[0] + d; m0 = m[0] + d
    t1 = (a[1] + a[2]).scale(64)
    na = [y.fma(8, a0.fma(64, a[2].fma(64, t1))),
          a0.fma(128, t1),
          a0.fma(64, a[1].fma(64, t1))]
    nb = [y.fma(8, b0.fma(-4, b[2].fma(32, b[1].scale(-8)))),
          b0.fma(-32, b[1].fma(-4, b[2].scale(-8))),
          b0.fma(8, b[1].fma(-32, b[2].scale(-4)))]
    nm = []
    for n in range(6):
        acc = m[1].scale(2 * NEG6[n][1])
        for j in range(2, 6):
            acc = m[j].fma(2 * NEG6[n][j], acc)
        acc = m0.fma(2 * NEG6[n][0], acc)
        if n == 0:
            acc = y.fma(8, acc)
        nm.append(acc)
    return na, nb, nm


def freq_lane0(a, b, m):
    """raw lane-0 value (a0 + b0 + 2*m0)/4: fma(a0 + b0, 0.25, 0.5*m0) — 0.5*m0 is an exact half-integer, the sum an integer"""
    e4 = a[0].v + b[0].v + 2 * m[0].v
    assert e4 % 4 == 0
    return BV(e4 // 4, (a[0].m + b[0].m + 2 * m[0].m + 3) // 4)


def rint4_div32(x):
    """x / 2^32 rounded to a multiple of 4 (what (x*2^-32 + 1.5*2^54) - 1.5*2^54 computes in fp64: ulp at 2^54 is 4)"""
    return 4 * rint_div(x, 34)


def freq_normalize(L, H):
    """(L, H) with value L + H*2^32 -> same value mod p, carries are multiples of 4; bounds are proven, not sampled:
    |L - cL*2^32| <= 2^33, |cL| <= |L|/2^32 + 2, then |H1 - cH*2^32| <= 2^33 and |cH| <= |H1|/2^32 + 2."""
    cL = rint4_div32(L.v); cLm = (L.m >> 32) + 2
    L1 = BV(L.v - (cL << 32), 1 << 33)
    H1 = BV(H.v + cL, H.m + cLm)
    cH = rint4_div32(H1.v); cHm = (H1.m >> 32) + 2
    H2 = BV(H1.v - (cH << 32) + cH, (1 << 33) + cHm)
    L2 = BV(L1.v - cH, L1.m + cHm)
    assert (L2.v + (H2.v << 32) - L.v - (H.v << 32)) % P == 0
    return L2, H2


def sbox_limbs_bv(x):
    lo, hi = sbox_limbs(x)
    return BV(lo, 1 << 33), BV(hi, 1 << 33)      # lo = z0 - z2 - z3 in (-2^33, 2^32), hi = z1 + z2 < 2^33


def permute_v4(state, rc, stats=None):
    """Mirror of csrc/poseidon.cuh · permute (v4) on exact integers with proven magnitude bounds."""
    lane0_c, tail_c = partial_constants(rc)
    s = [(state[i] + rc[i]) % P for i in range(12)]

    def full_layer(lo, hi, r):
        lanes = [rc[12 * (r + 1) + i] if r < 29 else 0 for i in range(12)]
        limbs = [rc_limbs(c) for c in lanes]
        out = []
        for limb, ins, off in ((0, lo, OFF_LO), (1, hi, OFF_HI)):
            k = [int(x) for x in fold_constants([l[limb] + BIAS + off for l in limbs])]
            o = mds_limb_checked(ins, k)
            assert all(BIAS <= x < 2 * BIAS for x in o)
            out.append([x - BIAS for x in o])
        return [(out[0][i] + (out[1][i] << 32)) % P for i in range(12)]

    def full_round(s, r):
        lo, hi = zip(*[sbox_limbs(x) for x in s])
        return full_layer(list(lo), list(hi), r)

    for r in range(0, 4):
        s = full_round(s, r)
    x0 = s[0]
    zero = BV(0, 0)
    fr = []                                   # per limb: [a, b, m]
    for limb in (0, 1):
        lanes = [zero] + [BV((v >> (32 * limb)) & 0xFFFFFFFF, (1 << 32) - 1) for v in s[1:]]
        fr.append(list(freq_forward(lanes)))
    e = [zero, zero]                          # raw lane-0 value of each limb (lane 0 is 0 in the initial transform)
    for r in range(4, 26):
        y = sbox_limbs_bv(x0)
        for limb in (0, 1):
            a, b, m = fr[limb]
            fr[limb] = list(freq_round(a, b, m, y[limb], e[limb]))
        if stats is not None:
            stats["max_unnorm"] = max(stats.get("max_unnorm", 0), max(x.m for l in fr for part in l for x in part))
        if r & 1:
            for part in range(3):
                for i in range(len(fr[0][part])):
                    fr[0][part][i], fr[1][part][i] = freq_normalize(fr[0][part][i], fr[1][part][i])
        xs = []
        for limb in (0, 1):
            e[limb] = freq_lane0(*fr[limb])
            c = rc_limbs(lane0_c[r + 1])[limb]
            X = e[limb] + bv_const(c + BIAS + (OFF4_LO, OFF4_HI)[limb])
            assert BIAS <= X.v < 2 * BIAS and e[limb].m < (1 << 51), "lane 0 read-out window"
            xs.append(X.v - BIAS)
        x0 = (xs[0] + (xs[1] << 32)) % P
    # leave the fp64 domain: inverse transform (exact: every component is F*(integer vector) + multiples of 4)
    out = [x0]
    lanes = [[None] * 12, [None] * 12]
    for limb in (0, 1):
        a, b, m = fr[limb]
        for i in range(3):
            tp, tm = a[i] + b[i], a[i] - b[i]
            for (t, j) in ((tp, i), (tm, i + 3)):
                for sign, lane in ((1, j), (-1, j + 6)):
                    v4 = t.v + 2 * sign * m[j].v
                    assert v4 % 4 == 0
                    lanes[limb][lane] = BV(v4 // 4, (t.m + 2 * m[j].m + 3) // 4)
    for i in range(1, 12):
        cl, ch = rc_limbs(tail_c[i - 1])
        al = lanes[0][i] + bv_const(cl + OFF4_LO + BIAS); ah = lanes[1][i] + bv_const(ch + OFF4_HI + BIAS)
        assert BIAS <= al.v < 2 * BIAS and BIAS <= ah.v < 2 * BIAS
        assert lanes[0][i].m < (1 << 51) and lanes[1][i].m < (1 << 51)
        out.append(((al.v - BIAS) + ((ah.v - BIAS) << 32)) % P)
    s = out
    for r in range(26, 30):
        s = full_round(s, r)
    return s


# ======================================================================================================================
# v5: the negacyclic-6 part is split once more, x^6+1 = (x^2+1)(x^4-x^2+1): p = m mod (x^2+1), q = m mod (x^4-x^2+1)
# ======================================================================================================================
# m' = m*G mod (x^6+1) with G = 2*[2,-4,16,1,-1,-1] becomes p' = p*(-30-12x) mod (x^2+1) (4 products) and
# q' = q*(6-6x+30x^2) mod (x^4-x^2+1) (13 products) instead of 36.  The CRT back-map divides by 3
# (3 m0 = p0 + 2 q0 + q2, ...), done exactly by rounding t/6 to the nearest half-integer; carries of p and q are
# multiples of 12 so that every limb stays divisible.
from fractions import Fraction as _Fr
C_INV_12_2P32 = 1.0 / (12.0 * 4294967296.0)      # the fp64 constant the CUDA code multiplies by
C_INV_6 = 1.0 / 6.0


def m_to_pq(m):
    return [m[0] - m[2] + m[4], m[1] - m[3] + m[5]], [m[0] - m[4], m[1] - m[5], m[2] + m[4], m[3] + m[5]]


def rint_half_div6(t):
    """what (t*fl(1/6) + 1.5*2^51) - 1.5*2^51 computes: t/6 rounded to a multiple of 1/2, returned as 2*value (an int).
    Exact whenever t is a multiple of 3 and |t| < 2^52 (error of the product < 1/4)."""
    x = _Fr(t.v) * _Fr(C_INV_6) * 2
    k = round(x)                                   # nearest integer of 2*t/6 (ties cannot occur for multiples of 3)
    assert t.v % 3 == 0 and k == t.v // 3 and abs(t.m) < (1 << 52)
    return BV(k, (t.m + 2) // 3)                   # = m0 (so that m0/2 is the half-integer)


def pq_lane0_m0(pp, q):
    t = q[0].fma(2, pp[0] + q[2])                  # 3*m0
    return rint_half_div6(t)


def freq5_round(a, b, pp, q, y, e):
    d = y - e
    a0 = a[0] + d; b0 = b[0] + d; p0 = pp[0] + d; q0 = q[0] + d
    t1 = (a[1] + a[2]).scale(64)
    na = [y.fma(8, a0.fma(64, a[2].fma(64, t1))), a0.fma(128, t1), a0.fma(64, a[1].fma(64, t1))]
    nb = [y.fma(8, b0.fma(-4, b[2].fma(32, b[1].scale(-8)))),
          b0.fma(-32, b[1].fma(-4, b[2].scale(-8))),
          b0.fma(8, b[1].fma(-32, b[2].scale(-4)))]
    np_ = [y.fma(8, p0.fma(-30, pp[1].scale(12))), p0.fma(-12, pp[1].scale(-30))]
    nq = [y.fma(8, q0.fma(6, q[2].fma(-30, q[3].scale(6)))),
          q0.fma(-6, q[1].fma(6, q[3].scale(-30))),
          q0.fma(30, q[1].fma(-6, q[2].fma(36, q[3].scale(-6)))),
          q[1].fma(30, q[2].fma(-6, q[3].scale(36)))]
    return na, nb, np_, nq


def freq5_lane0(a, b, pp, q):
    m0 = pq_lane0_m0(pp, q)
    e4 = a[0].v + b[0].v + 2 * m0.v
    assert e4 % 4 == 0
    return BV(e4 // 4, (a[0].m + b[0].m + 2 * m0.m + 3) // 4)


def normalize_k(L, H, mult):
    """carries are multiples of `mult` (4: magic rounding at 2^54; 12: k = rint(L*fl(2^-32/12)), carry 12k)"""
    def carry(x):
        if mult == 4:
            return rint4_div32(x.v), (x.m >> 32) + 2
        k = round(_Fr(x.v) * _Fr(C_INV_12_2P32))
        return 12 * k, (x.m >> 32) + 12
    half = (mult // 2) << 32
    cL, cLm = carry(L)
    L1 = BV(L.v - (cL << 32), half + (1 << 12)); H1 = BV(H.v + cL, H.m + cLm)
    cH, cHm = carry(H1)
    H2 = BV(H1.v - (cH << 32) + cH, half + (1 << 12) + cHm)
    L2 = BV(L1.v - cH, L1.m + cHm)
    assert (L2.v + (H2.v << 32) - L.v - (H.v << 32)) % P == 0
    return L2, H2


def permute_v5(state, rc, stats=None):
    """Mirror of csrc/poseidon.cuh · permute (v5) on exact integers with proven magnitude bounds."""
    lane0_c, tail_c = partial_constants(rc)
    s = [(state[i] + rc[i]) % P for i in range(12)]

    def full_round(s, r):
        lo, hi = zip(*[sbox_limbs(x) for x in s])
        lanes = [rc[12 * (r + 1) + i] if r < 29 else 0 for i in range(12)]
        limbs = [rc_limbs(c) for c in lanes]
        out = []
        for limb, ins, off in ((0, list(lo), OFF_LO), (1, list(hi), OFF_HI)):
            k = [int(x) for x in fold_constants([l[limb] + BIAS + off for l in limbs])]
            o = mds_limb_checked(ins, k)
            assert all(BIAS <= x < 2 * BIAS for x in o)
            out.append([x - BIAS for x in o])
        return [(out[0][i] + (out[1][i] << 32)) % P for i in range(12)]

    for r in range(0, 4):
        s = full_round(s, r)
    x0 = s[0]
    zero = BV(0, 0)
    fr = []
    for limb in (0, 1):
        lanes = [zero] + [BV((v >> (32 * limb)) & 0xFFFFFFFF, (1 << 32) - 1) for v in s[1:]]
        a, b, m = freq_forward(lanes)
        pp, q = m_to_pq(m)
        fr.append([a, b, pp, q])
    e = [zero, zero]
    for r in range(4, 26):
        y = sbox_limbs_bv(x0)
        for limb in (0, 1):
            fr[limb] = list(freq5_round(*fr[limb], y[limb], e[limb]))
        if stats is not None:
            stats["max_unnorm"] = max(stats.get("max_unnorm", 0), max(x.m for l in fr for part in l for x in part))
        if r & 1:
            for part, mult in ((0, 4), (1, 4), (2, 12), (3, 12)):
                for i in range(len(fr[0][part])):
                    fr[0][part][i], fr[1][part][i] = normalize_k(fr[0][part][i], fr[1][part][i], mult)
        xs = []
        for limb in (0, 1):
            e[limb] = freq5_lane0(*fr[limb])
            c = rc_limbs(lane0_c[r + 1])[limb]
            X = e[limb] + bv_const(c + BIAS + (OFF4_LO, OFF4_HI)[limb])
            assert BIAS <= X.v < 2 * BIAS and e[limb].m < (1 << 51), "lane 0 read-out window"
            xs.append(X.v - BIAS)
        x0 = (xs[0] + (xs[1] << 32)) % P
    out = [x0]
    lanes = [[None] * 12, [None] * 12]
    for limb in (0, 1):
        a, b, pp, q = fr[limb]
        # 3*m_j from the CRT back-map, then m_j = rint_half(t/6)*2
        t3 = [q[0].fma(2, pp[0] + q[2]), q[1].fma(2, pp[1] + q[3]), q[2].fma(2, q[0] - pp[0]),
              q[3].fma(2, q[1] - pp[1]), (pp[0] - q[0]) + q[2], (pp[1] - q[1]) + q[3]]
        m = [rint_half_div6(t) for t in t3]
        for i in range(3):
            tp, tm = a[i] + b[i], a[i] - b[i]
            for (t, j) in ((tp, i), (tm, i + 3)):
                for sign, lane in ((1, j), (-1, j + 6)):
                    v4 = t.v + 2 * sign * m[j].v
                    assert v4 % 4 == 0
                    lanes[limb][lane] = BV(v4 // 4, (t.m + 2 * m[j].m + 3) // 4)
    for i in range(1, 12):
        cl, ch = rc_limbs(tail_c[i - 1])
        al = lanes[0][i] + bv_const(cl + OFF4_LO + BIAS); ah = lanes[1][i] + bv_const(ch + OFF4_HI + BIAS)
        assert BIAS <= al.v < 2 * BIAS and BIAS <= ah.v < 2 * BIAS
        assert lanes[0][i].m < (1 << 51) and lanes[1][i].m < (1 << 51)
        out.append(((al.v - BIAS) + ((ah.v - BIAS) << 32)) % P)
    s = out
    for r in range(26, 30):
        s = full_round(s, r)
    return s


def selftest_v5():
    import os, sys, random
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from gen_poseidon_constants import constants
    rc = constants()
    rnd = random.Random(9)
    # one round of the split form against the time-domain layer (no normalisation involved)
    for _ in range(50):
        sv = [rnd.randrange(-(1 << 33), 1 << 33) for _ in range(12)]
        yv = rnd.randrange(-(1 << 33), 1 << 33)
        a, b, m = freq_forward([BV(v, 1 << 33) for v in sv])
        pp, q = m_to_pq(m)
        na, nb, np_, nq = freq5_round(a, b, pp, q, BV(yv, 1 << 33), BV(sv[0], 1 << 33))
        want = mds_direct([yv] + sv[1:])
        wa, wb, wm = freq_forward([BV(v, abs(v) + 1) for v in want])
        wp, wq = m_to_pq(wm)
        assert [x.v for x in na + nb + np_ + nq] == [x.v for x in wa + wb + wp + wq]
    stats = {}
    tests = [[0] * 12, list(range(12)), [P - 1] * 12] + [[rnd.randrange(P) for _ in range(60)]]
    tests = tests[:3] + [[rnd.randrange(P) for _ in range(12)] for _ in range(60)]
    for st in tests:
        assert permute_v5(st, rc, stats) == permute_naive(st, rc)
    print("permute_v5 model ok; proven bound on resident components = 2^%.2f" % (stats["max_unnorm"].bit_length()))


def selftest_v4():
    import os, sys, random
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from gen_poseidon_constants import constants
    rc = constants()
    rnd = random.Random(7)
    stats = {}
    tests = [[0] * 12, list(range(12)), [P - 1] * 12] + [[rnd.randrange(P) for _ in range(12)] for _ in range(60)]
    for st in tests:
        assert permute_v4(st, rc, stats) == permute_naive(st, rc)
    print("permute_v4 model ok; proven bound on resident components = 2^%.2f" % (stats["max_unnorm"].bit_length()))


if __name__ == "__main__":
    selftest_v4()
    selftest_v5()
