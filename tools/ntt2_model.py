#!/usr/bin/env python3
"""Word-level model + bound proof of the carry-save butterflies in csrc/ntt2.cuh (no GPU needed).

A value in flight is three 32-bit words (l, h, t) = l + h*2^32 + t*2^64, UNSIGNED, t small: an addition or subtraction is then
three carry-chain instructions (add.cc / addc.cc / addc) instead of the 8-10 of a fully folded 64-bit modular add.  Subtractions
stay non-negative because a few inputs are biased by a multiple of p up front (k*p = (k, 2^32 - k, k - 1) as words); values are
folded back to 64 bits ("any" representative) once per radix-8 / radix-16 step, or implicitly by the shift-multiplies, which
accept three-word inputs.

Every function mirrors the PTX of ntt2.cuh instruction by instruction (u32 words + one carry flag).  Run as a script it
  1. checks fold3 / reduce3 / the seven shift-multiplies against Python big integers on random and adversarial words,
  2. PROVES by interval propagation that with the bias constants below no subtraction goes negative, no third word exceeds its
     bit budget (the funnel shifts of the shift-multiplies need t < 2^(32 - r)), and reduce_words' precondition holds,
  3. checks dft8_cs / dft16_cs against the DFT definition with plonky2's roots (w_8 = 2^120, w_16 = 2^156) on random inputs.
"""
import random

B = 1 << 32
M32 = B - 1
P = 0xFFFFFFFF00000001


class F:
    cf = 0


def add_cc(f, a, b):
    s = a + b; f.cf = s >> 32; return s & M32
def addc_cc(f, a, b):
    s = a + b + f.cf; f.cf = s >> 32; return s & M32
def addc(f, a, b):
    return (a + b + f.cf) & M32
def sub_cc(f, a, b):
    d = a - b; f.cf = 1 if d < 0 else 0; return d & M32
def subc_cc(f, a, b):
    d = a - b - f.cf; f.cf = 1 if d < 0 else 0; return d & M32
def subc(f, a, b):
    return (a - b - f.cf) & M32
def funnel_l(lo, hi, r):            # __funnelshift_l(lo, hi, r): upper 32 bits of (hi:lo) << r, 0 < r < 32
    return (((hi << 32) | lo) << r >> 32) & M32


# ---- three-word values -------------------------------------------------------------------------------------------------------
def w3(x):                          # 64-bit -> (l, h, 0)
    return (x & M32, x >> 32, 0)
def val(a):
    return a[0] + (a[1] << 32) + (a[2] << 64)
def add3(a, b):
    f = F()
    l = add_cc(f, a[0], b[0]); h = addc_cc(f, a[1], b[1]); t = addc(f, a[2], b[2])
    assert val(a) + val(b) == val((l, h, t)), "third word overflow"
    return (l, h, t)
def sub3(a, b):
    f = F()
    l = sub_cc(f, a[0], b[0]); h = subc_cc(f, a[1], b[1]); t = subc(f, a[2], b[2])
    assert val(a) - val(b) == val((l, h, t)), "negative difference"
    return (l, h, t)
def bias3(a, k):                    # a + k*p, k*p = k + (2^32 - k)*2^32 + (k - 1)*2^64
    f = F()
    l = add_cc(f, a[0], k); h = addc_cc(f, a[1], B - k); t = addc(f, a[2], k - 1)
    assert val((l, h, t)) == val(a) + k * P
    return (l, h, t)


def reduce_words(z0, z1, z2, z3):   # gl_field.cuh · reduce_words (unchanged): needs (z3:z2) <= 2^64 - 2
    f = F()
    l = sub_cc(f, z0, z3); h = subc_cc(f, z1, 0); d = subc(f, 0, 0)
    l = sub_cc(f, l, z2); h = subc_cc(f, h, 0); d = subc(f, d, 0)
    h = add_cc(f, h, z2); d = addc(f, d, 0)
    assert d in (0, 1, M32)
    ds = M32 if d >> 31 else 0
    l = sub_cc(f, l, d); h = subc(f, h, ds)
    h = (h + d) & M32
    return l | h << 32


def reduce3(z0, z1, z2):            # (z2:z1:z0) -> 64 bits: X - z2 + z2*2^32, 9 instructions
    f = F()
    l = sub_cc(f, z0, z2); h = subc_cc(f, z1, 0); d = subc(f, 0, 0)           # d = -borrow
    h = add_cc(f, h, z2); d = addc(f, d, 0)                                   # d in {-1, 0, 1}
    assert d in (0, 1, M32)
    ds = M32 if d >> 31 else 0
    l = sub_cc(f, l, d); h = subc(f, h, ds)
    h = (h + d) & M32
    return l | h << 32


def fold3(a):                       # (l, h, t), t < 2^31 -> 64-bit "any": X + t*(2^32 - 1), 8 instructions
    l, h, t = a
    f = F()
    u0 = sub_cc(f, 0, t); u1 = subc(f, t, 0)                                  # u = t*2^32 - t
    l = add_cc(f, l, u0); h = addc_cc(f, h, u1); c = addc(f, 0, 0)
    m = (-c) & M32                                                            # c*(2^32 - 1) as a word
    l = add_cc(f, l, m); h = addc_cc(f, h, 0)
    assert f.cf == 0, "second wrap"
    return l | h << 32


def shift_mul(a, s):                # a * 2^s mod p -> 64-bit "any"; a three-word, s in {12, 24, 36, 48, 60, 72, 84}
    l, h, t = a
    r, w = s % 32, s // 32
    assert 0 < r < 32 and t < (1 << (32 - r)), ("third word too large for the funnel shift", s, t)
    y0, y1, y2 = (l << r) & M32, funnel_l(l, h, r), funnel_l(h, t, r)
    assert (y0 | y1 << 32 | y2 << 64) == val(a) << r
    if w == 0:
        return reduce3(y0, y1, y2)
    if w == 1:
        return reduce_words(0, y0, y1, y2)
    # w == 2: y * 2^64 = y * 2^32 - y  (non-negative four-word difference)
    f = F()
    r0 = sub_cc(f, 0, y0); r1 = subc_cc(f, y0, y1); r2 = subc_cc(f, y1, y2); r3 = subc(f, y2, 0)
    assert (r0 | r1 << 32 | r2 << 64 | r3 << 96) == (val(a) << r) * (B - 1)
    assert (r2 | r3 << 32) <= (1 << 64) - 2
    return reduce_words(r0, r1, r2, r3)


# ---- bias plans (multiples of p added before the subtractions); checked by the interval pass below --------------------------------
# dft8_cs: inputs may be three-word (from the radix-2 level of dft16) with values < IN8 * 2^64
K8 = {"x0": 22, "x2": 5, "x5": 10, "x7": 5, "a5": 2}
K16 = {"x0": 2}          # dft16: x0 (feeds s0 and d0, the x0 inputs of the two dft8s); every other pair biases its minuend by 2p


def dft8_cs(x, track=None):
    """x: 8 three-word values (natural order).  Returns 8 64-bit words, in-place-DIF order: out[e] = sum_j x_j w8^(j*bitrev3(e)).
    w8 = 2^120 = -2^24, w4 = 2^48, w8^3 = 2^168 = -2^72."""
    x = list(x)
    x[0] = bias3(x[0], K8["x0"]); x[2] = bias3(x[2], K8["x2"]); x[5] = bias3(x[5], K8["x5"]); x[7] = bias3(x[7], K8["x7"])
    a = [None] * 8
    a[0] = add3(x[0], x[4]); a[4] = sub3(x[0], x[4])
    a[1] = add3(x[1], x[5]); a[5] = w3(shift_mul(sub3(x[5], x[1]), 24))          # (x1 - x5) * w8   = (x5 - x1) * 2^24
    a[2] = add3(x[2], x[6]); a[6] = w3(shift_mul(sub3(x[2], x[6]), 48))          # (x2 - x6) * w8^2
    a[3] = add3(x[3], x[7]); a[7] = w3(shift_mul(sub3(x[7], x[3]), 72))          # (x3 - x7) * w8^3 = (x7 - x3) * 2^72
    a[5] = bias3(a[5], K8["a5"])
    b = [None] * 8
    b[0] = add3(a[0], a[2]); b[2] = sub3(a[0], a[2])
    b[1] = add3(a[1], a[3]); b[3] = w3(shift_mul(sub3(a[1], a[3]), 48))
    b[4] = add3(a[4], a[6]); b[6] = sub3(a[4], a[6])
    b[5] = add3(a[5], a[7]); b[7] = w3(shift_mul(sub3(a[5], a[7]), 48))
    c = [add3(b[0], b[1]), sub3(b[0], b[1]), add3(b[2], b[3]), sub3(b[2], b[3]),
         add3(b[4], b[5]), sub3(b[4], b[5]), add3(b[6], b[7]), sub3(b[6], b[7])]
    if track is not None:
        track.append(max(v[2] for v in c))
    return [fold3(v) for v in c]


# w16^j = 2^(156 j mod 192): j=1: -2^60, 2: -2^24, 3: 2^84, 4: 2^48, 5: 2^12, 6: -2^72, 7: -2^36
W16 = {1: (60, True), 2: (24, True), 3: (84, False), 4: (48, False), 5: (12, False), 6: (72, True), 7: (36, True)}


def dft16_cs(x, track=None):
    """x: 16 64-bit words (natural order) -> 16 64-bit words, in-place-DIF order: out[e] = sum_j x_j w16^(j*bitrev4(e))."""
    x = [w3(v) for v in x]
    s, d = [None] * 8, [None] * 8
    x[0] = bias3(x[0], K16["x0"])
    s[0] = add3(x[0], x[8]); d[0] = sub3(x[0], x[8])
    for j in range(1, 8):
        sh, neg = W16[j]
        lo, hi = (j + 8, j) if neg else (j, j + 8)      # minuend, subtrahend
        x[lo] = bias3(x[lo], 2)
        s[j] = add3(x[j], x[j + 8])
        d[j] = w3(shift_mul(sub3(x[lo], x[hi]), sh))
    return dft8_cs(s, track) + dft8_cs(d, track)


# ---- interval proof -------------------------------------------------------------------------------------------------------------
class Iv:
    """closed integer interval; mirrors the operations above on bounds"""
    def __init__(self, lo, hi): self.lo, self.hi = lo, hi
    def __add__(self, o): return Iv(self.lo + o.lo, self.hi + o.hi)
    def __sub__(self, o):
        r = Iv(self.lo - o.hi, self.hi - o.lo)
        assert r.lo >= 0, "a subtraction can go negative: %d" % r.lo
        return r
    def bias(self, k): return Iv(self.lo + k * P, self.hi + k * P)
    def tbits(self): return (self.hi >> 64).bit_length()


ANY = Iv(0, (1 << 64) - 1)


def iv_shift(v, s):
    assert v.tbits() <= 32 - (s % 32), ("funnel budget", s, v.tbits())
    if s // 32 == 2:
        assert ((v.hi << (s % 32)) * (B - 1)) >> 64 <= (1 << 64) - 2
    return ANY


def prove_dft8(inp):
    x = list(inp)
    for name, j in (("x0", 0), ("x2", 2), ("x5", 5), ("x7", 7)):
        x[j] = x[j].bias(K8[name])
    a = [None] * 8
    a[0] = x[0] + x[4]; a[4] = x[0] - x[4]
    a[1] = x[1] + x[5]; a[5] = iv_shift(x[5] - x[1], 24)
    a[2] = x[2] + x[6]; a[6] = iv_shift(x[2] - x[6], 48)
    a[3] = x[3] + x[7]; a[7] = iv_shift(x[7] - x[3], 72)
    a[5] = a[5].bias(K8["a5"])
    b = [None] * 8
    b[0] = a[0] + a[2]; b[2] = a[0] - a[2]
    b[1] = a[1] + a[3]; b[3] = iv_shift(a[1] - a[3], 48)
    b[4] = a[4] + a[6]; b[6] = a[4] - a[6]
    b[5] = a[5] + a[7]; b[7] = iv_shift(a[5] - a[7], 48)
    c = [b[0] + b[1], b[0] - b[1], b[2] + b[3], b[2] - b[3], b[4] + b[5], b[4] - b[5], b[6] + b[7], b[6] - b[7]]
    tmax = max(v.tbits() for v in c)
    assert tmax <= 31
    return tmax


def prove():
    t8 = prove_dft8([ANY] * 8)                                   # radix-8 rounds: fresh 64-bit inputs
    x = [ANY] * 16
    s, d = [None] * 8, [None] * 8
    x0 = x[0].bias(K16["x0"])
    s[0] = x0 + x[8]; d[0] = x0 - x[8]
    for j in range(1, 8):
        sh, _ = W16[j]
        s[j] = x[j].bias(2) + x[j + 8]                           # whichever of the two was biased, the sum carries 2p
        d[j] = iv_shift(x[j].bias(2) - x[j + 8], sh)
    t16s = prove_dft8(s)
    t16d = prove_dft8(d)
    return t8, t16s, t16d


def dft_ref(x, n_log, root_log2):
    n = 1 << n_log
    w = pow(2, root_log2, P)
    out = []
    for e in range(n):
        k = int(format(e, "0%db" % n_log)[::-1], 2)
        out.append(sum(x[j] * pow(w, j * k, P) for j in range(n)) % P)
    return out


if __name__ == "__main__":
    rnd = random.Random(11)
    edge = [0, 1, 2, M32 - 1, M32, B, B + 1, P - 1, P, P + 1, (1 << 64) - 1, (1 << 64) - 2, M32 << 32, (M32 << 32) | 1, 1 << 63]
    words = lambda: rnd.choice([0, 1, 2, M32, M32 - 1, rnd.getrandbits(32), rnd.getrandbits(32)])
    for _ in range(200000):
        a = (words(), words(), rnd.choice([0, 1, 2, 3, 7, 15, rnd.getrandbits(4)]))
        r = fold3(a)
        assert r < (1 << 64) and r % P == val(a) % P
        a = (words(), words(), rnd.choice([0, 1, 2, 31, 63, 127, rnd.getrandbits(7)]))
        r = fold3(a)
        assert r < (1 << 64) and r % P == val(a) % P
        z = (words(), words(), words())
        r = reduce3(*z)
        assert r < (1 << 64) and r % P == (z[0] + (z[1] << 32) + (z[2] << 64)) % P
        for s in (12, 24, 36, 48, 60, 72, 84):
            tb = min(32 - s % 32, 7)
            a = (words(), words(), rnd.getrandbits(tb) if rnd.random() < 0.7 else (1 << tb) - 1)
            r = shift_mul(a, s)
            assert r < (1 << 64) and r % P == (val(a) << s) % P, (s, a)
    print("fold3 / reduce3 / shift_mul ok")
    print("interval proof: max third-word bits (dft8 fresh, dft16 sums, dft16 diffs) =", prove())
    assert pow(2, 120, P) == pow(1753635133440165772, 1 << 29, P) and pow(2, 156, P) == pow(1753635133440165772, 1 << 28, P)
    for it in range(3000):
        xs = [rnd.choice(edge) if rnd.random() < 0.3 else rnd.getrandbits(64) for _ in range(16)]
        got = dft8_cs([w3(v) for v in xs[:8]])
        assert [g % P for g in got] == dft_ref(xs[:8], 3, 120), "dft8"
        got = dft16_cs(xs)
        assert [g % P for g in got] == dft_ref(xs, 4, 156), "dft16"
    # worst-case words through the word-level code itself (all-ones inputs maximise the third words)
    dft16_cs([(1 << 64) - 1] * 16); dft16_cs([0] * 8 + [(1 << 64) - 1] * 8); dft16_cs([(1 << 64) - 1] * 8 + [0] * 8)
    print("dft8_cs / dft16_cs ok")
