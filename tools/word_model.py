#!/usr/bin/env python3
"""32-bit word-level model of the carry logic in csrc/gl_field.cuh (mul) and csrc/poseidon.cuh (recombine).
Each function mirrors the PTX instruction by instruction (add.cc/addc/sub.cc/subc on u32 + one carry flag) so the
fix-up algebra can be checked against Python big ints without a GPU."""
import random

B = 1 << 32
M32 = B - 1
P = 0xFFFFFFFF00000001
EPS = M32


class Flags:
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


def mul_model(a, b):
    """gl::mul — 4 wide products, 5 carry adds, reduce (z3:z2:z1:z0) as X - z3 + z2*EPS with one signed fix-up."""
    a0, a1, b0, b1 = a & M32, a >> 32, b & M32, b >> 32
    p00, p01, p10, p11 = a0 * b0, a0 * b1, a1 * b0, a1 * b1
    z0, z1 = p00 & M32, p00 >> 32
    q0, q1, m0, m1, h0, h1 = p01 & M32, p01 >> 32, p10 & M32, p10 >> 32, p11 & M32, p11 >> 32
    f = Flags()
    z1 = add_cc(f, z1, q0); z2 = addc_cc(f, h0, q1); z3 = addc(f, h1, 0)
    z1 = add_cc(f, z1, m0); z2 = addc_cc(f, z2, m1); z3 = addc(f, z3, 0)
    assert (z0 | z1 << 32 | z2 << 64 | z3 << 96) == a * b
    return reduce_model(z0, z1, z2, z3)


def reduce_model(z0, z1, z2, z3):
    f = Flags()
    l = sub_cc(f, z0, z3); h = subc_cc(f, z1, 0); d = subc(f, 0, 0)          # d = -borrow
    # + z2*EPS = + z2*B - z2
    l = sub_cc(f, l, z2); h = subc_cc(f, h, 0); d = subc(f, d, 0)            # d -= borrow
    h = add_cc(f, h, z2); d = addc(f, d, 0)                                  # d += carry   -> d in {-1,0,1} (mod 2^32)
    assert d in (0, 1, M32), d
    # apply d*EPS = d*B - d  (d sign-extended)
    ds = M32 if d >> 31 else 0
    l = sub_cc(f, l, d); h = subc(f, h, ds)
    h = (h + d) & M32
    return l | h << 32


def add_any_model(a, b):
    f = Flags()
    l = add_cc(f, a & M32, b & M32); h = addc_cc(f, a >> 32, b >> 32); m = (-addc(f, 0, 0)) & M32
    l = add_cc(f, l, m); h = addc_cc(f, h, 0); m = (-addc(f, 0, 0)) & M32
    l = add_cc(f, l, m); h = addc_cc(f, h, 0)
    assert f.cf == 0
    return l | h << 32


def sub_any_model(a, b):
    f = Flags()
    l = sub_cc(f, a & M32, b & M32); h = subc_cc(f, a >> 32, b >> 32); m = subc(f, 0, 0)
    l = sub_cc(f, l, m); h = subc_cc(f, h, 0); m = subc(f, 0, 0)
    l = sub_cc(f, l, m); h = subc_cc(f, h, 0)
    assert f.cf == 0
    return l | h << 32


def mul_2_24_model(x):
    x0, x1 = x & M32, x >> 32
    return reduce_model((x0 << 24) & M32, ((x1 << 24) | (x0 >> 8)) & M32, x1 >> 8, 0)


def mul_2_48_model(x):
    x0, x1 = x & M32, x >> 32
    return reduce_model(0, (x0 << 16) & M32, ((x1 << 16) | (x0 >> 16)) & M32, x1 >> 16)


def recombine_model(al, ah):
    """value = al + ah*B (al, ah < 2^52 as they sit in the mantissa of 2^52 + x) -> 64-bit 'any'."""
    a0, a1, b0, b1 = al & M32, al >> 32, ah & M32, ah >> 32
    f = Flags()
    u = a1 + b1
    g = add_cc(f, u, b0); T = addc(f, b1, 0); g2 = addc(f, g, 0)
    lo = sub_cc(f, a0, T); hi = subc(f, g2, 0)
    return lo | hi << 32


if __name__ == "__main__":
    rnd = random.Random(7)
    edge = [0, 1, 2, M32 - 1, M32, B, B + 1, P - 1, P, P + 1, (1 << 64) - 1, (1 << 64) - 2, EPS << 32, (EPS << 32) | 1, 1 << 63]
    cases = [(a, b) for a in edge for b in edge] + [(rnd.getrandbits(64), rnd.getrandbits(64)) for _ in range(200000)]
    cases += [(rnd.choice(edge) ^ rnd.getrandbits(3), rnd.getrandbits(64)) for _ in range(20000)]
    for a, b in cases:
        r = mul_model(a, b)
        assert r < (1 << 64) and r % P == a * b % P, (hex(a), hex(b), hex(r))
    # reduce on adversarial words
    for _ in range(200000):
        z = [rnd.choice([0, 1, 2, M32, M32 - 1, rnd.getrandbits(32)]) for _ in range(4)]
        if (z[2] | z[3] << 32) > (M32 - 1) << 32 | 1: continue      # product of two u64 is < 2^128 - 2^65 + 1: hi <= (2^64-2)
        r = reduce_model(*z)
        assert r < (1 << 64) and r % P == (z[0] + (z[1] << 32) + (z[2] << 64) + (z[3] << 96)) % P, z
    for _ in range(300000):
        al = rnd.choice([0, 1, M32, B, (1 << 52) - 1, rnd.getrandbits(52), rnd.getrandbits(42), rnd.getrandbits(33)])
        ah = rnd.choice([0, 1, M32, B, (1 << 52) - 1, rnd.getrandbits(52), rnd.getrandbits(42), rnd.getrandbits(33)])
        r = recombine_model(al, ah)
        assert r < (1 << 64) and r % P == (al + (ah << 32)) % P, (hex(al), hex(ah))
    for a, b in cases:
        for fn, want in ((add_any_model, a + b), (sub_any_model, a - b)):
            r = fn(a, b)
            assert 0 <= r < (1 << 64) and r % P == want % P, (fn.__name__, hex(a), hex(b), hex(r))
        for fn, sh in ((mul_2_24_model, 24), (mul_2_48_model, 48)):
            r = fn(a)
            assert 0 <= r < (1 << 64) and r % P == (a << sh) % P, (fn.__name__, hex(a))
    assert (0xFFFFFFFF00 - pow(2, 72, P)) % P == 0
    print("word model ok")
