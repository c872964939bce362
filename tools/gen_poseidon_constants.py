#!/usr/bin/env python3
"""Regenerate plonky2's 360 Poseidon round constants offline and emit them as C headers.

Recipe (plonky2/src/bin/generate_constants.rs @ 3de92d9, restated in SURVEY.md A.5):
ChaCha8Rng::seed_from_u64(0), then 360 x gen_range(0..p) with rand-0.8 widening-multiply rejection.
Writes  oracle/poseidon_constants.h            (checker side)
and     plonky2.5_b200/csrc/poseidon_constants.cuh  (product side: __constant__ tables).
The table is pinned by sha256 and by the Poseidon KATs at
/root/reference/src/common/poseidon2/poseidon2_goldilocks.rs:190-211.
"""
import hashlib, os, struct

P = 0xFFFFFFFF00000001
M64 = (1 << 64) - 1
SHA = "d2fcbb5be293c50ab4b1ddcd9c81005b12d689816a54c91a054f97f6588a20a8"


def seed_bytes(state):
    out = b""
    for _ in range(8):
        state = (state * 6364136223846793005 + 11634580027462260723) & M64
        xs = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
        rot = state >> 59
        out += (((xs >> rot) | (xs << ((32 - rot) & 31))) & 0xFFFFFFFF).to_bytes(4, "little")
    return out


def chacha8_block(key, counter):
    rotl = lambda v, c: ((v << c) | (v >> (32 - c))) & 0xFFFFFFFF
    init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574, *key, counter & 0xFFFFFFFF, counter >> 32, 0, 0]
    x = list(init)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = rotl(x[b] ^ x[c], 7)

    for _ in range(4):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(a + b) & 0xFFFFFFFF for a, b in zip(x, init)]


def constants():
    sb = seed_bytes(0)
    key = [int.from_bytes(sb[4 * i:4 * i + 4], "little") for i in range(8)]
    words, ctr, out = [], 0, []
    while len(out) < 360:
        if len(words) < 2:
            words += chacha8_block(key, ctr); ctr += 1
        v = words[0] | (words[1] << 32); words = words[2:]
        m = v * P
        if (m & M64) <= P - 1:      # zone = (p << lzcnt(p)) - 1 with lzcnt(p) = 0
            out.append(m >> 64)
    return out


def mds_tables(rc):
    """Folded fp64 constants of the MDS layer that follows round r (tools/mds_model.py): per layer 2 limbs x
    (uu0..2, uv0..2, v0..5).  The constant of lane i is RC[12(r+1)+i] (0 after the last round), split into two limbs
    that are multiples of 4 (4*limbs of rc/4 mod p) so that every folded value is an integer; 2^52 is added to every
    lane so that the sums leave the fp64 pipe as 2^52 + integer, i.e. with the integer in the mantissa bits, plus the
    positivity offsets OFF_LO / OFF_HI (OFF_LO + OFF_HI*2^32 = 2^17 p) because the S-box hands over a signed low limb."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from mds_model import fold_constants, rc_limbs, OFF_LO, OFF_HI
    rows = []
    for r in range(30):
        lanes = [rc[12 * (r + 1) + i] if r < 29 else 0 for i in range(12)]
        limbs = [rc_limbs(c) for c in lanes]
        row = []
        for limb in (0, 1):
            k = fold_constants([l[limb] + (1 << 52) + (OFF_LO, OFF_HI)[limb] for l in limbs])
            assert all(x.denominator == 1 and abs(x) < (1 << 53) for x in k)
            row += [int(x) for x in k]
        rows.append(row)
    txt = "// MDS_K[r][limb*12 + j]: folded constants of the fp64 MDS layer after round r (see tools/mds_model.py)\n"
    txt += "__constant__ double POSEIDON_MDS_K[30][24] = {\n"
    for row in rows:
        txt += "    {" + ", ".join("%d.0" % v for v in row) + "},\n"
    txt += "};\n"
    from mds_model import partial_constants, BIAS
    lane0, tail = partial_constants(rc)
    # v4: partial rounds in the CRT domain of the circulant (mds_model.permute_v4): the constant is added at read-out
    from mds_model import OFF4_LO, OFF4_HI
    txt += ("// PARTIAL4_Q[r-4][limb] = limb of the pushed-forward lane-0 constant of round r+1 + 2^52 + positivity offset (2^19 p split)\n"
            "__constant__ double POSEIDON_PARTIAL4_Q[22][2] = {\n")
    for r in range(4, 26):
        cl, ch = rc_limbs(lane0[r + 1])
        assert 0 <= cl + OFF4_LO < BIAS and 0 <= ch + OFF4_HI < BIAS
        txt += "    {%d.0, %d.0},\n" % (cl + BIAS + OFF4_LO, ch + BIAS + OFF4_HI)
    txt += ("};\n// PARTIAL4_TAIL[i-1][limb]: what lane i (1..11) receives when it leaves the fp64 domain before round 26: pushed-forward\n"
            "// constant limb + positivity offset + 2^52\n__constant__ double POSEIDON_PARTIAL4_TAIL[11][2] = {\n")
    for i in range(11):
        cl, ch = rc_limbs(tail[i])
        txt += "    {%d.0, %d.0},\n" % (cl + OFF4_LO + BIAS, ch + OFF4_HI + BIAS)
    return txt + "};\n"


def main():
    rc = constants()
    assert hashlib.sha256(b"".join(struct.pack("<Q", c) for c in rc)).hexdigest() == SHA
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    body = ",\n".join("    " + ", ".join("0x%016xULL" % c for c in rc[i:i + 4]) for i in range(0, 360, 4))
    with open(os.path.join(root, "oracle", "poseidon_constants.h"), "w") as f:
        f.write("/* GENERATED by tools/gen_poseidon_constants.py — plonky2 Poseidon ALL_ROUND_CONSTANTS (sha256 %s) */\n" % SHA)
        f.write("#pragma once\n#include <stdint.h>\nstatic const uint64_t GLO_ROUND_CONSTANTS[360] = {\n%s\n};\n" % body)
    with open(os.path.join(root, "plonky2.5_b200", "csrc", "poseidon_constants.cuh"), "w") as f:
        f.write("// GENERATED by tools/gen_poseidon_constants.py — plonky2 Poseidon ALL_ROUND_CONSTANTS (sha256 %s)\n" % SHA)
        f.write("#pragma once\n#include <stdint.h>\n__constant__ uint64_t POSEIDON_RC[360] = {\n%s\n};\n" % body)
        f.write(mds_tables(rc))
    print("ok")


if __name__ == "__main__":
    main()
