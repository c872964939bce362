#!/usr/bin/env python3
"""Where the leaf hash's instructions go: static SASS of merkle::leaf_hash_kernel attributed, through the -lineinfo inline chains, to the
component of poseidon.cuh that produced it and weighted by how often that code runs per permutation (full-round body x8, partial-round
body x22, its re-normalisation x11, prologue/epilogue x1).  Development aid behind DESIGN.md §3.1; no GPU needed:
    python tools/sass_breakdown.py plonky2.5_b200/libgl_commit.so [kernel-substring]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

lib = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else "leaf_hash_kernel"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "plonky2.5_b200", "csrc", "poseidon.cuh")

# line ranges of poseidon.cuh's functions, found by name so that the tool follows edits of the file
src = open(SRC).read().splitlines()


def span(name, which=-1):
    """[first, last] line (1-based) of the LAST definition of `name` (the v5 forms come after the #else), up to its closing brace at column 0"""
    starts = [i for i, l in enumerate(src) if re.search(r"\b%s\s*\(" % re.escape(name), l) and "__device__" in l]
    i = starts[which]
    j = i
    while not src[j].startswith("}"):
        j += 1
    return i + 1, j + 1


SPANS = {k: span(k) for k in ("sbox_sqr", "sbox_mul_words", "sbox7_limbs", "mds_limb", "recombine", "full_round", "freq5_forward", "freq5_round",
                              "third_half", "freq5_lane0", "normalize_limbs4", "normalize_limbs12", "partial_rounds", "permute", "biased", "limb_to_double")}
pr0, pr1 = SPANS["partial_rounds"]
loop0 = next(i + 1 for i in range(pr0, pr1) if "for (int r = 0; r < N_PARTIAL" in src[i])
loop1 = next(i + 1 for i in range(loop0, pr1) if src[i].startswith("    }"))          # closing brace of the round loop
norm0 = next(i + 1 for i in range(loop0, loop1) if "if (r & 1)" in src[i])
norm1 = next(i + 1 for i in range(norm0, loop1) if src[i].startswith("        }"))


def inside(line, name):
    a, b = SPANS[name]
    return a <= line <= b


def classify(frames):
    """frames: [(file, line)] innermost first -> (region, component, weight per permutation)"""
    pos = [l for f, l in frames if f.endswith("poseidon.cuh")]
    if not pos:
        return "outside permute", "sponge: loads, overwrite, digest store", 0.0
    region, weight = "permute head (round-0 constants)", 1.0
    if any(inside(l, "full_round") for l in pos):
        region, weight = "full rounds (x8)", 8.0
    elif any(inside(l, "partial_rounds") for l in pos):
        body = [l for l in pos if inside(l, "partial_rounds")]
        l = body[-1]                                   # outermost frame inside partial_rounds = the statement of its body
        if norm0 <= l <= norm1:
            region, weight = "partial rounds: re-normalisation (x11)", 11.0
        elif loop0 <= l <= loop1:
            region, weight = "partial rounds: round body (x22)", 22.0
        else:
            region, weight = "partial rounds: into / out of the CRT domain (x1)", 1.0
    comp = "glue"
    for name, label in (("sbox7_limbs", "S-box x^7 (+ hand-off to fp64)"), ("sbox_sqr", "S-box x^7 (+ hand-off to fp64)"),
                        ("sbox_mul_words", "S-box x^7 (+ hand-off to fp64)"), ("mds_limb", "MDS layer (fp64 CRT convolution)"),
                        ("recombine", "recombine (fp64 limbs -> 64-bit word)"), ("freq5_round", "CRT-domain round (fp64)"),
                        ("freq5_lane0", "lane-0 read-out"), ("third_half", "lane-0 read-out"), ("normalize_limbs4", "limb re-normalisation"),
                        ("normalize_limbs12", "limb re-normalisation"), ("freq5_forward", "into the CRT domain"), ("limb_to_double", "into the CRT domain")):
        if any(inside(l, name) for l in pos):
            comp = label
            break
    return region, comp, weight


with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, check=True, capture_output=True)
    cubin = next(os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin"))
    sass = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()

PIPE = {"IMAD.WIDE": "imad.wide", "DADD": "fp64", "DFMA": "fp64", "DMUL": "fp64", "IMAD": "imad32 / mov", "IADD3": "alu", "LOP3": "alu", "SHF": "alu",
        "LEA": "alu", "SEL": "alu", "ISETP": "alu", "VIADD": "alu", "PRMT": "alu", "IADD": "alu", "MOV": "imad32 / mov", "CS2R": "imad32 / mov"}
cur, frames, pending = None, [], []
table = collections.defaultdict(lambda: collections.Counter())
for line in sass:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        cur = m.group(1) if pat in m.group(1) else None
        continue
    if cur is None:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        pending.append((m.group(1), int(m.group(2))))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", line)
    if not m:
        continue
    if pending:
        frames, pending = pending, []
    op = m.group(1)
    base = op.split(".")[0]
    key = "IMAD.WIDE" if op.startswith("IMAD.WIDE") else base
    pipe = PIPE.get(key, "other (ldc, bra, ...)")
    region, comp, w = classify(frames)
    table[(region, comp)][pipe] += w if w else 0
    table[(region, comp)]["_static"] += 1

pipes = ["imad.wide", "alu", "fp64", "imad32 / mov", "other (ldc, bra, ...)"]
print(f"{'region':52s} {'component':40s} " + " ".join(f"{p[:12]:>12s}" for p in pipes) + f" {'total':>8s} {'static':>7s}")
tot = collections.Counter()
for (region, comp), c in sorted(table.items(), key=lambda kv: -sum(v for k, v in kv[1].items() if k != "_static")):
    row = [c.get(p, 0) for p in pipes]
    print(f"{region:52s} {comp:40s} " + " ".join(f"{v:12.0f}" for v in row) + f" {sum(row):8.0f} {c['_static']:7d}")
    for p, v in zip(pipes, row):
        tot[p] += v
print(f"{'per permutation (static SASS x trip counts)':93s} " + " ".join(f"{tot[p]:12.0f}" for p in pipes) + f" {sum(tot.values()):8.0f}")
