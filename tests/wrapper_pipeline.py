"""The wrapper-circuit-shaped prove pipeline (BASELINE.json configs[0]/[3]) — the part of `data.prove(pw)`
(/root/reference/src/p3/mod.rs:258-262) that lies on the commitment hot path, as one timed unit:

    commits   constants+sigmas 86, wires 135, Z/partial products 20 (from_values), quotient chunks 16 (from_coeffs)   x 2^16, r=3, h=4
    openings  alpha <- transcript; batch zeta: all 257 polynomials, batch g*zeta: the 2 Z polynomials -> final_poly -> LDE (2^19)
    FRI       commit phase with arities [4, 4, 4] (trees of 2^15 / 2^11 / 2^7 leaves), final polynomial 2^4 coefficients
    PoW       16 bits (smallest witness), then 28 query rounds over the 4 initial trees + 3 commit-phase trees

The real wrapper proof cannot be produced here (no Rust toolchain; the circuit's witness generation, gate evaluation and verifier
are out of scope), so the COLUMN CONTENTS are synthetic (SplitMix64) and the transcript prefix is fixed; shapes, call order and every
byte that goes into the FRI proof are the reference's.  Two runners produce the same `proof` dictionary:

    run_product(g, ctx, ...)   the product: C ABI through the Python mirror of the reference interface, HOST columns in
    run_oracle(oc, ...)        the CPU oracle on the same transcript (test infrastructure: tests / bench.py's baseline legs only)

and `digest(proof)` hashes every proof field (caps, final polynomial, witness, query indices, every opened row and sibling).
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle_c import splitmix_columns  # noqa: E402

P = 0xFFFF_FFFF_0000_0001
WIDTHS = (86, 135, 20, 16)          # constants+sigmas, wires, Z + partial products, quotient chunks
IS_COEFFS = (False, False, False, True)
ARITIES = (4, 4, 4)
POW_BITS = 16
N_QUERIES = 28
PREFIX = [0x706C6F6E6B79, 2, 5]     # stands in for the circuit digest / public-inputs hash that open upstream's transcript


def make_columns(log_n, widths=WIDTHS, seed=7):
    return [splitmix_columns(seed + 101 * k, w, 1 << log_n) for k, w in enumerate(widths)]


def instance_for(widths):
    """FriInstanceInfo of a plonky2 proof: every polynomial opened at zeta, the Z polynomials (first two of oracle 2) also at g*zeta"""
    all_polys = [(k, i) for k, w in enumerate(widths) for i in range(w)]
    zs = [(2, i) for i in range(min(2, widths[2]))] if len(widths) > 2 else []
    return all_polys, zs


def _g_times(zeta, log_n):
    """g * zeta with g the generator of the size-2^log_n subgroup (extension element times base element)"""
    g = pow(1753635133440165772, 1 << (32 - log_n), P)
    return (zeta[0] * g % P, zeta[1] * g % P)


def digest(proof) -> str:
    h = hashlib.sha256()

    def put(a):
        h.update(np.ascontiguousarray(np.asarray(a, dtype=np.uint64)).tobytes())
    for cap in proof["commit_caps"]:
        put(cap)
    for cap in proof["commit_phase_caps"]:
        put(cap)
    put(proof["final_poly"])
    put([proof["pow_witness"]])
    put(proof["query_indices"])
    for rnd in proof["query_rounds"]:
        for row, sib in rnd["initial"]:
            put(row); put(sib)
        for ev, sib in rnd["steps"]:
            put(ev); put(sib)
    return h.hexdigest()


# ------------------------------------------------------------------------------------------------------------------ product
def run_product(g, ctx, cols, log_n, r=3, h=4, arities=ARITIES, pow_bits=POW_BITS, n_queries=N_QUERIES, widths=WIDTHS,
                is_coeffs=IS_COEFFS, timings=None):
    t = {}
    t0 = time.perf_counter()
    batches = []
    for k, c in enumerate(cols):
        ta = time.perf_counter()
        fn = g.PolynomialBatch.from_coeffs if is_coeffs[k] else g.PolynomialBatch.from_values
        batches.append(fn(list(c), r, False, h, ctx=ctx))
        t[f"commit_{widths[k]}"] = (time.perf_counter() - ta) * 1e3
    ta = time.perf_counter()
    ch = g.Challenger(ctx)
    ch.observe_elements(PREFIX)
    for b in batches:
        ch.observe_cap(b.merkle_tree.cap)
    zeta = ch.get_extension_challenge()
    all_polys, zs = instance_for(widths)
    instance = [g.FriBatchInfo(zeta, all_polys)] + ([g.FriBatchInfo(_g_times(zeta, log_n), zs)] if zs else [])
    t["transcript_head"] = (time.perf_counter() - ta) * 1e3
    ta = time.perf_counter()
    params = g.FriParams(r, h, list(arities))
    head = g.prove_openings(instance, batches, ch, params, proof_of_work_bits=pow_bits, ctx=ctx)
    t["prove_openings_fri_pow"] = (time.perf_counter() - ta) * 1e3
    ta = time.perf_counter()
    rounds = g.fri_prover_query_rounds([b.merkle_tree for b in batches], head.trees, ch, n_queries, params)
    t["query_rounds"] = (time.perf_counter() - ta) * 1e3
    t["total"] = (time.perf_counter() - t0) * 1e3
    proof = {"commit_caps": [b.merkle_tree.cap.hashes.copy() for b in batches],
             "commit_phase_caps": [tr.cap.hashes.copy() for tr in head.trees],
             "final_poly": np.asarray(head.final_poly).copy(), "pow_witness": int(head.pow_witness),
             "query_indices": [int(rn["x_index"]) for rn in rounds],
             "query_rounds": [{"initial": [(row, sib) for row, sib in rn["initial_trees_proof"]],
                               "steps": [(st["evals"], st["merkle_proof"]) for st in rn["steps"]]} for rn in rounds]}
    for b in batches:
        b.merkle_tree.free()
    for tr in head.trees:
        tr.free()
    if timings is not None:
        timings.update(t)
    return proof


# ------------------------------------------------------------------------------------------------------------------- oracle
def _prove_from_layout(digests, n_leaves, cap_height, index):
    """MerkleTree::prove on the flat digests vector (plonky2 hash/merkle_tree.rs layout): siblings bottom-up"""
    log_sub = (n_leaves.bit_length() - 1) - cap_height
    L = 1 << log_sub
    sub, j = index >> log_sub, index & (L - 1)
    base = sub * 2 * (L - 1)
    out = np.zeros((log_sub, 4), dtype=np.uint64)
    for layer in range(log_sub):
        s = j ^ 1
        out[layer] = digests[base + 2 * (((s >> 1) << (layer + 1)) + (1 << layer) - 1) + (s & 1)]
        j >>= 1
    return out


def run_oracle(oc, cols, log_n, r=3, h=4, arities=ARITIES, pow_bits=POW_BITS, n_queries=N_QUERIES, widths=WIDTHS,
               is_coeffs=IS_COEFFS, timings=None):
    t = {}
    t0 = time.perf_counter()
    n = 1 << log_n
    commits = []
    for k, c in enumerate(cols):
        ta = time.perf_counter()
        commits.append(oc.commit(c, r, h, is_coeffs=is_coeffs[k]))
        t[f"commit_{widths[k]}"] = (time.perf_counter() - ta) * 1e3
    ta = time.perf_counter()
    ch = oc.new_challenger()
    ch.observe_elements(PREFIX)
    for cm in commits:
        ch.observe_cap(cm["cap"])
    zeta = ch.get_extension_challenge()
    all_polys, zs = instance_for(widths)
    inst = [(zeta, all_polys)] + ([(_g_times(zeta, log_n), zs)] if zs else [])
    alpha = ch.get_extension_challenge()
    final, _ = oc.openings_final_poly(inst, [cm["coeffs"] for cm in commits], alpha)
    lde = np.zeros((n << r, 2), dtype=np.uint64)
    lde[:n] = final
    vals = np.stack([oc.coset_fft(lde[:, 0], 7), oc.coset_fft(lde[:, 1], 7)], axis=1)
    fri = oc.fri_committed_trees(lde, vals, list(arities), r, h, ch)
    w = ch.fri_proof_of_work(pow_bits)
    t["prove_openings_fri_pow"] = (time.perf_counter() - ta) * 1e3
    ta = time.perf_counter()
    R = n << r
    xs, rounds = [], []
    for _ in range(n_queries):
        x = ch.get_challenge() % R
        xs.append(x)
        rnd = {"initial": [(cm["leaves"][x].copy(), _prove_from_layout(cm["digests"], R, h, x)) for cm in commits], "steps": []}
        cur, n_leaves = x, R
        for a, leaves, digs in zip(arities, fri["leaves"], fri["digests"]):
            n_leaves >>= a
            idx = cur >> a
            rnd["steps"].append((leaves[idx].reshape(-1, 2).copy(), _prove_from_layout(digs, n_leaves, h, idx)))
            cur = idx
        rounds.append(rnd)
    t["query_rounds"] = (time.perf_counter() - ta) * 1e3
    t["total"] = (time.perf_counter() - t0) * 1e3
    if timings is not None:
        timings.update(t)
    return {"commit_caps": [cm["cap"] for cm in commits], "commit_phase_caps": fri["caps"], "final_poly": fri["final_poly"],
            "pow_witness": int(w), "query_indices": xs, "query_rounds": rounds}
