"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same seeded inputs — bit-exact."""
import hashlib
import random

import os

import numpy as np
import pytest

import gl_oracle as o
from oracle_c import P, splitmix_columns

pytestmark = pytest.mark.gpu


def _g():
    import plonky25_b200 as g
    return g


def test_poseidon_kat_on_gpu(ctx, golden):
    """KATs from /root/reference/src/common/poseidon2/poseidon2_goldilocks.rs:190-211 through gl_poseidon_permute."""
    vec = golden["poseidon_kat"]["vectors"]
    states = np.array([v["input"] for v in vec], dtype=np.uint64)
    out = ctx.poseidon_permute(states)
    assert out.tolist() == [v["output"] for v in vec]


def test_poseidon_random_and_non_canonical(ctx, oc):
    rng = np.random.default_rng(1)
    states = rng.integers(0, 2**64, size=(4096, 12), dtype=np.uint64)     # includes non-canonical words
    states[0, :] = np.uint64(2**64 - 1)
    states[1, :] = np.uint64(P)
    states[2, :] = np.uint64(P - 1)
    out = ctx.poseidon_permute(states)
    for i in range(states.shape[0]):                 # every one of the 4096 states against the C oracle
        assert out[i].tolist() == oc.poseidon(states[i] % np.uint64(P)).tolist(), i
    assert (out < np.uint64(P)).all()


@pytest.mark.parametrize("shape", [(1, 5, 0), (2, 1, 1), (8, 3, 0), (8, 4, 3), (16, 5, 2), (64, 7, 0), (64, 8, 3), (32, 9, 5),
                                   (256, 135, 4), (128, 32, 4), (1024, 20, 4), (512, 16, 0), (16, 135, 4), (4096, 17, 6)])
def test_merkle_new_matches_oracle(ctx, oc, shape):
    g = _g()
    n, ll, h = shape
    leaves = splitmix_columns(n * 7 + ll, n, ll, canonical=(ll % 2 == 0))   # odd widths get non-canonical words
    tree = g.MerkleTree.new(leaves, h, ctx=ctx)
    dig, cap = oc.merkle_new(leaves, h)
    assert np.array_equal(tree.cap.hashes, cap)
    assert np.array_equal(tree.digests, dig)
    assert np.array_equal(tree.leaves, np.where(leaves >= np.uint64(P), leaves - np.uint64(P), leaves))
    rng = random.Random(n)
    for idx in {0, n - 1, rng.randrange(n), rng.randrange(n)}:
        sib = tree.prove(idx)
        assert oc.verify_path(tree.get(idx), idx, sib, cap)
        assert sib.shape[0] == n.bit_length() - 1 - h


def test_merkle_errors(ctx):
    g = _g()
    with pytest.raises(ValueError, match="cap_height"):
        g.MerkleTree.new(np.zeros((4, 5), dtype=np.uint64), 3, ctx=ctx)
    with pytest.raises(ValueError, match="power of two"):
        g.MerkleTree.new(np.zeros((6, 5), dtype=np.uint64), 1, ctx=ctx)


COMMIT_SHAPES = [(0, 6, 2, 1), (1, 4, 1, 1), (2, 2, 1, 0), (3, 5, 2, 1), (4, 9, 3, 4), (5, 3, 1, 2), (3, 135, 3, 2), (3, 2, 3, 6),
                 (4, 4, 0, 0), (6, 8, 1, 0), (7, 16, 3, 4), (8, 135, 3, 4), (9, 7, 2, 3), (10, 20, 3, 4), (11, 9, 1, 4),
                 (12, 17, 3, 4), (13, 3, 2, 4), (12, 135, 1, 4)]


@pytest.mark.parametrize("shape", COMMIT_SHAPES)
def test_commit_from_values_matches_oracle(ctx, oc, shape):
    g = _g()
    log_n, n_cols, r, h = shape
    cols = splitmix_columns(1000 + log_n * 13 + n_cols, n_cols, 1 << log_n, canonical=(log_n % 2 == 0))
    ref = oc.commit(cols, r, h)
    pb = g.PolynomialBatch.from_values(list(cols), r, False, h, ctx=ctx, copy_back=True)
    assert np.array_equal(pb.polynomials, ref["coeffs"])
    assert np.array_equal(pb.merkle_tree.leaves, ref["leaves"])
    assert np.array_equal(pb.merkle_tree.digests, ref["digests"])
    assert np.array_equal(pb.merkle_tree.cap.hashes, ref["cap"])
    # device-resident variant returns the same through the read-back API
    pb2 = g.PolynomialBatch.from_values(list(cols), r, False, h, ctx=ctx)
    assert np.array_equal(pb2.merkle_tree.cap.hashes, ref["cap"])
    assert np.array_equal(pb2.polynomials, ref["coeffs"])
    R = (1 << log_n) << r
    rng = random.Random(log_n)
    for idx in {0, R - 1, rng.randrange(R)}:
        assert np.array_equal(pb2.merkle_tree.get(idx), ref["leaves"][idx])
        assert oc.verify_path(ref["leaves"][idx], idx, pb2.merkle_tree.prove(idx), ref["cap"])
    step = 1 << r
    for index in {0, (1 << log_n) - 1}:
        assert np.array_equal(pb2.get_lde_values(index, step), ref["leaves"][o.reverse_bits(index * step, log_n + r)])
    # from_coeffs on the coefficients reproduces the same tree
    pb3 = g.PolynomialBatch.from_coeffs(list(ref["coeffs"]), r, False, h, ctx=ctx)
    assert np.array_equal(pb3.merkle_tree.cap.hashes, ref["cap"])
    assert np.array_equal(pb3.merkle_tree.digests, ref["digests"])


def test_commit_golden_fixture_on_gpu(ctx, golden):
    g = _g()
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a, dtype="<u8").tobytes()).hexdigest()
    for c in golden["commit_small"]["cases"]:
        cols = splitmix_columns(c["seed"], c["n_cols"], 1 << c["log_n"])
        f = g.PolynomialBatch.from_coeffs if c["is_coeffs"] else g.PolynomialBatch.from_values
        pb = f(list(cols), c["rate_bits"], False, c["cap_height"], ctx=ctx, copy_back=True)
        assert pb.merkle_tree.cap.flatten().tolist() == c["cap"]
        assert sha(pb.polynomials) == c["sha256_coeffs"]
        assert sha(pb.merkle_tree.leaves) == c["sha256_leaves"]
        assert sha(pb.merkle_tree.digests) == c["sha256_digests"]


def test_tiny_commit_anchor(ctx, golden):
    g = _g()
    t = golden["derived_anchors"]["tiny_commit"]
    pb = g.PolynomialBatch.from_values([np.array(v, dtype=np.uint64) for v in t["values"]], t["rate_bits"], False, t["cap_height"],
                                       ctx=ctx, copy_back=True)
    assert pb.polynomials.tolist() == t["coeffs"]
    assert pb.merkle_tree.leaves.tolist() == t["leaves"]
    assert pb.merkle_tree.digests.tolist() == t["digests"]
    assert pb.merkle_tree.cap.hashes.tolist() == t["cap"]


def test_commit_errors(ctx):
    g = _g()
    cols = [np.zeros(8, dtype=np.uint64)] * 3
    with pytest.raises(ValueError, match="cap_height"):
        g.PolynomialBatch.from_values(cols, 1, False, 5, ctx=ctx)
    # the context stays usable after an error
    pb = g.PolynomialBatch.from_values(cols, 1, False, 4, ctx=ctx)
    assert pb.merkle_tree.cap.hashes.shape == (16, 4)
    assert (pb.merkle_tree.leaves == 0).all()


def test_commit_linearity_property(ctx):
    """LDE is linear: leaves(a + b) == leaves(a) + leaves(b) (mod p) — size-independent check."""
    g = _g()
    log_n, n_cols, r = 12, 6, 3
    a = splitmix_columns(1, n_cols, 1 << log_n)
    b = splitmix_columns(2, n_cols, 1 << log_n)
    s = ((a.astype(object) + b.astype(object)) % P).astype(np.uint64)
    la = g.PolynomialBatch.from_values(list(a), r, False, 4, ctx=ctx).merkle_tree.leaves.astype(object)
    lb = g.PolynomialBatch.from_values(list(b), r, False, 4, ctx=ctx).merkle_tree.leaves.astype(object)
    ls = g.PolynomialBatch.from_values(list(s), r, False, 4, ctx=ctx).merkle_tree.leaves.astype(object)
    assert ((la + lb) % P == ls).all()


def test_wrapper_shape_cfg2_properties(ctx, oc):
    """BASELINE config 2 (2^16 x 135, rate_bits 3, cap_height 4) at full size through size-independent properties:
    sampled leaves are Horner evaluations of the returned coefficients, coefficients interpolate the inputs, sampled
    Merkle paths verify with the verifier's rule, and the cap equals the oracle's cap recomputed from the GPU leaves'
    top levels."""
    g = _g()
    log_n, n_cols, r, h = 16, 135, 3, 4
    cols = splitmix_columns(42, n_cols, 1 << log_n)
    pb = g.PolynomialBatch.from_values(list(cols), r, False, h, ctx=ctx)
    tree = pb.merkle_tree
    coeffs = pb.polynomials
    bits = log_n + r
    w = o.primitive_root_of_unity(bits)
    rng = random.Random(9)
    R = 1 << bits
    for idx in [0, 1, R - 1] + [rng.randrange(R) for _ in range(5)]:
        row = tree.get(idx)
        x = 7 * pow(w, o.reverse_bits(idx, bits), P) % P
        for j in (0, 1, 67, 134):
            assert int(row[j]) == o.eval_poly(coeffs[j].tolist(), x)
        assert oc.verify_path(row, idx, tree.prove(idx), tree.cap.hashes)
    wn = o.primitive_root_of_unity(log_n)
    for k in (0, 1, 12345):
        for j in (0, 134):
            assert o.eval_poly(coeffs[j].tolist(), pow(wn, k, P)) == int(cols[j][k])
    # whole-tree check: digests recomputed by the oracle from the GPU's leaves
    dig, cap = oc.merkle_new(tree.leaves, h)
    assert np.array_equal(cap, tree.cap.hashes)
    assert np.array_equal(dig, tree.digests)


def _fri_inputs(oc, log_len, rate_bits, seed):
    n = 1 << log_len
    low = n >> rate_bits
    co = splitmix_columns(seed, n, 2)
    co[low:] = 0
    va = np.stack([oc.coset_fft(co[:, 0], 7), oc.coset_fft(co[:, 1], 7)], axis=1)
    return co, va


@pytest.mark.parametrize("cfg", [(8, 3, 1, [4]), (9, 1, 2, [2, 3]), (10, 3, 2, [4, 4]), (12, 3, 4, [4, 4]), (15, 3, 4, [4, 4]),
                                 (19, 3, 4, [4, 4, 4])])
def test_fri_committed_trees_matches_oracle(ctx, oc, cfg):
    """fri_committed_trees with the oracle's challenger as the caller-side transcript (wrapper shape = last case)."""
    g = _g()
    log_len, rate_bits, cap_h, arities = cfg
    co, va = _fri_inputs(oc, log_len, rate_bits, log_len)
    ch_ref, ch_gpu = oc.new_challenger(), oc.new_challenger()
    for ch in (ch_ref, ch_gpu):
        ch.observe_elements([1, 2, 3])
    ref = oc.fri_committed_trees(co, va, arities, rate_bits, cap_h, ch_ref)
    trees, final = g.fri_committed_trees(co, va, ch_gpu, g.FriParams(rate_bits, cap_h, arities), ctx=ctx)
    assert len(trees) == len(arities)
    for t, lv, dg, cap in zip(trees, ref["leaves"], ref["digests"], ref["caps"]):
        assert np.array_equal(t.cap.hashes, cap)
        assert np.array_equal(t.leaves, lv)
        if t.digests.size:
            assert np.array_equal(t.digests, dg)
    assert np.array_equal(final, ref["final_poly"])
    assert ch_ref.get_challenge() == ch_gpu.get_challenge()


def test_gpu_challenger_mirror_matches_oracle(ctx, oc):
    g = _g()
    rng = random.Random(3)
    a, b = g.Challenger(ctx), oc.new_challenger()
    for _ in range(12):
        es = [rng.randrange(P) for _ in range(rng.randrange(1, 13))]
        a.observe_elements(es)
        b.observe_elements(es)
        for _ in range(rng.randrange(0, 3)):
            assert a.get_challenge() == b.get_challenge()


def test_stage_times_and_launch_counts(ctx):
    g = _g()
    cols = splitmix_columns(5, 16, 1 << 10)
    g.PolynomialBatch.from_values(list(cols), 3, False, 4, ctx=ctx)
    ms, launches = ctx.stage_times()
    # host columns arrive in chunks of 8, 16, 32, ... columns (copy/compute overlap): 16 columns = 2 chunks, each with its own
    # iNTT (1 pass at N = 2^10), 8 coset NTTs and one launch of the streaming leaf sponge over its columns
    assert launches["leaf_hash"] == 2 and launches["tree"] == 13 - 4 and launches["lde"] == 16 and launches["intt"] == 2
    assert all(v >= 0 for v in ms.values())


def _edge_words():
    e = [0, 1, 2, 0xFFFFFFFE, 0xFFFFFFFF, 1 << 32, (1 << 32) + 1, P - 2, P - 1, P, P + 1, 2**64 - 2, 2**64 - 1, 0xFFFFFFFF << 32,
         (0xFFFFFFFF << 32) | 1, 1 << 63, (1 << 63) - 1, 0xFFFFFFFE00000001, 0xFFFFFFFE00000002, 0x00000001FFFFFFFF, 0xFFFFFF, 1 << 40]
    return e


def test_field_primitives_on_adversarial_words(ctx):
    """gl::mul / add_any / sub_any / shift multiplies / S-box exactly as the kernels use them, on every pair of edge words
    (carry and double-carry corners that random data reaches with probability 2^-32) and on random words."""
    from plonky25_b200 import _lib  # noqa: F401
    rng = np.random.default_rng(11)
    e = _edge_words()
    a = np.array([x for x in e for _ in e] + rng.integers(0, 2**64, 20000, dtype=np.uint64).tolist(), dtype=np.uint64)
    b = np.array([y for _ in e for y in e] + rng.integers(0, 2**64, 20000, dtype=np.uint64).tolist(), dtype=np.uint64)
    ai, bi = [int(x) for x in a], [int(y) for y in b]
    FOP = dict(mul=0, add_any=1, sub_any=2, mul_2_24=3, mul_2_48=4, mul_2_72=5, sbox7=6, add_any_c=7)
    want = {
        "mul": [x * y % P for x, y in zip(ai, bi)],
        "add_any": [(x + y) % P for x, y in zip(ai, bi)],
        "sub_any": [(x - y) % P for x, y in zip(ai, bi)],
        "mul_2_24": [(x << 24) % P for x in ai],
        "mul_2_48": [(x << 48) % P for x in ai],
        "mul_2_72": [(x << 72) % P for x in ai],
        "sbox7": [pow(x, 7, P) for x in ai],
        "add_any_c": [(x + y) % P for x, y in zip(ai, bi)],
    }
    for name, op in FOP.items():
        got = ctx.field_op(op, a, b).tolist()
        bad = [i for i, (g_, w_) in enumerate(zip(got, want[name])) if g_ != w_]
        assert not bad, f"{name}: {len(bad)} mismatches, first a={ai[bad[0]]:#x} b={bi[bad[0]]:#x} got={got[bad[0]]:#x} want={want[name][bad[0]]:#x}"


@pytest.mark.parametrize("bits", [0, 5, 12, 16, 20])
def test_fri_proof_of_work_matches_oracle(ctx, oc, bits):
    """gl_fri_pow (wrapper proof uses 16 bits) against the C restatement: same smallest witness, same transcript after."""
    g = _g()
    rng = random.Random(bits)
    a, b = g.Challenger(ctx), oc.new_challenger()
    es = [rng.randrange(P) for _ in range(bits % 7)]
    a.observe_elements(es)
    b.observe_elements(es)
    wa = g.fri_proof_of_work(a, bits, ctx=ctx)
    wb = b.fri_proof_of_work(bits)
    assert wa == wb
    assert a.get_challenge() == b.get_challenge()


# ---- prove_openings, front half ---------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [(3, [3, 2], 1, 1, []), (6, [9, 5, 2, 2], 3, 2, [2]), (10, [135, 20], 3, 4, [4]), (12, [86, 135, 20, 16], 3, 4, [4, 4]),
                                 (0, [2], 2, 0, [])])
def test_prove_openings_front_half_matches_oracle(ctx, oc, cfg):
    """commits -> prove_openings(instance) on the device: per-batch quotients, final_poly, its LDE, the FRI commit phase and the
    proof-of-work witness equal the oracle's (same transcript), for plonky2-shaped instances (all polynomials at zeta, a few at g*zeta)."""
    g = _g()
    log_n, widths, r, cap_h, arities = cfg
    n = 1 << log_n
    rng = random.Random(log_n * 31 + len(widths))
    cols = [splitmix_columns(900 + 17 * k + log_n, w, n) for k, w in enumerate(widths)]
    batches = [g.PolynomialBatch.from_values(list(c), r, False, min(cap_h, log_n + r), ctx=ctx) for c in cols]
    coeffs = [b.polynomials for b in batches]
    zeta = (rng.randrange(P), rng.randrange(P))
    gz = (rng.randrange(P), rng.randrange(P))
    all_polys = [(k, i) for k, w in enumerate(widths) for i in range(w)]
    instance = [g.FriBatchInfo(zeta, all_polys), g.FriBatchInfo(gz, [(len(widths) - 1, i) for i in range(min(2, widths[-1]))])]
    ch_ref, ch_gpu = oc.new_challenger(), g.Challenger(ctx)     # the product's own challenger mirror (needed by the PoW grind)
    for ch in (ch_ref, ch_gpu):
        ch.observe_elements([5, 6, 7, 8, 9])
    # reference transcript
    alpha = ch_ref.get_extension_challenge()
    final_ref, quot_ref = oc.openings_final_poly([(b.point, b.polynomials) for b in instance], coeffs, alpha)
    lde = np.zeros((n << r, 2), dtype=np.uint64)
    lde[:n] = final_ref
    vals = np.stack([oc.coset_fft(lde[:, 0], 7), oc.coset_fft(lde[:, 1], 7)], axis=1)
    ref = oc.fri_committed_trees(lde, vals, arities, r, min(cap_h, log_n + r), ch_ref)
    w_ref = ch_ref.fri_proof_of_work(8)
    # product
    dbg = {}
    head = g.prove_openings(instance, batches, ch_gpu, g.FriParams(r, min(cap_h, log_n + r), arities), proof_of_work_bits=8, ctx=ctx, debug=dbg)
    for q, qr in zip(dbg["quotients"], quot_ref):
        assert np.array_equal(q, qr)
    assert np.array_equal(dbg["final_poly"], final_ref)
    assert np.array_equal(dbg["lde_final_poly"], lde)
    bits = log_n + r
    rev = np.array([int(format(i, "0%db" % bits)[::-1], 2) if bits else 0 for i in range(n << r)])
    assert np.array_equal(dbg["lde_final_values_bitrev"], vals[rev])
    assert len(head.trees) == len(arities)
    for t, cap in zip(head.trees, ref["caps"]):
        assert np.array_equal(t.cap.hashes, cap)
    assert np.array_equal(head.final_poly, ref["final_poly"])
    assert head.pow_witness == w_ref
    assert ch_ref.get_challenge() == ch_gpu.get_challenge()


def test_prove_openings_errors(ctx):
    g = _g()
    a = g.PolynomialBatch.from_values(list(splitmix_columns(1, 3, 8)), 1, False, 1, ctx=ctx)
    b = g.PolynomialBatch.from_values(list(splitmix_columns(2, 3, 16)), 1, False, 1, ctx=ctx)
    ch = g.Challenger(ctx)
    with pytest.raises(ValueError, match="Polynomial degrees inconsistent"):
        g.prove_openings([g.FriBatchInfo((1, 2), [(0, 0), (1, 0)])], [a, b], ch, g.FriParams(1, 1, []), ctx=ctx)
    with pytest.raises(ValueError, match="out of range"):
        g.prove_openings([g.FriBatchInfo((1, 2), [(0, 3)])], [a], ch, g.FriParams(1, 1, []), ctx=ctx)


@pytest.mark.parametrize("cfg", [(6, [9, 5], 2, 2, [2, 2], 5), (10, [135, 20], 3, 4, [4, 4], 28), (4, [3], 1, 5, [], 3)])
def test_fri_query_rounds_open_and_verify(ctx, oc, cfg):
    """prove_openings -> fri_prover_query_rounds: every opened row equals the single-call readers (already checked against the
    oracle), every Merkle path verifies with the VERIFIER's rule against the committed cap, the commit-phase evaluations are the
    whole layer leaves (`arity` elements; the structure oracle/gl_oracle.py · fri_prover_query_rounds restates)."""
    g = _g()
    log_n, widths, r, cap_h, arities, n_rounds = cfg
    n = 1 << log_n
    cols = [splitmix_columns(300 + 13 * k + log_n, w, n) for k, w in enumerate(widths)]
    cap_h = min(cap_h, log_n + r)
    batches = [g.PolynomialBatch.from_values(list(c), r, False, cap_h, ctx=ctx) for c in cols]
    instance = [g.FriBatchInfo((11, 22), [(k, i) for k, w in enumerate(widths) for i in range(w)])]
    ch = g.Challenger(ctx)
    ch.observe_elements([1, 2, 3])
    head = g.prove_openings(instance, batches, ch, g.FriParams(r, cap_h, arities), proof_of_work_bits=4, ctx=ctx)
    ch2 = g.Challenger(ctx)                      # the query indices come from the transcript: replay them for the checker
    ch2.sponge_state, ch2.input_buffer, ch2.output_buffer = ch.sponge_state.copy(), list(ch.input_buffer), list(ch.output_buffer)
    rounds = g.fri_prover_query_rounds([b.merkle_tree for b in batches], head.trees, ch, n_rounds, g.FriParams(r, cap_h, arities))
    assert len(rounds) == n_rounds
    for rnd in rounds:
        x = int(ch2.get_challenge()) % (n << r)
        assert rnd["x_index"] == x
        for b, (row, sib) in zip(batches, rnd["initial_trees_proof"]):
            t = b.merkle_tree
            assert np.array_equal(row, t.get(x)) and np.array_equal(sib, t.prove(x))
            assert oc.verify_path(row, x, sib, t.cap.hashes)
        for arity_bits, t, step in zip(arities, head.trees, rnd["steps"]):
            leaf = t.get(x >> arity_bits)
            # as the verifier does (fri/verifier.rs · fri_verifier_query_round): evals has `arity` elements, the queried one is
            # evals[x & (arity-1)], and the Merkle proof is checked over flatten(evals)
            ev = np.asarray(step["evals"])
            assert ev.shape == (1 << arity_bits, 2)
            assert np.array_equal(ev[x & ((1 << arity_bits) - 1)], leaf.reshape(-1, 2)[x & ((1 << arity_bits) - 1)])
            assert np.array_equal(ev.reshape(-1), leaf)
            assert np.array_equal(step["merkle_proof"], t.prove(x >> arity_bits))
            assert oc.verify_path(ev.reshape(-1), x >> arity_bits, step["merkle_proof"], t.cap.hashes)
            x >>= arity_bits


def test_tree_open_batch_edges(ctx):
    g = _g()
    lv = splitmix_columns(9, 8, 5)                       # 8 leaves of 5 words
    t = g.MerkleTree.new(lv, 3, ctx=ctx)                 # cap_height = log2(n_leaves): no digests, empty proofs
    rows, sib = t.open_batch([7, 0, 7])
    assert np.array_equal(rows, lv[[7, 0, 7]]) and sib.shape == (3, 0, 4)
    rows, sib = t.open_batch([])
    assert rows.shape == (0, 5)
    with pytest.raises(ValueError, match="out of range"):
        t.open_batch([8])


def test_openings_fri_golden_fixture_on_gpu(ctx, golden):
    """commits -> prove_openings -> query indices reproduce tests/golden/openings_fri.json on the device."""
    g = _g()
    f = golden["openings_fri"]
    n = 1 << f["log_n"]
    batches = [g.PolynomialBatch.from_values(list(splitmix_columns(s, w, n)), f["rate_bits"], False, f["cap_height"], ctx=ctx)
               for s, w in zip(f["col_seeds"], f["widths"])]
    ch = g.Challenger(ctx)
    ch.observe_elements(f["transcript_prefix"])
    instance = [g.FriBatchInfo(b["point"], [tuple(x) for x in b["polynomials"]]) for b in f["batches"]]
    dbg = {}
    params = g.FriParams(f["rate_bits"], f["cap_height"], f["arity_bits"])
    head = g.prove_openings(instance, batches, ch, params, proof_of_work_bits=f["pow_bits"], ctx=ctx, debug=dbg)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a, dtype="<u8").tobytes()).hexdigest()
    assert sha(dbg["final_poly"]) == f["sha256_final_poly"]
    assert [sha(q) for q in dbg["quotients"]] == f["sha256_quotients"]
    assert [[int(x) for x in t.cap.flatten()] for t in head.trees] == f["commit_phase_caps"]
    assert [[int(a), int(b)] for a, b in head.final_poly] == f["fri_final_poly"]
    assert head.pow_witness == f["pow_witness"]
    rounds = g.fri_prover_query_rounds([b.merkle_tree for b in batches], head.trees, ch, len(f["query_indices"]), params)
    assert [r["x_index"] for r in rounds] == f["query_indices"]


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [(2, 8, 20, 1, 1), (2, 10, 135, 3, 4), (4, 9, 33, 2, 2), (4, 10, 135, 3, 4), (8, 7, 9, 3, 3), (2, 3, 2, 0, 1)])
def test_commit_multi_single_process(oc, cfg):
    """gl_commit_multi with n contexts (here all on device 0: the plan, the shipments and the per-range hashing are the same as
    with one device per context): cap, every shard's leaves and digests, rows and paths equal the single commit of the oracle."""
    g = _g()
    world, log_n, n_cols, r, cap_h = cfg
    # one device per context when the box has them (peer access over NVLink), else the contexts share devices
    n_dev = int(os.environ.get("GL_TEST_COMMIT_MULTI_DEVICES", "0")) or max(1, min(world, g._lib.load().gl_device_count()))
    ctxs = [g.Context(k % n_dev) for k in range(world)]
    try:
        cols = splitmix_columns(700 + world + log_n, n_cols, 1 << log_n)
        ref = oc.commit(cols, r, cap_h)
        for rep in range(2):                       # buffers and peer mappings are reused by the second call
            cap, trees = g.commit_multi(ctxs, list(cols), r, cap_h)
            assert np.array_equal(cap.hashes, ref["cap"])
            rows = (1 << (log_n + r)) // world
            dig_per = 2 * (rows - (1 << (cap_h - (world.bit_length() - 1))))
            for q, t in enumerate(trees):
                assert (t.n_leaves, t.leaf_len, t.cap_height) == (rows, n_cols, cap_h - (world.bit_length() - 1))
                assert np.array_equal(t.leaves, ref["leaves"][q * rows:(q + 1) * rows])
                assert np.array_equal(t.digests, ref["digests"][q * dig_per:(q + 1) * dig_per])
                for i in {0, rows - 1, rows // 3}:
                    assert oc.verify_path(t.get(i), q * rows + i, t.prove(i), ref["cap"])
            for t in trees:
                t.free()
        with pytest.raises(ValueError, match="cap_height too small"):
            g.commit_multi(ctxs, list(cols), r, 0)
        with pytest.raises(ValueError, match="twice"):
            g.commit_multi([ctxs[0], ctxs[0]], list(cols), r, cap_h)
    finally:
        for c in ctxs:
            c.close()
