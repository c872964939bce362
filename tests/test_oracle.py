"""CPU tests of the oracle itself: golden vectors the reference holds, Python-vs-C cross-restatement, and the
oracle-independent algebraic invariants of SURVEY.md A.9.  No GPU needed."""
import hashlib
import random
import struct

import numpy as np
import pytest

import gl_oracle as o
from oracle_c import P, splitmix_columns


def test_round_constants_pinned():
    rc = o.ALL_ROUND_CONSTANTS
    assert len(rc) == 360 and all(c < o.P for c in rc)
    assert rc[0] == 0xB585F766F2144405 and rc[359] == 0xBC8DFB627FE558FC
    assert hashlib.sha256(b"".join(struct.pack("<Q", c) for c in rc)).hexdigest() == o.ROUND_CONSTANTS_SHA256


def test_generated_headers_match_constants():
    import os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for rel in ("oracle/poseidon_constants.h", "plonky2.5_b200/csrc/poseidon_constants.cuh"):
        vals = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{16})ULL", open(os.path.join(root, rel)).read())]
        assert vals == o.ALL_ROUND_CONSTANTS, rel


def test_poseidon_kat_reference_vectors(golden, oc):
    """/root/reference/src/common/poseidon2/poseidon2_goldilocks.rs:190-211"""
    for v in golden["poseidon_kat"]["vectors"]:
        assert o.poseidon(v["input"]) == v["output"]
        assert oc.poseidon(v["input"]).tolist() == v["output"]


def test_field_constants(golden, oc):
    fc = golden["field_constants"]
    assert fc["modulus"] == o.P == P
    g2 = fc["power_of_two_generator"]
    assert pow(7, (o.P - 1) >> 32, o.P) == g2
    assert pow(g2, 1 << 31, o.P) == o.P - 1 and pow(g2, 1 << 32, o.P) == 1
    assert oc.lib.glo_root_of_unity(3) == o.primitive_root_of_unity(3) == pow(2, 24 * 5, o.P)  # w_8 = 2^120


def test_derived_anchors(golden, oc):
    a = golden["derived_anchors"]
    h = o.hash_no_pad(list(range(135)))
    assert h == a["hash_no_pad_0_134"]
    assert oc.hash_or_noop(np.arange(135, dtype=np.uint64)).tolist() == h
    assert o.two_to_one(h, h) == a["two_to_one_h_h"] == oc.two_to_one(h, h).tolist()
    t = a["tiny_commit"]
    pb = o.PolynomialBatch.from_values(t["values"], t["rate_bits"], t["cap_height"])
    assert pb.polynomials == t["coeffs"] and pb.merkle_tree.leaves == t["leaves"]
    # naive O(n^2) evaluation agrees with the radix-2 recursion
    for col, cf in zip(t["values"], t["coeffs"]):
        assert o.fft_naive(cf) == [v % o.P for v in col]


def test_hash_or_noop_modes(oc):
    rng = random.Random(5)
    for n in (0, 1, 3, 4):
        row = [rng.randrange(o.P) for _ in range(n)]
        assert o.hash_or_noop(row) == row + [0] * (4 - n)
        assert oc.hash_or_noop(np.array(row, dtype=np.uint64)).tolist() == row + [0] * (4 - n)
    for n in (5, 7, 8, 9, 16, 17, 135):
        row = [rng.randrange(o.P) for _ in range(n)]
        assert oc.hash_or_noop(np.array(row, dtype=np.uint64)).tolist() == o.hash_no_pad(row)
    # overwrite mode: a short last chunk keeps lanes of the previous state => differs from zero-padding
    row = [rng.randrange(o.P) for _ in range(9)]
    assert o.hash_no_pad(row) != o.hash_no_pad(row + [0] * 7)


def test_non_canonical_inputs_are_reduced(oc):
    row = np.array([o.P, o.P + 5, 2**64 - 1, 0, 1, 2, 3, 4, 5], dtype=np.uint64)
    assert oc.hash_or_noop(row).tolist() == o.hash_no_pad([int(x) % o.P for x in row])


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 8])
def test_fft_ifft_c_vs_python(oc, log_n):
    rng = random.Random(log_n)
    v = [rng.randrange(o.P) for _ in range(1 << log_n)]
    assert oc.fft(v).tolist() == o.fft(v)
    assert oc.ifft(v).tolist() == o.ifft(v)
    assert oc.fft(oc.ifft(v)).tolist() == v
    assert oc.coset_fft(v, 7).tolist() == o.coset_fft(v, 7)
    if log_n <= 5:
        assert o.fft_naive(v) == o.fft(v)


@pytest.mark.parametrize("shape", [(0, 6, 2, 1), (1, 4, 1, 1), (2, 2, 1, 0), (3, 5, 2, 1), (4, 9, 3, 4), (5, 3, 1, 2),
                                   (3, 135, 3, 2), (3, 2, 3, 6), (4, 4, 0, 0)])
def test_commit_c_vs_python(oc, shape):
    log_n, n_cols, r, h = shape
    cols = splitmix_columns(100 + log_n, n_cols, 1 << log_n)
    pb = o.PolynomialBatch.from_values([c.tolist() for c in cols], r, h)
    res = oc.commit(cols, r, h)
    assert res["coeffs"].tolist() == pb.polynomials
    assert res["leaves"].tolist() == pb.merkle_tree.leaves
    assert res["digests"].tolist() == pb.merkle_tree.digests
    assert res["cap"].tolist() == pb.merkle_tree.cap
    res2 = oc.commit(res["coeffs"], r, h, is_coeffs=True)
    assert np.array_equal(res2["leaves"], res["leaves"]) and np.array_equal(res2["cap"], res["cap"])


def test_commit_golden_fixture(oc, golden):
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a, dtype="<u8").tobytes()).hexdigest()
    for c in golden["commit_small"]["cases"]:
        cols = splitmix_columns(c["seed"], c["n_cols"], 1 << c["log_n"])
        res = oc.commit(cols, c["rate_bits"], c["cap_height"], bool(c["is_coeffs"]))
        assert res["cap"].reshape(-1).tolist() == c["cap"]
        assert sha(res["coeffs"]) == c["sha256_coeffs"] and sha(res["leaves"]) == c["sha256_leaves"]
        assert sha(res["digests"]) == c["sha256_digests"]


def test_invariant_leaves_are_horner_evaluations(oc):
    """A.9(i): leaves[i][j] == P_j(7 * w_R^bitrev(i)) — independent of any FFT code."""
    log_n, n_cols, r = 6, 5, 3
    cols = splitmix_columns(7, n_cols, 1 << log_n)
    res = oc.commit(cols, r, 2)
    bits = log_n + r
    w = o.primitive_root_of_unity(bits)
    rng = random.Random(1)
    for i in [0, 1, (1 << bits) - 1] + [rng.randrange(1 << bits) for _ in range(20)]:
        x = 7 * pow(w, o.reverse_bits(i, bits), o.P) % o.P
        for j in range(n_cols):
            assert int(res["leaves"][i][j]) == o.eval_poly(res["coeffs"][j].tolist(), x)
    # and the coefficients interpolate the input values on the subgroup
    wn = o.primitive_root_of_unity(log_n)
    for k in (0, 1, 17, 63):
        for j in range(n_cols):
            assert o.eval_poly(res["coeffs"][j].tolist(), pow(wn, k, o.P)) == int(cols[j][k])


def test_invariant_merkle_paths_verify_with_verifier_rule(oc):
    """A.9(iii): closed-form prove indices + verifier rule, for the C layout, incl. cap_height == log2(n)."""
    for (log_l, ll, h) in [(3, 7, 0), (4, 135, 2), (6, 32, 4), (4, 5, 4), (7, 9, 3), (5, 3, 1)]:
        leaves = splitmix_columns(log_l * 31 + ll, 1 << log_l, ll)
        dig, cap = oc.merkle_new(leaves, h)
        tree = o.MerkleTree([r.tolist() for r in leaves], h) if ll <= 9 or log_l <= 4 else None
        if tree is not None:
            assert dig.tolist() == tree.digests and cap.tolist() == tree.cap
        sub = (1 << log_l) >> h
        per = 2 * (sub - 1)
        for idx in range(1 << log_l):
            t, j = divmod(idx, sub)
            sibs = []
            for layer in range(sub.bit_length() - 1):
                sibs.append(dig[t * per + o.digest_index(layer, j ^ 1)])
                j >>= 1
            assert oc.verify_path(leaves[idx], idx, sibs, cap)
        if sub > 1:
            bad = leaves[0].copy()
            bad[0] ^= np.uint64(1)
            t, j = 0, 0
            sibs = [dig[o.digest_index(layer, (0 >> layer) ^ 1)] for layer in range(sub.bit_length() - 1)]
            assert not oc.verify_path(bad, 0, sibs, cap)


def test_merkle_rejects_oversized_cap(oc):
    with pytest.raises(ValueError):
        oc.merkle_new(np.zeros((4, 5), dtype=np.uint64), 3)
    with pytest.raises(AssertionError):
        o.MerkleTree([[1, 2, 3, 4, 5]] * 4, 3)


def test_challenger_c_vs_python(oc):
    rng = random.Random(3)
    cp, cc = o.Challenger(), oc.new_challenger()
    for step in range(40):
        k = rng.randrange(1, 13)
        es = [rng.randrange(o.P) for _ in range(k)]
        cp.observe_elements(es)
        cc.observe_elements(es)
        for _ in range(rng.randrange(0, 4)):
            assert cp.get_challenge() == cc.get_challenge()


def _fri_inputs(log_len, rate_bits, seed):
    """coeffs with the upper (1 - 2^-r) fraction zero, values = coset_fft(coeffs, 7)."""
    n = 1 << log_len
    low = n >> rate_bits
    rng = random.Random(seed)
    coeffs = [(rng.randrange(o.P), rng.randrange(o.P)) if i < low else (0, 0) for i in range(n)]
    values = o.ext_coset_fft(coeffs, 7)
    return coeffs, values


@pytest.mark.parametrize("cfg", [(8, 3, 1, [4]), (9, 1, 2, [2, 3]), (10, 3, 2, [4, 4])])
def test_fri_committed_trees_c_vs_python_and_fold_invariant(oc, cfg):
    log_len, rate_bits, cap_h, arities = cfg
    coeffs, values = _fri_inputs(log_len, rate_bits, log_len)
    cp, cc = o.Challenger(), oc.new_challenger()
    seed_obs = [11, 22, 33]
    cp.observe_elements(seed_obs); cc.observe_elements(seed_obs)
    trees, final = o.fri_committed_trees(coeffs, values, cp, arities, rate_bits, cap_h)
    res = oc.fri_committed_trees(np.array(coeffs, dtype=np.uint64), np.array(values, dtype=np.uint64), arities, rate_bits, cap_h, cc)
    for t, lv, dg, cap in zip(trees, res["leaves"], res["digests"], res["caps"]):
        assert lv.tolist() == t.leaves and cap.tolist() == t.cap
        if t.digests:
            assert dg.tolist() == t.digests
    assert res["final_poly"].tolist() == [list(c) for c in final]
    assert cp.get_challenge() == cc.get_challenge()          # transcripts stayed in lock-step
    # A.9(iv): the folded codeword evaluates the beta-combination of the decimated polynomials: check layer 1 by Horner
    beta = tuple(int(x) for x in res["betas"][0])
    arity = 1 << arities[0]
    folded = []
    for m in range(len(coeffs) // arity):
        acc = (0, 0)
        for c in reversed(coeffs[arity * m:arity * m + arity]):
            acc = o.ext_add(o.ext_mul(acc, beta), c)
        folded.append(acc)
    if len(arities) > 1:
        shift = pow(7, arity, o.P)
        bits = log_len - arities[0]
        w = o.primitive_root_of_unity(bits)
        lv1 = res["leaves"][1].reshape(-1, 2)
        for i in (0, 1, 5, (1 << bits) - 1):
            x = shift * pow(w, o.reverse_bits(i, bits), o.P) % o.P
            assert tuple(int(v) for v in lv1[i]) == o.ext_eval_poly(folded, x)


def test_reduction_arity_bits_standard_recursion():
    # ConstantArityBits(4, 5), rate_bits 3, cap_height 4 (standard_recursion_config, src/p3/mod.rs:231)
    assert o.reduction_arity_bits_constant(4, 5, 16, 3, 4) == [4, 4, 4]
    assert o.reduction_arity_bits_constant(4, 5, 13, 3, 4) == [4, 4]


def test_pow_c_matches_python_and_is_smallest(oc):
    """fri_proof_of_work: C restatement == Python restatement, the witness is the smallest one, transcript advances alike."""
    import random
    rng = random.Random(4)
    for bits in (0, 3, 7, 9):
        a, b = o.Challenger(), oc.new_challenger()
        es = [rng.randrange(o.P) for _ in range(rng.randrange(0, 7))]
        a.observe_elements(es)
        b.observe_elements(es)
        base = list(a.sponge_state)
        for i, v in enumerate(a.input_buffer):
            base[i] = v
        pos = len(a.input_buffer)
        wa = o.fri_proof_of_work(a, bits)
        wb = b.fri_proof_of_work(bits)
        assert wa == wb
        for w in range(wa):      # nothing smaller qualifies
            st = list(base); st[pos] = w
            assert 64 - o.poseidon(st)[7].bit_length() < bits
        assert a.get_challenge() == b.get_challenge()


# ---- prove_openings, front half: the two restatements agree and satisfy the defining identities ------------------------
def _openings_case(seed, log_n, widths, batches_spec):
    rng = random.Random(seed)
    n = 1 << log_n
    oracles = [splitmix_columns(seed * 7 + k, w, n) for k, w in enumerate(widths)]
    batches = []
    for polys in batches_spec:
        point = (rng.randrange(o.P), rng.randrange(o.P))
        batches.append((point, polys))
    alpha = (rng.randrange(o.P), rng.randrange(o.P))
    return oracles, batches, alpha


OPENINGS_CASES = [
    (1, 3, [3, 2], [[(0, 0), (0, 1), (0, 2), (1, 0), (1, 1)], [(1, 1)]]),
    (2, 5, [4], [[(0, 3), (0, 1)], [(0, 0)], [(0, 2), (0, 2), (0, 0)]]),
    (3, 0, [2], [[(0, 0), (0, 1)]]),
    (4, 6, [9, 5, 2, 2], [[(0, i) for i in range(9)] + [(1, i) for i in range(5)] + [(2, 0), (2, 1), (3, 0), (3, 1)], [(2, 0), (2, 1)]]),
]


@pytest.mark.parametrize("case", OPENINGS_CASES)
def test_openings_front_half_restatements_agree(oc, case):
    oracles, batches, alpha = _openings_case(*case)
    final_py, quot_py = o.prove_openings_final_poly(batches, [[list(map(int, c)) for c in orc] for orc in oracles], alpha)
    final_c, quot_c = oc.openings_final_poly(batches, oracles, alpha)
    assert [list(map(int, r)) for r in final_c] == [list(e) for e in final_py]
    for qc, qp in zip(quot_c, quot_py):
        assert [list(map(int, r)) for r in qc] == [list(e) for e in qp]


@pytest.mark.parametrize("case", OPENINGS_CASES)
def test_openings_front_half_identities(case):
    """(X - z) * quotient + F(z) == F with F = sum_j alpha^j f_j, and final = sum_i alpha^(polys after batch i) * quotient_i."""
    oracles, batches, alpha = _openings_case(*case)
    orc = [[list(map(int, c)) for c in x] for x in oracles]
    final, quots = o.prove_openings_final_poly(batches, orc, alpha)
    n = len(orc[0][0])
    later = 0
    expect = [(0, 0)] * n
    for (point, polys), q in reversed(list(zip(batches, quots))):
        F = [(0, 0)] * n
        for j, (oi, pi) in enumerate(polys):
            F = [o.ext_add(a, o.ext_scale(o.ext_pow(alpha, j), c)) for a, c in zip(F, orc[oi][pi])]
        Fz = o.ext_eval_poly_ext(F, point)
        assert q[-1] == (0, 0)
        back = [(0, 0)] * n                      # (X - z) * q
        for k in range(n - 1):
            back[k + 1] = o.ext_add(back[k + 1], q[k])
            back[k] = o.ext_sub(back[k], o.ext_mul(point, q[k]))
        back[0] = o.ext_add(back[0], Fz)
        assert back == F
        s = o.ext_pow(alpha, later)
        expect = [o.ext_add(e, o.ext_mul(s, c)) for e, c in zip(expect, q)]
        later += len(polys)
    assert expect == final
    lde, vals = o.prove_openings_lde(final, 2)
    x = o.MULTIPLICATIVE_GROUP_GENERATOR * pow(o.primitive_root_of_unity(o.log2_strict(len(lde))), 3, o.P) % o.P if len(lde) > 4 else None
    if x is not None:
        assert vals[3] == o.ext_eval_poly(final, x)


def test_fri_query_rounds_restatement_verifies():
    """oracle-only: prove_openings front half -> fri_committed_trees -> fri_prover_query_rounds; every opened path verifies with the
    verifier's rule and the commit-phase evaluations re-assemble the committed leaf."""
    log_n, r, cap_h, arities = 4, 2, 2, [2, 2]
    n = 1 << log_n
    cols = [[int(v) for v in c] for c in splitmix_columns(55, 5, n)]
    batch = o.PolynomialBatch.from_values(cols, r, cap_h)
    ch = o.Challenger()
    ch.observe_elements([4, 5, 6])
    alpha = ch.get_extension_challenge()
    final, _ = o.prove_openings_final_poly([((3, 9), [(0, i) for i in range(5)])], [batch.polynomials], alpha)
    lde, vals = o.prove_openings_lde(final, r)
    trees, final_poly = o.fri_committed_trees(lde, vals, ch, arities, r, cap_h)
    o.fri_proof_of_work(ch, 3)
    rounds = o.fri_prover_query_rounds([batch.merkle_tree], trees, ch, 6, arities, n << r)
    assert len(rounds) == 6
    for rnd in rounds:
        x = rnd["x_index"]
        row, proof = rnd["initial_trees_proof"][0]
        assert o.verify_merkle_proof_to_cap(row, x, batch.merkle_tree.cap, proof)
        for arity_bits, tree, step in zip(arities, trees, rnd["steps"]):
            leaf = tree.get(x >> arity_bits)
            evals = list(step["evals"])
            assert len(evals) == 1 << arity_bits                       # validate_shape: evals.len() == arity
            flat = [w for e in evals for w in e]
            assert flat == list(leaf)
            assert o.verify_merkle_proof_to_cap(flat, x >> arity_bits, tree.cap, step["merkle_proof"])
            x >>= arity_bits


def test_openings_fri_golden_fixture_python_oracle(golden):
    """The Python big-int restatement reproduces the committed openings/FRI fixture (generated by the C oracle)."""
    f = golden["openings_fri"]
    n = 1 << f["log_n"]
    cols = [splitmix_columns(s, w, n) for s, w in zip(f["col_seeds"], f["widths"])]
    coeffs = [[o.ifft([int(v) for v in c]) for c in x] for x in cols]
    ch = o.Challenger()
    ch.observe_elements(f["transcript_prefix"])
    alpha = ch.get_extension_challenge()
    assert list(alpha) == f["alpha"]
    batches = [(tuple(b["point"]), [tuple(x) for x in b["polynomials"]]) for b in f["batches"]]
    final, quots = o.prove_openings_final_poly(batches, coeffs, alpha)
    sha = lambda a: hashlib.sha256(np.array(a, dtype="<u8").tobytes()).hexdigest()
    assert sha(final) == f["sha256_final_poly"] and [sha(q) for q in quots] == f["sha256_quotients"]
    lde, vals = o.prove_openings_lde(final, f["rate_bits"])
    trees, fp = o.fri_committed_trees(lde, vals, ch, f["arity_bits"], f["rate_bits"], f["cap_height"])
    assert [[w for h in t.cap for w in h] for t in trees] == f["commit_phase_caps"]
    assert [list(e) for e in fp] == f["fri_final_poly"]
    assert o.fri_proof_of_work(ch, f["pow_bits"]) == f["pow_witness"]
    assert [ch.get_challenge() % (n << f["rate_bits"]) for _ in f["query_indices"]] == f["query_indices"]


# ---- the prover-side restatement against a restatement of the VERIFIER (oracle/fri_verifier.py) --------------------------------------
def _prove_for_verifier(log_n, r, cap_h, arities, pow_bits, n_queries, seed):
    """oracle-only prover run shaped like plonky2's: two oracles, openings at zeta (all polynomials) and g*zeta (the second oracle's)"""
    n = 1 << log_n
    cols_a = [[int(v) for v in c] for c in splitmix_columns(seed, 5, n)]
    cols_b = [[int(v) for v in c] for c in splitmix_columns(seed + 1, 3, n)]
    batch_a, batch_b = o.PolynomialBatch.from_values(cols_a, r, cap_h), o.PolynomialBatch.from_values(cols_b, r, cap_h)
    ch = o.Challenger()
    ch.observe_cap(batch_a.merkle_tree.cap)
    ch.observe_cap(batch_b.merkle_tree.cap)
    zeta = ch.get_extension_challenge()
    g = o.primitive_root_of_unity(log_n)
    batches = [(zeta, [(0, i) for i in range(5)] + [(1, i) for i in range(3)]), (o.ext_scale(zeta, g), [(1, i) for i in range(3)])]
    oracles = [batch_a.polynomials, batch_b.polynomials]
    opened = [[o.ext_eval_poly_ext([(c, 0) for c in oracles[oi][pi]], z) for (oi, pi) in polys] for z, polys in batches]
    for vals in opened:                                          # the openings are observed before alpha is drawn
        for v in vals:
            ch.observe_extension_element(v)
    import copy
    verifier_ch = copy.deepcopy(ch)                              # the verifier reaches the same transcript state on its own
    alpha = ch.get_extension_challenge()
    final, _ = o.prove_openings_final_poly(batches, oracles, alpha)
    lde, vals = o.prove_openings_lde(final, r)
    trees, final_poly = o.fri_committed_trees(lde, vals, ch, arities, r, cap_h)
    pow_witness = o.fri_proof_of_work(ch, pow_bits)
    rounds = o.fri_prover_query_rounds([batch_a.merkle_tree, batch_b.merkle_tree], trees, ch, n_queries, arities, n << r)
    return dict(batches=batches, opened_values=opened, initial_caps=[batch_a.merkle_tree.cap, batch_b.merkle_tree.cap],
                commit_caps=[t.cap for t in trees], final_poly=final_poly, pow_witness=pow_witness, query_rounds=rounds,
                reduction_arity_bits=arities, log_n=log_n, rate_bits=r, pow_bits=pow_bits), verifier_ch


@pytest.mark.parametrize("cfg", [(5, 2, 1, [2, 1], 4, 5), (6, 1, 2, [3], 3, 4), (6, 3, 0, [1, 2, 1], 2, 4), (4, 2, 2, [], 3, 3)])
def test_prover_restatement_passes_the_verifier_restatement(cfg):
    """What verify_fri_proof enforces holds for the prover restatement's proof: Merkle paths of the opened rows, the first layer =
    sum_b alpha^(k_b) (F_b(x) - F_b(z_b)) / (x - z_b) from the opened rows, every fold = the coset interpolated and evaluated at beta,
    the last value = final_poly(x), proof of work, transcript order."""
    import copy
    import fri_verifier as fv
    proof, ch = _prove_for_verifier(*cfg, seed=91)
    fv.verify_fri_proof(challenger=copy.deepcopy(ch), **proof)
    # and it is a check: every one of these single-word changes is rejected
    def tampered(mut):
        p = copy.deepcopy(proof)
        mut(p)
        with pytest.raises(AssertionError):
            fv.verify_fri_proof(challenger=copy.deepcopy(ch), **p)
    def bump(e):
        return ((e[0] + 1) % o.P, e[1])
    tampered(lambda p: p["opened_values"][0].__setitem__(2, bump(p["opened_values"][0][2])))          # a claimed opening
    tampered(lambda p: p["opened_values"][1].__setitem__(0, bump(p["opened_values"][1][0])))
    tampered(lambda p: p["final_poly"].__setitem__(0, bump(p["final_poly"][0])))                        # the final polynomial
    tampered(lambda p: p.__setitem__("pow_witness", p["pow_witness"] + 1))                              # transcript / proof of work
    if cfg[3]:
        def step_eval(p):
            ev = list(p["query_rounds"][0]["steps"][0]["evals"])
            ev[0] = bump(ev[0])
            p["query_rounds"][0]["steps"][0]["evals"] = ev
        tampered(step_eval)                                                                            # a commit-phase evaluation
    def leaf_word(p):
        row, path = p["query_rounds"][1]["initial_trees_proof"][0]
        row = list(row)
        row[0] = (row[0] + 1) % o.P
        p["query_rounds"][1]["initial_trees_proof"][0] = (row, path)
    tampered(leaf_word)                                                                                # an opened leaf row
