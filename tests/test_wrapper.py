"""The wrapper-shaped prove pipeline (tests/wrapper_pipeline.py): product (GPU, C ABI through the Python mirror) against the CPU
oracle on the same transcript — every proof field, at a small degree and at the wrapper's own 2^16."""
import numpy as np
import pytest

import wrapper_pipeline as wp


def test_wrapper_oracle_is_self_consistent(oc):
    """CPU only: the oracle's pipeline runs, every opened path verifies with the verifier's rule, and evals[x & 15] sits where the
    verifier looks for it."""
    log_n = 6
    cols = wp.make_columns(log_n)
    proof = wp.run_oracle(oc, cols, log_n, arities=(4,), pow_bits=4, n_queries=5)
    assert len(proof["query_rounds"]) == 5 and len(proof["commit_phase_caps"]) == 1
    for x, rnd in zip(proof["query_indices"], proof["query_rounds"]):
        for (row, sib), cap in zip(rnd["initial"], proof["commit_caps"]):
            assert oc.verify_path(row, x, sib, cap)
        ev, sib = rnd["steps"][0]
        assert ev.shape == (16, 2)
        assert oc.verify_path(ev.reshape(-1), x >> 4, sib, proof["commit_phase_caps"][0])
    assert wp.digest(proof) == wp.digest(wp.run_oracle(oc, cols, log_n, arities=(4,), pow_bits=4, n_queries=5))


@pytest.mark.gpu
@pytest.mark.parametrize("log_n,arities,pow_bits", [(10, (4, 4), 8), (16, wp.ARITIES, wp.POW_BITS)])
def test_wrapper_pipeline_matches_oracle(ctx, oc, log_n, arities, pow_bits):
    import plonky25_b200 as g
    cols = wp.make_columns(log_n)
    want = wp.run_oracle(oc, cols, log_n, arities=arities, pow_bits=pow_bits)
    got = wp.run_product(g, ctx, cols, log_n, arities=arities, pow_bits=pow_bits)
    for a, b in zip(got["commit_caps"], want["commit_caps"]):
        assert np.array_equal(a, b)
    for a, b in zip(got["commit_phase_caps"], want["commit_phase_caps"]):
        assert np.array_equal(a, b)
    assert np.array_equal(got["final_poly"], want["final_poly"])
    assert got["pow_witness"] == want["pow_witness"]
    assert got["query_indices"] == want["query_indices"]
    for ra, rb in zip(got["query_rounds"], want["query_rounds"]):
        for (r1, s1), (r2, s2) in zip(ra["initial"], rb["initial"]):
            assert np.array_equal(r1, r2) and np.array_equal(s1, s2)
        for (e1, s1), (e2, s2) in zip(ra["steps"], rb["steps"]):
            assert np.array_equal(e1, e2) and np.array_equal(s1, s2)
    assert wp.digest(got) == wp.digest(want)


def test_wrapper_oracle_proof_passes_the_verifier_restatement(oc):
    """CPU only: the wrapper-shaped proof of the C oracle (the one the GPU pipeline is compared with word for word) satisfies what a
    verifier enforces (oracle/fri_verifier.py: opened rows -> first FRI layer through the alpha-reduced openings at zeta and g*zeta,
    coset interpolation at beta per layer, final polynomial, proof of work, transcript order) — 257 + 2 openings over four oracles."""
    import gl_oracle as o
    import fri_verifier as fv
    log_n, r, h, arities, pow_bits, nq = 6, 3, 4, (4,), 4, 5
    cols = wp.make_columns(log_n)
    proof = wp.run_oracle(oc, cols, log_n, r=r, h=h, arities=arities, pow_bits=pow_bits, n_queries=nq)
    commits = [oc.commit(c, r, h, is_coeffs=wp.IS_COEFFS[k], want=("coeffs",)) for k, c in enumerate(cols)]
    ch = o.Challenger()
    ch.observe_elements([int(v) for v in wp.PREFIX])
    for cm in commits:
        ch.observe_cap(cm["cap"].tolist())
    zeta = ch.get_extension_challenge()
    all_polys, zs = wp.instance_for(wp.WIDTHS)
    batches = [(zeta, all_polys), (wp._g_times(zeta, log_n), zs)]
    coeffs = [cm["coeffs"] for cm in commits]
    opened = [[o.ext_eval_poly_ext([(int(c), 0) for c in coeffs[oi][pi]], z) for (oi, pi) in polys] for z, polys in batches]
    rounds = []
    for x, rnd in zip(proof["query_indices"], proof["query_rounds"]):
        rounds.append({"x_index": x,
                       "initial_trees_proof": [([int(v) for v in row], [[int(w) for w in s] for s in sib]) for row, sib in rnd["initial"]],
                       "steps": [{"evals": [(int(e[0]), int(e[1])) for e in ev], "merkle_proof": [[int(w) for w in s] for s in sib]}
                                 for ev, sib in rnd["steps"]]})
    fv.verify_fri_proof(batches=batches, opened_values=opened, initial_caps=[cm["cap"].tolist() for cm in commits],
                        commit_caps=[c.tolist() for c in proof["commit_phase_caps"]],
                        final_poly=[(int(c[0]), int(c[1])) for c in proof["final_poly"]], pow_witness=proof["pow_witness"], query_rounds=rounds,
                        challenger=ch, reduction_arity_bits=list(arities), log_n=log_n, rate_bits=r, pow_bits=pow_bits)
