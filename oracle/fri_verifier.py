"""TEST INFRASTRUCTURE — CPU restatement of the VERIFIER side of plonky2's FRI opening protocol, used to check the prover-side
restatement (oracle/gl_oracle.py, which the CUDA path matches bit for bit) against the relations a verifier enforces.

Restates (from memory of plonky2 @ 3de92d9ed1721cec133e4e1e1b3ec7facb756ccf, the un-vendored dependency at
/root/reference/Cargo.toml:15-19; parity unpinned — no verifier vectors exist in /root/reference):
    plonky2/src/fri/verifier.rs · verify_fri_proof, fri_verify_proof_of_work, PrecomputedReducedOpenings::from_os_and_alpha,
                                  fri_combine_initial, fri_verifier_query_round, compute_evaluation
    plonky2/src/fri/validate_shape.rs · validate_fri_proof_shape (the parts that concern the query rounds)
    plonky2/src/hash/merkle_proofs.rs · verify_merkle_proof_to_cap   (gl_oracle.verify_merkle_proof_to_cap)

Why it is worth having although it cannot be pinned either: every relation below is forced by the mathematics once the prover's
objects are fixed — the leaf at index i holds the batch evaluated at 7·w^bitrev(i), the first FRI layer is
sum_b alpha^(k_b) (F_b(x) - F_b(z_b)) / (x - z_b), a layer's coset interpolated and evaluated at beta is the next layer's value at
x^arity, the last one equals final_poly(x) — so a proof produced by the prover restatement only verifies if index order, coset
order, the alpha bookkeeping of ReducingFactor and the fold agree with each other.  Only tests/ imports this module.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import gl_oracle as o

Ext = Tuple[int, int]
P = o.P


def ext_inv(a: Ext) -> Ext:
    """1 / (a0 + a1 X) in F_p[X] / (X^2 - 7): conj / norm"""
    a0, a1 = a
    norm = (a0 * a0 - 7 * a1 * a1) % P
    ni = o.inv(norm)
    return (a0 * ni % P, (P - a1) % P * ni % P)


def ext_div(a: Ext, b: Ext) -> Ext:
    return o.ext_mul(a, ext_inv(b))


def reduce_ext(base: Ext, values: Sequence[Ext]) -> Ext:
    """ReducingFactor::reduce: sum_j base^j * values[j] (Horner from the last element)"""
    acc = (0, 0)
    for v in reversed(list(values)):
        acc = o.ext_add(o.ext_mul(acc, base), v)
    return acc


def precomputed_reduced_openings(opened_values: Sequence[Sequence[Ext]], alpha: Ext) -> List[Ext]:
    """PrecomputedReducedOpenings::from_os_and_alpha: per batch, sum_j alpha^j * f_j(z)"""
    return [reduce_ext(alpha, vals) for vals in opened_values]


def fri_combine_initial(batches, initial_rows: Sequence[Sequence[int]], alpha: Ext, subgroup_x: int, reduced_openings: Sequence[Ext]) -> Ext:
    """fri_combine_initial: batches = [(point, [(oracle_index, polynomial_index), ...])]; initial_rows[o] = the opened leaf row of oracle o.
    sum = shift(sum) + (reduce(evals) - reduced_openings) / (x - point), the shift being alpha^(number of evals of THIS batch)."""
    x = (subgroup_x % P, 0)
    total = (0, 0)
    for (point, polys), ro in zip(batches, reduced_openings):
        evals = [(initial_rows[oi][pi] % P, 0) for (oi, pi) in polys]
        reduced = reduce_ext(alpha, evals)
        numerator = o.ext_sub(reduced, ro)
        denominator = o.ext_sub(x, point)
        total = o.ext_mul(total, o.ext_pow(alpha, len(evals)))
        total = o.ext_add(total, ext_div(numerator, denominator))
    return total


def compute_evaluation(x: int, x_index_within_coset: int, arity_bits: int, evals: Sequence[Ext], beta: Ext) -> Ext:
    """compute_evaluation: interpolate {(coset_start * g^k, evals_bitreversed[k])} and evaluate at beta"""
    arity = 1 << arity_bits
    assert len(evals) == arity
    g = o.primitive_root_of_unity(arity_bits)
    ev = [evals[o.reverse_bits(k, arity_bits)] for k in range(arity)]          # reverse_index_bits_in_place
    rev = o.reverse_bits(x_index_within_coset, arity_bits)
    coset_start = x * pow(g, arity - rev, P) % P
    pts = [coset_start * pow(g, k, P) % P for k in range(arity)]
    # Lagrange form (upstream uses barycentric weights: the same polynomial)
    acc = (0, 0)
    for k in range(arity):
        num: Ext = (1, 0)
        den = 1
        for m in range(arity):
            if m != k:
                num = o.ext_mul(num, o.ext_sub(beta, (pts[m], 0)))
                den = den * ((pts[k] - pts[m]) % P) % P
        acc = o.ext_add(acc, o.ext_mul(ev[k], o.ext_scale(num, o.inv(den))))
    return acc


def verify_pow(challenger: o.Challenger, pow_witness: int, min_leading_zeros: int) -> bool:
    """get_challenges: observe the witness, draw fri_pow_response; fri_verify_proof_of_work: enough leading zeros (64-bit view)"""
    challenger.observe_element(pow_witness)
    resp = challenger.get_challenge()
    return 64 - resp.bit_length() >= min_leading_zeros


def verify_fri_proof(*, batches, opened_values, initial_caps, commit_caps, final_poly: Sequence[Ext], pow_witness: int, query_rounds,
                     challenger: o.Challenger, reduction_arity_bits: Sequence[int], log_n: int, rate_bits: int, pow_bits: int) -> None:
    """verify_fri_proof with the challenger in the state the PROVER had when prove_openings started (openings already observed).
    Raises AssertionError with the upstream ensure!() text on the first violated relation."""
    lde_bits = log_n + rate_bits
    n = 1 << lde_bits
    # ---- get_challenges ----
    alpha = challenger.get_extension_challenge()
    betas = []
    for cap in commit_caps:
        challenger.observe_cap(cap)
        betas.append(challenger.get_extension_challenge())
    for c in final_poly:
        challenger.observe_extension_element(c)
    assert verify_pow(challenger, pow_witness, pow_bits), "Invalid proof of work witness."
    indices = [challenger.get_challenge() % n for _ in range(len(query_rounds))]
    # ---- validate shape (what concerns this restatement) ----
    assert len(final_poly) == (1 << log_n) >> sum(reduction_arity_bits), "Final polynomial has wrong degree."
    reduced = precomputed_reduced_openings(opened_values, alpha)
    for x_index, rnd in zip(indices, query_rounds):
        assert rnd["x_index"] == x_index
        assert len(rnd["steps"]) == len(reduction_arity_bits)
        rows = []
        for (row, proof), cap in zip(rnd["initial_trees_proof"], initial_caps):
            assert o.verify_merkle_proof_to_cap(row, x_index, cap, proof), "Invalid Merkle proof."
            rows.append(row)
        subgroup_x = o.MULTIPLICATIVE_GROUP_GENERATOR * pow(o.primitive_root_of_unity(lde_bits), o.reverse_bits(x_index, lde_bits), P) % P
        old_eval = fri_combine_initial(batches, rows, alpha, subgroup_x, reduced)
        x = x_index
        for i, arity_bits in enumerate(reduction_arity_bits):
            arity = 1 << arity_bits
            step = rnd["steps"][i]
            evals = list(step["evals"])
            assert len(evals) == arity                                           # validate_shape
            coset_index, within = x >> arity_bits, x & (arity - 1)
            assert tuple(evals[within]) == tuple(old_eval), "Old evaluation does not match."
            old_eval = compute_evaluation(subgroup_x, within, arity_bits, evals, betas[i])
            flat = [w for e in evals for w in e]
            assert o.verify_merkle_proof_to_cap(flat, coset_index, commit_caps[i], step["merkle_proof"]), "Invalid Merkle proof."
            subgroup_x = pow(subgroup_x, arity, P)
            x = coset_index
        assert o.ext_eval_poly(list(final_poly), subgroup_x) == tuple(old_eval), "Final polynomial evaluation is invalid."
