"""CPU oracle for the gate layer of SURVEY.md §8(f) ranks 3-4 — TEST INFRASTRUCTURE ONLY (imported by tests/ and bench.py's
baseline legs; the product never touches oracle/).

Unlike the commitment path, the source of these functions IS in the reference tree, so this is a line-by-line restatement of
Rust that can be read next to it ("parity pinned by the reference source"):

  poseidon2()                      /root/reference/src/common/poseidon2/poseidon2.rs:59-91      Poseidon2::poseidon2
  matmul_external / matmul_m4      poseidon2.rs:127-146, 185-245
  matmul_internal                  poseidon2.rs:164-182
  Poseidon2Gate wire layout        /root/reference/src/common/poseidon2/poseidon2_gate.rs:82-142
  poseidon2_gate_eval()            poseidon2_gate.rs:233-310   eval_unfiltered_base_one (123 constraints, degree 7)
  poseidon2_gate_witness()         poseidon2_gate.rs:447-523   Poseidon2Generator::run_once
  u32_arithmetic_eval()            /root/reference/src/common/u32/gates/arithmetic_u32.rs:103-166 (eval_unfiltered; the packed base
                                   evaluator :285-340 computes the same values), num_ops = 3 for the 135-wire / 80-routed config (:52-55)
  u32_arithmetic_witness()         arithmetic_u32.rs · U32ArithmeticGenerator::run_once (output split into halves, 2-bit limbs, inverse)
  reduce_with_powers()             plonky2 plonk/plonk_common.rs · reduce_with_powers (sum_i alpha^i * c_i), restated

Constants come from oracle/poseidon2_constants.py, generated from the reference source by tools/gen_poseidon2_constants.py.
The reference holds no Poseidon2 known-answer vector (the block at poseidon2_goldilocks.rs:190-211 is plonky2's *Poseidon*, SURVEY F6,
and its checker is tautological); these constants are not the HorizenLabs instance either, so no published vector applies.  What pins
the restatement is the source itself plus the gate's own consistency: a row filled by the generator satisfies all 123 constraints and
its output wires are poseidon2(inputs) (tests/test_gates.py).
"""
from typing import List, Sequence

from poseidon2_constants import MAT_DIAG_M_1, RC, RC_MID

P = 0xFFFF_FFFF_0000_0001
WIDTH = 12
ROUND_F_BEGIN = 4
ROUND_F_END = 8
ROUND_P = 22


def sbox_monomial(x: int) -> int:
    x2 = x * x % P
    x4 = x2 * x2 % P
    x3 = x * x2 % P
    return x3 * x4 % P


def matmul_m4(s: List[int]) -> None:
    for i in range(WIDTH // 4):
        a = 4 * i
        t0 = (s[a] + s[a + 1]) % P
        t1 = (s[a + 2] + s[a + 3]) % P
        t2 = (t1 + 2 * s[a + 1]) % P          # t_1.multiply_accumulate(input[1], 2)
        t3 = (t0 + 2 * s[a + 3]) % P
        t4 = (t3 + 4 * t1) % P
        t5 = (t2 + 4 * t0) % P
        s[a], s[a + 1], s[a + 2], s[a + 3] = (t3 + t5) % P, t5, (t2 + t4) % P, t4


def matmul_external(s: List[int]) -> None:
    matmul_m4(s)
    stored = [sum(s[4 * j + l] for j in range(WIDTH // 4)) % P for l in range(4)]
    for i in range(WIDTH):
        s[i] = (s[i] + stored[i % 4]) % P


def matmul_internal(s: List[int]) -> None:
    total = sum(s)
    for i in range(WIDTH):
        s[i] = ((MAT_DIAG_M_1[i] - 1) * s[i] + total) % P


def poseidon2(inp: Sequence[int]) -> List[int]:
    s = [int(x) % P for x in inp]
    matmul_external(s)
    for r in range(ROUND_F_BEGIN):
        s = [sbox_monomial((s[i] + RC[r][i]) % P) for i in range(WIDTH)]
        matmul_external(s)
    for r in range(ROUND_P):
        s[0] = sbox_monomial((s[0] + RC_MID[r]) % P)
        matmul_internal(s)
    for r in range(ROUND_F_BEGIN, ROUND_F_END):
        s = [sbox_monomial((s[i] + RC[r][i]) % P) for i in range(WIDTH)]
        matmul_external(s)
    return s


# ---- Poseidon2Gate -------------------------------------------------------------------------------------------------------------
WIRE_SWAP = 2 * WIDTH
START_DELTA = 2 * WIDTH + 1
START_ROUND_F_BEGIN = START_DELTA + 4
START_PARTIAL = START_ROUND_F_BEGIN + WIDTH * (ROUND_F_BEGIN - 1)
START_ROUND_F_END = START_PARTIAL + ROUND_P
POSEIDON2_NUM_WIRES = START_ROUND_F_END + WIDTH * ROUND_F_BEGIN                         # 135
POSEIDON2_NUM_CONSTRAINTS = WIDTH * (ROUND_F_END - 1) + ROUND_P + WIDTH + 1 + 4         # 123


def wire_input(i): return i
def wire_output(i): return WIDTH + i
def wire_delta(i): return START_DELTA + i
def wire_full_round_begin(r, i): return START_ROUND_F_BEGIN + WIDTH * (r - 1) + i
def wire_partial_round(r): return START_PARTIAL + r
def wire_full_round_end(r, i): return START_ROUND_F_END + WIDTH * r + i


def poseidon2_gate_eval(w: Sequence[int]) -> List[int]:
    """eval_unfiltered_base_one: the 123 constraint values of one row (all zero iff the row is a valid Poseidon2 evaluation)"""
    w = [int(x) % P for x in w]
    c = []
    swap = w[WIRE_SWAP]
    c.append(swap * (swap - 1) % P)
    for i in range(4):
        c.append((swap * (w[wire_input(i + 4)] - w[wire_input(i)]) - w[wire_delta(i)]) % P)
    s = [0] * WIDTH
    for i in range(4):
        d = w[wire_delta(i)]
        s[i] = (w[wire_input(i)] + d) % P
        s[i + 4] = (w[wire_input(i + 4)] - d) % P
    for i in range(8, WIDTH):
        s[i] = w[wire_input(i)]
    matmul_external(s)
    for r in range(ROUND_F_BEGIN):
        s = [(s[i] + RC[r][i]) % P for i in range(WIDTH)]
        if r != 0:
            for i in range(WIDTH):
                sbox_in = w[wire_full_round_begin(r, i)]
                c.append((s[i] - sbox_in) % P)
                s[i] = sbox_in
        s = [sbox_monomial(x) for x in s]
        matmul_external(s)
    for r in range(ROUND_P):
        s[0] = (s[0] + RC_MID[r]) % P
        sbox_in = w[wire_partial_round(r)]
        c.append((s[0] - sbox_in) % P)
        s[0] = sbox_monomial(sbox_in)
        matmul_internal(s)
    for r in range(ROUND_F_BEGIN, ROUND_F_END):
        s = [(s[i] + RC[r][i]) % P for i in range(WIDTH)]
        for i in range(WIDTH):
            sbox_in = w[wire_full_round_end(r - ROUND_F_BEGIN, i)]
            c.append((s[i] - sbox_in) % P)
            s[i] = sbox_in
        s = [sbox_monomial(x) for x in s]
        matmul_external(s)
    for i in range(WIDTH):
        c.append((s[i] - w[wire_output(i)]) % P)
    assert len(c) == POSEIDON2_NUM_CONSTRAINTS
    return c


def poseidon2_gate_witness(inputs: Sequence[int], swap: int) -> List[int]:
    """Poseidon2Generator::run_once: the full 135-wire row from the 12 inputs and the swap flag"""
    w = [0] * POSEIDON2_NUM_WIRES
    s = [int(x) % P for x in inputs]
    for i in range(WIDTH):
        w[wire_input(i)] = s[i]
    w[WIRE_SWAP] = swap
    for i in range(4):
        w[wire_delta(i)] = swap * (s[i + 4] - s[i]) % P
    if swap == 1:
        for i in range(4):
            s[i], s[i + 4] = s[i + 4], s[i]
    matmul_external(s)
    for r in range(ROUND_F_BEGIN):
        s = [(s[i] + RC[r][i]) % P for i in range(WIDTH)]
        if r != 0:
            for i in range(WIDTH):
                w[wire_full_round_begin(r, i)] = s[i]
        s = [sbox_monomial(x) for x in s]
        matmul_external(s)
    for r in range(ROUND_P):
        s[0] = (s[0] + RC_MID[r]) % P
        w[wire_partial_round(r)] = s[0]
        s[0] = sbox_monomial(s[0])
        matmul_internal(s)
    for r in range(ROUND_F_BEGIN, ROUND_F_END):
        s = [(s[i] + RC[r][i]) % P for i in range(WIDTH)]
        for i in range(WIDTH):
            w[wire_full_round_end(r - ROUND_F_BEGIN, i)] = s[i]
        s = [sbox_monomial(x) for x in s]
        matmul_external(s)
    for i in range(WIDTH):
        w[wire_output(i)] = s[i]
    return w


# ---- U32ArithmeticGate (num_ops = 3 at 135 wires / 80 routed) ---------------------------------------------------------------------------
U32_LIMB_BITS = 2
U32_NUM_LIMBS = 64 // U32_LIMB_BITS
U32_ROUTED_PER_OP = 6


def u32_arith_num_ops(num_wires=135, num_routed=80):
    return min(num_wires // (U32_ROUTED_PER_OP + U32_NUM_LIMBS), num_routed // U32_ROUTED_PER_OP)


def u32_arithmetic_eval(w: Sequence[int], num_ops: int = 3) -> List[int]:
    w = [int(x) % P for x in w]
    c = []
    for i in range(num_ops):
        m0, m1, addend = w[6 * i], w[6 * i + 1], w[6 * i + 2]
        out_lo, out_hi, inverse = w[6 * i + 3], w[6 * i + 4], w[6 * i + 5]
        computed = (m0 * m1 + addend) % P
        diff = (0xFFFFFFFF - out_hi) % P
        hi_not_max = (inverse * diff - 1) % P
        c.append(hi_not_max * out_lo % P)
        combined = (out_hi * (1 << 32) + out_lo) % P
        c.append((combined - computed) % P)
        lo = hi = 0
        for j in reversed(range(U32_NUM_LIMBS)):
            limb = w[U32_ROUTED_PER_OP * num_ops + U32_NUM_LIMBS * i + j]
            prod = 1
            for x in range(1 << U32_LIMB_BITS):
                prod = prod * (limb - x) % P
            c.append(prod)
            if j < U32_NUM_LIMBS // 2:
                lo = ((1 << U32_LIMB_BITS) * lo + limb) % P
            else:
                hi = ((1 << U32_LIMB_BITS) * hi + limb) % P
        c.append((lo - out_lo) % P)
        c.append((hi - out_hi) % P)
    assert len(c) == num_ops * (4 + U32_NUM_LIMBS)
    return c


def u32_arithmetic_witness(ops: Sequence[Sequence[int]], num_wires: int = 135) -> List[int]:
    """ops: num_ops triples (multiplicand_0, multiplicand_1, addend) of u32 values -> the full row (U32ArithmeticGenerator::run_once)"""
    num_ops = len(ops)
    w = [0] * num_wires
    for i, (m0, m1, a) in enumerate(ops):
        out = m0 * m1 + a
        lo, hi = out & 0xFFFFFFFF, out >> 32
        w[6 * i:6 * i + 5] = [m0, m1, a, lo, hi]
        w[6 * i + 5] = pow((0xFFFFFFFF - hi) % P, P - 2, P)          # inverse of (u32::MAX - output_high); 0 when that is 0
        for j in range(U32_NUM_LIMBS):
            w[U32_ROUTED_PER_OP * num_ops + U32_NUM_LIMBS * i + j] = (out >> (U32_LIMB_BITS * j)) & ((1 << U32_LIMB_BITS) - 1)
    return w


def reduce_with_powers(terms: Sequence[int], alpha: int, start_power: int = 0) -> int:
    acc, pw = 0, pow(alpha, start_power, P)
    for t in terms:
        acc = (acc + t * pw) % P
        pw = pw * alpha % P
    return acc


# ---- permutation argument: partial products and Z (plonky2 plonk/prover.rs · wires_permutation_partial_products_and_zs) -------------
# Upstream source is NOT under /root/reference (plonky2 @ 3de92d9, un-vendored): restated from the published algorithm — parity
# unpinned.  The independent check is the defining identity of the argument (tests/test_gates.py): for wires that satisfy a copy
# permutation sigma, Z closes (the running product returns to 1 after the last row).
def partial_products_and_zs(wires, sigmas, k_is, betas, gammas, degree, log_n):
    """wires / sigmas: [n_routed][N] columns (values on the subgroup, natural order).  Returns the polynomials in prove()'s order:
    [Z_0, .., Z_{c-1}, pp_0 of challenge 0, .., pp of challenge c-1 ..], each a list of N values."""
    n_routed, n = len(wires), 1 << log_n
    w = pow(1753635133440165772, 1 << (32 - log_n), P)
    xs = [pow(w, i, P) for i in range(n)]
    n_chunks = -(-n_routed // degree)
    zs, pps = [], []
    for beta, gamma in zip(betas, gammas):
        z_x = 1
        z_col, pp_cols = [], [[] for _ in range(n_chunks - 1)]
        for i in range(n):
            quotients = []
            for j in range(n_routed):
                num = (wires[j][i] + beta * k_is[j] % P * xs[i] + gamma) % P
                den = (wires[j][i] + beta * sigmas[j][i] + gamma) % P
                quotients.append(num * pow(den, P - 2, P) % P)
            chunk_products = []
            for c in range(n_chunks):
                prod = 1
                for q in quotients[c * degree:(c + 1) * degree]:
                    prod = prod * q % P
                chunk_products.append(prod)
            acc, res = z_x, []
            for q in chunk_products:                      # partial_products_and_z_gx
                acc = acc * q % P
                res.append(acc)
            z_col.append(z_x)                             # "the last term is Z(gx), but we replace it with Z(x)"
            z_x = res[-1]
            for c in range(n_chunks - 1):
                pp_cols[c].append(res[c])
        zs.append(z_col)
        pps.extend(pp_cols)
    return zs + pps


# ---- the reference's other u32 / b32 gates (all under /root/reference/src/common/u32/gates/) ----------------------------------------
# Each *_eval follows the gate's eval_unfiltered line by line (the packed base evaluators compute the same values); each *_witness fills a
# valid row the way the gate's generator does, for tests.  plonky2's reduce_with_powers(terms, a) = sum_i terms[i] * a^i.
def _range_product(v, n):
    prod = 1
    for x in range(n):
        prod = prod * (v - x) % P
    return prod


def _rwp(terms, alpha):
    acc = 0
    for t in reversed(list(terms)):
        acc = (acc * alpha + t) % P
    return acc


def add_many_num_ops(num_addends, num_wires=135, num_routed=80):
    """add_many_u32.rs:52-57"""
    return min(num_wires // (num_addends + 3 + 18), num_routed // (num_addends + 3))


def add_many_eval(w, num_addends, num_ops):
    """add_many_u32.rs:103-143 (result limbs: 16 x 2 bits, carry limbs: 2 x 2 bits)"""
    w = [int(x) % P for x in w]
    c, stride = [], num_addends + 3
    for i in range(num_ops):
        addends = w[stride * i:stride * i + num_addends]
        carry, out_res, out_carry = w[stride * i + num_addends], w[stride * i + num_addends + 1], w[stride * i + num_addends + 2]
        computed = (sum(addends) + carry) % P
        c.append((out_carry * (1 << 32) + out_res - computed) % P)
        res_l = car_l = 0
        for j in reversed(range(18)):
            limb = w[stride * num_ops + 18 * i + j]
            c.append(_range_product(limb, 4))
            if j < 16:
                res_l = (4 * res_l + limb) % P
            else:
                car_l = (4 * car_l + limb) % P
        c.append((res_l - out_res) % P)
        c.append((car_l - out_carry) % P)
    return c


def add_many_witness(ops, num_addends, num_wires=135):
    """ops: per op (addends[num_addends] u32, carry u32)"""
    num_ops, stride = len(ops), num_addends + 3
    w = [0] * num_wires
    for i, (addends, carry) in enumerate(ops):
        out = sum(addends) + carry
        res, car = out & 0xFFFFFFFF, out >> 32
        w[stride * i:stride * i + num_addends] = list(addends)
        w[stride * i + num_addends:stride * i + num_addends + 3] = [carry, res, car]
        for j in range(18):
            w[stride * num_ops + 18 * i + j] = ((res >> (2 * j)) & 3) if j < 16 else ((car >> (2 * (j - 16))) & 3)
    return w


def subtraction_eval(w, num_ops):
    """subtraction_u32.rs:100-134"""
    w = [int(x) % P for x in w]
    c = []
    for i in range(num_ops):
        x, y, borrow, out_res, out_borrow = w[5 * i:5 * i + 5]
        initial = (x - y - borrow) % P
        c.append((out_res - (initial + (1 << 32) * out_borrow)) % P)
        comb = 0
        for j in reversed(range(16)):
            limb = w[5 * num_ops + 16 * i + j]
            c.append(_range_product(limb, 4))
            comb = (4 * comb + limb) % P
        c.append((comb - out_res) % P)
        c.append(out_borrow * (1 - out_borrow) % P)
    return c


def subtraction_witness(ops, num_wires=135):
    num_ops = len(ops)
    w = [0] * num_wires
    for i, (x, y, b) in enumerate(ops):
        d = x - y - b
        res, bo = (d, 0) if d >= 0 else (d + (1 << 32), 1)
        w[5 * i:5 * i + 5] = [x, y, b, res, bo]
        for j in range(16):
            w[5 * num_ops + 16 * i + j] = (res >> (2 * j)) & 3
    return w


def range_check_eval(w, num_input_limbs):
    """range_check_u32.rs:70-92"""
    w = [int(x) % P for x in w]
    c = []
    for i in range(num_input_limbs):
        aux = w[num_input_limbs + 16 * i:num_input_limbs + 16 * i + 16]
        c.append((_rwp(aux, 4) - w[i]) % P)
        for a in aux:
            c.append(_range_product(a, 4))
    return c


def range_check_witness(limbs, num_wires=135):
    n = len(limbs)
    w = [0] * num_wires
    for i, v in enumerate(limbs):
        w[i] = v
        for j in range(16):
            w[n + 16 * i + j] = (v >> (2 * j)) & 3
    return w


def interleave_eval(w, num_ops):
    """interleave_u32.rs:104-140: bits are big-endian"""
    w = [int(x) % P for x in w]
    c = []
    for i in range(num_ops):
        x, x_int = w[2 * i], w[2 * i + 1]
        bits = w[2 * num_ops + 32 * i:2 * num_ops + 32 * (i + 1)]
        c.append((_rwp(reversed(bits), 2) - x) % P)
        c.append((_rwp(reversed(bits), 4) - x_int) % P)
        for b in bits:
            c.append(_range_product(b, 2))
    return c


def interleave_witness(xs, num_wires=135):
    num_ops = len(xs)
    w = [0] * num_wires
    for i, x in enumerate(xs):
        bits = [(x >> (31 - k)) & 1 for k in range(32)]
        w[2 * i] = x
        w[2 * i + 1] = sum(b << (2 * (31 - k)) for k, b in enumerate(bits))
        w[2 * num_ops + 32 * i:2 * num_ops + 32 * (i + 1)] = bits
    return w


def uninterleave_eval(w, num_ops, to_b32):
    """uninterleave_to_u32.rs:91-134 (to_b32=False) / uninterleave_to_b32.rs:115-168 (True)"""
    w = [int(x) % P for x in w]
    c = []
    for i in range(num_ops):
        x_int, evens, odds = w[3 * i:3 * i + 3]
        bits = w[3 * num_ops + 64 * i:3 * num_ops + 64 * (i + 1)]
        c.append((_rwp(reversed(bits), 2) - x_int) % P)
        ce = co = 0
        for j in range(32):
            coeff = (1 << (2 * (32 - j - 1))) if to_b32 else (1 << (32 - j - 1))
            ce = (ce + coeff * bits[2 * j]) % P
            co = (co + coeff * bits[2 * j + 1]) % P
        c.append((ce - evens) % P)
        c.append((co - odds) % P)
        for b in bits:
            c.append(_range_product(b, 2))
    return c


def uninterleave_witness(xs, to_b32, num_wires=135):
    num_ops = len(xs)
    w = [0] * num_wires
    for i, x in enumerate(xs):
        bits = [(x >> (63 - k)) & 1 for k in range(64)]
        ce = co = 0
        for j in range(32):
            coeff = (1 << (2 * (32 - j - 1))) if to_b32 else (1 << (32 - j - 1))
            ce += coeff * bits[2 * j]
            co += coeff * bits[2 * j + 1]
        w[3 * i:3 * i + 3] = [x, ce, co]
        w[3 * num_ops + 64 * i:3 * num_ops + 64 * (i + 1)] = bits
    return w


def comparison_eval(w, num_bits, num_chunks):
    """comparison.rs:112-190 (ComparisonGate::eval_unfiltered)"""
    w = [int(x) % P for x in w]
    chunk_bits = -(-num_bits // num_chunks)
    c = []
    first_input, second_input = w[0], w[1]
    fc = w[4:4 + num_chunks]
    sc = w[4 + num_chunks:4 + 2 * num_chunks]
    c.append((_rwp(fc, 1 << chunk_bits) - first_input) % P)
    c.append((_rwp(sc, 1 << chunk_bits) - second_input) % P)
    msd = 0
    for i in range(num_chunks):
        c.append(_range_product(fc[i], 1 << chunk_bits))
        c.append(_range_product(sc[i], 1 << chunk_bits))
        diff = (sc[i] - fc[i]) % P
        eq_dummy, chunks_equal = w[4 + 2 * num_chunks + i], w[4 + 3 * num_chunks + i]
        c.append((diff * eq_dummy - (1 - chunks_equal)) % P)
        c.append(chunks_equal * diff % P)
        inter = w[4 + 4 * num_chunks + i]
        c.append((inter - chunks_equal * msd) % P)
        msd = (inter + (1 - chunks_equal) * diff) % P
    c.append((w[3] - msd) % P)
    bits = w[4 + 5 * num_chunks:4 + 5 * num_chunks + chunk_bits + 1]
    for b in bits:
        c.append(b * (1 - b) % P)
    c.append(((1 << chunk_bits) + w[3] - _rwp(bits, 2)) % P)
    c.append((w[2] - bits[chunk_bits]) % P)
    return c


def comparison_witness(a, b, num_bits, num_chunks, num_wires=135):
    """ComparisonGenerator::run_once: result = (a <= b)"""
    chunk_bits = -(-num_bits // num_chunks)
    w = [0] * num_wires
    w[0], w[1] = a, b
    fc = [(a >> (chunk_bits * i)) & ((1 << chunk_bits) - 1) for i in range(num_chunks)]
    sc = [(b >> (chunk_bits * i)) & ((1 << chunk_bits) - 1) for i in range(num_chunks)]
    msd = 0
    for i in range(num_chunks):
        diff = (sc[i] - fc[i]) % P
        eq = 1 if diff == 0 else 0
        w[4 + i], w[4 + num_chunks + i] = fc[i], sc[i]
        w[4 + 2 * num_chunks + i] = pow(diff, P - 2, P) if diff else 1        # equality_dummy: 1/diff, or 1 when the chunks are equal
        w[4 + 3 * num_chunks + i] = eq
        w[4 + 4 * num_chunks + i] = eq * msd % P
        msd = (w[4 + 4 * num_chunks + i] + (1 - eq) * diff) % P
    w[3] = msd
    v = ((1 << chunk_bits) + msd) % P
    for k in range(chunk_bits + 1):
        w[4 + 5 * num_chunks + k] = (v >> k) & 1
    w[2] = w[4 + 5 * num_chunks + chunk_bits]
    return w


def vanishing_permutation_terms(wires_row, sigmas_row, zs_row, zs_next_row, x, k_is, betas, gammas, degree, log_n):
    """plonky2 plonk/vanishing_poly.rs · eval_vanishing_poly (the permutation-argument part; restated, upstream source absent): at a point x
    with the routed wire values, the sigma values, the Z / partial-product values at x (columns: Z of every challenge, then each challenge's
    partial products) and at g*x.  Returns vanishing_terms[0 .. n_ch*(1 + n_chunks)): the L_0(x)(Z_i - 1) terms, then check_partial_products
    of every challenge."""
    n = 1 << log_n
    n_routed, n_ch = len(wires_row), len(betas)
    n_chunks = -(-n_routed // degree)
    n_pp = n_chunks - 1
    l0 = (pow(x, n, P) - 1) * pow(n * (x - 1) % P, P - 2, P) % P
    z1, pp_terms = [], []
    for i in range(n_ch):
        z_x, z_gx = zs_row[i], zs_next_row[i]
        z1.append(l0 * (z_x - 1) % P)
        nums = [(wires_row[j] + betas[i] * k_is[j] % P * x + gammas[i]) % P for j in range(n_routed)]
        dens = [(wires_row[j] + betas[i] * sigmas_row[j] + gammas[i]) % P for j in range(n_routed)]
        accs = [z_x] + [zs_row[n_ch + i * n_pp + c] for c in range(n_pp)] + [z_gx]
        for c in range(n_chunks):
            pn = pd = 1
            for j in range(c * degree, min((c + 1) * degree, n_routed)):
                pn = pn * nums[j] % P
                pd = pd * dens[j] % P
            pp_terms.append((accs[c] * pn - accs[c + 1] * pd) % P)
    return z1 + pp_terms
