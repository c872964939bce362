"""CPU oracle (Python big-int) for the plonky2 commitment hot path that plonky2.5 drives.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it,
and only as the checker.  The product path (``plonky2.5_b200``) never falls back to this code.

What is restated here
---------------------
The arithmetic of this path lives in a third-party dependency that is NOT vendored under /root/reference:
``plonky2`` / ``plonky2_field`` / ``plonky2_util`` @ git rev 3de92d9ed1721cec133e4e1e1b3ec7facb756ccf
(pinned at /root/reference/Cargo.toml:15-19).  This file restates its published algorithm (SURVEY.md
Appendix A) in the most obvious form possible (naive O(n^2) DFT available, naive dense-MDS Poseidon,
recursive ``fill_subtree``).  Call sites in the reference that drive it: /root/reference/src/p3/mod.rs:250
(``builder.build`` -> commit #0) and :260 (``data.prove`` -> commits #1-#3 + FRI commit phase).

Parity pinning status
---------------------
* Poseidon permutation + its 360 round constants + MDS: PINNED by the four known-answer vectors at
  /root/reference/src/common/poseidon2/poseidon2_goldilocks.rs:190-211 (tests/golden/poseidon_kat.json).
* Field constants: PINNED (/root/reference/src/p3/mod.rs:55, src/p3/extension.rs:149,155,
  src/p3/serde/two_adic.rs:19,35,66).
* Sponge absorb mode: corroborated (not pinned) by the reference's implementation of plonky2's PlonkyPermutation trait
  (/root/reference/src/common/poseidon2/poseidon2.rs:528-565: set_elt / set_from_slice / set_from_iter overwrite lanes, squeeze is
  state[..RATE]); plonky2's generic hash_n_to_hash_no_pad / compress are called through that interface at :575-581.
* LDE order, sponge mode, digest layout, FRI fold: "parity unpinned" — the reference tree holds no golden
  vectors for them and cannot be run here (no Rust toolchain).  They are cross-checked by two independent
  restatements (this file vs oracle/gl_oracle.c) and by oracle-independent algebraic invariants
  (SURVEY.md A.9), see tests/test_oracle_*.py.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

# --------------------------------------------------------------------------------------------------------
# A.1 field  (plonky2 field/src/goldilocks_field.rs; constants visible at /root/reference/src/p3/mod.rs:55,
#             src/p3/extension.rs:149,155, src/p3/serde/two_adic.rs:19,35,66)
# --------------------------------------------------------------------------------------------------------
P = 0xFFFF_FFFF_0000_0001
EPSILON = 0xFFFF_FFFF
MULTIPLICATIVE_GROUP_GENERATOR = 7
POWER_OF_TWO_GENERATOR = 1753635133440165772
TWO_ADICITY = 32
M64 = (1 << 64) - 1


def canon(x: int) -> int:
    return x % P


def inv(x: int) -> int:
    return pow(x, P - 2, P)


def primitive_root_of_unity(n_log: int) -> int:
    """plonky2 field/src/types.rs · Field::primitive_root_of_unity: g2^(2^(32-n_log))."""
    assert n_log <= TWO_ADICITY
    return pow(POWER_OF_TWO_GENERATOR, 1 << (TWO_ADICITY - n_log), P)


def coset_shift() -> int:
    return MULTIPLICATIVE_GROUP_GENERATOR


def reverse_bits(i: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def log2_strict(n: int) -> int:
    assert n > 0 and n & (n - 1) == 0, "not a power of two"
    return n.bit_length() - 1


# --------------------------------------------------------------------------------------------------------
# A.5 Poseidon constants: plonky2/src/bin/generate_constants.rs regenerated offline
#     ChaCha8Rng::seed_from_u64(0) then 360 x gen_range(0..p)   (rand_chacha 0.3 / rand 0.8 semantics)
# --------------------------------------------------------------------------------------------------------
def _pcg32_seed_bytes(state: int, n_bytes: int = 32) -> bytes:
    """rand_core::SeedableRng::seed_from_u64 — PCG32 expansion of a u64 into the seed."""
    MUL, INC = 6364136223846793005, 11634580027462260723
    out = b""
    while len(out) < n_bytes:
        state = (state * MUL + INC) & M64
        xorshifted = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
        rot = state >> 59
        x = ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & 0xFFFFFFFF
        out += x.to_bytes(4, "little")
    return out[:n_bytes]


def _chacha_block(key_words: Sequence[int], counter: int, stream: int, double_rounds: int) -> List[int]:
    def rotl(v, c):
        return ((v << c) | (v >> (32 - c))) & 0xFFFFFFFF

    init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574, *key_words,
            counter & 0xFFFFFFFF, counter >> 32, stream & 0xFFFFFFFF, stream >> 32]
    x = list(init)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = rotl(x[b] ^ x[c], 7)

    for _ in range(double_rounds):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(a + b) & 0xFFFFFFFF for a, b in zip(x, init)]


class ChaCha8Rng:
    """rand_chacha::ChaCha8Rng: 8 rounds, 64-bit block counter from 0, stream 0, u64 = low word then high."""

    def __init__(self, seed_u64: int):
        seed = _pcg32_seed_bytes(seed_u64)
        self.key = [int.from_bytes(seed[4 * i:4 * i + 4], "little") for i in range(8)]
        self.counter = 0
        self.buf: List[int] = []

    def next_u32(self) -> int:
        if not self.buf:
            self.buf = _chacha_block(self.key, self.counter, 0, 4)
            self.counter += 1
        return self.buf.pop(0)

    def next_u64(self) -> int:
        lo = self.next_u32()
        hi = self.next_u32()
        return lo | (hi << 32)

    def gen_range_u64(self, high: int) -> int:
        """rand 0.8 UniformInt<u64>::sample_single(0, high): widening-multiply rejection."""
        lz = 64 - high.bit_length()
        zone = ((high << lz) & M64) - 1
        while True:
            v = self.next_u64()
            m = v * high
            if (m & M64) <= zone:
                return m >> 64


N_ROUND_CONSTANTS = 360
SPONGE_WIDTH = 12
SPONGE_RATE = 8
HALF_N_FULL_ROUNDS = 4
N_PARTIAL_ROUNDS = 22
MDS_CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
MDS_DIAG = [8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]


def generate_round_constants() -> List[int]:
    rng = ChaCha8Rng(0)
    return [rng.gen_range_u64(P) for _ in range(N_ROUND_CONSTANTS)]


ALL_ROUND_CONSTANTS: List[int] = generate_round_constants()
ROUND_CONSTANTS_SHA256 = "d2fcbb5be293c50ab4b1ddcd9c81005b12d689816a54c91a054f97f6588a20a8"


# --------------------------------------------------------------------------------------------------------
# A.5 Poseidon permutation, naive form (plonky2/src/hash/poseidon.rs · Poseidon::poseidon, with
#     full_rounds / partial_rounds_naive / mds_layer / sbox_monomial)
# --------------------------------------------------------------------------------------------------------
def _mds_layer(state: Sequence[int]) -> List[int]:
    out = []
    for r in range(SPONGE_WIDTH):
        acc = 0
        for i in range(SPONGE_WIDTH):
            acc += state[(i + r) % SPONGE_WIDTH] * MDS_CIRC[i]
        acc += state[r] * MDS_DIAG[r]
        out.append(acc % P)
    return out


def poseidon(inp: Sequence[int]) -> List[int]:
    assert len(inp) == SPONGE_WIDTH
    state = [x % P for x in inp]
    rc = 0
    for phase, n_rounds in ((0, HALF_N_FULL_ROUNDS), (1, N_PARTIAL_ROUNDS), (2, HALF_N_FULL_ROUNDS)):
        for _ in range(n_rounds):
            state = [(s + ALL_ROUND_CONSTANTS[SPONGE_WIDTH * rc + i]) % P for i, s in enumerate(state)]
            if phase == 1:
                state[0] = pow(state[0], 7, P)
            else:
                state = [pow(s, 7, P) for s in state]
            state = _mds_layer(state)
            rc += 1
    return state


# --------------------------------------------------------------------------------------------------------
# A.5 hashing (plonky2/src/hash/hashing.rs · hash_n_to_m_no_pad, compress; plonk/config.rs · hash_or_noop)
# --------------------------------------------------------------------------------------------------------
def hash_no_pad(inputs: Sequence[int]) -> List[int]:
    state = [0] * SPONGE_WIDTH
    for off in range(0, len(inputs), SPONGE_RATE):
        chunk = inputs[off:off + SPONGE_RATE]
        state[:len(chunk)] = [c % P for c in chunk]      # overwrite mode, shorter last chunk keeps the rest
        state = poseidon(state)
    return state[:4]


def hash_or_noop(inputs: Sequence[int]) -> List[int]:
    if len(inputs) <= 4:
        return [x % P for x in inputs] + [0] * (4 - len(inputs))
    return hash_no_pad(inputs)


def two_to_one(left: Sequence[int], right: Sequence[int]) -> List[int]:
    return poseidon(list(left) + list(right) + [0, 0, 0, 0])[:4]


# --------------------------------------------------------------------------------------------------------
# A.2 FFT / iFFT   (plonky2 field/src/fft.rs, field/src/polynomial/mod.rs)
# --------------------------------------------------------------------------------------------------------
def fft_naive(coeffs: Sequence[int]) -> List[int]:
    n = len(coeffs)
    w = primitive_root_of_unity(log2_strict(n))
    return [sum(c * pow(w, j * k, P) for j, c in enumerate(coeffs)) % P for k in range(n)]


def fft(coeffs: Sequence[int]) -> List[int]:
    """Radix-2 recursion with the same root choice; identical values to fft_naive (field-exact)."""
    n = len(coeffs)
    if n == 1:
        return [coeffs[0] % P]
    w = primitive_root_of_unity(log2_strict(n))
    ev = fft(coeffs[0::2])
    od = fft(coeffs[1::2])
    out = [0] * n
    t = 1
    for k in range(n // 2):
        u = od[k] * t % P
        out[k] = (ev[k] + u) % P
        out[k + n // 2] = (ev[k] - u) % P
        t = t * w % P
    return out


def ifft(values: Sequence[int]) -> List[int]:
    """ifft_with_options: forward FFT, then out[i] = b[(n-i) mod n] / n."""
    n = len(values)
    b = fft(values)
    n_inv = inv(n)
    return [b[(n - i) % n] * n_inv % P for i in range(n)]


def coset_fft(coeffs: Sequence[int], shift: int) -> List[int]:
    s = 1
    scaled = []
    for c in coeffs:
        scaled.append(c * s % P)
        s = s * shift % P
    return fft(scaled)


# A.3
def lde_values(coeffs: Sequence[int], rate_bits: int) -> List[int]:
    padded = list(coeffs) + [0] * (len(coeffs) * ((1 << rate_bits) - 1))
    return coset_fft(padded, coset_shift())


def eval_poly(coeffs: Sequence[int], x: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % P
    return acc


# --------------------------------------------------------------------------------------------------------
# A.6 Merkle tree (plonky2/src/hash/merkle_tree.rs · MerkleTree::new, fill_digests_buf, fill_subtree, prove;
#                  hash/merkle_proofs.rs · verify_merkle_proof_to_cap)
# --------------------------------------------------------------------------------------------------------
def _fill_subtree(digests: List, base: int, leaves: Sequence[Sequence[int]]) -> List[int]:
    """Recursive layout: [left subtree | left digest | right digest | right subtree]; returns the root."""
    n = len(leaves)
    if n == 1:
        return hash_or_noop(leaves[0])
    half_buf = n - 2          # digests in each child subtree: 2*(n/2 - 1)
    left = _fill_subtree(digests, base, leaves[:n // 2])
    right = _fill_subtree(digests, base + half_buf + 2, leaves[n // 2:])
    digests[base + half_buf] = left
    digests[base + half_buf + 1] = right
    return two_to_one(left, right)


def digest_index(layer: int, j: int) -> int:
    """Closed form: node j of layer `layer` (0 = leaf digests) inside one subtree's digest buffer."""
    return 2 * (((j >> 1) << (layer + 1)) + (1 << layer) - 1) + (j & 1)


class MerkleTree:
    def __init__(self, leaves: Sequence[Sequence[int]], cap_height: int):
        n = len(leaves)
        log_n = log2_strict(n)
        assert cap_height <= log_n, "cap_height should be at most log2(leaves.len())"
        self.leaves = [list(l) for l in leaves]
        self.cap_height = cap_height
        n_sub = 1 << cap_height
        sub = n // n_sub
        per = 2 * (sub - 1)
        self.digests: List = [None] * (2 * (n - n_sub))
        self.cap: List[List[int]] = []
        for t in range(n_sub):
            self.cap.append(_fill_subtree(self.digests, t * per, self.leaves[t * sub:(t + 1) * sub]))

    def get(self, i: int) -> List[int]:
        return self.leaves[i]

    def prove(self, leaf_index: int) -> List[List[int]]:
        n = len(self.leaves)
        n_sub = 1 << self.cap_height
        sub = n // n_sub
        per = 2 * (sub - 1)
        t, j = divmod(leaf_index, sub)
        siblings = []
        for layer in range(log2_strict(sub)):
            siblings.append(self.digests[t * per + digest_index(layer, j ^ 1)])
            j >>= 1
        return siblings


def verify_merkle_proof_to_cap(leaf: Sequence[int], leaf_index: int, cap: Sequence[Sequence[int]],
                               siblings: Sequence[Sequence[int]]) -> bool:
    cur = hash_or_noop(leaf)
    idx = leaf_index
    for sib in siblings:
        cur = two_to_one(sib, cur) if idx & 1 else two_to_one(cur, sib)
        idx >>= 1
    return [c % P for c in cur] == [c % P for c in cap[idx]]


# --------------------------------------------------------------------------------------------------------
# A.3/A.4 PolynomialBatch (plonky2/src/fri/oracle.rs · from_values / from_coeffs / lde_values)
# --------------------------------------------------------------------------------------------------------
class PolynomialBatch:
    def __init__(self, polynomials, merkle_tree, degree_log, rate_bits):
        self.polynomials = polynomials
        self.merkle_tree = merkle_tree
        self.degree_log = degree_log
        self.rate_bits = rate_bits
        self.blinding = False

    @classmethod
    def from_values(cls, values: Sequence[Sequence[int]], rate_bits: int, cap_height: int) -> "PolynomialBatch":
        return cls.from_coeffs([ifft(v) for v in values], rate_bits, cap_height)

    @classmethod
    def from_coeffs(cls, polynomials: Sequence[Sequence[int]], rate_bits: int, cap_height: int) -> "PolynomialBatch":
        degree = len(polynomials[0])
        assert all(len(p) == degree for p in polynomials), "Polynomial degrees inconsistent"
        lde = [lde_values(p, rate_bits) for p in polynomials]
        rows = degree << rate_bits
        bits = log2_strict(rows)
        # transpose + reverse_index_bits_in_place
        leaves = [[col[reverse_bits(i, bits)] for col in lde] for i in range(rows)]
        tree = MerkleTree(leaves, cap_height)
        return cls([list(p) for p in polynomials], tree, log2_strict(degree), rate_bits)

    def get_lde_values(self, index: int, step: int) -> List[int]:
        rows = len(self.merkle_tree.leaves)
        return self.merkle_tree.leaves[reverse_bits(index * step, log2_strict(rows))]


# --------------------------------------------------------------------------------------------------------
# A.7 quadratic extension F_p[X]/(X^2-7) (plonky2 field/src/extension/quadratic.rs; same W=7 as
#     /root/reference/src/p3/extension.rs:147-151,458-470)
# --------------------------------------------------------------------------------------------------------
EXT_W = 7
Ext = Tuple[int, int]


def ext_add(a: Ext, b: Ext) -> Ext:
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def ext_sub(a: Ext, b: Ext) -> Ext:
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def ext_mul(a: Ext, b: Ext) -> Ext:
    return ((a[0] * b[0] + EXT_W * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def ext_scale(a: Ext, s: int) -> Ext:
    return (a[0] * s % P, a[1] * s % P)


def ext_fft(coeffs: Sequence[Ext]) -> List[Ext]:
    """FFT over the base-field subgroup applied to extension coefficients = component-wise base FFT."""
    c0 = fft([c[0] for c in coeffs])
    c1 = fft([c[1] for c in coeffs])
    return list(zip(c0, c1))


def ext_coset_fft(coeffs: Sequence[Ext], shift: int) -> List[Ext]:
    c0 = coset_fft([c[0] for c in coeffs], shift)
    c1 = coset_fft([c[1] for c in coeffs], shift)
    return list(zip(c0, c1))


def ext_eval_poly(coeffs: Sequence[Ext], x: int) -> Ext:
    acc = (0, 0)
    for c in reversed(coeffs):
        acc = ext_add(ext_scale(acc, x), c)
    return acc


# --------------------------------------------------------------------------------------------------------
# A.8 Challenger (plonky2/src/iop/challenger.rs) — duplex sponge over Poseidon
# --------------------------------------------------------------------------------------------------------
class Challenger:
    def __init__(self):
        self.sponge_state = [0] * SPONGE_WIDTH
        self.input_buffer: List[int] = []
        self.output_buffer: List[int] = []

    def observe_element(self, e: int):
        self.output_buffer = []
        self.input_buffer.append(e % P)
        if len(self.input_buffer) == SPONGE_RATE:
            self._duplexing()

    def observe_elements(self, es: Sequence[int]):
        for e in es:
            self.observe_element(e)

    def observe_hash(self, h: Sequence[int]):
        self.observe_elements(h)

    def observe_cap(self, cap: Sequence[Sequence[int]]):
        for h in cap:
            self.observe_hash(h)

    def observe_extension_element(self, e: Ext):
        self.observe_elements(list(e))

    def get_challenge(self) -> int:
        if self.input_buffer or not self.output_buffer:
            self._duplexing()
        return self.output_buffer.pop()

    def get_extension_challenge(self) -> Ext:
        c0 = self.get_challenge()
        c1 = self.get_challenge()
        return (c0, c1)

    def _duplexing(self):
        assert len(self.input_buffer) <= SPONGE_RATE
        for i, v in enumerate(self.input_buffer):
            self.sponge_state[i] = v
        self.input_buffer = []
        self.sponge_state = poseidon(self.sponge_state)
        self.output_buffer = list(self.sponge_state[:SPONGE_RATE])


def fri_proof_of_work(challenger: "Challenger", min_leading_zeros: int) -> int:
    """plonky2 fri/prover.rs · fri_proof_of_work, smallest witness (serial `find`, as the reference ships
    plonky2 with `parallel` off).  Advances the challenger exactly as upstream (observe witness, draw the response)."""
    base = list(challenger.sponge_state)
    for i, v in enumerate(challenger.input_buffer):
        base[i] = v
    pos = len(challenger.input_buffer)
    assert pos < SPONGE_RATE
    w = 0
    while True:
        st = list(base)
        st[pos] = w
        r = poseidon(st)[SPONGE_RATE - 1]
        if 64 - r.bit_length() >= min_leading_zeros:
            break
        w += 1
    challenger.observe_element(w)
    resp = challenger.get_challenge()
    assert resp == r
    return w


# --------------------------------------------------------------------------------------------------------
# A.7 FRI commit phase (plonky2/src/fri/prover.rs · fri_committed_trees; fri/reduction_strategies.rs;
#     plonk/plonk_common.rs · reduce_with_powers)
# --------------------------------------------------------------------------------------------------------
def reduction_arity_bits_constant(arity_bits: int, final_poly_bits: int, degree_bits: int,
                                  rate_bits: int, cap_height: int) -> List[int]:
    """FriReductionStrategy::ConstantArityBits(arity_bits, final_poly_bits).reduction_arity_bits."""
    out = []
    while degree_bits > final_poly_bits and degree_bits + rate_bits - arity_bits >= cap_height:
        out.append(arity_bits)
        assert degree_bits >= arity_bits
        degree_bits -= arity_bits
    return out


def fri_committed_trees(coeffs: List[Ext], values: List[Ext], challenger: Challenger,
                        reduction_arity_bits: Sequence[int], rate_bits: int, cap_height: int):
    """Returns (trees, final_poly_coeffs). `coeffs` has len R (upper part zero), `values` = coset_fft(coeffs, 7)."""
    trees = []
    shift = MULTIPLICATIVE_GROUP_GENERATOR
    coeffs = list(coeffs)
    values = list(values)
    for arity_bits in reduction_arity_bits:
        arity = 1 << arity_bits
        bits = log2_strict(len(values))
        values = [values[reverse_bits(i, bits)] for i in range(len(values))]
        leaves = []
        for off in range(0, len(values), arity):
            flat: List[int] = []
            for e in values[off:off + arity]:
                flat.extend(e)
            leaves.append(flat)
        tree = MerkleTree(leaves, cap_height)
        challenger.observe_cap(tree.cap)
        trees.append(tree)
        beta = challenger.get_extension_challenge()
        folded = []
        for off in range(0, len(coeffs), arity):
            acc = (0, 0)
            for c in reversed(coeffs[off:off + arity]):      # reduce_with_powers: sum c_t * beta^t
                acc = ext_add(ext_mul(acc, beta), c)
            folded.append(acc)
        coeffs = folded
        shift = pow(shift, arity, P)
        values = ext_coset_fft(coeffs, shift)
    final = coeffs[:len(coeffs) >> rate_bits]
    for c in final:
        challenger.observe_extension_element(c)
    return trees, final


# --------------------------------------------------------------------------------------------------------
# prove_openings, front half (plonky2/src/fri/oracle.rs · PolynomialBatch::prove_openings; util/reducing.rs ·
# ReducingFactor; field/src/polynomial/division.rs · divide_by_linear) — restated from memory of upstream @ 3de92d9
# (parity unpinned: no vectors in /root/reference; tests check the algebraic identities instead)
# --------------------------------------------------------------------------------------------------------
def ext_pow(a: Ext, e: int) -> Ext:
    r = (1, 0)
    while e:
        if e & 1:
            r = ext_mul(r, a)
        a = ext_mul(a, a)
        e >>= 1
    return r


class ReducingFactor:
    def __init__(self, base: Ext):
        self.base = base
        self.count = 0

    def reduce_polys_base(self, polys: Sequence[Sequence[int]]) -> List[Ext]:
        """sum_j base^j * poly_j (powers restart at base^0 on every call; count += number of polynomials)"""
        n = max((len(p) for p in polys), default=0)
        acc = [(0, 0)] * n
        pw = (1, 0)
        for p in polys:
            self.count += 1
            acc = [ext_add(a, ext_scale(pw, c)) for a, c in zip(acc, list(p) + [0] * (n - len(p)))]
            pw = ext_mul(pw, self.base)
        return acc

    def shift_poly(self, p: List[Ext]) -> List[Ext]:
        s = ext_pow(self.base, self.count)
        self.count = 0
        return [ext_mul(c, s) for c in p]


def ext_divide_by_linear(coeffs: Sequence[Ext], z: Ext) -> List[Ext]:
    """(p(X) - p(z)) / (X - z): Horner scan from the top coefficient, remainder dropped (one coefficient shorter)."""
    bs = []
    acc = (0, 0)
    for c in reversed(coeffs):
        acc = ext_add(ext_mul(acc, z), c)
        bs.append(acc)
    bs.pop()
    bs.reverse()
    return bs


def ext_eval_poly_ext(coeffs: Sequence[Ext], x: Ext) -> Ext:
    acc = (0, 0)
    for c in reversed(coeffs):
        acc = ext_add(ext_mul(acc, x), c)
    return acc


def prove_openings_final_poly(batches, oracles, alpha: Ext):
    """batches: [(point: Ext, [(oracle_index, polynomial_index), ...])]; oracles: lists of coefficient vectors per oracle.
    Returns (final_poly coefficients [N], the per-batch quotients).
    Memory note (parity unpinned): early-2022 plonky2 multiplied final_poly by X here (`coeffs.insert(0, ZERO)`, PR #436) and its
    verifier multiplied fri_combine_initial's sum by subgroup_x; the batch-structured code this restates pads each quotient back to a
    power of two instead (`quotient.coeffs.push(ZERO)`) and has no such factor on either side.  oracle/fri_verifier.py mirrors that
    choice, so prover and verifier restatements agree with each other; if the pinned commit did carry the factor, both (and
    csrc/openings.cuh) would need the same one-line shift."""
    rf = ReducingFactor(alpha)
    final: List[Ext] = []
    quotients = []
    for point, polys in batches:
        comp = rf.reduce_polys_base([oracles[o][i] for (o, i) in polys])
        q = ext_divide_by_linear(comp, point)
        q.append((0, 0))                                  # pad back to a power of two
        quotients.append(q)
        final = rf.shift_poly(final)
        if not final:
            final = [(0, 0)] * len(q)
        final = [ext_add(a, b) for a, b in zip(final, q)]
    return final, quotients


def prove_openings_lde(final_poly: Sequence[Ext], rate_bits: int):
    """lde_final_poly = final_poly.lde(rate_bits); lde_final_values = lde_final_poly.coset_fft(7)"""
    lde = list(final_poly) + [(0, 0)] * (len(final_poly) * ((1 << rate_bits) - 1))
    return lde, ext_coset_fft(lde, MULTIPLICATIVE_GROUP_GENERATOR)


def fri_prover_query_rounds(initial_trees: Sequence[MerkleTree], trees: Sequence[MerkleTree], challenger: Challenger,
                            n_query_rounds: int, reduction_arity_bits: Sequence[int], lde_size: int):
    """plonky2/src/fri/prover.rs · fri_prover_query_rounds / fri_prover_query_round (restated from memory of upstream @ 3de92d9):
    per step `evals = unflatten(tree.get(x_index >> arity_bits))` — the whole coset, `arity` extension elements; the verifier
    (fri/verifier.rs · fri_verifier_query_round) reads evals[x_index & (arity-1)] and Merkle-verifies flatten(evals), and
    validate_shape requires evals.len() == arity."""
    rounds = []
    for _ in range(n_query_rounds):
        x_index = challenger.get_challenge() % lde_size
        rnd = {"x_index": x_index, "initial_trees_proof": [(t.get(x_index), t.prove(x_index)) for t in initial_trees], "steps": []}
        x = x_index
        for arity_bits, tree in zip(reduction_arity_bits, trees):
            arity = 1 << arity_bits
            flat = tree.get(x >> arity_bits)
            evals = [(flat[2 * i], flat[2 * i + 1]) for i in range(arity)]   # unflatten(tree.get(..)): ALL arity elements; dropping the
            # queried one is FriProof::compress's job (fri/proof.rs), and decompress re-inserts it before the verifier runs
            rnd["steps"].append({"evals": evals, "merkle_proof": tree.prove(x >> arity_bits)})
            x >>= arity_bits
        rounds.append(rnd)
    return rounds
