"""ctypes wrapper around oracle/libgl_oracle.so (the C restatement of the reference algorithm) — test harness only."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
P = 0xFFFF_FFFF_0000_0001
u64 = ctypes.c_uint64
vp = ctypes.c_void_p


def build_oracle():
    lib = os.path.join(ORACLE_DIR, "libgl_oracle.so")
    src = os.path.join(ORACLE_DIR, "gl_oracle.c")
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, env=env, stdout=subprocess.DEVNULL)
    return lib


def splitmix_columns(seed: int, n_cols: int, n: int, canonical: bool = True) -> np.ndarray:
    """SplitMix64 stream (seed 0x706c6f6e6b7932 ^ seed), reduced mod p (or raw u64 when canonical=False): [n_cols][n]."""
    total = n_cols * n
    idx = np.arange(1, total + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(0x706C6F6E6B7932 ^ seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    if canonical:
        z = np.where(z >= np.uint64(P), z - np.uint64(P), z)
    return z.reshape(n_cols, n)


class OracleC:
    def __init__(self):
        self.lib = ctypes.CDLL(build_oracle())
        L = self.lib
        L.glo_commit.restype = ctypes.c_int
        L.glo_commit.argtypes = [ctypes.POINTER(vp), ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                 ctypes.c_int, vp, vp, vp, vp, ctypes.POINTER(ctypes.c_double)]
        L.glo_merkle_new.restype = ctypes.c_int
        L.glo_merkle_new.argtypes = [vp, u64, u64, ctypes.c_uint, vp, vp]
        L.glo_poseidon.argtypes = [vp]
        L.glo_hash_or_noop.argtypes = [vp, u64, vp]
        L.glo_two_to_one.argtypes = [vp, vp, vp]
        L.glo_fft.argtypes = [vp, ctypes.c_uint]
        L.glo_ifft.argtypes = [vp, ctypes.c_uint]
        L.glo_coset_fft.argtypes = [vp, ctypes.c_uint, u64]
        L.glo_challenger_sizeof.restype = ctypes.c_uint
        L.glo_challenger_init.argtypes = [vp]
        L.glo_challenger_observe.argtypes = [vp, vp, u64]
        L.glo_challenger_get.restype = u64
        L.glo_challenger_get.argtypes = [vp]
        L.glo_fri_committed_trees.restype = ctypes.c_int
        L.glo_fri_committed_trees.argtypes = [vp, vp, u64, vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, vp,
                                              ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp), vp, vp]
        L.glo_fri_proof_of_work.restype = u64
        L.glo_fri_proof_of_work.argtypes = [vp, ctypes.c_uint32]
        L.glo_num_threads.restype = ctypes.c_int
        L.glo_root_of_unity.restype = u64
        L.glo_root_of_unity.argtypes = [ctypes.c_uint]
        L.glo_mul.restype = u64
        L.glo_mul.argtypes = [u64, u64]
        L.glo_openings_add_batch.restype = ctypes.c_int
        L.glo_openings_add_batch.argtypes = [ctypes.POINTER(vp), ctypes.c_uint32, u64, vp, vp, vp, vp]

    def set_threads(self, n):
        self.lib.glo_set_num_threads(int(n))

    def num_threads(self):
        return self.lib.glo_num_threads()

    def poseidon(self, state):
        s = np.array(state, dtype=np.uint64).copy()
        self.lib.glo_poseidon(s.ctypes.data)
        return s

    def hash_or_noop(self, row):
        r = np.ascontiguousarray(row, dtype=np.uint64)
        out = np.zeros(4, dtype=np.uint64)
        self.lib.glo_hash_or_noop(r.ctypes.data, len(r), out.ctypes.data)
        return out

    def two_to_one(self, l, r):
        l = np.ascontiguousarray(l, dtype=np.uint64); r = np.ascontiguousarray(r, dtype=np.uint64)
        out = np.zeros(4, dtype=np.uint64)
        self.lib.glo_two_to_one(l.ctypes.data, r.ctypes.data, out.ctypes.data)
        return out

    def fft(self, a):
        a = np.array(a, dtype=np.uint64).copy()
        self.lib.glo_fft(a.ctypes.data, int(len(a)).bit_length() - 1)
        return a

    def ifft(self, a):
        a = np.array(a, dtype=np.uint64).copy()
        self.lib.glo_ifft(a.ctypes.data, int(len(a)).bit_length() - 1)
        return a

    def coset_fft(self, a, shift):
        a = np.array(a, dtype=np.uint64).copy()
        self.lib.glo_coset_fft(a.ctypes.data, int(len(a)).bit_length() - 1, shift)
        return a

    def commit(self, cols, rate_bits, cap_height, is_coeffs=False, want=("coeffs", "leaves", "digests")):
        cols = [np.ascontiguousarray(c, dtype=np.uint64) for c in cols]
        n_cols, n = len(cols), len(cols[0])
        log_n = n.bit_length() - 1
        rows = n << rate_bits
        ptrs = (vp * n_cols)(*[c.ctypes.data for c in cols])
        cap = np.zeros(((1 << cap_height), 4), dtype=np.uint64)
        oc = np.zeros((n_cols, n), dtype=np.uint64) if "coeffs" in want else None
        ol = np.zeros((rows, n_cols), dtype=np.uint64) if "leaves" in want else None
        od = np.zeros((max(2 * (rows - (1 << cap_height)), 0), 4), dtype=np.uint64) if "digests" in want else None
        st = (ctypes.c_double * 4)()
        p = lambda a: None if a is None else a.ctypes.data
        rc = self.lib.glo_commit(ptrs, n_cols, log_n, rate_bits, cap_height, int(is_coeffs), p(oc), p(ol), p(od), cap.ctypes.data, st)
        if rc != 0:
            raise ValueError("oracle commit rejected the shape")
        return {"coeffs": oc, "leaves": ol, "digests": od, "cap": cap, "stage_s": list(st)}

    def merkle_new(self, leaves, cap_height):
        lv = np.ascontiguousarray(leaves, dtype=np.uint64)
        n, ll = lv.shape
        if n < (1 << cap_height):
            raise ValueError("cap_height should be at most log2(leaves.len())")
        dig = np.zeros((2 * (n - (1 << cap_height)), 4), dtype=np.uint64)
        cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
        rc = self.lib.glo_merkle_new(lv.ctypes.data, n, ll, cap_height, dig.ctypes.data if dig.size else None, cap.ctypes.data)
        if rc != 0:
            raise ValueError("oracle merkle_new rejected the shape")
        return dig, cap

    def verify_path(self, leaf, index, siblings, cap):
        """Verifier rule (plonky2 hash/merkle_proofs.rs · verify_merkle_proof_to_cap)."""
        cur = self.hash_or_noop(leaf)
        for sib in siblings:
            cur = self.two_to_one(sib, cur) if index & 1 else self.two_to_one(cur, sib)
            index >>= 1
        return bool(np.array_equal(cur, np.asarray(cap).reshape(-1, 4)[index]))

    def new_challenger(self):
        return ChallengerC(self)

    def openings_final_poly(self, batches, oracles, alpha):
        """batches: [(point (a0, a1), [(oracle_index, polynomial_index), ...])]; oracles: [n_cols][N] uint64 arrays.
        Returns (final_poly [N][2], [quotient [N][2] per batch]) — prove_openings front half."""
        n = np.asarray(oracles[0]).shape[1]
        final = np.zeros((n, 2), dtype=np.uint64)
        quotients = []
        al = np.array(alpha, dtype=np.uint64)
        for point, polys in batches:
            cols = [np.ascontiguousarray(oracles[o][i], dtype=np.uint64) for (o, i) in polys]
            ptrs = (vp * max(len(cols), 1))(*[c.ctypes.data for c in cols])
            q = np.zeros((n, 2), dtype=np.uint64)
            pt = np.array(point, dtype=np.uint64)
            rc = self.lib.glo_openings_add_batch(ptrs, len(cols), n, al.ctypes.data, pt.ctypes.data, final.ctypes.data, q.ctypes.data)
            assert rc == 0
            quotients.append(q)
        return final, quotients

    def fri_committed_trees(self, coeffs, values, arity_bits, rate_bits, cap_height, challenger):
        co = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 2)
        va = np.ascontiguousarray(values, dtype=np.uint64).reshape(-1, 2)
        ln = co.shape[0]
        ab = np.array(arity_bits, dtype=np.uint32)
        leaves, digs, caps = [], [], []
        cur = ln
        for a in arity_bits:
            nl = cur >> a
            leaves.append(np.zeros((nl, 2 << a), dtype=np.uint64))
            digs.append(np.zeros((max(2 * (nl - (1 << cap_height)), 1), 4), dtype=np.uint64))
            caps.append(np.zeros((1 << cap_height, 4), dtype=np.uint64))
            cur = nl
        n_final = cur >> rate_bits
        final = np.zeros((max(n_final, 1), 2), dtype=np.uint64)
        betas = np.zeros((max(len(arity_bits), 1), 2), dtype=np.uint64)
        mk = lambda arrs: (vp * max(len(arrs), 1))(*[a.ctypes.data for a in arrs])
        rc = self.lib.glo_fri_committed_trees(co.ctypes.data, va.ctypes.data, ln, ab.ctypes.data, len(arity_bits), rate_bits,
                                              cap_height, challenger.buf, mk(leaves), mk(digs), mk(caps), betas.ctypes.data,
                                              final.ctypes.data)
        if rc != 0:
            raise ValueError("oracle fri_committed_trees rejected the shape")
        return {"leaves": leaves, "digests": digs, "caps": caps, "betas": betas[:len(arity_bits)], "final_poly": final[:n_final]}


class ChallengerC:
    """oracle Challenger (plonky2 iop/challenger.rs) — duck-type compatible with the product's fri_committed_trees."""

    def __init__(self, oc: OracleC):
        self.oc = oc
        self._mem = ctypes.create_string_buffer(oc.lib.glo_challenger_sizeof())
        self.buf = ctypes.addressof(self._mem)
        oc.lib.glo_challenger_init(self.buf)

    def observe_elements(self, es):
        a = np.ascontiguousarray(es, dtype=np.uint64).reshape(-1)
        self.oc.lib.glo_challenger_observe(self.buf, a.ctypes.data, a.size)

    def observe_cap(self, cap):
        self.observe_elements(cap)

    def observe_extension_elements(self, es):
        self.observe_elements(es)

    def get_challenge(self):
        return int(self.oc.lib.glo_challenger_get(self.buf))

    def get_extension_challenge(self):
        c0 = self.get_challenge()
        c1 = self.get_challenge()
        return (c0, c1)

    def fri_proof_of_work(self, min_leading_zeros):
        return int(self.oc.lib.glo_fri_proof_of_work(self.buf, min_leading_zeros))
