// §8(f) ranks 3-4: gate-constraint evaluation over the resident LDE rows, and witness generation, for the reference's own gates.
//
//   Poseidon2Gate           /root/reference/src/common/poseidon2/poseidon2_gate.rs:233-310 (eval_unfiltered_base_one: 135 wires ->
//                           123 constraints, degree 7) and :447-523 (Poseidon2Generator::run_once); the permutation pieces are
//                           /root/reference/src/common/poseidon2/poseidon2.rs:127-245 (matmul_external / matmul_m4 / matmul_internal)
//   U32ArithmeticGate       /root/reference/src/common/u32/gates/arithmetic_u32.rs:103-166 (num_ops x (4 + 32) constraints, degree 4)
//
// This is what plonky2 plonk/vanishing_poly.rs · evaluate_gate_constraints_base_batch does per gate inside compute_quotient_polys:
// one thread per LDE row (a leaf row of the wires commit, already in HBM), every constraint value multiplied by alpha_k^(offset + i)
// and accumulated per challenge k (plonk_common.rs · reduce_with_powers), optionally times a per-row filter value.
// All arithmetic is canonical Goldilocks (gl::add / gl::sub / gl::mulc): results are bit-exact field values.
#pragma once
#include "gl_field.cuh"
#include "poseidon2_constants.cuh"

namespace gates {

constexpr int P2_WIDTH = 12, P2_RF_BEGIN = 4, P2_RF_END = 8, P2_RP = 22;
constexpr int P2_WIRE_SWAP = 24, P2_START_DELTA = 25, P2_START_RF_BEGIN = 29, P2_START_PARTIAL = 29 + 12 * 3, P2_START_RF_END = 29 + 36 + 22;
constexpr int P2_NUM_WIRES = P2_START_RF_END + 12 * 4;                 // 135
constexpr int P2_NUM_CONSTRAINTS = 12 * 7 + 22 + 12 + 1 + 4;           // 123
constexpr int U32_LIMBS = 32, U32_ROUTED = 6;
constexpr int MAX_CONSTRAINTS = 160, MAX_CHALLENGES = 4;

__device__ __forceinline__ uint64_t dbl(uint64_t x) { return gl::add(x, x); }

__device__ __forceinline__ uint64_t sbox7(uint64_t x) {
    const uint64_t x2 = gl::sqr(x), x4 = gl::sqr(x2), x3 = gl::mul(x, x2);
    return gl::mulc(x3, x4);
}

// poseidon2.rs:185-245
__device__ __forceinline__ void matmul_m4(uint64_t (&s)[P2_WIDTH]) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int a = 4 * i;
        const uint64_t t0 = gl::add(s[a], s[a + 1]), t1 = gl::add(s[a + 2], s[a + 3]);
        const uint64_t t2 = gl::add(t1, dbl(s[a + 1])), t3 = gl::add(t0, dbl(s[a + 3]));
        const uint64_t t4 = gl::add(t3, dbl(dbl(t1))), t5 = gl::add(t2, dbl(dbl(t0)));
        s[a] = gl::add(t3, t5); s[a + 1] = t5; s[a + 2] = gl::add(t2, t4); s[a + 3] = t4;
    }
}
// poseidon2.rs:127-146
__device__ __forceinline__ void matmul_external(uint64_t (&s)[P2_WIDTH]) {
    matmul_m4(s);
    uint64_t st[4];
#pragma unroll
    for (int l = 0; l < 4; l++) st[l] = gl::add(gl::add(s[l], s[4 + l]), s[8 + l]);
#pragma unroll
    for (int i = 0; i < P2_WIDTH; i++) s[i] = gl::add(s[i], st[i & 3]);
}
// poseidon2.rs:164-182: s_i <- (MAT_DIAG_M_1[i] - 1) * s_i + sum(s)
__device__ __forceinline__ void matmul_internal(uint64_t (&s)[P2_WIDTH]) {
    uint64_t sum = s[0];
#pragma unroll
    for (int i = 1; i < P2_WIDTH; i++) sum = gl::add(sum, s[i]);
#pragma unroll
    for (int i = 0; i < P2_WIDTH; i++) s[i] = gl::add(gl::mulc(s[i], poseidon2::DIAG_M_2[i]), sum);
}

// The constraint stream of one Poseidon2Gate row, in the reference's order; `w(i)` reads wire i (canonical), `emit(v)` takes constraint values.
template <class Wires, class Emit>
__device__ __forceinline__ void poseidon2_gate(Wires&& w, Emit&& emit) {
    const uint64_t swap = w(P2_WIRE_SWAP);
    emit(gl::mulc(swap, gl::sub(swap, 1)));
    uint64_t s[P2_WIDTH];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint64_t lhs = w(i), rhs = w(i + 4), d = w(P2_START_DELTA + i);
        emit(gl::sub(gl::mulc(swap, gl::sub(rhs, lhs)), d));
        s[i] = gl::add(lhs, d);
        s[i + 4] = gl::sub(rhs, d);
    }
#pragma unroll
    for (int i = 8; i < P2_WIDTH; i++) s[i] = w(i);
    matmul_external(s);
#pragma unroll 1
    for (int r = 0; r < P2_RF_BEGIN; r++) {
#pragma unroll
        for (int i = 0; i < P2_WIDTH; i++) {
            uint64_t v = gl::add(s[i], poseidon2::RC[r][i]);
            if (r != 0) {
                const uint64_t in = w(P2_START_RF_BEGIN + P2_WIDTH * (r - 1) + i);
                emit(gl::sub(v, in));
                v = in;
            }
            s[i] = sbox7(v);
        }
        matmul_external(s);
    }
#pragma unroll 1
    for (int r = 0; r < P2_RP; r++) {
        const uint64_t v = gl::add(s[0], poseidon2::RC_MID[r]);
        const uint64_t in = w(P2_START_PARTIAL + r);
        emit(gl::sub(v, in));
        s[0] = sbox7(in);
        matmul_internal(s);
    }
#pragma unroll 1
    for (int r = P2_RF_BEGIN; r < P2_RF_END; r++) {
#pragma unroll
        for (int i = 0; i < P2_WIDTH; i++) {
            const uint64_t v = gl::add(s[i], poseidon2::RC[r][i]);
            const uint64_t in = w(P2_START_RF_END + P2_WIDTH * (r - P2_RF_BEGIN) + i);
            emit(gl::sub(v, in));
            s[i] = sbox7(in);
        }
        matmul_external(s);
    }
#pragma unroll
    for (int i = 0; i < P2_WIDTH; i++) emit(gl::sub(s[i], w(P2_WIDTH + i)));
}

// U32ArithmeticGate, eval_unfiltered order: per op [hi_not_max_or_lo_zero, combined - computed, 32 limb range products (j = 31..0), low, high]
template <class Wires, class Emit>
__device__ __forceinline__ void u32_arithmetic_gate(Wires&& w, Emit&& emit, uint32_t num_ops) {
    for (uint32_t i = 0; i < num_ops; i++) {
        const uint64_t m0 = w(6 * i), m1 = w(6 * i + 1), addend = w(6 * i + 2), lo = w(6 * i + 3), hi = w(6 * i + 4), inv = w(6 * i + 5);
        const uint64_t computed = gl::add(gl::mulc(m0, m1), addend);
        const uint64_t diff = gl::sub(0xFFFFFFFFULL, hi);
        const uint64_t hi_not_max = gl::sub(gl::mulc(inv, diff), 1);
        emit(gl::mulc(hi_not_max, lo));
        emit(gl::sub(gl::add(gl::mulc(hi, 1ULL << 32), lo), computed));
        uint64_t clo = 0, chi = 0;
#pragma unroll 1
        for (int j = U32_LIMBS - 1; j >= 0; j--) {
            const uint64_t limb = w(U32_ROUTED * num_ops + U32_LIMBS * i + j);
            uint64_t prod = limb;                                   // (limb - 0)
            prod = gl::mulc(prod, gl::sub(limb, 1));
            prod = gl::mulc(prod, gl::sub(limb, 2));
            prod = gl::mulc(prod, gl::sub(limb, 3));
            emit(prod);
            if (j < U32_LIMBS / 2) clo = gl::add(dbl(dbl(clo)), limb);
            else chi = gl::add(dbl(dbl(chi)), limb);
        }
        emit(gl::sub(clo, lo));
        emit(gl::sub(chi, hi));
    }
}

// ---- the reference's other u32 / b32 gates (/root/reference/src/common/u32/gates/*.rs · eval_unfiltered, same constraint order) -------
__device__ __forceinline__ uint64_t range4(uint64_t v) {   // v (v-1) (v-2) (v-3)
    return gl::mulc(gl::mulc(gl::mulc(v, gl::sub(v, 1)), gl::sub(v, 2)), gl::sub(v, 3));
}
__device__ __forceinline__ uint64_t range2(uint64_t v) { return gl::mulc(v, gl::sub(v, 1)); }
__device__ __forceinline__ uint64_t times4(uint64_t v) { return dbl(dbl(v)); }

// add_many_u32.rs:103-143; param = num_addends | num_ops << 8; 16 result limbs + 2 carry limbs of 2 bits
template <class Wires, class Emit>
__device__ __forceinline__ void add_many_gate(Wires&& w, Emit&& emit, uint32_t param) {
    const uint32_t na = param & 0xFF, num_ops = param >> 8, stride = na + 3;
    for (uint32_t i = 0; i < num_ops; i++) {
        uint64_t sum = w(stride * i + na);                                    // carry
        for (uint32_t j = 0; j < na; j++) sum = gl::add(sum, w(stride * i + j));
        const uint64_t res = w(stride * i + na + 1), car = w(stride * i + na + 2);
        emit(gl::sub(gl::add(gl::mulc(car, 1ULL << 32), res), sum));
        uint64_t rl = 0, cl = 0;
#pragma unroll 1
        for (int j = 17; j >= 0; j--) {
            const uint64_t limb = w(stride * num_ops + 18 * i + j);
            emit(range4(limb));
            if (j < 16) rl = gl::add(times4(rl), limb);
            else cl = gl::add(times4(cl), limb);
        }
        emit(gl::sub(rl, res));
        emit(gl::sub(cl, car));
    }
}
// subtraction_u32.rs:100-134; param = num_ops
template <class Wires, class Emit>
__device__ __forceinline__ void subtraction_gate(Wires&& w, Emit&& emit, uint32_t num_ops) {
    for (uint32_t i = 0; i < num_ops; i++) {
        const uint64_t x = w(5 * i), y = w(5 * i + 1), b = w(5 * i + 2), res = w(5 * i + 3), bo = w(5 * i + 4);
        const uint64_t initial = gl::sub(gl::sub(x, y), b);
        emit(gl::sub(res, gl::add(initial, gl::mulc(bo, 1ULL << 32))));
        uint64_t comb = 0;
#pragma unroll 1
        for (int j = 15; j >= 0; j--) {
            const uint64_t limb = w(5 * num_ops + 16 * i + j);
            emit(range4(limb));
            comb = gl::add(times4(comb), limb);
        }
        emit(gl::sub(comb, res));
        emit(gl::mulc(bo, gl::sub(1, bo)));
    }
}
// range_check_u32.rs:70-92; param = num_input_limbs
template <class Wires, class Emit>
__device__ __forceinline__ void range_check_gate(Wires&& w, Emit&& emit, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) {
        uint64_t sum = 0;
        for (int j = 15; j >= 0; j--) sum = gl::add(times4(sum), w(n + 16 * i + j));   // reduce_with_powers(aux, 4)
        emit(gl::sub(sum, w(i)));
#pragma unroll 1
        for (int j = 0; j < 16; j++) emit(range4(w(n + 16 * i + j)));
    }
}
// interleave_u32.rs:104-140; param = num_ops; 32 big-endian bits per op
template <class Wires, class Emit>
__device__ __forceinline__ void interleave_gate(Wires&& w, Emit&& emit, uint32_t num_ops) {
    for (uint32_t i = 0; i < num_ops; i++) {
        const uint32_t b0 = 2 * num_ops + 32 * i;
        uint64_t v2 = 0, v4 = 0;
        for (int k = 0; k < 32; k++) {                                        // bits[0] is the most significant
            const uint64_t bit = w(b0 + k);
            v2 = gl::add(dbl(v2), bit);
            v4 = gl::add(times4(v4), bit);
        }
        emit(gl::sub(v2, w(2 * i)));
        emit(gl::sub(v4, w(2 * i + 1)));
#pragma unroll 1
        for (int k = 0; k < 32; k++) emit(range2(w(b0 + k)));
    }
}
// uninterleave_to_u32.rs:91-134 (B32 = false) / uninterleave_to_b32.rs:115-168 (B32 = true); param = num_ops; 64 big-endian bits per op
template <bool B32, class Wires, class Emit>
__device__ __forceinline__ void uninterleave_gate(Wires&& w, Emit&& emit, uint32_t num_ops) {
    for (uint32_t i = 0; i < num_ops; i++) {
        const uint32_t b0 = 3 * num_ops + 64 * i;
        uint64_t v2 = 0, ce = 0, co = 0;
        for (int k = 0; k < 64; k++) v2 = gl::add(dbl(v2), w(b0 + k));
        for (int j = 0; j < 32; j++) {                                        // Horner over j: coefficient 2^(31-j) or 4^(31-j)
            ce = gl::add(B32 ? times4(ce) : dbl(ce), w(b0 + 2 * j));
            co = gl::add(B32 ? times4(co) : dbl(co), w(b0 + 2 * j + 1));
        }
        emit(gl::sub(v2, w(3 * i)));
        emit(gl::sub(ce, w(3 * i + 1)));
        emit(gl::sub(co, w(3 * i + 2)));
#pragma unroll 1
        for (int k = 0; k < 64; k++) emit(range2(w(b0 + k)));
    }
}
// comparison.rs:112-190; param = num_bits | num_chunks << 8
template <class Wires, class Emit>
__device__ __forceinline__ void comparison_gate(Wires&& w, Emit&& emit, uint32_t param) {
    const uint32_t num_bits = param & 0xFF, nc = param >> 8, chunk_bits = (num_bits + nc - 1) / nc;
    const uint64_t base = 1ULL << chunk_bits;
    uint64_t f = 0, s = 0;
    for (int i = (int)nc - 1; i >= 0; i--) {
        f = gl::add(gl::mulc(f, base), w(4 + i));
        s = gl::add(gl::mulc(s, base), w(4 + nc + i));
    }
    emit(gl::sub(f, w(0)));
    emit(gl::sub(s, w(1)));
    uint64_t msd = 0;
    for (uint32_t i = 0; i < nc; i++) {
        const uint64_t fc = w(4 + i), sc = w(4 + nc + i);
        uint64_t pf = 1, ps = 1;
        for (uint64_t x = 0; x < base; x++) { pf = gl::mulc(pf, gl::sub(fc, x)); ps = gl::mulc(ps, gl::sub(sc, x)); }
        emit(pf);
        emit(ps);
        const uint64_t diff = gl::sub(sc, fc), dummy = w(4 + 2 * nc + i), eq = w(4 + 3 * nc + i), inter = w(4 + 4 * nc + i);
        emit(gl::sub(gl::mulc(diff, dummy), gl::sub(1, eq)));
        emit(gl::mulc(eq, diff));
        emit(gl::sub(inter, gl::mulc(eq, msd)));
        msd = gl::add(inter, gl::mulc(gl::sub(1, eq), diff));
    }
    emit(gl::sub(w(3), msd));
    uint64_t comb = 0;
    for (uint32_t k = 0; k <= chunk_bits; k++) {
        const uint64_t bit = w(4 + 5 * nc + k);
        emit(gl::mulc(bit, gl::sub(1, bit)));
    }
    for (int k = (int)chunk_bits; k >= 0; k--) comb = gl::add(dbl(comb), w(4 + 5 * nc + k));
    emit(gl::sub(gl::add(base, w(3)), comb));
    emit(gl::sub(w(2), w(4 + 5 * nc + chunk_bits)));
}

enum { GATE_POSEIDON2 = 0, GATE_U32_ARITHMETIC = 1, GATE_U32_ADD_MANY = 2, GATE_U32_SUBTRACTION = 3, GATE_U32_RANGE_CHECK = 4,
       GATE_U32_INTERLEAVE = 5, GATE_UNINTERLEAVE_TO_U32 = 6, GATE_UNINTERLEAVE_TO_B32 = 7, GATE_COMPARISON = 8 };

template <class Wires, class Emit>
__device__ __forceinline__ void eval_gate(int kind, uint32_t param, Wires&& w, Emit&& emit) {
    switch (kind) {
        case GATE_POSEIDON2: poseidon2_gate(w, emit); break;
        case GATE_U32_ARITHMETIC: u32_arithmetic_gate(w, emit, param); break;
        case GATE_U32_ADD_MANY: add_many_gate(w, emit, param); break;
        case GATE_U32_SUBTRACTION: subtraction_gate(w, emit, param); break;
        case GATE_U32_RANGE_CHECK: range_check_gate(w, emit, param); break;
        case GATE_U32_INTERLEAVE: interleave_gate(w, emit, param); break;
        case GATE_UNINTERLEAVE_TO_U32: uninterleave_gate<false>(w, emit, param); break;
        case GATE_UNINTERLEAVE_TO_B32: uninterleave_gate<true>(w, emit, param); break;
        default: comparison_gate(w, emit, param); break;
    }
}

// every constraint of every row, uncombined (tests): out[row][i]
__global__ void gate_constraints_kernel(int kind, uint32_t param, const uint64_t* __restrict__ rows, uint32_t pitch, uint64_t n_rows,
                                        uint32_t n_constraints, uint64_t* __restrict__ out) {
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const uint64_t* r = rows + row * pitch;
    uint64_t* o = out + row * n_constraints;
    uint32_t k = 0;
    eval_gate(kind, param, [&](int i) { return gl::canon(__ldg(r + i)); }, [&](uint64_t v) { o[k++] = v; });
}

struct QuotientArgs {
    const uint64_t* rows;          // wires: [n_rows][pitch] (the leaf matrix of the wires commit)
    const uint64_t* filter;        // optional per-row multiplier (column `filter_col` of a [n_rows][filter_pitch] matrix), or nullptr
    uint64_t* acc;                 // [n_challenges][n_rows], += in place
    const uint64_t* powers;        // [n_challenges][n_constraints]: alpha_k^(offset + i), canonical
    uint64_t n_rows;
    uint32_t pitch, filter_pitch, filter_col, n_constraints, n_challenges, kind, param;
};

// acc[k][row] += filter(row) * sum_i powers[k][i] * constraint_i(row)
template <int NCH>
__global__ void __launch_bounds__(128) gate_quotient_kernel(const QuotientArgs a) {
    extern __shared__ uint64_t pw[];   // [NCH][n_constraints]
    for (uint32_t i = threadIdx.x; i < NCH * a.n_constraints; i += blockDim.x) pw[i] = a.powers[i];
    __syncthreads();
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= a.n_rows) return;
    const uint64_t* r = a.rows + row * a.pitch;
    uint64_t acc[NCH];
#pragma unroll
    for (int k = 0; k < NCH; k++) acc[k] = 0;
    uint32_t i = 0;
    eval_gate((int)a.kind, a.param, [&](int j) { return gl::canon(__ldg(r + j)); },
              [&](uint64_t v) {
#pragma unroll
                  for (int k = 0; k < NCH; k++) acc[k] = gl::add(acc[k], gl::mulc(v, pw[k * a.n_constraints + i]));
                  i++;
              });
    uint64_t f = 1;
    if (a.filter) f = gl::canon(__ldg(a.filter + row * a.filter_pitch + a.filter_col));
#pragma unroll
    for (int k = 0; k < NCH; k++) {
        uint64_t v = a.filter ? gl::mulc(acc[k], f) : acc[k];
        uint64_t* dst = a.acc + (uint64_t)k * a.n_rows + row;
        *dst = gl::add(*dst, v);
    }
}

// tail of compute_quotient_polys: vals[i][k] = acc[k][bitrev(i)] / Z_H(g w_R^i), natural point order, row-major [R][pitch];
// Z_H on the coset takes 2^rate_bits values: zh_inv[i mod 2^rate_bits] (plonky2 plonk/plonk_common.rs · ZeroPolyOnCoset::eval_inverse)
__global__ void quotient_gather_kernel(const uint64_t* __restrict__ acc, uint64_t* __restrict__ out, uint32_t pitch, uint32_t n_ch, uint32_t bits,
                                       uint32_t rate_bits, const uint64_t* __restrict__ zh_inv) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, R = 1ULL << bits;
    if (i >= R) return;
    const uint64_t src = gl::bitrev32((uint32_t)i, bits);
    const uint64_t zi = zh_inv[i & ((1u << rate_bits) - 1)];
    for (uint32_t k = 0; k < pitch; k++) out[i * pitch + k] = k < n_ch ? gl::mulc(acc[(uint64_t)k * R + src], zi) : 0;
}
// coefficients of the coset iFFT -> the quotient chunk polynomials: out[(k * n_chunks + c)][m] = coeffs[c * N + m][k] * shift^-(c N + m)
__global__ void quotient_chunks_kernel(const uint64_t* __restrict__ coeffs, uint32_t pitch, uint32_t n_ch, uint32_t log_n, uint32_t rate_bits,
                                       const uint64_t* __restrict__ shift_inv_pow, uint64_t* __restrict__ out) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, R = 1ULL << (log_n + rate_bits), N = 1ULL << log_n;
    if (j >= R) return;
    const uint64_t c = j >> log_n, m = j & (N - 1), s = shift_inv_pow[j];
    for (uint32_t k = 0; k < n_ch; k++) out[((uint64_t)k << rate_bits | c) * N + m] = gl::mulc(coeffs[j * pitch + k], s);
}

// Poseidon2Generator::run_once for a batch of rows: in[n][13] = 12 inputs + swap flag -> out[n][out_pitch] (135 wires)
__global__ void poseidon2_witness_kernel(const uint64_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ out, uint32_t out_pitch) {
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const uint64_t* x = in + row * 13;
    uint64_t* w = out + row * out_pitch;
    uint64_t s[P2_WIDTH];
#pragma unroll
    for (int i = 0; i < P2_WIDTH; i++) { s[i] = gl::canon(x[i]); w[i] = s[i]; }
    const uint64_t swap = gl::canon(x[12]);
    w[P2_WIRE_SWAP] = swap;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        w[P2_START_DELTA + i] = gl::mulc(swap, gl::sub(s[i + 4], s[i]));
        if (swap == 1) { const uint64_t t = s[i]; s[i] = s[i + 4]; s[i + 4] = t; }   // `if swap_value == F::ONE { state.swap(i, 4 + i) }`
    }
    matmul_external(s);
#pragma unroll 1
    for (int r = 0; r < P2_RF_BEGIN; r++) {
#pragma unroll
        for (int i = 0; i < P2_WIDTH; i++) {
            const uint64_t v = gl::add(s[i], poseidon2::RC[r][i]);
            if (r != 0) w[P2_START_RF_BEGIN + P2_WIDTH * (r - 1) + i] = v;
            s[i] = sbox7(v);
        }
        matmul_external(s);
    }
#pragma unroll 1
    for (int r = 0; r < P2_RP; r++) {
        const uint64_t v = gl::add(s[0], poseidon2::RC_MID[r]);
        w[P2_START_PARTIAL + r] = v;
        s[0] = sbox7(v);
        matmul_internal(s);
    }
#pragma unroll 1
    for (int r = P2_RF_BEGIN; r < P2_RF_END; r++) {
#pragma unroll
        for (int i = 0; i < P2_WIDTH; i++) {
            const uint64_t v = gl::add(s[i], poseidon2::RC[r][i]);
            w[P2_START_RF_END + P2_WIDTH * (r - P2_RF_BEGIN) + i] = v;
            s[i] = sbox7(v);
        }
        matmul_external(s);
    }
#pragma unroll
    for (int i = 0; i < P2_WIDTH; i++) w[P2_WIDTH + i] = s[i];
}

}  // namespace gates
