// §8(f) rank 4: the permutation argument's partial products and Z polynomials.
//
// Replaces plonky2 plonk/prover.rs · all_wires_permutation_partial_products / wires_permutation_partial_products_and_zs and
// plonk/permutation_argument... · quotient_chunk_products / partial_products_and_z_gx (upstream source is NOT in /root/reference —
// restated from the published algorithm; "parity unpinned", oracle/gates_oracle.py restates the same text).  Driven in
// the reference from /root/reference/src/p3/mod.rs:260 (`data.prove(pw)`, between the wires commit and the Z commit).
//
// Per challenge (beta, gamma) and row i (x_i = w^i, the subgroup in natural order), with n_routed wires in chunks of `degree`:
//     numerator_j   = wire_ij + beta * k_j * x_i + gamma            denominator_j = wire_ij + beta * sigma_ij + gamma
//     q_ic          = prod_{j in chunk c} numerator_j / denominator_j
//     running product over (row-major) (i, c):  acc <- acc * q_ic, starting at Z(x_0) = 1;  partial product pp_c(x_i) = acc after chunk
//     c < n_chunks - 1, and the value after the last chunk is Z(x_{i+1}).
// Output polynomials (values on the subgroup) in prove()'s order: the Z of every challenge first, then each challenge's partial products.
#pragma once
#include "gl_field.cuh"

namespace perm {

constexpr int MAX_CHUNKS = 16;

// x^(p - 2), p - 2 = 0xFFFFFFFE_FFFFFFFF = (2^31 - 1) * 2^33 + (2^32 - 1): 64 squarings + 10 multiplies; inv(0) = 0
__device__ __forceinline__ uint64_t inverse(uint64_t x) {
    auto sqn = [](uint64_t v, int n) { for (int i = 0; i < n; i++) v = gl::sqr(v); return v; };
    const uint64_t t2 = gl::mul(gl::sqr(x), x);
    const uint64_t t4 = gl::mul(sqn(t2, 2), t2);
    const uint64_t t8 = gl::mul(sqn(t4, 4), t4);
    const uint64_t t16 = gl::mul(sqn(t8, 8), t8);
    const uint64_t t24 = gl::mul(sqn(t16, 8), t8);
    const uint64_t t28 = gl::mul(sqn(t24, 4), t4);
    const uint64_t t30 = gl::mul(sqn(t28, 2), t2);
    const uint64_t t31 = gl::mul(gl::sqr(t30), x);          // x^(2^31 - 1)
    const uint64_t t32 = gl::mul(gl::sqr(t31), x);          // x^(2^32 - 1)
    return gl::mulc(sqn(t31, 33), t32);
}

struct Args {
    const uint64_t* wires;      // [N][pitch] row-major, canonical
    const uint64_t* sigmas;     // [N][pitch]
    const uint64_t* k_is;       // [n_routed]
    const uint64_t* xs;         // [N]: w^i
    uint64_t* q;                // [n_ch][N][n_chunks] chunk quotients
    uint64_t* out;              // [(n_ch * n_chunks)][N] column-major result
    uint64_t beta[4], gamma[4];
    uint32_t n, pitch, n_routed, degree, n_chunks, n_ch;
};

// thread (i, k): the n_chunks chunk quotients of row i under challenge k (one field inversion per thread: batch inverse of the chunk denominators)
__global__ void __launch_bounds__(128) chunk_quotients_kernel(const Args a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (i >= a.n) return;
    const uint64_t beta = a.beta[k], gamma = a.gamma[k];
    const uint64_t bx = gl::mulc(beta, a.xs[i]);
    const uint64_t* w = a.wires + (size_t)i * a.pitch;
    const uint64_t* s = a.sigmas + (size_t)i * a.pitch;
    uint64_t num[MAX_CHUNKS], den[MAX_CHUNKS];
    for (uint32_t c = 0; c < a.n_chunks; c++) {
        uint64_t pn = 1, pd = 1;
        const uint32_t j1 = min((c + 1) * a.degree, a.n_routed);
        for (uint32_t j = c * a.degree; j < j1; j++) {
            const uint64_t wv = w[j];
            const uint64_t nj = gl::add(gl::add(wv, gl::mulc(bx, a.k_is[j])), gamma);
            const uint64_t dj = gl::add(gl::add(wv, gl::mulc(beta, s[j])), gamma);
            pn = gl::mulc(pn, nj);
            pd = gl::mulc(pd, dj);
        }
        num[c] = pn; den[c] = pd;
    }
    // batch inverse of den[0..n_chunks): prefix products, one inversion, back-substitution
    uint64_t pre[MAX_CHUNKS];
    uint64_t acc = 1;
    for (uint32_t c = 0; c < a.n_chunks; c++) { pre[c] = acc; acc = gl::mulc(acc, den[c]); }
    uint64_t inv = inverse(acc);
    uint64_t* q = a.q + ((size_t)k * a.n + i) * a.n_chunks;
    for (uint32_t c = a.n_chunks; c-- > 0;) {
        const uint64_t dinv = gl::mulc(inv, pre[c]);
        inv = gl::mulc(inv, den[c]);
        q[c] = gl::mulc(num[c], dinv);
    }
}

// one CTA per challenge: running product over (i, c) in row-major order.  Thread t owns rows [t*L, (t+1)*L): (1) product of its rows'
// quotients, (2) exclusive scan of the 1024 segment products in shared memory, (3) walk the segment writing Z(x_i) and the partial products.
__global__ void __launch_bounds__(1024) running_product_kernel(const Args a) {
    __shared__ uint64_t seg[1024];
    const uint32_t k = blockIdx.x, T = blockDim.x, t = threadIdx.x;
    const uint32_t L = (a.n + T - 1) / T, i0 = min(t * L, a.n), i1 = min(i0 + L, a.n);
    const uint64_t* q = a.q + (size_t)k * a.n * a.n_chunks;
    uint64_t prod = 1;
    for (uint32_t i = i0; i < i1; i++)
        for (uint32_t c = 0; c < a.n_chunks; c++) prod = gl::mulc(prod, q[(size_t)i * a.n_chunks + c]);
    seg[t] = prod;
    __syncthreads();
    for (uint32_t d = 1; d < T; d <<= 1) {          // inclusive Hillis-Steele scan (multiplication)
        const uint64_t v = t >= d ? gl::mulc(seg[t], seg[t - d]) : seg[t];
        __syncthreads();
        seg[t] = v;
        __syncthreads();
    }
    uint64_t z = t ? seg[t - 1] : 1;                // Z(x_{i0}) = product of everything before this segment
    const uint32_t n_pp = a.n_chunks - 1;
    uint64_t* out_z = a.out + (size_t)k * a.n;
    uint64_t* out_pp = a.out + ((size_t)a.n_ch + (size_t)k * n_pp) * a.n;
    for (uint32_t i = i0; i < i1; i++) {
        out_z[i] = z;
        for (uint32_t c = 0; c < a.n_chunks; c++) {
            z = gl::mulc(z, q[(size_t)i * a.n_chunks + c]);
            if (c < n_pp) out_pp[(size_t)c * a.n + i] = z;
        }
    }
}

// The permutation-argument terms of plonky2 plonk/vanishing_poly.rs · eval_vanishing_poly_base_batch (restated; upstream source absent),
// evaluated at every LDE row and folded into the quotient accumulator with the challenges' alpha powers.  vanishing_terms is ordered
//     [ L_0(x) (Z_i(x) - 1)  for every challenge i ]  ++  [ check_partial_products of challenge 0, of challenge 1, ... ]  ++  gate constraints
// and EVERY alpha_k reduces the whole list (reduce_with_powers), so term t contributes alpha_k^t * term_t to acc[k].
// check_partial_products: with product_accs = [Z_i(x), pp_i0, .., pp_i(m-2), Z_i(g x)], chunk c gives
//     accs[c] * prod_{j in chunk c} (w_j + beta_i k_j x + gamma_i)  -  accs[c+1] * prod_{j in chunk c} (w_j + beta_i sigma_j(x) + gamma_i).
// Rows are leaf rows (row = LDE point index bitrev(row)); Z_i(g x) is the row of LDE point index + 2^rate_bits.
struct VanishArgs {
    const uint64_t* wires;  uint32_t wires_pitch;       // [R][pitch]: routed wires are columns [0, n_routed)
    const uint64_t* sigmas; uint32_t sigmas_pitch, sigma_col0;
    const uint64_t* zs;     uint32_t zs_pitch;          // columns: Z_0..Z_{n_ch-1}, then (n_chunks-1) partial products per challenge
    const uint64_t* k_is;                               // [n_routed]
    const uint64_t* W;                                  // w_R^e, e < R/2
    const uint64_t* powers;                             // [n_ch][n_terms]: alpha_k^t
    uint64_t* acc;                                      // [n_ch][R]
    uint64_t beta[4], gamma[4], zh[64], n_inv;          // zh[s] = (7 w_R^s)^N - 1, s < 2^rate_bits
    uint32_t log_n, rate_bits, n_routed, degree, n_chunks, n_ch;
};

template <int NCH>
__global__ void __launch_bounds__(128) vanishing_perm_kernel(const VanishArgs a) {
    extern __shared__ uint64_t pw[];   // [NCH][n_terms]
    const uint32_t n_terms = NCH * (1 + a.n_chunks);
    for (uint32_t i = threadIdx.x; i < NCH * n_terms; i += blockDim.x) pw[i] = a.powers[i];
    __syncthreads();
    const uint32_t bits = a.log_n + a.rate_bits;
    const uint64_t R = 1ULL << bits, row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= R) return;
    const uint32_t idx = gl::bitrev32((uint32_t)row, bits);                       // LDE point index
    const uint32_t idx_next = (idx + (1u << a.rate_bits)) & (uint32_t)(R - 1);
    const uint64_t row_next = gl::bitrev32(idx_next, bits);
    // x = 7 * w_R^idx
    const uint32_t half = (uint32_t)(R >> 1);
    const uint64_t wr = idx >= half ? gl::P - a.W[idx - half] : a.W[idx];
    const uint64_t x = gl::mulc(wr, gl::COSET_SHIFT);
    // L_0(x) = (x^N - 1) / (N (x - 1))
    const uint64_t zh = a.zh[idx & ((1u << a.rate_bits) - 1)];
    const uint64_t l0 = gl::mulc(gl::mulc(zh, a.n_inv), inverse(gl::sub(x, 1)));
    const uint64_t* w = a.wires + row * a.wires_pitch;
    const uint64_t* s = a.sigmas + row * a.sigmas_pitch + a.sigma_col0;
    const uint64_t* z = a.zs + row * a.zs_pitch;
    const uint64_t* zn = a.zs + row_next * a.zs_pitch;
    uint64_t acc[NCH];
#pragma unroll
    for (int k = 0; k < NCH; k++) acc[k] = 0;
    auto emit = [&](uint32_t t, uint64_t v) {
#pragma unroll
        for (int k = 0; k < NCH; k++) acc[k] = gl::add(acc[k], gl::mulc(v, pw[k * n_terms + t]));
    };
    const uint32_t n_pp = a.n_chunks - 1;
#pragma unroll 1
    for (uint32_t i = 0; i < (uint32_t)NCH; i++) {
        const uint64_t z_x = gl::canon(z[i]), z_gx = gl::canon(zn[i]);
        emit(i, gl::mulc(l0, gl::sub(z_x, 1)));
        const uint64_t bx = gl::mulc(a.beta[i], x);
        uint64_t prev = z_x;
#pragma unroll 1
        for (uint32_t c = 0; c < a.n_chunks; c++) {
            uint64_t pn = 1, pd = 1;
            const uint32_t j1 = min((c + 1) * a.degree, a.n_routed);
            for (uint32_t j = c * a.degree; j < j1; j++) {
                const uint64_t wv = gl::canon(w[j]);
                pn = gl::mulc(pn, gl::add(gl::add(wv, gl::mulc(bx, a.k_is[j])), a.gamma[i]));
                pd = gl::mulc(pd, gl::add(gl::add(wv, gl::mulc(a.beta[i], gl::canon(s[j]))), a.gamma[i]));
            }
            const uint64_t next = c + 1 < a.n_chunks ? gl::canon(z[NCH + i * n_pp + c]) : z_gx;
            emit(NCH + i * a.n_chunks + c, gl::sub(gl::mulc(prev, pn), gl::mulc(next, pd)));
            prev = next;
        }
    }
#pragma unroll
    for (int k = 0; k < NCH; k++) {
        uint64_t* dst = a.acc + (uint64_t)k * R + row;
        *dst = gl::add(*dst, acc[k]);
    }
}

}  // namespace perm
