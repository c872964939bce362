// K9: front half of PolynomialBatch::prove_openings (plonky2 fri/oracle.rs) on device-resident coefficient matrices:
//   composition  F = sum_j alpha^j f_j            (util/reducing.rs · ReducingFactor::reduce_polys_base)
//   quotient     (F(X) - F(z)) / (X - z)          (field/polynomial/division.rs · divide_by_linear, padded back to N)
//   final_poly <- final_poly * alpha^count + quotient   (ReducingFactor::shift_poly, then +=)
// driven in the reference from /root/reference/src/p3/mod.rs:260 (`data.prove`).  Polynomials are columns of the
// row-major coefficient matrices [N][pitch] the commits left in HBM, so coefficient k of every polynomial of a batch is
// one contiguous run of a row — the same access pattern as the leaf hash.  All arithmetic is exact field arithmetic, so
// any association order gives the reference's canonical result.
#pragma once
#include "fri.cuh"

namespace openings {

struct PolyRef {
    const uint64_t* base;   // coefficient matrix of the oracle
    uint32_t pitch, col;
};

// out[k] = sum_j pw[j] * f_j[k]  (pw[j] = alpha^j, extension; f_j base field).  One thread per coefficient index.
__global__ void reduce_polys_base_kernel(const PolyRef* __restrict__ polys, const ulonglong2* __restrict__ pw, uint32_t n_polys,
                                         uint32_t n, ulonglong2* __restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint64_t a0 = 0, a1 = 0;
    for (uint32_t j = 0; j < n_polys; j++) {
        const PolyRef r = polys[j];
        const uint64_t v = r.base[(size_t)k * r.pitch + r.col];
        const ulonglong2 w = pw[j];
        a0 = gl::add(a0, gl::mulc(v, w.x));
        a1 = gl::add(a1, gl::mulc(v, w.y));
    }
    out[k] = make_ulonglong2(a0, a1);
}

__device__ __forceinline__ fri::Ext ext_add(fri::Ext a, fri::Ext b) { return {gl::add(a.a0, b.a0), gl::add(a.a1, b.a1)}; }

// In place: f (n extension coefficients) -> q with q[k] = f[k+1] + z*q[k+1], q[n-1] = 0  (synthetic division by X - z).
// One CTA; thread t owns the chunk [t*L, (t+1)*L): (1) Horner value of its chunk, (2) a serial right-to-left carry over
// the <= 1024 chunk values by thread 0, (3) the chunk's quotient coefficients from its carry-in.
__global__ void divide_by_linear_kernel(ulonglong2* __restrict__ f, uint32_t n, uint64_t z0, uint64_t z1) {
    __shared__ fri::Ext chunk_val[1024];
    __shared__ fri::Ext carry_in[1024];
    const uint32_t T = blockDim.x, t = threadIdx.x, L = n / T;   // n, T powers of two, T <= n
    const fri::Ext z = {z0, z1};
    // (1) h_t = sum_{i in chunk} f[i] z^(i - t*L)
    fri::Ext h = {0, 0};
    for (uint32_t i = L; i-- > 0;) {
        const ulonglong2 c = f[(size_t)t * L + i];
        h = fri::ext_mul(h, z);
        h = ext_add(h, {c.x, c.y});
    }
    chunk_val[t] = h;
    __syncthreads();
    // (2) carry_in[t] = sum_{i >= (t+1)L} f[i] z^(i - (t+1)L)  (= q[(t+1)L - 1], the quotient coefficient just left of chunk t+1)
    if (t == 0) {
        fri::Ext zL = {1, 0};
        for (uint32_t i = 0; i < L; i++) zL = fri::ext_mul(zL, z);
        fri::Ext acc = {0, 0};
        for (uint32_t u = T; u-- > 0;) {
            carry_in[u] = acc;
            acc = ext_add(fri::ext_mul(acc, zL), chunk_val[u]);
        }
    }
    __syncthreads();
    // (3) q[k] = f[k+1] + z*q[k+1] inside the chunk, starting from q[(t+1)L - 1] = carry_in[t]
    fri::Ext q = carry_in[t];
    for (uint32_t i = L; i-- > 0;) {
        const size_t k = (size_t)t * L + i;
        const ulonglong2 c = f[k];                      // f[k] is needed for q[k-1]; read before overwriting
        f[k] = make_ulonglong2(q.a0, q.a1);
        q = ext_add(fri::ext_mul(q, z), {c.x, c.y});
    }
}

// final[k] = final[k] * s + q[k]
__global__ void shift_add_kernel(ulonglong2* __restrict__ fin, const ulonglong2* __restrict__ q, uint32_t n, uint64_t s0, uint64_t s1) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const ulonglong2 a = fin[k], b = q[k];
    fri::Ext r = fri::ext_mul({a.x, a.y}, {s0, s1});
    fin[k] = make_ulonglong2(gl::add(r.a0, b.x), gl::add(r.a1, b.y));
}

}  // namespace openings
