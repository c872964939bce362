// Goldilocks field p = 2^64 - 2^32 + 1 on the sm_100a 32-bit integer datapath.
//
// Replaces plonky2 field/src/goldilocks_field.rs (add/sub/mul/reduce128) — semantics per SURVEY.md A.1; the
// constants are the ones visible at /root/reference/src/p3/mod.rs:55 (modulus) and
// /root/reference/src/p3/extension.rs:149,155 (generator 7, 2^32-th root 1753635133440165772).
//
// Representation rules used by the kernels:
//   * "canonical"  : value in [0, p)
//   * "any"        : value in [0, 2^64) congruent to the element (what reduce128 yields)
//   mul()/sqr() accept "any" operands and return "any"; add()/sub() need canonical operands and return
//   canonical; add_any_c()/sub_any_c() take (any, canonical) and return "any".
#pragma once
#include <stdint.h>

namespace gl {

constexpr uint64_t P = 0xFFFFFFFF00000001ULL;
constexpr uint64_t EPS = 0xFFFFFFFFULL;  // 2^64 mod p

__host__ __device__ __forceinline__ uint64_t canon(uint64_t x) { return x >= P ? x - P : x; }

// 128 -> 64 ("any"): lo + 2^64*hi, with 2^64 = 2^32 - 1 and 2^96 = -1 (mod p)
__device__ __forceinline__ uint64_t reduce128(uint64_t lo, uint64_t hi) {
    uint32_t hl = (uint32_t)hi;
    uint64_t hh = hi >> 32;
    uint64_t t0 = lo - hh;
    if (lo < hh) t0 -= EPS;                 // borrow: wrapped value is 2^64 too big, 2^64 = EPS
    uint64_t t1 = (uint64_t)hl * (uint32_t)EPS;  // mul.wide.u32
    uint64_t t2 = t0 + t1;
    if (t2 < t0) t2 += EPS;                 // carry: cannot overflow again
    return t2;
}

__device__ __forceinline__ uint64_t mul(uint64_t a, uint64_t b) { return reduce128(a * b, __umul64hi(a, b)); }
__device__ __forceinline__ uint64_t sqr(uint64_t a) { return mul(a, a); }
__device__ __forceinline__ uint64_t mulc(uint64_t a, uint64_t b) { return canon(mul(a, b)); }

// canonical (+/-) canonical -> canonical
__device__ __forceinline__ uint64_t add(uint64_t a, uint64_t b) {
    uint64_t nb = P - b;                    // in (0, p]
    uint64_t d = a - nb;
    if (a < nb) d += P;
    return d;
}
__device__ __forceinline__ uint64_t sub(uint64_t a, uint64_t b) {
    uint64_t d = a - b;
    if (a < b) d += P;
    return d;
}
// any + canonical -> any
__device__ __forceinline__ uint64_t add_any_c(uint64_t a, uint64_t b) {
    uint64_t s = a + b;
    if (s < a) s += EPS;                    // a+b < 2^64 + p  =>  wrapped < p  =>  no second overflow
    return s;
}

__host__ __device__ __forceinline__ uint32_t bitrev32(uint32_t x, uint32_t bits) {
#ifdef __CUDA_ARCH__
    return bits ? (__brev(x) >> (32 - bits)) : 0;
#else
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
#endif
}

// ---- host-side helpers (table generation in the context; not on any hot path) -------------------------
__host__ inline uint64_t h_mul(uint64_t a, uint64_t b) {
    unsigned __int128 x = (unsigned __int128)a * b;
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hh = hi >> 32, hl = hi & EPS;
    uint64_t t0 = lo - hh;
    if (lo < hh) t0 -= EPS;
    uint64_t t1 = hl * EPS;
    uint64_t t2 = t0 + t1;
    if (t2 < t0) t2 += EPS;
    return canon(t2);
}
__host__ inline uint64_t h_pow(uint64_t b, uint64_t e) {
    uint64_t r = 1;
    while (e) {
        if (e & 1) r = h_mul(r, b);
        b = h_mul(b, b);
        e >>= 1;
    }
    return r;
}
__host__ inline uint64_t h_inv(uint64_t a) { return h_pow(a, P - 2); }
constexpr uint64_t POWER_OF_TWO_GENERATOR = 1753635133440165772ULL;  // order 2^32
constexpr uint64_t COSET_SHIFT = 7;                                  // MULTIPLICATIVE_GROUP_GENERATOR
__host__ inline uint64_t h_root_of_unity(uint32_t n_log) { return h_pow(POWER_OF_TWO_GENERATOR, 1ULL << (32 - n_log)); }

}  // namespace gl
