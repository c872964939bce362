// Register-resident Poseidon (width 12, x^7, 4 + 22 + 4 rounds) over Goldilocks for sm_100a.
//
// Replaces plonky2 hash/src/poseidon.rs · Poseidon::poseidon and hash/hashing.rs · hash_n_to_m_no_pad / compress
// (semantics SURVEY.md A.5).  Pinned by the known-answer vectors found at
// /root/reference/src/common/poseidon2/poseidon2_goldilocks.rs:190-211 (tests/golden/poseidon_kat.json).
//
// Formulation chosen for the 32-bit integer pipe:
//   * state = 12 x u64 in registers, kept in "any" (non-canonical) form between layers;
//   * S-box x^7 = 2 squarings + 2 multiplies, each 64x64->128 (4 IMAD.WIDE.U32) + reduce128;
//   * MDS layer on the 32-bit halves of every lane: 2 x 144 small-constant IMAD.WIDE.U32 accumulations
//     (sums stay below 2^42), recombined once per lane; the NEXT round's constant is folded into the accumulator
//     initial value so there is no separate constant layer.
#pragma once
#include "gl_field.cuh"
#include "poseidon_constants.cuh"

namespace poseidon {

constexpr int WIDTH = 12;
constexpr int RATE = 8;
constexpr int N_FULL_HALF = 4;
constexpr int N_PARTIAL = 22;
constexpr int N_ROUNDS = 30;

// circulant first row and diagonal (plonky2 hash/poseidon_goldilocks.rs · MDS_MATRIX_CIRC / MDS_MATRIX_DIAG)
#define PSD_C0 17u
#define PSD_C1 15u
#define PSD_C2 41u
#define PSD_C3 16u
#define PSD_C4 2u
#define PSD_C5 28u
#define PSD_C6 13u
#define PSD_C7 13u
#define PSD_C8 39u
#define PSD_C9 18u
#define PSD_C10 34u
#define PSD_C11 20u
#define PSD_DIAG0 8u

__device__ __forceinline__ uint64_t sbox7(uint64_t x) {
    uint64_t x2 = gl::sqr(x);
    uint64_t x4 = gl::sqr(x2);
    uint64_t x3 = gl::mul(x, x2);
    return gl::mul(x3, x4);
}

// out[r] = sum_i s[(i+r)%12]*C[i] + (r==0)*8*s[0] + rc[r]   (mod p), inputs "any", outputs "any".
// rc == nullptr -> no constant.
template <bool HAS_RC>
__device__ __forceinline__ void mds_layer(uint64_t (&s)[WIDTH], const uint64_t* __restrict__ rc) {
    constexpr uint32_t C[WIDTH] = {PSD_C0, PSD_C1, PSD_C2, PSD_C3, PSD_C4, PSD_C5,
                                   PSD_C6, PSD_C7, PSD_C8, PSD_C9, PSD_C10, PSD_C11};
    uint32_t lo[WIDTH], hi[WIDTH];
#pragma unroll
    for (int i = 0; i < WIDTH; i++) {
        lo[i] = (uint32_t)s[i];
        hi[i] = (uint32_t)(s[i] >> 32);
    }
#pragma unroll
    for (int r = 0; r < WIDTH; r++) {
        uint64_t al = 0, ah = 0;
        if (HAS_RC) {
            uint64_t k = rc[r];
            al = (uint32_t)k;
            ah = k >> 32;
        }
#pragma unroll
        for (int i = 0; i < WIDTH; i++) {
            al += (uint64_t)lo[(i + r) % WIDTH] * C[i];
            ah += (uint64_t)hi[(i + r) % WIDTH] * C[i];
        }
        if (r == 0) {
            al += (uint64_t)lo[0] * PSD_DIAG0;
            ah += (uint64_t)hi[0] * PSD_DIAG0;
        }
        // value = al + ah*2^32,  al, ah < 2^42.   ah*2^32 = (ah_lo << 32) + ah_hi * 2^64,  2^64 = EPS
        uint64_t x = al + ((uint64_t)(uint32_t)ah << 32);
        uint32_t top = (uint32_t)(ah >> 32) + (x < al ? 1u : 0u);   // multiples of 2^64, < 2^11
        uint64_t y = (uint64_t)top * (uint32_t)gl::EPS;
        uint64_t z = x + y;
        if (z < x) z += gl::EPS;   // wrapped z < 2^43, cannot overflow again
        s[r] = z;
    }
}

// Full permutation. in/out "any" -> "any" (callers canonicalise what they store).
__device__ __forceinline__ void permute(uint64_t (&s)[WIDTH]) {
    // round-0 constants (all later constant layers are folded into the preceding MDS layer)
#pragma unroll
    for (int i = 0; i < WIDTH; i++) s[i] = gl::add_any_c(s[i], POSEIDON_RC[i]);
    int r = 0;
#pragma unroll 1
    for (; r < N_FULL_HALF; r++) {
#pragma unroll
        for (int i = 0; i < WIDTH; i++) s[i] = sbox7(s[i]);
        mds_layer<true>(s, &POSEIDON_RC[WIDTH * (r + 1)]);
    }
#pragma unroll 1
    for (; r < N_FULL_HALF + N_PARTIAL; r++) {
        s[0] = sbox7(s[0]);
        mds_layer<true>(s, &POSEIDON_RC[WIDTH * (r + 1)]);
    }
#pragma unroll 1
    for (; r < N_ROUNDS - 1; r++) {
#pragma unroll
        for (int i = 0; i < WIDTH; i++) s[i] = sbox7(s[i]);
        mds_layer<true>(s, &POSEIDON_RC[WIDTH * (r + 1)]);
    }
#pragma unroll
    for (int i = 0; i < WIDTH; i++) s[i] = sbox7(s[i]);
    mds_layer<false>(s, nullptr);
}

}  // namespace poseidon
