// Register-resident Poseidon (width 12, x^7, 4 + 22 + 4 rounds) over Goldilocks for sm_100a.
//
// Replaces plonky2 hash/src/poseidon.rs · Poseidon::poseidon and hash/hashing.rs · hash_n_to_m_no_pad / compress
// (semantics SURVEY.md A.5).  Pinned by the known-answer vectors found at
// /root/reference/src/common/poseidon2/poseidon2_goldilocks.rs:190-211 (tests/golden/poseidon_kat.json).
//
// Formulation.  Measured on B200 (tools/ubench.cu): IMAD.WIDE.U32 issues at 32 lanes/clk/SM (fma-heavy
// only), IADD3/LOP3/SHF at 64, DFMA/DADD at 64 on a pipe of their own, and the scheduler sustains ~120 warp-lanes/clk/SM
// across the three.  So the permutation is split over all three pipes and written for the lowest instruction count:
//   * S-box x^7 = 4 multiplies (gl::mul: 4 IMAD.WIDE.U32 + carry chains in PTX) on the fma + alu pipes;
//   * the MDS layer runs on the fp64 pipe.  Each 64-bit lane is cut into two 32-bit limbs; a limb becomes a double for
//     free (register pair {limb, 0x43300000} is the double 2^52 + limb), every sum of the layer stays below 2^53 and is
//     therefore exact, and the results are read back from the mantissa bits.  The circulant is evaluated as a length-12
//     cyclic convolution split by x^12-1 = (x^3-1)(x^3+1)(x^6+1): 103 DFMA/DADD per limb instead of 146 (plonky2 chose
//     the matrix so that every folded constant is an integer — tools/mds_model.py);
//   * the next round's constants ride in the accumulator initial values (POSEIDON_MDS_K), so there is no constant layer.
#pragma once
#include "gl_field.cuh"
#include "poseidon_constants.cuh"

namespace poseidon {

constexpr int WIDTH = 12;
constexpr int RATE = 8;
constexpr int N_FULL_HALF = 4;
constexpr int N_PARTIAL = 22;
constexpr int N_ROUNDS = 30;

__device__ __forceinline__ uint64_t sbox7(uint64_t x) {
    uint64_t x2 = gl::sqr(x);
    uint64_t x4 = gl::sqr(x2);
    uint64_t x3 = gl::mul(x, x2);
    return gl::mul(x3, x4);
}

constexpr double TWO52 = 4503599627370496.0;

// {w, 0x43300000} is the double 2^52 + w
__device__ __forceinline__ double biased(uint32_t w) { return __hiloint2double(0x43300000, (int)w); }
// exact integer value of a 32-bit word as a double
__device__ __forceinline__ double limb_to_double(uint32_t w) { return __dsub_rn(biased(w), TWO52); }

// x^7 handed to the fp64 MDS as two limbs (value = lo + hi*2^32 mod p).  The last product x^3 * x^4 = (z3:z2:z1:z0) is
// not reduced to 64 bits on the integer pipe: 2^64 = 2^32 - 1 and 2^96 = -1 give lo = z0 - z2 - z3 (signed, |lo| < 2^33)
// and hi = z1 + z2 (< 2^33), six exact DADDs — cheaper than reduce_words + conversion, and off the alu pipe.
__device__ __forceinline__ void sbox7_limbs(uint64_t x, double& lo, double& hi) {
    const uint64_t x2 = gl::sqr(x);
    const uint64_t x4 = gl::sqr(x2);
    const uint64_t x3 = gl::mul(x, x2);
    uint32_t z0, z1, z2, z3;
    gl::mul_words(x3, x4, z0, z1, z2, z3);
    const double d2 = biased(z2);
    lo = __dsub_rn(__dsub_rn(biased(z0), d2), __dsub_rn(biased(z3), TWO52));
    hi = __dsub_rn(__dadd_rn(biased(z1), __dsub_rn(d2, TWO52)), TWO52);
}

// One limb of the MDS layer.  s[12]: exact integers (|s| < 2^41); k[12]: folded constants (uu0..2, uv0..2, v0..5);
// o[r] = sum_i s[(i+r)%12]*CIRC[i] + (r==0)*8*s[0] + c[r], with c (and the 2^52 read-out bias) folded into k.
// Mirrors tools/mds_model.py · mds_limb operation by operation.
__device__ __forceinline__ void mds_limb(const double (&s)[WIDTH], const double* __restrict__ k, double (&o)[WIDTH]) {
    double sp[6], sm[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        sp[i] = __dadd_rn(s[i], s[i + 6]);
        sm[i] = __dsub_rn(s[i], s[i + 6]);
    }
    double a[3], b[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        a[i] = __dadd_rn(sp[i], sp[i + 3]);
        b[i] = __dsub_rn(sp[i], sp[i + 3]);
    }
    // cyclic-3 with [16, 32, 16]:  UU[j] = 16*(a0 + a1 + a2) + 16*a[(j+2)%3]
    const double S = __dadd_rn(__dadd_rn(a[0], a[1]), a[2]);
    double U[6], V[6];
    {
        double UU[3], UV[3];
#pragma unroll
        for (int j = 0; j < 3; j++) UU[j] = __fma_rn(a[(j + 2) % 3], 16.0, __fma_rn(S, 16.0, k[j]));
        // negacyclic-3 with [-1, -8, 2]
        UV[0] = __fma_rn(b[2], 8.0, __fma_rn(b[1], -2.0, __dsub_rn(k[3], b[0])));
        UV[1] = __fma_rn(b[2], -2.0, __fma_rn(b[0], -8.0, __dsub_rn(k[4], b[1])));
        UV[2] = __fma_rn(b[1], -8.0, __fma_rn(b[0], 2.0, __dsub_rn(k[5], b[2])));
#pragma unroll
        for (int j = 0; j < 3; j++) {
            U[j] = __dadd_rn(UU[j], UV[j]);
            U[j + 3] = __dsub_rn(UU[j], UV[j]);
        }
    }
    // negacyclic-6 with f = [2, -4, 16, 1, -1, -1]:  V[n] = sum_{j<=n} sm[j] f[n-j] - sum_{j>n} sm[j] f[6+n-j]
    constexpr double f[6] = {2.0, -4.0, 16.0, 1.0, -1.0, -1.0};
#pragma unroll
    for (int n = 0; n < 6; n++) {
        double acc = k[6 + n];
#pragma unroll
        for (int j = 0; j < 6; j++) acc = __fma_rn(sm[j], (j <= n) ? f[n - j] : -f[6 + n - j], acc);
        V[n] = acc;
    }
    // diagonal 8*s[0] reaches o[0] = U0 + V0 but not o[6] = U0 - V0
    U[0] = __fma_rn(s[0], 4.0, U[0]);
    V[0] = __fma_rn(s[0], 4.0, V[0]);
#pragma unroll
    for (int n = 0; n < 6; n++) {
        o[n] = __dadd_rn(U[n], V[n]);
        o[n + 6] = __dsub_rn(U[n], V[n]);
    }
}

// (2^52 + al, 2^52 + ah) -> the 64-bit "any" word congruent to al + ah*2^32 (al, ah < 2^52; here < 2^50).
// With a1 = al >> 32, b1 = ah >> 32:  al + ah*2^32 = a0 + (a1 + b1 + b0)*2^32 + b1*(2^64 - 2^32), and 2^64 = 2^32 - 1
// turns the last term into -b1; the single possible carry k out of the middle word is folded the same way.  No
// conditional fix-up is needed (tools/word_model.py · recombine_model).
__device__ __forceinline__ uint64_t recombine(double AL, double AH) {
    uint32_t lo, hi;
    asm("{\n\t.reg .u32 u, b1, T, g;\n\t"
        "sub.u32     b1, %5, 0x43300000;\n\t"
        "add.u32     u, %3, b1;\n\t"
        "sub.u32     u, u, 0x43300000;\n\t"
        "add.cc.u32  g, u, %4;\n\t"
        "addc.u32    T, b1, 0;\n\t"
        "addc.u32    g, g, 0;\n\t"
        "sub.cc.u32  %0, %2, T;\n\t"
        "subc.u32    %1, g, 0;\n\t"
        "}" : "=&r"(lo), "=&r"(hi)
            : "r"((uint32_t)__double2loint(AL)), "r"((uint32_t)__double2hiint(AL)), "r"((uint32_t)__double2loint(AH)),
              "r"((uint32_t)__double2hiint(AH)));
    return ((uint64_t)hi << 32) | lo;
}

// S-boxes on every lane, then s <- MDS * s + (constants of layer r): "any" in, "any" out
__device__ __forceinline__ void full_round(uint64_t (&s)[WIDTH], int r) {
    const double* __restrict__ k = POSEIDON_MDS_K[r];
    double lo[WIDTH], hi[WIDTH], AL[WIDTH], AH[WIDTH];
#pragma unroll
    for (int i = 0; i < WIDTH; i++) sbox7_limbs(s[i], lo[i], hi[i]);
    mds_limb(lo, k, AL);
    mds_limb(hi, k + WIDTH, AH);
#pragma unroll
    for (int i = 0; i < WIDTH; i++) s[i] = recombine(AL[i], AH[i]);
}

// ---- partial rounds: lanes 1..11 never leave the fp64 pipe ----------------------------------------------------------
// Only lane 0 goes through the S-box in a partial round, so lanes 1..11 stay as (lo, hi) limb doubles — exact, signed,
// un-reduced — from round 4 to round 25.  A layer multiplies magnitudes by <= 272, so limbs are renormalised every
// second round (2^32.1 -> 2^40.2 -> 2^48.3 < 2^53).  The constants of those rounds are pushed forward through the MDS
// so that only lane 0 receives one per round (tools/mds_model.py · partial_constants); lane 0's sums carry 2^52 plus an
// offset = 0 (mod p) that keeps them positive, and are read back from the mantissa as before.

// mds_limb with the only non-zero folded constants uu0 = uv0 = q, v0 = 2q (lane-0 constant 4q)
__device__ __forceinline__ void mds_limb_partial(const double (&s)[WIDTH], const double q, double (&o)[WIDTH]) {
    double sp[6], sm[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        sp[i] = __dadd_rn(s[i], s[i + 6]);
        sm[i] = __dsub_rn(s[i], s[i + 6]);
    }
    double a[3], b[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        a[i] = __dadd_rn(sp[i], sp[i + 3]);
        b[i] = __dsub_rn(sp[i], sp[i + 3]);
    }
    const double S16 = __dmul_rn(__dadd_rn(__dadd_rn(a[0], a[1]), a[2]), 16.0);
    double U[6], V[6];
    {
        double UU[3], UV[3];
        UU[0] = __fma_rn(a[2], 16.0, __dadd_rn(S16, q));
        UU[1] = __fma_rn(a[0], 16.0, S16);
        UU[2] = __fma_rn(a[1], 16.0, S16);
        UV[0] = __fma_rn(b[2], 8.0, __fma_rn(b[1], -2.0, __dsub_rn(q, b[0])));
        UV[1] = __fma_rn(b[2], -2.0, __fma_rn(b[0], -8.0, -b[1]));
        UV[2] = __fma_rn(b[1], -8.0, __fma_rn(b[0], 2.0, -b[2]));
#pragma unroll
        for (int j = 0; j < 3; j++) {
            U[j] = __dadd_rn(UU[j], UV[j]);
            U[j + 3] = __dsub_rn(UU[j], UV[j]);
        }
    }
    // negacyclic-6 rows, coefficient of sm[j] in V[n] is f[n-j] (j <= n) or -f[6+n-j]; f = [2,-4,16,1,-1,-1]
    V[0] = __fma_rn(sm[0], 2.0, __fma_rn(sm[4], -16.0, __fma_rn(sm[5], 4.0, __dadd_rn(__dadd_rn(sm[1], sm[2]), __fma_rn(q, 2.0, -sm[3])))));
    V[1] = __fma_rn(sm[0], -4.0, __fma_rn(sm[1], 2.0, __fma_rn(sm[5], -16.0, __dsub_rn(__dadd_rn(sm[2], sm[3]), sm[4]))));
    V[2] = __fma_rn(sm[0], 16.0, __fma_rn(sm[1], -4.0, __fma_rn(sm[2], 2.0, __dsub_rn(__dadd_rn(sm[3], sm[4]), sm[5]))));
    V[3] = __fma_rn(sm[1], 16.0, __fma_rn(sm[2], -4.0, __fma_rn(sm[3], 2.0, __dadd_rn(__dadd_rn(sm[0], sm[4]), sm[5]))));
    V[4] = __fma_rn(sm[2], 16.0, __fma_rn(sm[3], -4.0, __fma_rn(sm[4], 2.0, __dadd_rn(__dsub_rn(sm[1], sm[0]), sm[5]))));
    V[5] = __fma_rn(sm[3], 16.0, __fma_rn(sm[4], -4.0, __fma_rn(sm[5], 2.0, __dsub_rn(__dsub_rn(sm[2], sm[0]), sm[1]))));
    U[0] = __fma_rn(s[0], 4.0, U[0]);
    V[0] = __fma_rn(s[0], 4.0, V[0]);
#pragma unroll
    for (int n = 0; n < 6; n++) {
        o[n] = __dadd_rn(U[n], V[n]);
        o[n + 6] = __dsub_rn(U[n], V[n]);
    }
}

// (L, H) with value L + H*2^32 -> same value mod p with |L|, |H| <= 2^31 + 2^17  (inputs up to 2^49)
__device__ __forceinline__ void normalize_limbs(double& L, double& H) {
    constexpr double MAGIC = 6755399441055744.0;     // 1.5 * 2^52: adding it rounds to an integer
    constexpr double INV32 = 2.3283064365386963e-10;  // 2^-32
    constexpr double B32 = 4294967296.0;
    const double cL = __dsub_rn(__fma_rn(L, INV32, MAGIC), MAGIC);
    L = __fma_rn(cL, -B32, L);
    H = __dadd_rn(H, cL);
    const double cH = __dsub_rn(__fma_rn(H, INV32, MAGIC), MAGIC);   // cH * 2^64 = cH * (2^32 - 1)
    H = __fma_rn(cH, -(B32 - 1.0), H);                               // H - cH*2^32 + cH, one exact operation
    L = __dsub_rn(L, cH);
}

__device__ __forceinline__ void partial_rounds(uint64_t (&s)[WIDTH]) {
    double L[WIDTH], H[WIDTH];
#pragma unroll
    for (int i = 1; i < WIDTH; i++) {
        L[i] = limb_to_double((uint32_t)s[i]);
        H[i] = limb_to_double((uint32_t)(s[i] >> 32));
    }
    uint64_t x0 = s[0];
#pragma unroll 1
    for (int r = N_FULL_HALF; r < N_FULL_HALF + N_PARTIAL; r++) {
        sbox7_limbs(x0, L[0], H[0]);
        double OL[WIDTH], OH[WIDTH];
        mds_limb_partial(L, POSEIDON_PARTIAL_Q[r - N_FULL_HALF][0], OL);
        mds_limb_partial(H, POSEIDON_PARTIAL_Q[r - N_FULL_HALF][1], OH);
        x0 = recombine(OL[0], OH[0]);
#pragma unroll
        for (int i = 1; i < WIDTH; i++) {
            L[i] = OL[i];
            H[i] = OH[i];
        }
        if (r & 1) {
#pragma unroll
            for (int i = 1; i < WIDTH; i++) normalize_limbs(L[i], H[i]);
        }
    }
    s[0] = x0;
#pragma unroll
    for (int i = 1; i < WIDTH; i++)
        s[i] = recombine(__dadd_rn(L[i], POSEIDON_PARTIAL_TAIL[i - 1][0]), __dadd_rn(H[i], POSEIDON_PARTIAL_TAIL[i - 1][1]));
}

// Full permutation. in/out "any" -> "any" (callers canonicalise what they store).
// One copy of the full-round body serves both halves (the code must stay inside the 32 KB L1.5 instruction cache).
__device__ __forceinline__ void permute(uint64_t (&s)[WIDTH]) {
    // round-0 constants (all later constant layers are folded into the preceding MDS layer)
#pragma unroll
    for (int i = 0; i < WIDTH; i++) s[i] = gl::add_any_c(s[i], POSEIDON_RC[i]);
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
        const int r0 = half ? N_FULL_HALF + N_PARTIAL : 0;
#pragma unroll 1
        for (int r = r0; r < r0 + N_FULL_HALF; r++) full_round(s, r);
        if (half == 0) partial_rounds(s);
    }
}

}  // namespace poseidon
