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

// Per-product choice of the 128-bit product form inside the S-box (bit 0: x^2, bit 1: x^4, bit 2: x^3, bit 3: x^3*x^4):
// 0 = PTX carry chains (3 IMAD.WIDE for a square, 6 alu adds), 1 = nvcc's 128-bit multiply (4 IMAD.WIDE, 1 alu add).
// The full rounds are alu-pipe-bound with form 0 everywhere and fma-pipe-bound with form 1 everywhere.
#ifndef POSEIDON_SBOX_FORM
#define POSEIDON_SBOX_FORM 0x8
#endif
template <int NEW>
__device__ __forceinline__ uint64_t sbox_sqr(uint64_t x) {
    if (NEW) {
        uint32_t z0, z1, z2, z3;
        gl::mul_words_c(x, x, z0, z1, z2, z3);
        return gl::reduce_words(z0, z1, z2, z3);
    }
    return gl::sqr(x);
}
template <int NEW>
__device__ __forceinline__ void sbox_mul_words(uint64_t a, uint64_t b, uint32_t& z0, uint32_t& z1, uint32_t& z2, uint32_t& z3) {
    if (NEW) gl::mul_words_c(a, b, z0, z1, z2, z3);
    else gl::mul_words(a, b, z0, z1, z2, z3);
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
    const uint64_t x2 = sbox_sqr<(POSEIDON_SBOX_FORM >> 0) & 1>(x);
    const uint64_t x4 = sbox_sqr<(POSEIDON_SBOX_FORM >> 1) & 1>(x2);
    uint32_t z0, z1, z2, z3;
    sbox_mul_words<(POSEIDON_SBOX_FORM >> 2) & 1>(x, x2, z0, z1, z2, z3);
    const uint64_t x3 = gl::reduce_words(z0, z1, z2, z3);
    sbox_mul_words<(POSEIDON_SBOX_FORM >> 3) & 1>(x3, x4, z0, z1, z2, z3);
#ifdef POSEIDON_HANDOFF_FP64
    const double d2 = biased(z2);
    lo = __dsub_rn(__dsub_rn(biased(z0), d2), __dsub_rn(biased(z3), TWO52));
    hi = __dsub_rn(__dadd_rn(biased(z1), __dsub_rn(d2, TWO52)), TWO52);
#else
    // z0 - z2 - z3 + 2^33 and z1 + z2 are formed on the integer pipe directly in the mantissa of 2^52 + v (the borrows /
    // the carry run into the high word 0x43300002 / 0x43300000), so each limb costs one DADD to remove the bias
    uint32_t ll, lh, hl, hh;
    asm("{\n\t"
        "sub.cc.u32  %0, %4, %6;\n\t"
        "subc.u32    %1, 0x43300002, 0;\n\t"
        "sub.cc.u32  %0, %0, %7;\n\t"
        "subc.u32    %1, %1, 0;\n\t"
        "add.cc.u32  %2, %5, %6;\n\t"
        "addc.u32    %3, 0x43300000, 0;\n\t"
        "}" : "=&r"(ll), "=&r"(lh), "=&r"(hl), "=&r"(hh) : "r"(z0), "r"(z1), "r"(z2), "r"(z3));
    lo = __dsub_rn(__hiloint2double((int)lh, (int)ll), TWO52 + 8589934592.0);
    hi = __dsub_rn(__hiloint2double((int)hh, (int)hl), TWO52);
#endif
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
// un-reduced — from round 4 to round 25.  The constants of those rounds are pushed forward through the MDS so that only
// lane 0 receives one per round (tools/mds_model.py · partial_constants); lane 0's read-out carries 2^52 plus an offset
// = 0 (mod p) that keeps it positive, and is read back from the mantissa.  (The first version of this idea kept the lanes
// in the time domain, 89 fp64 operations per limb and round: tools/mds_model.py · permute_v3.)

#if defined(POSEIDON_PARTIAL_V4)
// ---- v4: the 11 resident lanes live in the CRT domain of the circulant ------------------------------------------------
// With a_i, b_i, m_i as in mds_limb (cyclic-3, negacyclic-3, negacyclic-6 parts of one limb vector) the MDS layer acts
// blockwise, a' = 4*UU(a), b' = 4*UV(b), m' = 2*V(m): the 30 add/sub butterflies of the time-domain layer disappear
// when the state stays in this domain for all 22 partial rounds.  Lane 0 is read out as (a0 + b0 + 2 m0)/4 (an integer
// identity) and written back by adding (y - lane0) to a0, b0, m0; the diagonal 8*s0 reaches a0', b0', m0'; the round
// constant is only needed at the read-out.  Re-normalisation carries are multiples of 4, so the division by 4 stays
// exact on each limb.  63 fp64 operations per limb and round instead of 89, and everything that does not depend on the
// S-box output is off the dependency chain.  tools/mds_model.py · permute_v4 mirrors this operation by operation and
// proves (by bound propagation) that no intermediate reaches 2^53.
struct Freq {
    double a[3], b[3], m[6];
};

// time -> CRT domain of lanes 1..11 (lane 0 enters as 0: it is replaced by the S-box output in the first round)
__device__ __forceinline__ void freq_forward(const double (&s)[WIDTH], Freq& f) {
    double sp[6];
    sp[0] = s[6];
    f.m[0] = -s[6];
#pragma unroll
    for (int i = 1; i < 6; i++) {
        sp[i] = __dadd_rn(s[i], s[i + 6]);
        f.m[i] = __dsub_rn(s[i], s[i + 6]);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        f.a[i] = __dadd_rn(sp[i], sp[i + 3]);
        f.b[i] = __dsub_rn(sp[i], sp[i + 3]);
    }
}

// lane 0 (raw value e) <- y, then the MDS layer.  The component that carries y is the last link of every chain.
__device__ __forceinline__ void freq_round(Freq& f, const double y, const double e) {
    const double d = __dsub_rn(y, e);
    const double a0 = __dadd_rn(f.a[0], d), b0 = __dadd_rn(f.b[0], d), m0 = __dadd_rn(f.m[0], d);
    const double t1 = __dmul_rn(__dadd_rn(f.a[1], f.a[2]), 64.0);
    const double na0 = __fma_rn(y, 8.0, __fma_rn(a0, 64.0, __fma_rn(f.a[2], 64.0, t1)));
    const double na1 = __fma_rn(a0, 128.0, t1);
    const double na2 = __fma_rn(a0, 64.0, __fma_rn(f.a[1], 64.0, t1));
    const double nb0 = __fma_rn(y, 8.0, __fma_rn(b0, -4.0, __fma_rn(f.b[2], 32.0, __dmul_rn(f.b[1], -8.0))));
    const double nb1 = __fma_rn(b0, -32.0, __fma_rn(f.b[1], -4.0, __dmul_rn(f.b[2], -8.0)));
    const double nb2 = __fma_rn(b0, 8.0, __fma_rn(f.b[1], -32.0, __dmul_rn(f.b[2], -4.0)));
    // negacyclic-6 with 2*[2, -4, 16, 1, -1, -1]: coefficient of m[j] in m'[n] is 2*f[n-j] (j <= n) or -2*f[6+n-j]
    constexpr double g[6] = {4.0, -8.0, 32.0, 2.0, -2.0, -2.0};
    double nm[6];
#pragma unroll
    for (int n = 0; n < 6; n++) {
        double acc = __dmul_rn(f.m[1], (1 <= n) ? g[n - 1] : -g[5 + n]);
#pragma unroll
        for (int j = 2; j < 6; j++) acc = __fma_rn(f.m[j], (j <= n) ? g[n - j] : -g[6 + n - j], acc);
        nm[n] = __fma_rn(m0, g[n], acc);
    }
    nm[0] = __fma_rn(y, 8.0, nm[0]);
    f.a[0] = na0; f.a[1] = na1; f.a[2] = na2;
    f.b[0] = nb0; f.b[1] = nb1; f.b[2] = nb2;
#pragma unroll
    for (int n = 0; n < 6; n++) f.m[n] = nm[n];
}

// raw lane-0 value (a0 + b0 + 2 m0)/4; 0.5*m0 is an exact half-integer, the result an exact integer
__device__ __forceinline__ double freq_lane0(const Freq& f) {
    return __fma_rn(__dadd_rn(f.a[0], f.b[0]), 0.25, __dmul_rn(f.m[0], 0.5));
}

// (L, H) with value L + H*2^32 -> same value mod p with |L|, |H| <= 2^33 + 2^20, carries multiples of 4 (inputs < 2^52)
__device__ __forceinline__ void normalize_limbs4(double& L, double& H) {
    constexpr double MAGIC4 = 27021597764222976.0;    // 1.5 * 2^54: adding it rounds to a multiple of 4
    constexpr double INV32 = 2.3283064365386963e-10;  // 2^-32
    constexpr double B32 = 4294967296.0;
    const double cL = __dsub_rn(__fma_rn(L, INV32, MAGIC4), MAGIC4);
    L = __fma_rn(cL, -B32, L);
    H = __dadd_rn(H, cL);
    const double cH = __dsub_rn(__fma_rn(H, INV32, MAGIC4), MAGIC4);   // cH * 2^64 = cH * (2^32 - 1)
    H = __fma_rn(cH, -(B32 - 1.0), H);
    L = __dsub_rn(L, cH);
}

__device__ __forceinline__ void partial_rounds(uint64_t (&s)[WIDTH]) {
    Freq FL, FH;
    {
        double l[WIDTH], h[WIDTH];
        l[0] = h[0] = 0.0;
#pragma unroll
        for (int i = 1; i < WIDTH; i++) {
            l[i] = limb_to_double((uint32_t)s[i]);
            h[i] = limb_to_double((uint32_t)(s[i] >> 32));
        }
        freq_forward(l, FL);
        freq_forward(h, FH);
    }
    uint64_t x0 = s[0];
    double eL = 0.0, eH = 0.0;
#pragma unroll 1
    for (int r = 0; r < N_PARTIAL; r++) {
        double yL, yH;
        sbox7_limbs(x0, yL, yH);
        freq_round(FL, yL, eL);
        freq_round(FH, yH, eH);
        if (r & 1) {   // rounds 5, 7, .., 25
#pragma unroll
            for (int i = 0; i < 3; i++) {
                normalize_limbs4(FL.a[i], FH.a[i]);
                normalize_limbs4(FL.b[i], FH.b[i]);
            }
#pragma unroll
            for (int i = 0; i < 6; i++) normalize_limbs4(FL.m[i], FH.m[i]);
        }
        eL = freq_lane0(FL);
        eH = freq_lane0(FH);
        x0 = recombine(__dadd_rn(eL, POSEIDON_PARTIAL4_Q[r][0]), __dadd_rn(eH, POSEIDON_PARTIAL4_Q[r][1]));
    }
    s[0] = x0;
    // back to the time domain: s_i = (a_i + b_i)/4 + m_i/2, s_{i+6} = (a_i + b_i)/4 - m_i/2, s_{i+3}, s_{i+9} with a_i - b_i
    double l[WIDTH], h[WIDTH];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double tpl = __dadd_rn(FL.a[i], FL.b[i]), tml = __dsub_rn(FL.a[i], FL.b[i]);
        const double tph = __dadd_rn(FH.a[i], FH.b[i]), tmh = __dsub_rn(FH.a[i], FH.b[i]);
        const double ml = __dmul_rn(FL.m[i], 0.5), ml3 = __dmul_rn(FL.m[i + 3], 0.5);
        const double mh = __dmul_rn(FH.m[i], 0.5), mh3 = __dmul_rn(FH.m[i + 3], 0.5);
        l[i] = __fma_rn(tpl, 0.25, ml);      l[i + 6] = __fma_rn(tpl, 0.25, -ml);
        l[i + 3] = __fma_rn(tml, 0.25, ml3); l[i + 9] = __fma_rn(tml, 0.25, -ml3);
        h[i] = __fma_rn(tph, 0.25, mh);      h[i + 6] = __fma_rn(tph, 0.25, -mh);
        h[i + 3] = __fma_rn(tmh, 0.25, mh3); h[i + 9] = __fma_rn(tmh, 0.25, -mh3);
    }
#pragma unroll
    for (int i = 1; i < WIDTH; i++)
        s[i] = recombine(__dadd_rn(l[i], POSEIDON_PARTIAL4_TAIL[i - 1][0]), __dadd_rn(h[i], POSEIDON_PARTIAL4_TAIL[i - 1][1]));
}
#else
// ---- v5 = v4 with the negacyclic-6 part split once more: x^6+1 = (x^2+1)(x^4-x^2+1) ------------------------------------
// As in v4 the 11 resident lanes live in the CRT domain of the circulant (a: cyclic-3, b: negacyclic-3; the MDS layer acts
// blockwise, lane 0 is read out as (a0 + b0 + 2 m0)/4 and written back by adding (y - lane0) to the components that
// contain it, the diagonal 8*s0 reaches the same components, carries are multiples of 4).  The negacyclic-6 part m is
// kept as p = m mod (x^2+1) and q = m mod (x^4-x^2+1): m' = m*G becomes p' = p*(-30-12x) (4 products) and
// q' = q*(6-6x+30x^2) (13 products) instead of 36.  The back-map divides by 3 (3 m0 = p0 + 2 q0 + q2): t/6 is rounded
// to the nearest half-integer with a magic constant, which is exact because t is a multiple of 3; p and q carry in
// multiples of 12 so that every limb stays divisible.  49 fp64 operations per limb and round (v4: 63, v3: 89).
// tools/mds_model.py · permute_v5 mirrors this operation by operation and proves the 2^53 bounds.
struct Freq5 {
    double a[3], b[3], p[2], q[4];
};

__device__ __forceinline__ void freq5_forward(const double (&s)[WIDTH], Freq5& f) {
    double sp[6], m[6];
    sp[0] = s[6];
    m[0] = -s[6];
#pragma unroll
    for (int i = 1; i < 6; i++) {
        sp[i] = __dadd_rn(s[i], s[i + 6]);
        m[i] = __dsub_rn(s[i], s[i + 6]);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        f.a[i] = __dadd_rn(sp[i], sp[i + 3]);
        f.b[i] = __dsub_rn(sp[i], sp[i + 3]);
    }
    f.p[0] = __dadd_rn(__dsub_rn(m[0], m[2]), m[4]);
    f.p[1] = __dadd_rn(__dsub_rn(m[1], m[3]), m[5]);
    f.q[0] = __dsub_rn(m[0], m[4]);
    f.q[1] = __dsub_rn(m[1], m[5]);
    f.q[2] = __dadd_rn(m[2], m[4]);
    f.q[3] = __dadd_rn(m[3], m[5]);
}

// lane 0 (raw value e) <- y, then the MDS layer.  The component that carries y is the last link of every chain.
__device__ __forceinline__ void freq5_round(Freq5& f, const double y, const double e) {
    const double d = __dsub_rn(y, e);
    const double a0 = __dadd_rn(f.a[0], d), b0 = __dadd_rn(f.b[0], d), p0 = __dadd_rn(f.p[0], d), q0 = __dadd_rn(f.q[0], d);
    const double t1 = __dmul_rn(__dadd_rn(f.a[1], f.a[2]), 64.0);
    const double na0 = __fma_rn(y, 8.0, __fma_rn(a0, 64.0, __fma_rn(f.a[2], 64.0, t1)));
    const double na1 = __fma_rn(a0, 128.0, t1);
    const double na2 = __fma_rn(a0, 64.0, __fma_rn(f.a[1], 64.0, t1));
    const double nb0 = __fma_rn(y, 8.0, __fma_rn(b0, -4.0, __fma_rn(f.b[2], 32.0, __dmul_rn(f.b[1], -8.0))));
    const double nb1 = __fma_rn(b0, -32.0, __fma_rn(f.b[1], -4.0, __dmul_rn(f.b[2], -8.0)));
    const double nb2 = __fma_rn(b0, 8.0, __fma_rn(f.b[1], -32.0, __dmul_rn(f.b[2], -4.0)));
    const double np0 = __fma_rn(y, 8.0, __fma_rn(p0, -30.0, __dmul_rn(f.p[1], 12.0)));
    const double np1 = __fma_rn(p0, -12.0, __dmul_rn(f.p[1], -30.0));
    const double nq0 = __fma_rn(y, 8.0, __fma_rn(q0, 6.0, __fma_rn(f.q[2], -30.0, __dmul_rn(f.q[3], 6.0))));
    const double nq1 = __fma_rn(q0, -6.0, __fma_rn(f.q[1], 6.0, __dmul_rn(f.q[3], -30.0)));
    const double nq2 = __fma_rn(q0, 30.0, __fma_rn(f.q[1], -6.0, __fma_rn(f.q[2], 36.0, __dmul_rn(f.q[3], -6.0))));
    const double nq3 = __fma_rn(f.q[1], 30.0, __fma_rn(f.q[2], -6.0, __dmul_rn(f.q[3], 36.0)));
    f.a[0] = na0; f.a[1] = na1; f.a[2] = na2;
    f.b[0] = nb0; f.b[1] = nb1; f.b[2] = nb2;
    f.p[0] = np0; f.p[1] = np1;
    f.q[0] = nq0; f.q[1] = nq1; f.q[2] = nq2; f.q[3] = nq3;
}

constexpr double MAGIC_HALF = 3377699720527872.0;   // 1.5 * 2^51: adding it rounds to a multiple of 1/2
constexpr double INV6 = 1.0 / 6.0;

// t = 3 m  ->  m/2 (a half-integer), exact for |t| < 2^52
__device__ __forceinline__ double third_half(double t) { return __dsub_rn(__fma_rn(t, INV6, MAGIC_HALF), MAGIC_HALF); }

// raw lane-0 value (a0 + b0 + 2 m0)/4 with 3 m0 = p0 + 2 q0 + q2
__device__ __forceinline__ double freq5_lane0(const Freq5& f) {
    const double t = __fma_rn(f.q[0], 2.0, __dadd_rn(f.p[0], f.q[2]));
    return __fma_rn(__dadd_rn(f.a[0], f.b[0]), 0.25, third_half(t));
}

// (L, H) with value L + H*2^32 -> same value mod p, carries multiples of 4 / of 12 (inputs < 2^52)
__device__ __forceinline__ void normalize_limbs4(double& L, double& H) {
    constexpr double MAGIC4 = 27021597764222976.0;    // 1.5 * 2^54: adding it rounds to a multiple of 4
    constexpr double INV32 = 2.3283064365386963e-10;  // 2^-32
    constexpr double B32 = 4294967296.0;
    const double cL = __dsub_rn(__fma_rn(L, INV32, MAGIC4), MAGIC4);
    L = __fma_rn(cL, -B32, L);
    H = __dadd_rn(H, cL);
    const double cH = __dsub_rn(__fma_rn(H, INV32, MAGIC4), MAGIC4);   // cH * 2^64 = cH * (2^32 - 1)
    H = __fma_rn(cH, -(B32 - 1.0), H);
    L = __dsub_rn(L, cH);
}
__device__ __forceinline__ void normalize_limbs12(double& L, double& H) {
    constexpr double MAGIC = 6755399441055744.0;              // 1.5 * 2^52: adding it rounds to an integer
    constexpr double C12 = 1.0 / (12.0 * 4294967296.0);        // any integer k near L / (12 * 2^32) is a valid carry count
    constexpr double B32 = 4294967296.0;
    const double kL = __dsub_rn(__fma_rn(L, C12, MAGIC), MAGIC);
    L = __fma_rn(kL, -12.0 * B32, L);
    H = __fma_rn(kL, 12.0, H);
    const double kH = __dsub_rn(__fma_rn(H, C12, MAGIC), MAGIC);
    H = __fma_rn(kH, -12.0 * (B32 - 1.0), H);
    L = __fma_rn(kH, -12.0, L);
}

__device__ __forceinline__ void partial_rounds(uint64_t (&s)[WIDTH]) {
    Freq5 FL, FH;
    {
        double l[WIDTH], h[WIDTH];
        l[0] = h[0] = 0.0;
#pragma unroll
        for (int i = 1; i < WIDTH; i++) {
            l[i] = limb_to_double((uint32_t)s[i]);
            h[i] = limb_to_double((uint32_t)(s[i] >> 32));
        }
        freq5_forward(l, FL);
        freq5_forward(h, FH);
    }
    uint64_t x0 = s[0];
    double eL = 0.0, eH = 0.0;
#pragma unroll 1
    for (int r = 0; r < N_PARTIAL; r++) {
        double yL, yH;
        sbox7_limbs(x0, yL, yH);
        freq5_round(FL, yL, eL);
        freq5_round(FH, yH, eH);
        if (r & 1) {   // rounds 5, 7, .., 25
#pragma unroll
            for (int i = 0; i < 3; i++) {
                normalize_limbs4(FL.a[i], FH.a[i]);
                normalize_limbs4(FL.b[i], FH.b[i]);
            }
            normalize_limbs12(FL.p[0], FH.p[0]);
            normalize_limbs12(FL.p[1], FH.p[1]);
#pragma unroll
            for (int i = 0; i < 4; i++) normalize_limbs12(FL.q[i], FH.q[i]);
        }
        eL = freq5_lane0(FL);
        eH = freq5_lane0(FH);
        x0 = recombine(__dadd_rn(eL, POSEIDON_PARTIAL4_Q[r][0]), __dadd_rn(eH, POSEIDON_PARTIAL4_Q[r][1]));
    }
    s[0] = x0;
    // back to the time domain: m_j/2 from the CRT back-map, then s_i = (a_i + b_i)/4 + m_i/2, s_{i+6} = (a_i + b_i)/4 - m_i/2,
    // s_{i+3}, s_{i+9} with a_i - b_i and m_{i+3}
    double l[WIDTH], h[WIDTH];
    {
        const Freq5* F[2] = {&FL, &FH};
        double* out[2] = {l, h};
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const Freq5& f = *F[k];
            double mh[6];
            mh[0] = third_half(__fma_rn(f.q[0], 2.0, __dadd_rn(f.p[0], f.q[2])));
            mh[1] = third_half(__fma_rn(f.q[1], 2.0, __dadd_rn(f.p[1], f.q[3])));
            mh[2] = third_half(__fma_rn(f.q[2], 2.0, __dsub_rn(f.q[0], f.p[0])));
            mh[3] = third_half(__fma_rn(f.q[3], 2.0, __dsub_rn(f.q[1], f.p[1])));
            mh[4] = third_half(__dadd_rn(__dsub_rn(f.p[0], f.q[0]), f.q[2]));
            mh[5] = third_half(__dadd_rn(__dsub_rn(f.p[1], f.q[1]), f.q[3]));
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const double tp = __dadd_rn(f.a[i], f.b[i]), tm = __dsub_rn(f.a[i], f.b[i]);
                out[k][i] = __fma_rn(tp, 0.25, mh[i]);
                out[k][i + 6] = __fma_rn(tp, 0.25, -mh[i]);
                out[k][i + 3] = __fma_rn(tm, 0.25, mh[i + 3]);
                out[k][i + 9] = __fma_rn(tm, 0.25, -mh[i + 3]);
            }
        }
    }
#pragma unroll
    for (int i = 1; i < WIDTH; i++)
        s[i] = recombine(__dadd_rn(l[i], POSEIDON_PARTIAL4_TAIL[i - 1][0]), __dadd_rn(h[i], POSEIDON_PARTIAL4_TAIL[i - 1][1]));
}
#endif

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
