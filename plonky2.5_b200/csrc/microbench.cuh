// K7: integer-pipe calibration kernels (denominators for the integer roofline).
// Every thread runs 8 independent dependent-chains so the pipes, not latency, are what is measured.
#pragma once
#include "poseidon.cuh"

namespace microbench {

constexpr int CHAINS = 8;

// 0: IMAD.WIDE.U32 (fma pipe)
__global__ void imad_wide_kernel(uint64_t* out, uint32_t iters, uint32_t b) {
    uint64_t acc[CHAINS];
    uint32_t a[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) { acc[k] = threadIdx.x + k; a[k] = threadIdx.x * 7 + k; }
#pragma unroll 1
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < CHAINS; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a[k]), "r"(b));
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) s ^= acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 1: LOP3 (alu pipe)
__global__ void alu_kernel(uint64_t* out, uint32_t iters, uint32_t b) {
    uint32_t acc[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) acc[k] = threadIdx.x + k;
#pragma unroll 1
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < CHAINS; k++)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(acc[k]) : "r"(b), "r"(i));   // 3-input xor
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) s ^= acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 2: 1:1 mix of the two
__global__ void mixed_kernel(uint64_t* out, uint32_t iters, uint32_t b) {
    uint64_t acc[CHAINS];
    uint32_t x[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) { acc[k] = threadIdx.x + k; x[k] = threadIdx.x * 3 + k; }
#pragma unroll 1
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < CHAINS; k++) {
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x[k]), "r"(b));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(b), "r"(i));
            }
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) s ^= acc[k] ^ x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 3: Goldilocks modmul chains
__global__ void modmul_kernel(uint64_t* out, uint32_t iters, uint64_t y) {
    uint64_t x[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) x[k] = 0x9E3779B97F4A7C15ULL * (threadIdx.x + 1 + k);
#pragma unroll 1
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < CHAINS; k++) x[k] = gl::mul(x[k], y);
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) s ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 4: back-to-back Poseidon permutations, one state per thread
__global__ void poseidon_kernel(uint64_t* out, uint32_t iters) {
    uint64_t s[poseidon::WIDTH];
#pragma unroll
    for (int k = 0; k < poseidon::WIDTH; k++) s[k] = 0x9E3779B97F4A7C15ULL * (blockIdx.x * blockDim.x + threadIdx.x + 1 + k);
#pragma unroll 1
    for (uint32_t i = 0; i < iters; i++) poseidon::permute(s);
    uint64_t x = 0;
#pragma unroll
    for (int k = 0; k < poseidon::WIDTH; k++) x ^= s[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

}  // namespace microbench
