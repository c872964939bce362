// Is a 32x32->64 product cheaper as IMAD (lo) + IMAD.HI than as one IMAD.WIDE.U32 on sm_100a?  (IMAD.WIDE occupies
// the dispatch port 3.56 clk per warp, IMAD 0.78.)  Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench3 tools/ubench3.cu
// Every thread runs CH independent two-variable chains (ptxas cannot strength-reduce them); reports lane-ops per clock per SM from
// clock64() and Gops/s from CUDA events, like tools/ubench.cu.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CH 8
enum { M_WIDE = 0, M_HI, M_LO, M_LO_HI, M_HI_IADD3, M_WIDE_IADD3, M_N };
static const char* names[M_N] = {"imad.wide.u32", "imad.hi.u32", "imad (lo)", "imad lo + imad.hi", "imad.hi + iadd3", "imad.wide + iadd3"};
static const int ops_per[M_N] = {1, 1, 1, 2, 2, 2};

template <int MODE>
__global__ void __launch_bounds__(256) k(uint64_t* out, long long* cyc, uint32_t iters, uint32_t b, uint32_t c) {
    uint64_t acc[CH];
    uint32_t x[CH], y[CH], z[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { acc[i] = threadIdx.x + i + b; x[i] = threadIdx.x * 7 + i + c; y[i] = threadIdx.x * 3 + i + b; z[i] = threadIdx.x * 5 + i + c; }
    long long t0 = clock64();
#pragma unroll 1
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                if (MODE == M_WIDE || MODE == M_WIDE_IADD3) asm volatile("{.reg .b32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, hi, %0;}" : "+l"(acc[i]));
                if (MODE == M_HI || MODE == M_HI_IADD3) { if (r & 1) asm volatile("mad.hi.u32 %0, %0, %1, %1;" : "+r"(x[i]) : "r"(y[i])); else asm volatile("mad.hi.u32 %0, %0, %1, %1;" : "+r"(y[i]) : "r"(x[i])); }
                if (MODE == M_LO) { if (r & 1) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x[i]) : "r"(y[i])); else asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(y[i]) : "r"(x[i])); }
                if (MODE == M_LO_HI) {   // both halves of x*y, as a multi-word product would need them
                    uint32_t lo, hi;
                    asm volatile("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(x[i]), "r"(y[i]));
                    x[i] = lo | 1; y[i] = hi | 3;
                }
                if (MODE == M_HI_IADD3 || MODE == M_WIDE_IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(z[i]) : "r"(c));
            }
        }
    }
    long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s ^= acc[i] ^ x[i] ^ y[i] ^ z[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(int sms, int bps, uint32_t iters, uint64_t* out, long long* cyc) {
    int blocks = sms * bps;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(out, cyc, iters, 0x9E3779B9u, 12345u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    long long* h = (long long*)malloc(blocks * sizeof(long long));
    cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    free(h);
    double thread_ops = (double)ops_per[MODE] * 4 * CH * iters * 256.0 * bps;   // per SM
    printf("%-20s warps/SM=%2d  lane-ops/clk/SM(clock64)=%6.1f  Gops/s(event)=%8.1f\n", names[MODE], bps * 8, thread_ops / avg, thread_ops * sms / (ms * 1e6));
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount;
    uint64_t* out; long long* cyc;
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 8); cudaMalloc(&cyc, (size_t)sms * 8 * 8);
    printf("%s, %d SMs\n", pr.name, sms);
    for (int bps : {2, 4, 8}) {
        run<M_WIDE>(sms, bps, 4000, out, cyc); run<M_HI>(sms, bps, 4000, out, cyc); run<M_LO>(sms, bps, 4000, out, cyc);
        run<M_LO_HI>(sms, bps, 4000, out, cyc); run<M_HI_IADD3>(sms, bps, 4000, out, cyc); run<M_WIDE_IADD3>(sms, bps, 4000, out, cyc);
        printf("\n");
    }
    return cudaDeviceSynchronize() != cudaSuccess;
}
