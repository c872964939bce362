// Integer-pipe issue/throughput calibration for sm_100a (K7 extended).  Standalone: nvcc -o build/ubench tools/ubench.cu
// Each kernel: every thread runs CH independent chains; per-iteration op mix given by MODE.  Reports thread-instr/clk/SM
// from in-kernel clock64() and from CUDA events.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CH 8
enum { M_IMADW = 0, M_IMAD32, M_IADD3, M_LOP3, M_DFMA, M_IMADW_IADD3, M_IMADW_LOP3, M_IMAD32_IADD3, M_DFMA_IMADW, M_DFMA_IADD3,
       M_DFMA_IMADW_IADD3, M_IMADW_2IADD3, M_ADD64, M_IMADW_DEP_LOP3, M_SHF, M_IMADW_SHF, M_ISETP_SEL, M_N };
static const char* names[M_N] = {"imad.wide", "imad32", "iadd3", "lop3", "dfma", "imad.wide+iadd3", "imad.wide+lop3", "imad32+iadd3",
                                 "dfma+imad.wide", "dfma+iadd3", "dfma+imad.wide+iadd3", "imad.wide+2*iadd3", "add64(cc)",
                                 "imad.wide<-lop3 dep", "shf", "imad.wide+shf", "isetp+sel"};
static const int ops_per[M_N] = {1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 3, 3, 2, 2, 1, 2, 2};

template <int MODE>
__global__ void __launch_bounds__(256) k(uint64_t* out, long long* cyc, uint32_t iters, uint32_t b, uint32_t c) {
    uint64_t acc[CH];
    uint32_t x[CH], y[CH], z[CH], w[CH];
    double d[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { acc[i] = threadIdx.x + i; x[i] = threadIdx.x * 7 + i; y[i] = threadIdx.x * 3 + i; z[i] = threadIdx.x * 5 + i + c; w[i] = threadIdx.x * 11 + i + b; d[i] = 1.0 + threadIdx.x * 1e-9 + i; }
    double db = 1.0 + b * 1e-12, dc = c * 1e-12;
    long long t0 = clock64();
#pragma unroll 1
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                constexpr bool W = MODE == M_IMADW || MODE == M_IMADW_IADD3 || MODE == M_IMADW_LOP3 || MODE == M_DFMA_IMADW ||
                                   MODE == M_DFMA_IMADW_IADD3 || MODE == M_IMADW_2IADD3 || MODE == M_IMADW_SHF;
                constexpr bool A = MODE == M_IADD3 || MODE == M_IMADW_IADD3 || MODE == M_IMAD32_IADD3 || MODE == M_DFMA_IADD3 ||
                                   MODE == M_DFMA_IMADW_IADD3 || MODE == M_IMADW_2IADD3;
                // non-linear / two-variable recurrences so that ptxas cannot strength-reduce the chains
                if (W) asm volatile("{.reg .b32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, hi, %0;}" : "+l"(acc[i]));
                if (MODE == M_IMADW_DEP_LOP3) asm volatile("mad.wide.u32 %0, %1, %1, %0;" : "+l"(acc[i]) : "r"(x[i]));
                if (MODE == M_IMAD32 || MODE == M_IMAD32_IADD3) asm volatile("mad.lo.u32 %0, %0, %0, %1;" : "+r"(y[i]) : "r"(c));
                if (A) { if (r & 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(z[i])); else asm volatile("add.u32 %0, %0, %1;" : "+r"(z[i]) : "r"(x[i])); }
                if (MODE == M_IMADW_2IADD3) { if (r & 1) asm volatile("sub.u32 %0, %0, %1;" : "+r"(z[i]) : "r"(x[i])); else asm volatile("sub.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(z[i])); }
                if (MODE == M_LOP3 || MODE == M_IMADW_LOP3 || MODE == M_IMADW_DEP_LOP3)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(c));
                if (MODE == M_DFMA || MODE == M_DFMA_IMADW || MODE == M_DFMA_IADD3 || MODE == M_DFMA_IMADW_IADD3)
                    asm volatile("fma.rn.f64 %0, %0, %0, %1;" : "+d"(d[i]) : "d"(dc));
                if (MODE == M_ADD64) { if (r & 1) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(z[i]), "r"(w[i]));
                                       else asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(z[i]), "+r"(w[i]) : "r"(x[i]), "r"(y[i])); }
                if (MODE == M_SHF || MODE == M_IMADW_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));
                if (MODE == M_ISETP_SEL) asm volatile("{.reg .pred p; setp.ge.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(x[i]) : "r"(b), "r"(c));
            }
        }
    }
    long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s ^= acc[i] ^ x[i] ^ y[i] ^ z[i] ^ w[i] ^ (uint64_t)__double_as_longlong(d[i]);
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
    printf("%-24s warps/SM=%2d  instr/clk/SM(clock64)=%6.1f  Gops/s(event)=%8.1f  implied_MHz=%6.0f\n", names[MODE], bps * 8,
           thread_ops / avg, thread_ops * sms / (ms * 1e6), avg / (ms * 1e3));
}

int main(int argc, char** argv) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint64_t* out; long long* cyc;
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 8);
    cudaMalloc(&cyc, (size_t)sms * 8 * 8);
    uint32_t iters = argc > 1 ? atoi(argv[1]) : 4000;
    for (int bps : {2, 4, 8}) {
        run<M_IMADW>(sms, bps, iters, out, cyc); run<M_IMAD32>(sms, bps, iters, out, cyc); run<M_IADD3>(sms, bps, iters, out, cyc);
        run<M_LOP3>(sms, bps, iters, out, cyc); run<M_DFMA>(sms, bps, iters, out, cyc); run<M_IMADW_IADD3>(sms, bps, iters, out, cyc);
        run<M_IMADW_LOP3>(sms, bps, iters, out, cyc); run<M_IMAD32_IADD3>(sms, bps, iters, out, cyc);
        run<M_DFMA_IMADW>(sms, bps, iters, out, cyc); run<M_DFMA_IADD3>(sms, bps, iters, out, cyc);
        run<M_DFMA_IMADW_IADD3>(sms, bps, iters, out, cyc); run<M_IMADW_2IADD3>(sms, bps, iters, out, cyc);
        run<M_ADD64>(sms, bps, iters, out, cyc); run<M_IMADW_DEP_LOP3>(sms, bps, iters, out, cyc); run<M_SHF>(sms, bps, iters, out, cyc);
        run<M_IMADW_SHF>(sms, bps, iters, out, cyc); run<M_ISETP_SEL>(sms, bps, iters, out, cyc);
        printf("\n");
    }
    return 0;
}
