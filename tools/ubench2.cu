// Issue-model calibration #2 (sm_100a): how fp64, alu and fma-heavy instructions share the issue port when their
// register operands look like the Poseidon kernel's (3 distinct 64-bit sources, carry chains), at 4..6 warps per SMSP.
// Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -o build/ubench2 tools/ubench2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CH 6
enum { DFMA_RRR = 0, DFMA_RIR, DADD_RR, LOP3, CARRY, WIDE, DFMA_LOP3, DFMA_CARRY, DFMA_WIDE, CARRY_WIDE, MIX_FULL, MIX_PART, DADD_CARRY, DFMA_RRR_LOP3, M_N };
static const char* names[M_N] = {"dfma rrr", "dfma r,imm,r", "dadd rr", "lop3", "add.cc/addc.cc", "imad.wide", "dfma(imm)+lop3", "dfma(imm)+carry",
                                 "dfma(imm)+wide", "carry+wide", "mix full-round (2 dadd:5 carry:2 wide)", "mix partial (6 dfma:1 carry)", "dadd+carry", "dfma rrr+lop3"};
// instructions per inner step for each mode: {fp64, alu, wide}
static const int cnt[M_N][3] = {{1,0,0},{1,0,0},{1,0,0},{0,1,0},{0,2,0},{0,0,1},{1,1,0},{1,2,0},{1,0,1},{0,2,1},{2,5,2},{6,2,0},{1,2,0},{1,1,0}};

template <int MODE>
__global__ void __launch_bounds__(128) k(uint64_t* out, long long* cyc, uint32_t iters, uint32_t b, uint32_t c) {
    double d[CH], e[CH], f[CH];
    uint32_t x[CH], y[CH], z[CH], w[CH];
    uint64_t acc[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) {
        d[i] = 1.0 + threadIdx.x * 1e-9 + i; e[i] = 1.0 + 1e-12 * (b + i); f[i] = 1e-13 * (c + i);
        x[i] = threadIdx.x * 7 + i; y[i] = threadIdx.x * 3 + i; z[i] = threadIdx.x * 5 + i + c; w[i] = threadIdx.x * 11 + i + b;
        acc[i] = threadIdx.x + i;
    }
    long long t0 = clock64();
#pragma unroll 1
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                auto dfma_rrr = [&]() { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e[i]), "d"(f[i])); };
                auto dfma_rir = [&]() { asm volatile("fma.rn.f64 %0, %0, 0d3FF0000000000001, %1;" : "+d"(d[i]) : "d"(f[i])); };
                auto dadd = [&]() { asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(f[i])); };
                auto lop = [&]() { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(z[i])); };
                auto carry = [&]() { asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.cc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(z[i]), "r"(w[i])); };
                auto wide = [&]() { asm volatile("{.reg .b32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(acc[i]) : "r"(w[i])); };
                if (MODE == DFMA_RRR) dfma_rrr();
                if (MODE == DFMA_RIR) dfma_rir();
                if (MODE == DADD_RR) dadd();
                if (MODE == LOP3) lop();
                if (MODE == CARRY) carry();
                if (MODE == WIDE) wide();
                if (MODE == DFMA_LOP3) { dfma_rir(); lop(); }
                if (MODE == DFMA_RRR_LOP3) { dfma_rrr(); lop(); }
                if (MODE == DFMA_CARRY) { dfma_rir(); carry(); }
                if (MODE == DADD_CARRY) { dadd(); carry(); }
                if (MODE == DFMA_WIDE) { dfma_rir(); wide(); }
                if (MODE == CARRY_WIDE) { carry(); wide(); }
                if (MODE == MIX_FULL) { wide(); carry(); dadd(); carry(); lop(); wide(); dadd(); }
                if (MODE == MIX_PART) { dfma_rir(); dfma_rir(); dadd(); carry(); dfma_rir(); dfma_rir(); dadd(); }
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
        k<MODE><<<blocks, 128>>>(out, cyc, iters, 0x9E3779B9u, 12345u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    const double steps = 4.0 * CH * iters;                           // inner steps per thread
    const double warps_per_smsp = bps * 4 / 4.0;                      // 128-thread blocks: one warp per SMSP each
    const double cyc_total = ms * 1e-3 * 1.965e9;                     // assumes the max SM clock (check the clocks line)
    const double steps_per_smsp = steps * warps_per_smsp;             // warp-steps each SMSP executed
    const int* c = cnt[MODE];
    const double ipc = steps_per_smsp * (c[0] + c[1] + c[2]) / cyc_total;
    printf("%-40s warps/SMSP=%d  IPC/SMSP=%.3f  (fp64 %.3f alu %.3f wide %.3f)  cycles per step=%.2f\n", names[MODE], (int)warps_per_smsp, ipc,
           steps_per_smsp * c[0] / cyc_total, steps_per_smsp * c[1] / cyc_total, steps_per_smsp * c[2] / cyc_total, cyc_total / steps_per_smsp);
}

int main(int argc, char** argv) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint64_t* out; long long* cyc;
    cudaMalloc(&out, (size_t)sms * 8 * 128 * 8);
    cudaMalloc(&cyc, (size_t)sms * 8 * 8);
    uint32_t iters = argc > 1 ? atoi(argv[1]) : 20000;
    for (int bps : {5, 8}) {
        run<DFMA_RRR>(sms, bps, iters, out, cyc); run<DFMA_RIR>(sms, bps, iters, out, cyc); run<DADD_RR>(sms, bps, iters, out, cyc);
        run<LOP3>(sms, bps, iters, out, cyc); run<CARRY>(sms, bps, iters, out, cyc); run<WIDE>(sms, bps, iters, out, cyc);
        run<DFMA_LOP3>(sms, bps, iters, out, cyc); run<DFMA_RRR_LOP3>(sms, bps, iters, out, cyc); run<DFMA_CARRY>(sms, bps, iters, out, cyc);
        run<DADD_CARRY>(sms, bps, iters, out, cyc); run<DFMA_WIDE>(sms, bps, iters, out, cyc); run<CARRY_WIDE>(sms, bps, iters, out, cyc);
        run<MIX_FULL>(sms, bps, iters, out, cyc); run<MIX_PART>(sms, bps, iters, out, cyc);
        printf("\n");
    }
    return 0;
}
