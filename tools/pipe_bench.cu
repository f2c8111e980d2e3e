// pipe_bench.cu -- issue-rate microbenchmark for the integer ops the DP / sketch kernels are built from
// (which SASS pipe an op uses decides whether it competes with LOP3/SHF for the ALU pipe).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench tools/pipe_bench.cu ; run on the GPU box
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2048
template <int MODE>
__global__ void __launch_bounds__(256) bench(uint32_t *out, uint32_t a, uint32_t b) {
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 8 + i + a;
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
                if (MODE == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                if (MODE == 2) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                if (MODE == 3) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                if (MODE == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
                if (MODE == 5) { if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
                                 else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b)); }
                if (MODE == 6) { if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
                                 else asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b)); }
                if (MODE == 7) { if ((i & 3) == 0) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                                 else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b)); }
                if (MODE == 8) asm volatile("mul.wide.u32 %0, %1, %2;" : "+l"(*reinterpret_cast<unsigned long long *>(&x[i & 6])) : "r"(x[i]), "r"(a));
                if (MODE == 9) asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(x[(i + 1) & 7]));
                if (MODE == 11) asm volatile("{ .reg .u32 t; min.u32 t, %0, %1; min.u32 %0, t, %2; }" : "+r"(x[i]) : "r"(x[(i + 1) & 7]), "r"(x[(i + 2) & 7]));
                if (MODE == 12) asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %0, %1, p; }" : "+r"(x[i]) : "r"(x[(i + 1) & 7]));
                if (MODE == 10) { if (i & 1) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                                  else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b)); }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, uint32_t *d, int sms, int clock_khz) {
    const int blocks = sms * 8;
    bench<MODE><<<blocks, 256>>>(d, 3, 5);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<MODE><<<blocks, 256>>>(d, 3, 5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * 256 * ITER * 8;
    const double per_clk_sm = ops / (ms * 1e-3) / ((double)clock_khz * 1e3) / sms;
    printf("%-28s %8.3f ms  %7.1f thread-ops/clk/SM (at %d MHz nominal)\n", name, ms, per_clk_sm, clock_khz / 1000);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint32_t *d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    printf("%s, %d SMs, %d kHz\n", p.name, p.multiProcessorCount, khz);
    run<0>("LOP3", d, p.multiProcessorCount, khz);
    run<1>("IMAD (mad.lo)", d, p.multiProcessorCount, khz);
    run<2>("IMAD.HI (mad.hi)", d, p.multiProcessorCount, khz);
    run<3>("SHF (funnel)", d, p.multiProcessorCount, khz);
    run<4>("IADD (add.u32)", d, p.multiProcessorCount, khz);
    run<5>("LOP3 : IMAD 1:1", d, p.multiProcessorCount, khz);
    run<6>("LOP3 : IMAD.HI 1:1", d, p.multiProcessorCount, khz);
    run<7>("LOP3 : IMAD.HI 3:1", d, p.multiProcessorCount, khz);
    run<8>("IMAD.WIDE (mul.wide)", d, p.multiProcessorCount, khz);
    run<9>("VIMNMX (min.u32)", d, p.multiProcessorCount, khz);
    run<11>("VIMNMX3 (min of 3), per op", d, p.multiProcessorCount, khz);
    run<12>("ISETP + SEL (min by select)", d, p.multiProcessorCount, khz);
    run<10>("SHF : IMAD 1:1", d, p.multiProcessorCount, khz);
    return 0;
}
