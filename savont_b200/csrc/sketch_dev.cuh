// sketch_dev.cuh -- device helpers shared by the sketch kernels (sketch.cu: generic k/c; sketch_fast.cu: k=17, c=11).
#pragma once
#include "common.cuh"

namespace sb {

__device__ __forceinline__ uint32_t enc_base(uint32_t b) {
    const uint32_t u = b & 0xDFu;
    uint32_t code = (b <= 3u) ? b : 0u;
    code = (u == 'C') ? 1u : code;
    code = (u == 'G') ? 2u : code;
    code = (u == 'T' || u == 'U') ? 3u : code;
    return code;
}

// a * b + c as IMAD: the sketch kernels are bound by the INT32 ALU pipe (LOP3 / SHF / IADD3), IMAD issues beside it on the
// FMA pipe at the same rate (tools/pipe_bench.cu), so shift-and-merge steps whose bit fields cannot overlap are written as
// multiply-adds
__device__ __forceinline__ uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// 64 bits of the packed sequence starting at base a (base a in the top bit pair)
__device__ __forceinline__ uint64_t window64(const uint32_t *packed, uint32_t a) {
    const uint32_t w = a >> 4, sh = (a & 15u) * 2u;
    const uint32_t w0 = packed[w], w1 = packed[w + 1], w2 = packed[w + 2];
    const uint32_t hi = __funnelshift_l(w1, w0, sh);
    const uint32_t lo = __funnelshift_l(w2, w1, sh);
    return ((uint64_t)hi << 32) | lo;
}
// len (1..16) bases starting at base a, right-aligned, first base in the highest used pair
__device__ __forceinline__ uint32_t extract32(const uint32_t *packed, uint32_t a, uint32_t len) {
    const uint32_t w = a >> 4, sh = (a & 15u) * 2u;
    return __funnelshift_l(packed[w + 1], packed[w], sh) >> (32u - 2u * len);
}
__device__ __forceinline__ uint64_t swap_pairs64(uint64_t r) {
    return ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
}
// reverse complement of the n-mer held in the TOP 2n bits of x; result right-aligned (mask = low 2n bits)
__device__ __forceinline__ uint64_t revcomp_top(uint64_t x, uint64_t mask) {
    const uint64_t z = ~x;
    const uint64_t r = ((uint64_t)__brev((uint32_t)z) << 32) | (uint64_t)__brev((uint32_t)(z >> 32));
    return swap_pairs64(r) & mask;
}
__device__ __forceinline__ uint32_t revcomp_smer(uint32_t f, uint32_t s) {
    const uint32_t z = ~(f << (32u - 2u * s));
    const uint32_t r = __brev(z);
    const uint32_t sw = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    return (s == 16u) ? sw : (sw & ((1u << (2u * s)) - 1u));
}

}  // namespace sb
