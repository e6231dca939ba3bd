// pipes.cu -- micro-benchmark: can right/funnel shifts be moved off the ALU pipe (SHF) onto the FMA pipe
// (IMAD.HI / IMAD.WIDE with a multiplier ptxas cannot see through) while LOP3 keeps the ALU pipe busy?
// Prints warp-instructions per cycle per SM for a few instruction mixes.   nvcc -arch=sm_100a -O3 pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
template <int MODE>
__global__ void k(uint32_t* out, const uint32_t* cst, int n) {
    uint32_t c0 = cst[0], c1 = cst[1];          // opaque powers of two
    uint32_t a = threadIdx.x * 2654435761u, b = a ^ 0x9e3779b9u, c = b * 3, d = c + 7, e = a + 11, f = b + 13, g = c ^ 5, h = d ^ 9;
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0) {            // 8 LOP3
                a = (a ^ b) | c; b = (b ^ c) | d; c = (c ^ d) | e; d = (d ^ e) | f; e = (e ^ f) | g; f = (f ^ g) | h; g = (g ^ h) | a; h = (h ^ a) | b;
            } else if (MODE == 1) {     // 4 LOP3 + 4 SHF (funnel, constant)
                a = (a ^ b) | c; b = __funnelshift_r(b, c, 3); c = (c ^ d) | e; d = __funnelshift_r(d, e, 5);
                e = (e ^ f) | g; f = __funnelshift_r(f, g, 7); g = (g ^ h) | a; h = __funnelshift_r(h, a, 9);
            } else if (MODE == 2) {     // 4 LOP3 + 4 x (IMAD + IMAD.HI) with opaque multipliers = funnel shift on the FMA pipe
                uint32_t t;
                a = (a ^ b) | c; asm("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(c), "r"(c0)); asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(b) : "r"(b), "r"(c0), "r"(t));
                c = (c ^ d) | e; asm("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(e), "r"(c1)); asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(d), "r"(c1), "r"(t));
                e = (e ^ f) | g; asm("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(g), "r"(c0)); asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(f) : "r"(f), "r"(c0), "r"(t));
                g = (g ^ h) | a; asm("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(a), "r"(c1)); asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(h) : "r"(h), "r"(c1), "r"(t));
            } else if (MODE == 3) {     // 4 LOP3 + 4 IMAD.HI (plain right shift on the FMA pipe)
                a = (a ^ b) | c; asm("mul.hi.u32 %0, %1, %2;" : "=r"(b) : "r"(b), "r"(c0));
                c = (c ^ d) | e; asm("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(d), "r"(c1));
                e = (e ^ f) | g; asm("mul.hi.u32 %0, %1, %2;" : "=r"(f) : "r"(f), "r"(c0));
                g = (g ^ h) | a; asm("mul.hi.u32 %0, %1, %2;" : "=r"(h) : "r"(h), "r"(c1));
            } else if (MODE == 4) {     // 4 LOP3 + 4 IMAD.SHL (left shift; ptxas picks the pipe)
                a = (a ^ b) | c; b = (b << 3) + c; c = (c ^ d) | e; d = (d << 5) + e; e = (e ^ f) | g; f = (f << 7) + g; g = (g ^ h) | a; h = (h << 9) + a;
            } else if (MODE == 5) {     // 8 IMAD (multiply-add, opaque)
                a = a * c0 + b; b = b * c1 + c; c = c * c0 + d; d = d * c1 + e; e = e * c0 + f; f = f * c1 + g; g = g * c0 + h; h = h * c1 + a;
            } else if (MODE == 6) {     // 8 IMAD.HI
                asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(a), "r"(c0), "r"(b)); asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(b) : "r"(b), "r"(c1), "r"(c));
                asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(c) : "r"(c), "r"(c0), "r"(d)); asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(d), "r"(c1), "r"(e));
                asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(e) : "r"(e), "r"(c0), "r"(f)); asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(f) : "r"(f), "r"(c1), "r"(g));
                asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(g) : "r"(g), "r"(c0), "r"(h)); asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(h) : "r"(h), "r"(c1), "r"(a));
            } else if (MODE == 7) {     // 4 LOP3 + 4 PRMT
                a = (a ^ b) | c; b = __byte_perm(b, c, 0x4321); c = (c ^ d) | e; d = __byte_perm(d, e, 0x5432);
                e = (e ^ f) | g; f = __byte_perm(f, g, 0x6543); g = (g ^ h) | a; h = __byte_perm(h, a, 0x4321);
            } else if (MODE == 8) {     // 4 LOP3 + 4 POPC
                a = (a ^ b) | c; b = __popc(b) + c; c = (c ^ d) | e; d = __popc(d) + e; e = (e ^ f) | g; f = __popc(f) + g; g = (g ^ h) | a; h = __popc(h) + a;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

template <int MODE>
void run(const char* name, int per_iter, uint32_t* out, uint32_t* cst) {
    int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    int blocks = p.multiProcessorCount * 4, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, cst, 64);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, cst, ITERS);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    double cycles = ms * 1e-3 * clk * 1e3;
    double winst = (double)blocks * (threads / 32) * ITERS * 8.0 * per_iter / 8.0;   // per_iter = instructions per unrolled group of 8 statements
    printf("%-44s %8.3f ms  %6.2f warp-instr/cycle/SM (assuming %d instr per 8 statements)\n", name, ms, winst / cycles / p.multiProcessorCount, per_iter);
}

int main() {
    uint32_t *out, *cst; cudaMalloc(&out, 1 << 24); cudaMalloc(&cst, 64);
    uint32_t h[2] = {1u << 29, 1u << 27}; cudaMemcpy(cst, h, 8, cudaMemcpyHostToDevice);
    run<0>("8 LOP3", 8, out, cst);
    run<1>("4 LOP3 + 4 SHF", 8, out, cst);
    run<2>("4 LOP3 + 4 (IMAD + IMAD.HI) opaque", 12, out, cst);
    run<3>("4 LOP3 + 4 IMAD.HI opaque", 8, out, cst);
    run<4>("4 LOP3 + 4 (x<<s)+y", 8, out, cst);
    run<5>("8 IMAD opaque", 8, out, cst);
    run<6>("8 IMAD.HI opaque", 8, out, cst);
    run<7>("4 LOP3 + 4 PRMT", 8, out, cst);
    run<8>("4 LOP3 + 4 (POPC + IADD)", 12, out, cst);
    return 0;
}
