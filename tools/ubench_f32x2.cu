// Microbenchmark: issue rate of FADD/FFMA vs FADD2/FFMA2, and mixing with ALU-pipe ops, on one SM-full grid.
#include <cuda_runtime.h>
#include <cstdio>
#define ITERS 4096
template <int MODE>
__global__ void k(float* out, float s) {
    float a[8]; unsigned long long p[8]; unsigned int u[8];
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x + i; float2 f = make_float2(a[i], a[i] + 1); p[i] = *reinterpret_cast<unsigned long long*>(&f); u[i] = threadIdx.x * 7 + i; }
    float2 sf = make_float2(s, s); unsigned long long sp = *reinterpret_cast<unsigned long long*>(&sf);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) a[i] = a[i] + s;                                       // FADD
            if (MODE == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(sp));   // FADD2
            if (MODE == 2) a[i] = fmaf(a[i], s, s);                                // FFMA
            if (MODE == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(sp));  // FFMA2
            if (MODE == 4) { a[i] = a[i] + s; u[i] = (u[i] >> 3) ^ u[i]; }          // FADD + ALU
            if (MODE == 5) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(sp)); u[i] = (u[i] >> 3) ^ u[i]; u[(i+1)&7] += u[i]; }  // FADD2 + 2 ALU
            if (MODE == 6) { u[i] = (u[i] >> 3) ^ u[i]; }                           // ALU only (SHF+LOP3 or LOP3 w/ shift)
            if (MODE == 7) { a[i] = a[i] + s; a[i] = a[i] * s; u[i] = (u[i] >> 3) ^ u[i]; u[(i+1)&7] += u[i];}  // 2 FP + 2 ALU
        }
    }
    float r = 0; for (int i = 0; i < 8; i++) { float2 f = *reinterpret_cast<float2*>(&p[i]); r += a[i] + f.x + f.y + (float)u[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, float flops_per_iter) {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 256>>>(d, 1.0001f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 256>>>(d, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warps = 148.0 * 4 * 8;  // warps total; per SMSP: 8 warps
    double loopbody = (double)ITERS * 8;
    // cycles per loop-body-op per SMSP assuming 1.9 GHz
    double cyc = ms * 1e-3 * 1.9e9 / (loopbody * 8 /*warps per SMSP*/);
    printf("%-28s %.3f ms  ~%.2f cycles per unrolled slot per warp (at 1.9GHz)\n", name, ms, cyc);
    cudaFree(d);
}
int main() {
    run<0>("FADD", 1); run<1>("FADD2", 2); run<2>("FFMA", 2); run<3>("FFMA2", 4);
    run<4>("FADD+ALU", 1); run<5>("FADD2+2ALU", 1); run<6>("ALU", 1); run<7>("2FP+2ALU", 1);
    return 0;
}
