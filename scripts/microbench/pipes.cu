// throughput microbenchmark of the instructions the LIC kernel is made of (per-SM rate, all 148 SMs busy)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITERS 4096
#define UNROLL 8
template <int OP>
__global__ void bench(float *out, const float *in)
{
    float a[UNROLL]; u64 p[UNROLL]; unsigned int u[UNROLL];
    for (int i = 0; i < UNROLL; ++i) { a[i] = in[threadIdx.x + i]; u[i] = __float_as_uint(a[i]) | 1u; p[i] = ((u64)u[i] << 32) | u[i]; }
    const float f = in[64];
    const u64 f2 = ((u64)__float_as_uint(f) << 32) | __float_as_uint(f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (OP == 0) a[i] = fmaf(a[i], f, 1.0f);
            if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(f2));
            if (OP == 2) { unsigned short h = (unsigned short)u[i]; float r; asm volatile("cvt.f32.f16 %0, %1;" : "=f"(r) : "h"(h)); u[i] = __float_as_uint(r) + 1; }
            if (OP == 3) { unsigned short h = (unsigned short)u[i]; asm volatile("add.f32.f16 %0, %1, %0;" : "+f"(a[i]) : "h"(h)); }
            if (OP == 4) u[i] = __byte_perm(u[i], 0x4B000000u, 0x7651u);
            if (OP == 5) u[i] = u[i] * 3u + 7u;
            if (OP == 6) a[i] = a[i] + f;
            if (OP == 7) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(f2));
            if (OP == 8) a[i] = fminf(a[i], f) + 0.0f * a[i];
            if (OP == 9) { int v = __float2int_rd(a[i]); a[i] = a[i] + (float)(v & 1); }
        }
    }
    float s = 0;
    for (int i = 0; i < UNROLL; ++i) s += a[i] + __uint_as_float(u[i]) + __uint_as_float((unsigned int)p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char *name, float *out, float *in, int extra)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<OP><<<148 * 4, 512>>>(out, in);
    cudaEventRecord(e0);
    bench<OP><<<148 * 4, 512>>>(out, in);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_instr = 148.0 * 4 * 16 * (double)ITERS * UNROLL;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-22s %8.3f ms  %6.3f warp-instr/clk/SM (x%d instr per op counted as 1)\n", name, ms, warp_instr / cyc / 148.0, extra);
}
int main()
{
    float *out, *in; cudaMalloc(&out, 148 * 4 * 512 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 0x3c, 4096);
    run<0>("FFMA", out, in, 1); run<1>("FFMA2", out, in, 1); run<6>("FADD", out, in, 1); run<7>("FADD2", out, in, 1);
    run<2>("cvt.f32.f16 (+IADD)", out, in, 2); run<3>("FHADD add.f32.f16", out, in, 1); run<4>("PRMT", out, in, 1);
    run<5>("IMAD", out, in, 1); run<8>("FMNMX+FFMA", out, in, 2); run<9>("F2I+LOP+I2F+FADD", out, in, 4);
    return 0;
}
