// Issue cost of the lerp instruction patterns of lic_sample_kernel with REALISTIC operands (distinct source registers), to
// tell the FMA pipe's width from the register-file read bandwidth.  pipes.cu measures a = fma(a, f, imm) -- one register
// operand plus a reused one -- which is the best case.  Here every instruction reads distinct registers, as the lerps do:
//   x-lerp      t = fma2(fx, d[i], t0[i])          FFMA2 R, R.F32 (broadcast), Rpair, Rpair
//   y/z-lerp    t = fma2(fy, b - a, a)             FADD2 + FFMA2
//   widen       t0 = h + 0, d = h1 - t0            FHADD x 2
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes2 pipes2.cu && ./pipes2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITERS 2048
#define N 8
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fh_add(unsigned short h, float c) { float d; asm volatile("add.rn.f32.f16 %0, %1, %2;" : "=f"(d) : "h"(h), "f"(c)); return d; }

template <int OP>
__global__ void __launch_bounds__(256, 4) bench(float *out, const float *in)
{
    float a[N], b[N], c[N]; u64 p[N], q[N], r[N]; unsigned short h[N];
    for (int i = 0; i < N; ++i) {
        a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 32 + i]; c[i] = in[threadIdx.x + 64 + i];
        p[i] = pk2(a[i], b[i]); q[i] = pk2(b[i], c[i]); r[i] = pk2(c[i], a[i]); h[i] = (unsigned short)(__float_as_uint(a[i]) >> 16);
    }
    const float f = in[200 + (threadIdx.x & 1)], g = in[300 + (threadIdx.x & 1)];
    const u64 f2 = pk2(f, f), g2 = pk2(g, f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (OP == 0) a[i] = fmaf(f, b[i], a[i]);                       // FFMA, 3 registers, f reusable
            if (OP == 1) a[i] = fmaf(c[i], b[i], a[i]);                    // FFMA, 3 distinct registers
            if (OP == 2) p[i] = fma2(f2, q[i], p[i]);                      // FFMA2, broadcast scalar x pair + pair
            if (OP == 3) p[i] = fma2(g2, q[i], p[i]);                      // FFMA2, reusable pair x pair + pair
            if (OP == 4) p[i] = fma2(r[i], q[i], p[i]);                    // FFMA2, 3 distinct pairs
            if (OP == 5) p[i] = sub2(q[i], p[i]);                          // FADD2, 2 distinct pairs
            if (OP == 6) a[i] = fh_add(h[i], a[i]);                        // FHADD
            if (OP == 7) { p[i] = fma2(f2, sub2(q[i], p[i]), p[i]); }      // lerp2: FADD2 + FFMA2 (counted as 2)
            if (OP == 8) { const float t0 = fh_add(h[i], 0.0f); a[i] = fmaf(f, fh_add(h[(i + 1) % N], -t0), a[i] + t0); }   // 2 FHADD + FADD + FFMA
            if (OP == 9) { const float t0 = fh_add(h[i], a[i]), t1 = fh_add(h[(i + 1) % N], b[i]);                          // 4 FHADD + FFMA2 (the x-lerp of two channels)
                           const float d0 = fh_add(h[(i + 2) % N], -t0), d1 = fh_add(h[(i + 3) % N], -t1);
                           p[i] = fma2(f2, pk2(d0, d1), pk2(t0, t1)); a[i] = __uint_as_float((unsigned int)p[i]); b[i] = __uint_as_float((unsigned int)(p[i] >> 32)); }
        }
    }
    float s = 0;
    for (int i = 0; i < N; ++i) s += a[i] + b[i] + __uint_as_float((unsigned int)p[i]) + __uint_as_float((unsigned int)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char *name, float *out, float *in, int per)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<OP><<<148 * 4, 256>>>(out, in);
    cudaEventRecord(e0);
    bench<OP><<<148 * 4, 256>>>(out, in);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops_per_smsp = 8.0 * ITERS * N;               // 4 CTAs x 8 warps per SM = 8 warps per sub-partition
    const double cyc = ms * 1e-3 * 1.965e9;
    printf("%-46s %8.3f ms  %6.3f SMSP-cycles per op = %d instr -> %.3f cycles per instruction\n", name, ms, cyc / ops_per_smsp, per,
           cyc / ops_per_smsp / per);
}
int main()
{
    float *out, *in; cudaMalloc(&out, 148 * 4 * 256 * 4); cudaMalloc(&in, 8192); cudaMemset(in, 0x3c, 8192);
    run<0>("FFMA  f(reused) * b[i] + a[i]", out, in, 1);
    run<1>("FFMA  c[i] * b[i] + a[i]", out, in, 1);
    run<2>("FFMA2 f.F32 * q[i] + p[i]", out, in, 1);
    run<3>("FFMA2 g2(reused pair) * q[i] + p[i]", out, in, 1);
    run<4>("FFMA2 r[i] * q[i] + p[i]", out, in, 1);
    run<5>("FADD2 q[i] - p[i]", out, in, 1);
    run<6>("FHADD h[i] + a[i]", out, in, 1);
    run<7>("lerp2 = FADD2 + FFMA2", out, in, 2);
    run<8>("scalar x-lerp = 2 FHADD + FADD + FFMA", out, in, 4);
    run<9>("packed x-lerp of 2 channels = 4 FHADD + FFMA2", out, in, 5);
    return 0;
}
