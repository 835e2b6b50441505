// Gather-bandwidth microbenchmark ("L2 / L1 gather peaks ... the builder measures them once with a float4
// random-gather microbenchmark"): every lane issues independent 128-bit loads (LDG.E.128.CONSTANT, what the LIC samplers
// issue) at pseudo-random 16-byte slots of a working set, for three working-set sizes:
//   L1   : a private 32 KB window per CTA (4 CTAs x 256 threads per SM, as lic_sample_kernel runs), hit in L1
//   L2   : 64 MB shared by all CTAs (fits the 126 MB L2, misses L1)
//   HBM  : 4 GB
// and two access patterns: "random" (each lane its own slot) and "trilinear" (the 32 lanes of a warp read 32 neighbouring
// slots of one row, 4 rows per fetch -- the footprint of one software trilinear fetch of an 8x4 ray tile).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather gather.cu && ./gather
// Prints requested GB/s (bytes asked for by the lanes, the numerator of roofline.achieved) per pattern and working set.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;

template <int PATTERN>   // 0 random, 1 trilinear-like
__global__ void __launch_bounds__(256, 4) gather(const uint4 *__restrict__ base, u64 slots_per_cta, u64 cta_stride_slots, int iters, unsigned int *sink)
{
    const uint4 *p = base + (u64)blockIdx.x * cta_stride_slots;
    unsigned int s = (blockIdx.x * 256u + threadIdx.x) * 2654435761u + 12345u;
    const unsigned int lane = threadIdx.x & 31u;
    unsigned int acc = 0;
    const u64 mask = slots_per_cta - 1;                       // power of two
    for (int it = 0; it < iters; ++it) {
        s = s * 1664525u + 1013904223u;
        u64 i0;
        if (PATTERN == 0) i0 = (u64)(s >> 4);
        else i0 = (u64)(__shfl_sync(0xffffffffu, s, 0) >> 4) + lane;          // one random row start per warp, lanes adjacent
        // four independent 128-bit loads per iteration (corner rows of one trilinear fetch are 1 row / 1 plane apart)
        const uint4 a = __ldg(p + ((i0) & mask));
        const uint4 b = __ldg(p + ((i0 + 257) & mask));
        const uint4 c = __ldg(p + ((i0 + 66049) & mask));
        const uint4 d = __ldg(p + ((i0 + 66306) & mask));
        acc += (a.x ^ a.y ^ a.z ^ a.w) + (b.x ^ b.y ^ b.z ^ b.w) + (c.x ^ c.y ^ c.z ^ c.w) + (d.x ^ d.y ^ d.z ^ d.w);   // all 128 bits are used
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int PATTERN>
static void run(const char *what, const uint4 *buf, u64 slots_per_cta, u64 cta_stride, int grid, int iters, unsigned int *sink)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<PATTERN><<<grid, 256>>>(buf, slots_per_cta, cta_stride, iters / 8 + 1, sink);      // warm-up
    cudaEventRecord(e0);
    gather<PATTERN><<<grid, 256>>>(buf, slots_per_cta, cta_stride, iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)grid * 256 * (double)iters * 4 * 16;
    printf("%-34s %9.3f ms  %10.1f GB/s requested  (%s)\n", what, ms, bytes / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 4;
    const u64 hbm_slots = (u64)1 << 28;                      // 4 GB of 16-byte slots
    uint4 *buf;
    unsigned int *sink;
    if (cudaMalloc(&buf, hbm_slots * 16) != cudaSuccess) { printf("cudaMalloc failed\n"); return 1; }
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 0x11, hbm_slots * 16);
    cudaFuncSetAttribute(gather<0>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    cudaFuncSetAttribute(gather<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    // L1: 32 KB per CTA = 2048 slots, CTAs 1 MB apart
    run<0>("L1   random    (32 KB / CTA)", buf, 2048, 65536, grid, 1 << 14, sink);
    run<1>("L1   trilinear (32 KB / CTA)", buf, 2048, 65536, grid, 1 << 14, sink);
    // L2: all CTAs share 64 MB = 2^22 slots
    run<0>("L2   random    (64 MB)", buf, (u64)1 << 22, 0, grid, 1 << 12, sink);
    run<1>("L2   trilinear (64 MB)", buf, (u64)1 << 22, 0, grid, 1 << 12, sink);
    // HBM: 4 GB
    run<0>("HBM  random    (4 GB)", buf, hbm_slots, 0, grid, 1 << 10, sink);
    run<1>("HBM  trilinear (4 GB)", buf, hbm_slots, 0, grid, 1 << 10, sink);
    cudaFree(buf); cudaFree(sink);
    return 0;
}
