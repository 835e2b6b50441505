// vv_preprocess.cu -- bandwidth-bound pre-processing kernels: the GPU versions of the host loops the
// reference runs before it uploads its textures.  All arithmetic uses explicit round-to-nearest intrinsics in the
// reference's evaluation order, so the results are bit-identical to the CPU loops they replace (one stated exception: UCHAR3
// vector components are the signed (float)u - 128 of fillTexDataFloat, VV/dataset.cpp:474-490, not the in-place unsigned
// wrap-around of the interpolating path, VV/dataset.cpp:563-571).
//
//   pack_field_*        VectorDataSet::fillTexDataFloatInterp      VV/dataset.cpp:533-635
//   build_cell8/quad    sampler state baked into the layout         VV/dataset.cpp:1033-1038,1328-1335
//   sobel/filter/quant  computeGradients/filterGradients/quantize8  VV/gradient.cpp:190-532
#include "vv_device.cuh"
#include "vv_kernels.h"

namespace vvb200 {

// pass 1: interpolate between the two time steps, normalise the direction, keep |v| and the global max
template <bool U8>
__global__ void pack_field_pass1(const void *__restrict__ v0, const void *__restrict__ v1, size_t n, float frac,
                                 float4 *__restrict__ tmp, unsigned int *__restrict__ maxbits)
{
    float lmax = 0.0f;
    for (size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x; a < n; a += (size_t)gridDim.x * blockDim.x) {
        float t[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float a0, a1;
            if (U8) {
                a0 = (float)((const uint8_t *)v0)[3 * a + k] - 128.0f;
                a1 = a0;
            } else {
                a0 = ((const float *)v0)[3 * a + k];
                a1 = v1 ? ((const float *)v1)[3 * a + k] : a0;
            }
            t[k] = U8 ? a0 : __fadd_rn(a0, __fmul_rn(frac, __fsub_rn(a1, a0)));          // :590
        }
        float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(t[0], t[0]), __fmul_rn(t[1], t[1])), __fmul_rn(t[2], t[2])));   // :592
        float4 o;
        if (len < 1e-5f) {                                                                // EPS, VV/mmath.h:43
            len = 0.0f;
            o.x = o.y = o.z = 0.5f;
        } else {
            o.x = __fadd_rn(__fdiv_rn(__fmul_rn(0.5f, t[0]), len), 0.5f);                 // :605-607
            o.y = __fadd_rn(__fdiv_rn(__fmul_rn(0.5f, t[1]), len), 0.5f);
            o.z = __fadd_rn(__fdiv_rn(__fmul_rn(0.5f, t[2]), len), 0.5f);
        }
        o.w = len;
        tmp[a] = o;
        lmax = fmaxf(lmax, len);
    }
    // |v| >= 0: float order == unsigned order of the bit patterns
    lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, 16));
    lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, 8));
    lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, 4));
    lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, 2));
    lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, 1));
    if ((threadIdx.x & 31) == 0) atomicMax(maxbits, __float_as_uint(lmax));
}

__device__ __forceinline__ uint2 to_half4(float4 v, float maxLen)
{
    float a = __fdiv_rn(v.w, maxLen);                                                     // :629
    a = (a > 1.0f) ? 1.0f : ((a < 0.0f) ? 0.0f : a);
    __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, a);             // GL_RGBA16F_ARB upload, :329-347
    uint2 r;
    r.x = *reinterpret_cast<unsigned int *>(&lo);
    r.y = *reinterpret_cast<unsigned int *>(&hi);
    return r;
}

// pass 2: scale |v| by 1/max, round to fp16, write the x-pair layout and/or float4
// The x-pair layout carries CLAMP_TO_EDGE as replicated guard cells: the array is [nz+2G][ny+2G][fRow] with cell (x,y,z) at
// ((z+G) (ny+2G) + (y+G)) fRow + (x+gx); every cell holds texels (clamp(x), clamp(x+1)) of row clamp(y), plane clamp(z).
// G = 1 / gx = 0 is the minimum (the sampler's y / z neighbours are always one row / plane further, x is baked into the pair);
// larger guards let the walk drop the coordinate clamp altogether (XF_GUARD).
// quad != 0: every cell holds the pairs of rows y and y + 1 (2 x uint4, LAYOUT_QUAD)
__global__ void pack_field_pass2(const float4 *__restrict__ tmp, const unsigned int *__restrict__ maxbits, int nx, int ny, int nz,
                                 int guard, int gx, int frow, int quad, uint4 *__restrict__ out_pair, float4 *__restrict__ out_f4)
{
    const float maxLen = __uint_as_float(*maxbits);
    const size_t n = (size_t)nx * ny * nz;
    if (out_f4) {
        for (size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x; a < n; a += (size_t)gridDim.x * blockDim.x) {
            uint2 t0 = to_half4(tmp[a], maxLen);
            float2 rg = h2f(t0.x), ba = h2f(t0.y);
            out_f4[a] = make_float4(rg.x, rg.y, ba.x, ba.y);
        }
    }
    if (out_pair) {
        const int py = ny + 2 * guard, pz = nz + 2 * guard;
        const size_t np = (size_t)frow * py * pz;
        for (size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x; a < np; a += (size_t)gridDim.x * blockDim.x) {
            const int x = (int)(a % frow) - gx, y = (int)((a / frow) % py) - guard, z = (int)(a / ((size_t)frow * py)) - guard;
            const int x0 = min(max(x, 0), nx - 1), x1 = min(max(x + 1, 0), nx - 1);
            const size_t s = ((size_t)min(max(z, 0), nz - 1) * ny + min(max(y, 0), ny - 1)) * nx;
            uint2 t0 = to_half4(tmp[s + x0], maxLen);
            uint2 t1 = to_half4(tmp[s + x1], maxLen);
            if (quad) {
                const size_t s1 = ((size_t)min(max(z, 0), nz - 1) * ny + min(max(y + 1, 0), ny - 1)) * nx;
                uint2 u0 = to_half4(tmp[s1 + x0], maxLen);
                uint2 u1 = to_half4(tmp[s1 + x1], maxLen);
                out_pair[2 * a] = make_uint4(t0.x, t0.y, t1.x, t1.y);
                out_pair[2 * a + 1] = make_uint4(u0.x, u0.y, u1.x, u1.y);
                continue;
            }
            out_pair[a] = make_uint4(t0.x, t0.y, t1.x, t1.y);
        }
    }
}

cudaError_t launch_pack_field(const void *v0, const void *v1, int is_u8, int nx, int ny, int nz, float interp_frac,
                              float4 *tmp, unsigned int *maxbits, uint4 *out_pair, int guard, int gx, int frow, int quad, float4 *out_f4, cudaStream_t st)
{
    const size_t n = (size_t)nx * ny * nz;
    cudaError_t e = cudaMemsetAsync(maxbits, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    const int grid = 148 * 8;
    if (is_u8) pack_field_pass1<true><<<grid, 256, 0, st>>>(v0, v1, n, interp_frac, tmp, maxbits);
    else pack_field_pass1<false><<<grid, 256, 0, st>>>(v0, v1, n, interp_frac, tmp, maxbits);
    pack_field_pass2<<<grid, 256, 0, st>>>(tmp, maxbits, nx, ny, nz, guard, gx, frow, quad, out_pair, out_f4);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// pad = 1 (REPEAT volumes): the output is [nz+1][ny+1][nx+1] and entry (jx,jy,jz) is the cell whose low corner is
// texel (jx-1, jy-1, jz-1) mod n, so the sampler's cell index floor(u) in [-1, n-1] needs no wrap.
__global__ void build_cell8_kernel(const uint8_t *__restrict__ src, int stride, int off, int nx, int ny, int nz, int repeat, int pad,
                                   uint2 *__restrict__ out)
{
    const int px = nx + pad, py = ny + pad, pz = nz + pad;
    const size_t n = (size_t)px * py * pz;
    for (size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x; a < n; a += (size_t)gridDim.x * blockDim.x) {
        int x = (int)(a % px) - pad, y = (int)((a / px) % py) - pad, z = (int)(a / ((size_t)px * py)) - pad;
        if (x < 0) x += nx;
        if (y < 0) y += ny;
        if (z < 0) z += nz;
        const int x1 = (x + 1 < nx) ? x + 1 : (repeat ? 0 : x);
        const int y1 = (y + 1 < ny) ? y + 1 : (repeat ? 0 : y);
        const int z1 = (z + 1 < nz) ? z + 1 : (repeat ? 0 : z);
        auto at = [&](int xx, int yy, int zz) -> unsigned int {
            return src[(((size_t)zz * ny + yy) * nx + xx) * stride + off];
        };
        uint2 c;
        c.x = at(x, y, z) | (at(x1, y, z) << 8) | (at(x, y1, z) << 16) | (at(x1, y1, z) << 24);
        c.y = at(x, y, z1) | (at(x1, y, z1) << 8) | (at(x, y1, z1) << 16) | (at(x1, y1, z1) << 24);
        out[a] = c;
    }
}

cudaError_t launch_build_cell8(const uint8_t *src, int src_stride, int src_offset, int nx, int ny, int nz, int repeat, int pad,
                               uint2 *out, cudaStream_t st)
{
    build_cell8_kernel<<<148 * 8, 256, 0, st>>>(src, src_stride, src_offset, nx, ny, nz, repeat, pad, out);
    return cudaGetLastError();
}

__global__ void build_quad_kernel(const uchar4 *__restrict__ src, int nx, int ny, int nz, uint4 *__restrict__ out)
{
    const size_t n = (size_t)nx * ny * nz;
    const unsigned int *s = reinterpret_cast<const unsigned int *>(src);
    for (size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x; a < n; a += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(a % nx), y = (int)((a / nx) % ny), z = (int)(a / ((size_t)nx * ny));
        const int x1 = (x + 1 < nx) ? x + 1 : 0, y1 = (y + 1 < ny) ? y + 1 : 0;
        const size_t r0 = ((size_t)z * ny + y) * nx, r1 = ((size_t)z * ny + y1) * nx;
        out[a] = make_uint4(s[r0 + x], s[r0 + x1], s[r1 + x], s[r1 + x1]);
    }
}

cudaError_t launch_build_quad(const uchar4 *src, int nx, int ny, int nz, uint4 *out, cudaStream_t st)
{
    build_quad_kernel<<<148 * 8, 256, 0, st>>>(src, nx, ny, nz, out);
    return cudaGetLastError();
}

// RGBA8 volume -> x-pair layout {T[x mod nx], T[(x+1) mod nx]} with REPEAT baked in as wrapped guard cells:
// the array is [nz+2G][ny+2G][frow], cell (x,y,z) at ((z+G)(ny+2G) + (y+G)) frow + x + gx  (default G = 1, gx = 1, frow = nx+1;
// or the vector field's geometry, so that both arrays share one cell index, XF_NSHARE).
// bf16diff = 0: fp16 {half4 T0, half4 T1} (byte values are exact in fp16);  1: {bf16x4 T0, bf16x4 (T1 - T0)} (integers up to 255
// and their differences are exact in bf16)
__global__ void build_noise_pair_kernel(const uchar4 *__restrict__ src, int nx, int ny, int nz, int guard, int gx, int frow,
                                        uint4 *__restrict__ out, int bf16diff)
{
    const int py = ny + 2 * guard;
    const size_t n = (size_t)frow * py * (nz + 2 * guard);
    auto wrap = [](int v, int m) { v %= m; return v < 0 ? v + m : v; };
    for (size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x; a < n; a += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(a % frow) - gx, y = (int)((a / frow) % py) - guard, z = (int)(a / ((size_t)frow * py)) - guard;
        const int x0 = wrap(x, nx), x1 = wrap(x + 1, nx);
        const size_t r = ((size_t)wrap(z, nz) * ny + wrap(y, ny)) * nx;
        const uchar4 t0 = src[r + x0], t1 = src[r + x1];
        if (bf16diff) {
            auto bf = [](int lo, int hi) { return (__float_as_uint((float)lo) >> 16) | (__float_as_uint((float)hi) & 0xffff0000u); };
            out[a] = make_uint4(bf(t0.x, t0.y), bf(t0.z, t0.w), bf((int)t1.x - t0.x, (int)t1.y - t0.y), bf((int)t1.z - t0.z, (int)t1.w - t0.w));
            continue;
        }
        __half2 p0 = __floats2half2_rn((float)t0.x, (float)t0.y), p1 = __floats2half2_rn((float)t0.z, (float)t0.w);
        __half2 q0 = __floats2half2_rn((float)t1.x, (float)t1.y), q1 = __floats2half2_rn((float)t1.z, (float)t1.w);
        out[a] = make_uint4(*reinterpret_cast<unsigned int *>(&p0), *reinterpret_cast<unsigned int *>(&p1),
                            *reinterpret_cast<unsigned int *>(&q0), *reinterpret_cast<unsigned int *>(&q1));
    }
}

cudaError_t launch_build_noise_pair(const uchar4 *src, int nx, int ny, int nz, int guard, int gx, int frow, uint4 *out, int bf16diff, cudaStream_t st)
{
    build_noise_pair_kernel<<<148 * 8, 256, 0, st>>>(src, nx, ny, nz, guard, gx, frow, out, bf16diff);
    return cudaGetLastError();
}

__global__ void float_to_unorm8_kernel(const float *__restrict__ src, size_t n, uint8_t *__restrict__ out)
{
    for (size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x; a < n; a += (size_t)gridDim.x * blockDim.x)
        out[a] = (uint8_t)floorf(fminf(fmaxf(src[a], 0.0f), 1.0f) * 255.0f + 0.5f);
}

cudaError_t launch_float_to_unorm8(const float *src, size_t n, uint8_t *out, cudaStream_t st)
{
    float_to_unorm8_kernel<<<148 * 8, 256, 0, st>>>(src, n, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// computeGradients, VV/gradient.cpp:190-374 (SOBEL == 1)
__global__ void sobel_kernel(const uint8_t *__restrict__ vol, int nx, int ny, int nz, float sdx, float sdy, float sdz,
                             float *__restrict__ g)
{
    const size_t n = (size_t)nx * ny * nz;
    for (size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x; a < n; a += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(a % nx), y = (int)((a / nx) % ny), z = (int)(a / ((size_t)nx * ny));
        auto vox = [&](int xx, int yy, int zz) -> float { return (float)vol[((size_t)zz * ny + yy) * nx + xx]; };
        float gx, gy, gz;
        if (x > 0 && x < nx - 1 && y > 0 && y < ny - 1 && z > 0 && z < nz - 1) {
            // separable form of the 3x3x3 weights (1,3,1)x(1,3,1) smoothing, -1/0/+1 derivative; integer-valued
            // partial sums below 2^24, so any summation order gives the reference's float result exactly
            int sx = 0, sy = 0, sz = 0;
#pragma unroll
            for (int k = -1; k <= 1; ++k)
#pragma unroll
                for (int j = -1; j <= 1; ++j)
#pragma unroll
                    for (int i = -1; i <= 1; ++i) {
                        const int v = vol[((size_t)(z + k) * ny + (y + j)) * nx + (x + i)];
                        const int wi = (i == 0) ? 3 : 1, wj = (j == 0) ? 3 : 1, wk = (k == 0) ? 3 : 1;
                        // centre weight is 6 (not 9) in the reference table: 3*3 -> 6 when both orthogonal offsets are 0
                        const int wjk = (j == 0 && k == 0) ? 6 : wj * wk;
                        const int wik = (i == 0 && k == 0) ? 6 : wi * wk;
                        const int wij = (i == 0 && j == 0) ? 6 : wi * wj;
                        sx += i * wjk * v;
                        sy += j * wik * v;
                        sz += k * wij * v;
                    }
            gx = __fdiv_rn((float)sx, __fmul_rn(2.0f, sdx));
            gy = __fdiv_rn((float)sy, __fmul_rn(2.0f, sdy));
            gz = __fdiv_rn((float)sz, __fmul_rn(2.0f, sdz));
        } else {
            gx = (x < 1) ? __fdiv_rn(vox(x + 1, y, z) - vox(x, y, z), sdx) : __fdiv_rn(vox(x, y, z) - vox(x - 1, y, z), sdx);
            gy = (y < 1) ? __fdiv_rn(vox(x, y + 1, z) - vox(x, y, z), sdy) : __fdiv_rn(vox(x, y, z) - vox(x, y - 1, z), sdy);
            gz = (z < 1) ? __fdiv_rn(vox(x, y, z + 1) - vox(x, y, z), sdz) : __fdiv_rn(vox(x, y, z) - vox(x, y, z - 1), sdz);
        }
        g[3 * a] = gx; g[3 * a + 1] = gy; g[3 * a + 2] = gz;
    }
}

// filterGradients + quantize8 + RGBA packing, VV/gradient.cpp:377-478, VV/dataset.cpp:1264-1282.
// Q16: taps k,j,i = -fw .. fw-2 with the 5^3 table built for fw = 2 indexed by the border-shrunken fw.
__global__ void filter_quantize_kernel(const float *__restrict__ g, const uint8_t *__restrict__ noise, int nx, int ny, int nz,
                                       const float *__restrict__ filter125, uchar4 *__restrict__ out)
{
    __shared__ float s_f[125];
    for (int i = threadIdx.x; i < 125; i += blockDim.x) s_f[i] = filter125[i];
    __syncthreads();
    const size_t n = (size_t)nx * ny * nz;
    for (size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x; a < n; a += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(a % nx), y = (int)((a / nx) % ny), z = (int)(a / ((size_t)nx * ny));
        const int bx = min(x, nx - x - 1), by = min(y, ny - y - 1), bz = min(z, nz - z - 1);
        const int fw = min(2, min(min(bx, by), bz));
        float acc[3] = {0.0f, 0.0f, 0.0f};
        for (int k = -fw; k < fw - 1; ++k)
            for (int j = -fw; j < fw - 1; ++j)
                for (int i = -fw; i < fw - 1; ++i) {
                    const size_t o = 3 * (((size_t)(z + k) * ny + (y + j)) * nx + (x + i));
                    const float w = s_f[((fw + k) * 5 + fw + j) * 5 + fw + i];
                    acc[0] = __fadd_rn(acc[0], __fmul_rn(w, g[o]));
                    acc[1] = __fadd_rn(acc[1], __fmul_rn(w, g[o + 1]));
                    acc[2] = __fadd_rn(acc[2], __fmul_rn(w, g[o + 2]));
                }
        float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(acc[0], acc[0]), __fmul_rn(acc[1], acc[1])), __fmul_rn(acc[2], acc[2])));
        if (len < 1e-5f) {
            acc[0] = acc[1] = acc[2] = 0.0f;
        } else {
            acc[0] = __fdiv_rn(acc[0], len); acc[1] = __fdiv_rn(acc[1], len); acc[2] = __fdiv_rn(acc[2], len);
        }
        unsigned char q[3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
            q[c] = (unsigned char)__dmul_rn(__ddiv_rn(__dadd_rn((double)acc[c], 1.0), 2.0), 255.0);   // quantize8 :477
        out[a] = make_uchar4(q[0], q[1], q[2], noise[a]);
    }
}

cudaError_t launch_noise_gradients(const uint8_t *noise, int nx, int ny, int nz, const float sd[3], const float *filter125,
                                   float *grad_tmp, uchar4 *out_rgba, cudaStream_t st)
{
    sobel_kernel<<<148 * 8, 256, 0, st>>>(noise, nx, ny, nz, sd[0], sd[1], sd[2], grad_tmp);
    filter_quantize_kernel<<<148 * 8, 256, 0, st>>>(grad_tmp, noise, nx, ny, nz, filter125, out_rgba);
    return cudaGetLastError();
}

} // namespace vvb200
