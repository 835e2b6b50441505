// vv_kernels.h -- host-callable launchers of the sm_100a kernels (vv_kernels.cu, vv_preprocess.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vvb200 {

struct DevParams;

size_t shared_table_bytes();
cudaError_t launch_lic_raycast(const DevParams &P, int layout, int illum, bool noise_gate, bool speed_of_flow, int grid, cudaStream_t st);
// sample-parallel pipeline (ray_setup -> [lic_sample -> composite] per depth window)
cudaError_t launch_ray_setup(const DevParams &P, int grid, cudaStream_t st);
// march checkpoints (every kCkStride-th position of every ray), once the set-up's allocation has sized DevParams::rayCk
cudaError_t launch_ray_checkpoints(const DevParams &P, int grid, cudaStream_t st);
// same view as the previous frame: put back what a frame consumes (ray state, tile accumulators, per-frame counters)
cudaError_t launch_ray_reset(const DevParams &P, unsigned int *counters, int grid, cudaStream_t st);
cudaError_t launch_lic_sample(const DevParams &P, int layout, int illum, bool noise_gate, bool speed_of_flow, int grid, cudaStream_t st);
cudaError_t launch_composite(const DevParams &P, int grid, cudaStream_t st);
cudaError_t launch_item_buckets(const DevParams &P, int nBuckets, int grid, cudaStream_t st);
// diagnostic (vv_debug_walk): one direction of the LIC walk from one position, 8 floats per step
cudaError_t launch_debug_walk(const DevParams &P, int layout, bool grad, int xf, const float pos[3], int dirSign, int nSteps, float *out, cudaStream_t st);
cudaError_t launch_volume_raycast(const DevParams &P, int layout, int grid, cudaStream_t st);
cudaError_t launch_lic_volume(const DevParams &P, int layout, bool grad, bool noise_gate, bool speed_of_flow, int grid, cudaStream_t st);
cudaError_t launch_unblock(const float4 *tiles, int world, int blocksPerRank, int nBlocksX, int nBlocksY, int skew, int unit, int width, int height,
                           float4 *frame, uchar4 *frame8, uchar4 *display8, cudaStream_t st);

// ---- peer-to-peer tile exchange (multi-GPU) ----
constexpr int kMaxPeers = 16;
struct P2PArgs {
    const float4 *tiles;               // this rank's tile buffer, n float4
    float4 *peerTiles[kMaxPeers];      // gather buffer (current parity) of every rank, own included: [world][n]
    unsigned int *peerFlags[kMaxPeers];// arrival counter (current parity) in every rank's buffer
    unsigned int *doneCounter;         // local scratch, zero between launches
    int rank, world;
    size_t n;                          // blocksPerRank * 256
};
cudaError_t launch_scatter_tiles(const P2PArgs &a, int grid, cudaStream_t st);
cudaError_t launch_wait_arrivals(unsigned int *flag, unsigned int target, unsigned int *err, cudaStream_t st);

// ---- pre-processing (K6) ----
// VectorDataSet::fillTexDataFloatInterp (VV/dataset.cpp:533-635): raw FLOAT3 / UCHAR3 time steps -> packed field.
// tmp: float4 [n] scratch; maxbits: 1 uint scratch.  Writes the pair-packed fp16 layout (padded [nz+2G][ny+2G][frow], cell
// (x,y,z) at ((z+G)(ny+2G) + (y+G)) frow + x + gx, edge replicated) and/or the float4 layout.
cudaError_t launch_pack_field(const void *v0, const void *v1, int is_u8, int nx, int ny, int nz, float interp_frac,
                              float4 *tmp, unsigned int *maxbits, uint4 *out_pair, int guard, int gx, int frow, int quad, float4 *out_f4, cudaStream_t st);
// u8 volume -> cell8 layout (wrap: 0 = CLAMP_TO_EDGE, 1 = REPEAT); src_stride = bytes per voxel, src_offset = channel;
// pad = 1 adds the wrapped cell -1 on every axis: out is [nz+1][ny+1][nx+1]
cudaError_t launch_build_cell8(const uint8_t *src, int src_stride, int src_offset, int nx, int ny, int nz, int repeat, int pad,
                               uint2 *out, cudaStream_t st);
// RGBA8 volume -> xy-quad layout with REPEAT
cudaError_t launch_build_quad(const uchar4 *src, int nx, int ny, int nz, uint4 *out, cudaStream_t st);
// RGBA8 volume -> x-pair layout with wrapped guard cells: out is [nz+2G][ny+2G][frow], cell x at column x + gx (default geometry
// G = 1, gx = 1, frow = nx + 1); bf16diff = 0: {half4 T[x], half4 T[x+1]}, 1: {bf16x4 T[x], bf16x4 (T[x+1] - T[x])}
cudaError_t launch_build_noise_pair(const uchar4 *src, int nx, int ny, int nz, int guard, int gx, int frow, uint4 *out, int bf16diff, cudaStream_t st);
// float scalar volume -> u8 LUMINANCE (GL float->UNORM8 conversion on upload)
cudaError_t launch_float_to_unorm8(const float *src, size_t n, uint8_t *out, cudaStream_t st);
// noise gradients, VV/gradient.cpp:190-532: Sobel/one-sided -> 5^3 smoothing (Q16) -> normalise + quantise, and
// NoiseDataSet::createTexture's RGBA8 packing (VV/dataset.cpp:1264-1282)
cudaError_t launch_noise_gradients(const uint8_t *noise, int nx, int ny, int nz, const float slice_dist[3],
                                   const float *filter125, float *grad_tmp, uchar4 *out_rgba, cudaStream_t st);

} // namespace vvb200
