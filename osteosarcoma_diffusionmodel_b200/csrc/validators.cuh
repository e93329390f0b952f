// Kernels behind the validators (utils/validation.py): operand packing + row norms for the RBF Gram
// (the Gram itself is gemm_tc_kernel<EPI_RBF>) and the column-gathered moment reduction that
// replaces DataFrame.corr / Series.corr. See api_ops.inl.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"
#include "elem_kernels.cuh"

namespace osteo {

// One warp per row: x - center -> bf16 [hi | lo] (zero padded to kp columns) and the squared norm OF THE
// REPRESENTATION THE GEMM SEES (|hi|^2, or |hi|^2 + 2 hi.lo in split mode), so that ||a - a||^2 == 0 on the diagonal.
__global__ void pack_center_norm_kernel(const float* __restrict__ src, long long rows, int d, const float* __restrict__ center,
                                        __nv_bfloat16* __restrict__ dst, long long dst_rows, int kp, int lo_off, float* __restrict__ norms) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long r = warp; r < dst_rows; r += nwarps) {
        float acc = 0.0f;
        for (int c = lane * 2; c < kp; c += 64) {
            float v[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int cc = c + j;
                v[j] = (r < rows && cc < d) ? src[r * d + cc] - (center ? __ldg(center + cc) : 0.0f) : 0.0f;
            }
            const float h0 = bf16r(v[0]), h1 = bf16r(v[1]);
            *reinterpret_cast<uint32_t*>(dst + r * 2LL * kp + c) = pack2(v[0], v[1]);
            acc = fmaf(h0, h0, fmaf(h1, h1, acc));
            if (lo_off > 0) {
                const float l0 = bf16r(v[0] - h0), l1 = bf16r(v[1] - h1);
                *reinterpret_cast<uint32_t*>(dst + r * 2LL * kp + lo_off + c) = pack2(l0, l1);
                acc = fmaf(2.0f * h0, l0, fmaf(2.0f * h1, l1, acc));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) norms[r] = acc;
    }
}

// out[0] += rows, out[1 + j] += sum (x_j - s_j), out[1 + k + j*k + i] += sum (x_j - s_j)(x_i - s_i) over rows [rb, re)
// for the k <= 32 gathered columns. One warp per row: lane j holds column j of the row; fp64 accumulators.
__global__ void corr_moments_kernel(const float* __restrict__ data, int ld, const int* __restrict__ cols, int k, const float* __restrict__ shift,
                                    long long rb, long long re, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const int col = lane < k ? cols[lane] : 0;
    const float s = (lane < k && shift) ? shift[lane] : 0.0f;
    double s1 = 0.0;
    double s2[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) s2[i] = 0.0;
    long long count = 0;
    for (long long r = rb + warp; r < re; r += nwarps) {
        const float v = lane < k ? data[r * ld + col] - s : 0.0f;
        s1 += v;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (i < k) s2[i] += static_cast<double>(v) * static_cast<double>(__shfl_sync(0xffffffffu, v, i));
        }
        ++count;
    }
    // block-level reduction in shared memory (fp64), then ONE set of global atomics per block: the k + k*k addresses are
    // shared by every warp of the grid, so per-warp global atomics would serialise on them.
    __shared__ double red[1 + 32 + 32 * 32];
    for (int i = threadIdx.x; i < 1 + k + k * k; i += blockDim.x) red[i] = 0.0;
    __syncthreads();
    if (lane < k) {
        atomicAdd(&red[1 + lane], s1);
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < k) atomicAdd(&red[1 + k + lane * k + i], s2[i]);
    }
    if (lane == 0 && count) atomicAdd(&red[0], static_cast<double>(count));
    __syncthreads();
    for (int i = threadIdx.x; i < 1 + k + k * k; i += blockDim.x)
        if (red[i] != 0.0) atomicAdd(out + i, red[i]);
}

}  // namespace osteo
