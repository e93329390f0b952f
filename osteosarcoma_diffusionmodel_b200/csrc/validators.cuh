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

// All pathways in ONE pass over the cohort (validate_pathway_coherence gathers ~15 genes for each of 10 pathways out of 371 columns:
// ten separate gathers re-read the same rows ten times, one 4-byte element per 32-byte sector). A block = P warps, warp p owns
// pathway p for the whole launch (its fp64 moment block stays in registers); the block streams chunks of whole rows through shared
// memory with coalesced loads and every warp gathers its columns from there.
//   cols [P][32] int32, -1 padded; shift [P][32]; out [P][CM_STRIDE] fp64 = {count, s1[32], s2[32][32]} (accumulated: zero it first)
constexpr int CM_STRIDE = 1 + 32 + 32 * 32;
__global__ void corr_moments_batched_kernel(const float* __restrict__ data, int ld, int ncols, const int* __restrict__ cols, int P, const float* __restrict__ shift,
                                            long long rb, long long re, int chunk_rows, double* __restrict__ out) {
    extern __shared__ float rows_sm[];                       // [chunk_rows][ncols]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = cols[warp * 32 + lane];
    const bool has = col >= 0;
    int k = __popc(__ballot_sync(0xffffffffu, has));          // columns are packed at the front
    const float s = has ? shift[warp * 32 + lane] : 0.0f;
    double s1 = 0.0;
    double s2[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) s2[i] = 0.0;
    long long count = 0;
    const long long nchunks = (re - rb + chunk_rows - 1) / chunk_rows;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const long long r0 = rb + c * chunk_rows;
        const int nr = static_cast<int>(re - r0 < chunk_rows ? re - r0 : chunk_rows);
        __syncthreads();                                      // the previous chunk has been consumed
        if (ld == ncols) {
            // rows are contiguous: one flat coalesced copy, 128-bit and unrolled (8 independent loads per thread in flight: a scalar
            // load -> store loop has one, and the staging of a chunk then costs 70 DRAM round trips)
            const float* src = data + r0 * ld;
            const int total = nr * ncols;
            if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
                const int n4 = total >> 2;
                const float4* s4 = reinterpret_cast<const float4*>(src);
                float4* d4 = reinterpret_cast<float4*>(rows_sm);
                int i = threadIdx.x;
                for (; i + 7 * static_cast<int>(blockDim.x) < n4; i += 8 * blockDim.x) {
                    float4 t[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) t[u] = __ldg(s4 + i + u * blockDim.x);
#pragma unroll
                    for (int u = 0; u < 8; ++u) d4[i + u * blockDim.x] = t[u];
                }
                for (; i < n4; i += blockDim.x) d4[i] = __ldg(s4 + i);
                for (int j = (n4 << 2) + threadIdx.x; j < total; j += blockDim.x) rows_sm[j] = src[j];
            } else {
                for (int i = threadIdx.x; i < total; i += blockDim.x) rows_sm[i] = src[i];
            }
        } else {
            for (int i = threadIdx.x; i < nr * ncols; i += blockDim.x) rows_sm[i] = data[(r0 + i / ncols) * ld + i % ncols];
        }
        __syncthreads();
        // products are accumulated in fp32 over sub-blocks of 16 rows (shifted, O(1) values) and flushed to the fp64 moment block:
        // per-product fp32 -> fp64 conversions run on the 16-lane XU pipe and made the loop 10x slower than its FMAs
        for (int rs = 0; rs < nr; rs += 16) {
            const int re16 = rs + 16 < nr ? rs + 16 : nr;
            float a1 = 0.0f, a2[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) a2[i] = 0.0f;
            if (k <= 16) {
                for (int r = rs; r < re16; ++r) {
                    const float v = has ? rows_sm[r * ncols + col] - s : 0.0f;
                    a1 += v;
#pragma unroll
                    for (int i = 0; i < 16; ++i) a2[i] = fmaf(v, __shfl_sync(0xffffffffu, v, i), a2[i]);
                }
            } else {
                for (int r = rs; r < re16; ++r) {
                    const float v = has ? rows_sm[r * ncols + col] - s : 0.0f;
                    a1 += v;
#pragma unroll
                    for (int i = 0; i < 32; ++i) a2[i] = fmaf(v, __shfl_sync(0xffffffffu, v, i), a2[i]);
                }
            }
            s1 += static_cast<double>(a1);
#pragma unroll
            for (int i = 0; i < 32; ++i) s2[i] += static_cast<double>(a2[i]);
        }
        count += nr;
    }
    double* o = out + static_cast<size_t>(warp) * CM_STRIDE;
    if (lane == 0 && count) atomicAdd(o, static_cast<double>(count));
    if (has && count) {
        atomicAdd(o + 1 + lane, s1);
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < k) atomicAdd(o + 33 + lane * 32 + i, s2[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Differentiable forms of the two correlation validators (the reference only has stubs, models/cvae.py:262-302;
// the forward values are tied to validate_pathway_coherence / validate_mutation_expression_correlation, utils/validation.py:125-223).
//   mode 0  (pathway coherence):   loss_s = 1 - mean_{i<j} R_ij                      over the set's k >= 2 columns
//   mode +1 / -1 (required sign):  loss_s = max(0, -mode * R_01)                     (k == 2: mutation column, pathway-score column)
// with R the Pearson correlation matrix over the rows. Written S_s = sum_{i<j} R_ij, the gradient is
//   dS/dx[r, i] = (1 / (n sd_i)) * (sum_j z_j - z_i - rho_i z_i),   z = (x - mean) / sd,   rho_i = sum_{j != i} R_ij
// so the backward pass needs only {mean_i, 1/sd_i, rho_i} per column and dloss/dS per set.
constexpr int CL_COEF = 4;      // per column: mean, 1/sd, (dloss/dS) / (n sd), rho

// One warp per column set, from the shifted fp64 moments of corr_moments_batched_kernel.
__global__ void corr_loss_finish_kernel(const double* __restrict__ mom, const int* __restrict__ cols, const float* __restrict__ shift,
                                        const int* __restrict__ modes, int n_sets, float* __restrict__ loss_out, float* __restrict__ coef) {
    const int lane = threadIdx.x & 31;
    const int set = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (set >= n_sets) return;
    const double* m = mom + static_cast<size_t>(set) * CM_STRIDE;
    const bool has = cols[set * 32 + lane] >= 0;
    const int k = __popc(__ballot_sync(0xffffffffu, has));
    const double n = m[0];
    const double mean = (has && n > 0) ? m[1 + lane] / n : 0.0;
    const double var = (has && n > 0) ? m[33 + lane * 32 + lane] / n - mean * mean : 0.0;
    const double isd = var > 0.0 ? rsqrt(var) : 0.0;      // a constant column has no correlation: it contributes R = 0
    double rho = 0.0;
    for (int j = 0; j < k; ++j) {
        const double mean_j = __shfl_sync(0xffffffffu, mean, j), isd_j = __shfl_sync(0xffffffffu, isd, j);
        // n == 0 (every row of the batch filtered out): no correlation is defined -- rho stays 0 so that the set's loss is a finite
        // constant with zero gradient instead of 0 / 0
        if (has && j != lane && n > 0) rho += (m[33 + lane * 32 + j] / n - mean * mean_j) * isd * isd_j;
    }
    double S = rho;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) S += __shfl_xor_sync(0xffffffffu, S, o);
    S *= 0.5;
    const int mode = modes[set];
    const double npairs = 0.5 * k * (k - 1);
    double loss, dS;
    if (!(n > 0)) {
        loss = 0.0;       // empty batch: the term vanishes (value and gradient)
        dS = 0.0;
    } else if (mode == 0) {
        loss = npairs > 0 ? 1.0 - S / npairs : 0.0;
        dS = npairs > 0 ? -1.0 / npairs : 0.0;
    } else {
        const double v = -static_cast<double>(mode) * S;
        loss = v > 0.0 ? v : 0.0;
        dS = v > 0.0 ? -static_cast<double>(mode) : 0.0;
    }
    if (lane == 0) loss_out[set] = static_cast<float>(loss);
    float* c = coef + (static_cast<size_t>(set) * 32 + lane) * CL_COEF;
    c[0] = static_cast<float>(mean + (has ? static_cast<double>(shift[set * 32 + lane]) : 0.0));
    c[1] = static_cast<float>(isd);
    c[2] = (has && n > 0) ? static_cast<float>(dS * isd / n) : 0.0f;
    c[3] = static_cast<float>(rho);
}

// grad[r, col] += upstream[set] * dloss_set/dx[r, col]. Block = n_sets warps (warp <-> set), grid-stride over rows; columns shared by
// several sets collide only within a row, hence the atomics.
__global__ void corr_loss_bwd_kernel(const float* __restrict__ data, long long n, int ld, const int* __restrict__ cols, int n_sets,
                                     const float* __restrict__ coef, const float* __restrict__ upstream, float* __restrict__ grad) {
    const int lane = threadIdx.x & 31, set = threadIdx.x >> 5;
    const int col = cols[set * 32 + lane];
    const bool has = col >= 0;
    const float* c = coef + (static_cast<size_t>(set) * 32 + lane) * CL_COEF;
    const float mean = c[0], isd = c[1], a = c[2] * upstream[set], rho = c[3];
    if (__all_sync(0xffffffffu, a == 0.0f)) return;       // inactive rule / zero upstream
    for (long long r = blockIdx.x; r < n; r += gridDim.x) {
        const float z = has ? (data[r * ld + col] - mean) * isd : 0.0f;
        float sz = z;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sz += __shfl_xor_sync(0xffffffffu, sz, o);
        if (has) atomicAdd(grad + r * ld + col, a * (sz - z - rho * z));
    }
}

}  // namespace osteo
