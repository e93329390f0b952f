// Bandwidth-bound kernels around the GEMMs: operand repacking, state I/O, x_T fill,
// q_sample, the standalone reverse update, the hoisted condition / time paths.
// All are 128-bit vectorised along the feature axis; rows are independent.
#pragma once
#include <cuda_bf16.h>
#include <cstdint>
#include "philox.cuh"

namespace osteo {

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16r(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }

// fp32 [rows, cols] (ld = src_ld) -> bf16 [dst_rows, dst_ld]: hi at column c, residual at c + lo_off
// (lo_off = 0 skips the residual), zero padded. One thread per 4 destination columns of the hi half.
__global__ void pack_bf16_hilo_kernel(const float* __restrict__ src, long long rows, int cols, long long src_ld,
                                      __nv_bfloat16* __restrict__ dst, long long dst_rows, int kp, long long dst_ld, int lo_off) {
    const long long total = dst_rows * (kp / 4);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / (kp / 4);
        const int c = static_cast<int>(i % (kp / 4)) * 4;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (r < rows && c + j < cols) ? src[r * src_ld + c + j] : 0.0f;
        uint2 hi = make_uint2(pack2(v[0], v[1]), pack2(v[2], v[3]));
        *reinterpret_cast<uint2*>(dst + r * dst_ld + c) = hi;
        if (lo_off > 0) {
            uint2 lo = make_uint2(pack2(v[0] - bf16r(v[0]), v[1] - bf16r(v[1])), pack2(v[2] - bf16r(v[2]), v[3] - bf16r(v[3])));
            *reinterpret_cast<uint2*>(dst + r * dst_ld + lo_off + c) = lo;
        }
    }
}

// Same, but the source is read transposed: dst[r, c] = src[c, r] (src is [cols, rows] with ld src_ld).
// Used for the dgrad operand W^T. Tiled through shared memory so both sides are coalesced.
__global__ void pack_bf16_hilo_transposed_kernel(const float* __restrict__ src, int rows, int cols, long long src_ld,
                                                 __nv_bfloat16* __restrict__ dst, int dst_rows, int kp, long long dst_ld, int lo_off) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, r = r0 + threadIdx.x;   // src element (c, r)
        tile[j][threadIdx.x] = (c < cols && r < rows) ? src[static_cast<long long>(c) * src_ld + r] : 0.0f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        if (r < dst_rows && c < kp) {
            const float v = tile[threadIdx.x][j];
            dst[static_cast<long long>(r) * dst_ld + c] = __float2bfloat16_rn(v);
            if (lo_off > 0) dst[static_cast<long long>(r) * dst_ld + lo_off + c] = __float2bfloat16_rn(v - bf16r(v));
        }
    }
}

// Blocked state layouts (each TMA box is one contiguous 16 KB chunk, so a tile load walks DRAM pages sequentially instead of
// touching 128 pages 20 KB apart):
//   fp32 state  [m_tile][x_nbox ][128 rows][32 cols]     bf16 shadow [m_tile][xb_nbox][128 rows][64 cols]
// The fused bf16 step (fused_step.cuh) uses 8-column boxes instead ("c8": [m_tile][DP / 8][128 rows][8 cols], x_shift = 3), so that a
// warp-wide 256-bit access of 32 consecutive rows is 1 KB contiguous; the TMA-staged path keeps 32-column boxes (x_shift = 5).
__host__ __device__ __forceinline__ size_t x_blocked_off(long long r, int c, int x_nbox, int x_shift) {
    return (((static_cast<size_t>(r >> 7) * x_nbox + (c >> x_shift)) * 128 + (r & 127)) << x_shift) + (c & ((1 << x_shift) - 1));
}
__host__ __device__ __forceinline__ size_t xb_blocked_off(long long r, int c, int xb_nbox) {
    return ((static_cast<size_t>(r >> 7) * xb_nbox + (c >> 6)) * 128 + (r & 127)) * 64 + (c & 63);
}

// Caller x [n, d] -> padded fp32 state [n, x_ld] + bf16 shadow [n, xb_ld] (hi | lo). One thread per 4 columns.
__global__ void load_state_kernel(const float* __restrict__ src, long long n, int d, int dp, float* __restrict__ x, int x_nbox, int x_shift,
                                  __nv_bfloat16* __restrict__ xb, int xb_nbox, int lo_boxes) {
    const int q = dp / 4;
    const long long total = n * q;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / q;
        const int c = static_cast<int>(i % q) * 4;
        float v[4];
        if ((d & 1) == 0 && c + 3 < d) {
            // even row pitch: every (row, even column) is 8-byte aligned -> two 64-bit loads instead of four 32-bit ones
            const float2 a = *reinterpret_cast<const float2*>(src + r * d + c), b = *reinterpret_cast<const float2*>(src + r * d + c + 2);
            v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (c + j < d) ? src[r * d + c + j] : 0.0f;
        }
        if (x) *reinterpret_cast<float4*>(x + x_blocked_off(r, c, x_nbox, x_shift)) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<uint2*>(xb + xb_blocked_off(r, c, xb_nbox)) = make_uint2(pack2(v[0], v[1]), pack2(v[2], v[3]));
        if (lo_boxes > 0)
            *reinterpret_cast<uint2*>(xb + xb_blocked_off(r, c + lo_boxes * 64, xb_nbox)) =
                make_uint2(pack2(v[0] - bf16r(v[0]), v[1] - bf16r(v[1])), pack2(v[2] - bf16r(v[2]), v[3] - bf16r(v[3])));
    }
}

// bf16 shadow rebuilt from the fp32 state (hi half only): needed when the fused step, which does not maintain the shadow, is
// followed by a step that has to recompute input_proj from scratch (a different timestep, new conditions or new weights).
__global__ void reshadow_kernel(const float* __restrict__ x, int x_nbox, int x_shift, int dp, __nv_bfloat16* __restrict__ xb, int xb_nbox, long long n) {
    const int q = dp / 4;
    const long long total = n * q;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / q;
        const int c = static_cast<int>(i % q) * 4;
        const float4 v = *reinterpret_cast<const float4*>(x + x_blocked_off(r, c, x_nbox, x_shift));
        *reinterpret_cast<uint2*>(xb + xb_blocked_off(r, c, xb_nbox)) = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
    }
}

__global__ void store_state_kernel(const float* __restrict__ x, int x_nbox, int x_shift, float* __restrict__ dst, long long n, int d) {
    const int q = (d + 3) / 4;       // one thread per 4 columns: a 128-bit read of the blocked state, 64-bit writes when the row pitch is even
    const long long total = n * q;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / q;
        const int c = static_cast<int>(i % q) * 4;
        const float4 v = *reinterpret_cast<const float4*>(x + x_blocked_off(r, c, x_nbox, x_shift));
        float* o = dst + r * d + c;
        if ((d & 1) == 0 && c + 3 < d) {
            *reinterpret_cast<float2*>(o) = make_float2(v.x, v.y);
            *reinterpret_cast<float2*>(o + 2) = make_float2(v.z, v.w);
        } else {
            const float vv[4] = {v.x, v.y, v.z, v.w};
            for (int j = 0; j < 4; ++j)
                if (c + j < d) o[j] = vv[j];
        }
    }
}

// Egress of a finished cohort in the form SyntheticPatientGenerator.generate consumes it (utils/generate.py:130-135): the mutation block
// thresholded (x > thr) into 0/1 bytes and/or one bit per gene (LSB first, ceil(mut / 8) bytes per patient), and the remaining
// expression + pathway-score columns as a dense fp32 matrix [n, d - mut].
__global__ void store_split_kernel(const float* __restrict__ x, int x_nbox, int x_shift, long long n, int d, int mut, float thr, uint8_t* __restrict__ calls,
                                   uint8_t* __restrict__ bits, float* __restrict__ rest) {
    const int bpr = (mut + 7) / 8;
    const int rest_w = d - mut;
    const int per_row = bpr + rest_w;            // work items per patient: one per packed byte of calls, one per remaining column
    const long long total = n * per_row;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / per_row;
        const int j = static_cast<int>(i % per_row);
        if (j < bpr) {
            uint32_t b = 0;
            for (int k = 0; k < 8; ++k) {
                const int c = j * 8 + k;
                if (c >= mut) break;
                const bool on = x[x_blocked_off(r, c, x_nbox, x_shift)] > thr;
                if (calls) calls[r * mut + c] = on ? 1 : 0;
                b |= (on ? 1u : 0u) << k;
            }
            if (bits) bits[r * bpr + j] = static_cast<uint8_t>(b);
        } else if (rest) {
            const int c = mut + (j - bpr);
            rest[r * rest_w + (j - bpr)] = x[x_blocked_off(r, c, x_nbox, x_shift)];
        }
    }
}

// Training ingress: one batch of a device-resident dataset with the reference's mixup fused in
// (utils/train.py:85-126: mixed = lam * batch + (1 - lam) * batch[index]):
//   out[i, :] = lam * src[ia[i], :] + oml * src[ib[i], :]        ia = the batch's dataset rows, ib = ia permuted
// lam and oml = 1 - lam arrive as the fp32 roundings torch makes of the Python scalars, and the two products and the sum are rounded
// separately like torch's three elementwise kernels: bit-exact. ib == nullptr: plain gather (no mixup / validation batches).
__global__ void mixup_gather_kernel(const float* __restrict__ src, int d, const long long* __restrict__ ia, const long long* __restrict__ ib, float lam, float oml,
                                    float* __restrict__ out, long long n) {
    const bool vec = (d & 1) == 0;                 // even row pitch: 64-bit accesses
    const int q = vec ? d / 2 : d;
    const long long total = n * q;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / q;
        const int c = static_cast<int>(i % q);
        const long long ra = ia ? ia[r] : r;
        if (vec) {
            const float2 a = reinterpret_cast<const float2*>(src + ra * d)[c];
            float2 o = a;
            if (ib) {
                const float2 b = reinterpret_cast<const float2*>(src + ib[r] * d)[c];
                o = make_float2(__fadd_rn(__fmul_rn(lam, a.x), __fmul_rn(oml, b.x)), __fadd_rn(__fmul_rn(lam, a.y), __fmul_rn(oml, b.y)));
            }
            reinterpret_cast<float2*>(out + r * d)[c] = o;
        } else {
            const float a = src[ra * d + c];
            out[r * d + c] = ib ? __fadd_rn(__fmul_rn(lam, a), __fmul_rn(oml, src[ib[r] * d + c])) : a;
        }
    }
}

// x_T ~ N(0, I): models/diffusion.py:443. One Philox call per 4 columns.
__global__ void init_noise_kernel(float* __restrict__ x, int x_nbox, int x_shift, int dp, __nv_bfloat16* __restrict__ xb, int xb_nbox, int lo_boxes, long long n, int d,
                                  unsigned long long seed, long long row_base, uint32_t stream_id, uint32_t step) {
    const int q = dp / 4;
    const long long total = n * q;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / q;
        const int c4 = static_cast<int>(i % q);
        const int c = c4 * 4;
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < d) {
            z = philox_normal4(seed, static_cast<uint64_t>(row_base + r), static_cast<uint32_t>(c4), stream_id, step);
            if (c + 1 >= d) z.y = 0.f;
            if (c + 2 >= d) z.z = 0.f;
            if (c + 3 >= d) z.w = 0.f;
        }
        *reinterpret_cast<float4*>(x + x_blocked_off(r, c, x_nbox, x_shift)) = z;
        *reinterpret_cast<uint2*>(xb + xb_blocked_off(r, c, xb_nbox)) = make_uint2(pack2(z.x, z.y), pack2(z.z, z.w));
        if (lo_boxes > 0)
            *reinterpret_cast<uint2*>(xb + xb_blocked_off(r, c + lo_boxes * 64, xb_nbox)) =
                make_uint2(pack2(z.x - bf16r(z.x), z.y - bf16r(z.y)), pack2(z.z - bf16r(z.z), z.w - bf16r(z.w)));
    }
}

// Dense test hooks for the RNG.
__global__ void philox_normal_kernel(float* __restrict__ out, long long n, int d, unsigned long long seed, long long row_base, uint32_t stream_id, uint32_t step) {
    const int q = (d + 3) / 4;
    const long long total = n * q;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / q;
        const int c4 = static_cast<int>(i % q);
        const float4 z = philox_normal4(seed, static_cast<uint64_t>(row_base + r), static_cast<uint32_t>(c4), stream_id, step);
        const float zz[4] = {z.x, z.y, z.z, z.w};
        for (int j = 0; j < 4; ++j)
            if (c4 * 4 + j < d) out[r * d + c4 * 4 + j] = zz[j];
    }
}
__global__ void philox_words_kernel(uint32_t* __restrict__ out, long long n, int ncol4, unsigned long long seed, long long row_base, uint32_t stream_id, uint32_t step) {
    const long long total = n * ncol4;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / ncol4;
        const int c4 = static_cast<int>(i % ncol4);
        const uint4 w = philox_words_stream(seed, static_cast<uint64_t>(row_base + r), static_cast<uint32_t>(c4), stream_id, step);
        reinterpret_cast<uint4*>(out)[i] = w;
    }
}

// q_sample (models/diffusion.py:337-340): xt = sqrt_ab[t_r]*x0 + sqrt_1mab[t_r]*noise, dense [n, d] tensors.
// GEN: fill `noise` from Philox first (models/diffusion.py:335).
template <bool GEN>
__global__ void q_sample_kernel(const float* __restrict__ x0, const int* __restrict__ t_idx, float* __restrict__ noise, float* __restrict__ xt,
                                long long n, int d, const float* __restrict__ sqrt_ab, const float* __restrict__ sqrt_1mab,
                                unsigned long long seed, long long row_base, uint32_t salt) {
    const int q = (d + 3) / 4;
    const long long total = n * q;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / q;
        const int c4 = static_cast<int>(i % q);
        const int t = t_idx[r];
        const float a = __ldg(sqrt_ab + t), b = __ldg(sqrt_1mab + t);
        float z[4];
        if (GEN) {
            const float4 zz = philox_normal4(seed, static_cast<uint64_t>(row_base + r), static_cast<uint32_t>(c4), STREAM_QNOISE, salt);
            z[0] = zz.x; z[1] = zz.y; z[2] = zz.z; z[3] = zz.w;
        }
        const int c0 = c4 * 4;
        if ((d & 1) == 0 && c0 + 3 < d) {
            // even row pitch: 64-bit accesses (every (row, even column) is 8-byte aligned)
            const long long o = r * d + c0;
            const float2 xa = *reinterpret_cast<const float2*>(x0 + o), xb2 = *reinterpret_cast<const float2*>(x0 + o + 2);
            if (GEN) {
                *reinterpret_cast<float2*>(noise + o) = make_float2(z[0], z[1]);
                *reinterpret_cast<float2*>(noise + o + 2) = make_float2(z[2], z[3]);
            } else {
                const float2 na = *reinterpret_cast<const float2*>(noise + o), nb = *reinterpret_cast<const float2*>(noise + o + 2);
                z[0] = na.x; z[1] = na.y; z[2] = nb.x; z[3] = nb.y;
            }
            // two rounded products then a rounded add, as the reference's elementwise ops do (bit-exact)
            *reinterpret_cast<float2*>(xt + o) = make_float2(__fadd_rn(__fmul_rn(a, xa.x), __fmul_rn(b, z[0])), __fadd_rn(__fmul_rn(a, xa.y), __fmul_rn(b, z[1])));
            *reinterpret_cast<float2*>(xt + o + 2) = make_float2(__fadd_rn(__fmul_rn(a, xb2.x), __fmul_rn(b, z[2])), __fadd_rn(__fmul_rn(a, xb2.y), __fmul_rn(b, z[3])));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + j;
                if (c < d) {
                    const long long o = r * d + c;
                    if (GEN) noise[o] = z[j]; else z[j] = noise[o];
                    xt[o] = __fadd_rn(__fmul_rn(a, x0[o]), __fmul_rn(b, z[j]));
                }
            }
        }
    }
}

// Standalone reverse update on dense [n, d] tensors (models/diffusion.py:400-423 collapsed).
__global__ void reverse_update_kernel(float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ z, long long n, int d,
                                      float cx, float ce, float sg, unsigned long long seed, long long row_base, uint32_t t) {
    const int q = (d + 3) / 4;
    const long long total = n * q;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / q;
        const int c4 = static_cast<int>(i % q);
        float zz[4] = {0.f, 0.f, 0.f, 0.f};
        if (sg != 0.0f && !z) {
            const float4 g = philox_normal4(seed, static_cast<uint64_t>(row_base + r), static_cast<uint32_t>(c4), STREAM_REVERSE, t);
            zz[0] = g.x; zz[1] = g.y; zz[2] = g.z; zz[3] = g.w;
        }
        const int c0 = c4 * 4;
        if ((d & 1) == 0 && c0 + 3 < d) {
            // even row pitch: 64-bit accesses (every (row, even column) is 8-byte aligned); same arithmetic as the scalar tail
            const long long o = r * d + c0;
            if (sg != 0.0f && z) {
                const float2 za = *reinterpret_cast<const float2*>(z + o), zb = *reinterpret_cast<const float2*>(z + o + 2);
                zz[0] = za.x; zz[1] = za.y; zz[2] = zb.x; zz[3] = zb.y;
            }
            const float2 xa = *reinterpret_cast<const float2*>(x + o), xb2 = *reinterpret_cast<const float2*>(x + o + 2);
            const float2 ea = *reinterpret_cast<const float2*>(eps + o), eb = *reinterpret_cast<const float2*>(eps + o + 2);
            *reinterpret_cast<float2*>(x + o) = make_float2(fmaf(sg, zz[0], fmaf(cx, xa.x, -ce * ea.x)), fmaf(sg, zz[1], fmaf(cx, xa.y, -ce * ea.y)));
            *reinterpret_cast<float2*>(x + o + 2) = make_float2(fmaf(sg, zz[2], fmaf(cx, xb2.x, -ce * eb.x)), fmaf(sg, zz[3], fmaf(cx, xb2.y, -ce * eb.y)));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + j;
                if (c < d) {
                    const long long o = r * d + c;
                    const float zv = (sg != 0.0f && z) ? z[o] : zz[j];
                    x[o] = fmaf(sg, zv, fmaf(cx, x[o], -ce * eps[o]));
                }
            }
        }
    }
}

// time_proj(time_embed(t / T)) for every integer t (models/diffusion.py:222-223): table[t, n] = emb[t, :] . W[n, :] + b[n].
// wt is the TRANSPOSED weight [td, h0] (coalesced across n); one block = 8 timesteps, each thread owns output features.
__global__ void time_table_kernel(const float* __restrict__ emb, const float* __restrict__ wt, const float* __restrict__ b, float* __restrict__ table, int T, int td, int h0) {
    constexpr int R = 8;
    extern __shared__ float e[];          // [R, td]
    const int t0 = blockIdx.x * R;
    for (int i = threadIdx.x; i < R * td; i += blockDim.x) {
        const int t = t0 + i / td;
        e[i] = t < T ? emb[static_cast<long long>(t) * td + i % td] : 0.0f;
    }
    __syncthreads();
    for (int n = threadIdx.x; n < h0; n += blockDim.x) {
        float acc[R];
        const float bn = b[n];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = bn;
        for (int k = 0; k < td; ++k) {
            const float w = wt[static_cast<long long>(k) * h0 + n];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = fmaf(e[r * td + k], w, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (t0 + r < T) table[static_cast<long long>(t0 + r) * h0 + n] = acc[r];
    }
}

// dst[c, r] = src[r, c] (tiny weight matrices; run once per set_weights)
__global__ void transpose_f32_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * cols) dst[(i % cols) * rows + i / cols] = src[i];
}

// ConditionalEmbedding (Linear C->E, SiLU, Linear E->E; models/diffusion.py:101-114) followed by
// cond_proj (Linear E->h0; :226). fp32 on CUDA cores: once per sample() call, once per training step. One block = 16 rows.
// The three weight matrices are passed TRANSPOSED ([in, out]) so that consecutive threads (consecutive output features)
// read consecutive addresses. Optionally keeps the intermediate activations for the training backward.
__global__ void cond_path_kernel(const float* __restrict__ cond, long long n, int C, int E, int h0,
                                 const float* __restrict__ w0t, const float* __restrict__ b0, const float* __restrict__ w2t, const float* __restrict__ b2,
                                 const float* __restrict__ wct, const float* __restrict__ bc, float* __restrict__ cproj,
                                 float* __restrict__ save_pre0, float* __restrict__ save_emb) {
    constexpr int R = 16;
    extern __shared__ float sm[];
    float* s_c = sm;                 // [R, C]
    float* s_h = s_c + R * C;        // [R, E]
    float* s_e = s_h + R * E;        // [R, E]
    const long long r0 = static_cast<long long>(blockIdx.x) * R;
    for (int i = threadIdx.x; i < R * C; i += blockDim.x) {
        const long long r = r0 + i / C;
        s_c[i] = r < n ? cond[r * C + i % C] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < R * E; i += blockDim.x) {
        const int rr = i / E, j = i % E;
        float a = b0[j];
        for (int k = 0; k < C; ++k) a = fmaf(s_c[rr * C + k], w0t[k * E + j], a);
        if (save_pre0 && r0 + rr < n) save_pre0[(r0 + rr) * E + j] = a;
        s_h[i] = a / (1.0f + expf(-a));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < R * E; i += blockDim.x) {
        const int rr = i / E, j = i % E;
        float a = b2[j];
#pragma unroll 8
        for (int k = 0; k < E; ++k) a = fmaf(s_h[rr * E + k], w2t[k * E + j], a);
        if (save_emb && r0 + rr < n) save_emb[(r0 + rr) * E + j] = a;
        s_e[i] = a;
    }
    __syncthreads();
    // each thread owns one output feature j for all R rows: the weight column is read once per block
    for (int j = threadIdx.x; j < h0; j += blockDim.x) {
        float acc[R];
        const float bj = bc[j];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) acc[rr] = bj;
        for (int k = 0; k < E; ++k) {
            const float w = wct[k * h0 + j];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) acc[rr] = fmaf(s_e[rr * E + k], w, acc[rr]);
        }
#pragma unroll
        for (int rr = 0; rr < R; ++rr)
            if (r0 + rr < n) cproj[(r0 + rr) * h0 + j] = acc[rr];
    }
}

// bf16 [hi | lo] pair -> fp32 (hi + lo). Test helper for the GroupNorm epilogue.
__global__ void unpack_hilo_kernel(const __nv_bfloat16* __restrict__ src, long long ld, int lo_off, float* __restrict__ dst, long long m, int n) {
    const long long total = m * n;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / n;
        const int c = static_cast<int>(i % n);
        dst[i] = __bfloat162float(src[r * ld + c]) + __bfloat162float(src[r * ld + lo_off + c]);
    }
}

// one thread per word (the per-branch step words of the sampling graphs)
__global__ void set_int_kernel(int* p, int v) { p[threadIdx.x] = v; }
__global__ void add_int_kernel(int* p, int v) { p[threadIdx.x] += v; }
__global__ void set_u64_kernel(unsigned long long* p, unsigned long long v) { *p = v; }
__global__ void finish_loss_kernel(const double* acc, float* out, double scale) { *out = static_cast<float>(*acc * scale); }

}  // namespace osteo
