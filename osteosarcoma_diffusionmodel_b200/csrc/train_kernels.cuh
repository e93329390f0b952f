// Training-step workspace and the CUDA-core kernels of the backward pass that are not GEMMs:
// q_sample fused with operand packing, column-partial reduction, and the backward of the tiny
// condition / time embedding paths. The heavy lifting (dgrad / wgrad / GroupNorm backward) is
// in gemm_tc.cuh; the orchestration is api_train.inl.
#pragma once
#include <memory>
#include <vector>
#include "common.cuh"
#include "elem_kernels.cuh"

namespace osteo {

// Tensors saved by the training forward for the backward pass.
struct TrainWorkspace {
    long long cap = 0;
    long long fwd_n = -1;                        // rows of the pending forward pass (two-phase API), -1 = none
    bool fwd_graphed = false;                    // that forward was graph-replayed: the backward half may use the library's t / cond copies
    // per half block
    std::vector<std::unique_ptr<DevBuf>> xhat;   // bf16 [cap, 2*n]  normalised pre-affine activations [hi|lo]
    std::vector<std::unique_ptr<DevBuf>> rstd;   // fp32 [cap, 8]
    std::vector<std::unique_ptr<DevBuf>> dy;     // bf16 [cap, 2*n]  d(loss)/d(pre-norm Linear output) [hi|lo]
    std::vector<CUtensorMap> dy_tmap;            // K-major A-operand maps of dy (dgrad)
    DevBuf dh0_bf, dh0_f32;                      // d(loss)/d(h0): bf16 [cap, 2*h0], fp32 [cap, h0]
    DevBuf deps;                                 // bf16 [cap, 2*DP]  d(loss)/d(eps_hat)
    CUtensorMap deps_tmap;
    DevBuf xt_bf;                                // bf16 [cap, 2*DP]  x_t [hi|lo], row-major: A operand of input_proj AND MN-major wgrad operand
    CUtensorMap xt_tmap;
    DevBuf noise;                                // fp32 [cap, DP]    target of the MSE epilogue
    DevBuf pre0, cemb, h1, dcemb, dpre0;         // fp32 [cap, E] each: condition-embedding forward saves / backward temporaries
    std::vector<std::unique_ptr<DevBuf>> partials;   // fp32 [cap/32, nq, width] column partials, one buffer per producer (MSE epilogue, each
                                                     // GroupNorm-backward dgrad, d(h0)) so that their reductions can run beside the next dgrad
    // side streams of the backward pass: weight / bias gradients run beside the dgrad chain (api_train.inl)
    cudaStream_t side[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    DevBuf colsum_tmp;                           // fp32 [E] scratch
    DevBuf temb_bf, cemb_bf;                     // bf16 [cap, 2*TD], [cap, 2*E]: gathered time embedding / condition embedding as wgrad operands
    DevBuf t_copy, cond_copy, loss_tmp;          // graph-replayed step: library-owned copies of t [cap] / cond [cap, C], and the loss scalar
    size_t bytes() const {
        size_t b = dh0_bf.bytes + dh0_f32.bytes + deps.bytes + xt_bf.bytes + noise.bytes + pre0.bytes + cemb.bytes + h1.bytes + dcemb.bytes + dpre0.bytes;
        for (auto& p : partials) b += p->bytes;
        for (auto& p : xhat) b += p->bytes;
        for (auto& p : rstd) b += p->bytes;
        for (auto& p : dy) b += p->bytes;
        return b;
    }
    void release() {
        xhat.clear();
        rstd.clear();
        dy.clear();
        dy_tmap.clear();
        partials.clear();
        for (DevBuf* b : {&dh0_bf, &dh0_f32, &deps, &xt_bf, &noise, &pre0, &cemb, &h1, &dcemb, &dpre0, &colsum_tmp, &t_copy, &cond_copy, &loss_tmp, &temb_bf, &cemb_bf}) b->release();
        for (int i = 0; i < 2; ++i) {
            if (side[i]) cudaStreamDestroy(side[i]);
            if (ev_join[i]) cudaEventDestroy(ev_join[i]);
            side[i] = nullptr;
            ev_join[i] = nullptr;
        }
        if (ev_fork) cudaEventDestroy(ev_fork);
        ev_fork = nullptr;
        cap = 0;
    }
};

// q_sample (models/diffusion.py:337-340) fused with the packing the GEMMs need:
//   noise (injected, or Philox when noise_in == nullptr) -> fp32 [n, ld] (target of the MSE epilogue)
//   x_t = sqrt_ab[t]*x0 + sqrt_1mab[t]*noise            -> bf16 [hi|lo] shadow (A operand of input_proj)
__global__ void train_prepare_kernel(const float* __restrict__ x0, const float* __restrict__ noise_in, const int* __restrict__ t_idx, long long n, int d,
                                     const float* __restrict__ sqrt_ab, const float* __restrict__ sqrt_1mab, float* __restrict__ noise_out, int ld,
                                     __nv_bfloat16* __restrict__ xb, int xb_ld, int lo_off, unsigned long long seed, long long row_base) {
    const int q = ld / 4;
    const long long total = n * q;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / q;
        const int c4 = static_cast<int>(i % q);
        const int c = c4 * 4;
        const int t = t_idx[r];
        const float a = __ldg(sqrt_ab + t), b = __ldg(sqrt_1mab + t);
        float z[4] = {0.f, 0.f, 0.f, 0.f}, xt[4] = {0.f, 0.f, 0.f, 0.f};
        if (c < d) {
            if (!noise_in) {
                const float4 g = philox_normal4(seed, static_cast<uint64_t>(row_base + r), static_cast<uint32_t>(c4), STREAM_QNOISE, 0u);
                z[0] = g.x; z[1] = g.y; z[2] = g.z; z[3] = g.w;
            }
            if ((d & 1) == 0 && c + 3 < d) {
                // even row pitch: 64-bit loads of the caller's rows (every (row, even column) is 8-byte aligned)
                const float2 xa = *reinterpret_cast<const float2*>(x0 + r * d + c), xb2 = *reinterpret_cast<const float2*>(x0 + r * d + c + 2);
                if (noise_in) {
                    const float2 na = *reinterpret_cast<const float2*>(noise_in + r * d + c), nb = *reinterpret_cast<const float2*>(noise_in + r * d + c + 2);
                    z[0] = na.x; z[1] = na.y; z[2] = nb.x; z[3] = nb.y;
                }
                xt[0] = __fadd_rn(__fmul_rn(a, xa.x), __fmul_rn(b, z[0]));
                xt[1] = __fadd_rn(__fmul_rn(a, xa.y), __fmul_rn(b, z[1]));
                xt[2] = __fadd_rn(__fmul_rn(a, xb2.x), __fmul_rn(b, z[2]));
                xt[3] = __fadd_rn(__fmul_rn(a, xb2.y), __fmul_rn(b, z[3]));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (c + j < d) {
                        if (noise_in) z[j] = noise_in[r * d + c + j];
                        xt[j] = __fadd_rn(__fmul_rn(a, x0[r * d + c + j]), __fmul_rn(b, z[j]));
                    } else {
                        z[j] = 0.f;
                    }
                }
            }
        }
        *reinterpret_cast<float4*>(noise_out + r * ld + c) = make_float4(z[0], z[1], z[2], z[3]);
        *reinterpret_cast<uint2*>(xb + r * xb_ld + c) = make_uint2(pack2(xt[0], xt[1]), pack2(xt[2], xt[3]));
        if (lo_off > 0)
            *reinterpret_cast<uint2*>(xb + r * xb_ld + lo_off + c) =
                make_uint2(pack2(xt[0] - bf16r(xt[0]), xt[1] - bf16r(xt[1])), pack2(xt[2] - bf16r(xt[2]), xt[3] - bf16r(xt[3])));
    }
}

// out_q[c] = sum over slabs of part[slab, q, c]. grid = (ceil(N/32), nq), block = (32, 8).
__global__ void partials_finish_kernel(const float* __restrict__ part, int slabs, int nq, int N, float* out0, float* out1, float* out2) {
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int q = blockIdx.y;
    float acc = 0.0f;
    if (c < N)
        for (int s = threadIdx.y; s < slabs; s += 8) acc += part[(static_cast<size_t>(s) * nq + q) * N + c];
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < N) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
        float* out = q == 0 ? out0 : (q == 1 ? out1 : out2);
        if (out) out[c] = t;
    }
}

// fp32 rows (optionally gathered through `idx`: the sinusoidal time-embedding table, models/diffusion.py:124-139) -> bf16 [n, dst_ld] with the
// hi part at column c and the residual at c + lo_off (lo_off = 0: none): the MN-major X operand of a tensor-core weight gradient.
__global__ void pack_rows_hilo_kernel(const float* __restrict__ src, int src_ld, const int* __restrict__ idx, long long n, int k, __nv_bfloat16* __restrict__ dst,
                                      int dst_ld, int lo_off) {
    const int q = k / 4;
    const long long total = n * q;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / q;
        const int c = static_cast<int>(i % q) * 4;
        const long long sr = idx ? __ldg(idx + r) : r;
        const float4 v = *reinterpret_cast<const float4*>(src + sr * src_ld + c);
        *reinterpret_cast<uint2*>(dst + r * dst_ld + c) = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
        if (lo_off > 0)
            *reinterpret_cast<uint2*>(dst + r * dst_ld + lo_off + c) =
                make_uint2(pack2(v.x - bf16r(v.x), v.y - bf16r(v.y)), pack2(v.z - bf16r(v.z), v.w - bf16r(v.w)));
    }
}

// Column sums of a dense fp32 matrix g [n, m] -> out[m] (atomic accumulate; out zeroed by the caller).
__global__ void colsum_f32_kernel(const float* __restrict__ g, long long n, int m, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const long long rows_per = (n + gridDim.y - 1) / gridDim.y;
    const long long r0 = blockIdx.y * rows_per, r1 = r0 + rows_per < n ? r0 + rows_per : n;
    float acc = 0.0f;
    for (long long r = r0; r < r1; ++r) acc += g[r * m + c];
    atomicAdd(out + c, acc);
}

// dW[gm, xk] += G^T X over rows: G fp32 [n, gm], X fp32 [n_x, xk] read through an optional row index (time table
// gather). One block = 64 rows staged in shared memory; each thread owns a strided set of dW entries.
__global__ void outer_accum_kernel(const float* __restrict__ G, int gm, const float* __restrict__ X, int xk, const int* __restrict__ x_idx, long long n,
                                   float* __restrict__ dW) {
    constexpr int R = 64;
    extern __shared__ float sm[];
    float* sg = sm;             // [R, gm]
    float* sx = sm + R * gm;    // [R, xk]
    const long long r0 = static_cast<long long>(blockIdx.x) * R;
    const int rows = static_cast<int>((n - r0) < R ? (n - r0) : R);
    // staging: eight independent loads per thread in flight before the first store (a load -> store loop keeps ONE in flight, and the
    // 96 dependent DRAM / L2 round trips of such a loop were most of this kernel's 71 us)
    for (int i0 = threadIdx.x; i0 < R * gm; i0 += 8 * blockDim.x) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * blockDim.x;
            const int rr = i / gm;
            t[u] = (i < R * gm && rr < rows) ? __ldg(G + (r0 + rr) * gm + i % gm) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i < R * gm) sg[i] = t[u];
        }
    }
    for (int i0 = threadIdx.x; i0 < R * xk; i0 += 8 * blockDim.x) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * blockDim.x;
            const int rr = i / xk;
            float v = 0.0f;
            if (i < R * xk && rr < rows) {
                const long long xr = x_idx ? __ldg(x_idx + r0 + rr) : (r0 + rr);
                v = __ldg(X + xr * xk + i % xk);
            }
            t[u] = v;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i < R * xk) sx[i] = t[u];
        }
    }
    __syncthreads();
    if ((gm & 3) == 0 && (xk & 7) == 0) {
        // register-tiled: a thread owns 4 x 8 patches of dW, three 128-bit shared loads feed 32 FMAs per staged row (one entry per thread
        // and two scalar shared loads per FMA made the 256 x 128 time_proj gradient shared-memory bound: 111 us)
        const int tb = xk >> 3;
        for (int tile = threadIdx.x; tile < (gm >> 2) * tb; tile += blockDim.x) {
            const int a0 = (tile / tb) << 2, b0 = (tile % tb) << 3;
            float acc[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
#pragma unroll 4
            for (int rr = 0; rr < R; ++rr) {
                const float4 g = *reinterpret_cast<const float4*>(sg + rr * gm + a0);
                const float4 x0 = *reinterpret_cast<const float4*>(sx + rr * xk + b0);
                const float4 x1 = *reinterpret_cast<const float4*>(sx + rr * xk + b0 + 4);
                const float gg[4] = {g.x, g.y, g.z, g.w};
                const float xx[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(gg[i], xx[j], acc[i][j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) atomicAdd(dW + (a0 + i) * xk + b0 + j, acc[i][j]);
        }
        return;
    }
    for (int e = threadIdx.x; e < gm * xk; e += blockDim.x) {
        const int a = e / xk, b = e % xk;
        float acc = 0.0f;
#pragma unroll 8
        for (int rr = 0; rr < R; ++rr) acc = fmaf(sg[rr * gm + a], sx[rr * xk + b], acc);
        atomicAdd(dW + e, acc);
    }
}

// Per-row backward through cond_proj and ConditionalEmbedding (models/diffusion.py:101-114, :226):
//   dcemb = dh0 . Wc ; h1 = silu(pre0) ; dh1 = dcemb . W2 ; dpre0 = dh1 * silu'(pre0).  One block = 8 rows.
__global__ void cond_bwd_rows_kernel(const float* __restrict__ dh0, long long n, int h0, int E, const float* __restrict__ wc, const float* __restrict__ w2,
                                     const float* __restrict__ pre0, float* __restrict__ dcemb, float* __restrict__ h1, float* __restrict__ dpre0) {
    constexpr int R = 8;
    extern __shared__ float sm[];
    float* s_d = sm;              // [R, h0]
    float* s_e = s_d + R * h0;    // [R, E]
    const long long r0 = static_cast<long long>(blockIdx.x) * R;
    for (int i = threadIdx.x; i < R * h0; i += blockDim.x) {
        const long long r = r0 + i / h0;
        s_d[i] = r < n ? dh0[r * h0 + i % h0] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < R * E; i += blockDim.x) {
        const int rr = i / E, j = i % E;
        float a = 0.0f;
        for (int k = 0; k < h0; ++k) a = fmaf(s_d[rr * h0 + k], wc[k * E + j], a);   // Wc is [h0, E]
        s_e[i] = a;
        if (r0 + rr < n) dcemb[(r0 + rr) * E + j] = a;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < R * E; i += blockDim.x) {
        const int rr = i / E, j = i % E;
        if (r0 + rr >= n) continue;
        float a = 0.0f;
        for (int k = 0; k < E; ++k) a = fmaf(s_e[rr * E + k], w2[k * E + j], a);     // W2 is [E_out, E_in]: dh1[j] = sum_k dcemb[k] W2[k, j]
        const float p0 = pre0[(r0 + rr) * E + j];
        const float sg = 1.0f / (1.0f + expf(-p0));
        h1[(r0 + rr) * E + j] = p0 * sg;
        dpre0[(r0 + rr) * E + j] = a * sg * (1.0f + p0 * (1.0f - sg));
    }
}

__global__ void zero_double_kernel(double* p) { *p = 0.0; }

// ---------------------------------------------------------------------------------------------------------------------------
// clip_grad_norm_(max_norm) + AdamW.step() over a list of tensors in two launches (utils/train.py:242-244 run the torch foreach
// versions: ~25 launches and 0.9 ms of host time per step for 52 tensors). Work is cut into chunks of OPT_CHUNK elements of one
// tensor each (multi-tensor apply): chunk c = (tensor id, first element).
constexpr int OPT_CHUNK = 4096;
struct OptTables {
    float* const* params;
    float* const* grads;
    float* const* exp_avg;
    float* const* exp_avg_sq;
    const long long* numel;
    const int* chunk_tensor;
    const long long* chunk_start;
};

__global__ void opt_sumsq_kernel(OptTables t, double* __restrict__ acc) {
    const int ti = t.chunk_tensor[blockIdx.x];
    const long long s0 = t.chunk_start[blockIdx.x];
    const long long n = t.numel[ti] - s0 < OPT_CHUNK ? t.numel[ti] - s0 : OPT_CHUNK;
    const float* g = t.grads[ti] + s0;
    float a = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a = fmaf(g[i], g[i], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) s += static_cast<double>(red[w]);
        atomicAdd(acc, s);
    }
}

// torch.optim.AdamW (decoupled weight decay, no amsgrad) in torch's operation order, after the gradient has been scaled by
// clip = min(1, max_norm / (||g|| + 1e-6)) as clip_grad_norm_ does (the scaled gradient is written back, as torch does in place).
// decay = 1 - lr * weight_decay, w1 = 1 - beta1, w2 = 1 - beta2 are formed on the host in double and rounded once, as torch's Python scalars are
// (1.0f - 0.999f differs from float(1 - 0.999) by 1.3e-5 relative).
__global__ void opt_adamw_kernel(OptTables t, const double* __restrict__ sumsq, float max_norm, float decay, float w1, float beta2, float w2, float eps,
                                 float step_size, float bc2_sqrt, float* __restrict__ norm_out) {
    float clip = 1.0f;
    if (max_norm > 0.0f) {
        const float total = static_cast<float>(sqrt(*sumsq));
        clip = fminf(max_norm / (total + 1e-6f), 1.0f);
        if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = total;
    }
    const int ti = t.chunk_tensor[blockIdx.x];
    const long long s0 = t.chunk_start[blockIdx.x];
    const long long n = t.numel[ti] - s0 < OPT_CHUNK ? t.numel[ti] - s0 : OPT_CHUNK;
    float* p = t.params[ti] + s0;
    float* g = t.grads[ti] + s0;
    float* m = t.exp_avg[ti] + s0;
    float* v = t.exp_avg_sq[ti] + s0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float gi = max_norm > 0.0f ? __fmul_rn(g[i], clip) : g[i];
        if (max_norm > 0.0f) g[i] = gi;
        float pi = __fmul_rn(p[i], decay);
        const float mi = fmaf(gi - m[i], w1, m[i]);                       // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = fmaf(__fmul_rn(gi, gi), w2, __fmul_rn(v[i], beta2));   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float denom = __fdiv_rn(__fsqrt_rn(vi), bc2_sqrt) + eps;
        pi = fmaf(-step_size, __fdiv_rn(mi, denom), pi);                  // param.addcdiv_(exp_avg, denom, value=-step_size)
        m[i] = mi;
        v[i] = vi;
        p[i] = pi;
    }
}

}  // namespace osteo
