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

// Register-tiled form of the same reduction for sets of at most 16 columns (the Hallmark pathways of validate_pathway_coherence have
// 11-19 member genes, 15 in the benchmark cohort): the warp <-> set kernel above spends 16 shuffles + 16 FMAs per row and set and is
// shuffle-bound at 14 % of HBM. Here a set's 16 x 16 moment matrix is cut into 4 x 4 register blocks on or above the diagonal (10 per
// set); thread <-> (set, block) reads two float4 per row from a compact copy of the gathered, shifted columns in shared memory and does
// 16 FMAs -- no shuffles, and each loaded value feeds 4 FMAs. The block's threads form `groups` identical teams that take alternate
// rows of a chunk; chunks of whole rows are staged with coalesced 128-bit loads, two CTAs per SM overlap one's staging with the
// other's arithmetic. fp32 products over the rows of ONE chunk (shifted, O(1) values), fp64 across chunks; one shared-memory fp64
// reduction per CTA, one set of global atomics per CTA.
//   shift == nullptr: the shift of a column is its value in row 0 of `data` (every rank holds the whole cohort: all ranks agree).
constexpr int CT_THREADS = 512;
constexpr int CT_CHUNK = 32;                     // rows per chunk
constexpr int CT_MAX_STAGES = 4;                 // row buffers (2..4, chosen by the launcher to fit shared memory): stages - 1 chunks in flight per SM
constexpr int CT_SMEM_LIMIT = 226 * 1024;
constexpr int CT_FLUSH = 4;                      // chunks between fp32 -> fp64 flushes of the register blocks
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// One bulk copy (TMA engine, no tensor map) of `bytes` contiguous bytes global -> shared, completion on `bar` (complete_tx). 16-byte aligned
// source, destination and size.
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// NB = 4 x 4 register blocks per dimension: 4 (sets of <= 16 columns, 10 blocks per set) or 8 (<= 32 columns, 36 blocks per set).
// One CTA per SM; the rows of the next CT_STAGES - 1 chunks are in flight while chunk c is gathered and multiplied (with a single chunk
// ahead the loop ran at the latency of one 47 KB fetch per iteration). A chunk of whole contiguous rows is ONE bulk copy issued by one
// thread (cp.async.bulk + mbarrier): as 2 968 per-thread 16-byte cp.async requests the staging alone kept the LSU busy for a large part
// of the 3 800 cycles an iteration took.
template <int NB>
__global__ void __launch_bounds__(CT_THREADS, 1) corr_moments_tiled_kernel(const float* __restrict__ data, int ld, int ncols, const int* __restrict__ cols, int P,
                                                                            const float* __restrict__ shift, long long rb, long long re, int stages, double* __restrict__ out) {
    constexpr int W = 4 * NB;                        // padded set width
    constexpr int NBLK = NB * (NB + 1) / 2;          // blocks on or above the diagonal
    constexpr int SET_STRIDE = 2 + W + W * W;        // shared-memory reduction block per set: count, s1[W], s2[W][W], one pad (keeps what follows 16-byte aligned)
    extern __shared__ __align__(16) uint8_t ct_smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ct_smem);                          // [CT_MAX_STAGES] row buffer landed
    double* red = reinterpret_cast<double*>(ct_smem + 64);                              // [P][SET_STRIDE]
    float* rows_sm = reinterpret_cast<float*>(red + static_cast<size_t>(P) * SET_STRIDE);      // [stages][CT_CHUNK][ncols] (+ pad to 16 bytes per buffer)
    const int buf_floats = (CT_CHUNK * ncols + 3) & ~3;
    float* g_sm = rows_sm + stages * static_cast<size_t>(buf_floats);                     // [CT_CHUNK][P][W]
    int* col_sm = reinterpret_cast<int*>(g_sm + static_cast<size_t>(CT_CHUNK) * P * W);   // [P][W]
    float* shift_sm = reinterpret_cast<float*>(col_sm + P * W);                           // [P][W]
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < CT_MAX_STAGES; ++i) mbar_init(&full_bar[i], 1);
        fence_mbar_init();
    }
    for (int i = tid; i < P * SET_STRIDE; i += CT_THREADS) red[i] = 0.0;
    for (int i = tid; i < P * W; i += CT_THREADS) {
        const int c = cols[(i / W) * 32 + (i % W)];
        col_sm[i] = c;
        shift_sm[i] = c < 0 ? 0.0f : (shift ? shift[(i / W) * 32 + (i % W)] : data[c]);
    }
    const int items = P * NBLK;
    const int groups = CT_THREADS / items;            // >= 1 (checked by the launcher)
    const bool worker = tid < groups * items;
    const int grp = tid / items, item = tid % items;
    const int set = item / NBLK, blk = item % NBLK;
    // block index -> (bi, bj), bi <= bj: row bi of the triangle holds NB - bi blocks
    int bi = 0, rem = blk;
    while (rem >= NB - bi) { rem -= NB - bi; ++bi; }
    const int bj = bi + rem;
    double d2[16], d1[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) d2[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) d1[i] = 0.0;
    // fp32 partial sums over up to CT_FLUSH chunks (<= 32 * CT_FLUSH / groups shifted O(1) products each), then one fp64 add: the
    // fp32 -> fp64 conversions run on the 16-lane XU pipe
    float a2[16], a1[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) a2[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) a1[i] = 0.0f;
    int pending = 0;
    long long count = 0;
    const long long nchunks = (re - rb + CT_CHUNK - 1) / CT_CHUNK;
    // gather roles (see the loop): SL slots per row, rlanes rows in flight
    const int SL = P * W;
    const int rlanes = CT_THREADS / SL > 0 ? CT_THREADS / SL : 1;
    const int gslot = tid % SL, rlane = tid / SL;
    __syncthreads();                                  // col_sm / shift_sm are complete
    const int gcol = col_sm[gslot];
    const float gshift = shift_sm[gslot];
    // contiguous rows whose chunks start on 16-byte boundaries are one bulk copy per chunk; anything else (and a last chunk whose size is not
    // a multiple of 16 bytes) is copied by the threads, and thread 0 completes the barrier phase by hand
    const bool async_ok = ld == ncols && ((reinterpret_cast<uintptr_t>(data + rb * ld) & 15) == 0) && ((static_cast<size_t>(CT_CHUNK) * ncols * 4) & 15) == 0;
    auto stage = [&](int buf, long long c) {
        if (c >= nchunks) return;
        const long long r0 = rb + c * CT_CHUNK;
        const int nr = static_cast<int>(re - r0 < CT_CHUNK ? re - r0 : CT_CHUNK);
        float* dst = rows_sm + static_cast<size_t>(buf) * buf_floats;
        const uint32_t bytes = static_cast<uint32_t>(nr) * static_cast<uint32_t>(ncols) * 4u;
        if (async_ok && (bytes & 15u) == 0) {
            if (tid == 0) {
                mbar_arrive_expect_tx(&full_bar[buf], bytes);
                bulk_load(dst, data + r0 * ld, bytes, &full_bar[buf]);
            }
        } else {
            for (int i = tid; i < nr * ncols; i += CT_THREADS) dst[i] = data[(r0 + i / ncols) * ld + i % ncols];
            __syncthreads();                          // (block-uniform branch) the copy is complete before the phase is
            if (tid == 0) mbar_arrive(&full_bar[buf]);
        }
    };
    __syncthreads();
    for (int st = 0; st < stages - 1; ++st) stage(st, blockIdx.x + static_cast<long long>(st) * gridDim.x);
    int buf = 0;
    uint32_t use = 0;                                 // iteration counter: buffer = use % stages, barrier parity = (use / stages) & 1
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x, buf = buf + 1 == stages ? 0 : buf + 1, ++use) {
        const long long r0 = rb + c * CT_CHUNK;
        const int nr = static_cast<int>(re - r0 < CT_CHUNK ? re - r0 : CT_CHUNK);
        const float* cur = rows_sm + static_cast<size_t>(buf) * buf_floats;
        // the buffer consumed in the previous iteration was released by the barrier that ended it: refill it with chunk c + (STAGES - 1) grid
        stage((buf + stages - 1) % stages, c + static_cast<long long>(stages - 1) * gridDim.x);
        if (!mbar_wait(&full_bar[buf], (use / static_cast<uint32_t>(stages)) & 1u)) __trap();      // chunk c is in `cur`
        // ---- gather + shift into the compact set-major copy: thread <-> one (set, slot) of `rlanes` interleaved rows, its column index and
        // shift in registers (a flat index over rows x slots costs two integer divisions per element)
        if (tid < rlanes * SL) {
            for (int r = rlane; r < nr; r += rlanes) g_sm[r * SL + gslot] = gcol < 0 ? 0.0f : cur[r * ncols + gcol] - gshift;
        }
        __syncthreads();
        // ---- 4 x 4 register blocks
        if (worker) {
            for (int r = grp; r < nr; r += groups) {
                const float4 a = *reinterpret_cast<const float4*>(g_sm + (r * P + set) * W + 4 * bi);
                const float4 b = *reinterpret_cast<const float4*>(g_sm + (r * P + set) * W + 4 * bj);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    a1[i] += av[i];
#pragma unroll
                    for (int j = 0; j < 4; ++j) a2[4 * i + j] = fmaf(av[i], bv[j], a2[4 * i + j]);
                }
                if (blk == 0) ++count;
            }
            if (++pending == CT_FLUSH) {
#pragma unroll
                for (int i = 0; i < 16; ++i) { d2[i] += static_cast<double>(a2[i]); a2[i] = 0.0f; }
#pragma unroll
                for (int i = 0; i < 4; ++i) { d1[i] += static_cast<double>(a1[i]); a1[i] = 0.0f; }
                pending = 0;
            }
        }
        // g_sm (and, after the next iteration's wait, this row buffer) is reused: every worker must be done with this chunk
        __syncthreads();
    }
    if (worker) {
#pragma unroll
        for (int i = 0; i < 16; ++i) d2[i] += static_cast<double>(a2[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) d1[i] += static_cast<double>(a1[i]);
        double* rs = red + static_cast<size_t>(set) * SET_STRIDE;
        if (blk == 0 && count) atomicAdd(rs, static_cast<double>(count));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (bi == bj) atomicAdd(rs + 1 + 4 * bi + i, d1[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(rs + 1 + W + (4 * bi + i) * W + 4 * bj + j, d2[4 * i + j]);
        }
    }
    __syncthreads();
    // ---- one set of global atomics per CTA, into the [count, s1[32], s2[32][32]] layout of the batched kernel (both triangles)
    for (int i = tid; i < P * SET_STRIDE; i += CT_THREADS) {
        const int st = i / SET_STRIDE, e = i % SET_STRIDE;
        const double v = red[i];
        if (v == 0.0) continue;
        double* o = out + static_cast<size_t>(st) * CM_STRIDE;
        if (e == 0) atomicAdd(o, v);
        else if (e < 1 + W) atomicAdd(o + 1 + (e - 1), v);
        else if (e < 1 + W + W * W) {
            const int a = (e - 1 - W) / W, b = (e - 1 - W) % W;
            if ((a >> 2) > (b >> 2)) continue;            // blocks below the diagonal were never accumulated
            atomicAdd(o + 33 + a * 32 + b, v);
            if ((a >> 2) != (b >> 2)) atomicAdd(o + 33 + b * 32 + a, v);      // mirror an off-diagonal block
        }
    }
}

// Shared-memory bytes of one launch with n_sets sets of padded width W over `ncols`-column rows.
inline size_t moments_tiled_smem(int n_sets, int W, int ncols, int stages) {
    const size_t buf_floats = (static_cast<size_t>(CT_CHUNK) * ncols + 3) & ~static_cast<size_t>(3);
    return 64 /*barriers*/ + static_cast<size_t>(n_sets) * (2 + W + W * W) * sizeof(double) + static_cast<size_t>(stages) * buf_floats * 4 + static_cast<size_t>(CT_CHUNK) * n_sets * W * 4 +
           static_cast<size_t>(n_sets) * W * 8;
}

template <int NB>
inline int launch_moments_tiled(const float* data_dev, int ld, int ncols, const int* cols_dev, int n_sets, const float* shift_dev, long long row_begin, long long row_end,
                                double* out_dev, int sms, cudaStream_t s) {
    int stages = CT_MAX_STAGES;
    while (stages > 2 && moments_tiled_smem(n_sets, 4 * NB, ncols, stages) > static_cast<size_t>(CT_SMEM_LIMIT)) --stages;
    const size_t smem = moments_tiled_smem(n_sets, 4 * NB, ncols, stages);
    if (smem > CT_SMEM_LIMIT) return fail("corr_moments_tiled: %d columns x %d sets need %zu bytes of shared memory per block (limit %d)", ncols, n_sets, smem, CT_SMEM_LIMIT);
    static PerDevice dev_state;
    if (!dev_state.configured()) {
        OSTEO_CUDA(cudaFuncSetAttribute(corr_moments_tiled_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM_LIMIT));
        dev_state.set_configured();
    }
    const long long nchunks = (row_end - row_begin + CT_CHUNK - 1) / CT_CHUNK;
    const long long blocks = nchunks < sms ? nchunks : sms;
    corr_moments_tiled_kernel<NB><<<static_cast<unsigned>(blocks), CT_THREADS, smem, s>>>(data_dev, ld, ncols, cols_dev, n_sets, shift_dev, row_begin, row_end, stages, out_dev);
    OSTEO_CUDA(cudaGetLastError());
    return 0;
}

// Per-set coherence score from the (all-reduced) moment blocks: mean over i < j of the Pearson correlation, float64, the same algebra as
// DataFrame.corr() followed by the upper-triangle mean (utils/validation.py:152-157): cov = S2 - S1 S1^T / n, R = cov / (sd sd^T),
// IEEE semantics for constant columns (0 / 0 -> NaN, as pandas reports). One warp per set.
__global__ void coherence_finish_kernel(const double* __restrict__ mom, const int* __restrict__ cols, int n_sets, double* __restrict__ scores) {
    const int lane = threadIdx.x & 31;
    const int set = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (set >= n_sets) return;
    const double* m = mom + static_cast<size_t>(set) * CM_STRIDE;
    const bool has = cols[set * 32 + lane] >= 0;
    const int k = __popc(__ballot_sync(0xffffffffu, has));
    const double n = m[0];
    const double s1 = has ? m[1 + lane] : 0.0;
    const double var = has ? m[33 + lane * 32 + lane] - s1 * s1 / n : 1.0;
    const double sd = sqrt(var);
    double sum = 0.0;
    for (int j = 0; j < k; ++j) {
        const double s1j = __shfl_sync(0xffffffffu, s1, j), sdj = __shfl_sync(0xffffffffu, sd, j);
        if (has && j > lane) sum += (m[33 + lane * 32 + j] - s1 * s1j / n) / (sd * sdj);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) scores[set] = sum / (0.5 * k * (k - 1));
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
