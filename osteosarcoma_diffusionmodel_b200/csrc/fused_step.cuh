// Fused tail of one reverse step (bf16 throughput mode):
//
//   eps      = h_final . W_out^T + b_out                         (models/diffusion.py:254)
//   x_{t-1}  = c_x[t] x_t - c_eps[t] eps + sigma[t] z            (models/diffusion.py:400-423, collapsed)
//   h0_next  = bf16(x_{t-1}) . W_in^T + b_in + time_proj[t-1] + cond_proj     (models/diffusion.py:229-232 of the NEXT step)
//
// in ONE persistent kernel, so the state is read once and written once per step in fp32 (41 136 B per
// patient-step, the algorithmic minimum) and no bf16 shadow of x ever touches HBM: each 128 x 64 tile of
// x_{t-1} is rounded to bf16 into shared memory in the UMMA K-major layout and immediately contracted
// against W_in[:, 64 columns] into a per-row-block TMEM accumulator.
//
// Work unit = one 128-row block, all DP / 64 column tiles. Roles (640 threads, 1 CTA / SM):
//   (logical warps; physically the four role warps are 16..19 and the epilogue warps 0..15, see the kernel)
//   warp 0      TMA producer: resident A (h_final, 128 x h0 bf16) once per unit, W_out tile [64 x h0] per column tile
//   warp 1      MMA issuer: eps(j) = A . W_out[j]^T  (M128 N64),  acc_in += xbf(j-2) . W_in[j-2]^T  (M128 N=h0)
//   warp 2      TMEM allocator (512 columns: acc_in 256 | 4 eps stages of 64)
//   warp 3      TMA producer for W_in tiles [h0 x 64] + L2 prefetch of the state tile two tiles ahead
//   warps 4-19  epilogue: thread <-> one row x 16 columns of the tile: Philox + Box-Muller noise, update,
//               256-bit global load / store of the fp32 state, bf16 tile into shared memory
//
// State layout ("c8"): x[m_block][DP / 8][128 rows][8 cols] fp32, so one warp-wide 256-bit access (32 rows x 32 B)
// is 1 KB contiguous and a whole 128 x 64 tile is 32 KB contiguous.
#pragma once
#include "gemm_tc.cuh"

namespace osteo {

constexpr int FT = 64;                                  // state columns per tile
constexpr int F_EPI_WARPS = 16;
constexpr int F_THREADS = 128 + 32 * F_EPI_WARPS;       // 640
constexpr int F_ARES_BYTES = 4 * A_TILE_BYTES;          // 64 KB: up to 4 k-blocks of [128 x 64] bf16
constexpr int F_WOUT_KB_BYTES = FT * BK * 2;            // 8 KB: one k-block of a W_out tile [64 x 64]
constexpr int F_WOUT_STAGE = 4 * F_WOUT_KB_BYTES;       // 32 KB
constexpr int F_WIN_STAGE = 256 * BK * 2;               // 32 KB: [256 x 64]
constexpr int F_XBF_STAGE = BM * BK * 2;                // 16 KB: [128 x 64]
constexpr int F_STAGES = 2;
constexpr int F_EPS_ACC = 4;
constexpr int F_LAG = 2;                                // the next-step contraction trails the eps GEMM by two tiles
constexpr int F_TMEM_EPS0 = 256;                        // first TMEM column of the eps stages
constexpr int F_BIAS_SLOTS = 6;                         // output_proj bias of a tile (64 floats) rides along with its W_out tile
constexpr int F_BIAS_BYTES = FT * 4;
constexpr int F_SMEM_BYTES = F_ARES_BYTES + F_STAGES * (F_WOUT_STAGE + F_WIN_STAGE + F_XBF_STAGE) + 1024 /*align*/ + 256 /*barriers*/ + F_BIAS_SLOTS * F_BIAS_BYTES;

struct FusedParams {
    CUtensorMap tma_a;        // h_final bf16 [rows, 2*h0], box 128 x 64
    CUtensorMap tma_wout;     // output_proj weight, packed bf16 [np, 2*kp], box 64 x 64
    CUtensorMap tma_win;      // input_proj weight, packed bf16 [h0, 2*DP], box h0 x 64
    int M, N;                 // valid rows (absolute) / valid state columns (D)
    int m_tile0, m_tiles;     // row blocks [m_tile0, m_tile0 + m_tiles)
    int n_tiles;              // DP / 64
    int nkb;                  // h0 / 64
    int h0;
    int* status;
    const int* step;
    const float* coef_x;
    const float* coef_eps;
    const float* coef_sigma;
    float* x;                 // fp32 state, c8 layout
    int x_c8;                 // DP / 8
    const float* bias_out;    // [>= DP], zero padded
    const float* noise;       // optional injected z, dense [M, noise_ld]
    int noise_ld;
    float* eps_out;           // optional dense eps [M, eps_ld]
    int eps_ld;
    unsigned long long seed;
    long long row_base;
    const float* bias_in;     // [h0]
    const float* time_table;  // [T, h0]
    const float* cproj;       // [rows, h0]
    __nv_bfloat16* h0_out;    // [rows, h0_ld] bf16: the next step's first activation
    int h0_ld;
    int dbg;
};

__device__ __forceinline__ void ld_global_v8(const void* ptr, uint32_t (&w)[8]) {
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(ptr)
                 : "memory");
}
// 1-D bulk copy global -> shared (size a multiple of 16 B), completion counted on `bar` like a tensor load.
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void l2_prefetch_bulk(const void* ptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
}

// 16 scaled normals sg * z for columns [4*col4_0, 4*col4_0 + 16) of one row. Same Philox counters and uniforms as
// philox_normal_row (philox.cuh); sigma is folded into the Box-Muller radius: sg * sqrt(-2 ln u1) = sqrt(k2 * lg2 u1),
// k2 = -2 ln2 sg^2.
__device__ __forceinline__ void philox_scaled_normal16(uint64_t seed, uint64_t row, uint32_t col4_0, uint32_t stream, uint32_t step, float k2, float (&z)[16]) {
    uint4 c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        c[i] = make_uint4(col4_0 + i, static_cast<uint32_t>(row), static_cast<uint32_t>(row >> 32), (stream << 16) | (step & 0xFFFFu));
    philox4x32_10_batch<4>(c, make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t w[4] = {c[i].x, c[i].y, c[i].z, c[i].w};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            // u1 = 2 - f in (0, 1] is never subnormal (>= 2^-23): the .ftz forms drop the denormal pre-scaling __log2f emits.
            // The angle is 2 pi f with f in [1, 2): one full turn more than 2 pi (f - 1), same sine and cosine, one FMUL instead of an FFMA.
            const float u1 = 2.0f - unit_1_2(w[2 * h]);
            const float th = unit_1_2(w[2 * h + 1]) * 6.283185307179586f;
            float l2, r, s, co;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(k2 * l2));
            asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
            asm("cos.approx.ftz.f32 %0, %1;" : "=f"(co) : "f"(th));
            z[4 * i + 2 * h] = r * co;
            z[4 * i + 2 * h + 1] = r * s;
        }
    }
}

__global__ void __launch_bounds__(F_THREADS, 1) ddpm_fused_kernel(const __grid_constant__ FusedParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* s_a = smem;
    uint8_t* s_wout = s_a + F_ARES_BYTES;
    uint8_t* s_win = s_wout + F_STAGES * F_WOUT_STAGE;
    uint8_t* s_xbf = s_win + F_STAGES * F_WIN_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_xbf + F_STAGES * F_XBF_STAGE);
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + 1;
    uint64_t* wout_full = bars + 2;
    uint64_t* wout_empty = bars + 4;
    uint64_t* win_full = bars + 6;
    uint64_t* win_empty = bars + 8;
    uint64_t* xbf_full = bars + 10;
    uint64_t* xbf_empty = bars + 12;
    uint64_t* tfull = bars + 14;
    uint64_t* tempty = bars + 18;
    uint64_t* accin_full = bars + 22;
    uint64_t* accin_empty = bars + 23;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
    uint8_t* s_bias = reinterpret_cast<uint8_t*>(bars) + 256;      // F_BIAS_SLOTS x 256 B

    // Role warps take the HIGHEST warp ids: the SMSP arbiter favours high warp ids, and a starved MMA issuer / TMA producer stalls
    // all sixteen epilogue warps. Epilogue warps are 0..15 (TMEM lane quadrant = warp % 4).
    const int warp_phys = threadIdx.x >> 5;
    const int warp = warp_phys < F_EPI_WARPS ? warp_phys + 4 : warp_phys - F_EPI_WARPS;      // logical: 0..3 roles, 4..19 epilogue
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tma_a);
        tma_prefetch_desc(&p.tma_wout);
        tma_prefetch_desc(&p.tma_win);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int i = 0; i < F_STAGES; ++i) {
            mbar_init(&wout_full[i], 1);
            mbar_init(&wout_empty[i], 1);
            mbar_init(&win_full[i], 1);
            mbar_init(&win_empty[i], 1);
            mbar_init(&xbf_full[i], F_EPI_WARPS);
            mbar_init(&xbf_empty[i], 1);
        }
        for (int i = 0; i < F_EPS_ACC; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], F_EPI_WARPS);
        }
        mbar_init(accin_full, 1);
        mbar_init(accin_empty, F_EPI_WARPS);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int nt = p.n_tiles;

    if (warp == 0) {
        // ------------------------------------------------ producer: resident A + W_out tiles
        if (lane == 0) {
            int itw = 0, k = 0, bslot = 0;
            bool ok = true;
            for (int u = blockIdx.x; ok && u < p.m_tiles; u += gridDim.x, ++k) {
                const int m_blk = p.m_tile0 + u;
                if (!mbar_wait_relaxed(a_empty, (static_cast<uint32_t>(k) & 1u) ^ 1u)) { ok = false; break; }
                mbar_arrive_expect_tx(a_full, p.nkb * A_TILE_BYTES);
                for (int kb = 0; kb < p.nkb; ++kb) tma_load_2d(&p.tma_a, s_a + kb * A_TILE_BYTES, a_full, kb * BK, m_blk * BM);
                for (int j = 0; j < nt; ++j, ++itw) {
                    const int s = itw & 1;
                    // stage s was last read by eps(itw - 2): its completion is that tile's tfull barrier (no separate "empty" commit)
                    if (itw >= 2 && !mbar_wait_relaxed(&tfull[(itw - 2) & (F_EPS_ACC - 1)], (static_cast<uint32_t>(itw - 2) >> 2) & 1u)) { ok = false; break; }
                    // the bias slot of tile itw - 6 is free: eps(itw - 2) has completed, so the epilogue of tile itw - 6 has arrived on
                    // tempty, which it does only after reading its bias
                    mbar_arrive_expect_tx(&wout_full[s], p.nkb * F_WOUT_KB_BYTES + F_BIAS_BYTES);
                    for (int kb = 0; kb < p.nkb; ++kb)
                        tma_load_2d(&p.tma_wout, s_wout + s * F_WOUT_STAGE + kb * F_WOUT_KB_BYTES, &wout_full[s], kb * BK, j * FT);
                    bulk_load_1d(s_bias + bslot * F_BIAS_BYTES, p.bias_out + j * FT, F_BIAS_BYTES, &wout_full[s]);
                    bslot = bslot + 1 == F_BIAS_SLOTS ? 0 : bslot + 1;
                }
            }
            if (!ok) atomicExch(p.status, ERR_PRODUCER_TIMEOUT);
        }
    } else if (warp == 3) {
        // ------------------------------------------------ producer: W_in tiles + L2 prefetch of the state
        if (lane == 0) {
            int iti = 0;
            bool ok = true;
            const size_t tile_floats = static_cast<size_t>(FT / 8) * BM * 8;      // 32 KB per tile
            for (int u = blockIdx.x; ok && u < p.m_tiles; u += gridDim.x) {
                const int m_blk = p.m_tile0 + u;
                const float* xblk = p.x + static_cast<size_t>(m_blk) * p.x_c8 * BM * 8;
                if (!(p.dbg & 1)) {
                    l2_prefetch_bulk(xblk, static_cast<uint32_t>(tile_floats * 4));
                    if (nt > 1) l2_prefetch_bulk(xblk + tile_floats, static_cast<uint32_t>(tile_floats * 4));
                }
                for (int j = 0; j < nt; ++j, ++iti) {
                    const int s = iti & 1;
                    // stage s was last read by the next-step MMA of tile iti - 2, which also releases the bf16 tile buffer: one barrier
                    if (!mbar_wait_relaxed(&xbf_empty[s], ((static_cast<uint32_t>(iti) >> 1) & 1u) ^ 1u)) { ok = false; break; }
                    if (j + 2 < nt && !(p.dbg & 1)) l2_prefetch_bulk(xblk + static_cast<size_t>(j + 2) * tile_floats, static_cast<uint32_t>(tile_floats * 4));
                    mbar_arrive_expect_tx(&win_full[s], p.h0 * BK * 2);
                    tma_load_2d(&p.tma_win, s_win + s * F_WIN_STAGE, &win_full[s], j * FT, 0);
                }
            }
            if (!ok) atomicExch(p.status, ERR_PRODUCER_TIMEOUT);
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer
        // The WHOLE warp runs this loop with warp-uniform control flow and only the tcgen05 instructions are predicated on one
        // elected lane: inside an `if (lane == 0)` region the compiler wraps every UTCHMMA / UTCBAR in an ELECT + branch loop and
        // rebuilds the descriptors through the vector registers, and the single issuing thread (~10 dependent instructions per MMA)
        // cannot keep up with 32-cycle MMAs.
        const uint32_t idesc_eps = make_idesc_bf16(BM, FT, 0, 0);
        const uint32_t idesc_in = make_idesc_bf16(BM, p.h0, 0, 0);
        const bool leader = elect_one();
        const uint64_t adesc0 = make_kmajor_sw128_desc(smem_u32(s_a));
        const uint64_t wout_desc0 = make_kmajor_sw128_desc(smem_u32(s_wout));
        const uint64_t win_desc0 = make_kmajor_sw128_desc(smem_u32(s_win));
        const uint64_t xbf_desc0 = make_kmajor_sw128_desc(smem_u32(s_xbf));
        int it_eps = 0, it_in = 0, k = 0;
        bool ok = true;
        for (int u = blockIdx.x; ok && u < p.m_tiles; u += gridDim.x, ++k) {
            if (!mbar_wait_relaxed(a_full, static_cast<uint32_t>(k) & 1u)) { ok = false; break; }
            tc_fence_after_sync();
            for (int j = 0; j < nt + F_LAG && ok; ++j) {
                if (j < nt) {
                    const int s = it_eps & 1, acc = it_eps & (F_EPS_ACC - 1);
                    if (!mbar_wait_relaxed(&wout_full[s], (static_cast<uint32_t>(it_eps) >> 1) & 1u)) { ok = false; break; }
                    if (!mbar_wait_relaxed(&tempty[acc], ((static_cast<uint32_t>(it_eps) >> 2) & 1u) ^ 1u)) { ok = false; break; }
                    tc_fence_after_sync();
                    if (leader && !(p.dbg & 64)) {
                        const uint32_t d = tmem_base + F_TMEM_EPS0 + static_cast<uint32_t>(acc * FT);
                        const uint64_t bdesc_s = wout_desc0 + static_cast<uint64_t>((s * F_WOUT_STAGE) >> 4);
                        for (int kb = 0; kb < p.nkb; ++kb) {
                            const uint64_t adesc = adesc0 + static_cast<uint64_t>((kb * A_TILE_BYTES) >> 4);
                            const uint64_t bdesc = bdesc_s + static_cast<uint64_t>((kb * F_WOUT_KB_BYTES) >> 4);
#pragma unroll
                            for (int kk = 0; kk < BK / 16; ++kk) umma_bf16(d, adesc + 2u * kk, bdesc + 2u * kk, idesc_eps, (kb | kk) != 0 ? 1u : 0u);
                        }
                    }
                    if (leader) umma_commit(&tfull[acc]);
                    __syncwarp();
                    ++it_eps;
                }
                if (j >= F_LAG) {
                    const int jj = j - F_LAG;
                    const int s = it_in & 1;
                    if (jj == 0 && !mbar_wait_relaxed(accin_empty, (static_cast<uint32_t>(k) & 1u) ^ 1u)) { ok = false; break; }
                    if (!mbar_wait_relaxed(&xbf_full[s], (static_cast<uint32_t>(it_in) >> 1) & 1u)) { ok = false; break; }
                    if (!mbar_wait_relaxed(&win_full[s], (static_cast<uint32_t>(it_in) >> 1) & 1u)) { ok = false; break; }
                    tc_fence_after_sync();
                    if (leader) {
                        const uint64_t adesc = xbf_desc0 + static_cast<uint64_t>((s * F_XBF_STAGE) >> 4);
                        const uint64_t bdesc = win_desc0 + static_cast<uint64_t>((s * F_WIN_STAGE) >> 4);
                        if (!(p.dbg & 128)) {
#pragma unroll
                            for (int kk = 0; kk < BK / 16; ++kk) umma_bf16(tmem_base, adesc + 2u * kk, bdesc + 2u * kk, idesc_in, (jj | kk) != 0 ? 1u : 0u);
                        }
                        umma_commit(&xbf_empty[s]);
                    }
                    __syncwarp();
                    ++it_in;
                }
            }
            if (ok && leader) {
                umma_commit(accin_full);
                umma_commit(a_empty);
            }
            __syncwarp();
        }
        if (!ok && leader) atomicExch(p.status, ERR_MMA_TIMEOUT);
    } else if (warp >= 4) {
        // ------------------------------------------------ epilogue
        const int q = warp_phys & 3;            // TMEM lane quadrant (hardware rule: physical warp id % 4)
        const int part = warp_phys >> 2;        // 16-column quarter of the 64-column tile
        const int r_tile = q * 32 + lane;
        const int t = *p.step;
        const float cx = __ldg(p.coef_x + t), nce = -__ldg(p.coef_eps + t);
        const float sg = (p.dbg & 8) ? 0.0f : __ldg(p.coef_sigma + t);
        const float k2 = -1.3862943611198906f * sg * sg;
        const int t_next = t > 0 ? t - 1 : 0;
        int it = 0, k = 0, bslot = 0;
        bool ok = true;
        for (int u = blockIdx.x; ok && u < p.m_tiles; u += gridDim.x, ++k) {
            const int m_blk = p.m_tile0 + u;
            const int row = m_blk * BM + r_tile;
            const bool live = row < p.M;
            float* xrow = p.x + (static_cast<size_t>(m_blk) * p.x_c8 * BM + r_tile) * 8;      // + c8 * (BM * 8)
            for (int j = 0; j < nt && ok; ++j, ++it) {
                const int c0 = j * FT + part * 16;
                float* xp = xrow + static_cast<size_t>(c0 >> 3) * (BM * 8);
                const int nvalid = p.N - c0;            // >= 16: all columns valid; <= 0: all padding
                uint32_t xw[16];
                float z[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) { xw[i] = 0u; z[i] = 0.0f; }
                if (live && nvalid > 0 && !(p.dbg & 16)) {
                    ld_global_v8(xp, *reinterpret_cast<uint32_t(*)[8]>(&xw[0]));
                    ld_global_v8(xp + BM * 8, *reinterpret_cast<uint32_t(*)[8]>(&xw[8]));
                }
                if (live && nvalid > 0 && sg != 0.0f) {
                    if (p.noise) {
                        const float* nz = p.noise + static_cast<size_t>(row) * p.noise_ld + c0;
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (i < nvalid) z[i] = sg * nz[i];
                    } else {
                        philox_scaled_normal16(p.seed, static_cast<uint64_t>(p.row_base + row), static_cast<uint32_t>(c0 >> 2), STREAM_REVERSE, static_cast<uint32_t>(t), k2, z);
                    }
                }
                const int acc = it & (F_EPS_ACC - 1);
                if (!mbar_wait(&tfull[acc], (static_cast<uint32_t>(it) >> 2) & 1u)) { ok = false; break; }
                tc_fence_after_sync();
                uint32_t vr[16];
                tmem_ld_16_nowait(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(F_TMEM_EPS0 + acc * FT + part * 16), vr);
                tmem_ld_wait();
                // bias of this tile: staged in shared memory by the W_out producer (an L1 hit is not available: the whole L1 is carved out
                // as shared memory, and an L2 round trip per tile on the critical path cost ~1000 cycles)
                const float* sb = reinterpret_cast<const float*>(s_bias + bslot * F_BIAS_BYTES) + part * 16;
                bslot = bslot + 1 == F_BIAS_SLOTS ? 0 : bslot + 1;
                float xn[16];
                if (nvalid >= 16) {
                    const float4* b4 = reinterpret_cast<const float4*>(sb);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 b = b4[i];
                        const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float ev = __uint_as_float(vr[4 * i + e]) + bb[e];
                            xn[4 * i + e] = fmaf(nce, ev, fmaf(cx, __uint_as_float(xw[4 * i + e]), z[4 * i + e]));
                        }
                    }
                    if (p.eps_out && live) {
                        float* eo = p.eps_out + static_cast<size_t>(row) * p.eps_ld + c0;
#pragma unroll
                        for (int i = 0; i < 16; ++i) eo[i] = __uint_as_float(vr[i]) + sb[i];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float v = 0.0f;
                        if (i < nvalid) {
                            const float ev = __uint_as_float(vr[i]) + sb[i];
                            if (p.eps_out && live) p.eps_out[static_cast<size_t>(row) * p.eps_ld + c0 + i] = ev;
                            v = fmaf(nce, ev, fmaf(cx, __uint_as_float(xw[i]), z[i]));
                        }
                        xn[i] = v;
                    }
                }
                if (!live) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) xn[i] = 0.0f;
                }
                // accumulator and bias are in registers: hand the TMEM stage (and, transitively, the bias slot) back
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                uint32_t xo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) xo[i] = __float_as_uint(xn[i]);
                if (live && !(p.dbg & 4)) {
                    st_global_v8(xp, *reinterpret_cast<uint32_t(*)[8]>(&xo[0]));
                    st_global_v8(xp + BM * 8, *reinterpret_cast<uint32_t(*)[8]>(&xo[8]));
                }
                // bf16 tile for the next step's input_proj: row r_tile, 16-byte chunks 2*part and 2*part+1 of the 128-byte swizzled row
                const int s = it & 1;
                if (!mbar_wait(&xbf_empty[s], ((static_cast<uint32_t>(it) >> 1) & 1u) ^ 1u)) { ok = false; break; }
                {
                    uint8_t* rowp = s_xbf + s * F_XBF_STAGE + (r_tile >> 3) * 1024 + (r_tile & 7) * 128;
                    const int sw = r_tile & 7;
                    uint4 w0, w1;
                    w0.x = pack_bf16x2(xn[0], xn[1]);   w0.y = pack_bf16x2(xn[2], xn[3]);   w0.z = pack_bf16x2(xn[4], xn[5]);   w0.w = pack_bf16x2(xn[6], xn[7]);
                    w1.x = pack_bf16x2(xn[8], xn[9]);   w1.y = pack_bf16x2(xn[10], xn[11]); w1.z = pack_bf16x2(xn[12], xn[13]); w1.w = pack_bf16x2(xn[14], xn[15]);
                    *reinterpret_cast<uint4*>(rowp + (((2 * part) ^ sw) << 4)) = w0;
                    *reinterpret_cast<uint4*>(rowp + (((2 * part + 1) ^ sw) << 4)) = w1;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&xbf_full[s]);
            }
            if (!ok) break;
            // ---- unit end: h0 of the next step = acc_in + b_in + time_proj[t-1] + cond_proj  -> bf16
            if (!mbar_wait(accin_full, static_cast<uint32_t>(k) & 1u)) { ok = false; break; }
            tc_fence_after_sync();
            const int cpw = p.h0 >> 2;              // columns per warp quarter (32 or 64)
            for (int cc = 0; cc < cpw; cc += 16) {
                const int c = part * cpw + cc;
                uint32_t vr[16];
                tmem_ld_16_nowait(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c), vr);
                tmem_ld_wait();
                if (live) {
                    const float4* b4 = reinterpret_cast<const float4*>(p.bias_in + c);
                    const float4* t4 = reinterpret_cast<const float4*>(p.time_table + static_cast<size_t>(t_next) * p.h0 + c);
                    const float4* m4 = reinterpret_cast<const float4*>(p.cproj + static_cast<size_t>(row) * p.h0 + c);
                    uint32_t o[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 b = __ldg(b4 + i), tt = __ldg(t4 + i), mm = __ldg(m4 + i);
                        const float v0 = ((__uint_as_float(vr[4 * i + 0]) + b.x) + tt.x) + mm.x;
                        const float v1 = ((__uint_as_float(vr[4 * i + 1]) + b.y) + tt.y) + mm.y;
                        const float v2 = ((__uint_as_float(vr[4 * i + 2]) + b.z) + tt.z) + mm.z;
                        const float v3 = ((__uint_as_float(vr[4 * i + 3]) + b.w) + tt.w) + mm.w;
                        o[2 * i] = pack_bf16x2(v0, v1);
                        o[2 * i + 1] = pack_bf16x2(v2, v3);
                    }
                    st_global_v8(p.h0_out + static_cast<size_t>(row) * p.h0_ld + c, o);
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(accin_empty);
        }
        if (!ok && lane == 0) atomicExch(p.status, ERR_EPI_TIMEOUT);
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace osteo
