// RBF Gram reduction for the MMD check (utils/validation.py:284-298):  sum over a tile of exp(-gamma ||a_m - b_n||^2),
// ||a - b||^2 = |a|^2 + |b|^2 - 2 a.b clamped at 0, with the dot products on tcgen05 and 128 x 256 output tiles.
//
// Why its own kernel: with SS-mode MMAs every operand byte is written into shared memory once (TMA) and read from it once (MMA). A
// 128 x 128 tile moves 16 KB + 16 KB per 64-wide k-block through a 128 B/clk shared memory for 256 cycles of tensor work, so the generic
// kernel (gemm_tc.cuh, EPI_RBF) tops out at ~50 % of the tensor peak (measured 49-55 %). A 128 x 256 tile moves 48 KB + 48 KB per 512
// cycles: 2/3 of the peak is reachable without pairing CTAs. The exp2 epilogue is 5 % of a tile's MMA time at K = 5184.
//
// 384 threads, 1 CTA / SM: warps 0-7 epilogue (quadrant = warp % 4, 128 columns each, four passes of 32), 8 TMA producer, 9 MMA issuer
// (whole warp, one elected lane), 10 TMEM allocator (2 accumulator stages x 256 columns). Static round-robin over (row block, column
// block) tiles; in the symmetric case (X against itself, this call owns all rows) only 128-column halves on or above the diagonal are
// computed, halves above it weighted 2.
#pragma once
#include "gemm_tc.cuh"

namespace osteo {

constexpr int RB_BN = 256;
constexpr int RB_STAGES = 4;
constexpr int RB_B_BYTES = RB_BN * BK * 2;                       // 32 KB
constexpr int RB_STAGE_BYTES = A_TILE_BYTES + RB_B_BYTES;        // 48 KB
constexpr int RB_ACC = 2;
constexpr int RB_GROUP_M = 16;                                   // row blocks per tile group (L2 reuse of the B operand, see decode)
constexpr int RB_EPI_WARPS = 8;
constexpr int RB_THREADS = 32 * (RB_EPI_WARPS + 4);
constexpr int RB_SMEM_BYTES = RB_STAGES * RB_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

struct RbfParams {
    CUtensorMap tma_a;        // rows of A: bf16 [rows_a, 2*kp] = [hi | lo], box 128 x 64
    CUtensorMap tma_b;        // rows of B: same layout, box 256 x 64
    int M, N;                 // valid rows of A (absolute) / valid rows of B
    int m_tile0, m_tiles;     // 128-row blocks of A this call covers: m_tile0 + i * m_stride, i < m_tiles
    int m_stride;             // 1 = a contiguous range; world_size = block-cyclic sharding over ranks (balanced for the symmetric half-Gram)
    int n_tiles;              // 256-row blocks of B
    int nseg;
    KSeg seg[3];              // (a_col, b_col, nkb): hi.hi, and in split mode hi.lo, lo.hi
    const float* norm_a;      // [rows_a] squared norms of the representation the MMA sees
    const float* norm_b;
    float gamma;
    int symmetric;
    double* acc;
    int* status;
};

__global__ void __launch_bounds__(RB_THREADS, 1) rbf_gram_kernel(const __grid_constant__ RbfParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic ON the __shared__ array: an integer round trip would turn every later access into a
    // generic LD / ST (address-space lookup in the LSU, several times slower than LDS / STS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + RB_STAGES * RB_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + RB_STAGES;
    uint64_t* tfull_bar = empty_bar + RB_STAGES;
    uint64_t* tempty_bar = tfull_bar + RB_ACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + RB_ACC);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int W_PROD = RB_EPI_WARPS, W_MMA = RB_EPI_WARPS + 1, W_ALLOC = RB_EPI_WARPS + 2;
    const int num_tiles = p.m_tiles * p.n_tiles;
    int total_kb = 0;
    for (int s = 0; s < p.nseg; ++s) total_kb += p.seg[s].nkb;

    if (warp == W_PROD && lane == 0) {
        tma_prefetch_desc(&p.tma_a);
        tma_prefetch_desc(&p.tma_b);
    }
    if (warp == W_MMA && lane == 0) {
        for (int i = 0; i < RB_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < RB_ACC; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], RB_EPI_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == W_ALLOC) tmem_alloc(tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // tile -> (row block, 256-column block); a symmetric tile entirely below the diagonal is skipped by all three roles
    // Tiles are walked in groups of RB_GROUP_M row blocks: consecutive tile indices (= the CTAs of one wave) go DOWN the group's row
    // blocks for one 256-column block, then to the next column block. A wave of 148 CTAs then shares 16 A row blocks and ~9 B column
    // blocks (45 MB of operands for 148 tiles at K = 5184) instead of one A row block and 148 different B blocks (390 MB): at 1 M x 1 M
    // the B operand (10 GB) is far beyond the L2 and the row-major walk was HBM-bound at 49 % of the tensor peak
    // (profiles/r2_config4_n8.json).
    auto decode = [&](int tile, int& m_blk, int& n_blk) -> bool {
        const int per_group = RB_GROUP_M * p.n_tiles;
        const int group = tile / per_group, in_group = tile - group * per_group;
        const int first = group * RB_GROUP_M;
        const int rows_here = p.m_tiles - first < RB_GROUP_M ? p.m_tiles - first : RB_GROUP_M;
        const int mi = first + in_group % rows_here;
        n_blk = in_group / rows_here;
        m_blk = p.m_tile0 + mi * p.m_stride;
        return !(p.symmetric && 2 * n_blk + 1 < m_blk);
    };

    if (warp == W_PROD) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            for (int tile = blockIdx.x; ok && tile < num_tiles; tile += gridDim.x) {
                int m_blk, n_blk;
                if (!decode(tile, m_blk, n_blk)) continue;
                for (int s = 0; s < p.nseg && ok; ++s) {
                    const KSeg sg = p.seg[s];
                    for (int kb = 0; kb < sg.nkb; ++kb) {
                        if (!mbar_wait_relaxed(&empty_bar[stage], phase ^ 1u)) { ok = false; break; }
                        mbar_arrive_expect_tx(&full_bar[stage], RB_STAGE_BYTES);
                        uint8_t* st = smem + stage * RB_STAGE_BYTES;
                        tma_load_2d(&p.tma_a, st, &full_bar[stage], sg.a_col + kb * BK, m_blk * BM);
                        tma_load_2d(&p.tma_b, st + A_TILE_BYTES, &full_bar[stage], sg.b_col + kb * BK, n_blk * RB_BN);
                        if (++stage == RB_STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
            if (!ok) atomicExch(p.status, ERR_PRODUCER_TIMEOUT);
        }
    } else if (warp == W_MMA) {
        constexpr uint32_t idesc = make_idesc_bf16(BM, RB_BN, 0, 0);
        const bool leader = elect_one();
        const uint64_t desc0 = make_kmajor_sw128_desc(smem_u32(smem));
        int stage = 0, it = 0;
        uint32_t phase = 0;
        bool ok = true;
        for (int tile = blockIdx.x; ok && tile < num_tiles; tile += gridDim.x) {
            int m_blk, n_blk;
            if (!decode(tile, m_blk, n_blk)) continue;
            const int acc = it & (RB_ACC - 1);
            if (!mbar_wait_relaxed(&tempty_bar[acc], ((static_cast<uint32_t>(it) >> 1) & 1u) ^ 1u)) { ok = false; break; }
            ++it;
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * RB_BN);
            for (int idx = 0; idx < total_kb; ++idx) {
                if (!mbar_wait_relaxed(&full_bar[stage], phase)) { ok = false; break; }
                tc_fence_after_sync();
                if (leader) {
                    const uint64_t adesc = desc0 + static_cast<uint64_t>((stage * RB_STAGE_BYTES) >> 4);
                    const uint64_t bdesc = adesc + static_cast<uint64_t>(A_TILE_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (idx | k) != 0 ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == RB_STAGES) { stage = 0; phase ^= 1u; }
            }
            if (ok && leader) umma_commit(&tfull_bar[acc]);
            __syncwarp();
        }
        if (!ok && leader) atomicExch(p.status, ERR_MMA_TIMEOUT);
    } else if (warp < RB_EPI_WARPS) {
        const int q = warp & 3;                 // TMEM lane quadrant
        const int half = warp >> 2;             // 128-column half of the tile
        const float ng = -p.gamma * 1.4426950408889634f;      // exp(x) = exp2(x log2 e)
        double thread_acc = 0.0;
        int it = 0;
        bool ok = true;
        for (int tile = blockIdx.x; ok && tile < num_tiles; tile += gridDim.x) {
            int m_blk, n_blk;
            if (!decode(tile, m_blk, n_blk)) continue;
            const int acc = it & (RB_ACC - 1);
            if (!mbar_wait(&tfull_bar[acc], (static_cast<uint32_t>(it) >> 1) & 1u)) { ok = false; break; }
            ++it;
            tc_fence_after_sync();
            const int row = m_blk * BM + q * 32 + lane;
            const int hb = 2 * n_blk + half;                    // 128-column block index of this warp's half
            const float w = !p.symmetric ? 1.0f : (hb > m_blk ? 2.0f : (hb == m_blk ? 1.0f : 0.0f));
            const float na = row < p.M ? p.norm_a[row] : 0.0f;
            float local = 0.0f;
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
                const int c0 = n_blk * RB_BN + half * 128 + ch * 32;
                if (w == 0.0f || c0 >= p.N) continue;           // warp-uniform
                float v[32];
                tmem_ld_32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * RB_BN + half * 128 + ch * 32), v);
                if (row < p.M) {
                    if (c0 + 32 <= p.N) {
                        const float4* nb4 = reinterpret_cast<const float4*>(p.norm_b + c0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 nb = __ldg(nb4 + j);
                            local += exp2f(ng * fmaxf(na + nb.x - 2.0f * v[4 * j + 0], 0.0f));
                            local += exp2f(ng * fmaxf(na + nb.y - 2.0f * v[4 * j + 1], 0.0f));
                            local += exp2f(ng * fmaxf(na + nb.z - 2.0f * v[4 * j + 2], 0.0f));
                            local += exp2f(ng * fmaxf(na + nb.w - 2.0f * v[4 * j + 3], 0.0f));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (c0 + j < p.N) local += exp2f(ng * fmaxf(na + __ldg(p.norm_b + c0 + j) - 2.0f * v[j], 0.0f));
                    }
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            thread_acc += static_cast<double>(local * w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) thread_acc += __shfl_xor_sync(0xffffffffu, thread_acc, o);
        if (lane == 0 && thread_acc != 0.0) atomicAdd(p.acc, thread_acc);
        if (!ok && lane == 0) atomicExch(p.status, ERR_EPI_TIMEOUT);
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == W_ALLOC) tmem_dealloc(tmem_base, 512);
}

inline int launch_rbf_gram(const RbfParams& p, int num_sms, cudaStream_t stream) {
    static PerDevice dev_state;
    if (!dev_state.configured()) {
        OSTEO_CUDA(cudaFuncSetAttribute(rbf_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM_BYTES));
        dev_state.set_configured();
    }
    const int tiles = p.m_tiles * p.n_tiles;
    if (tiles <= 0) return 0;
    rbf_gram_kernel<<<tiles < num_sms ? tiles : num_sms, RB_THREADS, RB_SMEM_BYTES, stream>>>(p);
    OSTEO_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace osteo
