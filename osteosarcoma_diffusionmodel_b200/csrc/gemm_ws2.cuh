// CTA-PAIR (tcgen05 cta_group::2) variant of the weight-stationary Linear + GroupNorm + SiLU GEMM (gemm_ws.cuh) for the bf16 throughput mode.
//
// An SS-mode M128 N128 K16 MMA reads 4 KB of A and 4 KB of B from shared memory per 64 tensor cycles = 128 B / clk, exactly what one SM's
// shared memory delivers, so the TMA writes of the A ring and the epilogue's shared-memory traffic come straight out of the MMA rate
// (ncu: tensor pipe 23-38 % for the block GEMMs). Here two CTAs of a cluster (one TPC) form ONE MMA of M = 256, N = 256:
// each CTA holds its own 128 rows of A and HALF of the 256-column W slice (128 W rows, <= 128 KB resident), the tensor cores of both SMs
// read both halves, and per SM the operand traffic drops to 4 KB (A) + 4 KB (its B half) per 128 tensor cycles = 64 B / clk.
//
//   pair p of the grid  <->  256-column slice n_pair of the layer, row-pair blocks mp = first, first + step, ...
//   CTA rank r of the pair: rows [(m_tile0 + 2 mp + r) * 128, +128), W rows [n_pair * 256 + r * 128, +128)
//   TMEM (cta_group::2 allocation, 512 columns in each CTA): 2 accumulator stages of 256 columns; each CTA's TMEM holds ITS 128 rows
//
// Barriers (same offsets in both CTAs):
//   full[s]    leader only: the A tiles of BOTH CTAs for stage s have landed (every TMA load signals the leader's barrier, cta_group::2 form)
//   empty[s]   both CTAs:   the MMAs that read stage s have retired (multicast tcgen05.commit from the leader)
//   tfull[a]   both CTAs:   accumulator stage a is complete (multicast commit)
//   tempty[a]  leader only: the epilogue warps of BOTH CTAs have read stage a (the peer arrives remotely)
//   wres       leader only: both W halves are resident
// Only the leader (rank 0) issues MMAs. Epilogue = the 32-column GroupNorm epilogue of gemm_tc.cuh, run over the two 128-column halves of
// the 256-column accumulator one after the other by the same 16 warps.
#pragma once
#include "gemm_ws.cuh"

namespace osteo {

constexpr int WS2_ACC = 2;                    // accumulator stages of 256 columns
constexpr int WS2_BN = 256;                   // columns of the pair's tile

__device__ __forceinline__ uint32_t leader_addr(const void* local_smem) { return smem_u32(local_smem) & 0xFEFFFFFFu; }      // same offset in CTA rank 0 of the pair

// 2-D tiled load into THIS CTA's shared memory whose complete_tx goes to the mbarrier at the same offset in the pair's leader CTA.
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, void* smem_dst, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_addr(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// tcgen05.commit of the pair's MMAs arriving on the mbarrier at this offset in BOTH CTAs.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}
// arrive on the mbarrier at this offset in the pair's leader CTA (from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_addr(bar)) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

template <int GW>
__global__ void __launch_bounds__(WS_THREADS, 1) gemm_ws2_gn_silu_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    int total_kb = 0;
    for (int s = 0; s < p.nseg; ++s) total_kb += p.seg[s].nkb;
    // bf16 inference: the lean epilogue of gemm_ws.cuh (ws_gn_silu_lean) + per-warp TMA stores staged in the last two 16 KB slots. The pair
    // streams HALF the A bytes per output column of the single-CTA kernel, so K = 512 can afford the two slots too (3 ring stages left).
    const bool lean = p.out_lo_off == 0 && !p.xhat_bf && !p.rstd_out && p.drop_p == 0.0f && p.out_bf && !(p.dbg & 256);
    const int out_slots = (lean && p.out_tma) ? 2 : 0;
    const int stages = WS_RING_PLUS_RES - total_kb - out_slots;     // 5 (K = 512) .. 9 (K = 256) without the staging
    uint8_t* s_out = smem + (WS_RING_PLUS_RES - 2) * A_TILE_BYTES;  // [16 warps][32 rows][64 bytes, 64-byte swizzle] when out_slots
    uint8_t* s_w = smem;                                // resident W half: 128 rows, k-block major
    uint8_t* s_a = smem + total_kb * B_TILE_BYTES;      // A ring (this CTA's 128 rows)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WS_RING_PLUS_RES * A_TILE_BYTES);
    uint64_t* empty_bar = full_bar + WS_MAX_STAGES;
    uint64_t* tfull_bar = empty_bar + WS_MAX_STAGES;
    uint64_t* tempty_bar = tfull_bar + NUM_ACC;
    uint64_t* wres_bar = tempty_bar + NUM_ACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wres_bar + 1);
    float* gn_par = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 512);
    float* gn_xch = gn_par + 3 * GN_PAR_MAX;

    constexpr int NUM_EPI_WARPS = 16;
    const int warp_phys = threadIdx.x >> 5;
    const int warp = warp_phys < NUM_EPI_WARPS ? warp_phys + 4 : warp_phys - NUM_EPI_WARPS;      // logical: 0 producer, 1 MMA, 2 alloc, 4.. epilogue
    const int lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const bool leader_cta = crank == 0;

    // this pair's column slice and its row-pair blocks mp = mp_first, mp_first + mp_step, ...
    const int n_pairs = p.n_tiles / 2;                  // 256-column slices
    const int pair_id = static_cast<int>(blockIdx.x) >> 1;
    const int n_pair = pair_id % n_pairs;
    const int mp_first = pair_id / n_pairs;
    const int mp_step = (static_cast<int>(gridDim.x) >> 1) / n_pairs;
    const int mp_tiles = (p.m_tiles + 1) >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tma_a[0]);
        tma_prefetch_desc(&p.tma_a[1]);
        tma_prefetch_desc(&p.tma_b[0]);
        if (out_slots) tma_prefetch_desc(&p.tma_out);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(&full_bar[i], 1);                 // the leader's producer arrives once (expect_tx for both CTAs' tiles)
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < WS2_ACC; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 2 * NUM_EPI_WARPS);      // both CTAs' epilogue warps
        }
        mbar_init(wres_bar, 1);
        fence_mbar_init();
    }
    if (warp >= 4) {
        for (int i = threadIdx.x; i < p.N && i < GN_PAR_MAX; i += NUM_EPI_WARPS * 32) {
            const float sc = lean ? 0.5f : 1.0f;      // ws_gn_silu_lean wants 0.5 gamma, 0.5 beta
            gn_par[i] = p.bias[i];
            gn_par[GN_PAR_MAX + i] = sc * p.gamma[i];
            gn_par[2 * GN_PAR_MAX + i] = sc * p.beta[i];
        }
    }
    __syncthreads();
    cluster_sync_all();                                  // the peer's barriers exist before anything is signalled at them
    if (warp == 2) tmem_alloc_pair(tmem_slot, 512);      // one warp of EACH CTA, same warp id, same destination offset
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs: own A rows, own W half; the leader's barriers count)
        if (lane == 0) {
            bool ok = true;
            if (leader_cta) mbar_arrive_expect_tx(wres_bar, 2 * total_kb * B_TILE_BYTES);
            int idx = 0;
            for (int s = 0; s < p.nseg; ++s) {
                const KSeg sg = p.seg[s];
                for (int kb = 0; kb < sg.nkb; ++kb, ++idx)
                    tma_load_2d_pair(&p.tma_b[sg.b_sel], s_w + idx * B_TILE_BYTES, wres_bar, sg.b_col + kb * BK, sg.b_row0 + n_pair * WS2_BN + static_cast<int>(crank) * BN);
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int mp = mp_first; ok && mp < mp_tiles; mp += mp_step) {
                const int m_blk = p.m_tile0 + 2 * mp + static_cast<int>(crank);
                for (int s = 0; s < p.nseg && ok; ++s) {
                    const KSeg sg = p.seg[s];
                    const CUtensorMap* ta = &p.tma_a[sg.a_sel];
                    for (int kb = 0; kb < sg.nkb; ++kb) {
                        if (!mbar_wait_relaxed(&empty_bar[stage], phase ^ 1u)) { ok = false; break; }
                        if (leader_cta) mbar_arrive_expect_tx(&full_bar[stage], 2 * A_TILE_BYTES);
                        tma_load_2d_pair(ta, s_a + stage * A_TILE_BYTES, &full_bar[stage], sg.a_col + kb * BK, m_blk * BM);
                        if (++stage == stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
            if (!ok) atomicExch(p.status, ERR_PRODUCER_TIMEOUT);
        }
    } else if (warp == 1) {
        // -------------------------------------------------------------- MMA issuer: the leader CTA's warp only
        if (leader_cta) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * BM, WS2_BN, 0, 0);
            const bool leader = elect_one();
            const uint64_t wdesc0 = make_kmajor_sw128_desc(smem_u32(s_w));
            const uint64_t adesc0 = make_kmajor_sw128_desc(smem_u32(s_a));
            bool ok = mbar_wait_relaxed(wres_bar, 0);
            tc_fence_after_sync();
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int mp = mp_first; ok && mp < mp_tiles; mp += mp_step, ++it) {
                const int acc = it & (WS2_ACC - 1);
                if (!mbar_wait_relaxed(&tempty_bar[acc], (static_cast<uint32_t>(it / WS2_ACC) & 1u) ^ 1u)) { ok = false; break; }
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * WS2_BN);
                for (int idx = 0; idx < total_kb; ++idx) {
                    if (!mbar_wait_relaxed(&full_bar[stage], phase)) { ok = false; break; }
                    tc_fence_after_sync();
                    if (leader) {
                        const uint64_t adesc = adesc0 + static_cast<uint64_t>((stage * A_TILE_BYTES) >> 4);
                        const uint64_t bdesc = wdesc0 + static_cast<uint64_t>((idx * B_TILE_BYTES) >> 4);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) umma_bf16_pair(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (idx | k) != 0 ? 1u : 0u);
                        umma_commit_pair(&empty_bar[stage]);
                    }
                    __syncwarp();
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
                if (ok && leader) umma_commit_pair(&tfull_bar[acc]);
                __syncwarp();
            }
            if (!ok && leader) atomicExch(p.status, ERR_MMA_TIMEOUT);
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- epilogue: this CTA's 128 rows x 256 columns, as two 128-column halves
        const int q = warp_phys & 3;
        const int part = warp_phys >> 2;
        bool ok = true;
        int it = 0;
        if (lean) {
            const int me = (q * 4 + part) * 32 + lane, other = (q * 4 + (part ^ 1)) * 32 + lane;
            const int bar_id = 1 + q * 2 + (part >> 1);
            const uint32_t stage_row = smem_u32(s_out + warp_phys * 2048) + lane * 64;
            const int sw = (lane >> 1) & 3;
            const float eps = p.gn_eps;
            const uint32_t tcol = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(part * 32);
            for (int mp = mp_first; mp < mp_tiles; mp += mp_step, ++it) {
                const int acc = it & (WS2_ACC - 1);
                const int row0 = (p.m_tile0 + 2 * mp + static_cast<int>(crank)) * BM + q * 32;
                if (!mbar_wait(&tfull_bar[acc], static_cast<uint32_t>(it / WS2_ACC) & 1u)) { ok = false; break; }
                tc_fence_after_sync();
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    const int col = n_pair * WS2_BN + half * BN + part * 32;
                    float v[32];
                    tmem_ld_32(tcol + static_cast<uint32_t>(acc * WS2_BN + half * BN), v);
                    if (half == 1) {      // the whole accumulator stage has been read by this warp: hand it back to the leader's MMA warp
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
                    }
                    uint32_t o[16];
                    ws_gn_silu_lean<GW>(v, gn_par + col, gn_xch, me, other, bar_id, eps, o);
                    if (out_slots) {
                        if (lane == 0) tma_store_wait_read<0>();      // the previous bulk store has read the staging box
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 4; ++j) st_shared_v4(stage_row + ((j ^ sw) << 4), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                        fence_proxy_async_smem();
                        __syncwarp();
                        // an odd number of row blocks leaves the last pair's rank 1 without a block of its own: those rows belong to someone else
                        if (lane == 0 && row0 < p.M) {
                            tma_store_2d(&p.tma_out, s_out + warp_phys * 2048, col, row0);
                            tma_store_commit();
                        }
                    } else if (row0 + lane < p.M) {
                        __nv_bfloat16* dst = p.out_bf + static_cast<size_t>(row0 + lane) * p.out_bf_ld + col;
#pragma unroll
                        for (int j = 0; j < 2; ++j) st_global_v8(dst + 16 * j, *reinterpret_cast<const uint32_t(*)[8]>(&o[8 * j]));
                    }
                }
            }
            if (out_slots && lane == 0) tma_store_wait_all<0>();      // all bulk stores complete before the CTA (and its shared memory) goes away
        } else
        for (int mp = mp_first; ok && mp < mp_tiles; mp += mp_step, ++it) {
            const int acc = it & (WS2_ACC - 1);
            const int row = (p.m_tile0 + 2 * mp + static_cast<int>(crank)) * BM + q * 32 + lane;
            if (!mbar_wait(&tfull_bar[acc], static_cast<uint32_t>(it / WS2_ACC) & 1u)) { ok = false; break; }
            tc_fence_after_sync();
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * WS2_BN + half * BN + part * 32);
                const int col = n_pair * WS2_BN + half * BN + part * 32;
                float v[32];
                tmem_ld_32(taddr, v);
                if (half == 1) {      // the whole accumulator stage has been read by this warp: hand it back to the leader's MMA warp
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
                }
                if (!(p.dbg & 64)) Epilogue<EPI_GN_SILU>::template run32<GW>(p, row, col, v, gn_par, gn_xch, q, part, lane);      // bit 6: timing probe, no epilogue
            }
        }
        if (!ok && lane == 0) atomicExch(p.status, ERR_EPI_TIMEOUT);
    }

    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                                  // no CTA leaves (or frees tensor memory) while its peer may still use it
    if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
}

// Eligible: what gemm_ws_eligible accepts, with an even number of 128-column tiles (whole 256-column slices).
inline bool gemm_ws2_eligible(const GemmParams& p) {
    return gemm_ws_eligible(p) && (p.n_tiles & 1) == 0 && p.n_tiles >= 2;
}

template <int GW>
int launch_gemm_ws2_inst(const GemmParams& p, int num_sms, cudaStream_t stream) {
    static PerDevice dev_state;
    if (!dev_state.configured()) {
        OSTEO_CUDA(cudaFuncSetAttribute(gemm_ws2_gn_silu_kernel<GW>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BYTES));
        cudaLaunchConfig_t qc = {};
        qc.gridDim = dim3(static_cast<unsigned>(num_sms & ~1), 1, 1);
        qc.blockDim = dim3(WS_THREADS, 1, 1);
        qc.dynamicSmemBytes = WS_SMEM_BYTES;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = 2;
        qa[0].val.clusterDim.y = 1;
        qa[0].val.clusterDim.z = 1;
        qc.attrs = qa;
        qc.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, gemm_ws2_gn_silu_kernel<GW>, &qc) == cudaSuccess) dev_state.val() = n;
        else cudaGetLastError();
        dev_state.set_configured();
    }
    const int max_pairs = dev_state.val();      // co-resident 2-CTA clusters on this device (0 = clusters unavailable / not queried)
    if (p.m_tiles <= 0) return 0;
    const int n_pairs = p.n_tiles / 2;
    if (max_pairs < n_pairs) return -2;
    const int mp_tiles = (p.m_tiles + 1) / 2;
    int per_slice = max_pairs / n_pairs;               // pairs per 256-column slice
    if (per_slice > mp_tiles) per_slice = mp_tiles;
    if (per_slice < 1) return -2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(2 * per_slice * n_pairs), 1, 1);
    cfg.blockDim = dim3(WS_THREADS, 1, 1);
    cfg.dynamicSmemBytes = WS_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    OSTEO_CUDA(cudaLaunchKernelEx(&cfg, gemm_ws2_gn_silu_kernel<GW>, p));
    return 0;
}

inline int launch_gemm_ws2(int gw, const GemmParams& p, int num_sms, cudaStream_t stream) {
    switch (gw) {
        case 16: return launch_gemm_ws2_inst<16>(p, num_sms, stream);
        case 32: return launch_gemm_ws2_inst<32>(p, num_sms, stream);
        case 64: return launch_gemm_ws2_inst<64>(p, num_sms, stream);
        default: return -2;
    }
}

}  // namespace osteo
