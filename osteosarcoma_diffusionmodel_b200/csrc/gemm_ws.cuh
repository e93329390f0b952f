// Weight-stationary variant of the Linear + GroupNorm + SiLU GEMM (models/diffusion.py:198-208) for the bf16 throughput mode.
//
// The generic kernel (gemm_tc.cuh) streams a 128 x 64 A tile AND a 128 x 64 W tile per k-block for every 128 x 128 output tile:
// 64 FLOP per byte of L2 -> shared-memory traffic. At 100k rows the 512-wide layers then move ~800 MB per launch through the L2 at
// 10-12 TB/s, which is the L2's limit (ncu: DRAM 20-40 %, tensor pipe 21-38 %, nothing else saturated): they are L2-bandwidth bound.
// Here a CTA owns ONE 128-column slice of the layer for the whole launch and keeps that slice of W (128 x K bf16, K <= 512: <= 128 KB)
// resident in shared memory; only the A tiles of the row blocks it walks are streamed. That halves the L2 traffic and leaves the whole
// TMA ring (5 to 9 stages of 16 KB) to A.
//
// Roles and epilogue are those of gemm_tc_kernel<EPI_GN_SILU, GW> (16 epilogue warps + TMA producer, MMA issuer, TMEM allocator; the
// role warps have the highest warp ids). Launch parameters are the same GemmParams.
#pragma once
#include "gemm_tc.cuh"

namespace osteo {

constexpr int WS_MAX_KB = 8;                            // resident W slice: at most 8 k-blocks of [128 x 64] bf16 = 128 KB
constexpr int WS_RING_PLUS_RES = 13;                    // 16 KB slots shared by the resident slice and the A ring (208 KB)
constexpr int WS_MAX_STAGES = 12;
constexpr int WS_THREADS = 128 + 32 * 16;
constexpr int WS_SMEM_BYTES = WS_RING_PLUS_RES * A_TILE_BYTES + 1024 /*align*/ + 512 /*barriers*/ + GN_PAR_BYTES + GN_XCH_BYTES;

#ifdef OSTEO_WS_TRACE
__device__ long long g_ws_trace[3 * 32 * 4];      // [role][tile < 32][event] clock64 stamps of CTA 0 (diagnostics build only)
#define WS_TRACE(role, tile, ev) do { if (blockIdx.x == 0 && (tile) < 32 && lane == 0) g_ws_trace[((role) * 32 + (tile)) * 4 + (ev)] = clock64(); } while (0)
#else
#define WS_TRACE(role, tile, ev) do { } while (0)
#endif

// Lean GroupNorm + SiLU for the bf16 inference case of the weight-stationary kernel (no dropout, no saved x-hat / rstd): this thread owns 32
// consecutive columns of one row; `par` points at those columns' parameters in shared memory, stored by the kernel prologue as
// [bias | 0.5 gamma | 0.5 beta] (the 0.5 is the tanh form of SiLU: silu(y) = h + h tanh(h), h = y / 2). Per element: bias FADD, sum FADD, sum
// of squares FFMA, normalise FFMA (u = x rstd - mean rstd), affine FFMA (h = u g' + b'), MUFU.TANH, FFMA, half a pack = 7.5 instructions
// (the shared run32 spends 9.8 on it: it folds 0.5 into gamma / beta and builds a per-element scale and offset at run time). One-pass
// moments: the outputs are bf16 (2^-9), the one-pass variance moves them by ~1e-6. out[16] = the row's 32 outputs as packed bf16 pairs.
template <int GW>
__device__ __forceinline__ void ws_gn_silu_lean(float (&v)[32], const float* __restrict__ par, float* xch, int me, int other, int bar_id, float eps, uint32_t (&out)[16]) {
    constexpr int NG = GW >= 32 ? 1 : 32 / GW;
    constexpr int W = GW >= 32 ? 32 : GW;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 b = reinterpret_cast<const float4*>(par)[j];
        v[4 * j + 0] += b.x;
        v[4 * j + 1] += b.y;
        v[4 * j + 2] += b.z;
        v[4 * j + 3] += b.w;
    }
    float rstd[NG], nm[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        float s = 0.0f, ss = 0.0f;
#pragma unroll
        for (int j = 0; j < W; ++j) {
            s += v[g * W + j];
            ss = fmaf(v[g * W + j], v[g * W + j], ss);
        }
        if constexpr (GW == 64) {
            xch[me] = s;
            xch[512 + me] = ss;
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
            s += xch[other];
            ss += xch[512 + other];
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");      // the partner has read before the next tile overwrites
        }
        const float mu = s * (1.0f / GW);
        rstd[g] = rsqrtf(fmaxf(fmaf(-mu, mu, ss * (1.0f / GW)), 0.0f) + eps);
        nm[g] = -mu * rstd[g];
    }
    const float4* g2 = reinterpret_cast<const float4*>(par + GN_PAR_MAX);
    const float4* b2 = reinterpret_cast<const float4*>(par + 2 * GN_PAR_MAX);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 g = g2[j], b = b2[j];
        const float gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w};
        float y[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = 4 * j + e;
            const float h = fmaf(fmaf(v[i], rstd[i / W], nm[i / W]), gg[e], bb[e]);
            float t;
            asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
            y[e] = fmaf(h, t, h);
        }
        out[2 * j] = pack_bf16x2(y[0], y[1]);
        out[2 * j + 1] = pack_bf16x2(y[2], y[3]);
    }
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// CS = 2: the kernel is launched in clusters of two CTAs that own ADJACENT column slices and therefore walk the same row blocks: every A
// stage is fetched from the L2 once, by one of the two in turn, and multicast into both shared memories; a stage is recycled when BOTH
// CTAs' MMAs have consumed it (multicast tcgen05.commit onto both "empty" barriers). Without it every A tile crosses the L2 -> SM fabric
// N / 128 times (four times for the 512-wide layers), which is what bounds those layers.
template <int GW, int CS = 1>
__global__ void __launch_bounds__(WS_THREADS, 1) gemm_ws_gn_silu_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic ON the __shared__ array: an integer round trip would turn every later access into a
    // generic LD / ST (address-space lookup in the LSU, several times slower than LDS / STS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    int total_kb = 0;
    for (int s = 0; s < p.nseg; ++s) total_kb += p.seg[s].nkb;
    // fast path = bf16 inference (what run32 takes the one-pass route for): its output leaves through per-warp TMA stores staged in the last
    // 16 KB slots (2 KB per warp), so the epilogue registers are free the moment the values are in shared memory -- with st.global the next
    // tile's first register write waited for the previous tile's stores to drain (19 % of the block GEMMs, probe bit 8). K <= 384 gives up
    // two ring slots (all 16 warps staged, 7+ stages left); K = 512 cannot (3 stages starve the MMAs: measured 20 % slower): with
    // out_tma >= 2 it gives up one slot and stages the warps of column parts 0 and 1 only.
    const bool lean = p.out_lo_off == 0 && !p.xhat_bf && !p.rstd_out && p.drop_p == 0.0f && p.out_bf && !(p.dbg & 256);
    const bool tma_ok = CS == 1 && p.out_tma && lean;
    const int out_slots = !tma_ok ? 0 : total_kb <= 6 ? 2 : p.out_tma >= 2 ? 1 : 0;
    const int stages = WS_RING_PLUS_RES - total_kb - out_slots;     // 5 (K = 512) .. 9 (K = 256) without the staging
    uint8_t* s_w = smem;                                // resident W slice, k-block major
    uint8_t* s_a = smem + total_kb * B_TILE_BYTES;      // A ring
    uint8_t* s_out = smem + (WS_RING_PLUS_RES - out_slots) * A_TILE_BYTES;      // [8 * out_slots warps][32 rows][64 bytes]
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

    // this CTA's column slice and its row blocks m = m_first, m_first + m_step, ...
    const int n_blk = static_cast<int>(blockIdx.x) % p.n_tiles;
    const int m_first = static_cast<int>(blockIdx.x) / p.n_tiles;
    const int m_step = static_cast<int>(gridDim.x) / p.n_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tma_a[0]);
        tma_prefetch_desc(&p.tma_a[1]);
        tma_prefetch_desc(&p.tma_b[0]);
        if (out_slots) tma_prefetch_desc(&p.tma_out);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], CS);
        }
        for (int i = 0; i < NUM_ACC; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], NUM_EPI_WARPS);
        }
        mbar_init(wres_bar, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    if (warp >= 4) {
        for (int i = threadIdx.x; i < p.N && i < GN_PAR_MAX; i += NUM_EPI_WARPS * 32) {
            const float sc = lean ? 0.5f : 1.0f;      // ws_gn_silu_lean wants 0.5 gamma, 0.5 beta
            gn_par[i] = p.bias[i];
            gn_par[GN_PAR_MAX + i] = sc * p.gamma[i];
            gn_par[2 * GN_PAR_MAX + i] = sc * p.beta[i];
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if constexpr (CS > 1) cluster_sync_all();            // the peer's barriers exist before anything is multicast at them
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t crank = CS > 1 ? cluster_ctarank() : 0u;
    constexpr uint16_t CMASK = static_cast<uint16_t>((1u << CS) - 1u);

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            bool ok = true;
            mbar_arrive_expect_tx(wres_bar, total_kb * B_TILE_BYTES);
            int idx = 0;
            for (int s = 0; s < p.nseg; ++s) {
                const KSeg sg = p.seg[s];
                for (int kb = 0; kb < sg.nkb; ++kb, ++idx)
                    tma_load_2d(&p.tma_b[sg.b_sel], s_w + idx * B_TILE_BYTES, wres_bar, sg.b_col + kb * BK, sg.b_row0 + n_blk * BN);
            }
            int stage = 0;
            uint32_t phase = 0;
            int tcount = 0;
            uint32_t gstage = 0;                        // running stage count: CTA (gstage % CS) of the cluster fetches this one
            for (int m = m_first; ok && m < p.m_tiles; m += m_step, ++tcount) {
                const int m_blk = p.m_tile0 + m;
                WS_TRACE(0, tcount, 0);
                for (int s = 0; s < p.nseg && ok; ++s) {
                    const KSeg sg = p.seg[s];
                    const CUtensorMap* ta = &p.tma_a[sg.a_sel];
                    for (int kb = 0; kb < sg.nkb; ++kb) {
                        if (!mbar_wait_relaxed(&empty_bar[stage], phase ^ 1u)) { ok = false; break; }
                        if (p.dbg & 128) {      // timing probe: no A traffic, no MMAs (the epilogue alone; results are wrong)
                            mbar_arrive(&full_bar[stage]);
                            if (++stage == stages) { stage = 0; phase ^= 1u; }
                            continue;
                        }
                        mbar_arrive_expect_tx(&full_bar[stage], A_TILE_BYTES);
                        if constexpr (CS > 1) {
                            if (gstage % CS == crank) tma_load_2d_multicast(ta, s_a + stage * A_TILE_BYTES, &full_bar[stage], sg.a_col + kb * BK, m_blk * BM, CMASK);
                            ++gstage;
                        } else {
                            tma_load_2d(ta, s_a + stage * A_TILE_BYTES, &full_bar[stage], sg.a_col + kb * BK, m_blk * BM);
                        }
                        if (++stage == stages) { stage = 0; phase ^= 1u; }
                    }
                }
                WS_TRACE(0, tcount, 1);
            }
            if (!ok) atomicExch(p.status, ERR_PRODUCER_TIMEOUT);
        }
    } else if (warp == 1) {
        // -------------------------------------------------------------- MMA issuer (whole warp, one elected lane issues)
        constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
        const bool leader = elect_one();
        const uint64_t wdesc0 = make_kmajor_sw128_desc(smem_u32(s_w));
        const uint64_t adesc0 = make_kmajor_sw128_desc(smem_u32(s_a));
        bool ok = mbar_wait_relaxed(wres_bar, 0);
        tc_fence_after_sync();
        int stage = 0, it = 0;
        uint32_t phase = 0;
        for (int m = m_first; ok && m < p.m_tiles; m += m_step, ++it) {
            const int acc = it % NUM_ACC;
            if (!mbar_wait_relaxed(&tempty_bar[acc], (static_cast<uint32_t>(it / NUM_ACC) & 1u) ^ 1u)) { ok = false; break; }
            tc_fence_after_sync();
            WS_TRACE(1, it, 0);
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
            for (int idx = 0; idx < total_kb; ++idx) {
                if (!mbar_wait_relaxed(&full_bar[stage], phase)) { ok = false; break; }
                tc_fence_after_sync();
                if (idx == 0) WS_TRACE(1, it, 1);
                if (leader) {
                    const uint64_t adesc = adesc0 + static_cast<uint64_t>((stage * A_TILE_BYTES) >> 4);
                    const uint64_t bdesc = wdesc0 + static_cast<uint64_t>((idx * B_TILE_BYTES) >> 4);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        if (!(p.dbg & 128)) umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (idx | k) != 0 ? 1u : 0u);
                    if constexpr (CS > 1) umma_commit_multicast(&empty_bar[stage], CMASK);
                    else umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
            if (ok && leader) umma_commit(&tfull_bar[acc]);
            __syncwarp();
            WS_TRACE(1, it, 2);
        }
        if (!ok && leader) atomicExch(p.status, ERR_MMA_TIMEOUT);
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- epilogue (the generic kernel's 32-column GroupNorm epilogue)
        const int q = warp_phys & 3;
        const int part = warp_phys >> 2;
        const bool tma_out = warp_phys < 8 * out_slots;
        bool ok = true;
        int it = 0;
        if (lean) {
            // ---- bf16 inference: everything tile-invariant is hoisted, the epilogue is ws_gn_silu_lean, the output goes to the warp's staging
            // box + one TMA store (tma_out) or straight to global memory
            const float* par = gn_par + n_blk * BN + part * 32;
            const int me = (q * 4 + part) * 32 + lane, other = (q * 4 + (part ^ 1)) * 32 + lane;
            const int bar_id = 1 + q * 2 + (part >> 1);
            const uint32_t stage_warp = smem_u32(s_out + warp_phys * 2048);
            const uint32_t stage_row = stage_warp + lane * 64;
            const int sw = (lane >> 1) & 3;      // 64-byte swizzle of the staging box: a quarter warp's 16-byte chunks cover all 32 banks
            const float eps = p.gn_eps;
            const int col = n_blk * BN + part * 32;
            const uint32_t tcol = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(part * 32);
            for (int m = m_first; m < p.m_tiles; m += m_step, ++it) {
                const int acc = it % NUM_ACC;
                const int row0 = (p.m_tile0 + m) * BM + q * 32;
                if (warp_phys == 0) WS_TRACE(2, it, 0);
                if (!mbar_wait(&tfull_bar[acc], static_cast<uint32_t>(it / NUM_ACC) & 1u)) { ok = false; break; }
                tc_fence_after_sync();
                if (warp_phys == 0) WS_TRACE(2, it, 1);
                float v[32];
                tmem_ld_32(tcol + static_cast<uint32_t>(acc * BN), v);
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                if (warp_phys == 0) WS_TRACE(2, it, 2);
                uint32_t o[16];
                if (p.dbg & 64) {      // bit 6: timing probe, no epilogue math
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(v[2 * j]);
                } else {
                    ws_gn_silu_lean<GW>(v, par, gn_xch, me, other, bar_id, eps, o);
                }
                if (tma_out) {
                    // the bulk store of this warp's previous tile (issued a whole tile ago) has read the staging box
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; ++j) st_shared_v4(stage_row + ((j ^ sw) << 4), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&p.tma_out, s_out + warp_phys * 2048, col, row0);
                        tma_store_commit();
                    }
                } else if (row0 + lane < p.M) {
                    __nv_bfloat16* dst = p.out_bf + static_cast<size_t>(row0 + lane) * p.out_bf_ld + col;
#pragma unroll
                    for (int j = 0; j < 2; ++j) st_global_v8(dst + 16 * j, *reinterpret_cast<const uint32_t(*)[8]>(&o[8 * j]));
                }
                if (warp_phys == 0) WS_TRACE(2, it, 3);
            }
        } else
        for (int m = m_first; ok && m < p.m_tiles; m += m_step, ++it) {
            const int acc = it % NUM_ACC;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN + part * 32);
            const int row = (p.m_tile0 + m) * BM + q * 32 + lane;
            const int col = n_blk * BN + part * 32;
            if (warp_phys == 0) WS_TRACE(2, it, 0);
            if (!mbar_wait(&tfull_bar[acc], static_cast<uint32_t>(it / NUM_ACC) & 1u)) { ok = false; break; }
            tc_fence_after_sync();
            if (warp_phys == 0) WS_TRACE(2, it, 1);
            float v[32];
            tmem_ld_32(taddr, v);
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (warp_phys == 0) WS_TRACE(2, it, 2);
            if (!(p.dbg & 64)) Epilogue<EPI_GN_SILU>::template run32<GW>(p, row, col, v, gn_par, gn_xch, q, part, lane);      // bit 6: timing probe, no epilogue
            if (warp_phys == 0) WS_TRACE(2, it, 3);
        }
        if (tma_out && lane == 0) tma_store_wait_all<0>();      // all bulk stores complete before the CTA (and its shared memory) goes away
        if (!ok && lane == 0) atomicExch(p.status, ERR_EPI_TIMEOUT);
    }

    tc_fence_before_sync();
    __syncthreads();
    if constexpr (CS > 1) cluster_sync_all();            // no CTA leaves while its peer may still signal its barriers
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// Eligible: bf16 mode (no hi/lo passes), at most WS_MAX_KB k-blocks in total, the un-blocked A layout, enough row blocks to fill the grid.
inline bool gemm_ws_eligible(const GemmParams& p) {
    int total_kb = 0;
    for (int s = 0; s < p.nseg; ++s) total_kb += p.seg[s].nkb;
    return total_kb > 0 && total_kb <= WS_MAX_KB && p.a_blocked_nbox == 0 && p.out_lo_off == 0 && p.n_tiles >= 1 && p.n_tiles <= 8 && p.N <= GN_PAR_MAX;
}

template <int GW>
int launch_gemm_ws_inst(const GemmParams& p, int num_sms, cudaStream_t stream) {
    static PerDevice dev_state;
    static const bool want_cluster = getenv("OSTEO_WS_CLUSTER") && atoi(getenv("OSTEO_WS_CLUSTER")) != 0;      // opt-in until measured
    if (!dev_state.configured()) {
        OSTEO_CUDA(cudaFuncSetAttribute(gemm_ws_gn_silu_kernel<GW, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BYTES));
        OSTEO_CUDA(cudaFuncSetAttribute(gemm_ws_gn_silu_kernel<GW, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BYTES));
        if (want_cluster) {
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
            if (cudaOccupancyMaxActiveClusters(&n, gemm_ws_gn_silu_kernel<GW, 2>, &qc) == cudaSuccess) dev_state.val() = n;
            else cudaGetLastError();
        }
        dev_state.set_configured();
    }
    const int max_pairs = dev_state.val();      // co-resident 2-CTA clusters on this device (0 = clusters unavailable / not queried)
    if (p.m_tiles <= 0) return 0;
    int per_slice = num_sms / p.n_tiles;                // CTAs per column slice
    if (per_slice > p.m_tiles) per_slice = p.m_tiles;
    if (per_slice < 1) return -2;
    // clusters of two adjacent column slices: worth it only if (nearly) every SM can still be used
    if (max_pairs > 0 && (p.n_tiles & 1) == 0 && 2 * max_pairs >= (num_sms * 15) / 16) {
        int ps = (2 * max_pairs) / p.n_tiles;
        if (ps > per_slice) ps = per_slice;
        if (ps >= 1) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(static_cast<unsigned>(ps * p.n_tiles), 1, 1);
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
            OSTEO_CUDA(cudaLaunchKernelEx(&cfg, gemm_ws_gn_silu_kernel<GW, 2>, p));
            return 0;
        }
    }
    gemm_ws_gn_silu_kernel<GW, 1><<<per_slice * p.n_tiles, WS_THREADS, WS_SMEM_BYTES, stream>>>(p);
    OSTEO_CUDA(cudaGetLastError());
#ifdef OSTEO_WS_TRACE
    if (getenv("OSTEO_DDPM_TRACE")) {      // diagnostics build: print CTA 0's timeline of this launch (synchronises)
        static long long h[3 * 32 * 4];
        OSTEO_CUDA(cudaStreamSynchronize(stream));
        OSTEO_CUDA(cudaMemcpyFromSymbol(h, g_ws_trace, sizeof h));
        const long long t0 = h[0];
        long long g[4];
        OSTEO_CUDA(cudaMemcpyFromSymbol(g, g_gn_stamp, sizeof g));
        fprintf(stderr, "[ws trace] N=%d nseg=%d nkb0=%d GW=%d | last epilogue pass: bias+moments %lld  elementwise %lld  store %lld cycles\n", p.N, p.nseg, p.seg[0].nkb, GW,
                g[1] - g[0], g[2] - g[1], g[3] - g[2]);
        for (int t = 0; t < 12; ++t)
            fprintf(stderr, "[ws trace] tile %2d | PROD start %6lld end %6lld | MMA acc_free %6lld first_full %6lld issued %6lld | EPI arrive %6lld tfull %6lld tmem %6lld done %6lld\n", t,
                    h[(0 * 32 + t) * 4 + 0] - t0, h[(0 * 32 + t) * 4 + 1] - t0, h[(1 * 32 + t) * 4 + 0] - t0, h[(1 * 32 + t) * 4 + 1] - t0, h[(1 * 32 + t) * 4 + 2] - t0,
                    h[(2 * 32 + t) * 4 + 0] - t0, h[(2 * 32 + t) * 4 + 1] - t0, h[(2 * 32 + t) * 4 + 2] - t0, h[(2 * 32 + t) * 4 + 3] - t0);
    }
#endif
    return 0;
}

inline int launch_gemm_ws(int gw, const GemmParams& p, int num_sms, cudaStream_t stream) {
    switch (gw) {
        case 16: return launch_gemm_ws_inst<16>(p, num_sms, stream);
        case 32: return launch_gemm_ws_inst<32>(p, num_sms, stream);
        case 64: return launch_gemm_ws_inst<64>(p, num_sms, stream);
        default: return -2;
    }
}

}  // namespace osteo
