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
// Neither the noise NOR THE OLD STATE passes through the update warps: NOISE warps preload every eps accumulator stage in TMEM
// (tcgen05.st) with  - (sigma / c_eps) z - (c_x / c_eps) x_t , the eps MMAs accumulate on top of it, and the update is then
//   x_{t-1} = -c_eps (acc + b_out)         (two instructions per element; b_out comes from the constant bank -- it is part of the kernel
//                                           parameters -- so the update warps issue NO global load at all).
// The noise warps run up to four tiles ahead and wait ~1 000 cycles per tile for a free accumulator stage anyway (round-2 event trace),
// which hides the L2 latency of the state load; on the update warps that latency sat on the critical per-tile chain -- either under the
// fence.proxy.async that publishes the bf16 tile (it waits for every load the thread has in flight: 830 cycles) or at the first use.
// That splits the ALU-dense Philox / Box-Muller stream (16 warps, run up to four tiles ahead) from the latency-bound
// path (8 warps: state load, TMEM read, state store, bf16 tile publish).
//
// Work unit = one 128-row block, all DP / 64 column tiles. 896 threads, 1 CTA / SM, physical warp ids:
//   0-7    update warps   quadrant = warp % 4, 32 columns each (two passes of 16). Their per-tile chain (accumulator ready -> TMEM read ->
//                         bf16 tile published) is the kernel's critical path (round-2 event trace: ~3 000 cycles per tile, of which 830 were
//                         the fence.proxy.async waiting for the NEXT tile's state loads, issued just before it): the tile is published first,
//                         the prefetch loads of the next tile are issued after the arrive
//   8-23   noise warps    quadrant = warp % 4, 16 columns each (one ~1 500-cycle pass per tile and warp: sixteen of them are needed to hide it)
//   24     TMA producer: resident A (h_final, 128 x h0 bf16) once per unit, W_out tile [64 x h0] per column tile
//   25     MMA issuer: eps(j) += A . W_out[j]^T  (M128 N64),  acc_in += xbf(j-2) . W_in[j-2]^T  (M128 N=h0)
//   26     TMEM allocator (512 columns: acc_in 256 | 4 eps stages of 64)
//   27     TMA producer for W_in tiles [h0 x 64] + L2 prefetch of the state tile two tiles ahead
// The role warps have the highest warp ids on purpose: the scheduler arbiter favours high warp ids.
//
// State layout ("c8"): x[m_block][DP / 8][128 rows][8 cols] fp32, so one warp-wide 256-bit access (32 rows x 32 B)
// is 1 KB contiguous and a whole 128 x 64 tile is 32 KB contiguous.
#pragma once
#include "gemm_tc.cuh"

namespace osteo {

constexpr int FT = 64;                                  // state columns per tile
constexpr int F_UPD_WARPS = 8;
constexpr int F_NOISE_WARPS = 16;
constexpr int F_THREADS = 32 * (F_UPD_WARPS + F_NOISE_WARPS + 4);   // 896
constexpr int F_ARES_BYTES = 4 * A_TILE_BYTES;          // 64 KB: up to 4 k-blocks of [128 x 64] bf16
constexpr int F_WOUT_KB_BYTES = FT * BK * 2;            // 8 KB: one k-block of a W_out tile [64 x 64]
constexpr int F_WOUT_STAGE = 4 * F_WOUT_KB_BYTES;       // 32 KB
constexpr int F_WIN_STAGE = 256 * BK * 2;               // 32 KB: [256 x 64]
constexpr int F_XBF_STAGE = BM * BK * 2;                // 16 KB: [128 x 64]
constexpr int F_STAGES = 2;
constexpr int F_EPS_ACC = 4;
constexpr int F_BIAS_MAX = 6144;                        // b_out travels in the kernel parameters (constant bank): DP <= F_BIAS_MAX on the fused path
constexpr int F_PREFETCH = 6;                           // state tiles prefetched into L2 ahead of the next-step MMA position (sweep 2..16: 6 is best, -1.5 % eager, -0.5 % in the loop; >= 10 re-reads evicted lines)
constexpr int F_LAG = 2;                                // the next-step contraction trails the eps GEMM by two tiles
constexpr int F_TMEM_EPS0 = 256;                        // first TMEM column of the eps stages
constexpr int F_SMEM_BYTES = F_ARES_BYTES + F_STAGES * (F_WOUT_STAGE + F_WIN_STAGE + F_XBF_STAGE) + 1024 /*align*/ + 256 /*barriers*/;

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
    const float* bias_out;    // [>= DP], zero padded (device copy; the kernel reads bias_c)
    const float* noise;       // optional injected z, dense [M, noise_ld]
    int noise_ld;
    long long noise_step_stride;   // != 0: `noise` is a per-step stack; this step's slice starts at (noise_t0 - t) * stride (graph replays)
    int noise_t0;
    float* eps_out;           // optional dense eps [M, eps_ld]
    int eps_ld;
    unsigned long long seed;
    uint32_t rk[2 * PHILOX_ROUNDS_REVERSE];   // Philox round keys (seed_lo + r W0, seed_hi + r W1), r = 0..6: constant-bank operands of the XORs
    uint32_t one_bits;        // 0x3f800000 as a RUNTIME value: keeps it in a register, so (w & mask) | 1.0f is ONE LOP3 (an instruction has one immediate)
    long long row_base;
    const float* bias_in;     // [h0]
    const float* time_table;  // [T, h0]
    const float* cproj;       // [rows, h0]
    __nv_bfloat16* h0_out;    // [rows, h0_ld] bf16: the next step's first activation
    int h0_ld;
    // b_out, zero padded to DP, BY VALUE: kernel parameters live in the constant bank, the update warps read them with warp-uniform
    // indices (no global load on their chain, no registers held across the tile, no shared memory -- which is full). 24 KB of the 32 KB
    // a kernel may take (CUDA >= 12.1).
    float bias_c[F_BIAS_MAX];
    long long* trace;         // optional event trace of CTA 0 (diagnostics, OSTEO_DDPM_TRACE): [role 0..2][tile < 32][8] clock64 stamps
    int prefetch;             // state tiles prefetched into L2 ahead of the next-step MMA position (F_PREFETCH; OSTEO_FUSED_PF overrides)
    int dbg;                  // timing probes (scripts/ddpm_probe.py): 1 no L2 prefetch, 4 no state store, 8 no noise, 16 no state load
};

__device__ __forceinline__ void ld_global_v8(const void* ptr, uint32_t (&w)[8]) {
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(ptr)
                 : "memory");
}
// Streaming 256-bit load that does not allocate in the (28 KB, after the shared-memory carve-out) L1: the state is read exactly once per
// step, while the per-column parameters that ARE re-read by every row block (bias_out: 20 KB) should stay there.
__device__ __forceinline__ void ld_global_v8_stream(const void* ptr, uint32_t (&w)[8]) {
    asm volatile("ld.global.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(ptr)
                 : "memory");
}
__device__ __forceinline__ float4 ld_global_f4_keep(const float4* ptr) {
    float4 v;
    asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr));
    return v;
}
// Event trace of CTA 0 (role, tile, event -> clock64): compiled in only with -DOSTEO_FUSED_TRACE (build.py: OSTEO_NVCC_EXTRA), because even
// the predicated-off stamps cost ~7 instructions each in the per-tile loops. Printed by launch_fused when OSTEO_DDPM_TRACE is set.
#ifdef OSTEO_FUSED_TRACE
#define F_TRACE(role, tile, ev) do { if (p.trace && blockIdx.x == 0 && (tile) < 32 && lane == 0) p.trace[((role) * 32 + (tile)) * 8 + (ev)] = clock64(); } while (0)
#else
#define F_TRACE(role, tile, ev) do { } while (0)
#endif
__device__ __forceinline__ void l2_prefetch_bulk(const void* ptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
}
// registers -> TMEM: this warp's 32 lanes (rows) x 16 consecutive fp32 columns.
__device__ __forceinline__ void tmem_st_16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// out[i] = kb * base[i] + kr * z_i for the 16 columns [8*col8_0, 8*col8_0 + 16) of one row, z_i the PACKED Philox4x32-7 + Box-Muller normals of
// philox_normal_row_packed (philox.cuh: same counters, same uniforms; two Philox blocks, one word per pair). |kr| is folded into the
// Box-Muller radius, |kr| sqrt(-2 ln u1) = sqrt(k2 lg2 u1) with k2 = -2 ln2 kr^2; `neg` carries the sign of kr.
// rk = the round keys precomputed on the host (FusedParams::rk, read as constant-bank operands: no key-schedule adds in the loop);
// one = 0x3f800000 held in a register (see FusedParams::one_bits).
// `base` is consumed by the LAST instruction of each output (an FFMA whose addend is the finished r cos / r sin product), so when it
// comes straight from a global load that load's latency hides under the whole Philox / Box-Muller computation: written as
// fmaf(r, cos, kb * base) the compiler hoists the kb * base multiply to the top of the pass and stalls there for the load.
__device__ __forceinline__ void philox_axpy_normal16(const uint32_t (&rk)[2 * PHILOX_ROUNDS_REVERSE], uint32_t one, uint64_t row, uint32_t col8_0, uint32_t stream,
                                                     uint32_t step, float k2, bool neg, float kb, const uint32_t (&base)[16], float (&out)[16]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint4 c[2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
        c[i] = make_uint4(col8_0 + i, static_cast<uint32_t>(row), static_cast<uint32_t>(row >> 32), (stream << 16) | (step & 0xFFFFu));
#pragma unroll
    for (int r = 0; r < PHILOX_ROUNDS_REVERSE; ++r) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const uint32_t hi0 = __umulhi(M0, c[i].x), lo0 = M0 * c[i].x;
            const uint32_t hi1 = __umulhi(M1, c[i].z), lo1 = M1 * c[i].z;
            c[i] = make_uint4(hi1 ^ c[i].y ^ rk[2 * r], lo1, hi0 ^ c[i].w ^ rk[2 * r + 1], lo0);
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const uint32_t w[4] = {c[i].x, c[i].y, c[i].z, c[i].w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            // u1 in (0, 1] is never subnormal (>= 2^-20): the .ftz forms drop the denormal pre-scaling __log2f emits.
            // The angle is 2 pi f with f in [1, 2): one full turn more than 2 pi (f - 1), same sine and cosine, one FMUL instead of an FFMA.
            // Same bits as packed_radius_uniform / packed_angle_1_2 (philox.cuh): ((w >> 12) << 3) == (w >> 9) & 0x7ffff8.
            const float u1 = 2.0f - __uint_as_float(((w[h] >> 9) & 0x7FFFF8u) | one);
            const float th = __uint_as_float(((w[h] << 11) & 0x7FF800u) | one) * 6.283185307179586f;
            float l2, r, s, co;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(k2 * l2));
            asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
            asm("cos.approx.ftz.f32 %0, %1;" : "=f"(co) : "f"(th));
            r = neg ? -r : r;
            out[8 * i + 2 * h] = fmaf(kb, __uint_as_float(base[8 * i + 2 * h]), r * co);
            out[8 * i + 2 * h + 1] = fmaf(kb, __uint_as_float(base[8 * i + 2 * h + 1]), r * s);
        }
    }
}

// HOOKS = true: the parity-test variant that can also write eps out (p.eps_out); the production variant carries none of that code,
// which keeps the update warps inside the 72-register budget of an 896-thread CTA without spilling.
template <bool HOOKS>
__global__ void __launch_bounds__(F_THREADS, 1) ddpm_fused_kernel(const __grid_constant__ FusedParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic ON the __shared__ array: an integer round trip would turn every later access into a
    // generic LD / ST (address-space lookup in the LSU, several times slower than LDS / STS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* s_a = smem;
    uint8_t* s_wout = s_a + F_ARES_BYTES;
    uint8_t* s_win = s_wout + F_STAGES * F_WOUT_STAGE;
    uint8_t* s_xbf = s_win + F_STAGES * F_WIN_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_xbf + F_STAGES * F_XBF_STAGE);
    uint64_t* a_full = bars;                // resident A landed
    uint64_t* a_empty = bars + 1;           // every eps MMA of the unit has retired
    uint64_t* wout_full = bars + 2;         // [2] W_out tile landed
    uint64_t* win_full = bars + 4;          // [2] W_in tile landed
    uint64_t* xbf_full = bars + 6;          // [2] bf16 tile of x_{t-1} written by the update warps
    uint64_t* xbf_empty = bars + 8;         // [2] next-step MMA of that tile retired (frees the bf16 tile AND the W_in stage)
    uint64_t* tfull = bars + 10;            // [4] eps accumulator complete (also frees the W_out stage)
    uint64_t* tempty = bars + 14;           // [4] update warps have read the accumulator
    uint64_t* zfull = bars + 18;            // [4] noise warps have preloaded the accumulator stage
    uint64_t* accin_full = bars + 22;
    uint64_t* accin_empty = bars + 23;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int W_PROD = F_UPD_WARPS + F_NOISE_WARPS, W_MMA = W_PROD + 1, W_ALLOC = W_PROD + 2, W_WIN = W_PROD + 3;

    if (warp == W_PROD && lane == 0) {
        tma_prefetch_desc(&p.tma_a);
        tma_prefetch_desc(&p.tma_wout);
        tma_prefetch_desc(&p.tma_win);
    }
    if (warp == W_MMA && lane == 0) {
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int i = 0; i < F_STAGES; ++i) {
            mbar_init(&wout_full[i], 1);
            mbar_init(&win_full[i], 1);
            mbar_init(&xbf_full[i], F_UPD_WARPS);
            mbar_init(&xbf_empty[i], 1);
        }
        for (int i = 0; i < F_EPS_ACC; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], F_UPD_WARPS);
            mbar_init(&zfull[i], F_NOISE_WARPS);
        }
        mbar_init(accin_full, 1);
        mbar_init(accin_empty, F_UPD_WARPS);
        fence_mbar_init();
    }
    if (warp == W_ALLOC) tmem_alloc(tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int nt = p.n_tiles;

    if (warp == W_PROD) {
        // ------------------------------------------------ producer: resident A + W_out tiles
        if (lane == 0) {
            int itw = 0, k = 0;
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
                    if ((p.dbg & 32) && itw >= 2) { mbar_arrive(&wout_full[s]); continue; }      // timing probe: no weight traffic (wrong results)
                    mbar_arrive_expect_tx(&wout_full[s], p.nkb * F_WOUT_KB_BYTES);
                    for (int kb = 0; kb < p.nkb; ++kb)
                        tma_load_2d(&p.tma_wout, s_wout + s * F_WOUT_STAGE + kb * F_WOUT_KB_BYTES, &wout_full[s], kb * BK, j * FT);
                }
            }
            if (!ok) atomicExch(p.status, ERR_PRODUCER_TIMEOUT);
        }
    } else if (warp == W_WIN) {
        // ------------------------------------------------ producer: W_in tiles + L2 prefetch of the state
        if (lane == 0) {
            int iti = 0;
            bool ok = true;
            const size_t tile_floats = static_cast<size_t>(FT / 8) * BM * 8;      // 32 KB per tile
            for (int u = blockIdx.x; ok && u < p.m_tiles; u += gridDim.x) {
                const int m_blk = p.m_tile0 + u;
                const float* xblk = p.x + static_cast<size_t>(m_blk) * p.x_c8 * BM * 8;
                // the state is read by the NOISE warps, which run up to F_EPS_ACC tiles ahead of the tile whose next-step MMA paces this loop
                if (!(p.dbg & 1)) {
                    for (int j = 0; j < p.prefetch && j < nt; ++j) l2_prefetch_bulk(xblk + static_cast<size_t>(j) * tile_floats, static_cast<uint32_t>(tile_floats * 4));
                }
                for (int j = 0; j < nt; ++j, ++iti) {
                    const int s = iti & 1;
                    // stage s was last read by the next-step MMA of tile iti - 2, which also releases the bf16 tile buffer: one barrier
                    if (!mbar_wait_relaxed(&xbf_empty[s], ((static_cast<uint32_t>(iti) >> 1) & 1u) ^ 1u)) { ok = false; break; }
                    if (j + p.prefetch < nt && !(p.dbg & 1)) l2_prefetch_bulk(xblk + static_cast<size_t>(j + p.prefetch) * tile_floats, static_cast<uint32_t>(tile_floats * 4));
                    if ((p.dbg & 32) && iti >= 2) { mbar_arrive(&win_full[s]); continue; }
                    mbar_arrive_expect_tx(&win_full[s], p.h0 * BK * 2);
                    tma_load_2d(&p.tma_win, s_win + s * F_WIN_STAGE, &win_full[s], j * FT, 0);
                }
            }
            if (!ok) atomicExch(p.status, ERR_PRODUCER_TIMEOUT);
        }
    } else if (warp == W_MMA) {
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
                    // the stage holds b_out - (sigma / c_eps) z, written by the noise warps after the update warps released it
                    if (!mbar_wait_relaxed(&zfull[acc], (static_cast<uint32_t>(it_eps) >> 2) & 1u)) { ok = false; break; }
                    tc_fence_after_sync();
                    F_TRACE(0, it_eps, 0);
                    if (leader) {
                        const uint32_t d = tmem_base + F_TMEM_EPS0 + static_cast<uint32_t>(acc * FT);
                        const uint64_t bdesc_s = wout_desc0 + static_cast<uint64_t>((s * F_WOUT_STAGE) >> 4);
                        for (int kb = 0; kb < p.nkb; ++kb) {
                            const uint64_t adesc = adesc0 + static_cast<uint64_t>((kb * A_TILE_BYTES) >> 4);
                            const uint64_t bdesc = bdesc_s + static_cast<uint64_t>((kb * F_WOUT_KB_BYTES) >> 4);
#pragma unroll
                            for (int kk = 0; kk < BK / 16; ++kk) umma_bf16(d, adesc + 2u * kk, bdesc + 2u * kk, idesc_eps, 1u);
                        }
                        umma_commit(&tfull[acc]);
                    }
                    __syncwarp();
                    F_TRACE(0, it_eps, 1);
                    ++it_eps;
                }
                if (j >= F_LAG) {
                    const int jj = j - F_LAG;
                    const int s = it_in & 1;
                    if (jj == 0 && !mbar_wait_relaxed(accin_empty, (static_cast<uint32_t>(k) & 1u) ^ 1u)) { ok = false; break; }
                    if (!mbar_wait_relaxed(&xbf_full[s], (static_cast<uint32_t>(it_in) >> 1) & 1u)) { ok = false; break; }
                    if (!mbar_wait_relaxed(&win_full[s], (static_cast<uint32_t>(it_in) >> 1) & 1u)) { ok = false; break; }
                    tc_fence_after_sync();
                    F_TRACE(0, it_in, 2);
                    if (leader) {
                        const uint64_t adesc = xbf_desc0 + static_cast<uint64_t>((s * F_XBF_STAGE) >> 4);
                        const uint64_t bdesc = win_desc0 + static_cast<uint64_t>((s * F_WIN_STAGE) >> 4);
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk) umma_bf16(tmem_base, adesc + 2u * kk, bdesc + 2u * kk, idesc_in, (jj | kk) != 0 ? 1u : 0u);
                        umma_commit(&xbf_empty[s]);
                    }
                    __syncwarp();
                    F_TRACE(0, it_in, 3);
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
    } else if (warp >= F_UPD_WARPS && warp < W_PROD) {
        // ------------------------------------------------ noise warps: preload the eps accumulator stage
        const int q = warp & 3;                             // TMEM lane quadrant (hardware rule: warp id % 4)
        const int part = (warp - F_UPD_WARPS) >> 2;         // 16-column quarter of the 64-column tile
        const int r_tile = q * 32 + lane;
        const int t = *p.step;
        const float ce = __ldg(p.coef_eps + t);
        const float sg = (p.dbg & 8) ? 0.0f : __ldg(p.coef_sigma + t);
        // eps_out (parity hook) needs the raw eps in the accumulator: the update warps then add the noise themselves
        const float kr = (HOOKS && p.eps_out) ? 0.0f : -sg / ce;       // accumulator preload = kr z + kx x_t (the bias is added by the update warps)
        const float kx = (HOOKS && p.eps_out) ? 0.0f : -__ldg(p.coef_x + t) / ce;
        const float k2 = -1.3862943611198906f * kr * kr;
        const uint32_t one = p.one_bits;
        const float* nz_step = p.noise ? p.noise + static_cast<size_t>(p.noise_t0 - t) * p.noise_step_stride : nullptr;
        int it = 0;
        bool ok = true;
        uint32_t xn[16];                                    // x_t of the NEXT tile this warp will preload (see below)
#pragma unroll
        for (int i = 0; i < 16; ++i) xn[i] = 0u;
        if (static_cast<int>(blockIdx.x) < p.m_tiles && (p.m_tile0 + static_cast<int>(blockIdx.x)) * BM + r_tile < p.M && p.N > part * 16 && kx != 0.0f && !(p.dbg & 16)) {
            const float* xp = p.x + (static_cast<size_t>(p.m_tile0 + blockIdx.x) * p.x_c8 * BM + r_tile) * 8 + static_cast<size_t>((part * 16) >> 3) * (BM * 8);
            ld_global_v8_stream(xp, *reinterpret_cast<uint32_t(*)[8]>(&xn[0]));
            ld_global_v8_stream(xp + BM * 8, *reinterpret_cast<uint32_t(*)[8]>(&xn[8]));
        }
        for (int u = blockIdx.x; ok && u < p.m_tiles; u += gridDim.x) {
            const int m_blk = p.m_tile0 + u;
            const int row = m_blk * BM + r_tile;
            const bool live = row < p.M;
            for (int j = 0; j < nt && ok; ++j, ++it) {
                const int c0 = j * FT + part * 16;
                const int nvalid = p.N - c0;
                if (warp == F_UPD_WARPS) F_TRACE(1, it, 0);
                // base of the preload: kx x_t of these 16 columns, loaded ONE PASS AHEAD (xn holds the next tile's values while this tile is
                // computed): a DRAM round trip under load is ~2 us here, longer than a pass (padding columns of the state are exact zeros).
                uint32_t xr[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) xr[i] = xn[i];
                {
                    const bool more = j + 1 < nt;
                    const int c1 = more ? c0 + FT : part * 16;                              // next tile of this unit, or the first tile of the next unit
                    const int u1 = more ? u : u + static_cast<int>(gridDim.x);
                    const int row1 = (p.m_tile0 + u1) * BM + r_tile;
#pragma unroll
                    for (int i = 0; i < 16; ++i) xn[i] = 0u;
                    if (u1 < p.m_tiles && row1 < p.M && p.N > c1 && kx != 0.0f && !(p.dbg & 16)) {
                        const float* xp = p.x + (static_cast<size_t>(p.m_tile0 + u1) * p.x_c8 * BM + r_tile) * 8 + static_cast<size_t>(c1 >> 3) * (BM * 8);
                        ld_global_v8_stream(xp, *reinterpret_cast<uint32_t(*)[8]>(&xn[0]));
                        ld_global_v8_stream(xp + BM * 8, *reinterpret_cast<uint32_t(*)[8]>(&xn[8]));
                    }
                }
                float v[16];
                if (live && nvalid > 0 && kr != 0.0f && !nz_step) {
                    philox_axpy_normal16(p.rk, one, static_cast<uint64_t>(p.row_base + row), static_cast<uint32_t>(c0 >> 3), STREAM_REVERSE, static_cast<uint32_t>(t), k2,
                                         /*neg=*/true, kx, xr, v);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = kx * __uint_as_float(xr[i]);
                    if (live && nvalid > 0 && kr != 0.0f) {      // injected noise (parity runs)
                        const float* nz = nz_step + static_cast<size_t>(row) * p.noise_ld + c0;
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (i < nvalid) v[i] = fmaf(kr, nz[i], v[i]);
                    }
                }
                const int acc = it & (F_EPS_ACC - 1);
                if (warp == F_UPD_WARPS) F_TRACE(1, it, 1);
                if (!mbar_wait(&tempty[acc], ((static_cast<uint32_t>(it) >> 2) & 1u) ^ 1u)) { ok = false; break; }
                tc_fence_after_sync();
                if (warp == F_UPD_WARPS) F_TRACE(1, it, 2);
                tmem_st_16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(F_TMEM_EPS0 + acc * FT + part * 16), v);
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&zfull[acc]);
                if (warp == F_UPD_WARPS) F_TRACE(1, it, 3);
            }
        }
        if (!ok && lane == 0) atomicExch(p.status, ERR_EPI_TIMEOUT);
    } else if (warp < F_UPD_WARPS) {
        // ------------------------------------------------ update warps
        const int q = warp & 3;                 // TMEM lane quadrant
        const int half = warp >> 2;             // 32-column half of the 64-column tile
        const int r_tile = q * 32 + lane;
        const int t = *p.step;
        const float cx = __ldg(p.coef_x + t), nce = -__ldg(p.coef_eps + t);
        const float sg = (p.dbg & 8) ? 0.0f : __ldg(p.coef_sigma + t);
        const int t_next = t > 0 ? t - 1 : 0;
        const float* nz_step = (HOOKS && p.noise) ? p.noise + static_cast<size_t>(p.noise_t0 - t) * p.noise_step_stride : nullptr;
        int it = 0, k = 0;
        bool ok = true;
        for (int u = blockIdx.x; ok && u < p.m_tiles; u += gridDim.x, ++k) {
            const int m_blk = p.m_tile0 + u;
            const int row = m_blk * BM + r_tile;
            const bool live = row < p.M;
            float* xrow = p.x + (static_cast<size_t>(m_blk) * p.x_c8 * BM + r_tile) * 8;      // + c8 * (BM * 8)
            for (int j = 0; j < nt && ok; ++j, ++it) {
                if (warp == 0) F_TRACE(2, it, 0);
                const int cb = j * FT + half * 32;
                float* xp = xrow + static_cast<size_t>(cb >> 3) * (BM * 8);
                const int acc = it & (F_EPS_ACC - 1);
                if (!mbar_wait(&tfull[acc], (static_cast<uint32_t>(it) >> 2) & 1u)) { ok = false; break; }
                tc_fence_after_sync();
                if (warp == 0) F_TRACE(2, it, 1);
                const int s = it & 1;
                uint8_t* rowp = s_xbf + s * F_XBF_STAGE + (r_tile >> 3) * 1024 + (r_tile & 7) * 128;
                const int sw = r_tile & 7;
                // the whole 32-column accumulator slice in ONE TMEM round trip (no state registers compete for the space any more), then
                // hand the TMEM stage back to the noise warps
                uint32_t vr0[16], vr1[16];
                tmem_ld_16_nowait(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(F_TMEM_EPS0 + acc * FT + half * 32), vr0);
                tmem_ld_16_nowait(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(F_TMEM_EPS0 + acc * FT + half * 32 + 16), vr1);
                tmem_ld_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                if (warp == 0) F_TRACE(2, it, 2);
                if (!mbar_wait(&xbf_empty[s], ((static_cast<uint32_t>(it) >> 1) & 1u) ^ 1u)) { ok = false; break; }
                if (warp == 0) F_TRACE(2, it, 3);
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int c0 = cb + 16 * hh;
                    const int nvalid = p.N - c0;            // >= 16: all columns valid; <= 0: all padding
                    uint32_t vr[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) vr[i] = hh ? vr1[i] : vr0[i];
                    float xn[16];
                    if (!HOOKS || !p.eps_out) {
                        // the accumulator holds eps - (sigma / c_eps) z - (c_x / c_eps) x_t
#pragma unroll
                        for (int i = 0; i < 16; ++i) xn[i] = nce * (__uint_as_float(vr[i]) + p.bias_c[c0 + i]);
                    } else {
                        // parity hook: the accumulator holds the raw eps (+ bias); write it out, load the state and add the noise here
                        uint32_t xw[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) xw[i] = 0u;
                        if (live && nvalid > 0) {
                            ld_global_v8(xp + (2 * hh) * (BM * 8), *reinterpret_cast<uint32_t(*)[8]>(&xw[0]));
                            ld_global_v8(xp + (2 * hh + 1) * (BM * 8), *reinterpret_cast<uint32_t(*)[8]>(&xw[8]));
                        }
                        float z[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) z[i] = 0.0f;
                        if (live && nvalid > 0 && sg != 0.0f) {
                            if (nz_step) {
                                const float* nz = nz_step + static_cast<size_t>(row) * p.noise_ld + c0;
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    if (i < nvalid) z[i] = sg * nz[i];
                            } else {
                                const uint32_t zero16[16] = {};
                                philox_axpy_normal16(p.rk, p.one_bits, static_cast<uint64_t>(p.row_base + row), static_cast<uint32_t>(c0 >> 3), STREAM_REVERSE,
                                                     static_cast<uint32_t>(t), -1.3862943611198906f * sg * sg, /*neg=*/false, 0.0f, zero16, z);
                            }
                        }
                        if (live) {
                            float* eo = p.eps_out + static_cast<size_t>(row) * p.eps_ld + c0;
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (i < nvalid) eo[i] = __uint_as_float(vr[i]) + p.bias_c[c0 + i];
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i) xn[i] = fmaf(nce, __uint_as_float(vr[i]) + p.bias_c[c0 + i], fmaf(cx, __uint_as_float(xw[i]), z[i]));
                    }
                    if (nvalid < 16 || !live) {      // keep the padding columns (and rows past the batch) at exactly zero
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (!live || i >= nvalid) xn[i] = 0.0f;
                    }
                    // keep x_{t-1} in the accumulator's registers: it is stored AFTER the bf16 tile has been published
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (hh) vr1[i] = __float_as_uint(xn[i]);
                        else vr0[i] = __float_as_uint(xn[i]);
                    }
                    // bf16 tile for the next step's input_proj: row r_tile, 16-byte chunks of the 128-byte swizzled row
                    uint4 w0, w1;
                    w0.x = pack_bf16x2(xn[0], xn[1]);   w0.y = pack_bf16x2(xn[2], xn[3]);   w0.z = pack_bf16x2(xn[4], xn[5]);   w0.w = pack_bf16x2(xn[6], xn[7]);
                    w1.x = pack_bf16x2(xn[8], xn[9]);   w1.y = pack_bf16x2(xn[10], xn[11]); w1.z = pack_bf16x2(xn[12], xn[13]); w1.w = pack_bf16x2(xn[14], xn[15]);
                    const int ch = 4 * half + 2 * hh;
                    *reinterpret_cast<uint4*>(rowp + ((ch ^ sw) << 4)) = w0;
                    *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ sw) << 4)) = w1;
                }
                if (!ok) break;
                if (warp == 0) F_TRACE(2, it, 4);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&xbf_full[s]);
                if (warp == 0) F_TRACE(2, it, 5);
                // the fp32 state goes out last: nothing downstream in this kernel waits for it
                if (live && !(p.dbg & 4)) {
                    if (p.N > cb) {
                        st_global_v8(xp, *reinterpret_cast<uint32_t(*)[8]>(&vr0[0]));
                        st_global_v8(xp + (BM * 8), *reinterpret_cast<uint32_t(*)[8]>(&vr0[8]));
                    }
                    if (p.N > cb + 16) {
                        st_global_v8(xp + 2 * (BM * 8), *reinterpret_cast<uint32_t(*)[8]>(&vr1[0]));
                        st_global_v8(xp + 3 * (BM * 8), *reinterpret_cast<uint32_t(*)[8]>(&vr1[8]));
                    }
                }
            }
            if (!ok) break;
            // ---- unit end: h0 of the next step = acc_in + b_in + time_proj[t-1] + cond_proj  -> bf16
            if (!mbar_wait(accin_full, static_cast<uint32_t>(k) & 1u)) { ok = false; break; }
            tc_fence_after_sync();
            const int cpw = p.h0 >> 1;              // columns per warp half (64 or 128)
            for (int cc = 0; cc < cpw; cc += 16) {
                const int c = half * cpw + cc;
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
    if (warp == W_ALLOC) tmem_dealloc(tmem_base, 512);
}

}  // namespace osteo
