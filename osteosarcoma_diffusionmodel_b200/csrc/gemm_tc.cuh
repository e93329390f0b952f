// Persistent, warp-specialised tcgen05 GEMM for the denoiser's Linear layers
// (reference: models/diffusion.py:198-208 blocks, :229-232 input_proj + embeddings,
// :250 skip concat, :254 output_proj) with the layer's tail fused as the epilogue.
//
//   acc[m, n] = sum over K-segments s of  A_s[m, a_col_s : a_col_s + 64*nkb_s] . W[n, b_col_s : ...]
//
// A K-segment list expresses (i) the decoder's cat([h, skip]) as two segments
// accumulated into one TMEM accumulator and (ii) the split-bf16 "fp32x3" mode
// (hi*hi + hi*lo + lo*hi) as three segments over [hi | lo] operand buffers.
//
// Roles (384 threads, 1 CTA / SM, grid = min(#tiles, #SMs), static round-robin tiles):
//   warp 0      TMA producer     A tile 128x64 bf16 + W tile 128x64 bf16 per k-block, 6-stage ring
//   warp 1      MMA issuer       tcgen05.mma cta_group::1 kind::f16, M=128 N=128 K=16, fp32 accum in TMEM
//   warp 2      TMEM allocator   512 columns = 4 accumulator stages of 128 columns
//   warps 4-11  epilogue         thread <-> one accumulator row (TMEM lane), 64 columns per tile
// Row-per-thread is what makes GroupNorm(8) a purely in-register reduction.
#pragma once
#include <cuda_bf16.h>
#include "ptx.cuh"
#include "philox.cuh"

namespace osteo {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 64;
constexpr int STAGES_DEFAULT = 6;
constexpr int STAGES_DDPM = 3;            // K = 256 only: the freed shared memory stages the fp32 state tile instead
constexpr int X_BOX_COLS = 32;            // 32 fp32 = 128 B rows (SWIZZLE_128B)
constexpr int X_BOX_BYTES = BM * X_BOX_COLS * 4;          // 16 KB: [128 rows x 32 cols]
constexpr int X_TILE_BYTES = (BN / X_BOX_COLS) * X_BOX_BYTES;   // 64 KB: 4 boxes
constexpr int X_BUFFERS = 2;
constexpr int NUM_ACC = 4;
// Epilogue warps: 8 (thread <-> row x 64 columns: a whole GroupNorm group per thread) except for the RNG-heavy
// reverse-update epilogue, which runs 16 (thread <-> row x 32 columns) to double the warps per scheduler.
// ... and for Linear+GroupNorm+SiLU: with 8 warps (2 per scheduler) its epilogue is latency-bound at a fifth of the issue rate.
// A 64-wide group is then split over two warps (same rows, adjacent 32-column spans) that swap partial sums through shared memory.
template <int EPI, int GW>
__host__ __device__ constexpr int epi_warps_of() { return (EPI == 2 /*EPI_DDPM*/ || EPI == 1 /*EPI_GN_SILU*/) ? 16 : 8; }
constexpr int GN_PAR_MAX = 512;                         // widest layer whose bias / gamma / beta are kept in shared memory
constexpr int GN_PAR_BYTES = 3 * GN_PAR_MAX * 4;        // [bias | gamma | beta]
constexpr int GN_XCH_BYTES = 2 * 16 * 32 * 4;           // two exchange buffers, one float per epilogue thread
template <int EPI, int GW>
__host__ __device__ constexpr int gemm_threads() { return 128 + 32 * epi_warps_of<EPI, GW>(); }
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int B_TILE_BYTES = BN * BK * 2;
template <int EPI>
__host__ __device__ constexpr int stages_of() { return EPI == 2 /*EPI_DDPM*/ ? STAGES_DDPM : STAGES_DEFAULT; }
template <int EPI>
__host__ __device__ constexpr int gemm_smem_bytes() {
    return stages_of<EPI>() * (A_TILE_BYTES + B_TILE_BYTES) + (EPI == 2 ? X_BUFFERS * X_TILE_BYTES : 0) + 1024 /*align*/ + 256 /*barriers*/ +
           (EPI == 1 ? GN_PAR_BYTES + GN_XCH_BYTES : 0);
}
constexpr int MAX_KSEG = 8;

enum EpiKind : int {
    EPI_LINEAR = 0,   // bias + table-row add + matrix add -> bf16 [hi|lo] and/or fp32
    EPI_GN_SILU = 1,  // bias + GroupNorm(8) + affine + SiLU (+dropout) -> bf16 [hi|lo]
    EPI_DDPM = 2,     // eps = acc + bias; x <- c_x*x - c_eps*eps + sigma*z ; xb <- bf16(x)
    EPI_MSE = 3,      // eps = acc + bias; loss += sum (eps - noise)^2 ; grad <- scale*(eps - noise)
    EPI_RBF = 4,      // sum over tile of exp(-gamma*(na[m] + nb[n] - 2 acc))
    EPI_GN_BWD = 5,   // acc = d(block output): dropout/SiLU/GroupNorm backward -> d(pre-norm) bf16 + column partials
    EPI_WGRAD = 6     // acc = dY^T X (both operands MN-major, split over the batch): fp32 atomic accumulate
};

enum GemmError : int { ERR_NONE = 0, ERR_PRODUCER_TIMEOUT = 1, ERR_MMA_TIMEOUT = 2, ERR_EPI_TIMEOUT = 3 };

struct KSeg {
    int a_sel;   // which A tensor map (0 / 1)
    int a_col;   // first K column in A            (MN-major mode: first feature column of the A operand)
    int b_col;   // first K column in W            (MN-major mode: first feature column of the B operand)
    int nkb;     // number of 64-wide k-blocks     (MN-major mode: unused, the k range is the launch's row split)
    int b_sel;   // which B tensor map (0 / 1)
    int b_row0;  // row of W holding output column 0 (dgrad through a concatenated input reads a row range of W^T)
};

struct GemmParams {
    CUtensorMap tma_a[2];
    CUtensorMap tma_b[2];
    CUtensorMap tma_x_ld;     // EPI_DDPM: fp32 state, box 128 rows x 32 cols (load)
    CUtensorMap tma_x_st;     // EPI_DDPM: fp32 state, box  32 rows x 32 cols (per-warp store)
    int M, N;                 // valid rows / valid output columns
    int m_tiles, n_tiles;     // tiles this launch covers: m blocks [m_tile0, m_tile0 + m_tiles)
    int m_tile0;
    int nseg;
    KSeg seg[MAX_KSEG];
    int a_resident;           // EPI_DDPM, single K segment of <= 4 k-blocks: keep the A tile in shared memory across the n-tiles of an m-block
    int n_chunks;             //   ... and split each m-block's n-tiles into this many work units (load balance)
    int dbg;                  // diagnostic switches (0 in production; scripts/ddpm_probe.py)
    CUtensorMap tma_out;      // gemm_ws: output view (box 32 rows x 32 columns) for the per-warp TMA stores
    int out_tma;              // != 0: the weight-stationary kernels store their bf16 output through tma_out (OSTEO_WS_TMA_STORE; 2 = also the
                              // single-CTA kernel at K = 512 with half the warps staged, an experiment that measured slower)
    int a_blocked_nbox;       // > 0: tma_a[0] views a BLOCKED operand [m_tile][nbox][128 rows][64 cols] (16 KB contiguous per k-block)
    int* status;              // sticky error word (device)
    long long row_base;       // global row index of row 0 (RNG keying under row sharding)

    const float* bias;        // [N] or nullptr
    // EPI_LINEAR
    const float* add_tab;     // [rows, add_tab_ld] table; row = add_idx ? add_idx[m] : (step ? *step : 0)
    const int* add_idx;
    int add_tab_ld;
    const float* add_mat;     // [M, add_mat_ld] or nullptr
    int add_mat_ld;
    float* out_f32;           // optional
    int out_f32_ld;
    __nv_bfloat16* out_bf;    // optional; hi at col, lo at col + out_lo_off when out_lo_off > 0
    int out_bf_ld;
    int out_lo_off;
    // EPI_GN_SILU
    const float* gamma;
    const float* beta;
    float gn_eps;
    float drop_p;             // 0 = no dropout
    const uint8_t* drop_mask; // optional injected keep-mask [M, N] (parity tests)
    uint32_t drop_stream;
    __nv_bfloat16* xhat_bf;   // optional: normalised pre-affine activations saved for backward
    float* rstd_out;          // optional: [M, 8]
    // EPI_DDPM / shared
    const int* step;          // device-resident timestep (graph replay keeps the launch constant)
    const float* coef_x;      // [T]
    const float* coef_eps;    // [T]
    const float* coef_sigma;  // [T]
    float* x;                 // [M, x_ld] fp32 master state, updated in place
    int x_ld;
    __nv_bfloat16* xb;        // bf16 shadow of x, BLOCKED [m_tile][xb_nbox][128][64]: next step's input_proj operand
    int xb_nbox;              // boxes per m-tile (hi boxes, then lo boxes)
    int xb_lo_boxes;          // > 0: the bf16 residual of column box k goes to box k + xb_lo_boxes
    int x_nbox;               // fp32 state is BLOCKED [m_tile][x_nbox][128][32] (each TMA box = 16 KB contiguous)
    const float* noise;       // injected z [M, noise_ld] (parity) or nullptr (Philox)
    int noise_ld;
    long long noise_step_stride;   // != 0: `noise` is a per-step stack [steps][M, noise_ld]; the step's slice is (noise_t0 - t) * stride
    int noise_t0;                  //       (lets a replayed graph, whose only changing input is the device step word, consume injected noise)
    float* eps_out;           // optional fp32 eps [M, eps_ld]
    int eps_ld;
    unsigned long long seed;
    const unsigned long long* seed_dev;   // when set, the Philox key of the dropout masks is read from this device word (graph-replayed training step)
    // EPI_MSE
    const float* target;      // noise [M, target_ld]
    int target_ld;
    float grad_scale;
    double* loss_acc;
    // EPI_GN_BWD (reads gamma / beta / drop_* above)
    const __nv_bfloat16* xhat_in;   // saved normalised activations [M, xhat_ld] (hi at col, lo at col + xhat_lo_off)
    int xhat_ld;
    int xhat_lo_off;
    const float* rstd_in;     // [M, 8]
    float* col_partials;      // [ceil(M/32), nq, N] fp32 column sums over 32-row slabs; nq = 3 (dgamma, dbeta, dbias) for GN_BWD, 1 (dbias) otherwise
    // EPI_WGRAD
    int k_rows;               // contraction length (batch rows)
    int splits;               // row splits; tiles = splits * m_tiles * n_tiles
    int kb_per_split;
    // EPI_RBF
    const float* norm_a;      // [M]
    const float* norm_b;      // [N]
    float rbf_gamma;
    int rbf_symmetric;        // 1: only tiles with n_blk >= m_blk, off-diagonal weighted 2x
    double* rbf_acc;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_round(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }

// Store 32 consecutive values of this thread's row as bf16 (and the bf16 residual at +lo_off).
// 256-bit global store (sm_100+): one full 32-byte sector per lane.
__device__ __forceinline__ void st_global_v8(void* ptr, const uint32_t (&w)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]),
                 "r"(w[7])
                 : "memory");
}

// dst must be 32-byte aligned (activation pitches and column offsets are multiples of 16 elements). Two 256-bit stores: each lane
// writes full 32-byte sectors (its row is a different cache line from its neighbours', so narrower stores only half-fill sectors).
__device__ __forceinline__ void store_row32_bf16(__nv_bfloat16* dst, const float (&v)[32], int lo_off) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(v[16 * j + 2 * i], v[16 * j + 2 * i + 1]);
        st_global_v8(dst + 16 * j, w);
    }
    if (lo_off > 0) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float a = v[16 * j + 2 * i], b = v[16 * j + 2 * i + 1];
                w[i] = pack_bf16x2(a - bf16_round(a), b - bf16_round(b));
            }
            st_global_v8(dst + lo_off + 16 * j, w);
        }
    }
}

// v[j] += src[j], j < 32, through eight 128-bit read-only loads (src 16-byte aligned).
__device__ __forceinline__ void add_row32(float (&v)[32], const float* __restrict__ src) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 t = __ldg(s4 + j);
        v[4 * j + 0] += t.x;
        v[4 * j + 1] += t.y;
        v[4 * j + 2] += t.z;
        v[4 * j + 3] += t.w;
    }
}

__device__ __forceinline__ float silu_f(float y) { return __fdividef(y, 1.0f + __expf(-y)); }
// SiLU through one MUFU instead of two: y sigmoid(y) = h + h tanh(h), h = y / 2. tanh.approx is good to ~5e-4 absolute, far below the
// 2^-9 relative rounding of the bf16 value the result is stored as: used only when the output is plain bf16 (no [hi|lo] residual).
__device__ __forceinline__ float silu_fast(float y) {
    const float h = 0.5f * y;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

// One epilogue pass over 32 columns [col, col+32) of row `row` held in v[].
// GW = GroupNorm group width (16 / 32 / 64); a 64-wide group is handled by the caller
// passing precomputed statistics (mean / rstd over both halves).
template <int EPI>
struct Epilogue;

// Tile decode shared by the three roles. K-major mode: tile -> (m_blk, n_blk), the k range is the segment list.
// MN-major (wgrad) mode: tile -> (split, m_blk, n_blk), the k range is this split's slice of the batch rows.
struct TileInfo {
    int m_blk, n_blk, kb0, kb1;
    bool skip;
};
template <int EPI, bool MN>
__device__ __forceinline__ TileInfo decode_tile(const GemmParams& p, int tile) {
    TileInfo t;
    const int per = p.m_tiles * p.n_tiles;
    const int split = MN ? tile / per : 0;
    const int r = MN ? tile % per : tile;
    t.m_blk = p.m_tile0 + r / p.n_tiles;
    t.n_blk = r % p.n_tiles;
    t.kb0 = split * p.kb_per_split;
    const int total_kb = (p.k_rows + BK - 1) / BK;
    t.kb1 = t.kb0 + p.kb_per_split < total_kb ? t.kb0 + p.kb_per_split : total_kb;
    t.skip = (EPI == EPI_RBF) && p.rbf_symmetric && t.n_blk < t.m_blk;
    return t;
}

// The order in which one CTA walks its tiles; identical in the producer, MMA and epilogue roles.
//   default     tile = blockIdx.x, blockIdx.x + grid, ...  (round robin over (m, n) with n fastest)
//   A-resident  work unit = (m block, chunk of consecutive n tiles), units round robin over CTAs; inside a unit the n tiles
//               are consecutive so the A tile (128 rows x K) is loaded once and only W streams (halves L2 -> SMEM traffic).
template <int EPI, bool MN>
struct TileSeq {
    const GemmParams& p;
    bool ares;
    int num_tiles, tile;              // default order
    int units, unit, per_chunk, n, n_end;   // A-resident order
    __device__ TileSeq(const GemmParams& p_) : p(p_) {
        ares = (EPI == EPI_DDPM) && p.a_resident != 0;
        num_tiles = (MN ? p.splits : 1) * p.m_tiles * p.n_tiles;
        tile = static_cast<int>(blockIdx.x) - static_cast<int>(gridDim.x);
        per_chunk = (p.n_tiles + (p.n_chunks > 0 ? p.n_chunks : 1) - 1) / (p.n_chunks > 0 ? p.n_chunks : 1);
        units = p.m_tiles * (p.n_chunks > 0 ? p.n_chunks : 1);
        unit = static_cast<int>(blockIdx.x) - static_cast<int>(gridDim.x);
        n = n_end = 0;
    }
    __device__ bool next(TileInfo& ti, bool& first, bool& last) {
        if (!ares) {
            // the grid size is re-read from the special register every tile: kept in a register it was the value ptxas chose to spill,
            // and a local-memory reload per tile is an L2 round trip here (the L1 is carved out as shared memory)
            uint32_t nctas;
            asm volatile("mov.u32 %0, %%nctaid.x;" : "=r"(nctas));
            tile += static_cast<int>(nctas);
            if (tile >= num_tiles) return false;
            ti = decode_tile<EPI, MN>(p, tile);
            first = last = true;
            return true;
        }
        first = false;
        while (n >= n_end) {
            unit += gridDim.x;
            if (unit >= units) return false;
            const int chunk = unit % p.n_chunks;
            n = chunk * per_chunk;
            n_end = n + per_chunk < p.n_tiles ? n + per_chunk : p.n_tiles;
            first = true;
        }
        ti.m_blk = p.m_tile0 + unit / p.n_chunks;
        ti.n_blk = n;
        ti.kb0 = ti.kb1 = 0;
        ti.skip = false;
        ++n;
        last = n >= n_end;
        return true;
    }
};

template <int EPI, int GW, bool MN = false>
__global__ void __launch_bounds__(gemm_threads<EPI, GW>(), 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic ON the __shared__ array: an integer round trip would turn every later access into a
    // generic LD / ST (address-space lookup in the LSU, several times slower than LDS / STS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int STAGES = stages_of<EPI>();
    constexpr bool XSTAGE = (EPI == EPI_DDPM);
    constexpr int NUM_EPI_WARPS = epi_warps_of<EPI, GW>();
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_TILE_BYTES;
    uint8_t* smem_x = smem + STAGES * (A_TILE_BYTES + B_TILE_BYTES);      // 1024-aligned: stages are multiples of 32 KB
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_x + (XSTAGE ? X_BUFFERS * X_TILE_BYTES : 0));
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + NUM_ACC;
    uint64_t* xfull_bar = tempty_bar + NUM_ACC;
    uint64_t* xempty_bar = xfull_bar + X_BUFFERS;
    uint64_t* afull_bar = xempty_bar + X_BUFFERS;      // A-resident mode: A tile landed / A tile no longer read by any MMA
    uint64_t* aempty_bar = afull_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty_bar + 1);
    float* gn_par = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);      // EPI_GN_SILU: [bias | gamma | beta], GN_PAR_MAX each
    float* gn_xch = gn_par + 3 * GN_PAR_MAX;

    // Physical warp ids: epilogue warps first (0 .. NUM_EPI_WARPS-1), the four role warps LAST: the scheduler arbiter favours high warp
    // ids, and a TMA producer / MMA issuer starved by four busy epilogue warps on its scheduler stalls the whole pipeline.
    // `warp` is the logical id the code below uses (0 producer, 1 MMA, 2 TMEM allocator, 3 idle, 4.. epilogue); the TMEM lane quadrant
    // of an epilogue warp is physical id % 4 == logical id % 4.
    const int warp_phys = threadIdx.x >> 5;
    const int warp = warp_phys < NUM_EPI_WARPS ? warp_phys + 4 : warp_phys - NUM_EPI_WARPS;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tma_a[0]);
        tma_prefetch_desc(&p.tma_a[1]);
        tma_prefetch_desc(&p.tma_b[0]);
        tma_prefetch_desc(&p.tma_b[1]);
        if (XSTAGE) {
            tma_prefetch_desc(&p.tma_x_ld);
            tma_prefetch_desc(&p.tma_x_st);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < NUM_ACC; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], NUM_EPI_WARPS);
        }
        for (int i = 0; i < X_BUFFERS; ++i) {
            mbar_init(&xfull_bar[i], 1);
            mbar_init(&xempty_bar[i], NUM_EPI_WARPS);
        }
        mbar_init(afull_bar, 1);
        mbar_init(aempty_bar, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    if constexpr (EPI == EPI_GN_SILU) {
        // The layer's bias / gamma / beta (N <= 512) live in shared memory for the whole kernel: the L1 is carved out as shared memory,
        // so per-tile __ldg loads of them were L2 round trips on the epilogue's critical path.
        if (warp >= 4) {
            for (int i = threadIdx.x; i < p.N && i < GN_PAR_MAX; i += NUM_EPI_WARPS * 32) {
                gn_par[i] = p.bias[i];
                gn_par[GN_PAR_MAX + i] = p.gamma[i];
                gn_par[2 * GN_PAR_MAX + i] = p.beta[i];
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int xit = 0, uit = 0;
            bool ok = true;
            TileSeq<EPI, MN> seq(p);
            const int ring = seq.ares ? 2 : STAGES;      // A-resident: 64 KB resident A + two 16 KB W stages in the same 96 KB
            TileInfo ti;
            bool first, last;
            while (ok && seq.next(ti, first, last)) {
                if (ti.skip) continue;
                if (XSTAGE && !(p.dbg & 16)) {
                    // fp32 state tile of this output tile: 4 boxes of [128 rows x 32 cols], double buffered
                    const int xb = xit & 1;
                    const uint32_t xphase = static_cast<uint32_t>(xit >> 1) & 1u;
                    ++xit;
                    if (!mbar_wait_relaxed(&xempty_bar[xb], xphase ^ 1u)) { ok = false; break; }
                    mbar_arrive_expect_tx(&xfull_bar[xb], X_TILE_BYTES);
#pragma unroll
                    for (int b = 0; b < BN / X_BOX_COLS; ++b)
                        tma_load_2d(&p.tma_x_ld, smem_x + xb * X_TILE_BYTES + b * X_BOX_BYTES, &xfull_bar[xb], 0, (ti.m_blk * p.x_nbox + ti.n_blk * (BN / X_BOX_COLS) + b) * BM);
                }
                if (XSTAGE && seq.ares) {
                    const KSeg sg = p.seg[0];
                    if (first) {
                        if (!mbar_wait_relaxed(aempty_bar, (static_cast<uint32_t>(uit) & 1u) ^ 1u)) { ok = false; break; }
                        ++uit;
                        mbar_arrive_expect_tx(afull_bar, sg.nkb * A_TILE_BYTES);
                        for (int kb = 0; kb < sg.nkb; ++kb)
                            tma_load_2d(&p.tma_a[0], smem + kb * A_TILE_BYTES, afull_bar, sg.a_col + kb * BK, ti.m_blk * BM);
                    }
                    for (int kb = 0; kb < sg.nkb; ++kb) {
                        if (!mbar_wait_relaxed(&empty_bar[stage], phase ^ 1u)) { ok = false; break; }
                        mbar_arrive_expect_tx(&full_bar[stage], B_TILE_BYTES);
                        tma_load_2d(&p.tma_b[0], smem + 4 * A_TILE_BYTES + stage * B_TILE_BYTES, &full_bar[stage], sg.b_col + kb * BK, sg.b_row0 + ti.n_blk * BN);
                        if (++stage == ring) { stage = 0; phase ^= 1u; }
                    }
                    continue;
                }
                for (int s = 0; s < p.nseg && ok; ++s) {
                    const KSeg sg = p.seg[s];
                    const CUtensorMap* ta = &p.tma_a[sg.a_sel];
                    const CUtensorMap* tb = &p.tma_b[sg.b_sel];
                    const int kb_begin = MN ? ti.kb0 : 0, kb_end = MN ? ti.kb1 : sg.nkb;
                    for (int kb = kb_begin; kb < kb_end; ++kb) {
                        if (!mbar_wait_relaxed(&empty_bar[stage], phase ^ 1u)) { ok = false; break; }
                        mbar_arrive_expect_tx(&full_bar[stage], A_TILE_BYTES + B_TILE_BYTES);
                        uint8_t* sa = smem_a + stage * A_TILE_BYTES;
                        uint8_t* sb = smem_b + stage * B_TILE_BYTES;
                        if (MN) {
                            // operand tile = two boxes of [64 batch rows][64 features]; features are the MN dimension
                            tma_load_2d(ta, sa, &full_bar[stage], sg.a_col + ti.m_blk * BM, kb * BK);
                            tma_load_2d(ta, sa + A_TILE_BYTES / 2, &full_bar[stage], sg.a_col + ti.m_blk * BM + 64, kb * BK);
                            tma_load_2d(tb, sb, &full_bar[stage], sg.b_col + ti.n_blk * BN, kb * BK);
                            tma_load_2d(tb, sb + B_TILE_BYTES / 2, &full_bar[stage], sg.b_col + ti.n_blk * BN + 64, kb * BK);
                        } else {
                            if (p.a_blocked_nbox > 0 && sg.a_sel == 0)
                                tma_load_2d(ta, sa, &full_bar[stage], 0, (ti.m_blk * p.a_blocked_nbox + sg.a_col / BK + kb) * BM);
                            else
                                tma_load_2d(ta, sa, &full_bar[stage], sg.a_col + kb * BK, ti.m_blk * BM);
                            tma_load_2d(tb, sb, &full_bar[stage], sg.b_col + kb * BK, sg.b_row0 + ti.n_blk * BN);
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
            if (!ok) atomicExch(p.status, ERR_PRODUCER_TIMEOUT);
        }
    } else if (warp == 1) {
        // -------------------------------------------------------------- MMA issuer
        // The whole warp walks the tile sequence with warp-uniform control flow; only the tcgen05 instructions are predicated on
        // one elected lane. (Inside an `if (lane == 0)` region the compiler wraps every UTCHMMA / UTCBAR in an ELECT + branch loop
        // and rebuilds the descriptors through vector registers: ~10 dependent instructions per MMA from a single thread, which
        // cannot keep a 64-cycle MMA pipe full.)
        constexpr uint32_t idesc = make_idesc_bf16(BM, BN, MN ? 1 : 0, MN ? 1 : 0);
        const bool leader = elect_one();
        int stage = 0;
        uint32_t phase = 0;
        int it = 0, uit = 0;
        bool ok = true;
        TileSeq<EPI, MN> seq(p);
        TileInfo ti;
        bool first, last;
        while (ok && seq.next(ti, first, last)) {
            if (ti.skip) continue;
            const int acc = it % NUM_ACC;
            const uint32_t acc_phase = static_cast<uint32_t>(it / NUM_ACC) & 1u;
            ++it;
            if (!mbar_wait_relaxed(&tempty_bar[acc], acc_phase ^ 1u)) { ok = false; break; }
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
            uint32_t accumulate = 0;
            if (XSTAGE && seq.ares) {
                if (first) {
                    if (!mbar_wait_relaxed(afull_bar, static_cast<uint32_t>(uit) & 1u)) { ok = false; break; }
                    ++uit;
                    tc_fence_after_sync();
                }
                const int nkb = p.seg[0].nkb;
                for (int kb = 0; kb < nkb; ++kb) {
                    if (!mbar_wait_relaxed(&full_bar[stage], phase)) { ok = false; break; }
                    tc_fence_after_sync();
                    if (leader) {
                        const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(smem + kb * A_TILE_BYTES));
                        const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(smem + 4 * A_TILE_BYTES + stage * B_TILE_BYTES));
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (accumulate | static_cast<uint32_t>(k)) != 0 ? 1u : 0u);
                        umma_commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                    accumulate = 1;
                    if (++stage == 2) { stage = 0; phase ^= 1u; }
                }
                if (ok && leader) {
                    umma_commit(&tfull_bar[acc]);
                    if (last) umma_commit(aempty_bar);     // every MMA that reads the resident A tile has retired
                }
                __syncwarp();
                continue;
            }
            for (int s = 0; s < p.nseg && ok; ++s) {
                const int kb_begin = MN ? ti.kb0 : 0, kb_end = MN ? ti.kb1 : p.seg[s].nkb;
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    if (!mbar_wait_relaxed(&full_bar[stage], phase)) { ok = false; break; }
                    tc_fence_after_sync();
                    if (leader) {
                        const uint32_t a_addr = smem_u32(smem_a + stage * A_TILE_BYTES);
                        const uint32_t b_addr = smem_u32(smem_b + stage * B_TILE_BYTES);
                        if (MN) {
                            const uint64_t adesc = make_mnmajor_sw128_desc(a_addr, A_TILE_BYTES / 2);
                            const uint64_t bdesc = make_mnmajor_sw128_desc(b_addr, B_TILE_BYTES / 2);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                // 16 batch rows per UMMA_K step = 16 * 128 B inside each box
                                umma_bf16(d_tmem, adesc + 128u * k, bdesc + 128u * k, idesc, (accumulate | static_cast<uint32_t>(k)) != 0 ? 1u : 0u);
                            }
                        } else {
                            const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
                            const uint64_t bdesc = make_kmajor_sw128_desc(b_addr);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                // +32 B per UMMA_K step inside the 128-byte swizzle row (start-address field is >>4)
                                umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (accumulate | static_cast<uint32_t>(k)) != 0 ? 1u : 0u);
                            }
                        }
                        umma_commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                    accumulate = 1;
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
            if (ok && leader) umma_commit(&tfull_bar[acc]);
            __syncwarp();
        }
        if (!ok && leader) atomicExch(p.status, ERR_MMA_TIMEOUT);
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- epilogue
        const int q = warp & 3;                 // TMEM lane quadrant this warp may access
        const int part = (warp - 4) >> 2;       // which 64-column half (8 warps) / 32-column quarter (16 warps) of the tile
        constexpr int CPT = BN / (NUM_EPI_WARPS / 4);   // accumulator columns per thread
        int it = 0;
        bool ok = true;
        double thread_acc = 0.0;                // EPI_MSE / EPI_RBF partial sums
        int ddpm_t = 0;
        int pending_xb = -1;      // x buffer whose bulk store has been issued but not yet confirmed read
        float ddpm_cx = 0.f, ddpm_ce = 0.f, ddpm_sg = 0.f;
        if constexpr (XSTAGE) {                 // timestep and its three coefficients: once per kernel, not once per tile
            ddpm_t = *p.step;
            ddpm_cx = __ldg(p.coef_x + ddpm_t);
            ddpm_ce = __ldg(p.coef_eps + ddpm_t);
            ddpm_sg = __ldg(p.coef_sigma + ddpm_t);
        }
        TileSeq<EPI, MN> seq(p);
        TileInfo ti;
        bool first, last;
        while (ok && seq.next(ti, first, last)) {
            if (ti.skip) continue;
            const int acc = it % NUM_ACC;
            const uint32_t acc_phase = static_cast<uint32_t>(it / NUM_ACC) & 1u;
            ++it;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN + part * CPT);
            const int row = ti.m_blk * BM + q * 32 + lane;
            const int col = ti.n_blk * BN + part * CPT;
            if constexpr (XSTAGE) {
                // The noise of this tile depends only on (row, column, t): draw it BEFORE waiting for the accumulator and the
                // staged state tile, so the Philox / Box-Muller instruction stream hides those waits.
                float z[32];
                Epilogue<EPI>::draw_noise(p, row, col, ddpm_t, (p.dbg & 8) ? 0.0f : ddpm_sg, z);
                if (pending_xb >= 0) {
                    // the previous tile's bulk store has had the whole noise draw to read its shared-memory block: release it
                    if (lane == 0) {
                        tma_store_wait_read<0>();
                        mbar_arrive(&xempty_bar[pending_xb]);
                    }
                    pending_xb = -1;
                }
                if (!mbar_wait(&tfull_bar[acc], acc_phase)) { ok = false; break; }
                tc_fence_after_sync();
                const int xb = (it - 1) & 1;
                const uint32_t xphase = static_cast<uint32_t>((it - 1) >> 1) & 1u;
                if (!(p.dbg & 16) && !mbar_wait(&xfull_bar[xb], xphase)) { ok = false; break; }
                uint8_t* xt = smem_x + xb * X_TILE_BYTES;
                // two passes of 16 accumulator columns keep z[32] + v[16] + x[16] in registers without spilling
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t vr[16];
                    tmem_ld_16_nowait(taddr + 16 * hh, vr);
                    tmem_ld_wait();
                    if (hh == 1) {
                        // accumulator fully in registers: hand the TMEM stage back to the MMA warp
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                    }
                    if (!(p.dbg & 4))
                        Epilogue<EPI>::run_staged16(p, row, col + 16 * hh, q * 32 + lane, ti.m_blk, xt + part * X_BOX_BYTES, 4 * hh, vr, &z[16 * hh], ddpm_cx, ddpm_ce, ddpm_sg);
                }
                // make this warp's generic-proxy writes visible to the TMA engine, then store its 32 x 32 block
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&p.tma_x_st, xt + part * X_BOX_BYTES + q * 32 * 128, 0, (ti.m_blk * p.x_nbox + ti.n_blk * (BN / X_BOX_COLS) + part) * BM + q * 32);
                    tma_store_commit();
                }
                pending_xb = (p.dbg & 16) ? -1 : xb;      // released after the next tile's noise draw (or after the loop)
                __syncwarp();
            } else if constexpr (CPT == 32) {
                if (!mbar_wait(&tfull_bar[acc], acc_phase)) { ok = false; break; }
                tc_fence_after_sync();
                float v[32];
                tmem_ld_32(taddr, v);
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                Epilogue<EPI>::template run32<GW>(p, row, col, v, gn_par, gn_xch, q, part, lane);
            } else {
                if (!mbar_wait(&tfull_bar[acc], acc_phase)) { ok = false; break; }
                tc_fence_after_sync();
                float v0[32], v1[32];
                tmem_ld_32(taddr, v0);
                tmem_ld_32(taddr + 32, v1);
                // accumulator is in registers: hand the TMEM stage back to the MMA warp
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                Epilogue<EPI>::template run<GW>(p, row, col, v0, v1, thread_acc);
            }
        }
        if constexpr (XSTAGE) {
            if (lane == 0) {
                if (pending_xb >= 0) {
                    tma_store_wait_read<0>();
                    mbar_arrive(&xempty_bar[pending_xb]);
                }
                tma_store_wait_all<0>();   // all bulk stores complete before the CTA exits
            }
        }
        if (EPI == EPI_MSE || EPI == EPI_RBF) {
            // warp reduce then one atomic per warp
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) thread_acc += __shfl_xor_sync(0xffffffffu, thread_acc, o);
            if (lane == 0 && thread_acc != 0.0) atomicAdd(EPI == EPI_MSE ? p.loss_acc : p.rbf_acc, thread_acc);
        }
        if (!ok && lane == 0) atomicExch(p.status, ERR_EPI_TIMEOUT);
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// Column sums over the 32 rows a warp holds: each lane contributes v[0..31] (its row's 32 columns); after the
// butterfly (16+8+4+2+1 = 31 shuffles) lane L holds the total of column L. Rows that must not count pass zeros.
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
        const bool upper = (lane & w) != 0;
#pragma unroll
        for (int j = 0; j < w; ++j) {
            // keep the half of the columns whose bit `w` matches this lane's bit, send the other half across
            const float keep = upper ? v[j + w] : v[j];
            const float send = upper ? v[j] : v[j + w];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, w);
        }
    }
    return v[0];
}

// ------------------------------------------------------------------ epilogues
template <>
struct Epilogue<EPI_LINEAR> {
    template <int GW>
    __device__ static __forceinline__ void run(const GemmParams& p, int row, int col, float (&v0)[32], float (&v1)[32], double&) {
        const bool live = row < p.M;
        if (!live && !p.col_partials) return;      // warp-uniform only when partials are off; otherwise every lane continues
        const int lane = threadIdx.x & 31;
        const float* tab = nullptr;
        if (live && p.add_tab) {
            const int r = p.add_idx ? p.add_idx[row] : (p.step ? *p.step : 0);
            tab = p.add_tab + static_cast<size_t>(r) * p.add_tab_ld;
        }
        const float* mat = (live && p.add_mat) ? p.add_mat + static_cast<size_t>(row) * p.add_mat_ld : nullptr;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float(&v)[32] = h ? v1 : v0;
            const int c0 = col + 32 * h;
            if (c0 >= p.N) break;
            if (live && c0 + 32 <= p.N && ((p.add_tab_ld | p.add_mat_ld) & 3) == 0) {
                if (p.bias) add_row32(v, p.bias + c0);
                if (tab) add_row32(v, tab + c0);
                if (mat) add_row32(v, mat + c0);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int c = c0 + j;
                    if (live && c < p.N) {
                        float a = v[j];
                        if (p.bias) a += __ldg(p.bias + c);
                        if (tab) a += __ldg(tab + c);
                        if (mat) a += __ldg(mat + c);
                        v[j] = a;
                    } else {
                        v[j] = 0.0f;
                    }
                }
            }
            if (live && p.out_f32) {
                float* o = p.out_f32 + static_cast<size_t>(row) * p.out_f32_ld + c0;
                if (c0 + 32 <= p.N && (p.out_f32_ld & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(o)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < p.N) o[j] = v[j];
                }
            }
            if (live && p.out_bf) store_row32_bf16(p.out_bf + static_cast<size_t>(row) * p.out_bf_ld + c0, v, p.out_lo_off);
            if (p.col_partials) {
                const float sum = warp_column_sums(v, lane);
                if (c0 + lane < p.N) p.col_partials[static_cast<size_t>(row >> 5) * p.N + c0 + lane] = sum;
            }
        }
    }
};

#ifdef OSTEO_WS_TRACE
__device__ long long g_gn_stamp[4];      // diagnostics build: last writer wins, good enough to see the split inside one epilogue pass
#endif
template <>
struct Epilogue<EPI_GN_SILU> {
    // 16 epilogue warps: this thread owns 32 consecutive columns [col, col + 32) of one row.
    // GroupNorm(8, N) -> group width GW = N / 8 (16 / 32 / 64). Biased variance (two-pass), eps inside the sqrt
    // (torch.nn.GroupNorm, models/diffusion.py:202,206), then SiLU (:203,207) and the block's Dropout (:204) when drop_p > 0.
    // GW == 64: the group spans this warp and its neighbour (part ^ 1, same quadrant, same rows); the two swap their partial
    // sums through shared memory behind a 64-thread named barrier.
    template <int GW>
    __device__ static __forceinline__ void run32(const GemmParams& p, int row, int col, float (&v)[32], const float* par, float* xch, int q, int part, int lane) {
        const bool live = row < p.M;
        constexpr int NG = GW >= 32 ? 1 : 32 / GW;       // groups (or the half group) inside this thread's 32 columns
        constexpr int W = GW >= 32 ? 32 : GW;            // columns of one group held by this thread
        // the layer's parameters are ALWAYS in shared memory (N <= GN_PAR_MAX is enforced at launch): selecting between a shared and a
        // global pointer at run time would make every access a generic load
        const float* bias = par + col;
        const float* gamma = par + GN_PAR_MAX + col;
        const float* beta = par + 2 * GN_PAR_MAX + col;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 b = reinterpret_cast<const float4*>(bias)[j];
            v[4 * j + 0] += b.x;
            v[4 * j + 1] += b.y;
            v[4 * j + 2] += b.z;
            v[4 * j + 3] += b.w;
        }
        float mean[NG], rstd[NG];
        const int me = (q * 4 + part) * 32 + lane, other = (q * 4 + (part ^ 1)) * 32 + lane;
        const int bar_id = 1 + q * 2 + (part >> 1);
#ifdef OSTEO_WS_TRACE
#define GN_STAMP(i) g_gn_stamp[i] = clock64()
#else
#define GN_STAMP(i) do { } while (0)
#endif
        GN_STAMP(0);
        if (p.out_lo_off == 0 && !p.xhat_bf && !p.rstd_out && p.drop_p == 0.0f) {
            // bf16 inference fast path (warp-uniform): one-pass moments (sum and sum of squares in one sweep, one exchange for a 64-wide
            // group) and normalise + affine + the 0.5 of the tanh form of SiLU folded into one FFMA per element:
            //   h = v * (0.5 gamma rstd) + (0.5 beta - mean * 0.5 gamma rstd),   silu = h + h tanh(h).
            // 10 instructions per element instead of 13; the results go to bf16 (2^-9 relative), far above the ~1e-6 the one-pass
            // variance and the folded constants can move them.
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
                mean[g] = mu;
                rstd[g] = 0.5f * rsqrtf(fmaxf(fmaf(-mu, mu, ss * (1.0f / GW)), 0.0f) + p.gn_eps);      // 0.5 / sigma
            }
            GN_STAMP(1);
            if (!live) return;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 g = reinterpret_cast<const float4*>(gamma)[j], b = reinterpret_cast<const float4*>(beta)[j];
                const float gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i = 4 * j + e;
                    const float a = gg[e] * rstd[i / W];
                    const float h = fmaf(v[i], a, fmaf(-mean[i / W], a, 0.5f * bb[e]));
                    float t;
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
                    v[i] = fmaf(h, t, h);
                }
            }
            GN_STAMP(2);
            if (!(p.dbg & 256)) {      // bit 8: timing probe, no output store
                store_row32_bf16(p.out_bf + static_cast<size_t>(row) * p.out_bf_ld + col, v, 0);
            }
            GN_STAMP(3);
            return;
        }
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            float s = 0.0f;
#pragma unroll
            for (int j = 0; j < W; ++j) s += v[g * W + j];
            if constexpr (GW == 64) {
                xch[me] = s;
                asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
                s += xch[other];
            }
            const float mu = s * (1.0f / GW);
            float ss = 0.0f;
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const float d = v[g * W + j] - mu;
                ss = fmaf(d, d, ss);
            }
            if constexpr (GW == 64) {
                xch[512 + me] = ss;
                asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
                ss += xch[512 + other];
            }
            mean[g] = mu;
            rstd[g] = rsqrtf(ss * (1.0f / GW) + p.gn_eps);
        }
        if (!live) return;          // after the barriers: rows past the batch take part in the exchange but store nothing
        if (p.rstd_out) {
            if (GW < 64 || (part & 1) == 0) {
                const int g0 = col / GW;
#pragma unroll
                for (int g = 0; g < NG; ++g) p.rstd_out[static_cast<size_t>(row) * 8 + g0 + g] = rstd[g];
            }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (v[j] - mean[j / W]) * rstd[j / W];
        if (p.xhat_bf) store_row32_bf16(p.xhat_bf + static_cast<size_t>(row) * p.out_bf_ld + col, v, p.out_lo_off);
        if (p.out_lo_off == 0) {          // bf16 throughput mode (warp-uniform)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 g = reinterpret_cast<const float4*>(gamma)[j], b = reinterpret_cast<const float4*>(beta)[j];
                v[4 * j + 0] = silu_fast(fmaf(v[4 * j + 0], g.x, b.x));
                v[4 * j + 1] = silu_fast(fmaf(v[4 * j + 1], g.y, b.y));
                v[4 * j + 2] = silu_fast(fmaf(v[4 * j + 2], g.z, b.z));
                v[4 * j + 3] = silu_fast(fmaf(v[4 * j + 3], g.w, b.w));
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 g = reinterpret_cast<const float4*>(gamma)[j], b = reinterpret_cast<const float4*>(beta)[j];
                v[4 * j + 0] = silu_f(fmaf(v[4 * j + 0], g.x, b.x));
                v[4 * j + 1] = silu_f(fmaf(v[4 * j + 1], g.y, b.y));
                v[4 * j + 2] = silu_f(fmaf(v[4 * j + 2], g.z, b.z));
                v[4 * j + 3] = silu_f(fmaf(v[4 * j + 3], g.w, b.w));
            }
        }
        if (p.drop_p > 0.0f) {
            const float keep_scale = 1.0f / (1.0f - p.drop_p);
            if (p.drop_mask) {
                const uint8_t* mk = p.drop_mask + static_cast<size_t>(row) * p.N + col;
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = mk[j] ? v[j] * keep_scale : 0.0f;
            } else {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const uint4 w = philox_words(p.seed_dev ? *p.seed_dev : p.seed, static_cast<uint64_t>(p.row_base + row), static_cast<uint32_t>((col >> 2) + j4), p.drop_stream, p.step ? static_cast<uint32_t>(*p.step) : 0u);
                    v[4 * j4 + 0] = (u01(w.x) >= p.drop_p) ? v[4 * j4 + 0] * keep_scale : 0.0f;
                    v[4 * j4 + 1] = (u01(w.y) >= p.drop_p) ? v[4 * j4 + 1] * keep_scale : 0.0f;
                    v[4 * j4 + 2] = (u01(w.z) >= p.drop_p) ? v[4 * j4 + 2] * keep_scale : 0.0f;
                    v[4 * j4 + 3] = (u01(w.w) >= p.drop_p) ? v[4 * j4 + 3] * keep_scale : 0.0f;
                }
            }
        }
        store_row32_bf16(p.out_bf + static_cast<size_t>(row) * p.out_bf_ld + col, v, p.out_lo_off);
    }
};

template <>
struct Epilogue<EPI_DDPM> {
    // Reverse step (models/diffusion.py:400-423) collapsed to
    //   x <- c_x[t]*x - c_eps[t]*eps + sigma[t]*z ,  sigma[0] = 0 (the t == 0 branch returns x0_pred)
    // with the three fp32 tables derived in fp64 from the reference's fp32 buffers.
    // The fp32 state tile was brought in by TMA as four [128 rows x 32 cols] boxes with the 128-byte swizzle (16-byte chunk
    // index XOR (row & 7)); this thread owns row `r_tile` of ONE box (32 columns). x is updated IN PLACE in shared memory and
    // the caller TMA-stores the warp's 32 x 32 block. Global accesses left in here: the bf16 shadow (64 contiguous bytes per
    // thread) and the optional parity hooks.
    // z for this thread's (row, 32 columns): injected tensor (parity runs) or Philox4x32-10 + Box-Muller.
    __device__ static __forceinline__ void draw_noise(const GemmParams& p, int row, int c0, int t, float sg, float (&z)[32]) {
#pragma unroll
        for (int j = 0; j < 32; ++j) z[j] = 0.0f;
        if (row >= p.M || c0 >= p.N || sg == 0.0f) return;
        if (p.noise) {
            const float* nz = p.noise + static_cast<size_t>(p.noise_t0 - t) * p.noise_step_stride + static_cast<size_t>(row) * p.noise_ld + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c0 + j < p.N) z[j] = nz[j];
            return;
        }
        philox_normal_row_packed<4>(p.seed, static_cast<uint64_t>(p.row_base + row), static_cast<uint32_t>(c0 >> 3), STREAM_REVERSE, static_cast<uint32_t>(t), z);
    }

    // One pass = 16 columns [c0, c0 + 16) of row `row`: chunks j4_0 .. j4_0 + 3 of the 128-byte swizzled row in the staged box.
    // MASKED: the span crosses N (last column tile only).
    template <bool MASKED>
    __device__ static __forceinline__ void process16(const GemmParams& p, int row, int c0, int r_tile, int m_blk, uint8_t* xrow, int j4_0, const uint32_t (&vr)[16],
                                                     const float* z, float cx, float ce, float sg) {
        const int sw = r_tile & 7;
        float4 xv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) xv[j] = *reinterpret_cast<const float4*>(xrow + (((j4_0 + j) ^ sw) << 4));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + 4 * j;
            if (MASKED && c >= p.N) {
                xv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                continue;
            }
            float e[4];
            if (MASKED) {
#pragma unroll
                for (int i = 0; i < 4; ++i) e[i] = (c + i < p.N) ? __uint_as_float(vr[4 * j + i]) + __ldg(p.bias + c + i) : 0.0f;
            } else {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c));
                e[0] = __uint_as_float(vr[4 * j + 0]) + b4.x;
                e[1] = __uint_as_float(vr[4 * j + 1]) + b4.y;
                e[2] = __uint_as_float(vr[4 * j + 2]) + b4.z;
                e[3] = __uint_as_float(vr[4 * j + 3]) + b4.w;
            }
            if (p.eps_out) {
                float* eo = p.eps_out + static_cast<size_t>(row) * p.eps_ld + c;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (c + i < p.N) eo[i] = e[i];
            }
            float4 xn;
            xn.x = fmaf(sg, z[4 * j + 0], fmaf(cx, xv[j].x, -ce * e[0]));
            xn.y = fmaf(sg, z[4 * j + 1], fmaf(cx, xv[j].y, -ce * e[1]));
            xn.z = fmaf(sg, z[4 * j + 2], fmaf(cx, xv[j].z, -ce * e[2]));
            xn.w = fmaf(sg, z[4 * j + 3], fmaf(cx, xv[j].w, -ce * e[3]));
            if (MASKED) {
                if (c + 1 >= p.N) xn.y = 0.0f;
                if (c + 2 >= p.N) xn.z = 0.0f;
                if (c + 3 >= p.N) xn.w = 0.0f;
            }
            xv[j] = xn;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(xrow + (((j4_0 + j) ^ sw) << 4)) = xv[j];
        if (p.xb && !(p.dbg & 2)) {
            // blocked shadow: box (m_blk, c0 / 64), row r_tile, 16 consecutive bf16 = one full 32-byte sector
            __nv_bfloat16* xbrow = p.xb + ((static_cast<size_t>(m_blk) * p.xb_nbox + (c0 >> 6)) * BM + r_tile) * BK + (c0 & 63);
            if (!MASKED) {
                // one 256-bit store: a full 32-byte sector per lane
                uint32_t u[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    u[2 * j] = pack_bf16x2(xv[j].x, xv[j].y);
                    u[2 * j + 1] = pack_bf16x2(xv[j].z, xv[j].w);
                }
                st_global_v8(xbrow, u);
                if (p.xb_lo_boxes > 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        u[2 * j] = pack_bf16x2(xv[j].x - bf16_round(xv[j].x), xv[j].y - bf16_round(xv[j].y));
                        u[2 * j + 1] = pack_bf16x2(xv[j].z - bf16_round(xv[j].z), xv[j].w - bf16_round(xv[j].w));
                    }
                    st_global_v8(xbrow + static_cast<size_t>(p.xb_lo_boxes) * BM * BK, u);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (c0 + 8 * j >= p.N) break;
                    const float4 a = xv[2 * j], b = xv[2 * j + 1];
                    uint4 u;
                    u.x = pack_bf16x2(a.x, a.y);
                    u.y = pack_bf16x2(a.z, a.w);
                    u.z = pack_bf16x2(b.x, b.y);
                    u.w = pack_bf16x2(b.z, b.w);
                    reinterpret_cast<uint4*>(xbrow)[j] = u;
                    if (p.xb_lo_boxes > 0) {
                        uint4 l;
                        l.x = pack_bf16x2(a.x - bf16_round(a.x), a.y - bf16_round(a.y));
                        l.y = pack_bf16x2(a.z - bf16_round(a.z), a.w - bf16_round(a.w));
                        l.z = pack_bf16x2(b.x - bf16_round(b.x), b.y - bf16_round(b.y));
                        l.w = pack_bf16x2(b.z - bf16_round(b.z), b.w - bf16_round(b.w));
                        reinterpret_cast<uint4*>(xbrow + static_cast<size_t>(p.xb_lo_boxes) * BM * BK)[j] = l;
                    }
                }
            }
        }
    }

    __device__ static __forceinline__ void run_staged16(const GemmParams& p, int row, int col, int r_tile, int m_blk, uint8_t* xbox, int j4_0, const uint32_t (&vr)[16],
                                                        const float* z, float cx, float ce, float sg) {
        if (row >= p.M) return;            // rows past the batch: leave the staged tile as loaded
        uint8_t* xrow = xbox + r_tile * 128;
        if (col >= p.N) {                  // whole span is padding: keep it at zero
            const int sw = r_tile & 7;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(xrow + (((j4_0 + j) ^ sw) << 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
            return;
        }
        if (col + 16 <= p.N) process16<false>(p, row, col, r_tile, m_blk, xrow, j4_0, vr, z, cx, ce, sg);      // warp-uniform
        else process16<true>(p, row, col, r_tile, m_blk, xrow, j4_0, vr, z, cx, ce, sg);
    }

    template <int GW>
    __device__ static __forceinline__ void run(const GemmParams&, int, int, float (&)[32], float (&)[32], double&) {}
};

template <>
struct Epilogue<EPI_MSE> {
    // Training loss (models/diffusion.py:377): mean((eps_hat - noise)^2) over B*D; the epilogue
    // accumulates the sum in fp64 and emits d(loss)/d(eps_hat) = grad_scale * (eps_hat - noise)
    // plus its column sums over 32-row slabs (d(loss)/d(output_proj.bias)).
    template <int GW>
    __device__ static __forceinline__ void run(const GemmParams& p, int row, int col, float (&v0)[32], float (&v1)[32], double& acc) {
        const bool live = row < p.M;
        const int lane = threadIdx.x & 31;
        const float* trow = p.target + static_cast<size_t>(live ? row : 0) * p.target_ld;
        float local = 0.0f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float(&v)[32] = h ? v1 : v0;
            const int c0 = col + 32 * h;
            if (c0 >= p.N) break;      // warp-uniform
            if (live && c0 + 32 <= p.N && (p.target_ld & 3) == 0 && !p.eps_out) {
                // interior span: 128-bit loads of the bias and of this row's noise target
                add_row32(v, p.bias + c0);
                const float4* t4 = reinterpret_cast<const float4*>(trow + c0);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 t = __ldg(t4 + j);
                    const float d0 = v[4 * j + 0] - t.x, d1 = v[4 * j + 1] - t.y, d2 = v[4 * j + 2] - t.z, d3 = v[4 * j + 3] - t.w;
                    local = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, local))));
                    v[4 * j + 0] = d0 * p.grad_scale;
                    v[4 * j + 1] = d1 * p.grad_scale;
                    v[4 * j + 2] = d2 * p.grad_scale;
                    v[4 * j + 3] = d3 * p.grad_scale;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int c = c0 + j;
                    float d = 0.0f;
                    if (live && c < p.N) {
                        const float e = v[j] + __ldg(p.bias + c);
                        if (p.eps_out) p.eps_out[static_cast<size_t>(row) * p.eps_ld + c] = e;
                        d = e - trow[c];
                        local = fmaf(d, d, local);
                    }
                    v[j] = d * p.grad_scale;
                }
            }
            if (live && p.out_bf) store_row32_bf16(p.out_bf + static_cast<size_t>(row) * p.out_bf_ld + c0, v, p.out_lo_off);
            if (p.col_partials) {
                const float sum = warp_column_sums(v, lane);
                if (c0 + lane < p.N) p.col_partials[static_cast<size_t>(row >> 5) * p.N + c0 + lane] = sum;
            }
        }
        acc += static_cast<double>(local);
    }
};

template <>
struct Epilogue<EPI_RBF> {
    // RBF Gram tile reduction (utils/validation.py:288-294): sum exp(-gamma * ||a - b||^2) with
    // ||a - b||^2 = |a|^2 + |b|^2 - 2 a.b clamped at 0 (cdist never returns negatives).
    template <int GW>
    __device__ static __forceinline__ void run(const GemmParams& p, int row, int col, float (&v0)[32], float (&v1)[32], double& acc) {
        if (row >= p.M) return;
        const float na = p.norm_a[row];
        const float ng = -p.rbf_gamma * 1.4426950408889634f;   // exp(x) = exp2(x * log2 e)
        float local = 0.0f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float(&v)[32] = h ? v1 : v0;
            const int c0 = col + 32 * h;
            if (c0 >= p.N) break;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int c = c0 + j;
                if (c < p.N) {
                    const float d2 = fmaxf(na + __ldg(p.norm_b + c) - 2.0f * v[j], 0.0f);
                    local += exp2f(ng * d2);
                }
            }
        }
        float w = 1.0f;
        if (p.rbf_symmetric && (col / BN) != (row / BM)) w = 2.0f;
        acc += static_cast<double>(local * w);
    }
};

template <>
struct Epilogue<EPI_GN_BWD> {
    // Backward of Linear -> GroupNorm(8) -> SiLU (-> Dropout) for one row held by this thread.
    //   acc = dL/d(block output); ds = acc * mask / (1 - p); z = gamma * xhat + beta; dz = ds * silu'(z);
    //   dgamma += dz * xhat; dbeta += dz; dxhat = dz * gamma;
    //   dy = rstd * (dxhat - mean_g(dxhat) - xhat * mean_g(dxhat * xhat));  dbias += dy
    // (autograd of models/diffusion.py:201-207 with torch's native_group_norm_backward formulas).
    // dy goes out as bf16 [hi|lo]: the A operand of the next dgrad and the MN-major operand of this layer's wgrad.
    template <int GW>
    __device__ static __forceinline__ void run(const GemmParams& p, int row, int col, float (&v0)[32], float (&v1)[32], double&) {
        const bool live = row < p.M;
        const int lane = threadIdx.x & 31;
        constexpr int NG = 64 / GW;
        float xh0[32], xh1[32];
        if (live) {
            const __nv_bfloat16* xr = p.xhat_in + static_cast<size_t>(row) * p.xhat_ld + col;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float(&xh)[32] = h ? xh1 : xh0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint4 w = reinterpret_cast<const uint4*>(xr + 32 * h)[j];
                    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&ww[i]);
                        xh[8 * j + 2 * i] = __low2float(b2);
                        xh[8 * j + 2 * i + 1] = __high2float(b2);
                    }
                }
                if (p.xhat_lo_off > 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 w = reinterpret_cast<const uint4*>(xr + p.xhat_lo_off + 32 * h)[j];
                        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&ww[i]);
                            xh[8 * j + 2 * i] += __low2float(b2);
                            xh[8 * j + 2 * i + 1] += __high2float(b2);
                        }
                    }
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) { xh0[j] = 0.f; xh1[j] = 0.f; v0[j] = 0.f; v1[j] = 0.f; }
        }
        // dropout of the block output
        if (live && p.drop_p > 0.0f) {
            const float keep_scale = 1.0f / (1.0f - p.drop_p);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float(&v)[32] = h ? v1 : v0;
                const int c0 = col + 32 * h;
                if (p.drop_mask) {
                    const uint8_t* mk = p.drop_mask + static_cast<size_t>(row) * p.N + c0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = mk[j] ? v[j] * keep_scale : 0.0f;
                } else {
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const uint4 w = philox_words(p.seed_dev ? *p.seed_dev : p.seed, static_cast<uint64_t>(p.row_base + row), static_cast<uint32_t>((c0 >> 2) + j4), p.drop_stream, 0u);
                        v[4 * j4 + 0] = (u01(w.x) >= p.drop_p) ? v[4 * j4 + 0] * keep_scale : 0.0f;
                        v[4 * j4 + 1] = (u01(w.y) >= p.drop_p) ? v[4 * j4 + 1] * keep_scale : 0.0f;
                        v[4 * j4 + 2] = (u01(w.z) >= p.drop_p) ? v[4 * j4 + 2] * keep_scale : 0.0f;
                        v[4 * j4 + 3] = (u01(w.w) >= p.drop_p) ? v[4 * j4 + 3] * keep_scale : 0.0f;
                    }
                }
            }
        }
        // dz, then the two parameter-gradient column sums (dgamma, dbeta) over this warp's 32 rows
        const int slab = row >> 5;
        float* part = p.col_partials ? p.col_partials + static_cast<size_t>(slab) * 3 * p.N : nullptr;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float(&v)[32] = h ? v1 : v0;
            float(&xh)[32] = h ? xh1 : xh0;
            const int c0 = col + 32 * h;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float z = fmaf(xh[j], __ldg(p.gamma + c0 + j), __ldg(p.beta + c0 + j));
                const float sg = __fdividef(1.0f, 1.0f + __expf(-z));
                v[j] = live ? v[j] * sg * fmaf(z, 1.0f - sg, 1.0f) : 0.0f;     // dz
            }
            float t[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = v[j] * xh[j];
            const float sg_ = warp_column_sums(t, lane);
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = v[j];
            const float sb_ = warp_column_sums(t, lane);
            if (part) {
                part[c0 + lane] = sg_;
                part[p.N + c0 + lane] = sb_;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= __ldg(p.gamma + c0 + j);      // dxhat
        }
        // GroupNorm backward within each group of this row
        const float* rs = p.rstd_in + static_cast<size_t>(live ? row : 0) * 8 + col / GW;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            float m1 = 0.0f, m2 = 0.0f;
#pragma unroll
            for (int j = 0; j < GW; ++j) {
                const int idx = g * GW + j;
                const float dx = (idx < 32) ? v0[idx & 31] : v1[idx & 31];
                const float x = (idx < 32) ? xh0[idx & 31] : xh1[idx & 31];
                m1 += dx;
                m2 = fmaf(dx, x, m2);
            }
            m1 *= (1.0f / GW);
            m2 *= (1.0f / GW);
            const float r = live ? rs[g] : 0.0f;
#pragma unroll
            for (int j = 0; j < GW; ++j) {
                const int idx = g * GW + j;
                if (idx < 32) v0[idx & 31] = r * (v0[idx & 31] - m1 - xh0[idx & 31] * m2);
                else v1[idx & 31] = r * (v1[idx & 31] - m1 - xh1[idx & 31] * m2);
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float(&v)[32] = h ? v1 : v0;
            const int c0 = col + 32 * h;
            if (live) store_row32_bf16(p.out_bf + static_cast<size_t>(row) * p.out_bf_ld + c0, v, p.out_lo_off);
            float t[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = v[j];
            const float sdy = warp_column_sums(t, lane);
            if (part) part[2 * p.N + c0 + lane] = sdy;
        }
    }
};

template <>
struct Epilogue<EPI_WGRAD> {
    // dW[row = output feature, col = input feature] += acc. One launch may split the batch rows over several CTAs.
    template <int GW>
    __device__ static __forceinline__ void run(const GemmParams& p, int row, int col, float (&v0)[32], float (&v1)[32], double&) {
        if (row >= p.M) return;
        float* o = p.out_f32 + static_cast<size_t>(row) * p.out_f32_ld;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float(&v)[32] = h ? v1 : v0;
            const int c0 = col + 32 * h;
            if (c0 >= p.N) break;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c0 + j < p.N) atomicAdd(o + c0 + j, v[j]);
        }
    }
};

}  // namespace osteo
