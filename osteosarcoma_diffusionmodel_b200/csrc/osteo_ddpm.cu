// C-ABI of the B200-native DDPM hot path (see include/osteo_ddpm.h).
// Context = repacked weights + workspace + prebuilt TMA descriptors; every compute entry
// point enqueues hand-written sm_100a kernels on the caller's stream. No CPU fallback.
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>
#include "../../include/osteo_ddpm.h"
#include "common.cuh"
#include "elem_kernels.cuh"
#include "fused_step.cuh"
#include "gemm_host.cuh"
#include "gemm_ws.cuh"
#include "gemm_ws2.cuh"
#include "gemm_rbf.cuh"
#include "train_kernels.cuh"
#include "validators.cuh"

namespace osteo {

// ------------------------------------------------------------------ common.cuh impl
std::string& last_error_ref() {
    static thread_local std::string s;
    return s;
}
int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return -1;
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0) return fail("tensor map operand not 16-byte aligned");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu)", static_cast<int>(r),
                                       static_cast<unsigned long long>(rows), static_cast<unsigned long long>(cols), static_cast<unsigned long long>(ld));
    return 0;
}

// bf16 row-major [rows, cols], box = 32 rows x 32 columns (64-byte rows, 64-byte swizzle: 16-byte chunk index XOR ((row >> 1) & 3), so the
// 32 lanes of a warp, one row each, write a chunk without bank conflicts): the per-warp output box of the TMA-store epilogue.
int make_tmap_bf16_st32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0) return fail("tensor map operand not 16-byte aligned");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (store box) failed with CUresult %d", static_cast<int>(r));
    return 0;
}

int make_tmap_f32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 4) % 16 != 0) return fail("tensor map operand not 16-byte aligned");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 4};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(X_BOX_COLS), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (fp32) failed with CUresult %d", static_cast<int>(r));
    return 0;
}

int sm_count(int device) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
    return n;
}

static inline int grid_for(long long work_items, int threads, int sms) {
    long long g = (work_items + threads - 1) / threads;
    const long long cap = static_cast<long long>(sms) * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

// One Linear layer repacked for the tensor cores: W bf16 [np, 2*kp] = [hi | lo], zero padded.
struct PackedLinear {
    int n = 0, k = 0, np = 0, kp = 0;
    DevBuf w, bias;     // bias fp32 [np]
    DevBuf wt;          // W^T bf16 [kp128, 2*np] for dgrad (training only)
    int ktp = 0;        // rows of wt (k rounded to 128)
    CUtensorMap tmap, tmap_t;
    int init(int n_, int k_) {
        n = n_; k = k_;
        np = static_cast<int>(round_up(n, BN));
        kp = static_cast<int>(round_up(k, BK));
        OSTEO_TRY(w.alloc(static_cast<size_t>(np) * 2 * kp * 2));
        OSTEO_TRY(bias.alloc(static_cast<size_t>(np) * 4));
        OSTEO_TRY(make_tmap_bf16(&tmap, w.p, np, 2 * kp, 2 * kp, BN));
        return 0;
    }
    int init_transposed() {
        ktp = static_cast<int>(round_up(k, BN));
        OSTEO_TRY(wt.alloc(static_cast<size_t>(ktp) * 2 * np * 2));
        OSTEO_TRY(make_tmap_bf16(&tmap_t, wt.p, ktp, 2 * np, 2 * np, BN));
        return 0;
    }
    int upload(const float* w_dev, const float* b_dev, int sms, cudaStream_t s) {
        const long long items = static_cast<long long>(np) * (kp / 4);
        pack_bf16_hilo_kernel<<<grid_for(items, 256, sms), 256, 0, s>>>(w_dev, n, k, k, w.as<__nv_bfloat16>(), np, kp, 2LL * kp, kp);
        OSTEO_CUDA(cudaGetLastError());
        OSTEO_CUDA(cudaMemsetAsync(bias.p, 0, static_cast<size_t>(np) * 4, s));
        OSTEO_CUDA(cudaMemcpyAsync(bias.p, b_dev, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToDevice, s));
        if (wt.p) {
            dim3 grid((np + 31) / 32, (ktp + 31) / 32), block(32, 8);
            // wt[r = k index, c = n index] = W[c, r]
            pack_bf16_hilo_transposed_kernel<<<grid, block, 0, s>>>(w_dev, k, n, k, wt.as<__nv_bfloat16>(), ktp, np, 2LL * np, np);
            OSTEO_CUDA(cudaGetLastError());
        }
        return 0;
    }
};

// Activation buffer bf16 [cap, 2*width] = [hi | lo] with its A-operand tensor map.
struct ActBuf {
    int width = 0;
    DevBuf buf;
    CUtensorMap tmap;          // A-operand view: box 128 rows x 64 columns, 128-byte swizzle
    CUtensorMap tmap_st;       // output view of the TMA-store epilogue: box 32 rows x 32 columns, 64-byte swizzle
    int init(long long cap, int width_) {
        width = width_;
        OSTEO_TRY(buf.alloc(static_cast<size_t>(cap) * 2 * width * 2));
        OSTEO_TRY(make_tmap_bf16(&tmap, buf.p, cap, 2 * width, 2 * width, BM));
        OSTEO_TRY(make_tmap_bf16_st32(&tmap_st, buf.p, cap, 2 * width, 2 * width));
        return 0;
    }
    __nv_bfloat16* ptr() const { return buf.as<__nv_bfloat16>(); }
};

struct HalfBlock {
    PackedLinear lin;
    DevBuf gamma, beta;   // fp32 [n]
    int gw = 0;
    int src0 = -1, src1 = -1;   // activation indices feeding this Linear (src1 = concatenated skip)
    int dst = -1;               // activation index written
    bool dropout = false;       // first half of a block carries the Dropout
    int block = 0;
};

}  // namespace osteo

using namespace osteo;

// A cached executable graph of one host-side enqueue sequence (weight repack, training step). The FIRST call with a given key runs
// eagerly (it may allocate, set function attributes ...), the second one captures the same sequence on an internal stream, and
// every later call with that key is a single cudaGraphLaunch. A different key drops the graph and starts over.
struct GraphSlot {
    std::vector<unsigned long long> key;
    cudaGraphExec_t exec = nullptr;
    long long launches = 0;
    void reset() {
        if (exec) cudaGraphExecDestroy(exec);
        exec = nullptr;
        key.clear();
    }
    ~GraphSlot() { reset(); }
};

struct osteo_ddpm_ctx {
    int device = 0, sms = 0;
    int D = 0, C = 0, TD = 0, E = 0, T = 0;
    std::vector<int> hidden;
    float drop_p = 0.f;
    int precision = OSTEO_PREC_BF16;
    int DP = 0;                 // padded feature pitch (multiple of 64)
    long long cap = 0;
    int chunk_rows = 131072;
    long long launches = 0;

    // parameters
    DevBuf ce_w0, ce_b0, ce_w2, ce_b2, cp_w, cp_b, tp_w, tp_b;   // fp32 copies of the small layers
    DevBuf tp_wt;                                                // time_proj.weight^T [TD, h0]
    DevBuf ce_w0t, ce_w2t, cp_wt;                                // ... and [in, out] transposes for the coalesced forward
    PackedLinear in_proj, out_proj;
    std::vector<std::unique_ptr<HalfBlock>> halves;
    DevBuf emb_table, time_table;        // [T, TD], [T, h0] fp32
    DevBuf sqrt_ab, sqrt_1mab, coef_x, coef_eps, coef_sigma;   // [T] fp32
    std::vector<float> h_coef_x, h_coef_eps, h_coef_sigma;
    bool have_weights = false, have_schedule = false, have_emb = false;

    // workspace (capacity `cap` rows)
    DevBuf x;                            // fp32 state, blocked [cap/128][x_nbox][128][32]
    CUtensorMap x_tmap_ld, x_tmap_st;    // TMA views of x: 128x32 load boxes, 32x32 store boxes
    DevBuf xb;                           // bf16 shadow, blocked [cap/128][xb_nbox][128][64] (hi boxes then lo boxes)
    CUtensorMap xb_tmap;
    int x_nbox = 0, xb_nbox = 0;
    // fused bf16 step (fused_step.cuh): output_proj + reverse update + the NEXT step's input_proj in one kernel
    CUtensorMap wout_tmap64, win_tmap;   // W_out as [64 x 64] boxes, W_in as [h0 x 64] boxes
    int fused_enable = 1;
    int ws_enable = getenv("OSTEO_DDPM_NO_WS") ? 0 : 1;
    int ws_tma_store = getenv("OSTEO_WS_TMA_STORE") ? atoi(getenv("OSTEO_WS_TMA_STORE")) : 1;      // block GEMMs write their output through per-warp TMA stores
    int ws2_enable = getenv("OSTEO_WS2") ? atoi(getenv("OSTEO_WS2")) : 1;      // 1: CTA-pair kernel for the K = 512, N >= 512 block GEMMs (measured faster there only); 2: wherever eligible; 0: never
    DevBuf fused_trace;
    bool x_c8 = false;                   // layout the state was loaded in: c8 (fused path) or 32-column boxes (TMA-staged path)
    bool shadow_valid = false;           // xb == bf16(x)? (the fused step does not maintain the shadow)
    bool h0_primed = false;              // acts[0] holds input_proj(x) + embeddings for timestep h0_t over rows [0, h0_n)
    int h0_t = -1;
    long long h0_n = 0;
    std::vector<std::unique_ptr<ActBuf>> acts;   // [0] = h0, then one per half block
    DevBuf cproj;                        // fp32 [cap, h0]
    DevBuf step_dev, status_dev;      // step_dev: MAX_BRANCHES step words (equal outside a graph replay), one per row branch
    int branches = 2, cur_branch = 0;
    long long noise_step_stride = 0;     // != 0 while a sample_loop with an injected per-step noise stack is being enqueued / captured
    int noise_t0 = 0;
    DevBuf loss_acc;                     // fp64 scalar
    TrainWorkspace train;

    // graph cache for sample_loop
    cudaGraphExec_t graph_exec = nullptr, graph_exec_multi = nullptr;   // 1 step / GRAPH_UNROLL steps
    long long graph_n = -1;
    unsigned long long graph_seed = 0;
    long long graph_row_base = 0;
    int graph_precision = -1, graph_chunk = -1, graph_fused = -1, graph_branches = -1, graph_noise_t0 = 0;
    int graph_nb = 0;                    // row branches the cached sampling graph was captured with
    const float* graph_noise = nullptr;
    unsigned long long graph_generation = ~0ull, graph_weights_version = ~0ull;
    long long graph_launches_per_step = 0;
    std::vector<cudaEvent_t>* prof = nullptr;   // when set, an event is recorded after every GEMM launch
    // graph caches of the training path (api_train.inl): weight repack after an optimizer step, forward + backward
    GraphSlot weights_graph, train_graph, train_fwd_graph, train_bwd_graph;      // the last two: the halves of the two-phase step
    GraphSlot train_bwd_part_graph[2];   // the backward pass cut in two (osteo_ddpm_train_backward_part: gradient all-reduce overlap)
    unsigned long long generation = 0;   // bumped by every (re)allocation of device buffers: part of every graph key (an address can be reused)
    cudaStream_t aux[3] = {nullptr, nullptr, nullptr};      // lanes of the weight repack
    cudaEvent_t aux_fork = nullptr, aux_join[3] = {nullptr, nullptr, nullptr};
    int train_graph_enable = 1;
    DevBuf seed_dev;                     // u64: Philox key of the graph-replayed training step

    std::vector<float> h_bias_out;       // host copy of output_proj.bias (zero padded): the fused kernel takes it by value (FusedParams::bias_c)
    bool h_bias_valid = false;
    unsigned long long weights_version = 0;   // bumped by set_weights: part of the sampling-graph key (the graphs embed bias_c)

    bool x3() const { return precision == OSTEO_PREC_FP32X3; }
    bool fused_ok() const { return fused_enable && precision == OSTEO_PREC_BF16 && hidden[0] <= 256 && DP <= F_BIAS_MAX; }
    int xs_nbox() const { return x_c8 ? DP / 8 : x_nbox; }
    int xs_shift() const { return x_c8 ? 3 : 5; }
    int h0() const { return hidden[0]; }
    int lo(int width) const { return x3() ? width : 0; }
    int lo_boxes() const { return x3() ? DP / BK : 0; }
    __nv_bfloat16* xb_ptr() const { return xb.as<__nv_bfloat16>(); }
    ~osteo_ddpm_ctx() {
        for (int i = 0; i < 3; ++i) {
            if (aux[i]) cudaStreamDestroy(aux[i]);
            if (aux_join[i]) cudaEventDestroy(aux_join[i]);
        }
        if (aux_fork) cudaEventDestroy(aux_fork);
        if (graph_exec) cudaGraphExecDestroy(graph_exec);
        if (graph_exec_multi) cudaGraphExecDestroy(graph_exec_multi);
    }
};

namespace osteo {

int check_ctx(const osteo_ddpm_ctx* c) {
    if (!c) return fail("null context");
    return 0;
}

static void base_params(const osteo_ddpm_ctx* c, GemmParams& p) {
    std::memset(&p, 0, sizeof p);
    p.status = c->status_dev.as<int>();
    p.step = c->step_dev.as<int>() + c->cur_branch;
    p.gn_eps = 1e-5f;
}

static int after_launch(osteo_ddpm_ctx* c, int rc, cudaStream_t s) {
    ++c->launches;
    if (rc == 0 && c->prof) {
        cudaEvent_t e;
        OSTEO_CUDA(cudaEventCreate(&e));
        OSTEO_CUDA(cudaEventRecord(e, s));
        c->prof->push_back(e);
    }
    return rc;
}

static void set_rows(GemmParams& p, long long row0, long long row1) {
    p.M = static_cast<int>(row1);
    p.m_tile0 = static_cast<int>(row0 / BM);
    p.m_tiles = static_cast<int>((row1 - row0 + BM - 1) / BM);
}

// input_proj + time/cond embedding add (models/diffusion.py:229-232) for rows [row0, row1).
static int launch_input_proj(osteo_ddpm_ctx* c, long long row0, long long row1, const int* t_idx, cudaStream_t s, const CUtensorMap* rowmajor_a = nullptr) {
    GemmParams p;
    base_params(c, p);
    if (rowmajor_a) {            // training: x_t lives in a row-major [rows, 2*DP] buffer (it is also the MN-major wgrad operand)
        p.tma_a[0] = p.tma_a[1] = *rowmajor_a;
    } else {                     // sampling / denoise: the blocked bf16 shadow of the state
        p.tma_a[0] = p.tma_a[1] = c->xb_tmap;
        p.a_blocked_nbox = c->xb_nbox;
    }
    p.tma_b[0] = p.tma_b[1] = c->in_proj.tmap;
    OSTEO_TRY(add_segments(p, 0, 0, c->DP, 0, c->in_proj.kp, c->DP, c->x3()));
    set_rows(p, row0, row1);
    p.N = c->h0();
    p.n_tiles = c->in_proj.np / BN;
    p.bias = c->in_proj.bias.as<float>();
    p.add_tab = c->time_table.as<float>();
    p.add_tab_ld = c->h0();
    p.add_idx = t_idx;
    p.add_mat = c->cproj.as<float>();
    p.add_mat_ld = c->h0();
    p.out_bf = c->acts[0]->ptr();
    p.out_bf_ld = 2 * c->h0();
    p.out_lo_off = c->lo(c->h0());
    return after_launch(c, launch_gemm(EPI_LINEAR, 64, p, c->sms, s), s);
}

struct HalfOpts {
    bool train = false;
    const uint8_t* drop_mask = nullptr;
    unsigned long long seed = 0;
    const unsigned long long* seed_dev = nullptr;
    long long row_base = 0;
    bool save = false;
};

// Linear + GroupNorm + SiLU (+Dropout) half block (models/diffusion.py:201-207; concat at :250).
static int launch_half(osteo_ddpm_ctx* c, int hi, long long row0, long long row1, const HalfOpts& o, cudaStream_t s) {
    HalfBlock& hb = *c->halves[hi];
    GemmParams p;
    base_params(c, p);
    const ActBuf& a0 = *c->acts[hb.src0];
    p.tma_a[0] = a0.tmap;
    p.tma_a[1] = a0.tmap;
    p.tma_b[0] = p.tma_b[1] = hb.lin.tmap;
    if (c->x3()) {
        // keep the three passes of each source adjacent: hi*hi, hi*lo, lo*hi
        OSTEO_TRY(add_segments(p, 0, 0, a0.width, 0, hb.lin.kp, a0.width, true));
        if (hb.src1 >= 0) {
            const ActBuf& a1 = *c->acts[hb.src1];
            p.tma_a[1] = a1.tmap;
            OSTEO_TRY(add_segments(p, 1, 0, a1.width, a0.width, hb.lin.kp, a1.width, true));
        }
    } else {
        OSTEO_TRY(add_segments(p, 0, 0, 0, 0, 0, a0.width, false));
        if (hb.src1 >= 0) {
            const ActBuf& a1 = *c->acts[hb.src1];
            p.tma_a[1] = a1.tmap;
            OSTEO_TRY(add_segments(p, 1, 0, 0, a0.width, 0, a1.width, false));
        }
    }
    set_rows(p, row0, row1);
    p.N = hb.lin.n;
    p.n_tiles = hb.lin.np / BN;
    p.bias = hb.lin.bias.as<float>();
    p.gamma = hb.gamma.as<float>();
    p.beta = hb.beta.as<float>();
    ActBuf& dst = *c->acts[hb.dst];
    p.out_bf = dst.ptr();
    p.out_bf_ld = 2 * dst.width;
    p.out_lo_off = c->lo(dst.width);
    if (o.train && hb.dropout && c->drop_p > 0.f) {
        p.drop_p = c->drop_p;
        p.drop_mask = o.drop_mask;
        p.drop_stream = STREAM_DROPOUT + static_cast<uint32_t>(hb.block);
        p.seed = o.seed;
        p.seed_dev = o.seed_dev;
        p.row_base = o.row_base;
        p.step = nullptr;
    }
    if (o.save) {
        p.xhat_bf = c->train.xhat[hi]->as<__nv_bfloat16>();
        p.rstd_out = c->train.rstd[hi]->as<float>();
    }
    p.tma_out = dst.tmap_st;
    p.out_tma = c->ws_tma_store;
    // bf16 mode, K <= 512: weight-stationary kernel (the column slice of W stays in shared memory, only A streams: half the L2 traffic)
    if (const char* e = getenv("OSTEO_DDPM_DBG")) p.dbg = atoi(e);
    // ... as a CTA pair (cta_group::2, M = 256 x N = 256) for the K = 512, N >= 512 layers: half the per-SM operand and A-ring traffic, which is what
    // bounds them (0.077 -> 0.066 ms per 100k rows); the K = 256 layers are epilogue-bound and lose with the pair (0.027 -> 0.033 ms)
    if (c->ws2_enable && c->ws_enable && gemm_ws2_eligible(p)) {
        int total_kb = 0;
        for (int sg = 0; sg < p.nseg; ++sg) total_kb += p.seg[sg].nkb;
        if (c->ws2_enable >= 2 || (total_kb >= 7 && p.n_tiles >= 4)) {
            const int rc = launch_gemm_ws2(hb.gw, p, c->sms, s);
            if (rc != -2) return after_launch(c, rc, s);
        }
    }
    if (c->ws_enable && gemm_ws_eligible(p)) {
        const int rc = launch_gemm_ws(hb.gw, p, c->sms, s);
        if (rc != -2) return after_launch(c, rc, s);
    }
    return after_launch(c, launch_gemm(EPI_GN_SILU, hb.gw, p, c->sms, s), s);
}

static void out_proj_common(osteo_ddpm_ctx* c, GemmParams& p, long long row0, long long row1) {
    base_params(c, p);
    const ActBuf& a = *c->acts.back();
    p.tma_a[0] = a.tmap;
    p.tma_a[1] = a.tmap;
    p.tma_b[0] = p.tma_b[1] = c->out_proj.tmap;
    add_segments(p, 0, 0, a.width, 0, c->out_proj.kp, a.width, c->x3());
    set_rows(p, row0, row1);
    p.N = c->D;
    p.n_tiles = c->out_proj.np / BN;
    p.bias = c->out_proj.bias.as<float>();
}

// output_proj with the reverse update as its epilogue (models/diffusion.py:254, :400-423).
static int launch_output_ddpm(osteo_ddpm_ctx* c, long long row0, long long row1, const float* noise, float* eps_out,
                              unsigned long long seed, long long row_base, cudaStream_t s) {
    GemmParams p;
    out_proj_common(c, p, row0, row1);
    p.coef_x = c->coef_x.as<float>();
    p.coef_eps = c->coef_eps.as<float>();
    p.coef_sigma = c->coef_sigma.as<float>();
    p.x = c->x.as<float>();
    p.x_nbox = c->x_nbox;
    p.tma_x_ld = c->x_tmap_ld;
    p.tma_x_st = c->x_tmap_st;
    p.xb = c->xb_ptr();
    p.xb_nbox = c->xb_nbox;
    p.xb_lo_boxes = c->lo_boxes();
    if (p.nseg == 1 && p.seg[0].nkb <= 4) {     // bf16 mode, h0 <= 256: A tile (<= 64 KB) stays resident across an m-block's n-tiles
        p.a_resident = 1;
        p.n_chunks = 2;
    }
    if (const char* e = getenv("OSTEO_DDPM_DBG")) p.dbg = atoi(e);
    if (!(p.dbg & 32)) p.a_resident = 0;      // measured SLOWER than streaming A (2-stage W ring is latency-bound): off unless bit 5 is set
    p.noise = noise;
    p.noise_ld = c->D;
    p.noise_step_stride = noise ? c->noise_step_stride : 0;
    p.noise_t0 = c->noise_t0;
    p.eps_out = eps_out;
    p.eps_ld = c->D;
    p.seed = seed;
    p.row_base = row_base;
    return after_launch(c, launch_gemm(EPI_DDPM, 64, p, c->sms, s), s);
}

// output_proj writing eps as a dense fp32 tensor (forward(return_loss=False)).
static int launch_output_eps(osteo_ddpm_ctx* c, long long row0, long long row1, float* eps_out, cudaStream_t s) {
    GemmParams p;
    out_proj_common(c, p, row0, row1);
    p.out_f32 = eps_out;
    p.out_f32_ld = c->D;
    return after_launch(c, launch_gemm(EPI_LINEAR, 64, p, c->sms, s), s);
}

// Fused tail of a bf16 reverse step: output_proj + reverse update + next step's input_proj (fused_step.cuh).
static int launch_fused(osteo_ddpm_ctx* c, long long row0, long long row1, const float* noise, float* eps_out, unsigned long long seed, long long row_base,
                        cudaStream_t s) {
    static PerDevice dev_state;
    if (!dev_state.configured()) {
        OSTEO_CUDA(cudaFuncSetAttribute(ddpm_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_BYTES));
        OSTEO_CUDA(cudaFuncSetAttribute(ddpm_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_BYTES));
        dev_state.set_configured();
    }
    FusedParams p;
    std::memset(&p, 0, sizeof p);
    p.tma_a = c->acts.back()->tmap;
    p.tma_wout = c->wout_tmap64;
    p.tma_win = c->win_tmap;
    p.M = static_cast<int>(row1);
    p.N = c->D;
    p.m_tile0 = static_cast<int>(row0 / BM);
    p.m_tiles = static_cast<int>((row1 - row0 + BM - 1) / BM);
    p.n_tiles = c->DP / FT;
    p.nkb = c->h0() / BK;
    p.h0 = c->h0();
    p.status = c->status_dev.as<int>();
    p.step = c->step_dev.as<int>() + c->cur_branch;
    p.coef_x = c->coef_x.as<float>();
    p.coef_eps = c->coef_eps.as<float>();
    p.coef_sigma = c->coef_sigma.as<float>();
    p.x = c->x.as<float>();
    p.x_c8 = c->DP / 8;
    p.bias_out = c->out_proj.bias.as<float>();
    if (!c->h_bias_valid) return fail("internal: host copy of output_proj.bias is stale (ensure_host_bias must run before the fused step is enqueued)");
    std::memcpy(p.bias_c, c->h_bias_out.data(), sizeof(float) * static_cast<size_t>(c->DP));
    p.noise = noise;
    p.noise_ld = c->D;
    p.noise_step_stride = noise ? c->noise_step_stride : 0;
    p.noise_t0 = c->noise_t0;
    p.eps_out = eps_out;
    p.eps_ld = c->D;
    p.seed = seed;
    for (int r = 0; r < PHILOX_ROUNDS_REVERSE; ++r) {      // Philox key schedule, hoisted out of the kernel: round r uses key + r * (W0, W1)
        p.rk[2 * r] = static_cast<uint32_t>(seed) + static_cast<uint32_t>(r) * 0x9E3779B9u;
        p.rk[2 * r + 1] = static_cast<uint32_t>(seed >> 32) + static_cast<uint32_t>(r) * 0xBB67AE85u;
    }
    p.one_bits = 0x3f800000u;
    p.row_base = row_base;
    p.bias_in = c->in_proj.bias.as<float>();
    p.time_table = c->time_table.as<float>();
    p.cproj = c->cproj.as<float>();
    p.h0_out = c->acts[0]->ptr();
    p.h0_ld = 2 * c->h0();
    if (const char* e = getenv("OSTEO_DDPM_DBG")) p.dbg = atoi(e);
    p.prefetch = F_PREFETCH;
    if (const char* e = getenv("OSTEO_FUSED_PF")) p.prefetch = atoi(e);
    if (p.m_tiles <= 0) return 0;
    const bool want_trace = getenv("OSTEO_DDPM_TRACE") != nullptr;      // diagnostics: synchronises, prints CTA 0's event timeline
    if (want_trace) {
        if (!c->fused_trace.p) OSTEO_TRY(c->fused_trace.alloc(3 * 32 * 8 * sizeof(long long)));
        OSTEO_CUDA(cudaMemsetAsync(c->fused_trace.p, 0, c->fused_trace.bytes, s));
        p.trace = c->fused_trace.as<long long>();
    }
    const int grid = p.m_tiles < c->sms ? p.m_tiles : c->sms;
    if (eps_out) ddpm_fused_kernel<true><<<grid, F_THREADS, F_SMEM_BYTES, s>>>(p);
    else ddpm_fused_kernel<false><<<grid, F_THREADS, F_SMEM_BYTES, s>>>(p);
    OSTEO_CUDA(cudaGetLastError());
    if (want_trace) {
        static long long h[3 * 32 * 8];
        OSTEO_CUDA(cudaMemcpyAsync(h, c->fused_trace.p, sizeof h, cudaMemcpyDeviceToHost, s));
        OSTEO_CUDA(cudaStreamSynchronize(s));
        long long t0 = h[(2 * 32 + 0) * 8 + 0];
        for (int tile = 0; tile < 24; ++tile) {
            fprintf(stderr, "[trace] tile %2d | MMA eps_rdy %6lld eps_iss %6lld in_rdy %6lld in_iss %6lld | NOISE start %6lld vals %6lld tempty %6lld pub %6lld | "
                            "UPD start %6lld tfull %6lld tmem %6lld xbfe %6lld prefence %6lld pub %6lld\n", tile,
                    h[(0 * 32 + tile) * 8 + 0] - t0, h[(0 * 32 + tile) * 8 + 1] - t0, h[(0 * 32 + tile) * 8 + 2] - t0, h[(0 * 32 + tile) * 8 + 3] - t0,
                    h[(1 * 32 + tile) * 8 + 0] - t0, h[(1 * 32 + tile) * 8 + 1] - t0, h[(1 * 32 + tile) * 8 + 2] - t0, h[(1 * 32 + tile) * 8 + 3] - t0,
                    h[(2 * 32 + tile) * 8 + 0] - t0, h[(2 * 32 + tile) * 8 + 1] - t0, h[(2 * 32 + tile) * 8 + 2] - t0, h[(2 * 32 + tile) * 8 + 3] - t0,
                    h[(2 * 32 + tile) * 8 + 4] - t0, h[(2 * 32 + tile) * 8 + 5] - t0);
        }
    }
    return after_launch(c, 0, s);
}

static long long chunk_of(const osteo_ddpm_ctx* c, long long n) {
    long long ch = c->chunk_rows > 0 ? c->chunk_rows : n;
    ch = round_up(ch, BM);
    return ch < 1 ? BM : ch;
}

// Enqueue one full reverse step over rows [0, n); the timestep is read from the device word.
// Fused path: acts[0] must already hold this step's h0 (ensure_primed); the step leaves the NEXT step's h0 there.
static int enqueue_reverse_step(osteo_ddpm_ctx* c, long long n, const float* noise, float* eps_out, unsigned long long seed, long long row_base, cudaStream_t s,
                                long long rb = 0, long long re = -1) {
    if (c->x_c8 != c->fused_ok())
        return fail("the state was loaded under a different precision / fused setting: call load_state or init_noise again");
    if (re < 0) re = n;
    const long long ch = chunk_of(c, n);
    HalfOpts o;
    for (long long r0 = rb; r0 < re; r0 += ch) {
        const long long r1 = r0 + ch < re ? r0 + ch : re;
        if (!c->x_c8) OSTEO_TRY(launch_input_proj(c, r0, r1, nullptr, s));
        for (size_t i = 0; i < c->halves.size(); ++i) OSTEO_TRY(launch_half(c, static_cast<int>(i), r0, r1, o, s));
        if (c->x_c8) OSTEO_TRY(launch_fused(c, r0, r1, noise, eps_out, seed, row_base, s));
        else OSTEO_TRY(launch_output_ddpm(c, r0, r1, noise, eps_out, seed, row_base, s));
    }
    return 0;
}

// The fused kernel takes output_proj.bias by value: refresh the host copy after a weight change. Synchronises `s` (once per
// set_weights, and only when the fused path is used next); must not be called while `s` is capturing.
static int ensure_host_bias(osteo_ddpm_ctx* c, cudaStream_t s) {
    if (c->h_bias_valid || !c->fused_ok()) return 0;
    c->h_bias_out.assign(F_BIAS_MAX, 0.0f);
    const size_t n = static_cast<size_t>(c->out_proj.np) < static_cast<size_t>(F_BIAS_MAX) ? c->out_proj.np : F_BIAS_MAX;
    OSTEO_CUDA(cudaMemcpyAsync(c->h_bias_out.data(), c->out_proj.bias.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    OSTEO_CUDA(cudaStreamSynchronize(s));
    for (size_t i = static_cast<size_t>(c->D); i < static_cast<size_t>(F_BIAS_MAX); ++i) c->h_bias_out[i] = 0.0f;      // padding columns stay exact zeros
    c->h_bias_valid = true;
    return 0;
}

// Fused path only: make acts[0] = input_proj(x) + b + time_proj[t] + cond_proj for rows [0, n) at the timestep in the device
// step word (== t). The first step after load_state / init_noise contracts the bf16 shadow those kernels wrote; if fused steps
// have run since (the shadow is stale) it is rebuilt from the fp32 state first.
static int ensure_primed(osteo_ddpm_ctx* c, long long n, int t, cudaStream_t s) {
    if (!c->x_c8) return 0;
    if (c->h0_primed && c->h0_t == t && c->h0_n == n) return 0;
    if (!c->shadow_valid) {
        const long long items = n * (c->DP / 4);
        reshadow_kernel<<<grid_for(items, 256, c->sms), 256, 0, s>>>(c->x.as<float>(), c->xs_nbox(), c->xs_shift(), c->DP, c->xb_ptr(), c->xb_nbox, n);
        OSTEO_CUDA(cudaGetLastError());
        ++c->launches;
        c->shadow_valid = true;
    }
    const long long ch = chunk_of(c, n);
    for (long long r0 = 0; r0 < n; r0 += ch) OSTEO_TRY(launch_input_proj(c, r0, r0 + ch < n ? r0 + ch : n, nullptr, s));
    c->h0_primed = true;
    c->h0_t = t;
    c->h0_n = n;
    return 0;
}
// Bookkeeping after `steps` fused steps that started at timestep t.
static void after_steps(osteo_ddpm_ctx* c, int t, int steps) {
    if (!c->x_c8) return;
    c->shadow_valid = false;
    c->h0_t = t - steps;
}

constexpr int GRAPH_UNROLL = 10;
constexpr int MAX_BRANCHES = 4;

// Capture `count` consecutive reverse steps into an executable graph. Rows are independent, so the batch is split into
// `branches` contiguous row ranges (128-row aligned) captured as parallel graph branches: each branch is its own chain of
// count x (10 block GEMMs + fused tail) with its OWN device step word, decremented after each of its steps, and the branches only
// join at the end of the graph. A persistent kernel's tail wave (782 row tiles on 148 SMs = 5.28 waves), its prologue and the
// launch gap of one branch are then filled by the other branch's kernels instead of idling the SMs.
static int capture_steps(osteo_ddpm_ctx* c, long long n, const float* noise, unsigned long long seed, long long row_base, int count, cudaGraphExec_t* out) {
    const long long tiles = (n + BM - 1) / BM;
    int nb = c->branches < 1 ? 1 : (c->branches > MAX_BRANCHES ? MAX_BRANCHES : c->branches);
    if (const char* e = getenv("OSTEO_DDPM_BRANCHES")) nb = atoi(e) < 1 ? 1 : (atoi(e) > MAX_BRANCHES ? MAX_BRANCHES : atoi(e));
    while (nb > 1 && tiles < 2LL * c->sms * nb) --nb;
    c->graph_nb = nb;  // less than two waves per branch only adds launches (measured: 3 / 4 branches of 1.8 / 1.3 waves are 2.5 / 4.5 % slower than 2)
    cudaStream_t cs[MAX_BRANCHES] = {};
    cudaEvent_t fork = nullptr, join[MAX_BRANCHES] = {};
    cudaGraph_t graph = nullptr;
    int rc = 0;
    cudaError_t ce = cudaSuccess;
    for (int b = 0; b < nb && ce == cudaSuccess; ++b) ce = cudaStreamCreateWithFlags(&cs[b], cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
    for (int b = 1; b < nb && ce == cudaSuccess; ++b) ce = cudaEventCreateWithFlags(&join[b], cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaStreamBeginCapture(cs[0], cudaStreamCaptureModeThreadLocal);
    if (ce == cudaSuccess) {
        if (nb > 1) ce = cudaEventRecord(fork, cs[0]);
        for (int b = 1; b < nb && ce == cudaSuccess; ++b) ce = cudaStreamWaitEvent(cs[b], fork, 0);
        for (int b = 0; b < nb && rc == 0 && ce == cudaSuccess; ++b) {
            const long long rb = (tiles * b / nb) * BM, re_ = (tiles * (b + 1) / nb) * BM;
            const long long re = re_ < n ? re_ : n;
            c->cur_branch = b;
            for (int i = 0; i < count && rc == 0; ++i) {
                rc = enqueue_reverse_step(c, n, noise, nullptr, seed, row_base, cs[b], rb, re);
                if (rc == 0) {
                    add_int_kernel<<<1, 1, 0, cs[b]>>>(c->step_dev.as<int>() + b, -1);
                    ++c->launches;
                }
            }
            if (b > 0 && rc == 0) {
                ce = cudaEventRecord(join[b], cs[b]);
                if (ce == cudaSuccess) ce = cudaStreamWaitEvent(cs[0], join[b], 0);
            }
        }
        c->cur_branch = 0;
        if (rc == 0 && ce == cudaSuccess && nb < MAX_BRANCHES) {      // unused step words follow along so that all words stay equal
            add_int_kernel<<<1, MAX_BRANCHES - nb, 0, cs[0]>>>(c->step_dev.as<int>() + nb, -count);
            ++c->launches;
        }
        const cudaError_t ce2 = cudaStreamEndCapture(cs[0], &graph);
        if (ce == cudaSuccess) ce = ce2;
    }
    for (int b = 0; b < nb; ++b) if (cs[b]) cudaStreamDestroy(cs[b]);
    if (fork) cudaEventDestroy(fork);
    for (int b = 1; b < nb; ++b) if (join[b]) cudaEventDestroy(join[b]);
    if (rc != 0) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    if (ce != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        return fail("graph capture failed: %s", cudaGetErrorString(ce));
    }
    ce = cudaGraphInstantiate(out, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return fail("graph instantiate failed: %s", cudaGetErrorString(ce));
    return 0;
}

// See GraphSlot. `enqueue(stream)` must be a pure function of `key` and of device memory contents.
template <class F>
static int run_cached(osteo_ddpm_ctx* c, GraphSlot& slot, const std::vector<unsigned long long>& key, cudaStream_t s, F&& enqueue) {
    if (slot.exec && slot.key == key) {
        OSTEO_CUDA(cudaGraphLaunch(slot.exec, s));
        c->launches += slot.launches;
        return 0;
    }
    if (!slot.exec && !slot.key.empty() && slot.key == key && !c->prof) {
        cudaStream_t cs = nullptr;
        OSTEO_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
        int rc = 0;
        const long long before = c->launches;
        if (ce == cudaSuccess) {
            rc = enqueue(cs);
            ce = cudaStreamEndCapture(cs, &graph);
        }
        slot.launches = c->launches - before;
        c->launches = before;
        cudaStreamDestroy(cs);
        if (rc == 0 && ce == cudaSuccess) ce = cudaGraphInstantiate(&slot.exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (rc != 0 || ce != cudaSuccess) {
            slot.reset();
            cudaGetLastError();
            return rc != 0 ? rc : fail("graph capture failed: %s", cudaGetErrorString(ce));
        }
        OSTEO_CUDA(cudaGraphLaunch(slot.exec, s));
        c->launches += slot.launches;
        return 0;
    }
    slot.reset();
    slot.key = key;
    return enqueue(s);
}

static int require_ready(const osteo_ddpm_ctx* c, long long n) {
    if (!c->have_weights) return fail("weights not set (osteo_ddpm_set_weights)");
    if (!c->have_schedule) return fail("schedule not set (osteo_ddpm_set_schedule)");
    if (!c->have_emb) return fail("time embedding not set (osteo_ddpm_set_time_embedding)");
    if (n <= 0) return fail("row count must be positive");
    if (n > c->cap) return fail("%lld rows exceed the reserved capacity %lld (osteo_ddpm_reserve)", n, c->cap);
    return 0;
}

static int rebuild_time_table(osteo_ddpm_ctx* c, cudaStream_t s) {
    if (!c->have_weights || !c->have_emb) return 0;
    time_table_kernel<<<(c->T + 7) / 8, 256, 8 * c->TD * sizeof(float), s>>>(c->emb_table.as<float>(), c->tp_wt.as<float>(), c->tp_b.as<float>(),
                                                                              c->time_table.as<float>(), c->T, c->TD, c->h0());
    OSTEO_CUDA(cudaGetLastError());
    ++c->launches;
    return 0;
}

}  // namespace osteo

// ====================================================================== C-ABI
extern "C" {

const char* osteo_last_error(void) { return last_error_ref().c_str(); }
int osteo_version(void) { return 100; }
int osteo_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
int osteo_ddpm_num_weight_tensors(int n_hidden) { return 4 + 8 + 8 * (2 * (n_hidden - 1) + 1); }

int osteo_ddpm_create(osteo_ddpm_ctx** out, int device, int data_dim, int cond_dim, int time_dim, int cond_embed_dim, int n_hidden,
                      const int* hidden_dims, int num_steps, float dropout_p, int precision) {
    if (!out) return fail("null output pointer");
    *out = nullptr;
    if (osteo_device_count() <= 0) return fail("no CUDA device: this library has no CPU fallback");
    if (data_dim <= 0 || cond_dim <= 0 || time_dim <= 0 || cond_embed_dim <= 0 || num_steps <= 0 || num_steps > 65535) return fail("invalid dimensions");
    if (n_hidden < 2) return fail("hidden_dims needs at least 2 entries (models/diffusion.py:171-193)");
    for (int i = 0; i < n_hidden; ++i)
        if (hidden_dims[i] % 128 != 0 || hidden_dims[i] > 512 || hidden_dims[i] < 128)
            return fail("hidden dim %d unsupported: the tcgen05 path needs multiples of 128 in [128, 512]", hidden_dims[i]);
    if (time_dim % 2 != 0) return fail("time_dim must be even");
    if (dropout_p < 0.f || dropout_p >= 1.f) return fail("dropout must be in [0, 1)");
    DeviceGuard device_guard(device);      // allocations below land on `device`; the caller's current device is restored on return
    OSTEO_CUDA(cudaSetDevice(device));
    std::unique_ptr<osteo_ddpm_ctx> c(new osteo_ddpm_ctx);
    c->device = device;
    c->sms = sm_count(device);
    if (c->sms <= 0) return fail("cannot query SM count");
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (major != 10) return fail("device compute capability %d.x: kernels are built for sm_100a only", major);
    c->D = data_dim; c->C = cond_dim; c->TD = time_dim; c->E = cond_embed_dim; c->T = num_steps;
    c->hidden.assign(hidden_dims, hidden_dims + n_hidden);
    c->drop_p = dropout_p;
    c->precision = precision;
    c->DP = static_cast<int>(round_up(data_dim, BK));
    const int h0 = c->h0();

    OSTEO_TRY(c->ce_w0.alloc(sizeof(float) * c->E * c->C));
    OSTEO_TRY(c->ce_b0.alloc(sizeof(float) * c->E));
    OSTEO_TRY(c->ce_w2.alloc(sizeof(float) * c->E * c->E));
    OSTEO_TRY(c->ce_b2.alloc(sizeof(float) * c->E));
    OSTEO_TRY(c->cp_w.alloc(sizeof(float) * h0 * c->E));
    OSTEO_TRY(c->ce_w0t.alloc(sizeof(float) * c->E * c->C));
    OSTEO_TRY(c->ce_w2t.alloc(sizeof(float) * c->E * c->E));
    OSTEO_TRY(c->cp_wt.alloc(sizeof(float) * h0 * c->E));
    OSTEO_TRY(c->tp_wt.alloc(sizeof(float) * h0 * c->TD));
    OSTEO_TRY(c->cp_b.alloc(sizeof(float) * h0));
    OSTEO_TRY(c->tp_w.alloc(sizeof(float) * h0 * c->TD));
    OSTEO_TRY(c->tp_b.alloc(sizeof(float) * h0));
    OSTEO_TRY(c->emb_table.alloc(sizeof(float) * c->T * c->TD));
    OSTEO_TRY(c->time_table.alloc(sizeof(float) * c->T * h0));
    for (DevBuf* b : {&c->sqrt_ab, &c->sqrt_1mab, &c->coef_x, &c->coef_eps, &c->coef_sigma}) OSTEO_TRY(b->alloc(sizeof(float) * c->T));
    OSTEO_TRY(c->step_dev.alloc(MAX_BRANCHES * sizeof(int)));
    OSTEO_TRY(c->seed_dev.alloc(sizeof(unsigned long long)));
    OSTEO_TRY(c->status_dev.alloc(sizeof(int)));
    OSTEO_TRY(c->loss_acc.alloc(sizeof(double)));
    OSTEO_CUDA(cudaMemset(c->step_dev.p, 0, MAX_BRANCHES * sizeof(int)));
    OSTEO_CUDA(cudaMemset(c->status_dev.p, 0, sizeof(int)));

    OSTEO_TRY(c->in_proj.init(h0, data_dim));
    OSTEO_TRY(c->out_proj.init(data_dim, h0));
    if (h0 <= 256) {
        OSTEO_TRY(make_tmap_bf16(&c->wout_tmap64, c->out_proj.w.p, c->out_proj.np, 2 * c->out_proj.kp, 2 * c->out_proj.kp, FT));
        OSTEO_TRY(make_tmap_bf16(&c->win_tmap, c->in_proj.w.p, c->in_proj.np, 2 * c->in_proj.kp, 2 * c->in_proj.kp, h0));
    }

    // Block structure of DiffusionUNet.__init__ (models/diffusion.py:171-193). Activation 0 = h0.
    int act_count = 1;
    int block_id = 0;
    auto add_block = [&](int in0_act, int in0_dim, int in1_act, int in1_dim, int out_dim) -> int {
        for (int half = 0; half < 2; ++half) {
            std::unique_ptr<HalfBlock> hb(new HalfBlock);
            const int k = half == 0 ? in0_dim + in1_dim : out_dim;
            OSTEO_TRY(hb->lin.init(out_dim, k));
            OSTEO_TRY(hb->gamma.alloc(sizeof(float) * out_dim));
            OSTEO_TRY(hb->beta.alloc(sizeof(float) * out_dim));
            hb->gw = out_dim / 8;
            hb->src0 = half == 0 ? in0_act : act_count - 1;
            hb->src1 = half == 0 ? in1_act : -1;
            hb->dst = act_count++;
            hb->dropout = half == 0;
            hb->block = block_id;
            c->halves.push_back(std::move(hb));
        }
        ++block_id;
        return 0;
    };
    std::vector<int> skip_act;
    int cur_act = 0, cur_dim = h0;
    for (int i = 1; i < n_hidden; ++i) {
        OSTEO_TRY(add_block(cur_act, cur_dim, -1, 0, c->hidden[i]));
        cur_act = act_count - 1;
        cur_dim = c->hidden[i];
        skip_act.push_back(cur_act);
    }
    OSTEO_TRY(add_block(cur_act, cur_dim, -1, 0, cur_dim));
    cur_act = act_count - 1;
    for (int i = n_hidden - 2; i >= 0; --i) {
        const int sk = skip_act.back();
        skip_act.pop_back();
        const int skip_dim = c->hidden[i + 1];
        OSTEO_TRY(add_block(cur_act, cur_dim, sk, skip_dim, c->hidden[i]));
        cur_act = act_count - 1;
        cur_dim = c->hidden[i];
    }
    if (cur_dim != h0) return fail("internal: decoder does not end at hidden_dims[0]");
    *out = c.release();
    return 0;
}

int osteo_ddpm_destroy(osteo_ddpm_ctx* ctx) {
    if (!ctx) return 0;
    DeviceGuard device_guard(ctx->device);      // garbage collection of a model must not move the thread to another device
    cudaDeviceSynchronize();
    delete ctx;
    return 0;
}

long long osteo_ddpm_capacity(const osteo_ddpm_ctx* ctx) { return ctx ? ctx->cap : 0; }
long long osteo_ddpm_launch_count(const osteo_ddpm_ctx* ctx) { return ctx ? ctx->launches : 0; }

long long osteo_ddpm_workspace_bytes(const osteo_ddpm_ctx* c) {
    if (!c) return 0;
    size_t b = c->x.bytes + c->xb.bytes + c->cproj.bytes + c->in_proj.w.bytes + c->out_proj.w.bytes;
    for (auto& a : c->acts) b += a->buf.bytes;
    for (auto& h : c->halves) b += h->lin.w.bytes + h->lin.wt.bytes;
    b += c->train.bytes();
    return static_cast<long long>(b);
}

int osteo_ddpm_set_chunk_rows(osteo_ddpm_ctx* c, int chunk_rows) {
    OSTEO_CTX(c);
    if (chunk_rows < 0) return fail("chunk_rows must be >= 0");
    c->chunk_rows = chunk_rows;
    return 0;
}
int osteo_ddpm_set_branches(osteo_ddpm_ctx* c, int branches) {
    OSTEO_CTX(c);
    if (branches < 1 || branches > MAX_BRANCHES) return fail("branches must be in [1, %d]", MAX_BRANCHES);
    c->branches = branches;
    return 0;
}
int osteo_ddpm_set_precision(osteo_ddpm_ctx* c, int precision) {
    OSTEO_CTX(c);
    if (precision != OSTEO_PREC_BF16 && precision != OSTEO_PREC_FP32X3) return fail("unknown precision %d", precision);
    c->precision = precision;
    return 0;
}

int osteo_ddpm_set_fused(osteo_ddpm_ctx* c, int enable) {
    OSTEO_CTX(c);
    c->fused_enable = enable ? 1 : 0;
    return 0;
}
int osteo_ddpm_step_is_fused(const osteo_ddpm_ctx* c) { return c && c->fused_ok() ? 1 : 0; }
int osteo_ddpm_graph_branches(const osteo_ddpm_ctx* c) { return c && c->graph_exec ? c->graph_nb : 0; }

int osteo_ddpm_reserve(osteo_ddpm_ctx* c, long long rows) {
    OSTEO_CTX(c);
    if (rows <= c->cap) return 0;
    OSTEO_CUDA(cudaSetDevice(c->device));
    OSTEO_CUDA(cudaDeviceSynchronize());
    if (c->graph_exec) {
        cudaGraphExecDestroy(c->graph_exec);
        c->graph_exec = nullptr;
        if (c->graph_exec_multi) cudaGraphExecDestroy(c->graph_exec_multi);
        c->graph_exec_multi = nullptr;
        c->graph_n = -1;
    }
    const long long cap = round_up(rows, BM);
    const long long mt = cap / BM;
    c->x_nbox = (c->out_proj.np / BN) * (BN / X_BOX_COLS);      // 4 boxes of 32 columns per output tile
    c->xb_nbox = 2 * (c->DP / BK);                               // hi boxes then lo boxes
    OSTEO_TRY(c->x.alloc(static_cast<size_t>(mt) * c->x_nbox * BM * X_BOX_COLS * 4));
    OSTEO_CUDA(cudaMemset(c->x.p, 0, c->x.bytes));
    OSTEO_TRY(make_tmap_f32(&c->x_tmap_ld, c->x.p, static_cast<uint64_t>(mt) * c->x_nbox * BM, X_BOX_COLS, X_BOX_COLS, BM));
    OSTEO_TRY(make_tmap_f32(&c->x_tmap_st, c->x.p, static_cast<uint64_t>(mt) * c->x_nbox * BM, X_BOX_COLS, X_BOX_COLS, 32));
    OSTEO_TRY(c->xb.alloc(static_cast<size_t>(mt) * c->xb_nbox * BM * BK * 2));
    OSTEO_CUDA(cudaMemset(c->xb.p, 0, c->xb.bytes));
    OSTEO_TRY(make_tmap_bf16(&c->xb_tmap, c->xb.p, static_cast<uint64_t>(mt) * c->xb_nbox * BM, BK, BK, BM));
    OSTEO_TRY(c->cproj.alloc(static_cast<size_t>(cap) * c->h0() * 4));
    OSTEO_CUDA(cudaMemset(c->cproj.p, 0, c->cproj.bytes));
    c->acts.clear();
    {
        std::unique_ptr<ActBuf> a(new ActBuf);
        OSTEO_TRY(a->init(cap, c->h0()));
        c->acts.push_back(std::move(a));
    }
    for (auto& hb : c->halves) {
        std::unique_ptr<ActBuf> a(new ActBuf);
        OSTEO_TRY(a->init(cap, hb->lin.n));
        OSTEO_CUDA(cudaMemset(a->buf.p, 0, a->buf.bytes));
        c->acts.push_back(std::move(a));
    }
    c->train.release();
    c->cap = cap;
    ++c->generation;
    c->h0_primed = false;
    c->shadow_valid = false;
    return 0;
}

int osteo_ddpm_set_weights(osteo_ddpm_ctx* c, const float* const* w, int n_tensors, void* stream) {
    OSTEO_CTX(c);
    const int expect = osteo_ddpm_num_weight_tensors(static_cast<int>(c->hidden.size()));
    if (n_tensors != expect) return fail("expected %d weight tensors, got %d", expect, n_tensors);
    for (int i = 0; i < n_tensors; ++i)
        if (!w[i]) return fail("weight tensor %d is null", i);
    OSTEO_CUDA(cudaSetDevice(c->device));
    cudaStream_t s0 = static_cast<cudaStream_t>(stream);
    // After every optimizer step the same 52 tensors are repacked (bf16 [hi|lo] operands, transposes for the backward pass, the
    // time table): ~30 small launches, replayed as one graph from the third call with the same addresses on (see GraphSlot).
    if (!c->aux_fork) {
        for (int i = 0; i < 3; ++i) {
            OSTEO_CUDA(cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking));
            OSTEO_CUDA(cudaEventCreateWithFlags(&c->aux_join[i], cudaEventDisableTiming));
        }
        OSTEO_CUDA(cudaEventCreateWithFlags(&c->aux_fork, cudaEventDisableTiming));
    }
    // The repacks of different layers are independent: they are spread over four lanes (`s` + three library streams forked from it and
    // joined back), which become parallel branches of the captured graph -- ~30 dependent 3 us launches were 0.1 ms of every training step.
    auto enqueue = [&](cudaStream_t s) -> int {
        cudaStream_t lane[4] = {s, c->aux[0], c->aux[1], c->aux[2]};
        OSTEO_CUDA(cudaEventRecord(c->aux_fork, s));
        for (int i = 1; i < 4; ++i) OSTEO_CUDA(cudaStreamWaitEvent(lane[i], c->aux_fork, 0));
        auto copy = [&](DevBuf& dst, const float* src, cudaStream_t q) -> int {
            OSTEO_CUDA(cudaMemcpyAsync(dst.p, src, dst.bytes, cudaMemcpyDeviceToDevice, q));
            return 0;
        };
        // lane 0: the small fp32 layers, their transposes and the time table that depends on them
        OSTEO_TRY(copy(c->ce_w0, w[0], s));
        OSTEO_TRY(copy(c->ce_b0, w[1], s));
        OSTEO_TRY(copy(c->ce_w2, w[2], s));
        OSTEO_TRY(copy(c->ce_b2, w[3], s));
        OSTEO_TRY(copy(c->cp_w, w[6], s));
        OSTEO_TRY(copy(c->cp_b, w[7], s));
        OSTEO_TRY(copy(c->tp_w, w[8], s));
        OSTEO_TRY(copy(c->tp_b, w[9], s));
        transpose_f32_kernel<<<(c->E * c->C + 255) / 256, 256, 0, s>>>(c->ce_w0.as<float>(), c->E, c->C, c->ce_w0t.as<float>());
        transpose_f32_kernel<<<(c->E * c->E + 255) / 256, 256, 0, s>>>(c->ce_w2.as<float>(), c->E, c->E, c->ce_w2t.as<float>());
        transpose_f32_kernel<<<(c->h0() * c->E + 255) / 256, 256, 0, s>>>(c->cp_w.as<float>(), c->h0(), c->E, c->cp_wt.as<float>());
        transpose_f32_kernel<<<(c->h0() * c->TD + 255) / 256, 256, 0, s>>>(c->tp_w.as<float>(), c->h0(), c->TD, c->tp_wt.as<float>());
        OSTEO_CUDA(cudaGetLastError());
        OSTEO_TRY(rebuild_time_table(c, s));
        // lanes 1..3: the packed GEMM operands, the two 5142-wide projections on lanes of their own
        OSTEO_TRY(c->in_proj.upload(w[4], w[5], c->sms, lane[1]));
        int idx = 10, k = 0;
        for (auto& hb : c->halves) {
            cudaStream_t q = lane[3 - (k++ & 1) * 3];      // alternate lane 3 / lane 0
            OSTEO_TRY(hb->lin.upload(w[idx], w[idx + 1], c->sms, q));
            OSTEO_TRY(copy(hb->gamma, w[idx + 2], q));
            OSTEO_TRY(copy(hb->beta, w[idx + 3], q));
            idx += 4;
        }
        OSTEO_TRY(c->out_proj.upload(w[idx], w[idx + 1], c->sms, lane[2]));
        c->launches += 2 + static_cast<long long>(c->halves.size());
        for (int i = 1; i < 4; ++i) {
            OSTEO_CUDA(cudaEventRecord(c->aux_join[i - 1], lane[i]));
            OSTEO_CUDA(cudaStreamWaitEvent(s, c->aux_join[i - 1], 0));
        }
        return 0;
    };
    std::vector<unsigned long long> key{static_cast<unsigned long long>(c->precision), reinterpret_cast<unsigned long long>(c->out_proj.wt.p),
                                        static_cast<unsigned long long>(c->have_emb), c->generation};
    for (int i = 0; i < n_tensors; ++i) key.push_back(reinterpret_cast<unsigned long long>(w[i]));
    const bool had = c->have_weights;
    c->have_weights = true;       // rebuild_time_table (last step of `enqueue`) needs it
    const int rc = run_cached(c, c->weights_graph, key, s0, enqueue);
    if (rc != 0) c->have_weights = had;
    c->h0_primed = false;
    c->h_bias_valid = false;
    ++c->weights_version;
    return rc;
}

int osteo_ddpm_set_schedule(osteo_ddpm_ctx* c, const float* sqrt_ab, const float* sqrt_1mab, const float* coef_x, const float* coef_eps, const float* sigma) {
    OSTEO_CTX(c);
    OSTEO_CUDA(cudaSetDevice(c->device));
    const size_t b = sizeof(float) * c->T;
    OSTEO_CUDA(cudaMemcpy(c->sqrt_ab.p, sqrt_ab, b, cudaMemcpyHostToDevice));
    OSTEO_CUDA(cudaMemcpy(c->sqrt_1mab.p, sqrt_1mab, b, cudaMemcpyHostToDevice));
    OSTEO_CUDA(cudaMemcpy(c->coef_x.p, coef_x, b, cudaMemcpyHostToDevice));
    OSTEO_CUDA(cudaMemcpy(c->coef_eps.p, coef_eps, b, cudaMemcpyHostToDevice));
    OSTEO_CUDA(cudaMemcpy(c->coef_sigma.p, sigma, b, cudaMemcpyHostToDevice));
    c->h_coef_x.assign(coef_x, coef_x + c->T);
    c->h_coef_eps.assign(coef_eps, coef_eps + c->T);
    c->h_coef_sigma.assign(sigma, sigma + c->T);
    c->have_schedule = true;
    return 0;
}

int osteo_ddpm_set_time_embedding(osteo_ddpm_ctx* c, const float* emb_host) {
    OSTEO_CTX(c);
    OSTEO_CUDA(cudaSetDevice(c->device));
    OSTEO_CUDA(cudaMemcpy(c->emb_table.p, emb_host, sizeof(float) * c->T * c->TD, cudaMemcpyHostToDevice));
    c->have_emb = true;
    OSTEO_TRY(rebuild_time_table(c, nullptr));
    OSTEO_CUDA(cudaDeviceSynchronize());
    return 0;
}

int osteo_ddpm_load_state(osteo_ddpm_ctx* c, const float* x_dev, long long n, void* stream) {
    OSTEO_CTX(c);
    if (n <= 0 || n > c->cap) return fail("load_state: %lld rows outside (0, capacity %lld]", n, c->cap);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long items = n * (c->DP / 4);
    c->x_c8 = c->fused_ok();
    load_state_kernel<<<grid_for(items, 256, c->sms), 256, 0, s>>>(x_dev, n, c->D, c->DP, c->x.as<float>(), c->xs_nbox(), c->xs_shift(), c->xb_ptr(), c->xb_nbox,
                                                                  c->lo_boxes());
    OSTEO_CUDA(cudaGetLastError());
    ++c->launches;
    c->shadow_valid = true;
    c->h0_primed = false;
    return 0;
}

int osteo_ddpm_store_state(osteo_ddpm_ctx* c, float* out_dev, long long n, void* stream) {
    OSTEO_CTX(c);
    if (n <= 0 || n > c->cap) return fail("store_state: %lld rows outside (0, capacity %lld]", n, c->cap);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    store_state_kernel<<<grid_for(n * c->D, 256, c->sms), 256, 0, s>>>(c->x.as<float>(), c->xs_nbox(), c->xs_shift(), out_dev, n, c->D);
    OSTEO_CUDA(cudaGetLastError());
    ++c->launches;
    return 0;
}

int osteo_ddpm_store_split(osteo_ddpm_ctx* c, long long n, int mutation_dim, float threshold, uint8_t* calls_dev, uint8_t* call_bits_dev, float* rest_dev, void* stream) {
    OSTEO_CTX(c);
    if (n <= 0 || n > c->cap) return fail("store_split: %lld rows outside (0, capacity %lld]", n, c->cap);
    if (mutation_dim < 0 || mutation_dim > c->D) return fail("store_split: mutation_dim %d outside [0, %d]", mutation_dim, c->D);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long items = n * ((mutation_dim + 7) / 8 + (c->D - mutation_dim));
    store_split_kernel<<<grid_for(items, 256, c->sms), 256, 0, s>>>(c->x.as<float>(), c->xs_nbox(), c->xs_shift(), n, c->D, mutation_dim, threshold, calls_dev, call_bits_dev,
                                                                   rest_dev);
    OSTEO_CUDA(cudaGetLastError());
    ++c->launches;
    return 0;
}

int osteo_ddpm_init_noise(osteo_ddpm_ctx* c, long long n, uint64_t seed, long long row_base, void* stream) {
    OSTEO_CTX(c);
    if (n <= 0 || n > c->cap) return fail("init_noise: %lld rows outside (0, capacity %lld]", n, c->cap);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long items = n * (c->DP / 4);
    c->x_c8 = c->fused_ok();
    init_noise_kernel<<<grid_for(items, 256, c->sms), 256, 0, s>>>(c->x.as<float>(), c->xs_nbox(), c->xs_shift(), c->DP, c->xb_ptr(), c->xb_nbox, c->lo_boxes(), n, c->D,
                                                                  seed, row_base, STREAM_XT, 0u);
    OSTEO_CUDA(cudaGetLastError());
    ++c->launches;
    c->shadow_valid = true;
    c->h0_primed = false;
    return 0;
}

int osteo_ddpm_set_conditions(osteo_ddpm_ctx* c, const float* cond_dev, long long n, void* stream) {
    OSTEO_CTX(c);
    if (!c->have_weights) return fail("weights not set");
    if (n <= 0 || n > c->cap) return fail("set_conditions: %lld rows outside (0, capacity %lld]", n, c->cap);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t smem = sizeof(float) * 16 * (c->C + 2 * c->E);
    cond_path_kernel<<<static_cast<unsigned>((n + 15) / 16), 256, smem, s>>>(cond_dev, n, c->C, c->E, c->h0(), c->ce_w0t.as<float>(), c->ce_b0.as<float>(),
                                                                           c->ce_w2t.as<float>(), c->ce_b2.as<float>(), c->cp_wt.as<float>(), c->cp_b.as<float>(),
                                                                           c->cproj.as<float>(), nullptr, nullptr);
    OSTEO_CUDA(cudaGetLastError());
    ++c->launches;
    c->h0_primed = false;
    return 0;
}

int osteo_ddpm_reverse_step(osteo_ddpm_ctx* c, long long n, int t, const float* noise_dev, float* eps_out_dev, uint64_t seed, long long row_base, void* stream) {
    OSTEO_CTX(c);
    OSTEO_TRY(require_ready(c, n));
    if (t < 0 || t >= c->T) return fail("timestep %d outside [0, %d)", t, c->T);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    set_int_kernel<<<1, MAX_BRANCHES, 0, s>>>(c->step_dev.as<int>(), t);
    OSTEO_CUDA(cudaGetLastError());
    ++c->launches;
    OSTEO_TRY(ensure_host_bias(c, s));
    OSTEO_TRY(ensure_primed(c, n, t, s));
    OSTEO_TRY(enqueue_reverse_step(c, n, noise_dev, eps_out_dev, seed, row_base, s));
    after_steps(c, t, 1);
    return 0;
}

int osteo_ddpm_sample_loop(osteo_ddpm_ctx* c, long long n, int t_start, int t_end, const float* noise_dev, uint64_t seed, long long row_base, int use_graph,
                           void* stream) {
    OSTEO_CTX(c);
    OSTEO_TRY(require_ready(c, n));
    if (t_start >= c->T || t_end < 0 || t_end > t_start) return fail("bad step range [%d, %d]", t_start, t_end);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int steps = t_start - t_end + 1;
    set_int_kernel<<<1, MAX_BRANCHES, 0, s>>>(c->step_dev.as<int>(), t_start);
    OSTEO_CUDA(cudaGetLastError());
    ++c->launches;
    if (c->x_c8 != c->fused_ok()) return fail("the state was loaded under a different precision / fused setting: call load_state or init_noise again");
    OSTEO_TRY(ensure_host_bias(c, s));
    OSTEO_TRY(ensure_primed(c, n, t_start, s));
    // Injected noise is a per-step stack [steps][n, D]: each kernel picks its slice from the device step word (noise_t0 - t), so the
    // replayed graphs below consume it exactly like the in-kernel Philox stream (parity runs of the benchmarked graph path).
    c->noise_step_stride = noise_dev ? n * static_cast<long long>(c->D) : 0;
    c->noise_t0 = t_start;
    if (!use_graph) {
        for (int i = 0; i < steps; ++i) {
            const int rc = enqueue_reverse_step(c, n, noise_dev, nullptr, seed, row_base, s);
            if (rc != 0) {
                c->noise_step_stride = 0;
                return rc;
            }
            add_int_kernel<<<1, MAX_BRANCHES, 0, s>>>(c->step_dev.as<int>(), -1);
            OSTEO_CUDA(cudaGetLastError());
            ++c->launches;
        }
        c->noise_step_stride = 0;
        after_steps(c, t_start, steps);
        return 0;
    }
    // One step (and a block of GRAPH_UNROLL steps) captured once and replayed; the device-resident step word is the only thing that
    // changes between replays. The unrolled graph amortises the per-graph-launch gap over GRAPH_UNROLL steps.
    const bool reuse = c->graph_exec && c->graph_n == n && c->graph_seed == seed && c->graph_row_base == row_base &&
                       c->graph_precision == c->precision && c->graph_chunk == c->chunk_rows && c->graph_fused == (c->x_c8 ? 1 : 0) &&
                       c->graph_branches == c->branches && c->graph_noise == noise_dev && c->graph_noise_t0 == (noise_dev ? t_start : 0) &&
                       c->graph_generation == c->generation && c->graph_weights_version == c->weights_version;
    if (!reuse) {
        for (cudaGraphExec_t* g : {&c->graph_exec, &c->graph_exec_multi}) {
            if (*g) cudaGraphExecDestroy(*g);
            *g = nullptr;
        }
        const long long before = c->launches;
        int rc = capture_steps(c, n, noise_dev, seed, row_base, 1, &c->graph_exec);
        c->graph_launches_per_step = c->launches - before;
        if (rc == 0) rc = capture_steps(c, n, noise_dev, seed, row_base, GRAPH_UNROLL, &c->graph_exec_multi);
        c->launches = before;   // capture enqueued nothing; the replays below are what runs
        if (rc != 0) {
            c->noise_step_stride = 0;
            return rc;
        }
        c->graph_n = n;
        c->graph_seed = seed;
        c->graph_row_base = row_base;
        c->graph_precision = c->precision;
        c->graph_chunk = c->chunk_rows;
        c->graph_branches = c->branches;
        c->graph_fused = c->x_c8 ? 1 : 0;
        c->graph_noise = noise_dev;
        c->graph_noise_t0 = noise_dev ? t_start : 0;
        c->graph_generation = c->generation;
        c->graph_weights_version = c->weights_version;
    }
    c->noise_step_stride = 0;
    int left = steps;
    for (; left >= GRAPH_UNROLL; left -= GRAPH_UNROLL) OSTEO_CUDA(cudaGraphLaunch(c->graph_exec_multi, s));
    for (; left > 0; --left) OSTEO_CUDA(cudaGraphLaunch(c->graph_exec, s));
    c->launches += c->graph_launches_per_step * steps;
    after_steps(c, t_start, steps);
    return 0;
}

int osteo_ddpm_denoise(osteo_ddpm_ctx* c, const float* xt_dev, const int* t_idx_dev, long long n, float* eps_out_dev, void* stream) {
    OSTEO_CTX(c);
    OSTEO_TRY(require_ready(c, n));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long items = n * (c->DP / 4);
    load_state_kernel<<<grid_for(items, 256, c->sms), 256, 0, s>>>(xt_dev, n, c->D, c->DP, nullptr, c->x_nbox, 5, c->xb_ptr(), c->xb_nbox, c->lo_boxes());
    OSTEO_CUDA(cudaGetLastError());
    ++c->launches;
    c->shadow_valid = false;      // the shadow now holds the caller's x_t, and acts[0] its projection
    c->h0_primed = false;
    const long long ch = chunk_of(c, n);
    HalfOpts o;
    for (long long r0 = 0; r0 < n; r0 += ch) {
        const long long r1 = r0 + ch < n ? r0 + ch : n;
        OSTEO_TRY(launch_input_proj(c, r0, r1, t_idx_dev, s));
        for (size_t i = 0; i < c->halves.size(); ++i) OSTEO_TRY(launch_half(c, static_cast<int>(i), r0, r1, o, s));
        OSTEO_TRY(launch_output_eps(c, r0, r1, eps_out_dev, s));
    }
    return 0;
}

int osteo_ddpm_q_sample(osteo_ddpm_ctx* c, const float* x0_dev, const int* t_idx_dev, float* noise_dev, float* xt_dev, long long n, int gen_noise,
                        uint64_t seed, long long row_base, uint32_t salt, void* stream) {
    OSTEO_CTX(c);
    if (!c->have_schedule) return fail("schedule not set");
    if (n <= 0) return fail("row count must be positive");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long items = n * ((c->D + 3) / 4);
    const int grid = grid_for(items, 256, c->sms);
    if (gen_noise)
        q_sample_kernel<true><<<grid, 256, 0, s>>>(x0_dev, t_idx_dev, noise_dev, xt_dev, n, c->D, c->sqrt_ab.as<float>(), c->sqrt_1mab.as<float>(), seed, row_base, salt);
    else
        q_sample_kernel<false><<<grid, 256, 0, s>>>(x0_dev, t_idx_dev, noise_dev, xt_dev, n, c->D, c->sqrt_ab.as<float>(), c->sqrt_1mab.as<float>(), seed, row_base, salt);
    OSTEO_CUDA(cudaGetLastError());
    ++c->launches;
    return 0;
}

int osteo_ddpm_reverse_update(osteo_ddpm_ctx* c, float* x_dev, const float* eps_dev, const float* z_dev, long long n, int t, uint64_t seed, long long row_base,
                              void* stream) {
    OSTEO_CTX(c);
    if (!c->have_schedule) return fail("schedule not set");
    if (t < 0 || t >= c->T) return fail("timestep %d outside [0, %d)", t, c->T);
    if (n <= 0) return fail("row count must be positive");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long items = n * ((c->D + 3) / 4);
    reverse_update_kernel<<<grid_for(items, 256, c->sms), 256, 0, s>>>(x_dev, eps_dev, z_dev, n, c->D, c->h_coef_x[t], c->h_coef_eps[t], c->h_coef_sigma[t], seed,
                                                                      row_base, static_cast<uint32_t>(t));
    OSTEO_CUDA(cudaGetLastError());
    ++c->launches;
    return 0;
}

int osteo_ddpm_profile_step(osteo_ddpm_ctx* c, long long n, int t, uint64_t seed, long long row_base, float* ms_out, int max_out, void* stream) {
    OSTEO_CTX(c);
    OSTEO_TRY(require_ready(c, n));
    if (t < 0 || t >= c->T) return fail("timestep %d outside [0, %d)", t, c->T);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    set_int_kernel<<<1, MAX_BRANCHES, 0, s>>>(c->step_dev.as<int>(), t);
    OSTEO_CUDA(cudaGetLastError());
    OSTEO_TRY(ensure_host_bias(c, s));
    OSTEO_TRY(ensure_primed(c, n, t, s));
    std::vector<cudaEvent_t> ev;
    cudaEvent_t e0;
    OSTEO_CUDA(cudaEventCreate(&e0));
    OSTEO_CUDA(cudaEventRecord(e0, s));
    ev.push_back(e0);
    c->prof = &ev;
    const int rc = enqueue_reverse_step(c, n, nullptr, nullptr, seed, row_base, s);
    c->prof = nullptr;
    if (rc == 0) after_steps(c, t, 1);
    int count = -1;
    if (rc == 0 && cudaStreamSynchronize(s) == cudaSuccess) {
        count = static_cast<int>(ev.size()) - 1;
        for (int i = 0; i < count && i < max_out; ++i) cudaEventElapsedTime(&ms_out[i], ev[i], ev[i + 1]);
    }
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    if (rc != 0) return rc;
    if (count < 0) return fail("profile_step: stream synchronisation failed");
    return count;
}

int osteo_ddpm_status(osteo_ddpm_ctx* c, void* stream) {
    OSTEO_CTX(c);
    int h = 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    OSTEO_CUDA(cudaMemcpyAsync(&h, c->status_dev.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    OSTEO_CUDA(cudaStreamSynchronize(s));
    if (h != 0) return fail("kernel pipeline reported error %d (1 = TMA producer, 2 = MMA issuer, 3 = epilogue wait timed out)", h);
    return 0;
}

}  // extern "C"

#include "api_train.inl"
#include "api_ops.inl"
