// Training step of the C-ABI (included by osteo_ddpm.cu): forward with saved statistics, fused MSE,
// and the full backward (dgrad / wgrad on tcgen05, GroupNorm+SiLU+Dropout backward fused into the dgrad
// epilogues). Replaces forward(return_loss=True) + loss.backward() (models/diffusion.py:344-378,
// utils/train.py:236-239).
namespace osteo {

static int ensure_train_ws(osteo_ddpm_ctx* c) {
    TrainWorkspace& w = c->train;
    if (w.cap == c->cap && w.cap > 0) return 0;
    OSTEO_CUDA(cudaDeviceSynchronize());
    w.release();
    const long long cap = c->cap;
    int max_w = c->D;
    for (auto& hb : c->halves) {
        const int n = hb->lin.n;
        std::unique_ptr<DevBuf> a(new DevBuf), b(new DevBuf), d(new DevBuf);
        OSTEO_TRY(a->alloc(static_cast<size_t>(cap) * 2 * n * 2));
        OSTEO_TRY(b->alloc(static_cast<size_t>(cap) * 8 * 4));
        OSTEO_TRY(d->alloc(static_cast<size_t>(cap) * 2 * n * 2));
        OSTEO_CUDA(cudaMemset(d->p, 0, d->bytes));
        CUtensorMap tm;
        OSTEO_TRY(make_tmap_bf16(&tm, d->p, cap, 2 * n, 2 * n, BM));
        w.xhat.push_back(std::move(a));
        w.rstd.push_back(std::move(b));
        w.dy.push_back(std::move(d));
        w.dy_tmap.push_back(tm);
        if (n > max_w) max_w = n;
    }
    const int h0 = c->h0();
    OSTEO_TRY(w.dh0_bf.alloc(static_cast<size_t>(cap) * 2 * h0 * 2));
    OSTEO_TRY(w.dh0_f32.alloc(static_cast<size_t>(cap) * h0 * 4));
    OSTEO_TRY(w.deps.alloc(static_cast<size_t>(cap) * 2 * c->DP * 2));
    OSTEO_CUDA(cudaMemset(w.deps.p, 0, w.deps.bytes));
    OSTEO_TRY(make_tmap_bf16(&w.deps_tmap, w.deps.p, cap, 2 * c->DP, 2 * c->DP, BM));
    OSTEO_TRY(w.xt_bf.alloc(static_cast<size_t>(cap) * 2 * c->DP * 2));
    OSTEO_CUDA(cudaMemset(w.xt_bf.p, 0, w.xt_bf.bytes));
    OSTEO_TRY(make_tmap_bf16(&w.xt_tmap, w.xt_bf.p, cap, 2 * c->DP, 2 * c->DP, BM));
    OSTEO_TRY(w.noise.alloc(static_cast<size_t>(cap) * c->DP * 4));
    for (DevBuf* b : {&w.pre0, &w.cemb, &w.h1, &w.dcemb, &w.dpre0}) OSTEO_TRY(b->alloc(static_cast<size_t>(cap) * c->E * 4));
    {
        // column partials: [0] MSE epilogue (D columns), [1 + j] GroupNorm backward of half j (3 x width), [H + 1] d(h0)
        std::vector<size_t> cols{static_cast<size_t>(c->D)};
        for (auto& hb : c->halves) cols.push_back(static_cast<size_t>(3) * hb->lin.n);
        cols.push_back(static_cast<size_t>(h0));
        for (size_t pc : cols) {
            std::unique_ptr<DevBuf> b(new DevBuf);
            OSTEO_TRY(b->alloc(static_cast<size_t>(cap / 32) * pc * 4));
            w.partials.push_back(std::move(b));
        }
    }
    OSTEO_TRY(w.temb_bf.alloc(static_cast<size_t>(cap) * 2 * c->TD * 2));
    OSTEO_TRY(w.cemb_bf.alloc(static_cast<size_t>(cap) * 2 * c->E * 2));
    OSTEO_TRY(w.t_copy.alloc(static_cast<size_t>(cap) * sizeof(int)));
    OSTEO_TRY(w.cond_copy.alloc(static_cast<size_t>(cap) * c->C * sizeof(float)));
    OSTEO_TRY(w.loss_tmp.alloc(sizeof(float)));
    c->train_graph.reset();
    c->train_fwd_graph.reset();
    c->train_bwd_graph.reset();
    c->train_bwd_part_graph[0].reset();
    c->train_bwd_part_graph[1].reset();
    for (int i = 0; i < 2; ++i) {
        OSTEO_CUDA(cudaStreamCreateWithFlags(&w.side[i], cudaStreamNonBlocking));
        OSTEO_CUDA(cudaEventCreateWithFlags(&w.ev_join[i], cudaEventDisableTiming));
    }
    OSTEO_CUDA(cudaEventCreateWithFlags(&w.ev_fork, cudaEventDisableTiming));
    (void)max_w;
    w.cap = cap;
    ++c->generation;
    return 0;
}

static int finish_partials(osteo_ddpm_ctx* c, const float* partials, long long n, int nq, int N, float* o0, float* o1, float* o2, cudaStream_t s) {
    const int slabs = static_cast<int>((n + 31) / 32);
    dim3 grid((N + 31) / 32, nq), block(32, 8);
    partials_finish_kernel<<<grid, block, 0, s>>>(partials, slabs, nq, N, o0, o1, o2);
    OSTEO_CUDA(cudaGetLastError());
    ++c->launches;
    return 0;
}

static int outer_accum(osteo_ddpm_ctx* c, const float* G, int gm, const float* X, int xk, const int* idx, long long n, float* dW, cudaStream_t s) {
    const size_t smem = static_cast<size_t>(64) * (gm + xk) * sizeof(float);
    static PerDevice dev_state;
    if (!dev_state.configured()) {
        OSTEO_CUDA(cudaFuncSetAttribute(outer_accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        dev_state.set_configured();
    }
    if (smem > 160 * 1024) return fail("outer_accum: tile does not fit shared memory (gm=%d xk=%d)", gm, xk);
    outer_accum_kernel<<<static_cast<unsigned>((n + 63) / 64), 256, smem, s>>>(G, gm, X, xk, idx, n, dW);
    OSTEO_CUDA(cudaGetLastError());
    ++c->launches;
    return 0;
}

// Append the dgrad K-segments "d(src) += dy_c . W_c[:, col0 : col0 + width]" (B = rows [col0, ...) of W_c^T).
static int add_dgrad_segments(GemmParams& p, int sel, int k, int a_lo_off, int b_lo_off, int b_row0, bool x3) {
    const int nkb = k / BK;
    auto push = [&](int ac, int bc) -> int {
        if (p.nseg >= MAX_KSEG) return fail("too many K segments");
        p.seg[p.nseg++] = KSeg{sel, ac, bc, nkb, sel, b_row0};
        return 0;
    };
    OSTEO_TRY(push(0, 0));
    if (x3) {
        OSTEO_TRY(push(0, b_lo_off));
        OSTEO_TRY(push(a_lo_off, 0));
    }
    return 0;
}

// The kernels of the training step that read the caller's batch (addresses change from batch to batch: never part of a replayed
// graph): q_sample fused with operand packing, and the condition path (cond -> cproj, saved pre-activations).
static int train_pre(osteo_ddpm_ctx* c, const float* x0_dev, const float* cond_dev, long long n, const int* t_idx_dev, const float* noise_dev, unsigned long long seed,
                     long long row_base, cudaStream_t s) {
    TrainWorkspace& w = c->train;
    const int D = c->D, DP = c->DP, h0 = c->h0(), E = c->E;
    {
        // the condition path is independent of q_sample: it runs beside it on a side stream (HBM-bound 98 us next to a 33 us CUDA-core kernel)
        cudaStream_t sc = w.side[1] ? w.side[1] : s;
        if (sc != s) {
            OSTEO_CUDA(cudaEventRecord(w.ev_fork, s));
            OSTEO_CUDA(cudaStreamWaitEvent(sc, w.ev_fork, 0));
        }
        const size_t smem = sizeof(float) * 16 * (c->C + 2 * E);
        cond_path_kernel<<<static_cast<unsigned>((n + 15) / 16), 256, smem, sc>>>(cond_dev, n, c->C, E, h0, c->ce_w0t.as<float>(), c->ce_b0.as<float>(),
                                                                                c->ce_w2t.as<float>(), c->ce_b2.as<float>(), c->cp_wt.as<float>(), c->cp_b.as<float>(),
                                                                                c->cproj.as<float>(), w.pre0.as<float>(), w.cemb.as<float>());
        OSTEO_CUDA(cudaGetLastError());
        train_prepare_kernel<<<grid_for(n * (DP / 4), 256, c->sms), 256, 0, s>>>(x0_dev, noise_dev, t_idx_dev, n, D, c->sqrt_ab.as<float>(), c->sqrt_1mab.as<float>(),
                                                                                 w.noise.as<float>(), DP, w.xt_bf.as<__nv_bfloat16>(), 2 * DP, c->lo(DP), seed, row_base);
        OSTEO_CUDA(cudaGetLastError());
        if (sc != s) {
            OSTEO_CUDA(cudaEventRecord(w.ev_join[1], sc));
            OSTEO_CUDA(cudaStreamWaitEvent(s, w.ev_join[1], 0));
        }
    }
    c->launches += 2;
    c->h0_primed = false;      // acts[0] now belongs to this training batch, not to a sampling state
    return 0;
}

// Shared argument checks of the training entry points.
static int train_check(osteo_ddpm_ctx* c, long long n, const int* t_idx_dev, float* const* grads_dev, int n_tensors, bool want_grads) {
    OSTEO_CTX(c);
    OSTEO_TRY(require_ready(c, n));
    if (!t_idx_dev) return fail("train_step: t_idx_dev is required");
    if (want_grads) {
        const int expect = osteo_ddpm_num_weight_tensors(static_cast<int>(c->hidden.size()));
        if (n_tensors != expect) return fail("train_step: expected %d gradient tensors, got %d", expect, n_tensors);
        if (!c->out_proj.wt.p) return fail("train_step: call osteo_ddpm_enable_training(ctx, 1) and osteo_ddpm_set_weights before requesting gradients");
        for (int i = 0; i < n_tensors; ++i)
            if (!grads_dev[i]) return fail("train_step: gradient tensor %d is null", i);
    }
    OSTEO_CUDA(cudaSetDevice(c->device));
    return 0;
}

// x0hat[r, g] = x0[r, col_g] - k_r (eps_hat - noise)[r, col_g],  k_r = sqrt(1 - ab[t_r]) / sqrt(ab[t_r])  -- the predicted clean sample
// (models/diffusion.py:401-403) at the gathered columns, with (eps_hat - noise) = d(loss)/d(eps_hat) / grad_scale read back from the
// MSE epilogue's output.
__global__ void x0hat_gather_kernel(const float* __restrict__ x0, int D, const int* __restrict__ t_idx, const float* __restrict__ sqrt_ab, const float* __restrict__ sqrt_1mab,
                                    const __nv_bfloat16* __restrict__ deps, int deps_ld, int lo_off, float inv_scale, const int* __restrict__ cols, int G, long long n,
                                    float* __restrict__ out) {
    const long long total = n * G;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / G;
        const int col = cols[i % G];
        const int t = t_idx[r];
        float d = __bfloat162float(deps[r * deps_ld + col]);
        if (lo_off) d += __bfloat162float(deps[r * deps_ld + lo_off + col]);
        const float k = __ldg(sqrt_1mab + t) / __ldg(sqrt_ab + t);
        out[i] = x0[r * D + col] - k * d * inv_scale;
    }
}

// d(loss)/d(eps_hat)[r, col_g] += -k_r g[r, g] (chain rule through x0hat; g = d(aux loss)/d(x0hat)), and the same amounts into the
// 32-row column partials the output_proj bias gradient is reduced from. The columns must be distinct.
__global__ void x0hat_inject_kernel(const int* __restrict__ t_idx, const float* __restrict__ sqrt_ab, const float* __restrict__ sqrt_1mab, __nv_bfloat16* __restrict__ deps,
                                    int deps_ld, int lo_off, float* __restrict__ partials, int D, const int* __restrict__ cols, int G, long long n,
                                    const float* __restrict__ g) {
    const long long total = n * G;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float gv = g[i];
        if (gv == 0.0f) continue;
        const long long r = i / G;
        const int col = cols[i % G];
        const int t = t_idx[r];
        const float add = -(__ldg(sqrt_1mab + t) / __ldg(sqrt_ab + t)) * gv;
        __nv_bfloat16* hi = deps + r * deps_ld + col;
        float v = __bfloat162float(*hi) + add;
        if (lo_off) v += __bfloat162float(hi[lo_off]);
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        *hi = h;
        if (lo_off) hi[lo_off] = __float2bfloat16_rn(v - __bfloat162float(h));
        atomicAdd(partials + (r >> 5) * D + col, add);
    }
}

// Everything of the training step that touches only library-owned buffers plus the gradient tensors: forward from input_proj on,
// MSE, full backward. `t_idx` / `cond` are either the caller's tensors (eager) or the library's copies of them (graph replay: the
// captured kernels must not bake caller addresses that change from batch to batch); seed_dev, when set, overrides `seed` as the
// Philox key of the dropout masks.
static int train_body(osteo_ddpm_ctx* c, const float* cond_dev, long long n, const int* t_idx_dev, const uint8_t* const* drop_masks_dev, int train,
                      unsigned long long seed, const unsigned long long* seed_dev, long long row_base, float* loss_dev, float* const* grads_dev, int n_tensors,
                      cudaStream_t s, int phase = 0, int bwd_part = 0, int cut = 0) {
    // phase 0: forward + backward; 1: forward only, with the statistics the backward needs; 2: backward only (after a phase-1 call)
    // bwd_part (phase 2 only) cuts the backward pass in two launches so that a data-parallel caller can all-reduce the gradients of the
    // first part while the second runs: 1 = loss gradient, output_proj and the half blocks j >= cut (the TAIL of the gradient list),
    // 2 = the half blocks j < cut, input_proj and the embedding paths (the head); 0 = everything
    const bool want_grads = grads_dev != nullptr || phase == 1;
    const int H = static_cast<int>(c->halves.size());
    TrainWorkspace& w = c->train;
    const int D = c->D, DP = c->DP, h0 = c->h0(), E = c->E;
    const bool x3 = c->x3();

    if (phase != 2) {
        // ------------------------------------------------------------------ forward
        zero_double_kernel<<<1, 1, 0, s>>>(c->loss_acc.as<double>());
        OSTEO_CUDA(cudaGetLastError());
        ++c->launches;
        OSTEO_TRY(launch_input_proj(c, 0, n, t_idx_dev, s, &w.xt_tmap));
        for (int i = 0; i < H; ++i) {
            HalfOpts o;
            o.train = train != 0;
            o.seed = seed;
            o.seed_dev = seed_dev;
            o.row_base = row_base;
            o.save = want_grads;
            if (drop_masks_dev && c->halves[i]->dropout) o.drop_mask = drop_masks_dev[c->halves[i]->block];
            OSTEO_TRY(launch_half(c, i, 0, n, o, s));
        }
        {
            GemmParams p;
            out_proj_common(c, p, 0, n);
            p.target = w.noise.as<float>();
            p.target_ld = DP;
            p.grad_scale = static_cast<float>(2.0 / (static_cast<double>(n) * D));
            p.loss_acc = c->loss_acc.as<double>();
            if (want_grads) {
                p.out_bf = w.deps.as<__nv_bfloat16>();
                p.out_bf_ld = 2 * DP;
                p.out_lo_off = c->lo(DP);
                p.col_partials = w.partials[0]->as<float>();
            }
            OSTEO_TRY(after_launch(c, launch_gemm(EPI_MSE, 64, p, c->sms, s), s));
        }
        finish_loss_kernel<<<1, 1, 0, s>>>(c->loss_acc.as<double>(), loss_dev, 1.0 / (static_cast<double>(n) * D));
        OSTEO_CUDA(cudaGetLastError());
        ++c->launches;
    }
    if (!want_grads || phase == 1) return 0;

    // ------------------------------------------------------------------ backward
    auto numel = [&](int i) -> size_t {
        // element counts in state_dict order (see osteo_ddpm_set_weights)
        if (i == 0) return static_cast<size_t>(E) * c->C;
        if (i == 1 || i == 3) return E;
        if (i == 2) return static_cast<size_t>(E) * E;
        if (i == 4) return static_cast<size_t>(h0) * D;
        if (i == 5 || i == 7 || i == 9) return h0;
        if (i == 6) return static_cast<size_t>(h0) * E;
        if (i == 8) return static_cast<size_t>(h0) * c->TD;
        const int j = i - 10;
        if (j < 4 * H) {
            const HalfBlock& hb = *c->halves[j / 4];
            return (j % 4 == 0) ? static_cast<size_t>(hb.lin.n) * hb.lin.k : static_cast<size_t>(hb.lin.n);
        }
        return (j - 4 * H == 0) ? static_cast<size_t>(D) * h0 : static_cast<size_t>(D);
    };
    if (bwd_part != 2) {
        // weight gradients accumulate atomically (split-batch wgrad): zero them; contiguous tensors share one memset
        int i = 0;
        while (i < n_tensors) {
            int j = i;
            size_t total = numel(i);
            while (j + 1 < n_tensors && grads_dev[j + 1] == grads_dev[j] + numel(j)) total += numel(++j);
            OSTEO_CUDA(cudaMemsetAsync(grads_dev[i], 0, total * sizeof(float), s));
            i = j + 1;
        }
    }
    const int gi_out_w = 10 + 4 * H, gi_out_b = gi_out_w + 1;

    // The dgrad chain (one GEMM per activation, each needing the previous one's output) is the critical path of the backward pass;
    // the weight / bias gradients only consume its outputs. They run on two side streams that fork from `s` after the producing
    // kernel and join it again before this call returns, so at training batch sizes (64 row tiles per GEMM: less than one wave)
    // they fill the SMs the chain leaves idle instead of lengthening it. OSTEO_TRAIN_OVERLAP=0 keeps everything on `s`.
    static const bool overlap = !(getenv("OSTEO_TRAIN_OVERLAP") && atoi(getenv("OSTEO_TRAIN_OVERLAP")) == 0);
    cudaStream_t s1 = overlap ? w.side[0] : s, s2 = overlap ? w.side[1] : s;
    auto fork_to = [&](cudaStream_t side) -> int {      // work enqueued on `side` from here on starts after everything enqueued on `s` so far
        if (side == s) return 0;
        OSTEO_CUDA(cudaEventRecord(w.ev_fork, s));
        OSTEO_CUDA(cudaStreamWaitEvent(side, w.ev_fork, 0));
        return 0;
    };

    // output_proj: bias gradient from the MSE epilogue's column partials, weight gradient = deps^T . act_last
    if (bwd_part != 2) {
    OSTEO_TRY(fork_to(s1));
    OSTEO_TRY(finish_partials(c, w.partials[0]->as<float>(), n, 1, D, grads_dev[gi_out_b], nullptr, nullptr, s1));
    {
        const ActBuf& a = *c->acts.back();
        OSTEO_TRY(after_launch(c, launch_wgrad(w.deps.as<__nv_bfloat16>(), 2 * DP, DP, D, a.ptr(), 2 * a.width, 0, a.width, a.width, grads_dev[gi_out_w], a.width, n, x3,
                                               c->status_dev.as<int>(), c->sms, s1), s1));
    }
    }

    // consumers of an activation: (half index, first input column of that activation inside the consumer's Linear)
    auto consumers_of = [&](int act, std::vector<std::pair<int, int>>& out) {
        out.clear();
        for (int ci = 0; ci < H; ++ci) {
            const HalfBlock& hb = *c->halves[ci];
            if (hb.src0 == act) out.push_back({ci, 0});
            if (hb.src1 == act) out.push_back({ci, c->acts[hb.src0]->width});
        }
    };
    std::vector<std::pair<int, int>> cons;
    const int j_hi = bwd_part == 2 ? cut - 1 : H - 1, j_lo = bwd_part == 1 ? cut : -1;
    for (int j = j_hi; j >= j_lo; --j) {
        // d(activation j+1) summed over its consumers, then this half's GroupNorm/SiLU/Dropout backward (j >= 0)
        GemmParams p;
        base_params(c, p);
        const int act = j + 1;
        const int width = c->acts[act]->width;
        consumers_of(act, cons);
        int sel = 0;
        if (j == H - 1) {
            p.tma_a[0] = p.tma_a[1] = w.deps_tmap;
            p.tma_b[0] = p.tma_b[1] = c->out_proj.tmap_t;
            OSTEO_TRY(add_dgrad_segments(p, 0, DP, DP, c->out_proj.np, 0, x3));
            sel = 1;
        }
        for (auto& cn : cons) {
            if (sel >= 2) return fail("internal: activation %d has more than two consumers", act);
            const HalfBlock& hb = *c->halves[cn.first];
            p.tma_a[sel] = w.dy_tmap[cn.first];
            p.tma_b[sel] = hb.lin.tmap_t;
            if (sel == 0) {
                p.tma_a[1] = p.tma_a[0];
                p.tma_b[1] = p.tma_b[0];
            }
            OSTEO_TRY(add_dgrad_segments(p, sel, hb.lin.n, hb.lin.n, hb.lin.np, cn.second, x3));
            ++sel;
        }
        if (p.nseg == 0) return fail("internal: activation %d has no consumer", act);
        set_rows(p, 0, n);
        p.N = width;
        p.n_tiles = width / BN;
        p.col_partials = w.partials[j >= 0 ? 1 + j : H + 1]->as<float>();
        cudaStream_t sj = (j & 1) ? s2 : s1;      // alternate: two consecutive halves' weight gradients can overlap each other too
        if (j >= 0) {
            HalfBlock& hb = *c->halves[j];
            p.gamma = hb.gamma.as<float>();
            p.beta = hb.beta.as<float>();
            p.xhat_in = w.xhat[j]->as<__nv_bfloat16>();
            p.xhat_ld = 2 * width;
            p.xhat_lo_off = c->lo(width);
            p.rstd_in = w.rstd[j]->as<float>();
            p.out_bf = w.dy[j]->as<__nv_bfloat16>();
            p.out_bf_ld = 2 * width;
            p.out_lo_off = c->lo(width);
            if (train && hb.dropout && c->drop_p > 0.f) {
                p.drop_p = c->drop_p;
                p.drop_mask = drop_masks_dev ? drop_masks_dev[hb.block] : nullptr;
                p.drop_stream = STREAM_DROPOUT + static_cast<uint32_t>(hb.block);
                p.seed = seed;
                p.seed_dev = seed_dev;
                p.row_base = row_base;
            }
            OSTEO_TRY(after_launch(c, launch_gemm(EPI_GN_BWD, hb.gw, p, c->sms, s), s));
            const int gi = 10 + 4 * j;
            OSTEO_TRY(fork_to(sj));
            OSTEO_TRY(finish_partials(c, p.col_partials, n, 3, width, grads_dev[gi + 2], grads_dev[gi + 3], grads_dev[gi + 1], sj));
            // weight gradient of this half: one launch per concatenated source
            const ActBuf& a0 = *c->acts[hb.src0];
            OSTEO_TRY(after_launch(c, launch_wgrad(w.dy[j]->as<__nv_bfloat16>(), 2 * width, width, width, a0.ptr(), 2 * a0.width, 0, a0.width, a0.width, grads_dev[gi],
                                                   hb.lin.k, n, x3, c->status_dev.as<int>(), c->sms, sj), sj));
            if (hb.src1 >= 0) {
                const ActBuf& a1 = *c->acts[hb.src1];
                OSTEO_TRY(after_launch(c, launch_wgrad(w.dy[j]->as<__nv_bfloat16>(), 2 * width, width, width, a1.ptr(), 2 * a1.width, 0, a1.width, a1.width,
                                                       grads_dev[gi] + a0.width, hb.lin.k, n, x3, c->status_dev.as<int>(), c->sms, sj), sj));
            }
        } else {
            // d(h0): plain epilogue, kept as bf16 (wgrad operand) and fp32 (embedding paths)
            p.out_bf = w.dh0_bf.as<__nv_bfloat16>();
            p.out_bf_ld = 2 * h0;
            p.out_lo_off = c->lo(h0);
            p.out_f32 = w.dh0_f32.as<float>();
            p.out_f32_ld = h0;
            p.step = nullptr;
            OSTEO_TRY(after_launch(c, launch_gemm(EPI_LINEAR, 64, p, c->sms, s), s));
        }
    }
    if (bwd_part != 1) {
    // Everything below needs only d(h0). Side stream 1: input_proj weight gradient (dh0^T . x_t, the largest wgrad); side stream 2: the
    // three bias gradients that equal colsum(d(h0)) and the time_proj / cond_proj weight gradients; `s`: the ConditionalEmbedding chain.
    OSTEO_TRY(fork_to(s1));
    OSTEO_TRY(fork_to(s2));
    OSTEO_TRY(after_launch(c, launch_wgrad(w.dh0_bf.as<__nv_bfloat16>(), 2 * h0, h0, h0, w.xt_bf.as<__nv_bfloat16>(), 2 * DP, 0, DP, D, grads_dev[4], D, n, x3, c->status_dev.as<int>(),
                                           c->sms, s1), s1));
    OSTEO_TRY(finish_partials(c, w.partials[H + 1]->as<float>(), n, 1, h0, grads_dev[5], nullptr, nullptr, s2));
    OSTEO_CUDA(cudaMemcpyAsync(grads_dev[7], grads_dev[5], sizeof(float) * h0, cudaMemcpyDeviceToDevice, s2));
    OSTEO_CUDA(cudaMemcpyAsync(grads_dev[9], grads_dev[5], sizeof(float) * h0, cudaMemcpyDeviceToDevice, s2));
    // time_proj / cond_proj weight gradients = dh0^T . emb[t] and dh0^T . cemb: [h0 x 128] and [h0 x 64] contractions over the batch on the
    // tensor cores (as CUDA-core outer products they were the longest kernels of the tail: 108 + 52 us at batch 8192)
    if ((c->TD & 63) == 0 && (E & 63) == 0) {
        const int TD = c->TD;
        pack_rows_hilo_kernel<<<grid_for(n * (TD / 4), 256, c->sms), 256, 0, s2>>>(c->emb_table.as<float>(), TD, t_idx_dev, n, TD, w.temb_bf.as<__nv_bfloat16>(), 2 * TD, c->lo(TD));
        pack_rows_hilo_kernel<<<grid_for(n * (E / 4), 256, c->sms), 256, 0, s2>>>(w.cemb.as<float>(), E, nullptr, n, E, w.cemb_bf.as<__nv_bfloat16>(), 2 * E, c->lo(E));
        OSTEO_CUDA(cudaGetLastError());
        c->launches += 2;
        OSTEO_TRY(after_launch(c, launch_wgrad(w.dh0_bf.as<__nv_bfloat16>(), 2 * h0, h0, h0, w.temb_bf.as<__nv_bfloat16>(), 2 * TD, 0, TD, TD, grads_dev[8], TD, n, x3,
                                               c->status_dev.as<int>(), c->sms, s2), s2));
        OSTEO_TRY(after_launch(c, launch_wgrad(w.dh0_bf.as<__nv_bfloat16>(), 2 * h0, h0, h0, w.cemb_bf.as<__nv_bfloat16>(), 2 * E, 0, E, E, grads_dev[6], E, n, x3,
                                               c->status_dev.as<int>(), c->sms, s2), s2));
    } else {
        OSTEO_TRY(outer_accum(c, w.dh0_f32.as<float>(), h0, c->emb_table.as<float>(), c->TD, t_idx_dev, n, grads_dev[8], s2));
        OSTEO_TRY(outer_accum(c, w.dh0_f32.as<float>(), h0, w.cemb.as<float>(), E, nullptr, n, grads_dev[6], s2));
    }
    {
        const size_t smem = sizeof(float) * 8 * (h0 + E);
        // Wc^T view: cond_bwd needs Wc as [h0, E] row-major, which is exactly cond_proj.weight's layout.
        cond_bwd_rows_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, smem, s>>>(w.dh0_f32.as<float>(), n, h0, E, c->cp_w.as<float>(), c->ce_w2.as<float>(),
                                                                                 w.pre0.as<float>(), w.dcemb.as<float>(), w.h1.as<float>(), w.dpre0.as<float>());
        OSTEO_CUDA(cudaGetLastError());
        ++c->launches;
    }
    {
        // 16 rows per block: a block's loop is a serial chain of dependent loads (64 blocks of 128 rows took 54 us for 2 MB)
        long long gy = (n + 15) / 16;
        gy = gy < 1 ? 1 : (gy > 4096 ? 4096 : gy);
        dim3 grid((E + 63) / 64, static_cast<unsigned>(gy));
        colsum_f32_kernel<<<grid, 64, 0, s>>>(w.dcemb.as<float>(), n, E, grads_dev[3]);
        colsum_f32_kernel<<<grid, 64, 0, s>>>(w.dpre0.as<float>(), n, E, grads_dev[1]);
        OSTEO_CUDA(cudaGetLastError());
        c->launches += 2;
    }
    OSTEO_TRY(outer_accum(c, w.dcemb.as<float>(), E, w.h1.as<float>(), E, nullptr, n, grads_dev[2], s));
    OSTEO_TRY(outer_accum(c, w.dpre0.as<float>(), E, cond_dev, c->C, nullptr, n, grads_dev[0], s));
    }
    for (int i = 0; i < 2 && overlap; ++i) {
        OSTEO_CUDA(cudaEventRecord(w.ev_join[i], w.side[i]));
        OSTEO_CUDA(cudaStreamWaitEvent(s, w.ev_join[i], 0));
    }
    return 0;
}


}  // namespace osteo

extern "C" {

int osteo_ddpm_enable_training(osteo_ddpm_ctx* c, int enable) {
    OSTEO_CTX(c);
    OSTEO_CUDA(cudaSetDevice(c->device));
    if (!enable) return 0;
    if (c->out_proj.wt.p) return 0;
    OSTEO_CUDA(cudaDeviceSynchronize());
    for (auto& hb : c->halves) OSTEO_TRY(hb->lin.init_transposed());
    OSTEO_TRY(c->out_proj.init_transposed());
    ++c->generation;
    c->have_weights = false;   // W^T copies are filled by the next osteo_ddpm_set_weights
    return 0;
}

int osteo_ddpm_set_train_graph(osteo_ddpm_ctx* c, int enable) {
    OSTEO_CTX(c);
    c->train_graph_enable = enable ? 1 : 0;
    if (!enable) {
        c->train_graph.reset();
        c->train_fwd_graph.reset();
        c->train_bwd_graph.reset();
        c->train_bwd_part_graph[0].reset();
        c->train_bwd_part_graph[1].reset();
    c->train_bwd_part_graph[0].reset();
    c->train_bwd_part_graph[1].reset();
    }
    return 0;
}

int osteo_ddpm_train_step(osteo_ddpm_ctx* c, const float* x0_dev, const float* cond_dev, long long n, const int* t_idx_dev, const float* noise_dev,
                          const uint8_t* const* drop_masks_dev, int train, uint64_t seed, long long row_base, float* loss_dev, float* const* grads_dev,
                          int n_tensors, void* stream) {
    if (!loss_dev) return fail("train_step: loss_dev is required");
    const bool want_grads = grads_dev != nullptr;
    OSTEO_TRY(train_check(c, n, t_idx_dev, grads_dev, n_tensors, want_grads));
    OSTEO_TRY(ensure_train_ws(c));
    TrainWorkspace& w = c->train;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    w.fwd_n = want_grads ? n : -1;

    OSTEO_TRY(train_pre(c, x0_dev, cond_dev, n, t_idx_dev, noise_dev, seed, row_base, s));

    // ---- everything else: forward from input_proj on, MSE, backward. Replayed as ONE graph launch when nothing is injected: at
    // batch 8192 the ~120 launches of a step cost more host time than device time (scripts/train_probe.py).
    const bool graphable = want_grads && c->train_graph_enable && !noise_dev && !drop_masks_dev && !c->prof;
    if (!graphable) return train_body(c, cond_dev, n, t_idx_dev, drop_masks_dev, train, seed, nullptr, row_base, loss_dev, grads_dev, n_tensors, s);
    OSTEO_CUDA(cudaMemcpyAsync(w.t_copy.p, t_idx_dev, sizeof(int) * n, cudaMemcpyDeviceToDevice, s));
    OSTEO_CUDA(cudaMemcpyAsync(w.cond_copy.p, cond_dev, sizeof(float) * n * c->C, cudaMemcpyDeviceToDevice, s));
    set_u64_kernel<<<1, 1, 0, s>>>(c->seed_dev.as<unsigned long long>(), seed);
    OSTEO_CUDA(cudaGetLastError());
    c->launches += 3;
    std::vector<unsigned long long> key{static_cast<unsigned long long>(n), static_cast<unsigned long long>(train != 0), static_cast<unsigned long long>(row_base),
                                        static_cast<unsigned long long>(c->precision), static_cast<unsigned long long>(c->ws_enable),
                                        reinterpret_cast<unsigned long long>(c->acts[0]->ptr()), reinterpret_cast<unsigned long long>(w.deps.p),
                                        reinterpret_cast<unsigned long long>(c->out_proj.wt.p), c->generation};
    for (int i = 0; i < n_tensors; ++i) key.push_back(reinterpret_cast<unsigned long long>(grads_dev[i]));
    OSTEO_TRY(run_cached(c, c->train_graph, key, s, [&](cudaStream_t q) {
        return train_body(c, w.cond_copy.as<float>(), n, w.t_copy.as<int>(), nullptr, train, seed, c->seed_dev.as<unsigned long long>(), row_base, w.loss_tmp.as<float>(),
                          grads_dev, n_tensors, q);
    }));
    OSTEO_CUDA(cudaMemcpyAsync(loss_dev, w.loss_tmp.p, sizeof(float), cudaMemcpyDeviceToDevice, s));
    return 0;
}

// ---- the training step in two halves, for auxiliary losses on the predicted clean sample: forward, then the
// caller reads x0hat at the columns its losses need, injects d(aux)/d(x0hat), and runs the backward pass.
int osteo_ddpm_train_forward(osteo_ddpm_ctx* c, const float* x0_dev, const float* cond_dev, long long n, const int* t_idx_dev, const float* noise_dev,
                             const uint8_t* const* drop_masks_dev, int train, uint64_t seed, long long row_base, float* loss_dev, void* stream) {
    if (!loss_dev) return fail("train_forward: loss_dev is required");
    OSTEO_TRY(train_check(c, n, t_idx_dev, nullptr, 0, false));
    if (!c->out_proj.wt.p) return fail("train_forward: call osteo_ddpm_enable_training(ctx, 1) and osteo_ddpm_set_weights first");
    OSTEO_TRY(ensure_train_ws(c));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TrainWorkspace& w = c->train;
    OSTEO_TRY(train_pre(c, x0_dev, cond_dev, n, t_idx_dev, noise_dev, seed, row_base, s));
    // graph replay as in osteo_ddpm_train_step, one cached graph per half
    w.fwd_graphed = c->train_graph_enable && !noise_dev && !drop_masks_dev && !c->prof;
    if (!w.fwd_graphed) {
        OSTEO_TRY(train_body(c, cond_dev, n, t_idx_dev, drop_masks_dev, train, seed, nullptr, row_base, loss_dev, nullptr, 0, s, /*phase=*/1));
    } else {
        OSTEO_CUDA(cudaMemcpyAsync(w.t_copy.p, t_idx_dev, sizeof(int) * n, cudaMemcpyDeviceToDevice, s));
        OSTEO_CUDA(cudaMemcpyAsync(w.cond_copy.p, cond_dev, sizeof(float) * n * c->C, cudaMemcpyDeviceToDevice, s));
        set_u64_kernel<<<1, 1, 0, s>>>(c->seed_dev.as<unsigned long long>(), seed);
        OSTEO_CUDA(cudaGetLastError());
        c->launches += 3;
        std::vector<unsigned long long> key{static_cast<unsigned long long>(n), static_cast<unsigned long long>(train != 0), static_cast<unsigned long long>(row_base),
                                            static_cast<unsigned long long>(c->precision), static_cast<unsigned long long>(c->ws_enable),
                                            reinterpret_cast<unsigned long long>(c->acts[0]->ptr()), reinterpret_cast<unsigned long long>(w.deps.p),
                                            reinterpret_cast<unsigned long long>(c->out_proj.wt.p), c->generation};
        OSTEO_TRY(run_cached(c, c->train_fwd_graph, key, s, [&](cudaStream_t q) {
            return train_body(c, w.cond_copy.as<float>(), n, w.t_copy.as<int>(), nullptr, train, seed, c->seed_dev.as<unsigned long long>(), row_base,
                              w.loss_tmp.as<float>(), nullptr, 0, q, /*phase=*/1);
        }));
        OSTEO_CUDA(cudaMemcpyAsync(loss_dev, w.loss_tmp.p, sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    w.fwd_n = n;
    return 0;
}

int osteo_ddpm_train_x0hat(osteo_ddpm_ctx* c, const float* x0_dev, const int* t_idx_dev, long long n, const int* cols_dev, int n_cols, float* out_dev, void* stream) {
    OSTEO_CTX(c);
    TrainWorkspace& w = c->train;
    if (w.fwd_n != n || n <= 0) return fail("train_x0hat: no forward pass of %lld rows is pending (osteo_ddpm_train_forward)", n);
    if (n_cols <= 0) return fail("train_x0hat: n_cols must be positive");
    const float inv_scale = static_cast<float>(static_cast<double>(n) * c->D / 2.0);
    x0hat_gather_kernel<<<grid_for(n * n_cols, 256, c->sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(x0_dev, c->D, t_idx_dev, c->sqrt_ab.as<float>(), c->sqrt_1mab.as<float>(),
                                                                                                        w.deps.as<__nv_bfloat16>(), 2 * c->DP, c->lo(c->DP), inv_scale, cols_dev,
                                                                                                        n_cols, n, out_dev);
    OSTEO_CUDA(cudaGetLastError());
    ++c->launches;
    return 0;
}

int osteo_ddpm_train_inject(osteo_ddpm_ctx* c, const int* t_idx_dev, long long n, const int* cols_dev, int n_cols, const float* g_dev, void* stream) {
    OSTEO_CTX(c);
    TrainWorkspace& w = c->train;
    if (w.fwd_n != n || n <= 0) return fail("train_inject: no forward pass of %lld rows is pending (osteo_ddpm_train_forward)", n);
    if (n_cols <= 0) return fail("train_inject: n_cols must be positive");
    x0hat_inject_kernel<<<grid_for(n * n_cols, 256, c->sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(t_idx_dev, c->sqrt_ab.as<float>(), c->sqrt_1mab.as<float>(),
                                                                                                        w.deps.as<__nv_bfloat16>(), 2 * c->DP, c->lo(c->DP),
                                                                                                        w.partials[0]->as<float>(), c->D, cols_dev, n_cols, n, g_dev);
    OSTEO_CUDA(cudaGetLastError());
    ++c->launches;
    return 0;
}

int osteo_ddpm_train_backward(osteo_ddpm_ctx* c, const float* cond_dev, long long n, const int* t_idx_dev, const uint8_t* const* drop_masks_dev, int train, uint64_t seed,
                              long long row_base, float* const* grads_dev, int n_tensors, void* stream) {
    if (!grads_dev) return fail("train_backward: grads_dev is required");
    OSTEO_TRY(train_check(c, n, t_idx_dev, grads_dev, n_tensors, true));
    if (c->train.fwd_n != n) return fail("train_backward: no forward pass of %lld rows is pending (osteo_ddpm_train_forward)", n);
    TrainWorkspace& w = c->train;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!w.fwd_graphed || drop_masks_dev) {
        OSTEO_TRY(train_body(c, cond_dev, n, t_idx_dev, drop_masks_dev, train, seed, nullptr, row_base, nullptr, grads_dev, n_tensors, s, /*phase=*/2));
    } else {
        // t / cond copies and the seed word were written by the matching train_forward
        std::vector<unsigned long long> key{static_cast<unsigned long long>(n), static_cast<unsigned long long>(train != 0), static_cast<unsigned long long>(row_base),
                                            static_cast<unsigned long long>(c->precision), static_cast<unsigned long long>(c->ws_enable),
                                            reinterpret_cast<unsigned long long>(c->acts[0]->ptr()), reinterpret_cast<unsigned long long>(w.deps.p),
                                            reinterpret_cast<unsigned long long>(c->out_proj.wt.p), c->generation};
        for (int i = 0; i < n_tensors; ++i) key.push_back(reinterpret_cast<unsigned long long>(grads_dev[i]));
        OSTEO_TRY(run_cached(c, c->train_bwd_graph, key, s, [&](cudaStream_t q) {
            return train_body(c, w.cond_copy.as<float>(), n, w.t_copy.as<int>(), nullptr, train, seed, c->seed_dev.as<unsigned long long>(), row_base, nullptr,
                              grads_dev, n_tensors, q, /*phase=*/2);
        }));
    }
    w.fwd_n = -1;
    return 0;
}


// The backward pass in two launches (after osteo_ddpm_train_forward): part 1 = loss gradient, output_proj and the half blocks j >= cut,
// i.e. the gradient tensors [10 + 4 cut, n_tensors) -- the TAIL of the list; part 2 = the rest. A data-parallel caller starts the
// all-reduce of the tail while part 2 runs (utils/train.py:236-244 has one backward() and no overlap). Parts 1 and 2 together enqueue
// exactly the kernels of osteo_ddpm_train_backward, in the same order: the gradients are the same bits.
int osteo_ddpm_train_backward_part(osteo_ddpm_ctx* c, const float* cond_dev, long long n, const int* t_idx_dev, const uint8_t* const* drop_masks_dev, int train,
                                   uint64_t seed, long long row_base, float* const* grads_dev, int n_tensors, int part, int cut, void* stream) {
    if (!grads_dev) return fail("train_backward_part: grads_dev is required");
    OSTEO_TRY(train_check(c, n, t_idx_dev, grads_dev, n_tensors, true));
    if (part != 1 && part != 2) return fail("train_backward_part: part %d outside {1, 2}", part);
    if (cut < 0 || cut > static_cast<int>(c->halves.size())) return fail("train_backward_part: cut %d outside [0, %d]", cut, static_cast<int>(c->halves.size()));
    if (c->train.fwd_n != n) return fail("train_backward_part: no forward pass of %lld rows is pending (osteo_ddpm_train_forward)", n);
    TrainWorkspace& w = c->train;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!w.fwd_graphed || drop_masks_dev) {
        OSTEO_TRY(train_body(c, cond_dev, n, t_idx_dev, drop_masks_dev, train, seed, nullptr, row_base, nullptr, grads_dev, n_tensors, s, /*phase=*/2, part, cut));
    } else {
        std::vector<unsigned long long> key{static_cast<unsigned long long>(n), static_cast<unsigned long long>(train != 0), static_cast<unsigned long long>(row_base),
                                            static_cast<unsigned long long>(c->precision), static_cast<unsigned long long>(c->ws_enable),
                                            reinterpret_cast<unsigned long long>(c->acts[0]->ptr()), reinterpret_cast<unsigned long long>(w.deps.p),
                                            reinterpret_cast<unsigned long long>(c->out_proj.wt.p), c->generation, static_cast<unsigned long long>(cut)};
        for (int i = 0; i < n_tensors; ++i) key.push_back(reinterpret_cast<unsigned long long>(grads_dev[i]));
        OSTEO_TRY(run_cached(c, c->train_bwd_part_graph[part - 1], key, s, [&](cudaStream_t q) {
            return train_body(c, w.cond_copy.as<float>(), n, w.t_copy.as<int>(), nullptr, train, seed, c->seed_dev.as<unsigned long long>(), row_base, nullptr,
                              grads_dev, n_tensors, q, /*phase=*/2, part, cut);
        }));
    }
    if (part == 2) w.fwd_n = -1;
    return 0;
}


// ---- fused clip_grad_norm_ + AdamW (utils/train.py:242-244)
struct osteo_adamw {
    int device = -1;                         // device of the tables (the current device at create): every later call runs under a DeviceGuard
    int n = 0;
    std::vector<long long> numel;
    std::vector<const void*> host_ptrs;      // [4 * n]: last uploaded params | grads | exp_avg | exp_avg_sq
    void* pinned = nullptr;                  // staging for the pointer tables
    osteo::DevBuf ptrs, numel_dev, chunk_tensor, chunk_start, acc;
    int chunks = 0;
    ~osteo_adamw() {
        if (pinned) cudaFreeHost(pinned);
    }
};

int osteo_adamw_create(osteo_adamw** out, int n_tensors, const long long* numel_host) {
    if (!out || n_tensors <= 0 || !numel_host) return fail("adamw_create: bad arguments");
    if (osteo_device_count() <= 0) return fail("no CUDA device: this library has no CPU fallback");
    std::unique_ptr<osteo_adamw> h(new osteo_adamw);
    OSTEO_CUDA(cudaGetDevice(&h->device));
    h->n = n_tensors;
    h->numel.assign(numel_host, numel_host + n_tensors);
    std::vector<int> ct;
    std::vector<long long> cs;
    for (int i = 0; i < n_tensors; ++i) {
        if (numel_host[i] < 0) return fail("adamw_create: negative element count");
        for (long long s0 = 0; s0 < numel_host[i]; s0 += OPT_CHUNK) {
            ct.push_back(i);
            cs.push_back(s0);
        }
    }
    h->chunks = static_cast<int>(ct.size());
    if (h->chunks == 0) return fail("adamw_create: nothing to optimise");
    OSTEO_TRY(h->ptrs.alloc(sizeof(void*) * 4 * n_tensors));
    OSTEO_TRY(h->numel_dev.alloc(sizeof(long long) * n_tensors));
    OSTEO_TRY(h->chunk_tensor.alloc(sizeof(int) * ct.size()));
    OSTEO_TRY(h->chunk_start.alloc(sizeof(long long) * cs.size()));
    OSTEO_TRY(h->acc.alloc(sizeof(double)));
    OSTEO_CUDA(cudaMallocHost(&h->pinned, sizeof(void*) * 4 * n_tensors));
    OSTEO_CUDA(cudaMemcpy(h->numel_dev.p, numel_host, sizeof(long long) * n_tensors, cudaMemcpyHostToDevice));
    OSTEO_CUDA(cudaMemcpy(h->chunk_tensor.p, ct.data(), sizeof(int) * ct.size(), cudaMemcpyHostToDevice));
    OSTEO_CUDA(cudaMemcpy(h->chunk_start.p, cs.data(), sizeof(long long) * cs.size(), cudaMemcpyHostToDevice));
    *out = h.release();
    return 0;
}

int osteo_adamw_destroy(osteo_adamw* h) {
    if (!h) return 0;
    DeviceGuard device_guard(h->device);
    delete h;
    return 0;
}

int osteo_adamw_step(osteo_adamw* h, float* const* params_dev, float* const* grads_dev, float* const* exp_avg_dev, float* const* exp_avg_sq_dev, double lr, double beta1,
                     double beta2, double eps, double weight_decay, long long step, double max_norm, float* norm_out_dev, void* stream) {
    if (!h) return fail("adamw_step: null handle");
    if (!params_dev || !grads_dev || !exp_avg_dev || !exp_avg_sq_dev) return fail("adamw_step: null pointer table");
    if (step < 1) return fail("adamw_step: step counts from 1");
    DeviceGuard device_guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n = h->n;
    std::vector<const void*> now(4 * n);
    for (int i = 0; i < n; ++i) {
        now[i] = params_dev[i];
        now[n + i] = grads_dev[i];
        now[2 * n + i] = exp_avg_dev[i];
        now[3 * n + i] = exp_avg_sq_dev[i];
        if (h->numel[i] > 0 && (!params_dev[i] || !grads_dev[i] || !exp_avg_dev[i] || !exp_avg_sq_dev[i])) return fail("adamw_step: tensor %d has a null pointer", i);
    }
    if (now != h->host_ptrs) {
        // the staging buffer may still be in flight from the previous upload: wait for it (rare: addresses are stable in steady state)
        if (!h->host_ptrs.empty()) OSTEO_CUDA(cudaStreamSynchronize(s));
        std::memcpy(h->pinned, now.data(), sizeof(void*) * 4 * n);
        OSTEO_CUDA(cudaMemcpyAsync(h->ptrs.p, h->pinned, sizeof(void*) * 4 * n, cudaMemcpyHostToDevice, s));
        h->host_ptrs = now;
    }
    OptTables t;
    float** base = h->ptrs.as<float*>();
    t.params = base;
    t.grads = base + n;
    t.exp_avg = base + 2 * n;
    t.exp_avg_sq = base + 3 * n;
    t.numel = h->numel_dev.as<long long>();
    t.chunk_tensor = h->chunk_tensor.as<int>();
    t.chunk_start = h->chunk_start.as<long long>();
    if (max_norm > 0.0) {
        OSTEO_CUDA(cudaMemsetAsync(h->acc.p, 0, sizeof(double), s));
        opt_sumsq_kernel<<<h->chunks, 256, 0, s>>>(t, h->acc.as<double>());
        OSTEO_CUDA(cudaGetLastError());
    }
    const double bc1 = 1.0 - std::pow(beta1, static_cast<double>(step));
    const double bc2 = 1.0 - std::pow(beta2, static_cast<double>(step));
    opt_adamw_kernel<<<h->chunks, 256, 0, s>>>(t, h->acc.as<double>(), static_cast<float>(max_norm), static_cast<float>(1.0 - lr * weight_decay),
                                               static_cast<float>(1.0 - beta1), static_cast<float>(beta2), static_cast<float>(1.0 - beta2), static_cast<float>(eps),
                                               static_cast<float>(lr / bc1), static_cast<float>(std::sqrt(bc2)), norm_out_dev);
    OSTEO_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
