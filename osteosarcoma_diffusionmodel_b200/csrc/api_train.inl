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
    int max_hidden = 0;
    for (auto& hb : c->halves) max_hidden = hb->lin.n > max_hidden ? hb->lin.n : max_hidden;
    const size_t part_cols = static_cast<size_t>(3 * max_hidden > c->D ? 3 * max_hidden : c->D);
    OSTEO_TRY(w.partials.alloc(static_cast<size_t>(cap / 32) * part_cols * 4));
    (void)max_w;
    w.cap = cap;
    return 0;
}

static int finish_partials(osteo_ddpm_ctx* c, long long n, int nq, int N, float* o0, float* o1, float* o2, cudaStream_t s) {
    const int slabs = static_cast<int>((n + 31) / 32);
    dim3 grid((N + 31) / 32, nq), block(32, 8);
    partials_finish_kernel<<<grid, block, 0, s>>>(c->train.partials.as<float>(), slabs, nq, N, o0, o1, o2);
    OSTEO_CUDA(cudaGetLastError());
    ++c->launches;
    return 0;
}

static int outer_accum(osteo_ddpm_ctx* c, const float* G, int gm, const float* X, int xk, const int* idx, long long n, float* dW, cudaStream_t s) {
    const size_t smem = static_cast<size_t>(64) * (gm + xk) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        OSTEO_CUDA(cudaFuncSetAttribute(outer_accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        configured = true;
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

}  // namespace osteo

extern "C" {

int osteo_ddpm_enable_training(osteo_ddpm_ctx* c, int enable) {
    OSTEO_TRY(check_ctx(c));
    OSTEO_CUDA(cudaSetDevice(c->device));
    if (!enable) return 0;
    if (c->out_proj.wt.p) return 0;
    OSTEO_CUDA(cudaDeviceSynchronize());
    for (auto& hb : c->halves) OSTEO_TRY(hb->lin.init_transposed());
    OSTEO_TRY(c->out_proj.init_transposed());
    c->have_weights = false;   // W^T copies are filled by the next osteo_ddpm_set_weights
    return 0;
}

int osteo_ddpm_train_step(osteo_ddpm_ctx* c, const float* x0_dev, const float* cond_dev, long long n, const int* t_idx_dev, const float* noise_dev,
                          const uint8_t* const* drop_masks_dev, int train, uint64_t seed, long long row_base, float* loss_dev, float* const* grads_dev,
                          int n_tensors, void* stream) {
    OSTEO_TRY(check_ctx(c));
    OSTEO_TRY(require_ready(c, n));
    if (!t_idx_dev) return fail("train_step: t_idx_dev is required");
    if (!loss_dev) return fail("train_step: loss_dev is required");
    const bool want_grads = grads_dev != nullptr;
    const int H = static_cast<int>(c->halves.size());
    const int expect = osteo_ddpm_num_weight_tensors(static_cast<int>(c->hidden.size()));
    if (want_grads) {
        if (n_tensors != expect) return fail("train_step: expected %d gradient tensors, got %d", expect, n_tensors);
        if (!c->out_proj.wt.p) return fail("train_step: call osteo_ddpm_enable_training(ctx, 1) and osteo_ddpm_set_weights before requesting gradients");
        for (int i = 0; i < n_tensors; ++i)
            if (!grads_dev[i]) return fail("train_step: gradient tensor %d is null", i);
    }
    OSTEO_CUDA(cudaSetDevice(c->device));
    OSTEO_TRY(ensure_train_ws(c));
    TrainWorkspace& w = c->train;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int D = c->D, DP = c->DP, h0 = c->h0(), E = c->E;
    const bool x3 = c->x3();

    // ------------------------------------------------------------------ forward
    zero_double_kernel<<<1, 1, 0, s>>>(c->loss_acc.as<double>());
    train_prepare_kernel<<<grid_for(n * (DP / 4), 256, c->sms), 256, 0, s>>>(x0_dev, noise_dev, t_idx_dev, n, D, c->sqrt_ab.as<float>(), c->sqrt_1mab.as<float>(),
                                                                             w.noise.as<float>(), DP, w.xt_bf.as<__nv_bfloat16>(), 2 * DP, c->lo(DP), seed, row_base);
    OSTEO_CUDA(cudaGetLastError());
    {
        const size_t smem = sizeof(float) * 16 * (c->C + 2 * E);
        cond_path_kernel<<<static_cast<unsigned>((n + 15) / 16), 256, smem, s>>>(cond_dev, n, c->C, E, h0, c->ce_w0t.as<float>(), c->ce_b0.as<float>(),
                                                                               c->ce_w2t.as<float>(), c->ce_b2.as<float>(), c->cp_wt.as<float>(), c->cp_b.as<float>(),
                                                                               c->cproj.as<float>(), w.pre0.as<float>(), w.cemb.as<float>());
        OSTEO_CUDA(cudaGetLastError());
    }
    c->launches += 3;
    OSTEO_TRY(launch_input_proj(c, 0, n, t_idx_dev, s, &w.xt_tmap));
    c->h0_primed = false;      // acts[0] now belongs to this training batch, not to a sampling state
    for (int i = 0; i < H; ++i) {
        HalfOpts o;
        o.train = train != 0;
        o.seed = seed;
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
            p.col_partials = w.partials.as<float>();
        }
        OSTEO_TRY(after_launch(c, launch_gemm(EPI_MSE, 64, p, c->sms, s), s));
    }
    finish_loss_kernel<<<1, 1, 0, s>>>(c->loss_acc.as<double>(), loss_dev, 1.0 / (static_cast<double>(n) * D));
    OSTEO_CUDA(cudaGetLastError());
    ++c->launches;
    if (!want_grads) return 0;

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
    {
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

    // output_proj: bias gradient from the MSE epilogue's column partials, weight gradient = deps^T . act_last
    OSTEO_TRY(finish_partials(c, n, 1, D, grads_dev[gi_out_b], nullptr, nullptr, s));
    {
        const ActBuf& a = *c->acts.back();
        OSTEO_TRY(after_launch(c, launch_wgrad(w.deps.as<__nv_bfloat16>(), 2 * DP, DP, D, a.ptr(), 2 * a.width, 0, a.width, a.width, grads_dev[gi_out_w], a.width, n, x3,
                                               c->status_dev.as<int>(), c->sms, s), s));
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
    for (int j = H - 1; j >= -1; --j) {
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
        p.col_partials = w.partials.as<float>();
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
                p.row_base = row_base;
            }
            OSTEO_TRY(after_launch(c, launch_gemm(EPI_GN_BWD, hb.gw, p, c->sms, s), s));
            const int gi = 10 + 4 * j;
            OSTEO_TRY(finish_partials(c, n, 3, width, grads_dev[gi + 2], grads_dev[gi + 3], grads_dev[gi + 1], s));
            // weight gradient of this half: one launch per concatenated source
            const ActBuf& a0 = *c->acts[hb.src0];
            OSTEO_TRY(after_launch(c, launch_wgrad(w.dy[j]->as<__nv_bfloat16>(), 2 * width, width, width, a0.ptr(), 2 * a0.width, 0, a0.width, a0.width, grads_dev[gi],
                                                   hb.lin.k, n, x3, c->status_dev.as<int>(), c->sms, s), s));
            if (hb.src1 >= 0) {
                const ActBuf& a1 = *c->acts[hb.src1];
                OSTEO_TRY(after_launch(c, launch_wgrad(w.dy[j]->as<__nv_bfloat16>(), 2 * width, width, width, a1.ptr(), 2 * a1.width, 0, a1.width, a1.width,
                                                       grads_dev[gi] + a0.width, hb.lin.k, n, x3, c->status_dev.as<int>(), c->sms, s), s));
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
            OSTEO_TRY(finish_partials(c, n, 1, h0, grads_dev[5], nullptr, nullptr, s));
            OSTEO_CUDA(cudaMemcpyAsync(grads_dev[7], grads_dev[5], sizeof(float) * h0, cudaMemcpyDeviceToDevice, s));
            OSTEO_CUDA(cudaMemcpyAsync(grads_dev[9], grads_dev[5], sizeof(float) * h0, cudaMemcpyDeviceToDevice, s));
        }
    }
    // input_proj weight gradient: dh0^T . x_t
    OSTEO_TRY(after_launch(c, launch_wgrad(w.dh0_bf.as<__nv_bfloat16>(), 2 * h0, h0, h0, w.xt_bf.as<__nv_bfloat16>(), 2 * DP, 0, DP, D, grads_dev[4], D, n, x3, c->status_dev.as<int>(),
                                           c->sms, s), s));
    // time_proj / cond_proj / ConditionalEmbedding (fp32 CUDA-core kernels; tiny matrices)
    OSTEO_TRY(outer_accum(c, w.dh0_f32.as<float>(), h0, c->emb_table.as<float>(), c->TD, t_idx_dev, n, grads_dev[8], s));
    OSTEO_TRY(outer_accum(c, w.dh0_f32.as<float>(), h0, w.cemb.as<float>(), E, nullptr, n, grads_dev[6], s));
    {
        const size_t smem = sizeof(float) * 8 * (h0 + E);
        // Wc^T view: cond_bwd needs Wc as [h0, E] row-major, which is exactly cond_proj.weight's layout.
        cond_bwd_rows_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, smem, s>>>(w.dh0_f32.as<float>(), n, h0, E, c->cp_w.as<float>(), c->ce_w2.as<float>(),
                                                                                 w.pre0.as<float>(), w.dcemb.as<float>(), w.h1.as<float>(), w.dpre0.as<float>());
        OSTEO_CUDA(cudaGetLastError());
        ++c->launches;
    }
    {
        dim3 grid((E + 63) / 64, 64);
        colsum_f32_kernel<<<grid, 64, 0, s>>>(w.dcemb.as<float>(), n, E, grads_dev[3]);
        colsum_f32_kernel<<<grid, 64, 0, s>>>(w.dpre0.as<float>(), n, E, grads_dev[1]);
        OSTEO_CUDA(cudaGetLastError());
        c->launches += 2;
    }
    OSTEO_TRY(outer_accum(c, w.dcemb.as<float>(), E, w.h1.as<float>(), E, nullptr, n, grads_dev[2], s));
    OSTEO_TRY(outer_accum(c, w.dpre0.as<float>(), E, cond_dev, c->C, nullptr, n, grads_dev[0], s));
    return 0;
}

}  // extern "C"
