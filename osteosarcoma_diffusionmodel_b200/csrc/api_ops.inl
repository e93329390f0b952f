// Context-free building blocks of the C-ABI (included by osteo_ddpm.cu): the raw tcgen05 Linear
// (+GroupNorm+SiLU) used by the unit tests, the Philox test hooks and the validator kernels.
namespace osteo {

static int current_sms() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    return sm_count(dev);
}

// Shared driver of osteo_linear_tc / osteo_linear_gn_silu_tc: packs fp32 operands to bf16 [hi|lo],
// runs one GEMM launch, synchronises and frees the temporaries.
static int linear_tc_impl(const float* a_dev, const float* w_dev, const float* bias_dev, const float* gamma_dev, const float* beta_dev, float* out_dev, int m,
                          int n, int k, int precision, bool gn, cudaStream_t s) {
    if (osteo_device_count() <= 0) return fail("no CUDA device: this library has no CPU fallback");
    if (m <= 0 || n <= 0 || k <= 0) return fail("linear_tc: bad shape m=%d n=%d k=%d", m, n, k);
    if (gn && (n % 128 != 0 || (n / 8 != 16 && n / 8 != 32 && n / 8 != 64))) return fail("linear_gn_silu_tc: n=%d needs n/8 in {16,32,64} and n %% 128 == 0", n);
    const int sms = current_sms();
    const bool x3 = precision == OSTEO_PREC_FP32X3;
    const int kp = static_cast<int>(round_up(k, BK)), np = static_cast<int>(round_up(n, BN));
    const long long mp = round_up(m, BM);
    DevBuf a_bf, w_bf, bias_p, out_bf, status;
    OSTEO_TRY(a_bf.alloc(static_cast<size_t>(mp) * 2 * kp * 2));
    OSTEO_TRY(w_bf.alloc(static_cast<size_t>(np) * 2 * kp * 2));
    OSTEO_TRY(bias_p.alloc(static_cast<size_t>(np) * 4));
    OSTEO_TRY(status.alloc(sizeof(int)));
    OSTEO_CUDA(cudaMemsetAsync(status.p, 0, sizeof(int), s));
    OSTEO_CUDA(cudaMemsetAsync(bias_p.p, 0, static_cast<size_t>(np) * 4, s));
    if (bias_dev) OSTEO_CUDA(cudaMemcpyAsync(bias_p.p, bias_dev, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToDevice, s));
    pack_bf16_hilo_kernel<<<grid_for(mp * (kp / 4), 256, sms), 256, 0, s>>>(a_dev, m, k, k, a_bf.as<__nv_bfloat16>(), mp, kp, 2LL * kp, kp);
    OSTEO_CUDA(cudaGetLastError());
    pack_bf16_hilo_kernel<<<grid_for(static_cast<long long>(np) * (kp / 4), 256, sms), 256, 0, s>>>(w_dev, n, k, k, w_bf.as<__nv_bfloat16>(), np, kp, 2LL * kp, kp);
    OSTEO_CUDA(cudaGetLastError());

    GemmParams p;
    std::memset(&p, 0, sizeof p);
    OSTEO_TRY(make_tmap_bf16(&p.tma_a[0], a_bf.p, mp, 2 * kp, 2 * kp, BM));
    p.tma_a[1] = p.tma_a[0];
    OSTEO_TRY(make_tmap_bf16(&p.tma_b[0], w_bf.p, np, 2 * kp, 2 * kp, BN));
    p.tma_b[1] = p.tma_b[0];
    OSTEO_TRY(add_segments(p, 0, 0, kp, 0, kp, kp, x3));
    p.M = m;
    p.N = n;
    p.m_tile0 = 0;
    p.m_tiles = static_cast<int>(mp / BM);
    p.n_tiles = np / BN;
    p.status = status.as<int>();
    p.bias = bias_p.as<float>();
    p.gn_eps = 1e-5f;
    int rc;
    if (gn) {
        OSTEO_TRY(out_bf.alloc(static_cast<size_t>(mp) * 2 * np * 2));
        p.gamma = gamma_dev;
        p.beta = beta_dev;
        p.out_bf = out_bf.as<__nv_bfloat16>();
        p.out_bf_ld = 2 * np;
        p.out_lo_off = np;   // always keep the residual so the fp32 result can be reassembled
        rc = launch_gemm(EPI_GN_SILU, n / 8, p, sms, s);
        if (rc == 0) {
            unpack_hilo_kernel<<<grid_for(static_cast<long long>(m) * n, 256, sms), 256, 0, s>>>(out_bf.as<__nv_bfloat16>(), 2LL * np, np, out_dev, m, n);
            OSTEO_CUDA(cudaGetLastError());
        }
    } else {
        p.out_f32 = out_dev;
        p.out_f32_ld = n;
        rc = launch_gemm(EPI_LINEAR, 64, p, sms, s);
    }
    if (rc != 0) return rc;
    int h = 0;
    OSTEO_CUDA(cudaMemcpyAsync(&h, status.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    OSTEO_CUDA(cudaStreamSynchronize(s));
    if (h != 0) return fail("tcgen05 pipeline error %d (1 = TMA producer, 2 = MMA issuer, 3 = epilogue wait timed out)", h);
    return 0;
}

}  // namespace osteo

extern "C" {

int osteo_linear_tc(const float* a_dev, const float* w_dev, const float* bias_dev, float* out_dev, int m, int n, int k, int precision, void* stream) {
    return linear_tc_impl(a_dev, w_dev, bias_dev, nullptr, nullptr, out_dev, m, n, k, precision, false, static_cast<cudaStream_t>(stream));
}

int osteo_linear_gn_silu_tc(const float* a_dev, const float* w_dev, const float* bias_dev, const float* gamma_dev, const float* beta_dev, float* out_dev, int m,
                            int n, int k, int precision, void* stream) {
    if (!gamma_dev || !beta_dev || !bias_dev) return fail("linear_gn_silu_tc: bias, gamma and beta are required");
    return linear_tc_impl(a_dev, w_dev, bias_dev, gamma_dev, beta_dev, out_dev, m, n, k, precision, true, static_cast<cudaStream_t>(stream));
}

int osteo_wgrad_tc(const float* dy_dev, const float* x_dev, float* dw_dev, long long rows, int n_out, int k_in, int precision, void* stream) {
    if (osteo_device_count() <= 0) return fail("no CUDA device: this library has no CPU fallback");
    if (rows <= 0 || n_out <= 0 || k_in <= 0) return fail("wgrad_tc: bad shape");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int sms = current_sms();
    const bool x3 = precision == OSTEO_PREC_FP32X3;
    const int np = static_cast<int>(round_up(n_out, 64)), kp = static_cast<int>(round_up(k_in, 64));
    DevBuf dy_bf, x_bf, status;
    OSTEO_TRY(dy_bf.alloc(static_cast<size_t>(rows) * 2 * np * 2));
    OSTEO_TRY(x_bf.alloc(static_cast<size_t>(rows) * 2 * kp * 2));
    OSTEO_TRY(status.alloc(sizeof(int)));
    OSTEO_CUDA(cudaMemsetAsync(status.p, 0, sizeof(int), s));
    pack_bf16_hilo_kernel<<<grid_for(rows * (np / 4), 256, sms), 256, 0, s>>>(dy_dev, rows, n_out, n_out, dy_bf.as<__nv_bfloat16>(), rows, np, 2LL * np, np);
    pack_bf16_hilo_kernel<<<grid_for(rows * (kp / 4), 256, sms), 256, 0, s>>>(x_dev, rows, k_in, k_in, x_bf.as<__nv_bfloat16>(), rows, kp, 2LL * kp, kp);
    OSTEO_CUDA(cudaGetLastError());
    OSTEO_CUDA(cudaMemsetAsync(dw_dev, 0, static_cast<size_t>(n_out) * k_in * sizeof(float), s));
    OSTEO_TRY(launch_wgrad(dy_bf.as<__nv_bfloat16>(), 2 * np, np, n_out, x_bf.as<__nv_bfloat16>(), 2 * kp, 0, kp, k_in, dw_dev, k_in, rows, x3, status.as<int>(), sms, s));
    int h = 0;
    OSTEO_CUDA(cudaMemcpyAsync(&h, status.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    OSTEO_CUDA(cudaStreamSynchronize(s));
    if (h != 0) return fail("tcgen05 pipeline error %d in wgrad", h);
    return 0;
}

int osteo_philox_normal(float* out_dev, long long n, int d, uint64_t seed, long long row_base, uint32_t stream_id, uint32_t step, void* stream) {
    if (osteo_device_count() <= 0) return fail("no CUDA device: this library has no CPU fallback");
    if (n <= 0 || d <= 0) return fail("philox_normal: bad shape");
    const long long items = n * ((d + 3) / 4);
    philox_normal_kernel<<<grid_for(items, 256, current_sms()), 256, 0, static_cast<cudaStream_t>(stream)>>>(out_dev, n, d, seed, row_base, stream_id, step);
    OSTEO_CUDA(cudaGetLastError());
    return 0;
}

int osteo_philox_words(uint32_t* out_dev, long long n, int ncol4, uint64_t seed, long long row_base, uint32_t stream_id, uint32_t step, void* stream) {
    if (osteo_device_count() <= 0) return fail("no CUDA device: this library has no CPU fallback");
    if (n <= 0 || ncol4 <= 0) return fail("philox_words: bad shape");
    philox_words_kernel<<<grid_for(n * ncol4, 256, current_sms()), 256, 0, static_cast<cudaStream_t>(stream)>>>(out_dev, n, ncol4, seed, row_base, stream_id, step);
    OSTEO_CUDA(cudaGetLastError());
    return 0;
}

// Shared body of osteo_mmd_partial (contiguous Gram-row ranges) and osteo_mmd_partial_cyclic (cyc_world > 0: this rank owns the 128-row
// blocks b with b % cyc_world == cyc_rank of every Gram, and K(X,X) / K(Y,Y) are reduced as symmetric half-Grams -- a row block's tiles
// below the diagonal are skipped and the off-diagonal ones count twice, which is valid per row block, so the halves of all ranks add up
// to the whole sum while the block-cyclic assignment keeps the triangular work balanced).
static int mmd_partial_impl(const float* x_dev, long long n, const float* y_dev, long long m, int d, float gamma, const float* center_dev, long long row_begin,
                            long long row_end, long long yrow_begin, long long yrow_end, int cyc_rank, int cyc_world, int precision, double* sums_dev, void* stream) {
    if (osteo_device_count() <= 0) return fail("no CUDA device: this library has no CPU fallback");
    if (n <= 0 || m <= 0 || d <= 0) return fail("mmd_partial: bad shape");
    if (row_begin < 0 || row_end > n || row_begin > row_end || yrow_begin < 0 || yrow_end > m || yrow_begin > yrow_end) return fail("mmd_partial: bad row range");
    if ((row_begin < row_end && row_begin % BM != 0) || (yrow_begin < yrow_end && yrow_begin % BM != 0)) return fail("mmd_partial: shard starts must be multiples of %d", BM);
    if (n > 2000000000LL || m > 2000000000LL) return fail("mmd_partial: more than 2^31 rows");
    if (cyc_world > 0 && (cyc_rank < 0 || cyc_rank >= cyc_world)) return fail("mmd_partial_cyclic: rank %d outside [0, %d)", cyc_rank, cyc_world);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int sms = current_sms();
    const bool x3 = precision == OSTEO_PREC_FP32X3;
    const int kp = static_cast<int>(round_up(d, BK));
    const long long np = round_up(n, BM), mp = round_up(m, BM);
    // Packed operands live in a grow-only per-process workspace: cudaMalloc / cudaFree of hundreds of MB per call costs more
    // than the Gram itself at 16k rows (and synchronises the device). Not thread-safe; one validator per process.
    static DevBuf xb, yb, nx, ny, status;
    static int ws_device = -1;
    int dev_now = 0;
    OSTEO_CUDA(cudaGetDevice(&dev_now));
    if (dev_now != ws_device) {
        for (DevBuf* b : {&xb, &yb, &nx, &ny, &status}) b->release();
        ws_device = dev_now;
    }
    auto grow = [](DevBuf& b, size_t bytes) -> int {
        if (b.bytes >= bytes) return 0;
        OSTEO_CUDA(cudaDeviceSynchronize());
        return b.alloc(bytes);
    };
    OSTEO_TRY(grow(xb, static_cast<size_t>(np) * 2 * kp * 2));
    OSTEO_TRY(grow(yb, static_cast<size_t>(mp) * 2 * kp * 2));
    OSTEO_TRY(grow(nx, static_cast<size_t>(np) * 4));
    OSTEO_TRY(grow(ny, static_cast<size_t>(mp) * 4));
    OSTEO_TRY(grow(status, sizeof(int)));
    OSTEO_CUDA(cudaMemsetAsync(status.p, 0, sizeof(int), s));
    OSTEO_CUDA(cudaMemsetAsync(sums_dev, 0, 3 * sizeof(double), s));
    const int lo = x3 ? kp : 0;
    pack_center_norm_kernel<<<sms * 8, 256, 0, s>>>(x_dev, n, d, center_dev, xb.as<__nv_bfloat16>(), np, kp, lo, nx.as<float>());
    pack_center_norm_kernel<<<sms * 8, 256, 0, s>>>(y_dev, m, d, center_dev, yb.as<__nv_bfloat16>(), mp, kp, lo, ny.as<float>());
    OSTEO_CUDA(cudaGetLastError());
    // A-side maps (128-row boxes) and B-side maps (256-row boxes) of the same two packed operands: gemm_rbf.cuh, 128 x 256 tiles
    CUtensorMap tx, ty, txb, tyb;
    OSTEO_TRY(make_tmap_bf16(&tx, xb.p, np, 2 * kp, 2 * kp, BM));
    OSTEO_TRY(make_tmap_bf16(&ty, yb.p, mp, 2 * kp, 2 * kp, BM));
    OSTEO_TRY(make_tmap_bf16(&txb, xb.p, np, 2 * kp, 2 * kp, RB_BN));
    OSTEO_TRY(make_tmap_bf16(&tyb, yb.p, mp, 2 * kp, 2 * kp, RB_BN));
    auto gram = [&](const CUtensorMap& ta, const CUtensorMap& tb, const float* na, const float* nb, long long rb, long long re, long long cols, bool symmetric,
                    double* acc) -> int {
        if (re <= rb) return 0;
        RbfParams p;
        std::memset(&p, 0, sizeof p);
        p.tma_a = ta;
        p.tma_b = tb;
        const int nkb = kp / BK;
        p.seg[p.nseg++] = KSeg{0, 0, 0, nkb, 0, 0};                 // hi . hi
        if (x3) {
            p.seg[p.nseg++] = KSeg{0, 0, kp, nkb, 0, 0};            // hi . lo
            p.seg[p.nseg++] = KSeg{0, kp, 0, nkb, 0, 0};            // lo . hi
        }
        p.M = static_cast<int>(re);
        p.N = static_cast<int>(cols);
        p.m_tile0 = static_cast<int>(rb / BM);
        p.m_tiles = static_cast<int>((re - rb + BM - 1) / BM);
        p.m_stride = 1;
        if (cyc_world > 0) {      // block-cyclic: row blocks cyc_rank, cyc_rank + cyc_world, ... of [rb, re) = [0, rows)
            const int blocks = p.m_tiles;
            p.m_tile0 = cyc_rank;
            p.m_stride = cyc_world;
            p.m_tiles = blocks > cyc_rank ? (blocks - cyc_rank + cyc_world - 1) / cyc_world : 0;
            if (p.m_tiles == 0) return 0;
        }
        p.n_tiles = static_cast<int>((cols + RB_BN - 1) / RB_BN);
        p.status = status.as<int>();
        p.norm_a = na;
        p.norm_b = nb;
        p.gamma = gamma;
        p.symmetric = symmetric ? 1 : 0;
        p.acc = acc;
        return launch_rbf_gram(p, sms, s);
    };
    // K(X,X) rows [row_begin,row_end): with contiguous ranges the symmetric half-Gram is used only when this call owns every row (a
    // contiguous shard of a triangle is unbalanced); block-cyclic sharding always uses it
    const bool cyc = cyc_world > 0;
    OSTEO_TRY(gram(tx, txb, nx.as<float>(), nx.as<float>(), row_begin, row_end, n, cyc || (row_begin == 0 && row_end == n), sums_dev + 0));
    OSTEO_TRY(gram(ty, tyb, ny.as<float>(), ny.as<float>(), yrow_begin, yrow_end, m, cyc || (yrow_begin == 0 && yrow_end == m), sums_dev + 1));
    OSTEO_TRY(gram(tx, tyb, nx.as<float>(), ny.as<float>(), row_begin, row_end, m, false, sums_dev + 2));
    int h = 0;
    OSTEO_CUDA(cudaMemcpyAsync(&h, status.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    OSTEO_CUDA(cudaStreamSynchronize(s));
    if (h != 0) return fail("tcgen05 pipeline error %d in mmd_partial", h);
    return 0;
}

int osteo_mmd_partial(const float* x_dev, long long n, const float* y_dev, long long m, int d, float gamma, const float* center_dev, long long row_begin,
                      long long row_end, long long yrow_begin, long long yrow_end, int precision, double* sums_dev, void* stream) {
    return mmd_partial_impl(x_dev, n, y_dev, m, d, gamma, center_dev, row_begin, row_end, yrow_begin, yrow_end, 0, 0, precision, sums_dev, stream);
}

int osteo_mmd_partial_cyclic(const float* x_dev, long long n, const float* y_dev, long long m, int d, float gamma, const float* center_dev, int rank, int world,
                             int precision, double* sums_dev, void* stream) {
    if (world <= 0) return fail("mmd_partial_cyclic: world size must be positive");
    return mmd_partial_impl(x_dev, n, y_dev, m, d, gamma, center_dev, 0, n, 0, m, rank, world, precision, sums_dev, stream);
}

int osteo_corr_moments(const float* data_dev, long long n, int ld, const int* cols_dev, int k, const float* shift_dev, long long row_begin, long long row_end,
                       double* out_dev, void* stream) {
    if (osteo_device_count() <= 0) return fail("no CUDA device: this library has no CPU fallback");
    if (k <= 0 || k > 32) return fail("corr_moments: k=%d outside [1, 32]", k);
    if (row_begin < 0 || row_end > n || row_begin > row_end) return fail("corr_moments: bad row range");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    OSTEO_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double) * (1 + k + k * k), s));
    if (row_end == row_begin) return 0;
    long long warps = row_end - row_begin;
    const int sms = current_sms();
    long long blocks = (warps + 7) / 8;
    if (blocks > sms * 4LL) blocks = sms * 4LL;
    corr_moments_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(data_dev, ld, cols_dev, k, shift_dev, row_begin, row_end, out_dev);
    OSTEO_CUDA(cudaGetLastError());
    return 0;
}

int osteo_corr_moments_batched(const float* data_dev, long long n, int ld, int ncols, const int* cols_dev, int n_sets, const float* shift_dev, long long row_begin,
                               long long row_end, double* out_dev, void* stream) {
    if (osteo_device_count() <= 0) return fail("no CUDA device: this library has no CPU fallback");
    if (n_sets <= 0 || n_sets > 32) return fail("corr_moments_batched: %d column sets outside [1, 32]", n_sets);
    if (ncols <= 0 || ncols > ld) return fail("corr_moments_batched: ncols=%d outside [1, ld=%d]", ncols, ld);
    if (row_begin < 0 || row_end > n || row_begin > row_end) return fail("corr_moments_batched: bad row range");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    OSTEO_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double) * n_sets * CM_STRIDE, s));
    if (row_end == row_begin) return 0;
    // rows per chunk: as many whole rows as fit in 96 KB of shared memory, at most 64
    int chunk_rows = static_cast<int>((96 * 1024) / (static_cast<size_t>(ncols) * 4));
    if (chunk_rows > 64) chunk_rows = 64;
    if (chunk_rows < 1) return fail("corr_moments_batched: a row of %d columns does not fit the 96 KB staging buffer", ncols);
    const size_t smem = static_cast<size_t>(chunk_rows) * ncols * 4;
    static PerDevice dev_state;
    if (!dev_state.configured()) {
        OSTEO_CUDA(cudaFuncSetAttribute(corr_moments_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        dev_state.set_configured();
    }
    const long long nchunks = (row_end - row_begin + chunk_rows - 1) / chunk_rows;
    const int sms = current_sms();
    const long long blocks = nchunks < 2LL * sms ? nchunks : 2LL * sms;
    // the kernel holds a 32 x 32 fp64 moment block per warp in registers (128 per thread): at most 16 warps = 16 sets per launch
    for (int s0 = 0; s0 < n_sets; s0 += 16) {
        const int ns = n_sets - s0 < 16 ? n_sets - s0 : 16;
        corr_moments_batched_kernel<<<static_cast<unsigned>(blocks), 32 * ns, smem, s>>>(data_dev, ld, ncols, cols_dev + s0 * 32, ns, shift_dev + s0 * 32, row_begin, row_end,
                                                                                          chunk_rows, out_dev + static_cast<size_t>(s0) * CM_STRIDE);
        OSTEO_CUDA(cudaGetLastError());
    }
    return 0;
}

int osteo_corr_moments_tiled(const float* data_dev, long long n, int ld, int ncols, const int* cols_dev, int n_sets, int max_set_size, const float* shift_dev,
                             long long row_begin, long long row_end, double* out_dev, void* stream) {
    if (osteo_device_count() <= 0) return fail("no CUDA device: this library has no CPU fallback");
    if (n_sets <= 0 || n_sets > 32) return fail("corr_moments_tiled: %d column sets outside [1, 32]", n_sets);
    if (max_set_size <= 0 || max_set_size > 32) return fail("corr_moments_tiled: sets of up to %d columns (limit 32)", max_set_size);
    if (ncols <= 0 || ncols > ld) return fail("corr_moments_tiled: ncols=%d outside [1, ld=%d]", ncols, ld);
    if (row_begin < 0 || row_end > n || row_begin > row_end) return fail("corr_moments_tiled: bad row range");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    OSTEO_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double) * n_sets * CM_STRIDE, s));
    if (row_end == row_begin) return 0;
    // sets per launch: every (set, 4 x 4 block) pair needs a thread (51 sets of <= 16 columns, 14 of <= 32) and the launch has to fit its
    // 226 KB of shared memory; more sets than that take several passes over the rows
    const int W = max_set_size <= 16 ? 16 : 32;
    int per = max_set_size <= 16 ? CT_THREADS / 10 : CT_THREADS / 36;
    while (per > 1 && moments_tiled_smem(per, W, ncols, 2) > static_cast<size_t>(CT_SMEM_LIMIT)) --per;
    for (int s0 = 0; s0 < n_sets; s0 += per) {
        const int ns = n_sets - s0 < per ? n_sets - s0 : per;
        const int* c0 = cols_dev + s0 * 32;
        const float* sh0 = shift_dev ? shift_dev + s0 * 32 : nullptr;
        double* o0 = out_dev + static_cast<size_t>(s0) * CM_STRIDE;
        if (max_set_size <= 16) OSTEO_TRY(launch_moments_tiled<4>(data_dev, ld, ncols, c0, ns, sh0, row_begin, row_end, o0, current_sms(), s));
        else OSTEO_TRY(launch_moments_tiled<8>(data_dev, ld, ncols, c0, ns, sh0, row_begin, row_end, o0, current_sms(), s));
    }
    return 0;
}

int osteo_coherence_finish(const double* moments_dev, const int* cols_dev, int n_sets, double* scores_dev, void* stream) {
    if (osteo_device_count() <= 0) return fail("no CUDA device: this library has no CPU fallback");
    if (n_sets <= 0) return fail("coherence_finish: no column sets");
    coherence_finish_kernel<<<(n_sets * 32 + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(moments_dev, cols_dev, n_sets, scores_dev);
    OSTEO_CUDA(cudaGetLastError());
    return 0;
}

int osteo_corr_loss_finish(const double* moments_dev, const int* cols_dev, const float* shift_dev, const int* modes_dev, int n_sets, float* loss_out_dev,
                           float* coef_out_dev, void* stream) {
    if (osteo_device_count() <= 0) return fail("no CUDA device: this library has no CPU fallback");
    if (n_sets <= 0 || n_sets > 32) return fail("corr_loss_finish: %d column sets outside [1, 32]", n_sets);
    corr_loss_finish_kernel<<<(n_sets * 32 + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(moments_dev, cols_dev, shift_dev, modes_dev, n_sets, loss_out_dev,
                                                                                                       coef_out_dev);
    OSTEO_CUDA(cudaGetLastError());
    return 0;
}

int osteo_corr_loss_backward(const float* data_dev, long long n, int ld, const int* cols_dev, int n_sets, const float* coef_dev, const float* upstream_dev,
                             float* grad_dev, void* stream) {
    if (osteo_device_count() <= 0) return fail("no CUDA device: this library has no CPU fallback");
    if (n_sets <= 0 || n_sets > 32) return fail("corr_loss_backward: %d column sets outside [1, 32]", n_sets);
    if (n <= 0) return 0;
    const int sms = current_sms();
    const long long blocks = n < 8LL * sms ? n : 8LL * sms;
    corr_loss_bwd_kernel<<<static_cast<unsigned>(blocks), 32 * n_sets, 0, static_cast<cudaStream_t>(stream)>>>(data_dev, n, ld, cols_dev, n_sets, coef_dev, upstream_dev,
                                                                                                                grad_dev);
    OSTEO_CUDA(cudaGetLastError());
    return 0;
}

int osteo_mixup_rows(const float* src_dev, long long src_rows, int d, const long long* idx_a_dev, const long long* idx_b_dev, long long n, float lam, float one_minus_lam,
                     float* out_dev, void* stream) {
    if (osteo_device_count() <= 0) return fail("no CUDA device: this library has no CPU fallback");
    if (!src_dev || !out_dev || d <= 0 || src_rows <= 0) return fail("mixup_rows: bad arguments");
    if (n <= 0) return 0;
    if (!idx_a_dev && n > src_rows) return fail("mixup_rows: %lld rows requested from a %lld-row source without an index", n, src_rows);
    const long long items = n * ((d & 1) == 0 ? d / 2 : d);
    mixup_gather_kernel<<<grid_for(items, 256, current_sms()), 256, 0, static_cast<cudaStream_t>(stream)>>>(src_dev, d, idx_a_dev, idx_b_dev, lam, one_minus_lam, out_dev, n);
    OSTEO_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
