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

int osteo_mmd_partial(const float* x_dev, long long n, const float* y_dev, long long m, int d, float gamma, const float* center_dev, long long row_begin,
                      long long row_end, long long yrow_begin, long long yrow_end, int precision, double* sums_dev, void* stream) {
    return fail("osteo_mmd_partial: not built yet");
}

int osteo_corr_moments(const float* data_dev, long long n, int ld, const int* cols_dev, int k, const float* shift_dev, long long row_begin, long long row_end,
                       double* out_dev, void* stream) {
    return fail("osteo_corr_moments: not built yet");
}

}  // extern "C"
