// Training step C-ABI (included by osteo_ddpm.cu).
extern "C" int osteo_ddpm_train_step(osteo_ddpm_ctx* ctx, const float* x0_dev, const float* cond_dev, long long n, const int* t_idx_dev,
                                     const float* noise_dev, const uint8_t* const* drop_masks_dev, int train, uint64_t seed, long long row_base,
                                     float* loss_dev, float* const* grads_dev, int n_tensors, void* stream) {
    return osteo::fail("osteo_ddpm_train_step: not built yet");
}
