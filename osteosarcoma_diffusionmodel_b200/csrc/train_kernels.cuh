// Training-step workspace and kernels (backward of the denoiser). See api_train.inl.
#pragma once
#include <memory>
#include <vector>
#include "common.cuh"

namespace osteo {

// Tensors saved by the training forward for the backward pass, one entry per half block:
// normalised pre-affine activations x_hat (bf16 [cap, 2*n] = [hi|lo]) and 1/sigma per (row, group).
struct TrainWorkspace {
    std::vector<std::unique_ptr<DevBuf>> xhat, rstd;
    size_t bytes() const {
        size_t b = 0;
        for (auto& p : xhat) b += p->bytes;
        for (auto& p : rstd) b += p->bytes;
        return b;
    }
    void release() {
        xhat.clear();
        rstd.clear();
    }
};

}  // namespace osteo
