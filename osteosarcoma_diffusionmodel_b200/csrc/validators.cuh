// Kernels behind the validators (utils/validation.py): row norms for the RBF Gram and the
// column-gathered moment reduction for Pearson correlations. See api_ops.inl.
#pragma once
#include "common.cuh"

namespace osteo {
}  // namespace osteo
