// Host-side launch of the tcgen05 GEMM family.
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"

namespace osteo {

template <int EPI, int GW>
int launch_gemm_inst(const GemmParams& p, int num_sms, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        OSTEO_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<EPI, GW>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
        configured = true;
    }
    const int tiles = p.m_tiles * p.n_tiles;
    if (tiles <= 0) return 0;
    const int grid = tiles < num_sms ? tiles : num_sms;
    gemm_tc_kernel<EPI, GW><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(p);
    OSTEO_CUDA(cudaGetLastError());
    return 0;
}

inline int launch_gemm(int epi, int gw, const GemmParams& p, int num_sms, cudaStream_t stream) {
    switch (epi) {
        case EPI_LINEAR: return launch_gemm_inst<EPI_LINEAR, 64>(p, num_sms, stream);
        case EPI_DDPM: return launch_gemm_inst<EPI_DDPM, 64>(p, num_sms, stream);
        case EPI_MSE: return launch_gemm_inst<EPI_MSE, 64>(p, num_sms, stream);
        case EPI_RBF: return launch_gemm_inst<EPI_RBF, 64>(p, num_sms, stream);
        case EPI_GN_SILU:
            switch (gw) {
                case 16: return launch_gemm_inst<EPI_GN_SILU, 16>(p, num_sms, stream);
                case 32: return launch_gemm_inst<EPI_GN_SILU, 32>(p, num_sms, stream);
                case 64: return launch_gemm_inst<EPI_GN_SILU, 64>(p, num_sms, stream);
                default: return fail("GroupNorm group width %d not supported (need 16, 32 or 64)", gw);
            }
        default: return fail("unknown epilogue %d", epi);
    }
}

// Append the K-segments of one operand pair: A columns [a_col0, a_col0 + k) against W columns
// [b_col0, b_col0 + k). In FP32X3 mode the hi/lo halves (A lo at +a_lo_off, W lo at +b_lo_off)
// contribute hi*hi + hi*lo + lo*hi.
inline int add_segments(GemmParams& p, int a_sel, int a_col0, int a_lo_off, int b_col0, int b_lo_off, int k, bool x3) {
    if (k % BK != 0) return fail("segment K=%d is not a multiple of %d", k, BK);
    const int nkb = k / BK;
    auto push = [&](int ac, int bc) -> int {
        if (p.nseg >= MAX_KSEG) return fail("too many K segments");
        p.seg[p.nseg++] = KSeg{a_sel, ac, bc, nkb, 0, 0};
        return 0;
    };
    OSTEO_TRY(push(a_col0, b_col0));
    if (x3) {
        OSTEO_TRY(push(a_col0, b_col0 + b_lo_off));
        OSTEO_TRY(push(a_col0 + a_lo_off, b_col0));
    }
    return 0;
}

}  // namespace osteo
