// Host-side launch of the tcgen05 GEMM family.
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"

namespace osteo {

template <int EPI, int GW, bool MN = false>
int launch_gemm_inst(const GemmParams& p, int num_sms, cudaStream_t stream) {
    static PerDevice dev_state;
    if (!dev_state.configured()) {
        OSTEO_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<EPI, GW, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes<EPI>()));
        dev_state.set_configured();
    }
    int tiles = (MN ? p.splits : 1) * p.m_tiles * p.n_tiles;
    if (tiles <= 0) return 0;
    if (EPI == EPI_DDPM && p.a_resident) tiles = p.m_tiles * p.n_chunks;      // work units, not tiles
    const int grid = tiles < num_sms ? tiles : num_sms;
    gemm_tc_kernel<EPI, GW, MN><<<grid, gemm_threads<EPI, GW>(), gemm_smem_bytes<EPI>(), stream>>>(p);
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
            if (p.N > GN_PAR_MAX) return fail("Linear+GroupNorm layer of width %d: the epilogue keeps bias / gamma / beta of at most %d columns in shared memory", p.N, GN_PAR_MAX);
            switch (gw) {
                case 16: return launch_gemm_inst<EPI_GN_SILU, 16>(p, num_sms, stream);
                case 32: return launch_gemm_inst<EPI_GN_SILU, 32>(p, num_sms, stream);
                case 64: return launch_gemm_inst<EPI_GN_SILU, 64>(p, num_sms, stream);
                default: return fail("GroupNorm group width %d not supported (need 16, 32 or 64)", gw);
            }
        case EPI_GN_BWD:
            switch (gw) {
                case 16: return launch_gemm_inst<EPI_GN_BWD, 16>(p, num_sms, stream);
                case 32: return launch_gemm_inst<EPI_GN_BWD, 32>(p, num_sms, stream);
                case 64: return launch_gemm_inst<EPI_GN_BWD, 64>(p, num_sms, stream);
                default: return fail("GroupNorm group width %d not supported (need 16, 32 or 64)", gw);
            }
        case EPI_WGRAD: return launch_gemm_inst<EPI_WGRAD, 64, true>(p, num_sms, stream);
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

// dW[n_out, k_in] (+)= dY^T X over `rows` batch rows, both operands read MN-major straight from their row-major
// [rows, ld] bf16 [hi|lo] buffers (no transposed copies). dW must be zero-initialised: row splits accumulate atomically.
inline int launch_wgrad(const __nv_bfloat16* dy, int dy_ld, int dy_lo_off, int n_out, const __nv_bfloat16* x, int x_ld, int x_col0, int x_lo_off, int k_in,
                        float* dw, int dw_ld, long long rows, bool x3, int* status, int num_sms, cudaStream_t stream) {
    GemmParams p;
    std::memset(&p, 0, sizeof p);
    OSTEO_TRY(make_tmap_bf16(&p.tma_a[0], dy, rows, dy_ld, dy_ld, 64));
    p.tma_a[1] = p.tma_a[0];
    OSTEO_TRY(make_tmap_bf16(&p.tma_b[0], x, rows, x_ld, x_ld, 64));
    p.tma_b[1] = p.tma_b[0];
    p.seg[p.nseg++] = KSeg{0, 0, x_col0, 0, 0, 0};
    if (x3) {
        p.seg[p.nseg++] = KSeg{0, 0, x_col0 + x_lo_off, 0, 0, 0};
        p.seg[p.nseg++] = KSeg{0, dy_lo_off, x_col0, 0, 0, 0};
    }
    p.M = n_out;
    p.N = k_in;
    p.m_tile0 = 0;
    p.m_tiles = (n_out + BM - 1) / BM;
    p.n_tiles = (k_in + BN - 1) / BN;
    p.k_rows = static_cast<int>(rows);
    const int total_kb = static_cast<int>((rows + BK - 1) / BK);
    const int mn = p.m_tiles * p.n_tiles;
    int splits = (2 * num_sms + mn - 1) / mn;
    // every split ends in an atomic accumulation of its whole [128 x 64] tile: below ~8 k-blocks (512 rows) per split the atomics and the
    // pipeline fill cost more than the MMAs (the [256 x 128] time_proj gradient at batch 8192: 64 splits of 2 k-blocks, 48 us)
    if (splits > total_kb / 8) splits = total_kb / 8;
    if (splits > total_kb) splits = total_kb;
    if (splits < 1) splits = 1;
    p.kb_per_split = (total_kb + splits - 1) / splits;
    p.splits = (total_kb + p.kb_per_split - 1) / p.kb_per_split;
    p.out_f32 = dw;
    p.out_f32_ld = dw_ld;
    p.status = status;
    return launch_gemm(EPI_WGRAD, 64, p, num_sms, stream);
}

}  // namespace osteo
