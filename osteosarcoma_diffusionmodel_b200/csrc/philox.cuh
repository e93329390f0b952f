// Counter-based RNG for the DDPM noise: Philox4x32 keyed by the sampling seed,
// counter = (column/4, global_row_lo, global_row_hi, stream<<16 | step), so a
// row's noise is independent of how rows are sharded over GPUs.
// Rounds: 10 (the Random123 / cuRAND default) for everything drawn once per patient or training row; 7 -- the smallest round count
// Salmon et al. (SC'11, table 2) report as passing BigCrush ("Crush-resistant"), 10 being 7 plus a safety margin -- for the
// reverse-step noise stream, which is 5142 x 1000 normals per patient and the bulk of every integer instruction the sampling loop issues.
// The numpy restatement lives in oracle/philox_oracle.py and is compared bit-exactly
// (uint32 words) and within 2e-6 (Box-Muller normals) in tests/test_philox.py.
//
// Two word -> normal mappings:
//   * two words per Box-Muller pair (23-bit radius uniform, 23-bit angle), 4 normals per Philox block, counter word 0 = column / 4:
//     x_T, q_sample noise (drawn once per patient / training row);
//   * PACKED, for the reverse-step noise z (5142 x 1000 normals per patient: Philox was ~60 % of every instruction the sampling loop
//     executed): ONE word per pair -- radius uniform from the top 20 bits (u1 = 1 - k / 2^20 in (0, 1], |z| <= 5.27), angle from the
//     low 12 bits (4096 directions: every trigonometric moment below order 4096 is exact) -- so a block yields 8 normals and
//     counter word 0 = column / 8. Half the Philox blocks: +9 % patients/s at the board's power cap.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace osteo {

enum PhiloxStream : uint32_t {
    STREAM_REVERSE = 0,   // z in the reverse update (models/diffusion.py:409)
    STREAM_XT = 1,        // x_T = randn (models/diffusion.py:443)
    STREAM_QNOISE = 2,    // q_sample noise (models/diffusion.py:335)
    STREAM_DROPOUT = 3,   // + block index: dropout masks (models/diffusion.py:204)
    STREAM_TIMESTEP = 16  // randint t (models/diffusion.py:361)
};

constexpr int PHILOX_ROUNDS = 10;             // x_T, q_sample noise, dropout masks, timesteps
constexpr int PHILOX_ROUNDS_REVERSE = 7;      // STREAM_REVERSE (the per-step z)
__host__ __device__ constexpr int philox_rounds(uint32_t stream) { return stream == 0u /* STREAM_REVERSE */ ? PHILOX_ROUNDS_REVERSE : PHILOX_ROUNDS; }

template <int R>
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) { return philox4x32<PHILOX_ROUNDS>(c, k); }

// 10-round words of any stream but STREAM_REVERSE (dropout masks, timesteps, x_T, q_sample noise).
__device__ __forceinline__ uint4 philox_words(uint64_t seed, uint64_t row, uint32_t col4, uint32_t stream, uint32_t step) {
    const uint4 ctr = make_uint4(col4, static_cast<uint32_t>(row), static_cast<uint32_t>(row >> 32), (stream << 16) | (step & 0xFFFFu));
    const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    return philox4x32_10(ctr, key);
}
// Words of a stream with that stream's round count (philox_rounds): the generic form behind the test hook and philox_normal4.
__device__ __forceinline__ uint4 philox_words_stream(uint64_t seed, uint64_t row, uint32_t col, uint32_t stream, uint32_t step) {
    const uint4 ctr = make_uint4(col, static_cast<uint32_t>(row), static_cast<uint32_t>(row >> 32), (stream << 16) | (step & 0xFFFFu));
    const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    return stream == STREAM_REVERSE ? philox4x32<PHILOX_ROUNDS_REVERSE>(ctr, key) : philox4x32<PHILOX_ROUNDS>(ctr, key);
}

// 23-bit uniforms built with integer ops only (no I2F on the XU pipe): the mantissa trick gives f in [1, 2).
__device__ __forceinline__ float unit_1_2(uint32_t w) { return __uint_as_float(0x3f800000u | (w >> 9)); }
// [0, 1): used for Bernoulli draws (dropout keep-masks).
__device__ __forceinline__ float u01(uint32_t w) { return unit_1_2(w) - 1.0f; }

// Box-Muller on the MUFU approximations (lg2 / sqrt / sin / cos). u1 = 2 - f is in (0, 1] (never 0: the log is finite),
// u2 = f - 1 is in [0, 1); both are exact in fp32, so the fp64 oracle sees the same uniforms.
__device__ __forceinline__ void box_muller(uint32_t w0, uint32_t w1, float& z0, float& z1) {
    const float u1 = 2.0f - unit_1_2(w0), u2 = unit_1_2(w1) - 1.0f;
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * __log2f(u1)));   // sqrt(-2 ln u1)
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);
    z0 = r * c;
    z1 = r * s;
}

// PACKED mapping: one word -> two normals (see the header of this file).
__device__ __forceinline__ float packed_radius_uniform(uint32_t w) { return 2.0f - __uint_as_float(0x3f800000u | ((w >> 12) << 3)); }   // (0, 1], 2^20 levels
__device__ __forceinline__ float packed_angle_1_2(uint32_t w) { return __uint_as_float(0x3f800000u | ((w & 0xFFFu) << 11)); }           // [1, 2), 2^12 levels
__device__ __forceinline__ void box_muller_packed(uint32_t w, float& z0, float& z1) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * __log2f(packed_radius_uniform(w))));   // sqrt(-2 ln u1)
    float s, c;
    __sincosf(6.283185307179586f * (packed_angle_1_2(w) - 1.0f), &s, &c);
    z0 = r * c;
    z1 = r * s;
}
__device__ __forceinline__ bool stream_is_packed(uint32_t stream) { return stream == STREAM_REVERSE; }

// N independent Philox4x32-R blocks advanced round by round (round loop outermost): N independent dependency
// chains in flight instead of one, which is what keeps the integer pipes busy with few warps per scheduler.
template <int N, int R>
__device__ __forceinline__ void philox4x32_batch(uint4 (&c)[N], uint2 k) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const uint32_t hi0 = __umulhi(M0, c[i].x), lo0 = M0 * c[i].x;
            const uint32_t hi1 = __umulhi(M1, c[i].z), lo1 = M1 * c[i].z;
            c[i] = make_uint4(hi1 ^ c[i].y ^ k.x, lo1, hi0 ^ c[i].w ^ k.y, lo0);
        }
        k.x += W0;
        k.y += W1;
    }
}

// 4*N normals for columns [4*col4_0, 4*(col4_0 + N)) of one row: same values as N calls of philox_normal4.
template <int N>
__device__ __forceinline__ void philox_normal_row(uint64_t seed, uint64_t row, uint32_t col4_0, uint32_t stream, uint32_t step, float (&z)[4 * N]) {
    uint4 c[N];
#pragma unroll
    for (int i = 0; i < N; ++i)
        c[i] = make_uint4(col4_0 + i, static_cast<uint32_t>(row), static_cast<uint32_t>(row >> 32), (stream << 16) | (step & 0xFFFFu));
    if (stream == STREAM_REVERSE) philox4x32_batch<N, PHILOX_ROUNDS_REVERSE>(c, make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
    else philox4x32_batch<N, PHILOX_ROUNDS>(c, make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
#pragma unroll
    for (int i = 0; i < N; ++i) {
        box_muller(c[i].x, c[i].y, z[4 * i + 0], z[4 * i + 1]);
        box_muller(c[i].z, c[i].w, z[4 * i + 2], z[4 * i + 3]);
    }
}

// PACKED: 8*N normals for columns [8*col8_0, 8*(col8_0 + N)) of one row.
template <int N>
__device__ __forceinline__ void philox_normal_row_packed(uint64_t seed, uint64_t row, uint32_t col8_0, uint32_t stream, uint32_t step, float (&z)[8 * N]) {
    uint4 c[N];
#pragma unroll
    for (int i = 0; i < N; ++i)
        c[i] = make_uint4(col8_0 + i, static_cast<uint32_t>(row), static_cast<uint32_t>(row >> 32), (stream << 16) | (step & 0xFFFFu));
    if (stream == STREAM_REVERSE) philox4x32_batch<N, PHILOX_ROUNDS_REVERSE>(c, make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
    else philox4x32_batch<N, PHILOX_ROUNDS>(c, make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
#pragma unroll
    for (int i = 0; i < N; ++i) {
        box_muller_packed(c[i].x, z[8 * i + 0], z[8 * i + 1]);
        box_muller_packed(c[i].y, z[8 * i + 2], z[8 * i + 3]);
        box_muller_packed(c[i].z, z[8 * i + 4], z[8 * i + 5]);
        box_muller_packed(c[i].w, z[8 * i + 6], z[8 * i + 7]);
    }
}

// The 4 normals of columns [4*col4, 4*col4 + 4) of one row, in the stream's mapping (elementwise kernels: 4 columns per thread).
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t row, uint32_t col4, uint32_t stream, uint32_t step) {
    float4 z;
    if (stream_is_packed(stream)) {
        const uint4 w = philox_words_stream(seed, row, col4 >> 1, stream, step);
        box_muller_packed((col4 & 1u) ? w.z : w.x, z.x, z.y);
        box_muller_packed((col4 & 1u) ? w.w : w.y, z.z, z.w);
        return z;
    }
    const uint4 w = philox_words_stream(seed, row, col4, stream, step);
    box_muller(w.x, w.y, z.x, z.y);
    box_muller(w.z, w.w, z.z, z.w);
    return z;
}

}  // namespace osteo
