// Host-side helpers shared by the C-ABI translation units: error reporting,
// TMA tensor-map construction (driver entry point fetched through the runtime,
// so the library does not link against libcuda) and launch bookkeeping.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <cuda.h>
#include <cuda_runtime.h>

namespace osteo {

std::string& last_error_ref();
int fail(const char* fmt, ...);

#define OSTEO_CUDA(expr)                                                                              \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) return ::osteo::fail("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
    } while (0)

#define OSTEO_TRY(expr)            \
    do {                           \
        int _r = (expr);           \
        if (_r != 0) return _r;    \
    } while (0)

inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

// RAII: make `device` the calling thread's current CUDA device for the duration of a C-ABI call and put the caller's device back on
// exit. Every entry point that takes a context (or a handle that remembers its device) opens one, so a model on cuda:1 works while
// the process's current device is cuda:0, and no call -- destroy included -- leaves the thread on a different device.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int device) {
        if (device >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
// context check + device guard: the first statement of every entry point that takes an osteo_ddpm_ctx*
#define OSTEO_CTX(c)                   \
    OSTEO_TRY(::osteo::check_ctx(c));  \
    ::osteo::DeviceGuard _osteo_device_guard((c)->device)

// cudaFuncSetAttribute (the opt-in to > 48 KB of dynamic shared memory) and occupancy queries are PER DEVICE: a launcher keeps one of
// these per kernel instead of a process-wide `static bool configured`, or a model on cuda:1 launches an unconfigured kernel
// ("invalid argument") after a model on cuda:0 configured it. `value` carries a per-device result (e.g. co-resident cluster count).
struct PerDevice {
    static constexpr int MAX_DEVICES = 64;
    bool done[MAX_DEVICES] = {};
    int value[MAX_DEVICES] = {};
    static int current() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= MAX_DEVICES) d = 0;
        return d;
    }
    bool configured() const { return done[current()]; }
    void set_configured() { done[current()] = true; }
    int& val() { return value[current()]; }
};

// bf16 row-major [rows, cols] with leading dimension ld (elements); box = box_rows x 64 columns,
// 128-byte swizzle (matches make_kmajor_sw128_desc). Out-of-bounds elements read as zero.
int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

// bf16 row-major [rows, cols]; box = 32 rows x 32 columns (64-byte rows), 64-byte swizzle: the per-warp TMA-store box of gemm_ws.cuh.
int make_tmap_bf16_st32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld);

// fp32 row-major [rows, cols] (ld in elements); box = box_rows x 32 columns (128-byte rows), 128-byte swizzle.
int make_tmap_f32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

int sm_count(int device);

// Device buffer owned by the context.
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int alloc(size_t n) {
        release();
        if (n == 0) return 0;
        OSTEO_CUDA(cudaMalloc(&p, n));
        bytes = n;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    ~DevBuf() { release(); }
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};


}  // namespace osteo
