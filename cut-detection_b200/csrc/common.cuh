// Shared helpers for libcutdet_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/cutdet_b200.h"

namespace cutdet {

// Thread-local error text behind cutdet_last_error().
void set_error(const char *fmt, ...);
int fail(int code, const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define CUTDET_CUDA(expr)                                           \
    do {                                                            \
        cudaError_t _e = (expr);                                    \
        if (_e != cudaSuccess) return ::cutdet::cuda_fail(_e, #expr); \
    } while (0)

#define CUTDET_LAUNCH_CHECK(name)                                   \
    do {                                                            \
        cudaError_t _e = cudaGetLastError();                        \
        if (_e != cudaSuccess) return ::cutdet::cuda_fail(_e, name); \
    } while (0)

#define CUTDET_REQUIRE(cond, ...)                                   \
    do {                                                            \
        if (!(cond)) return ::cutdet::fail(CUTDET_EINVAL, __VA_ARGS__); \
    } while (0)

static inline cudaStream_t as_stream(cutdet_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();

// Wrap every kernel launch: counts launches (cutdet_launch_count) and, while profiling is on
// (cutdet_profile_begin/end), brackets the launch with CUDA events on its own stream.
class KernelScope {
public:
    KernelScope(const char *name, cudaStream_t stream);
    ~KernelScope();
private:
    int slot_;
    cudaStream_t stream_;
};

}  // namespace cutdet
