// common.cuh -- shared helpers for libb2n (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b2n.h"

#define B2N_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void b2n_set_error(const char *fmt, ...);

#define B2N_CHECK_ARG(cond, msg)                                     \
    do {                                                             \
        if (!(cond)) {                                               \
            b2n_set_error("%s: %s", __func__, msg);                  \
            return 1;                                                \
        }                                                            \
    } while (0)

// async launch errors (bad configuration etc.) are surfaced right after the launch
#define B2N_LAUNCH_CHECK()                                                        \
    do {                                                                          \
        cudaError_t e_ = cudaPeekAtLastError();                                   \
        if (e_ != cudaSuccess) {                                                  \
            b2n_set_error("%s: CUDA error %s", __func__, cudaGetErrorString(e_)); \
            (void)cudaGetLastError();                                             \
            return 2;                                                             \
        }                                                                         \
    } while (0)

static inline unsigned b2n_blocks(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }
// persistent-style grid: enough CTAs to cover n, capped at `per_sm` resident CTAs on every SM
static inline unsigned b2n_grid(int64_t n_ctas_needed, int per_sm) {
    int64_t cap = (int64_t)B2N_SMS * per_sm;
    if (n_ctas_needed < 1) n_ctas_needed = 1;
    return (unsigned)(n_ctas_needed < cap ? n_ctas_needed : cap);
}

// cudaLaunchKernelEx wrapper used by the gather / scatter / Adam kernels (argument conversion to the kernel's types)
template <class... KArgs, class... Args>
static inline cudaError_t b2n_launch(void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t stream,
                                     Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ uint32_t b2n_expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ uint32_t b2n_morton3D(uint32_t x, uint32_t y, uint32_t z) {
    return b2n_expand_bits(x) | (b2n_expand_bits(y) << 1) | (b2n_expand_bits(z) << 2);
}
__device__ __forceinline__ uint32_t b2n_compact_bits(uint32_t x) {
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

// effective element count: min(n, *n_dev) when a device-side count is supplied (no host sync)
__device__ __forceinline__ int64_t b2n_eff_n(int64_t n, const int32_t *n_dev) {
    if (n_dev == nullptr) return n;
    int64_t m = (int64_t)__ldg(n_dev);
    return m < n ? m : n;
}
