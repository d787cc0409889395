// encoding.cu -- multiresolution hash grid (fw / table-bw), frequency and spherical-harmonics encodings.
// Replaces the tiny-cuda-nn encodings NGP builds (ngp_pl/models/networks.py:34-53,63-70).
//
// Hash grid design: gather-bound on the L1-miss path (one sector request per clock and SM).
// Forward: TWO lanes per sample (lane parity = x-corner, so the pair's gathers fall into one 128-byte line), a warp
// walks all levels of 16 consecutive samples of a ray, four levels (16 independent 4-byte gathers) in flight per lane,
// fp32 accumulation, one fp16 rounding, the 64-byte output row written as four 16-byte stores.  The fp16 table
// (21.8 MiB at T=2^19) stays L2-resident on B200 (126 MB L2); at T=2^22 (185 MiB) the fine levels come from HBM.
// Backward: one lane per (alive) sample, runs of lanes in one cell are summed with a segmented shuffle reduction on
// the coarse levels, then vector red.global.add.v2.f32 straight into the fp32 gradient table (no fp16 underflow, no
// separate cast pass).  Entry indices avoid the integer division of `index % size`: hashed levels have a
// power-of-two size (mask), dense levels overshoot the size by less than one period (conditional subtract).
// Algorithmic bytes per sample: 12 in + 16 levels x 8 corners x 4 B gathered + 64 out = 588 B.
#include "common.cuh"
#include "hashgrid.cuh"

extern "C" int b2n_hashgrid_layout(int n_levels, int n_features, int log2_hashmap_size, int base_resolution,
                                   double per_level_scale, b2n_grid_layout *layout) {
    B2N_CHECK_ARG(layout && n_levels >= 1 && n_levels <= B2N_MAX_LEVELS && n_features == 2 &&
                  log2_hashmap_size >= 3 && log2_hashmap_size <= 30 && base_resolution >= 1, "bad hash grid config");
    layout->n_levels = n_levels; layout->n_features = n_features;
    uint64_t off = 0;
    for (int l = 0; l < n_levels; ++l) {
        // b^l in double, snapped to an integer when within 1e-6 so that the intended resolutions
        // (N_min ... 2048*scale) do not hinge on one ulp of exp2f/log2f (DESIGN.md "Hash grid").
        double v = pow(per_level_scale, (double)l) * base_resolution;
        if (fabs(v - nearbyint(v)) < 1e-6 * v) v = nearbyint(v);
        const float s = (float)(v - 1.0);
        const uint32_t res = (uint32_t)ceilf(s) + 1;
        uint64_t n = (uint64_t)res * res * res;
        if (n > 0x7fffffffull) n = 0x7fffffffull;
        n = (n + 7) / 8 * 8;
        if (n > (1ull << log2_hashmap_size)) n = 1ull << log2_hashmap_size;
        layout->scale[l] = s; layout->resolution[l] = res; layout->size[l] = (uint32_t)n;
        layout->offset[l] = (uint32_t)off;
        off += n;
        B2N_CHECK_ARG(off < 0xffffffffull, "hash table too large");
    }
    layout->offset[n_levels] = (uint32_t)off;
    layout->x_offset = 0.0f; layout->x_scale = 1.0f;
    return 0;
}

// Lane mapping of the forward kernel: TWO lanes per sample, lane parity = x-corner (x0 or x0+1); each lane
// handles the four (y,z) corners of its x.  Entries of x-neighbours are adjacent (dense levels) or, with the
// coherent prime hash, differ only in their low bits, so the two lanes of a pair hit the same 128-byte line and a
// warp-level gather touches half as many L1 wavefronts as a one-lane-per-sample mapping (the kernel is L1
// wavefront bound on the fine levels: 10 hashed levels x 8 corners x 32 distinct lines per warp instruction).
// The two half-results are combined with one xor-shuffle.
#define B2N_PRAGMA_(x) _Pragma(#x)
#define B2N_PRAGMA(x) B2N_PRAGMA_(x)
#ifndef HG_FW_UNROLL
#define HG_FW_UNROLL 4       // levels in flight per lane (16 independent gathers)
#endif
#ifndef HG_FW_CTAS
#define HG_FW_CTAS 16
#endif
#ifndef HG_BW_CTAS
#define HG_BW_CTAS 16
#endif
#ifndef B2N_BW_PASS_BYTES
#define B2N_BW_PASS_BYTES (72ull << 20)      // gradient bytes one level-major scatter pass may keep resident in the 126 MB L2
#endif
__global__ void __launch_bounds__(128) hashgrid_fw_kernel(const float *__restrict__ x,
                                                          const __half2 *__restrict__ table,
                                                          const __grid_constant__ GridLevels g, int64_t n,
                                                          const int32_t *__restrict__ n_dev,
                                                          __half *__restrict__ out, int out_stride, int l_begin,
                                                          int l_end) {
    n = b2n_eff_n(n, n_dev);
    const int lane = threadIdx.x & 31, cx = lane & 1;
    const bool vec16 = (out_stride % 8 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);   // rows 16-byte aligned
    const int64_t pairs_per_grid = ((int64_t)gridDim.x * blockDim.x) >> 1;
    // warp-uniform loop: a warp covers 16 consecutive samples per iteration
    for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) >> 1; base < n; base += pairs_per_grid) {
        const int64_t i = base + (lane >> 1);
        const bool live = i < n;
        const int64_t ii = live ? i : n - 1;
        const float px = (__ldg(x + 3 * ii) - g.x_offset) * g.x_scale, py = (__ldg(x + 3 * ii + 1) - g.x_offset) * g.x_scale,
                    pz = (__ldg(x + 3 * ii + 2) - g.x_offset) * g.x_scale;
        __half *row = out + ii * out_stride;
        // four levels per block: 16 independent gathers in flight per lane, one 16-byte store per block
        for (int l0 = l_begin; l0 < l_end; l0 += 4) {
            uint32_t packed[4];
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int l = l0 + q;
                packed[q] = 0u;
                if (l < l_end) {
                    Corner4 c;
                    level_corners4(px, py, pz, g.scale[l], g.resolution[l], g.size[l], g.offset[l], g.mode[l], cx, c);
                    __half2 v[4];
                    #pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = __ldg(table + c.idx[k]);
                    float a0 = 0.f, a1 = 0.f;
                    #pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 f = __half22float2(v[k]);
                        a0 = fmaf(c.w[k], f.x, a0);
                        a1 = fmaf(c.w[k], f.y, a1);
                    }
                    a0 += __shfl_xor_sync(0xffffffffu, a0, 1);
                    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
                    const __half2 r = __floats2half2_rn(a0, a1);
                    packed[q] = *reinterpret_cast<const uint32_t *>(&r);
                }
            }
            if (cx == 0 && live) {
                if (vec16 && l0 + 4 <= l_end) {
                    *reinterpret_cast<uint4 *>(row + 2 * l0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                } else {
                    #pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (l0 + q < l_end) *reinterpret_cast<uint32_t *>(row + 2 * (l0 + q)) = packed[q];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(128) hashgrid_bw_kernel(const float *__restrict__ x,
                                                          const __half *__restrict__ dy, int dy_stride,
                                                          const __grid_constant__ GridLevels g, int64_t n,
                                                          const int32_t *__restrict__ n_dev, float grad_scale,
                                                          float2 *__restrict__ grad_table,
                                                          const int32_t *__restrict__ sample_idx, int l_begin, int l_end) {
    // warp-aggregated scatter: hashgrid.cuh::scatter_level
    n = b2n_eff_n(n, n_dev);
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < n; base += stride) {  // warp-uniform
        const int64_t i = base + lane;
        const bool live = i < n;
        const int64_t ii = live ? i : n - 1;
        const int64_t xi = sample_idx ? (int64_t)__ldg(sample_idx + ii) : ii;     // compacted backward: row ii <-> sample xi
        const float px = (__ldg(x + 3 * xi) - g.x_offset) * g.x_scale, py = (__ldg(x + 3 * xi + 1) - g.x_offset) * g.x_scale,
                    pz = (__ldg(x + 3 * xi + 2) - g.x_offset) * g.x_scale;
        const __half2 *row = reinterpret_cast<const __half2 *>(dy + ii * dy_stride);
        #pragma unroll 1
        for (int l = l_begin; l < l_end; ++l) {
            const float2 gr = __half22float2(__ldg(row + l));
            scatter_level(px, py, pz, gr.x * grad_scale, gr.y * grad_scale, l, g, grad_table, live, lane);
        }
    }
}

extern "C" int b2n_hashgrid_fw(const float *x, const b2n_half *table, const b2n_grid_layout *layout, int64_t n,
                               const int32_t *n_dev, b2n_half *out, int out_stride, void *stream) {
    GridLevels g;
    if (to_levels(layout, g)) return 1;
    B2N_CHECK_ARG(out_stride >= 2 * g.n_levels && out_stride % 2 == 0, "out_stride too small / odd");
    if (n <= 0) return 0;
    // A table that does not fit the L2 (T = 2^22: 185 MiB fp16) is walked LEVEL-MAJOR: one launch per block of four
    // levels (4 x 16.8 MB), whose slice of the table is fetched from HBM once and then served by the L2 while every
    // sample gathers from it (measured at 1.45 M samples: 512 us in one pass, 444 / 385 / 478 with 2 / 4 / 8 levels a pass).
#ifndef HG_FW_PASS_LEVELS
#define HG_FW_PASS_LEVELS 4
#endif
    const bool level_major = HG_FW_PASS_LEVELS > 0 && (uint64_t)layout->offset[g.n_levels] * 4 > (96ull << 20);
    const int step = level_major ? HG_FW_PASS_LEVELS : g.n_levels;
    for (int l0 = 0; l0 < g.n_levels; l0 += step)
        b2n_launch(hashgrid_fw_kernel, b2n_grid(b2n_blocks(2 * n, 128), HG_FW_CTAS), 128, (cudaStream_t)stream,
                   x, (const __half2 *)table, g, n, n_dev, (__half *)out, out_stride, l0,
                   l0 + step < g.n_levels ? l0 + step : g.n_levels);
    B2N_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2n_hashgrid_bw(const float *x, const b2n_half *dL_dout, int dy_stride,
                               const b2n_grid_layout *layout, int64_t n, const int32_t *n_dev,
                               float grad_scale, float *grad_table, const int32_t *sample_idx, void *stream) {
    GridLevels g;
    if (to_levels(layout, g)) return 1;
    B2N_CHECK_ARG(dy_stride >= 2 * g.n_levels && dy_stride % 2 == 0, "dy_stride too small / odd");
    if (n <= 0) return 0;
    // Level-major passes when the fp32 gradient table does not fit the L2 (T = 2^22: 388 MB): consecutive levels are
    // grouped while their gradient slices stay under the budget (a full hashed level is 33.5 MB), so the atomics of a
    // pass resolve in the L2 and each line is written back once instead of a DRAM read-modify-write per update.
    const bool level_major = (uint64_t)layout->offset[g.n_levels] * 8 > (96ull << 20);
    int l0 = 0;
    while (l0 < g.n_levels) {
        int l1 = l0 + 1;
        if (!level_major) l1 = g.n_levels;
        else
            while (l1 < g.n_levels && (uint64_t)(layout->offset[l1 + 1] - layout->offset[l0]) * 8 <= B2N_BW_PASS_BYTES) ++l1;
        b2n_launch(hashgrid_bw_kernel, b2n_grid(b2n_blocks(n, 128), HG_BW_CTAS), 128, (cudaStream_t)stream,
                   x, (const __half *)dL_dout, dy_stride, g, n, n_dev, grad_scale, (float2 *)grad_table, sample_idx, l0, l1);
        l0 = l1;
    }
    B2N_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Frequency encoding: column j of the 3*F*2 encoded columns is sin(2^f * pi * x_d + (j&1) * pi/2) with
// d = j / (2F), f = (j/2) % F; the remaining columns up to a multiple of 16 are ones.
__global__ void __launch_bounds__(256) frequency_fw_kernel(const float *__restrict__ x, int n_freq, int width,
                                                           int64_t n, const int32_t *__restrict__ n_dev,
                                                           __half *__restrict__ out, int out_stride, float x_min,
                                                           float x_extent) {
    n = b2n_eff_n(n, n_dev);
    const int enc = 3 * n_freq * 2;
    const int64_t total = n * width;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / width;
        const int j = (int)(e - i * width);
        float v = 1.0f;
        if (j < enc) {
            const int d = j / (2 * n_freq), f = (j >> 1) % n_freq;
            // box normalisation of NGP.density (networks.py:96) folded in: (x - xyz_min) / (xyz_max - xyz_min)
            const float xs = scalbnf(__fdiv_rn(__fsub_rn(__ldg(x + 3 * i + d), x_min), x_extent), f);
            v = sinf(__fadd_rn(__fmul_rn(xs, 3.14159265358979323846f), (j & 1) ? 1.57079632679489661923f : 0.0f));
        }
        out[i * out_stride + j] = __float2half_rn(v);
    }
}

extern "C" int b2n_frequency_fw(const float *x, int n_frequencies, int64_t n, const int32_t *n_dev,
                                b2n_half *out, int out_stride, float x_min, float x_extent, void *stream) {
    const int width = (3 * n_frequencies * 2 + 15) / 16 * 16;
    B2N_CHECK_ARG(n_frequencies >= 1 && out_stride >= width && x_extent != 0.f, "bad frequency config");
    if (n <= 0) return 0;
    frequency_fw_kernel<<<b2n_grid(b2n_blocks(n * width, 256), 8), 256, 0, (cudaStream_t)stream>>>(
        x, n_frequencies, width, n, n_dev, (__half *)out, out_stride, x_min, x_extent);
    B2N_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Real spherical harmonics up to degree 4 (16 coefficients) of a unit vector.
__device__ __forceinline__ void sh4_eval(float x, float y, float z, float *o) {
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    o[0] = 0.28209479177387814f;
    o[1] = -0.48860251190291987f * y;
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    o[10] = 2.8906114426405538f * xy * z;
    o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    o[14] = 1.4453057213202769f * z * (x2 - y2);
    o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

__global__ void __launch_bounds__(256) sh4_fw_kernel(const float *__restrict__ d, int normalize, int64_t n,
                                                     const int32_t *__restrict__ n_dev,
                                                     __half *__restrict__ out, int out_stride) {
    n = b2n_eff_n(n, n_dev);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float x = __ldg(d + 3 * i), y = __ldg(d + 3 * i + 1), z = __ldg(d + 3 * i + 2);
        if (normalize) {
            const float inv = 1.0f / sqrtf(x * x + y * y + z * z);
            x *= inv; y *= inv; z *= inv;
        } else {  // tcnn convention: input in [0,1] -> [-1,1]
            x = x * 2.0f - 1.0f; y = y * 2.0f - 1.0f; z = z * 2.0f - 1.0f;
        }
        float o[16];
        sh4_eval(x, y, z, o);
        __half2 *row = reinterpret_cast<__half2 *>(out + i * out_stride);
        #pragma unroll
        for (int k = 0; k < 8; ++k) row[k] = __floats2half2_rn(o[2 * k], o[2 * k + 1]);
    }
}

extern "C" int b2n_sh4_fw(const float *d, int normalize, int64_t n, const int32_t *n_dev, b2n_half *out,
                          int out_stride, void *stream) {
    B2N_CHECK_ARG(out_stride >= 16 && out_stride % 2 == 0, "out_stride must be even and >= 16");
    if (n <= 0) return 0;
    sh4_fw_kernel<<<b2n_grid(b2n_blocks(n, 256), 8), 256, 0, (cudaStream_t)stream>>>(
        d, normalize, n, n_dev, (__half *)out, out_stride);
    B2N_LAUNCH_CHECK();
    return 0;
}
