// hashgrid.cuh -- device code of the multiresolution hash grid shared by encoding.cu (stand-alone gather / scatter
// kernels) and render_tc.cu (the gather inside the whole-ray test-time renderer).
#pragma once
#include "common.cuh"

struct GridLevels {
    int n_levels;
    float scale[B2N_MAX_LEVELS];
    uint32_t resolution[B2N_MAX_LEVELS];
    uint32_t size[B2N_MAX_LEVELS];
    uint32_t offset[B2N_MAX_LEVELS];
    uint8_t mode[B2N_MAX_LEVELS];       // 0 dense (index < 2 * size), 1 hashed with a power-of-two size, 2 generic
    float x_offset, x_scale;
};

static inline int to_levels(const b2n_grid_layout *l, GridLevels &g) {
    B2N_CHECK_ARG(l != nullptr && l->n_features == 2 && l->n_levels >= 1 && l->n_levels <= B2N_MAX_LEVELS,
                  "hash grid needs n_features == 2 and 1..32 levels");
    g.n_levels = l->n_levels;
    g.x_offset = l->x_offset; g.x_scale = l->x_scale;
    for (int i = 0; i < l->n_levels; ++i) {
        g.scale[i] = l->scale[i]; g.resolution[i] = l->resolution[i];
        g.size[i] = l->size[i]; g.offset[i] = l->offset[i];
        // the stride walk of grid_index on the host: does this level fall through to the hash?
        const uint32_t res = l->resolution[i], size = l->size[i];
        uint32_t stride = 1;
        for (int d = 0; d < 3; ++d)
            if (stride <= size) stride *= res;
        const bool hashed = size < stride, pow2 = size && (size & (size - 1)) == 0;
        const bool dense_ok = !hashed && res >= 2 && (uint64_t)res * res * res <= size;   // index <= res+res^2+res^3 < 2*size
        g.mode[i] = hashed ? (pow2 ? 1 : 2) : (dense_ok ? 0 : 2);
    }
    return 0;
}

// entry index of one corner (tiny-cuda-nn grid_index: dense x + y*res + z*res^2 while it fits, otherwise the
// coherent prime hash), all in wrapping uint32 arithmetic
__device__ __forceinline__ uint32_t grid_index(uint32_t x, uint32_t y, uint32_t z, uint32_t res, uint32_t size) {
    uint32_t stride = 1, index = 0;
    // unrolled over the 3 dims with the early exit of the reference loop
    if (stride <= size) { index += x * stride; stride *= res; }
    if (stride <= size) { index += y * stride; stride *= res; }
    if (stride <= size) { index += z * stride; stride *= res; }
    if (size < stride) index = x ^ (y * 2654435761u) ^ (z * 805459861u);
    return index % size;
}

// the same index without the division (mode from to_levels)
__device__ __forceinline__ uint32_t grid_index_m(uint32_t x, uint32_t y, uint32_t z, uint32_t res, uint32_t size, int mode) {
    if (mode == 1) return (x ^ (y * 2654435761u) ^ (z * 805459861u)) & (size - 1);
    if (mode == 0) {
        uint32_t index = x + (y + z * res) * res;
        if (index >= size) {                          // only the x/y/z = res boundary corners; positions inside the box
            index -= size;                            // overshoot by less than one period
            if (index >= size) index %= size;         // out-of-box positions stay in bounds like the generic form
        }
        return index;
    }
    return grid_index(x, y, z, res, size);
}

struct Corner8 {
    uint32_t idx[8];
    float w[8];
};

__device__ __forceinline__ void level_corners(float px, float py, float pz, float scale, uint32_t res,
                                              uint32_t size, uint32_t offset, int mode, Corner8 &c) {
    const float fx = fmaf(scale, px, 0.5f), fy = fmaf(scale, py, 0.5f), fz = fmaf(scale, pz, 0.5f);
    const float gx = floorf(fx), gy = floorf(fy), gz = floorf(fz);
    const float wx = fx - gx, wy = fy - gy, wz = fz - gz;
    const uint32_t x0 = (uint32_t)gx, y0 = (uint32_t)gy, z0 = (uint32_t)gz;
    #pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t x = x0 + (k & 1), y = y0 + ((k >> 1) & 1), z = z0 + ((k >> 2) & 1);
        float w = 1.0f;
        w *= (k & 1) ? wx : 1.0f - wx;
        w *= (k & 2) ? wy : 1.0f - wy;
        w *= (k & 4) ? wz : 1.0f - wz;
        c.idx[k] = offset + grid_index_m(x, y, z, res, size, mode);
        c.w[k] = w;
    }
}

// Half a cell: the four (y, z) corners at x0 + cx.  Two adjacent lanes (cx = lane & 1) share a sample, so their
// gathers fall into one sector; the halves are combined with one xor-shuffle (hashgrid_fw_kernel, render_tc.cu).
struct Corner4 {
    uint32_t idx[4];
    float w[4];
    uint32_t key;     // packed integer lattice position of the cell (for run detection in the backward pass)
};

__device__ __forceinline__ void level_corners4(float px, float py, float pz, float scale, uint32_t res, uint32_t size,
                                               uint32_t offset, int mode, int cx, Corner4 &c) {
    const float fx = fmaf(scale, px, 0.5f), fy = fmaf(scale, py, 0.5f), fz = fmaf(scale, pz, 0.5f);
    const float gx = floorf(fx), gy = floorf(fy), gz = floorf(fz);
    const float wx = fx - gx, wy = fy - gy, wz = fz - gz;
    const uint32_t x0 = (uint32_t)gx, y0 = (uint32_t)gy, z0 = (uint32_t)gz;
    c.key = x0 | (y0 << 10) | (z0 << 20);
    const uint32_t x = x0 + (uint32_t)cx;
    const float wxc = cx ? wx : 1.0f - wx;
    #pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t y = y0 + (k & 1), z = z0 + ((k >> 1) & 1);
        float w = 1.0f;
        w *= wxc;                                  // same product order as the 8-corner form: ((1*wx)*wy)*wz
        w *= (k & 1) ? wy : 1.0f - wy;
        w *= (k & 2) ? wz : 1.0f - wz;
        c.idx[k] = offset + grid_index_m(x, y, z, res, size, mode);
        c.w[k] = w;
    }
}

// The two x-corners of a cell: entries i0 (x0) and i1 (x0 + 1).  Whenever they form an aligned pair {2m, 2m + 1} -- dense
// levels with an even entry index, hashed levels with an even x0 (the prime of the x axis is 1, so x0 and x0 + 1 then
// differ in bit 0 only) -- both go out as ONE 16-byte red.global.add.v4.f32: the scatter is bound by the number of L2
// requests, not by their size.
__device__ __forceinline__ void red_pair(float2 *__restrict__ grad_table, uint32_t i0, uint32_t i1, float ax, float ay,
                                         float bx, float by) {
#ifndef HG_BW_NO_V4
    if ((i0 ^ i1) == 1u) {
        const bool swap = i0 & 1u;                   // i1 is the even one
        atomicAdd(reinterpret_cast<float4 *>(grad_table + (i0 & ~1u)),
                  swap ? make_float4(bx, by, ax, ay) : make_float4(ax, ay, bx, by));
        return;
    }
#endif
    if (ax != 0.0f || ay != 0.0f) atomicAdd(grad_table + i0, make_float2(ax, ay));
    if (bx != 0.0f || by != 0.0f) atomicAdd(grad_table + i1, make_float2(bx, by));
}
__device__ __forceinline__ void scatter8(float2 *__restrict__ grad_table, const Corner8 &c, float gx, float gy) {
    #pragma unroll
    for (int j = 0; j < 4; ++j)
        red_pair(grad_table, c.idx[2 * j], c.idx[2 * j + 1], gx * c.w[2 * j], gy * c.w[2 * j], gx * c.w[2 * j + 1],
                 gy * c.w[2 * j + 1]);
}

#ifndef AGG_RES_DEF
#define AGG_RES_DEF 200      // resolution up to which runs of lanes in one cell are pre-summed (measured 128 / 200 / 264 / 350 / 460 -> 109 / 102.6 / 102.8 / 104.5 / 108.4 us)
#endif

// One level of the table-gradient scatter for the 32 lanes of a warp, lane = one sample (gx, gy = its dL/d(feature
// pair) of level l, already scaled).  The lanes of a warp hold CONSECUTIVE packed samples, i.e. neighbours on a ray
// (0.0017 apart), so on the coarse levels most lanes fall into the same cell and would hit the same 8 table entries:
// L2 serialises atomics per address.  For levels whose cells are wide enough (resolution <= AGG_RES_DEF) runs of lanes
// with the same cell are summed with a segmented shuffle reduction and only the run's head lane issues the reds; fine
// levels (one sample per cell) go straight to red.global.add.  Must be called by all 32 lanes (live = false for idle ones).
__device__ __forceinline__ void scatter_level(float px, float py, float pz, float gx, float gy, int l, const GridLevels &g,
                                              float2 *__restrict__ grad_table, bool live, int lane) {
    const uint32_t FULLM = 0xffffffffu;
    const uint32_t res = g.resolution[l];
    Corner8 c;
    level_corners(px, py, pz, g.scale[l], res, g.size[l], g.offset[l], g.mode[l], c);
    if (!live) { gx = 0.f; gy = 0.f; }
    if (res > (uint32_t)AGG_RES_DEF) {
        if (gx != 0.0f || gy != 0.0f) scatter8(grad_table, c, gx, gy);
        return;
    }
    // cell key: the integer lattice position (10 bits per axis is enough for res <= 1023)
    const float s = g.scale[l];
    const uint32_t kx = (uint32_t)floorf(fmaf(s, px, 0.5f)), ky = (uint32_t)floorf(fmaf(s, py, 0.5f)),
                   kz = (uint32_t)floorf(fmaf(s, pz, 0.5f));
    const uint32_t key = live ? (kx | (ky << 10) | (kz << 20)) : (0xC0000000u | (uint32_t)lane);
    const uint32_t prev = __shfl_up_sync(FULLM, key, 1);
    const bool head = (lane == 0) || (key != prev);
    const uint32_t heads = __ballot_sync(FULLM, head);
    const uint32_t above = (lane == 31) ? 0u : (heads & ~((2u << lane) - 1u));   // heads at positions > lane
    const int seg_end = above ? (__ffs(above) - 2) : 31;
    float vx[8], vy[8];
    #pragma unroll
    for (int k = 0; k < 8; ++k) { vx[k] = gx * c.w[k]; vy[k] = gy * c.w[k]; }
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const bool take = lane + d <= seg_end;
        #pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float ox = __shfl_down_sync(FULLM, vx[k], d), oy = __shfl_down_sync(FULLM, vy[k], d);
            if (take) { vx[k] += ox; vy[k] += oy; }
        }
    }
    if (head && live) {
        #pragma unroll
        for (int j = 0; j < 4; ++j)
            red_pair(grad_table, c.idx[2 * j], c.idx[2 * j + 1], vx[2 * j], vy[2 * j], vx[2 * j + 1], vy[2 * j + 1]);
    }
}
