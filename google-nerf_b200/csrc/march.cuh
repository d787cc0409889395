// march.cuh -- the DDA loop body of the occupancy-bitfield marcher, shared by march.cu (train / test marchers) and
// render_tc.cu (whole rays in one kernel).  Every floating-point operation is spelled as ONE IEEE binary32 operation
// (__fmul_rn / __fadd_rn / __fsub_rn are never contracted into FMAs), so sample positions are bit-identical whatever
// -fmad setting the including translation unit is compiled with (oracle/SPEC.md sections 1-4).
#pragma once
#include "common.cuh"

#define SQRT3 1.73205080757f
#define MAX_MIPS 16

struct MarchParams {
    const uint8_t *bitfield;
    int cascades, grid_size, max_samples;
    float scale, esf, dt_lo, dt_hi, dt0, g_inv;  // dt0 = calc_dt for esf == 0 (constant step)
    uint32_t g3;
    float mip_bound[MAX_MIPS], mip_bound_inv[MAX_MIPS];   // min(2^(mip-1), scale) and its IEEE reciprocal
};

__device__ __forceinline__ float calc_dt(float t, const MarchParams &p) {
    return fminf(p.dt_hi, fmaxf(p.dt_lo, __fmul_rn(t, p.esf)));
}

struct Ray {
    float ox, oy, oz, dx, dy, dz, ix, iy, iz;
};

// far face of cell n along one axis, as a ray parameter:  (((n + 0.5 + 0.5 sign(d)) / G * 2 - 1) * bound - x) / d
__device__ __forceinline__ float axis_exit(int n, float d, float inv_d, float x, float g_inv, float bound) {
    const float cell = __fadd_rn(__fadd_rn((float)n, 0.5f), __fmul_rn(0.5f, copysignf(1.0f, d)));
    const float face = __fsub_rn(__fmul_rn(__fmul_rn(cell, g_inv), 2.0f), 1.0f);
    return __fmul_rn(__fsub_rn(__fmul_rn(face, bound), x), inv_d);
}
__device__ __forceinline__ int axis_cell(float x, float bound_inv, float G) {
    const float u = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(__fmul_rn(x, bound_inv), 1.0f)), G);
    return __float2int_rz(fminf(G - 1.0f, fmaxf(0.0f, u)));
}

// One DDA loop body at parameter t, in two halves so that a caller can put other work between the address and the
// use of the occupancy byte: probe_cell = position, step, bitfield index of the cell and (for an empty cell) the skip
// target; cell_occupied = the bit.  probe = both.
__device__ __forceinline__ uint32_t probe_cell(const Ray &r, float t, const MarchParams &p, float &dt, float &x,
                                               float &y, float &z, float &t_target) {
    const float G = (float)p.grid_size;
    x = __fadd_rn(r.ox, __fmul_rn(t, r.dx));
    y = __fadd_rn(r.oy, __fmul_rn(t, r.dy));
    z = __fadd_rn(r.oz, __fmul_rn(t, r.dz));
    dt = calc_dt(t, p);
    int mip = 0;
    if (p.cascades > 1) {
        int e;
        frexpf(fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z))), &e);
        mip = min(p.cascades - 1, max(0, e + 1));
        frexpf(__fmul_rn(dt, G), &e);
        mip = max(mip, min(p.cascades - 1, max(0, e)));
    }
    const float mip_bound = p.mip_bound[mip];
    const float mip_bound_inv = p.mip_bound_inv[mip];
    const int nx = axis_cell(x, mip_bound_inv, G), ny = axis_cell(y, mip_bound_inv, G), nz = axis_cell(z, mip_bound_inv, G);
    const float tx = axis_exit(nx, r.dx, r.ix, x, p.g_inv, mip_bound);
    const float ty = axis_exit(ny, r.dy, r.iy, y, p.g_inv, mip_bound);
    const float tz = axis_exit(nz, r.dz, r.iz, z, p.g_inv, mip_bound);
    t_target = __fadd_rn(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
    return (uint32_t)mip * p.g3 + b2n_morton3D(nx, ny, nz);
}
__device__ __forceinline__ bool cell_occupied(const MarchParams &p, uint32_t idx) {
    return (__ldg(p.bitfield + (idx >> 3)) >> (idx & 7)) & 1;
}
__device__ __forceinline__ bool probe(const Ray &r, float t, const MarchParams &p, float &dt, float &x,
                                      float &y, float &z, float &t_target) {
    const uint32_t idx = probe_cell(r, t, p, dt, x, y, z, t_target);
    return cell_occupied(p, idx);
}

__device__ __forceinline__ Ray load_ray(const float *rays_o, const float *rays_d, int64_t r) {
    Ray q;
    q.ox = __ldg(rays_o + 3 * r); q.oy = __ldg(rays_o + 3 * r + 1); q.oz = __ldg(rays_o + 3 * r + 2);
    q.dx = __ldg(rays_d + 3 * r); q.dy = __ldg(rays_d + 3 * r + 1); q.dz = __ldg(rays_d + 3 * r + 2);
    q.ix = 1.0f / q.dx; q.iy = 1.0f / q.dy; q.iz = 1.0f / q.dz;
    return q;
}

static inline int fill_params(MarchParams &p, const uint8_t *bitfield, int cascades, float scale, float esf,
                              int grid_size, int max_samples) {
    B2N_CHECK_ARG(cascades >= 1 && cascades <= MAX_MIPS && grid_size >= 1 && grid_size <= 1024 && max_samples >= 1,
                  "bad marcher config");
    p.bitfield = bitfield; p.cascades = cascades; p.grid_size = grid_size; p.max_samples = max_samples;
    p.scale = scale; p.esf = esf;
    p.dt_lo = SQRT3 / max_samples;
    p.dt_hi = SQRT3 * 2 * scale / grid_size;
    p.dt0 = fminf(p.dt_hi, fmaxf(p.dt_lo, 0.0f));
    p.g_inv = 1.0f / grid_size;
    p.g3 = (uint32_t)grid_size * grid_size * grid_size;
    for (int m = 0; m < MAX_MIPS; ++m) {
        const float b = fminf(scalbnf(1.0f, m - 1), scale);   // host fp32 = IEEE, same values as the device would compute
        p.mip_bound[m] = b;
        p.mip_bound_inv[m] = 1.0f / b;
    }
    return 0;
}
