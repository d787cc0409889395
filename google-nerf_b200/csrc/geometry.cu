// geometry.cu -- ray/AABB, ray/sphere, Morton codes, packbits.
// Compiled with -fmad=false: the numerics contract for bit-exact geometry is one IEEE fp32 operation per
// source operation, no fused multiply-add (DESIGN.md "Numerics"); the oracle is built the same way.
// Replaces vren.ray_aabb_intersect / ray_sphere_intersect / morton3D / morton3D_invert / packbits
// (ngp_pl/models/custom_functions.py:29,52; ngp_pl/models/networks.py:128,147,153,251-252).
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void b2n_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
extern "C" const char *b2n_last_error(void) { return g_err; }
extern "C" int b2n_version(void) { return B2N_VERSION; }

// ------------------------------------------------------------------------------------------------
// One thread per ray; boxes/spheres are looped (render() uses exactly one box, max_hits = 1, so this is
// launch-latency bound: 32 B/ray of HBM traffic).  Hits are kept sorted near-to-far by stable insertion.
__device__ __forceinline__ void insert_hit(float t1, float t2, int64_t v, int max_hits, int &cnt,
                                           float *ht, int64_t *hv) {
    int pos = cnt < max_hits ? cnt : max_hits;
    while (pos > 0 && ht[2 * (pos - 1)] > t1) --pos;
    if (pos >= max_hits) return;
    int last = cnt < max_hits ? cnt : max_hits - 1;
    for (int k = last; k > pos; --k) {
        ht[2 * k] = ht[2 * (k - 1)];
        ht[2 * k + 1] = ht[2 * (k - 1) + 1];
        hv[k] = hv[k - 1];
    }
    ht[2 * pos] = t1;
    ht[2 * pos + 1] = t2;
    hv[pos] = v;
    if (cnt < max_hits) ++cnt;
}

__global__ void __launch_bounds__(256) ray_aabb_kernel(const float *__restrict__ rays_o,
                                                       const float *__restrict__ rays_d,
                                                       const float *__restrict__ centers,
                                                       const float *__restrict__ half_sizes,
                                                       int64_t n_rays, int64_t n_voxels, int max_hits,
                                                       int32_t *__restrict__ hits_cnt,
                                                       float *__restrict__ hits_t,
                                                       int64_t *__restrict__ hits_idx) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    float *ht = hits_t + r * max_hits * 2;
    int64_t *hv = hits_idx + r * max_hits;
    for (int k = 0; k < max_hits; ++k) {
        ht[2 * k] = -1.0f;
        ht[2 * k + 1] = -1.0f;
        hv[k] = -1;
    }
    const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
    const float ix = 1.0f / rays_d[3 * r], iy = 1.0f / rays_d[3 * r + 1], iz = 1.0f / rays_d[3 * r + 2];
    int cnt = 0, total = 0;
    for (int64_t v = 0; v < n_voxels; ++v) {
        const float cx = __ldg(centers + 3 * v), cy = __ldg(centers + 3 * v + 1), cz = __ldg(centers + 3 * v + 2);
        const float hx = __ldg(half_sizes + 3 * v), hy = __ldg(half_sizes + 3 * v + 1), hz = __ldg(half_sizes + 3 * v + 2);
        const float ax = (cx - hx - ox) * ix, bx = (cx + hx - ox) * ix;
        const float ay = (cy - hy - oy) * iy, by = (cy + hy - oy) * iy;
        const float az = (cz - hz - oz) * iz, bz = (cz + hz - oz) * iz;
        const float t1 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
        const float t2 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
        if (t1 > t2) continue;
        if (t2 > 0.0f) {
            ++total;
            insert_hit(fmaxf(t1, 0.0f), t2, v, max_hits, cnt, ht, hv);
        }
    }
    hits_cnt[r] = total;
}

__global__ void __launch_bounds__(256) ray_sphere_kernel(const float *__restrict__ rays_o,
                                                         const float *__restrict__ rays_d,
                                                         const float *__restrict__ centers,
                                                         const float *__restrict__ radii, int64_t n_rays,
                                                         int64_t n_spheres, int max_hits,
                                                         int32_t *__restrict__ hits_cnt,
                                                         float *__restrict__ hits_t,
                                                         int64_t *__restrict__ hits_idx) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    float *ht = hits_t + r * max_hits * 2;
    int64_t *hv = hits_idx + r * max_hits;
    for (int k = 0; k < max_hits; ++k) {
        ht[2 * k] = -1.0f;
        ht[2 * k + 1] = -1.0f;
        hv[k] = -1;
    }
    const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
    const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
    const float a = dx * dx + dy * dy + dz * dz;
    int cnt = 0, total = 0;
    for (int64_t s = 0; s < n_spheres; ++s) {
        const float px = ox - __ldg(centers + 3 * s), py = oy - __ldg(centers + 3 * s + 1),
                    pz = oz - __ldg(centers + 3 * s + 2);
        const float rad = __ldg(radii + s);
        const float hb = px * dx + py * dy + pz * dz;
        const float c = px * px + py * py + pz * pz - rad * rad;
        const float disc = hb * hb - a * c;
        if (disc < 0.0f) continue;
        const float sq = sqrtf(disc);
        const float t1 = (-hb - sq) / a, t2 = (-hb + sq) / a;
        if (t2 > 0.0f) {
            ++total;
            insert_hit(fmaxf(t1, 0.0f), t2, s, max_hits, cnt, ht, hv);
        }
    }
    hits_cnt[r] = total;
}

__global__ void clamp_near_kernel(float *hits_t, int64_t n_rays, float near_distance) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const float t1 = hits_t[2 * r];
    if (t1 >= 0.0f && t1 < near_distance) hits_t[2 * r] = near_distance;
}

extern "C" int b2n_ray_aabb_intersect(const float *rays_o, const float *rays_d, const float *centers,
                                      const float *half_sizes, int64_t n_rays, int64_t n_voxels,
                                      int max_hits, int32_t *hits_cnt, float *hits_t,
                                      int64_t *hits_voxel_idx, void *stream) {
    B2N_CHECK_ARG(max_hits >= 1 && n_rays >= 0 && n_voxels >= 0, "bad sizes");
    if (n_rays == 0) return 0;
    ray_aabb_kernel<<<b2n_blocks(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, centers, half_sizes, n_rays, n_voxels, max_hits, hits_cnt, hits_t, hits_voxel_idx);
    B2N_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2n_ray_sphere_intersect(const float *rays_o, const float *rays_d, const float *centers,
                                        const float *radii, int64_t n_rays, int64_t n_spheres,
                                        int max_hits, int32_t *hits_cnt, float *hits_t,
                                        int64_t *hits_sphere_idx, void *stream) {
    B2N_CHECK_ARG(max_hits >= 1 && n_rays >= 0 && n_spheres >= 0, "bad sizes");
    if (n_rays == 0) return 0;
    ray_sphere_kernel<<<b2n_blocks(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, centers, radii, n_rays, n_spheres, max_hits, hits_cnt, hits_t, hits_sphere_idx);
    B2N_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2n_clamp_near(float *hits_t, int64_t n_rays, float near_distance, void *stream) {
    if (n_rays <= 0) return 0;
    clamp_near_kernel<<<b2n_blocks(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(hits_t, n_rays, near_distance);
    B2N_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------ ray generation
// get_rays (ngp_pl/datasets/ray_utils.py:152-175) for a batch drawn as (img_idxs, pix_idxs) (datasets/base.py:24-40):
// rays_d = directions[pix] @ c2w[img][:, :3]^T, rays_o = c2w[img][:, 3].  One thread per ray; no FMA contraction in
// this file, the three products are summed left to right.
__global__ void __launch_bounds__(256) rays_from_indices_kernel(const float *__restrict__ directions,
                                                                const float *__restrict__ poses,
                                                                const int64_t *__restrict__ img_idxs,
                                                                const int64_t *__restrict__ pix_idxs, int64_t n,
                                                                float *__restrict__ rays_o, float *__restrict__ rays_d) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *d = directions + 3 * pix_idxs[i];
    const float *c = poses + 12 * img_idxs[i];
    const float d0 = __ldg(d), d1 = __ldg(d + 1), d2 = __ldg(d + 2);
    #pragma unroll
    for (int a = 0; a < 3; ++a) {
        rays_d[3 * i + a] = d0 * __ldg(c + 4 * a) + d1 * __ldg(c + 4 * a + 1) + d2 * __ldg(c + 4 * a + 2);
        rays_o[3 * i + a] = __ldg(c + 4 * a + 3);
    }
}

extern "C" int b2n_rays_from_indices(const float *directions, const float *poses, const int64_t *img_idxs,
                                     const int64_t *pix_idxs, int64_t n_rays, float *rays_o, float *rays_d,
                                     void *stream) {
    if (n_rays <= 0) return 0;
    rays_from_indices_kernel<<<b2n_blocks(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(
        directions, poses, img_idxs, pix_idxs, n_rays, rays_o, rays_d);
    B2N_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void morton3D_kernel(const int32_t *__restrict__ coords, int64_t n, int32_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = (int32_t)b2n_morton3D((uint32_t)coords[3 * i], (uint32_t)coords[3 * i + 1], (uint32_t)coords[3 * i + 2]);
}
__global__ void morton3D_invert_kernel(const int32_t *__restrict__ idx, int64_t n, int32_t *__restrict__ coords) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = (uint32_t)idx[i];
    coords[3 * i + 0] = (int32_t)b2n_compact_bits(v);
    coords[3 * i + 1] = (int32_t)b2n_compact_bits(v >> 1);
    coords[3 * i + 2] = (int32_t)b2n_compact_bits(v >> 2);
}

extern "C" int b2n_morton3D(const int32_t *coords, int64_t n, int32_t *indices, void *stream) {
    if (n <= 0) return 0;
    morton3D_kernel<<<b2n_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(coords, n, indices);
    B2N_LAUNCH_CHECK();
    return 0;
}
extern "C" int b2n_morton3D_invert(const int32_t *indices, int64_t n, int32_t *coords, void *stream) {
    if (n <= 0) return 0;
    morton3D_invert_kernel<<<b2n_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(indices, n, coords);
    B2N_LAUNCH_CHECK();
    return 0;
}

// packbits: one thread per output byte reads 32 B (two float4) and writes 1 B; HBM-bound, 33 B/byte.
__global__ void __launch_bounds__(256) packbits_kernel(const float4 *__restrict__ grid, int64_t n_bytes,
                                                       float threshold, const float *__restrict__ threshold_dev,
                                                       uint8_t *__restrict__ bitfield) {
    const float thr = threshold_dev ? __ldg(threshold_dev) : threshold;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < n_bytes;
         n += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = __ldg(grid + 2 * n), b = __ldg(grid + 2 * n + 1);
        uint32_t bits = 0;
        bits |= (a.x > thr) ? 1u : 0u;
        bits |= (a.y > thr) ? 2u : 0u;
        bits |= (a.z > thr) ? 4u : 0u;
        bits |= (a.w > thr) ? 8u : 0u;
        bits |= (b.x > thr) ? 16u : 0u;
        bits |= (b.y > thr) ? 32u : 0u;
        bits |= (b.z > thr) ? 64u : 0u;
        bits |= (b.w > thr) ? 128u : 0u;
        bitfield[n] = (uint8_t)bits;
    }
}

extern "C" int b2n_packbits(const float *density_grid, int64_t n_bytes, float threshold,
                            const float *threshold_dev, uint8_t *density_bitfield, void *stream) {
    if (n_bytes <= 0) return 0;
    B2N_CHECK_ARG(((uintptr_t)density_grid & 15) == 0, "density_grid must be 16-byte aligned");
    packbits_kernel<<<b2n_grid(b2n_blocks(n_bytes, 256), 8), 256, 0, (cudaStream_t)stream>>>(
        (const float4 *)density_grid, n_bytes, threshold, threshold_dev, density_bitfield);
    B2N_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// NGP.mark_invisible_cells (ngp_pl/models/networks.py:159-214) as one launch: a thread per (cascade, cell) projects
// the cell centre into every training camera; the cell is valid (density 0) when at least one camera sees it at
// depth >= near and no camera has it in its image closer than near, otherwise it gets -1 and is never updated.
// The reference does this with chunked bmm over (N_img, 3, 64^3) tensors.
__global__ void __launch_bounds__(256) mark_invisible_cells_kernel(const float *__restrict__ K,
                                                                   const float *__restrict__ poses, int n_img,
                                                                   float img_w, float img_h, float near_d, int G,
                                                                   int cascades, float scale,
                                                                   float *__restrict__ density_grid) {
    const int64_t cells = (int64_t)G * G * G;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cells * cascades) return;
    const int c = (int)(t / cells);
    const int64_t cell = t - (int64_t)c * cells;
    const uint32_t cx = (uint32_t)(cell % G), cy = (uint32_t)((cell / G) % G), cz = (uint32_t)(cell / ((int64_t)G * G));
    const float s = fminf(scalbnf(1.0f, c - 1), scale);           // min(2^(c-1), scale)
    const float span = s - s / (float)G;
    const float gm1 = (float)(G - 1);
    const float xw = ((float)cx / gm1 * 2.0f - 1.0f) * span, yw = ((float)cy / gm1 * 2.0f - 1.0f) * span,
                zw = ((float)cz / gm1 * 2.0f - 1.0f) * span;
    const float k00 = K[0], k01 = K[1], k02 = K[2], k10 = K[3], k11 = K[4], k12 = K[5], k20 = K[6], k21 = K[7], k22 = K[8];
    bool covered = false, too_near = false;
    for (int i = 0; i < n_img && !too_near; ++i) {
        const float *P = poses + 12 * i;                            // (3,4) row-major camera-to-world
        // world -> camera: R^T (x - t) written like the reference, R^T x + (-R^T t)
        float pc[3];
        #pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float tr = -(P[r] * P[3] + P[4 + r] * P[7] + P[8 + r] * P[11]);
            pc[r] = P[r] * xw + P[4 + r] * yw + P[8 + r] * zw + tr;
        }
        const float u3 = k00 * pc[0] + k01 * pc[1] + k02 * pc[2], v3 = k10 * pc[0] + k11 * pc[1] + k12 * pc[2],
                    d = k20 * pc[0] + k21 * pc[1] + k22 * pc[2];
        const float u = u3 / d, v = v3 / d;
        const bool in_image = (d >= 0.0f) && (u >= 0.0f) && (u < img_w) && (v >= 0.0f) && (v < img_h);
        covered |= in_image && (d >= near_d);
        too_near |= in_image && (d < near_d);
    }
    density_grid[(int64_t)c * cells + b2n_morton3D(cx, cy, cz)] = (covered && !too_near) ? 0.0f : -1.0f;
}

extern "C" int b2n_mark_invisible_cells(const float *K, const float *poses, int n_img, int img_w, int img_h,
                                        float near_distance, int grid_size, int cascades, float scale,
                                        float *density_grid, void *stream) {
    B2N_CHECK_ARG(K && poses && density_grid && n_img >= 1 && grid_size >= 2 && grid_size <= 1024 && cascades >= 1,
                  "bad arguments");
    const int64_t n = (int64_t)grid_size * grid_size * grid_size * cascades;
    mark_invisible_cells_kernel<<<b2n_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(
        K, poses, n_img, (float)img_w, (float)img_h, near_distance, grid_size, cascades, scale, density_grid);
    B2N_LAUNCH_CHECK();
    return 0;
}
