/*
 * oracle/c/vren_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Serial, scalar C restatement of the arithmetic behind the ten `vren` entry
 * points that ngp_pl calls (reference call sites:
 *   ngp_pl/models/custom_functions.py:29,52,86-90,140-142,153-158
 *   ngp_pl/models/rendering.py:79-83,97-100
 *   ngp_pl/models/networks.py:128,147,153,251-252).
 *
 * PARITY UNPINNED: the reference repository git-ignores models/csrc (its
 * .gitignore:23-25) and ships no tests, so there is no golden vector to pin
 * this file against.  The arithmetic restated here is the published upstream
 * kwea123/ngp_pl `vren` algorithm as recorded in SURVEY.md Appendix A, checked
 * against every shape / sentinel / in-place contract visible at the call
 * sites above.  Two independent restatements exist (this serial C one and the
 * vectorised torch one in oracle/ python files); the tests under tests/ cross-checks
 * them bit for bit.
 *
 * Floating-point contract (DESIGN.md "Numerics"): every float operation is a
 * single IEEE-754 binary32 operation with round-to-nearest-even and NO fused
 * multiply-add.  Build with -ffp-contract=off (oracle/Makefile does).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define SQRT3 1.73205080757f

static inline float clampf(float x, float lo, float hi) {
    return fminf(hi, fmaxf(lo, x));
}
static inline float signf(float x) { return copysignf(1.0f, x); }

/* ---- Morton codes (networks.py:128,147,153) ------------------ */
static inline uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3D_1(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
static inline uint32_t morton3D_invert_1(uint32_t x) {
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

void orc_morton3D(const int32_t *coords, int64_t n, int32_t *out) {
    for (int64_t i = 0; i < n; ++i)
        out[i] = (int32_t)morton3D_1((uint32_t)coords[3 * i], (uint32_t)coords[3 * i + 1],
                                     (uint32_t)coords[3 * i + 2]);
}
void orc_morton3D_invert(const int32_t *idx, int64_t n, int32_t *coords) {
    for (int64_t i = 0; i < n; ++i) {
        uint32_t v = (uint32_t)idx[i];
        coords[3 * i + 0] = (int32_t)morton3D_invert_1(v >> 0);
        coords[3 * i + 1] = (int32_t)morton3D_invert_1(v >> 1);
        coords[3 * i + 2] = (int32_t)morton3D_invert_1(v >> 2);
    }
}

/* ---- packbits (networks.py:251-252) -------------------------- */
void orc_packbits(const float *grid, int64_t n_bytes, float thr, uint8_t *bitfield) {
    for (int64_t n = 0; n < n_bytes; ++n) {
        uint8_t bits = 0;
        for (int i = 0; i < 8; ++i)
            bits |= (grid[8 * n + i] > thr) ? ((uint8_t)1 << i) : 0;
        bitfield[n] = bits;
    }
}

/* ---- ray / AABB and ray / sphere (custom_functions.py:8-52) --
 * Output order follows the docstring at custom_functions.py:20-24: hits sorted
 * near to far, unused slots = -1.  (Only max_hits=1, one box is exercised by
 * render(), rendering.py:27-28.)                                            */
static void insert_hit(float t1, float t2, int64_t v, int max_hits, int *cnt, float *ht,
                       int64_t *hv) {
    /* stable insertion by t1; keeps the nearest max_hits */
    int pos = *cnt < max_hits ? *cnt : max_hits;
    while (pos > 0 && ht[2 * (pos - 1)] > t1) --pos;
    if (pos >= max_hits) return;
    int last = (*cnt < max_hits ? *cnt : max_hits - 1);
    for (int k = last; k > pos; --k) {
        ht[2 * k] = ht[2 * (k - 1)];
        ht[2 * k + 1] = ht[2 * (k - 1) + 1];
        hv[k] = hv[k - 1];
    }
    ht[2 * pos] = t1;
    ht[2 * pos + 1] = t2;
    hv[pos] = v;
    if (*cnt < max_hits) ++*cnt;
}

void orc_ray_aabb_intersect(const float *rays_o, const float *rays_d, const float *centers,
                            const float *half_sizes, int64_t n_rays, int64_t n_voxels,
                            int max_hits, int32_t *hits_cnt, float *hits_t,
                            int64_t *hits_voxel_idx) {
    for (int64_t r = 0; r < n_rays; ++r) {
        float *ht = hits_t + r * max_hits * 2;
        int64_t *hv = hits_voxel_idx + r * max_hits;
        for (int k = 0; k < max_hits; ++k) { ht[2 * k] = ht[2 * k + 1] = -1.0f; hv[k] = -1; }
        int cnt = 0, total = 0;
        const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
        const float ix = 1.0f / rays_d[3 * r], iy = 1.0f / rays_d[3 * r + 1],
                    iz = 1.0f / rays_d[3 * r + 2];
        for (int64_t v = 0; v < n_voxels; ++v) {
            const float *c = centers + 3 * v, *h = half_sizes + 3 * v;
            const float ax = (c[0] - h[0] - ox) * ix, bx = (c[0] + h[0] - ox) * ix;
            const float ay = (c[1] - h[1] - oy) * iy, by = (c[1] + h[1] - oy) * iy;
            const float az = (c[2] - h[2] - oz) * iz, bz = (c[2] + h[2] - oz) * iz;
            const float t1 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
            const float t2 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
            if (t1 > t2) continue; /* no intersection */
            if (t2 > 0.0f) {
                ++total;
                insert_hit(fmaxf(t1, 0.0f), t2, v, max_hits, &cnt, ht, hv);
            }
        }
        hits_cnt[r] = total;
    }
}

void orc_ray_sphere_intersect(const float *rays_o, const float *rays_d, const float *centers,
                              const float *radii, int64_t n_rays, int64_t n_spheres,
                              int max_hits, int32_t *hits_cnt, float *hits_t,
                              int64_t *hits_sphere_idx) {
    for (int64_t r = 0; r < n_rays; ++r) {
        float *ht = hits_t + r * max_hits * 2;
        int64_t *hv = hits_sphere_idx + r * max_hits;
        for (int k = 0; k < max_hits; ++k) { ht[2 * k] = ht[2 * k + 1] = -1.0f; hv[k] = -1; }
        int cnt = 0, total = 0;
        const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
        const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
        const float a = dx * dx + dy * dy + dz * dz;
        for (int64_t s = 0; s < n_spheres; ++s) {
            const float px = ox - centers[3 * s], py = oy - centers[3 * s + 1],
                        pz = oz - centers[3 * s + 2];
            const float hb = px * dx + py * dy + pz * dz; /* half b */
            const float c = px * px + py * py + pz * pz - radii[s] * radii[s];
            const float disc = hb * hb - a * c;
            if (disc < 0.0f) continue;
            const float sq = sqrtf(disc);
            const float t1 = (-hb - sq) / a, t2 = (-hb + sq) / a;
            if (t2 > 0.0f) {
                ++total;
                insert_hit(fmaxf(t1, 0.0f), t2, s, max_hits, &cnt, ht, hv);
            }
        }
        hits_cnt[r] = total;
    }
}

/* ---- marcher helpers ---------------------------------------- */
static inline float calc_dt(float t, float exp_step_factor, int max_samples, int grid_size,
                            float scale) {
    return clampf(t * exp_step_factor, SQRT3 / max_samples, SQRT3 * 2 * scale / grid_size);
}
static inline int mip_from_pos(float x, float y, float z, int cascades) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int e;
    frexpf(mx, &e);
    int m = e + 1;
    if (m < 0) m = 0;
    if (m > cascades - 1) m = cascades - 1;
    return m;
}
static inline int mip_from_dt(float dt, int grid_size, int cascades) {
    int e;
    frexpf(dt * grid_size, &e);
    int m = e;
    if (m < 0) m = 0;
    if (m > cascades - 1) m = cascades - 1;
    return m;
}

typedef struct {
    float ox, oy, oz, dx, dy, dz, dx_inv, dy_inv, dz_inv;
    const uint8_t *bitfield;
    int cascades, grid_size, max_samples;
    float scale, exp_step_factor;
} march_ctx;

/* One loop body of the DDA: test the cell at parameter t.
 * Returns 1 and *dt_out if occupied; otherwise returns 0 and *t_target. */
static inline int probe(const march_ctx *m, float t, float *dt_out, float *x, float *y,
                        float *z, float *t_target) {
    const int G = m->grid_size;
    const uint32_t G3 = (uint32_t)G * G * G;
    const float G_inv = 1.0f / G;
    *x = m->ox + t * m->dx;
    *y = m->oy + t * m->dy;
    *z = m->oz + t * m->dz;
    const float dt = calc_dt(t, m->exp_step_factor, m->max_samples, G, m->scale);
    *dt_out = dt;
    int mip = mip_from_pos(*x, *y, *z, m->cascades);
    const int mip2 = mip_from_dt(dt, G, m->cascades);
    if (mip2 > mip) mip = mip2;
    const float mip_bound = fminf(scalbnf(1.0f, mip - 1), m->scale);
    const float mip_bound_inv = 1.0f / mip_bound;
    const int nx = (int)clampf(0.5f * (*x * mip_bound_inv + 1) * G, 0.0f, G - 1.0f);
    const int ny = (int)clampf(0.5f * (*y * mip_bound_inv + 1) * G, 0.0f, G - 1.0f);
    const int nz = (int)clampf(0.5f * (*z * mip_bound_inv + 1) * G, 0.0f, G - 1.0f);
    const uint32_t idx = (uint32_t)mip * G3 + morton3D_1(nx, ny, nz);
    const int occ = m->bitfield[idx / 8] & (1 << (idx % 8));
    if (occ) return 1;
    const float tx = (((nx + 0.5f + 0.5f * signf(m->dx)) * G_inv * 2 - 1) * mip_bound - *x) * m->dx_inv;
    const float ty = (((ny + 0.5f + 0.5f * signf(m->dy)) * G_inv * 2 - 1) * mip_bound - *y) * m->dy_inv;
    const float tz = (((nz + 0.5f + 0.5f * signf(m->dz)) * G_inv * 2 - 1) * mip_bound - *z) * m->dz_inv;
    *t_target = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    return 0;
}

static void ctx_init(march_ctx *m, const float *o, const float *d, const uint8_t *bitfield,
                     int cascades, float scale, float esf, int grid_size, int max_samples) {
    m->ox = o[0]; m->oy = o[1]; m->oz = o[2];
    m->dx = d[0]; m->dy = d[1]; m->dz = d[2];
    m->dx_inv = 1.0f / d[0]; m->dy_inv = 1.0f / d[1]; m->dz_inv = 1.0f / d[2];
    m->bitfield = bitfield; m->cascades = cascades; m->scale = scale;
    m->exp_step_factor = esf; m->grid_size = grid_size; m->max_samples = max_samples;
}

/* ---- raymarching_train (custom_functions.py:78-101) ----------
 * Deterministic packing: ray r owns row r of rays_a and its samples start at
 * the exclusive prefix sum of the per-ray counts in ray order (DESIGN.md
 * "Sample order").  Pass xyzs==NULL to only count (fills rays_a, counter).
 * counter[0] = total samples, counter[1] = n_rays.                          */
void orc_raymarching_train(const float *rays_o, const float *rays_d, const float *hits_t,
                           const uint8_t *bitfield, int cascades, float scale, float esf,
                           const float *noise, int grid_size, int max_samples, int64_t n_rays,
                           int64_t *rays_a, float *xyzs, float *dirs, float *deltas, float *ts,
                           int32_t *counter) {
    int64_t start = 0;
    for (int64_t r = 0; r < n_rays; ++r) {
        march_ctx m;
        ctx_init(&m, rays_o + 3 * r, rays_d + 3 * r, bitfield, cascades, scale, esf, grid_size,
                 max_samples);
        float t1 = hits_t[2 * r], t2 = hits_t[2 * r + 1];
        if (t1 >= 0.0f) {
            const float dt = calc_dt(t1, esf, max_samples, grid_size, scale);
            t1 += dt * noise[r];
        }
        float t = t1;
        int n = 0;
        while (0.0f <= t && t < t2 && n < max_samples) {
            float dt, x, y, z, t_target;
            if (probe(&m, t, &dt, &x, &y, &z, &t_target)) {
                if (xyzs) {
                    const int64_t s = start + n;
                    xyzs[3 * s] = x; xyzs[3 * s + 1] = y; xyzs[3 * s + 2] = z;
                    dirs[3 * s] = m.dx; dirs[3 * s + 1] = m.dy; dirs[3 * s + 2] = m.dz;
                    ts[s] = t; deltas[s] = dt;
                }
                t += dt;
                ++n;
            } else {
                do {
                    t += calc_dt(t, esf, max_samples, grid_size, scale);
                } while (t < t_target);
            }
        }
        rays_a[3 * r] = r; rays_a[3 * r + 1] = start; rays_a[3 * r + 2] = n;
        start += n;
    }
    counter[0] = (int32_t)start;
    counter[1] = (int32_t)n_rays;
}

/* ---- raymarching_test (rendering.py:79-83) ------------------- */
void orc_raymarching_test(const float *rays_o, const float *rays_d, float *hits_t,
                          const int64_t *alive_indices, const uint8_t *bitfield, int cascades,
                          float scale, float esf, int grid_size, int max_samples, int n_samples,
                          int64_t n_alive, float *xyzs, float *dirs, float *deltas, float *ts,
                          int32_t *n_eff) {
    memset(xyzs, 0, sizeof(float) * 3 * n_alive * n_samples);
    memset(dirs, 0, sizeof(float) * 3 * n_alive * n_samples);
    memset(deltas, 0, sizeof(float) * n_alive * n_samples);
    memset(ts, 0, sizeof(float) * n_alive * n_samples);
    for (int64_t n = 0; n < n_alive; ++n) {
        const int64_t r = alive_indices[n];
        march_ctx m;
        ctx_init(&m, rays_o + 3 * r, rays_d + 3 * r, bitfield, cascades, scale, esf, grid_size,
                 max_samples);
        float t = hits_t[2 * r];
        const float t2 = hits_t[2 * r + 1];
        int s = 0;
        while (t < t2 && s < n_samples) {
            float dt, x, y, z, t_target;
            if (probe(&m, t, &dt, &x, &y, &z, &t_target)) {
                const int64_t k = n * n_samples + s;
                xyzs[3 * k] = x; xyzs[3 * k + 1] = y; xyzs[3 * k + 2] = z;
                dirs[3 * k] = m.dx; dirs[3 * k + 1] = m.dy; dirs[3 * k + 2] = m.dz;
                ts[k] = t; deltas[k] = dt;
                t += dt;
                hits_t[2 * r] = t;
                ++s;
            } else {
                do {
                    t += calc_dt(t, esf, max_samples, grid_size, scale);
                } while (t < t_target);
            }
        }
        n_eff[n] = s;
    }
}

/* ---- compositing (custom_functions.py:116-159) ---------------
 * The oracle uses libm expf; the CUDA path documents its own exp and is held
 * to 1e-5 relative (BASELINE.json north_star).                              */
void orc_composite_train_fw(const float *sigmas, const float *rgbs, const float *deltas,
                            const float *ts, const int64_t *rays_a, float T_threshold,
                            int64_t n_rays, float *opacity, float *depth, float *depth_sq,
                            float *rgb) {
    for (int64_t n = 0; n < n_rays; ++n) {
        const int64_t ray = rays_a[3 * n], start = rays_a[3 * n + 1], N = rays_a[3 * n + 2];
        float T = 1.0f, r = 0, g = 0, b = 0, op = 0, d = 0, d2 = 0;
        for (int64_t k = 0; k < N; ++k) {
            const int64_t s = start + k;
            const float a = 1.0f - expf(-sigmas[s] * deltas[s]);
            const float w = a * T;
            r += w * rgbs[3 * s]; g += w * rgbs[3 * s + 1]; b += w * rgbs[3 * s + 2];
            d += w * ts[s]; d2 += w * ts[s] * ts[s];
            op += w;
            T *= 1.0f - a;
            if (T <= T_threshold) break;
        }
        opacity[ray] = op; depth[ray] = d; depth_sq[ray] = d2;
        rgb[3 * ray] = r; rgb[3 * ray + 1] = g; rgb[3 * ray + 2] = b;
    }
}

void orc_composite_train_bw(const float *dL_dopacity, const float *dL_ddepth,
                            const float *dL_ddepth_sq, const float *dL_drgb, const float *sigmas,
                            const float *rgbs, const float *deltas, const float *ts,
                            const int64_t *rays_a, const float *opacity, const float *depth,
                            const float *depth_sq, const float *rgb, float T_threshold,
                            int64_t n_rays, int64_t n_total, float *dL_dsigmas, float *dL_drgbs) {
    memset(dL_dsigmas, 0, sizeof(float) * n_total);
    memset(dL_drgbs, 0, sizeof(float) * 3 * n_total);
    for (int64_t n = 0; n < n_rays; ++n) {
        const int64_t ray = rays_a[3 * n], start = rays_a[3 * n + 1], N = rays_a[3 * n + 2];
        const float R = rgb[3 * ray], G = rgb[3 * ray + 1], B = rgb[3 * ray + 2];
        const float O = opacity[ray], D = depth[ray], D2 = depth_sq[ray];
        const float gr = dL_drgb[3 * ray], gg = dL_drgb[3 * ray + 1], gb = dL_drgb[3 * ray + 2];
        float T = 1.0f, r = 0, g = 0, b = 0, d = 0, d2 = 0;
        for (int64_t k = 0; k < N; ++k) {
            const int64_t s = start + k;
            const float a = 1.0f - expf(-sigmas[s] * deltas[s]);
            const float w = a * T;
            r += w * rgbs[3 * s]; g += w * rgbs[3 * s + 1]; b += w * rgbs[3 * s + 2];
            d += w * ts[s]; d2 += w * ts[s] * ts[s];
            T *= 1.0f - a;
            dL_drgbs[3 * s] = gr * w; dL_drgbs[3 * s + 1] = gg * w; dL_drgbs[3 * s + 2] = gb * w;
            dL_dsigmas[s] = deltas[s] * (gr * (rgbs[3 * s] * T - (R - r)) +
                                         gg * (rgbs[3 * s + 1] * T - (G - g)) +
                                         gb * (rgbs[3 * s + 2] * T - (B - b)) +
                                         dL_dopacity[ray] * (1 - O) +
                                         dL_ddepth[ray] * (ts[s] * T - (D - d)) +
                                         dL_ddepth_sq[ray] * (ts[s] * ts[s] * T - (D2 - d2)));
            if (T <= T_threshold) break;
        }
    }
}

void orc_composite_test_fw(const float *sigmas, const float *rgbs, const float *deltas,
                           const float *ts, const float *hits_t, int64_t *alive_indices,
                           float T_threshold, const int32_t *n_eff, int n_samples,
                           int64_t n_alive, float *opacity, float *depth, float *rgb) {
    (void)hits_t;
    for (int64_t n = 0; n < n_alive; ++n) {
        if (n_eff[n] == 0) { alive_indices[n] = -1; continue; }
        const int64_t r = alive_indices[n];
        float T = 1.0f - opacity[r];
        for (int s = 0; s < n_eff[n]; ++s) {
            const int64_t k = n * n_samples + s;
            const float a = 1.0f - expf(-sigmas[k] * deltas[k]);
            const float w = a * T;
            rgb[3 * r] += w * rgbs[3 * k]; rgb[3 * r + 1] += w * rgbs[3 * k + 1];
            rgb[3 * r + 2] += w * rgbs[3 * k + 2];
            depth[r] += w * ts[k];
            opacity[r] += w;
            T *= 1.0f - a;
            if (T <= T_threshold) { alive_indices[n] = -1; break; }
        }
    }
}
