// march.cu -- occupancy-bitfield DDA ray marcher (train + test), warp-cooperative.
// Replaces vren.raymarching_train / vren.raymarching_test
// (ngp_pl/models/custom_functions.py:86-90, ngp_pl/models/rendering.py:79-83).
// Compiled with -fmad=false (bit-exact sample positions / indices against the oracle).
//
// Design (DESIGN.md "Marcher"): the reference walks each ray serially, t <- t + calc_dt(t), both when it
// emits a sample and inside the empty-cell skip loop.  The candidate parameters therefore form a
// path-independent "ladder" t_0, t_1, ...; a rung is probed iff no earlier probed-empty rung set a skip
// target beyond it.  One warp owns one ray: the 32 lanes hold 32 consecutive rungs, probe the bitfield in
// parallel (one L2-latency per 32 rungs instead of one per rung), then resolve which rungs the serial
// loop would have visited with ballots and a short uniform walk over the empty rungs; emitted samples are
// compacted with a popcount rank.  Results are bit-identical to the serial loop (oracle cross-check).
#include "common.cuh"

#define SQRT3 1.73205080757f
#define FULL 0xffffffffu

struct MarchParams {
    const uint8_t *bitfield;
    int cascades, grid_size, max_samples;
    float scale, esf, dt_lo, dt_hi, dt0, g_inv;  // dt0 = calc_dt for esf == 0 (constant step)
    uint32_t g3;
};

__device__ __forceinline__ float calc_dt(float t, const MarchParams &p) {
    return fminf(p.dt_hi, fmaxf(p.dt_lo, t * p.esf));
}

struct Ray {
    float ox, oy, oz, dx, dy, dz, ix, iy, iz;
};

// One DDA loop body at parameter t: occupancy of the cell, the step and (if empty) the skip target.
__device__ __forceinline__ bool probe(const Ray &r, float t, const MarchParams &p, float &dt, float &x,
                                      float &y, float &z, float &t_target) {
    const float G = (float)p.grid_size;
    x = r.ox + t * r.dx;
    y = r.oy + t * r.dy;
    z = r.oz + t * r.dz;
    dt = calc_dt(t, p);
    int e;
    frexpf(fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z))), &e);
    int mip = min(p.cascades - 1, max(0, e + 1));
    frexpf(dt * G, &e);
    mip = max(mip, min(p.cascades - 1, max(0, e)));
    const float mip_bound = fminf(__int_as_float((126 + mip) << 23), p.scale);  // scalbnf(1, mip-1)
    const float mip_bound_inv = 1.0f / mip_bound;
    const int nx = __float2int_rz(fminf(G - 1.0f, fmaxf(0.0f, 0.5f * (x * mip_bound_inv + 1) * G)));
    const int ny = __float2int_rz(fminf(G - 1.0f, fmaxf(0.0f, 0.5f * (y * mip_bound_inv + 1) * G)));
    const int nz = __float2int_rz(fminf(G - 1.0f, fmaxf(0.0f, 0.5f * (z * mip_bound_inv + 1) * G)));
    const uint32_t idx = (uint32_t)mip * p.g3 + b2n_morton3D(nx, ny, nz);
    const bool occ = (__ldg(p.bitfield + (idx >> 3)) >> (idx & 7)) & 1;
    const float tx = (((nx + 0.5f + 0.5f * copysignf(1.0f, r.dx)) * p.g_inv * 2 - 1) * mip_bound - x) * r.ix;
    const float ty = (((ny + 0.5f + 0.5f * copysignf(1.0f, r.dy)) * p.g_inv * 2 - 1) * mip_bound - y) * r.iy;
    const float tz = (((nz + 0.5f + 0.5f * copysignf(1.0f, r.dz)) * p.g_inv * 2 - 1) * mip_bound - z) * r.iz;
    t_target = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    return occ;
}

// Sample sink interfaces: count only / packed train output / (n_alive, n_samples) test output.
struct CountSink {
    __device__ __forceinline__ void put(int, float, float, float, float, float, const Ray &) const {}
};
struct TrainSink {
    float *xyzs, *dirs, *deltas, *ts;
    int64_t start;
    __device__ __forceinline__ void put(int k, float x, float y, float z, float t, float dt, const Ray &r) const {
        const int64_t s = start + k;
        xyzs[3 * s] = x; xyzs[3 * s + 1] = y; xyzs[3 * s + 2] = z;
        dirs[3 * s] = r.dx; dirs[3 * s + 1] = r.dy; dirs[3 * s + 2] = r.dz;
        ts[s] = t; deltas[s] = dt;
    }
};

// Marches one ray with the whole warp.  Emits at most `limit` samples.  Returns the number emitted;
// t_after_last = parameter after the last emitted sample (test marcher state), unchanged if none.
template <bool ESF_ZERO, class Sink>
__device__ __forceinline__ int march_ray(const Ray &r, float t_start, float t2, int limit,
                                         const MarchParams &p, const Sink &sink, float &t_after_last) {
    const int lane = threadIdx.x & 31;
    const float NEG_INF = __int_as_float(0xff800000);
    float t_base = t_start;
    float pending = NEG_INF;  // skip target carried over from the previous chunk
    int n = 0;
    while (true) {
        // ---- ladder: lane j gets t_j = f^j(t_base); lane 31 also produces the next chunk's base
        float t = t_base, tj = t_base, t_next;
        #pragma unroll
        for (int i = 0; i < 32; ++i) {
            t = t + (ESF_ZERO ? p.dt0 : calc_dt(t, p));
            if (i < lane) tj = t;
        }
        t_next = t;  // f^32(t_base), identical on every lane
        // ---- probe all 32 rungs in parallel
        const bool valid = tj < t2;
        float dt = 0.f, x = 0.f, y = 0.f, z = 0.f, target = NEG_INF;
        bool occ = false;
        if (valid) occ = probe(r, tj, p, dt, x, y, z, target);
        const uint32_t valid_m = __ballot_sync(FULL, valid);
        const uint32_t occ_m = __ballot_sync(FULL, occ) & valid_m;
        // next rung the serial loop probes after an empty rung j: first k > j with t_k >= target_j
        int lo = 0, hi = 32;
        #pragma unroll
        for (int s = 0; s < 5; ++s) {
            const int mid = (lo + hi) >> 1;
            const float tm = __shfl_sync(FULL, tj, mid);
            if (tm < target) lo = mid + 1; else hi = mid;
        }
        const int next = max(lane + 1, lo);
        // ---- which rungs does the serial loop visit?  (uniform walk; one iteration per visited empty rung)
        const uint32_t reach_m = __ballot_sync(FULL, tj >= pending);
        int cur = reach_m ? (__ffs(reach_m) - 1) : 32;
        uint32_t emit_m = 0;
        bool done = false, carried = (cur == 32);
        while (cur < 32) {
            if (!((valid_m >> cur) & 1)) { done = true; break; }
            // run of occupied rungs starting at cur
            const uint32_t stop_m = (~occ_m) & (FULL << cur);   // first non-occupied (or invalid) rung >= cur
            const int run_end = stop_m ? (__ffs(stop_m) - 1) : 32;
            if (run_end > cur) emit_m |= (run_end == 32 ? FULL : ((1u << run_end) - 1)) & (FULL << cur);
            cur = run_end;
            if (cur == 32) break;
            if (!((valid_m >> cur) & 1)) { done = true; break; }
            const int nx = __shfl_sync(FULL, next, cur);
            if (nx >= 32) { pending = __shfl_sync(FULL, target, cur); carried = true; cur = 32; }
            else cur = nx;
        }
        if (!carried) pending = NEG_INF;
        // ---- compaction: rank among emitted rungs, honour the per-ray limit
        const int rank = __popc(emit_m & ((1u << lane) - 1));
        const int room = limit - n;
        const int cnt = __popc(emit_m);
        const bool mine = ((emit_m >> lane) & 1) && rank < room;
        if (mine) sink.put(n + rank, x, y, z, tj, dt, r);
        const int took = min(cnt, room);
        if (took > 0) {
            // parameter after the last emitted sample: the lane holding rank == took-1
            const uint32_t last_m = __ballot_sync(FULL, mine && rank == took - 1);
            t_after_last = __shfl_sync(FULL, tj + dt, __ffs(last_m) - 1);
        }
        n += took;
        if (done || n >= limit) break;
        t_base = t_next;
        if (!(t_base < t2)) break;  // the next chunk's first rung already fails the loop test
    }
    return n;
}

__device__ __forceinline__ Ray load_ray(const float *rays_o, const float *rays_d, int64_t r) {
    Ray q;
    q.ox = __ldg(rays_o + 3 * r); q.oy = __ldg(rays_o + 3 * r + 1); q.oz = __ldg(rays_o + 3 * r + 2);
    q.dx = __ldg(rays_d + 3 * r); q.dy = __ldg(rays_d + 3 * r + 1); q.dz = __ldg(rays_d + 3 * r + 2);
    q.ix = 1.0f / q.dx; q.iy = 1.0f / q.dy; q.iz = 1.0f / q.dz;
    return q;
}

// ------------------------------------------------------------------------------------------------ train
// WRITE=false: count pass (rays_a[r] = [r, -, N]); WRITE=true: write pass using rays_a[r] = [r, start, N].
template <bool ESF_ZERO, bool WRITE>
__global__ void __launch_bounds__(256) march_train_kernel(const float *__restrict__ rays_o,
                                                          const float *__restrict__ rays_d,
                                                          const float *__restrict__ hits_t,
                                                          const float *__restrict__ noise, MarchParams p,
                                                          int64_t n_rays, int64_t *rays_a, float *xyzs,
                                                          float *dirs, float *deltas, float *ts) {
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rays; r += warps) {
        const Ray q = load_ray(rays_o, rays_d, r);
        float t1 = __ldg(hits_t + 2 * r);
        const float t2 = __ldg(hits_t + 2 * r + 1);
        int n = 0;
        if (t1 >= 0.0f) {  // the reference loop needs 0 <= t; a miss (-1) emits nothing
            t1 += calc_dt(t1, p) * __ldg(noise + r);
            float unused;
            if (t1 < t2) {
                if (WRITE) {
                    const int limit = (int)rays_a[3 * r + 2];
                    TrainSink sink{xyzs, dirs, deltas, ts, rays_a[3 * r + 1]};
                    if (limit > 0) n = march_ray<ESF_ZERO>(q, t1, t2, limit, p, sink, unused);
                } else {
                    n = march_ray<ESF_ZERO>(q, t1, t2, p.max_samples, p, CountSink{}, unused);
                }
            }
        }
        if (!WRITE && (threadIdx.x & 31) == 0) {
            rays_a[3 * r] = r;
            rays_a[3 * r + 2] = n;
        }
    }
}

// Exclusive scan of the per-ray counts in ray order (deterministic packing) + capacity clamp.  One CTA:
// n_rays is a batch (8192 ... a few 100k), the scan reads 8 B/ray and is latency-bound.
__global__ void __launch_bounds__(1024) march_scan_kernel(int64_t *rays_a, int64_t n_rays, int64_t capacity,
                                                          int32_t *counter) {
    __shared__ int64_t warp_sums[32];
    __shared__ int64_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t per = (n_rays + blockDim.x - 1) / blockDim.x;
    const int64_t begin = (int64_t)tid * per, end = min(n_rays, begin + per);
    int64_t sum = 0;
    for (int64_t r = begin; r < end; ++r) sum += rays_a[3 * r + 2];
    int64_t incl = sum;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int64_t w = warp_sums[lane], wi = w;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t v = __shfl_up_sync(FULL, wi, o);
            if (lane >= o) wi += v;
        }
        warp_sums[lane] = wi - w;  // exclusive
        if (lane == 31) carry_s = wi;
    }
    __syncthreads();
    int64_t run = warp_sums[wid] + incl - sum;
    bool overflow = false;
    for (int64_t r = begin; r < end; ++r) {
        int64_t n = rays_a[3 * r + 2];
        int64_t start = run;
        run += n;
        if (capacity >= 0) {
            if (start >= capacity) { start = capacity; if (n > 0) overflow = true; n = 0; }
            else if (start + n > capacity) { n = capacity - start; overflow = true; }
            rays_a[3 * r + 2] = n;
        }
        rays_a[3 * r + 1] = start;
    }
    const int any_over = __syncthreads_or(overflow ? 1 : 0);
    if (tid == 0) {
        const int64_t total = carry_s;
        counter[0] = (int32_t)((capacity >= 0 && total > capacity) ? capacity : total);
        counter[1] = (int32_t)n_rays;
        counter[2] = any_over;
        counter[3] = (int32_t)total;
    }
}

static int fill_params(MarchParams &p, const uint8_t *bitfield, int cascades, float scale, float esf,
                       int grid_size, int max_samples) {
    B2N_CHECK_ARG(cascades >= 1 && grid_size >= 1 && grid_size <= 1024 && max_samples >= 1, "bad marcher config");
    p.bitfield = bitfield; p.cascades = cascades; p.grid_size = grid_size; p.max_samples = max_samples;
    p.scale = scale; p.esf = esf;
    p.dt_lo = SQRT3 / max_samples;
    p.dt_hi = SQRT3 * 2 * scale / grid_size;
    p.dt0 = fminf(p.dt_hi, fmaxf(p.dt_lo, 0.0f));
    p.g_inv = 1.0f / grid_size;
    p.g3 = (uint32_t)grid_size * grid_size * grid_size;
    return 0;
}

static inline unsigned march_grid(int64_t n_warps) {
    // 8 warps per CTA; all warps resident at once when they fit (148 SMs x 8 CTAs x 8 warps = 9472 rays)
    return b2n_grid((n_warps + 7) / 8, 8);
}

extern "C" int b2n_raymarching_train_count(const float *rays_o, const float *rays_d, const float *hits_t,
                                           const uint8_t *density_bitfield, int cascades, float scale,
                                           float exp_step_factor, const float *noise, int grid_size,
                                           int max_samples, int64_t n_rays, int64_t capacity,
                                           int64_t *rays_a, int32_t *counter, void *stream) {
    MarchParams p;
    if (fill_params(p, density_bitfield, cascades, scale, exp_step_factor, grid_size, max_samples)) return 1;
    B2N_CHECK_ARG(n_rays >= 0, "n_rays < 0");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rays > 0) {
        if (exp_step_factor == 0.0f)
            march_train_kernel<true, false><<<march_grid(n_rays), 256, 0, st>>>(
                rays_o, rays_d, hits_t, noise, p, n_rays, rays_a, nullptr, nullptr, nullptr, nullptr);
        else
            march_train_kernel<false, false><<<march_grid(n_rays), 256, 0, st>>>(
                rays_o, rays_d, hits_t, noise, p, n_rays, rays_a, nullptr, nullptr, nullptr, nullptr);
        B2N_LAUNCH_CHECK();
    }
    march_scan_kernel<<<1, 1024, 0, st>>>(rays_a, n_rays, capacity, counter);
    B2N_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2n_raymarching_train_write(const float *rays_o, const float *rays_d, const float *hits_t,
                                           const uint8_t *density_bitfield, int cascades, float scale,
                                           float exp_step_factor, const float *noise, int grid_size,
                                           int max_samples, int64_t n_rays, const int64_t *rays_a,
                                           float *xyzs, float *dirs, float *deltas, float *ts, void *stream) {
    MarchParams p;
    if (fill_params(p, density_bitfield, cascades, scale, exp_step_factor, grid_size, max_samples)) return 1;
    if (n_rays <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (exp_step_factor == 0.0f)
        march_train_kernel<true, true><<<march_grid(n_rays), 256, 0, st>>>(
            rays_o, rays_d, hits_t, noise, p, n_rays, (int64_t *)rays_a, xyzs, dirs, deltas, ts);
    else
        march_train_kernel<false, true><<<march_grid(n_rays), 256, 0, st>>>(
            rays_o, rays_d, hits_t, noise, p, n_rays, (int64_t *)rays_a, xyzs, dirs, deltas, ts);
    B2N_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------ test
struct TestSink {
    float *xyzs, *dirs, *deltas, *ts;
    int64_t base;  // n * n_samples
    __device__ __forceinline__ void put(int k, float x, float y, float z, float t, float dt, const Ray &r) const {
        const int64_t s = base + k;
        xyzs[3 * s] = x; xyzs[3 * s + 1] = y; xyzs[3 * s + 2] = z;
        dirs[3 * s] = r.dx; dirs[3 * s + 1] = r.dy; dirs[3 * s + 2] = r.dz;
        ts[s] = t; deltas[s] = dt;
    }
};

template <bool ESF_ZERO>
__global__ void __launch_bounds__(256) march_test_kernel(const float *__restrict__ rays_o,
                                                         const float *__restrict__ rays_d, float *hits_t,
                                                         const int64_t *__restrict__ alive, MarchParams p,
                                                         int n_samples, int64_t n_alive, float *xyzs,
                                                         float *dirs, float *deltas, float *ts,
                                                         int32_t *n_eff) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < n_alive; n += warps) {
        const int64_t r = alive[n];
        const Ray q = load_ray(rays_o, rays_d, r);
        const float t1 = hits_t[2 * r], t2 = hits_t[2 * r + 1];
        TestSink sink{xyzs, dirs, deltas, ts, n * n_samples};
        float t_after = t1;
        int s = 0;
        if (t1 < t2) s = march_ray<ESF_ZERO>(q, t1, t2, n_samples, p, sink, t_after);
        // unused slots stay zero (rendering.py:87 relies on dirs == 0 to find them)
        for (int k = s + lane; k < n_samples; k += 32) {
            const int64_t o = n * n_samples + k;
            xyzs[3 * o] = 0.f; xyzs[3 * o + 1] = 0.f; xyzs[3 * o + 2] = 0.f;
            dirs[3 * o] = 0.f; dirs[3 * o + 1] = 0.f; dirs[3 * o + 2] = 0.f;
            ts[o] = 0.f; deltas[o] = 0.f;
        }
        if (lane == 0) {
            if (s > 0) hits_t[2 * r] = t_after;
            n_eff[n] = s;
        }
    }
}

extern "C" int b2n_raymarching_test(const float *rays_o, const float *rays_d, float *hits_t,
                                    const int64_t *alive_indices, const uint8_t *density_bitfield,
                                    int cascades, float scale, float exp_step_factor, int grid_size,
                                    int max_samples, int n_samples, int64_t n_alive, float *xyzs,
                                    float *dirs, float *deltas, float *ts, int32_t *n_eff, void *stream) {
    MarchParams p;
    if (fill_params(p, density_bitfield, cascades, scale, exp_step_factor, grid_size, max_samples)) return 1;
    B2N_CHECK_ARG(n_samples >= 1, "n_samples < 1");
    if (n_alive <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (exp_step_factor == 0.0f)
        march_test_kernel<true><<<march_grid(n_alive), 256, 0, st>>>(
            rays_o, rays_d, hits_t, alive_indices, p, n_samples, n_alive, xyzs, dirs, deltas, ts, n_eff);
    else
        march_test_kernel<false><<<march_grid(n_alive), 256, 0, st>>>(
            rays_o, rays_d, hits_t, alive_indices, p, n_samples, n_alive, xyzs, dirs, deltas, ts, n_eff);
    B2N_LAUNCH_CHECK();
    return 0;
}
