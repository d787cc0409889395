// composite.cu -- front-to-back transmittance compositing (train fw/bw, test fw).
// Replaces vren.composite_train_fw / composite_train_bw / composite_test_fw
// (ngp_pl/models/custom_functions.py:140-142,153-158; ngp_pl/models/rendering.py:97-100).
//
// Design (DESIGN.md "Compositing"): one warp per ray; the ray's packed samples are contiguous, so the 32
// lanes read 32 consecutive samples (coalesced 128 B lines for sigma/delta/t, 384 B for rgb) and a
// shuffle product-scan gives each lane its transmittance.  HBM-bound: 24 B/sample + 48 B/ray forward,
// 40 B/sample + 96 B/ray backward.
#include "common.cuh"

#define FULL 0xffffffffu

__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v *= u;
    }
    return v;
}
__device__ __forceinline__ float warp_scan_add(float v, int lane) {
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v += u;
    }
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

struct RayTotals { float O, D, D2, R, G, B; };

// A ray is consumed CU chunks of 32 samples at a time: the loads of the CU chunks are issued together and their
// product / sum scans are independent instruction streams, so a long ray (the kernel's critical path: half of the
// rays of a batch are empty, the rest carry 100-1000 samples) costs one memory latency and one scan latency per
// 32*CU samples.  The arithmetic per chunk -- and so every result bit -- is that of the one-chunk-at-a-time form.
#ifndef CU
#define CU 4
#endif

__device__ __forceinline__ void warp_scan_mul_cu(float (&v)[CU], int lane) {
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        #pragma unroll
        for (int u = 0; u < CU; ++u) {
            const float w = __shfl_up_sync(FULL, v[u], o);
            if (lane >= o) v[u] *= w;
        }
    }
}
__device__ __forceinline__ void warp_scan_add_cu(float (&v)[CU], int lane) {
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        #pragma unroll
        for (int u = 0; u < CU; ++u) {
            const float w = __shfl_up_sync(FULL, v[u], o);
            if (lane >= o) v[u] += w;
        }
    }
}

// Transmittance entering each of the CU chunks and after the last one.  The serial form carries
// T_run <- (T_run * P)[lane 31] from chunk to chunk; the same products are formed here from the chunk tails, so the
// values are bit-identical while the CU chunks no longer wait for one another.
__device__ __forceinline__ float chunk_entries(const float (&P)[CU], float T_run, float (&Tr)[CU]) {
    float E[CU];
    #pragma unroll
    for (int u = 0; u < CU; ++u) E[u] = __shfl_sync(FULL, P[u], 31);
    #pragma unroll
    for (int u = 0; u < CU; ++u) { Tr[u] = T_run; T_run = T_run * E[u]; }
    return T_run;
}

// Forward over one ray's packed samples (all 32 lanes of the warp take part; the totals are warp-uniform).
// n_incl receives the number of samples composited (everything up to and including the one that triggers the early
// stop): they form a prefix of the ray.
__device__ __forceinline__ RayTotals ray_forward(const float *__restrict__ sigmas, const float *__restrict__ rgbs,
                                                 const float *__restrict__ deltas, const float *__restrict__ ts,
                                                 int64_t start, int N, float T_threshold, int lane, int &n_incl) {
    float T_run = 1.0f, aO = 0.f, aD = 0.f, aD2 = 0.f, aR = 0.f, aG = 0.f, aB = 0.f;
    bool stopped = false;
    n_incl = N;
    for (int base = 0; base < N && !stopped; base += 32 * CU) {
        float a[CU], t[CU], cr[CU], cg[CU], cb[CU], P[CU], Tr[CU];
        bool active[CU];
        #pragma unroll
        for (int u = 0; u < CU; ++u) {
            const int k = base + 32 * u + lane;
            active[u] = k < N;
            const int64_t s = start + k;
            a[u] = 0.f; t[u] = 0.f; cr[u] = 0.f; cg[u] = 0.f; cb[u] = 0.f;
            if (active[u]) {
                a[u] = 1.0f - expf(-__ldg(sigmas + s) * __ldg(deltas + s));
                t[u] = __ldg(ts + s);
                cr[u] = __ldg(rgbs + 3 * s); cg[u] = __ldg(rgbs + 3 * s + 1); cb[u] = __ldg(rgbs + 3 * s + 2);
            }
            P[u] = 1.0f - a[u];
        }
        warp_scan_mul_cu(P, lane);                           // P[u] = prod_{i<=lane} (1-a_i) within chunk u
        T_run = chunk_entries(P, T_run, Tr);
        float T_before[CU];
        uint32_t dead_m[CU];
        #pragma unroll
        for (int u = 0; u < CU; ++u) {
            float Pprev = __shfl_up_sync(FULL, P[u], 1);
            if (lane == 0) Pprev = 1.0f;
            T_before[u] = Tr[u] * Pprev;
            dead_m[u] = __ballot_sync(FULL, active[u] && !(Tr[u] * P[u] > T_threshold));
        }
        #pragma unroll
        for (int u = 0; u < CU; ++u) {
            if (stopped) break;
            const int first_dead = dead_m[u] ? (__ffs(dead_m[u]) - 1) : 32;
            if (active[u] && lane <= first_dead) {
                const float w = a[u] * T_before[u];
                aO += w; aD += w * t[u]; aD2 += w * t[u] * t[u];
                aR += w * cr[u]; aG += w * cg[u]; aB += w * cb[u];
            }
            if (first_dead < 32) { stopped = true; n_incl = base + 32 * u + first_dead + 1; }
        }
    }
    RayTotals r;
    r.O = warp_sum(aO); r.D = warp_sum(aD); r.D2 = warp_sum(aD2);
    r.R = warp_sum(aR); r.G = warp_sum(aG); r.B = warp_sum(aB);
    return r;
}

// Backward over one ray: g* are dL/d(ray outputs), tot the forward totals.  Returns the number of samples that
// carry gradient (the composited prefix of the ray); the rest get exact zeros.
__device__ __forceinline__ int ray_backward(const float *__restrict__ sigmas, const float *__restrict__ rgbs,
                                            const float *__restrict__ deltas, const float *__restrict__ ts,
                                            int64_t start, int N, float T_threshold, int lane, float gO, float gD,
                                            float gD2, float gR, float gG, float gB, const RayTotals &tot,
                                            float *__restrict__ dL_dsigmas, float *__restrict__ dL_drgbs) {
    // sum_c g_c*(C_c - c_c) + gD*(D-d) + gD2*(D2-d2) = Q_total - q_prefix  (the reference's six terms,
    // regrouped so that one sum-scan per chunk suffices)
    const float Q_total = gR * tot.R + gG * tot.G + gB * tot.B + gD * tot.D + gD2 * tot.D2;
    const float opa_term = gO * (1.0f - tot.O);
    float T_run = 1.0f, q_run = 0.0f;
    bool stopped = false;
    int n_incl = N;
    for (int base = 0; base < N; base += 32 * CU) {
        if (stopped) {  // samples after an early stop get zero gradient
            #pragma unroll
            for (int u = 0; u < CU; ++u) {
                const int k = base + 32 * u + lane;
                const int64_t s = start + k;
                if (k < N) { dL_dsigmas[s] = 0.f; dL_drgbs[3 * s] = 0.f; dL_drgbs[3 * s + 1] = 0.f; dL_drgbs[3 * s + 2] = 0.f; }
            }
            continue;
        }
        float a[CU], dl[CU], gc[CU], P[CU], Tr[CU], wq[CU], w[CU], T_after[CU], T_before[CU];
        bool active[CU], incl[CU];
        uint32_t dead_m[CU];
        #pragma unroll
        for (int u = 0; u < CU; ++u) {
            const int k = base + 32 * u + lane;
            active[u] = k < N;
            const int64_t s = start + k;
            float t = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
            a[u] = 0.f; dl[u] = 0.f;
            if (active[u]) {
                dl[u] = __ldg(deltas + s);
                a[u] = 1.0f - expf(-__ldg(sigmas + s) * dl[u]);
                t = __ldg(ts + s);
                cr = __ldg(rgbs + 3 * s); cg = __ldg(rgbs + 3 * s + 1); cb = __ldg(rgbs + 3 * s + 2);
            }
            gc[u] = gR * cr + gG * cg + gB * cb + gD * t + gD2 * t * t;
            P[u] = 1.0f - a[u];
        }
        warp_scan_mul_cu(P, lane);
        T_run = chunk_entries(P, T_run, Tr);
        #pragma unroll
        for (int u = 0; u < CU; ++u) {
            float Pprev = __shfl_up_sync(FULL, P[u], 1);
            if (lane == 0) Pprev = 1.0f;
            T_after[u] = Tr[u] * P[u];
            T_before[u] = Tr[u] * Pprev;
            dead_m[u] = __ballot_sync(FULL, active[u] && !(T_after[u] > T_threshold));
        }
        #pragma unroll
        for (int u = 0; u < CU; ++u) {
            const int first_dead = dead_m[u] ? (__ffs(dead_m[u]) - 1) : 32;
            incl[u] = !stopped && active[u] && lane <= first_dead;
            w[u] = incl[u] ? a[u] * T_before[u] : 0.0f;
            wq[u] = w[u] * gc[u];
            if (!stopped && first_dead < 32) { stopped = true; n_incl = base + 32 * u + first_dead + 1; }
        }
        warp_scan_add_cu(wq, lane);
        float tails[CU];
        #pragma unroll
        for (int u = 0; u < CU; ++u) tails[u] = __shfl_sync(FULL, wq[u], 31);
        #pragma unroll
        for (int u = 0; u < CU; ++u) {
            const int64_t s = start + base + 32 * u + lane;
            const float q_incl = q_run + wq[u];
            if (active[u]) {
                float ds = 0.f, dr = 0.f, dg = 0.f, db = 0.f;
                if (incl[u]) {
                    dr = gR * w[u]; dg = gG * w[u]; db = gB * w[u];
                    ds = dl[u] * (gc[u] * T_after[u] - (Q_total - q_incl) + opa_term);
                }
                dL_dsigmas[s] = ds;
                dL_drgbs[3 * s] = dr; dL_drgbs[3 * s + 1] = dg; dL_drgbs[3 * s + 2] = db;
            }
            q_run = q_run + tails[u];                        // = q_incl of lane 31
        }
    }
    return n_incl;
}

// Appends the composited prefix [start, start + n_incl) of a ray to the compacted list of gradient-carrying samples
// (one atomicAdd per ray; the field / hash-grid backward kernels then skip the dead samples).
__device__ __forceinline__ int alive_reserve(int32_t *alive_count, int n_incl, int lane) {
    int base_i = 0;
    if (lane == 0 && n_incl > 0) base_i = atomicAdd(alive_count, n_incl);
    return base_i;                                           // lane 0's value is the one that counts
}
__device__ __forceinline__ void alive_emit(int32_t *__restrict__ alive_idx, int base_lane0, int64_t start, int n_incl,
                                           int lane) {
    const int base_i = __shfl_sync(FULL, base_lane0, 0);
    for (int k = lane; k < n_incl; k += 32) alive_idx[base_i + k] = (int32_t)(start + k);
}

__global__ void __launch_bounds__(256) composite_train_fw_kernel(
    const float *__restrict__ sigmas, const float *__restrict__ rgbs, const float *__restrict__ deltas,
    const float *__restrict__ ts, const int64_t *__restrict__ rays_a, float T_threshold, int64_t n_rays,
    float *__restrict__ opacity, float *__restrict__ depth, float *__restrict__ depth_sq,
    float *__restrict__ rgb) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < n_rays; n += warps) {
        const int64_t ray = rays_a[3 * n], start = rays_a[3 * n + 1];
        const int N = (int)rays_a[3 * n + 2];
        int n_incl;
        const RayTotals t = ray_forward(sigmas, rgbs, deltas, ts, start, N, T_threshold, lane, n_incl);
        if (lane == 0) {
            opacity[ray] = t.O; depth[ray] = t.D; depth_sq[ray] = t.D2;
            rgb[3 * ray] = t.R; rgb[3 * ray + 1] = t.G; rgb[3 * ray + 2] = t.B;
        }
    }
}

__global__ void __launch_bounds__(256) composite_train_bw_kernel(
    const float *__restrict__ dL_dopacity, const float *__restrict__ dL_ddepth,
    const float *__restrict__ dL_ddepth_sq, const float *__restrict__ dL_drgb,
    const float *__restrict__ sigmas, const float *__restrict__ rgbs, const float *__restrict__ deltas,
    const float *__restrict__ ts, const int64_t *__restrict__ rays_a, const float *__restrict__ opacity,
    const float *__restrict__ depth, const float *__restrict__ depth_sq, const float *__restrict__ rgb,
    float T_threshold, int64_t n_rays, float *__restrict__ dL_dsigmas, float *__restrict__ dL_drgbs,
    int32_t *__restrict__ alive_idx, int32_t *alive_count) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < n_rays; n += warps) {
        const int64_t ray = rays_a[3 * n], start = rays_a[3 * n + 1];
        const int N = (int)rays_a[3 * n + 2];
        RayTotals tot;
        tot.O = __ldg(opacity + ray); tot.D = __ldg(depth + ray); tot.D2 = __ldg(depth_sq + ray);
        tot.R = __ldg(rgb + 3 * ray); tot.G = __ldg(rgb + 3 * ray + 1); tot.B = __ldg(rgb + 3 * ray + 2);
        const int n_incl = ray_backward(sigmas, rgbs, deltas, ts, start, N, T_threshold, lane, __ldg(dL_dopacity + ray),
                                        __ldg(dL_ddepth + ray), __ldg(dL_ddepth_sq + ray), __ldg(dL_drgb + 3 * ray),
                                        __ldg(dL_drgb + 3 * ray + 1), __ldg(dL_drgb + 3 * ray + 2), tot, dL_dsigmas,
                                        dL_drgbs);
        if (alive_idx != nullptr) alive_emit(alive_idx, alive_reserve(alive_count, n_incl, lane), start, n_incl, lane);
    }
}

// Training fast path: compositing forward, NeRFLoss (losses.py:20-40: MSE + opacity entropy, random-background blend
// of rendering.py:163-164) and compositing backward of one ray in one warp.  The loss gradient of a ray depends on
// that ray's totals only, so nothing crosses rays except the scalar loss (one atomicAdd per block).  Saves two
// launches and the round trip of the per-ray tensors; the second pass over the ray's samples hits L1/L2.
__global__ void __launch_bounds__(256) composite_loss_fwbw_kernel(
    const float *__restrict__ sigmas, const float *__restrict__ rgbs, const float *__restrict__ deltas,
    const float *__restrict__ ts, const int64_t *__restrict__ rays_a, const float *__restrict__ target,
    float T_threshold, int64_t n_rays, float bg, float lambda_opa, float loss_scale, float *__restrict__ opacity,
    float *__restrict__ depth, float *__restrict__ rgb_out, float *loss, float *__restrict__ dL_dsigmas,
    float *__restrict__ dL_drgbs, int32_t *__restrict__ alive_idx, int32_t *alive_count,
    const float *__restrict__ loss_scale_dev) {
    if (loss_scale_dev != nullptr) loss_scale = __ldg(loss_scale_dev);    // the trainer's device-side loss scaler
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float inv3n = 1.0f / (3.0f * (float)n_rays), invn = 1.0f / (float)n_rays;
    float part = 0.f;
    for (int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < n_rays; n += warps) {
        const int64_t ray = rays_a[3 * n], start = rays_a[3 * n + 1];
        const int N = (int)rays_a[3 * n + 2];
        int n_incl;
        const RayTotals tot = ray_forward(sigmas, rgbs, deltas, ts, start, N, T_threshold, lane, n_incl);
        // reserve the ray's slots in the alive list now; the returned offset is only needed after the backward pass
        const int alive_base = alive_idx != nullptr ? alive_reserve(alive_count, n_incl, lane) : 0;
        const float c[3] = {tot.R, tot.G, tot.B};
        float g[3], dsum = 0.f, lray = 0.f;
        #pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = c[k] + bg * (1.0f - tot.O);
            const float e = v - __ldg(target + 3 * ray + k);
            if (lane == 0 && rgb_out) rgb_out[3 * ray + k] = v;
            lray += e * e * inv3n;
            g[k] = 2.0f * e * inv3n * loss_scale;
            dsum += g[k];
        }
        const float o = tot.O + 1e-10f;
        const float lo = logf(o);
        lray += lambda_opa * (-o * lo) * invn;
        const float gO = -bg * dsum + lambda_opa * (-(lo + 1.0f)) * invn * loss_scale;
        part += lray;                                        // warp-uniform; lane 0's copy is the one that is used
        if (lane == 0) { opacity[ray] = tot.O; depth[ray] = tot.D; }
        ray_backward(sigmas, rgbs, deltas, ts, start, N, T_threshold, lane, gO, 0.f, 0.f, g[0], g[1], g[2], tot,
                     dL_dsigmas, dL_drgbs);
        if (alive_idx != nullptr) alive_emit(alive_idx, alive_base, start, n_incl, lane);
    }
    __shared__ float sp[8];
    if (lane == 0) sp[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int k = 0; k < 8; ++k) s += sp[k];
        atomicAdd(loss, s);
    }
}

// Test-time compositor: one thread per alive ray, serial over its <= 64 new samples (in place).
// Device-driven form (ctl != nullptr): ray count / samples per ray come from ctl[0] / ctl[1]; rays that stay alive are
// appended to alive_next (warp-aggregated atomicAdd on ctl[4]) so that the host never compacts the list, and the
// samples consumed are added to the 64-bit counter at ctl[6..7].
__global__ void __launch_bounds__(256) composite_test_fw_kernel(
    const float *__restrict__ sigmas, const float *__restrict__ rgbs, const float *__restrict__ deltas,
    const float *__restrict__ ts, int64_t *alive_indices, float T_threshold,
    const int32_t *__restrict__ n_eff, int n_samples, int64_t n_alive, float *opacity, float *depth,
    float *rgb, int32_t *ctl, int64_t *alive_next) {
    if (ctl != nullptr) {
        n_alive = ctl[0];
        n_samples = ctl[1];
    }
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t keep = -1;
    int used = 0;
    if (n < n_alive) {
        const int ne = n_eff[n];
        used = ne;
        if (ne == 0) {
            alive_indices[n] = -1;
        } else {
            const int64_t r = alive_indices[n];
            keep = r;
            float op = opacity[r], dp = depth[r], cr = rgb[3 * r], cg = rgb[3 * r + 1], cb = rgb[3 * r + 2];
            float T = 1.0f - op;
            for (int s = 0; s < ne; ++s) {
                const int64_t k = n * n_samples + s;
                const float a = 1.0f - expf(-__ldg(sigmas + k) * __ldg(deltas + k));
                const float w = a * T;
                cr += w * __ldg(rgbs + 3 * k); cg += w * __ldg(rgbs + 3 * k + 1); cb += w * __ldg(rgbs + 3 * k + 2);
                dp += w * __ldg(ts + k);
                op += w;
                T *= 1.0f - a;
                if (T <= T_threshold) { alive_indices[n] = -1; keep = -1; break; }
            }
            opacity[r] = op; depth[r] = dp; rgb[3 * r] = cr; rgb[3 * r + 1] = cg; rgb[3 * r + 2] = cb;
        }
    }
    if (ctl != nullptr) {                                    // (all threads of the warp reach this point)
        const int lane = threadIdx.x & 31;
        const uint32_t keep_m = __ballot_sync(FULL, keep >= 0);
        int base = 0;
        if (lane == 0 && keep_m) base = atomicAdd(ctl + 4, __popc(keep_m));
        base = __shfl_sync(FULL, base, 0);
        if (keep >= 0) alive_next[base + __popc(keep_m & ((1u << lane) - 1))] = keep;
        int tot = used;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
        if (lane == 0 && tot) atomicAdd(reinterpret_cast<unsigned long long *>(ctl + 6), (unsigned long long)tot);
    }
}

// Round controller of the device-driven render loop (one thread).  ctl (8 x i32, 8-byte aligned):
// [0] live rays this round, [1] samples per ray, [2] slots = [0]*[1], [3] samples marched per ray so far,
// [4] live rays found by the compositor (input of the next round), [5] rounds run, [6..7] u64 samples consumed.
// The schedule is the reference's (rendering.py:68-71): N_samples = max(min(N_rays // N_alive, 64), min_samples),
// stop once max_samples have been marched.
__global__ void render_schedule_kernel(int32_t *ctl, int n_rays, int min_samples, int max_samples) {
    int alive = ctl[4];
    if (ctl[3] >= max_samples) alive = 0;
    int s = 0;
    if (alive > 0) {
        s = n_rays / alive;
        s = s < 64 ? s : 64;
        s = s > min_samples ? s : min_samples;
    }
    ctl[0] = alive; ctl[1] = s; ctl[2] = alive * s; ctl[3] += s; ctl[4] = 0; ctl[5] += 1;
}

extern "C" int b2n_render_schedule(int32_t *ctl, int64_t n_rays, int min_samples, int max_samples, void *stream) {
    B2N_CHECK_ARG(ctl != nullptr && n_rays >= 0 && n_rays < (1ll << 31) && min_samples >= 1, "bad arguments");
    render_schedule_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ctl, (int)n_rays, min_samples, max_samples);
    B2N_LAUNCH_CHECK();
    return 0;
}

static inline unsigned warp_grid(int64_t n_warps) { return b2n_grid((n_warps + 7) / 8, 8); }

extern "C" int b2n_composite_train_fw(const float *sigmas, const float *rgbs, const float *deltas,
                                      const float *ts, const int64_t *rays_a, float T_threshold,
                                      int64_t n_rays, float *opacity, float *depth, float *depth_sq,
                                      float *rgb, void *stream) {
    if (n_rays <= 0) return 0;
    composite_train_fw_kernel<<<warp_grid(n_rays), 256, 0, (cudaStream_t)stream>>>(
        sigmas, rgbs, deltas, ts, rays_a, T_threshold, n_rays, opacity, depth, depth_sq, rgb);
    B2N_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2n_composite_train_bw(const float *dL_dopacity, const float *dL_ddepth,
                                      const float *dL_ddepth_sq, const float *dL_drgb, const float *sigmas,
                                      const float *rgbs, const float *deltas, const float *ts,
                                      const int64_t *rays_a, const float *opacity, const float *depth,
                                      const float *depth_sq, const float *rgb, float T_threshold,
                                      int64_t n_rays, float *dL_dsigmas, float *dL_drgbs, int32_t *alive_idx,
                                      int32_t *alive_count, void *stream) {
    B2N_CHECK_ARG((alive_idx == nullptr) == (alive_count == nullptr), "alive_idx and alive_count go together");
    if (alive_count != nullptr) cudaMemsetAsync(alive_count, 0, sizeof(int32_t), (cudaStream_t)stream);
    if (n_rays <= 0) return 0;
    composite_train_bw_kernel<<<warp_grid(n_rays), 256, 0, (cudaStream_t)stream>>>(
        dL_dopacity, dL_ddepth, dL_ddepth_sq, dL_drgb, sigmas, rgbs, deltas, ts, rays_a, opacity, depth,
        depth_sq, rgb, T_threshold, n_rays, dL_dsigmas, dL_drgbs, alive_idx, alive_count);
    B2N_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2n_composite_loss_fwbw(const float *sigmas, const float *rgbs, const float *deltas, const float *ts,
                                       const int64_t *rays_a, const float *target, float T_threshold,
                                       int64_t n_rays, float bg, float lambda_opa, float loss_scale,
                                       float *opacity, float *depth, float *rgb_out, float *loss_dev,
                                       float *dL_dsigmas, float *dL_drgbs, int32_t *alive_idx,
                                       int32_t *alive_count, const float *loss_scale_dev, void *stream) {
    B2N_CHECK_ARG((alive_idx == nullptr) == (alive_count == nullptr), "alive_idx and alive_count go together");
    B2N_CHECK_ARG(loss_dev != nullptr, "loss_dev is required");
    if (alive_count != nullptr && (void *)loss_dev == (void *)(alive_count + 1)) {
        cudaMemsetAsync(alive_count, 0, 8, (cudaStream_t)stream);     // adjacent words: one clear for both
    } else {
        if (alive_count != nullptr) cudaMemsetAsync(alive_count, 0, sizeof(int32_t), (cudaStream_t)stream);
        cudaMemsetAsync(loss_dev, 0, sizeof(float), (cudaStream_t)stream);
    }
    if (n_rays <= 0) return 0;
    composite_loss_fwbw_kernel<<<warp_grid(n_rays), 256, 0, (cudaStream_t)stream>>>(
        sigmas, rgbs, deltas, ts, rays_a, target, T_threshold, n_rays, bg, lambda_opa, loss_scale, opacity, depth,
        rgb_out, loss_dev, dL_dsigmas, dL_drgbs, alive_idx, alive_count, loss_scale_dev);
    B2N_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2n_composite_test_fw(const float *sigmas, const float *rgbs, const float *deltas, const float *ts,
                                     const float *hits_t, int64_t *alive_indices,
                                     float T_threshold, const int32_t *n_eff, int n_samples,
                                     int64_t n_alive, float *opacity, float *depth, float *rgb, void *stream) {
    (void)hits_t;
    if (n_alive <= 0) return 0;
    composite_test_fw_kernel<<<b2n_blocks(n_alive, 256), 256, 0, (cudaStream_t)stream>>>(
        sigmas, rgbs, deltas, ts, alive_indices, T_threshold, n_eff, n_samples, n_alive, opacity, depth, rgb, nullptr,
        nullptr);
    B2N_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2n_composite_test_fw_dev(const float *sigmas, const float *rgbs, const float *deltas, const float *ts,
                                         int64_t *alive_indices, int64_t *alive_next, float T_threshold,
                                         const int32_t *n_eff, int64_t max_alive, int32_t *ctl, float *opacity,
                                         float *depth, float *rgb, void *stream) {
    B2N_CHECK_ARG(ctl != nullptr && alive_next != nullptr, "ctl and alive_next are required");
    if (max_alive <= 0) return 0;
    composite_test_fw_kernel<<<b2n_blocks(max_alive, 256), 256, 0, (cudaStream_t)stream>>>(
        sigmas, rgbs, deltas, ts, alive_indices, T_threshold, n_eff, 1, max_alive, opacity, depth, rgb, ctl, alive_next);
    B2N_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// RayMarcher.backward (ngp_pl/models/custom_functions.py:103-113, the --optimize_ext path): per-ray segment sums
//   dL_drays_o = sum_s dL_dxyzs[s]      dL_drays_d = sum_s (dL_dxyzs[s] * ts[s] + dL_ddirs[s])
// over the ray's packed samples [start, start + N).  The reference builds a CSR pointer from rays_a and calls
// torch_scatter.segment_csr twice; here a warp owns a ray (coalesced reads, shuffle reduction, no atomics).
__global__ void __launch_bounds__(256) raymarcher_bw_kernel(const float *__restrict__ dL_dxyzs,
                                                            const float *__restrict__ dL_ddirs,
                                                            const float *__restrict__ ts,
                                                            const int64_t *__restrict__ rays_a, int64_t n_rays,
                                                            float *__restrict__ dL_drays_o, float *__restrict__ dL_drays_d) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < n_rays; n += warps) {
        const int64_t ray = rays_a[3 * n], start = rays_a[3 * n + 1];
        const int N = (int)rays_a[3 * n + 2];
        float o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
        for (int i = lane; i < N; i += 32) {
            const int64_t s = start + i;
            const float t = __ldg(ts + s);
            #pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float g = __ldg(dL_dxyzs + 3 * s + k);
                o[k] += g;
                d[k] += g * t + (dL_ddirs != nullptr ? __ldg(dL_ddirs + 3 * s + k) : 0.f);
            }
        }
        #pragma unroll
        for (int k = 0; k < 3; ++k) {
            #pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                o[k] += __shfl_xor_sync(0xffffffffu, o[k], off);
                d[k] += __shfl_xor_sync(0xffffffffu, d[k], off);
            }
        }
        if (lane < 3) {
            dL_drays_o[3 * ray + lane] = lane == 0 ? o[0] : (lane == 1 ? o[1] : o[2]);
            dL_drays_d[3 * ray + lane] = lane == 0 ? d[0] : (lane == 1 ? d[1] : d[2]);
        }
    }
}

extern "C" int b2n_raymarcher_bw(const float *dL_dxyzs, const float *dL_ddirs, const float *ts, const int64_t *rays_a,
                                 int64_t n_rays, float *dL_drays_o, float *dL_drays_d, void *stream) {
    B2N_CHECK_ARG(dL_dxyzs && ts && rays_a && dL_drays_o && dL_drays_d, "null argument");
    if (n_rays <= 0) return 0;
    raymarcher_bw_kernel<<<warp_grid(n_rays), 256, 0, (cudaStream_t)stream>>>(dL_dxyzs, dL_ddirs, ts, rays_a, n_rays,
                                                                            dL_drays_o, dL_drays_d);
    B2N_LAUNCH_CHECK();
    return 0;
}
