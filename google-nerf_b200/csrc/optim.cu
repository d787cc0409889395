// optim.cu -- optimiser step, occupancy-grid maintenance and the fused NeRF loss.
// Replaces apex FusedAdam as used at ngp_pl/train.py:112, the torch ops of NGP.update_density_grid
// (ngp_pl/models/networks.py:225-252) and NeRFLoss + background blend
// (ngp_pl/losses.py:26-40, ngp_pl/models/rendering.py:159-164).  All streaming, HBM-bound kernels:
// 128-bit accesses, grid-stride loops on a grid sized to the 148 SMs.
#include "common.cuh"

#define FULL 0xffffffffu

// ------------------------------------------------------------------------------------------------ Adam
// Per parameter: read p, g, m, v (16 B), write p, m, v (12 B), zero g (4 B, optional), write fp16 copy (2 B) = 34 B.
// hyper (b2n_hyper, optional): learning rate, step count and the loss-scaler state live on the device so that a captured
// graph replays with fresh values.  found_inf != 0 (a backward kernel saw a gradient leave the fp16 range) turns the
// step into "clear the gradient, keep everything else" -- what torch.cuda.amp.GradScaler does for the reference's
// precision=16 training (ngp_pl/train.py:265): the optimiser step is skipped and does not count for the bias correction.
__global__ void __launch_bounds__(256) adam_kernel(float4 *__restrict__ p, float4 *__restrict__ g,
                                                   float4 *__restrict__ m, float4 *__restrict__ v,
                                                   __half2 *__restrict__ h, int64_t n4, float lr, float b1,
                                                   float b2, float eps, float inv_scale, int step,
                                                   const b2n_hyper *__restrict__ hyper, int zero_grad) {
    bool skip = false;
    if (hyper != nullptr) {
        lr = hyper->lr;
        step = hyper->step - hyper->skipped;
        skip = hyper->found_inf != 0;
        if (hyper->loss_scale > 0.f) inv_scale /= hyper->loss_scale;
    }
    const float c1 = 1.0f - powf(b1, (float)step), c2 = 1.0f - powf(b2, (float)step);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        if (skip) {
            if (zero_grad) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
        float *P = &pp.x, *G = &gg.x, *M = &mm.x, *V = &vv.x;
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gr = G[k] * inv_scale;
            M[k] = b1 * M[k] + (1.0f - b1) * gr;
            V[k] = b2 * V[k] + (1.0f - b2) * gr * gr;
            P[k] -= lr * (M[k] / c1) / (sqrtf(V[k] / c2) + eps);
        }
        p[i] = pp; m[i] = mm; v[i] = vv;
        if (zero_grad) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (h != nullptr) {
            h[2 * i] = __floats2half2_rn(pp.x, pp.y);
            h[2 * i + 1] = __floats2half2_rn(pp.z, pp.w);
        }
    }
}

extern "C" int b2n_adam_step(float *param, float *grad, float *exp_avg, float *exp_avg_sq, b2n_half *half_copy,
                             int64_t n, float lr, float beta1, float beta2, float eps, float inv_scale,
                             int step, const b2n_hyper *hyper_dev, int zero_grad, void *stream) {
    B2N_CHECK_ARG(n % 4 == 0, "parameter count must be a multiple of 4");
    B2N_CHECK_ARG(step >= 1 || hyper_dev != nullptr, "step is 1-based");
    if (n == 0) return 0;
#ifndef ADAM_CTAS
#define ADAM_CTAS 32         // CTAs per SM the grid is capped at: 4 / 8 / 16 / 32 / 64 / 512 -> 88.5 / 84 / 77.8 / 76.6 / 82 / 88 us
#endif
    b2n_launch(adam_kernel, b2n_grid(b2n_blocks(n / 4, 256), ADAM_CTAS), 256, (cudaStream_t)stream,
               (float4 *)param, (float4 *)grad, (float4 *)exp_avg, (float4 *)exp_avg_sq, (__half2 *)half_copy, n / 4,
               lr, beta1, beta2, eps, inv_scale, step, hyper_dev, zero_grad);
    B2N_LAUNCH_CHECK();
    return 0;
}

// GradScaler bookkeeping after the optimiser step (one thread): found_inf (OR over the ranks' blocks when a pointer
// table is given; every rank computes the same value) halves the loss scale and counts a skipped step, otherwise
// growth_interval clean steps in a row double it (0 = fixed scale).  found_inf itself is cleared by the caller once
// every rank has read it.
struct HyperPtrs { const b2n_hyper *p[16]; };
__global__ void scaler_update_kernel(b2n_hyper *hyper, const __grid_constant__ HyperPtrs all, int world) {
    int bad = hyper->found_inf;
    for (int r = 0; r < world; ++r) bad |= all.p[r] != nullptr ? *reinterpret_cast<const volatile int32_t *>(&all.p[r]->found_inf) : 0;
    if (bad) {
        hyper->skipped += 1;
        hyper->good_steps = 0;
        hyper->loss_scale = fmaxf(hyper->loss_scale * 0.5f, 1.0f);
    } else if (hyper->growth_interval > 0 && ++hyper->good_steps >= hyper->growth_interval) {
        hyper->good_steps = 0;
        hyper->loss_scale = fminf(hyper->loss_scale * 2.0f, 65536.0f);
    }
}
extern "C" int b2n_scaler_update(b2n_hyper *hyper_dev, void *const *hyper_ptrs, int world, void *stream) {
    B2N_CHECK_ARG(hyper_dev != nullptr && world >= 1 && world <= 16, "bad arguments");
    HyperPtrs all;
    for (int r = 0; r < 16; ++r) all.p[r] = (hyper_ptrs != nullptr && r < world) ? (const b2n_hyper *)hyper_ptrs[r] : nullptr;
    scaler_update_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(hyper_dev, all, hyper_ptrs != nullptr ? world : 0);
    B2N_LAUNCH_CHECK();
    return 0;
}

__global__ void __launch_bounds__(256) cast_half_kernel(const float4 *__restrict__ src, __half2 *__restrict__ dst,
                                                        int64_t n4, const float *__restrict__ tail_src,
                                                        __half *__restrict__ tail_dst, int n_tail) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(src + i);
        dst[2 * i] = __floats2half2_rn(v.x, v.y);
        dst[2 * i + 1] = __floats2half2_rn(v.z, v.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < n_tail) tail_dst[threadIdx.x] = __float2half_rn(tail_src[threadIdx.x]);
}

extern "C" int b2n_cast_half(const float *src, b2n_half *dst, int64_t n, void *stream) {
    if (n <= 0) return 0;
    B2N_CHECK_ARG(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 3) == 0, "misaligned buffers");
    const int64_t n4 = n / 4;
    cast_half_kernel<<<b2n_grid(b2n_blocks(n4 > 0 ? n4 : 1, 256), 8), 256, 0, (cudaStream_t)stream>>>(
        (const float4 *)src, (__half2 *)dst, n4, src + 4 * n4, (__half *)dst + 4 * n4, (int)(n - 4 * n4));
    B2N_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------ grid
__global__ void __launch_bounds__(256) cell_positions_kernel(const int32_t *__restrict__ coords,
                                                             const float *__restrict__ noise, int64_t n,
                                                             float gm1_inv2, float span, float half, float xyz_min,
                                                             float inv_extent, int unit_cube,
                                                             float *__restrict__ xyz) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < 3 * n; e += (int64_t)gridDim.x * blockDim.x) {
        // (coords/(G-1)*2-1)*(s-s/G) + (noise*2-1)*s/G      (networks.py:227-231)
        float v = ((float)__ldg(coords + e) * gm1_inv2 - 1.0f) * span + (__ldg(noise + e) * 2.0f - 1.0f) * half;
        if (unit_cube) v = (v - xyz_min) * inv_extent;     // networks.py:96
        xyz[e] = v;
    }
}

extern "C" int b2n_grid_cell_positions(const int32_t *coords, const float *noise, int64_t n, int grid_size, float s,
                                       float xyz_min, float xyz_max, int unit_cube, float *xyz, void *stream) {
    B2N_CHECK_ARG(grid_size >= 2, "grid_size < 2");
    if (n <= 0) return 0;
    const float half = s / grid_size;
    cell_positions_kernel<<<b2n_grid(b2n_blocks(3 * n, 256), 8), 256, 0, (cudaStream_t)stream>>>(
        coords, noise, n, 2.0f / (grid_size - 1), s - half, half, xyz_min, 1.0f / (xyz_max - xyz_min), unit_cube, xyz);
    B2N_LAUNCH_CHECK();
    return 0;
}

__global__ void __launch_bounds__(256) grid_scatter_kernel(const int64_t *__restrict__ indices,
                                                           const float *__restrict__ sigmas, int64_t n,
                                                           float *__restrict__ tmp, int64_t n_cells) {
    // tmp[indices] = sigmas (networks.py:233).  Duplicate indices: one of the written values survives, as with
    // torch's index_put; indices outside [0, n_cells) are ignored instead of writing out of bounds.
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = indices[i];
        if (c >= 0 && c < n_cells) tmp[c] = sigmas[i];
    }
}
extern "C" int b2n_grid_scatter(const int64_t *indices, const float *sigmas, int64_t n, float *tmp, int64_t n_cells,
                                void *stream) {
    if (n <= 0) return 0;
    grid_scatter_kernel<<<b2n_grid(b2n_blocks(n, 256), 8), 256, 0, (cudaStream_t)stream>>>(indices, sigmas, n, tmp, n_cells);
    B2N_LAUNCH_CHECK();
    return 0;
}

__global__ void __launch_bounds__(256) grid_ema_kernel(float *__restrict__ grid, const float *__restrict__ tmp,
                                                       int64_t n, float decay) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = grid[i];
        grid[i] = (g < 0.0f) ? g : fmaxf(g * decay, tmp[i]);   // networks.py:234-237
    }
}
extern "C" int b2n_grid_ema(float *density_grid, const float *tmp, int64_t n_cells, float decay, void *stream) {
    if (n_cells <= 0) return 0;
    grid_ema_kernel<<<b2n_grid(b2n_blocks(n_cells, 256), 8), 256, 0, (cudaStream_t)stream>>>(density_grid, tmp, n_cells, decay);
    B2N_LAUNCH_CHECK();
    return 0;
}

// mean over grid > 0 (networks.py:249) kept on the device: block partial sums in double, one atomic per
// CTA, and the last CTA to finish publishes min(mean, threshold).
__global__ void __launch_bounds__(256) grid_threshold_kernel(const float *__restrict__ grid, int64_t n,
                                                             float density_threshold, double *ws,
                                                             unsigned int *ticket, float *stats) {
    double sum = 0.0, cnt = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = __ldg(grid + i);
        if (g > 0.0f) { sum += g; cnt += 1.0; }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(FULL, sum, o);
        cnt += __shfl_xor_sync(FULL, cnt, o);
    }
    __shared__ double ss[8], sc[8];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { ss[wid] = sum; sc[wid] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0, c = 0;
        for (int k = 0; k < 8; ++k) { s += ss[k]; c += sc[k]; }
        atomicAdd(ws, s);
        atomicAdd(ws + 1, c);
        __threadfence();
        last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        const double s = atomicAdd(ws, 0.0), c = atomicAdd(ws + 1, 0.0);
        const float mean = (float)(s / c);          // NaN when no cell is positive, like torch's empty mean
        stats[0] = fminf(mean, density_threshold);  // python min(nan, thr) would keep nan; fminf keeps thr
        stats[1] = mean;
        stats[2] = (float)c;
    }
}

__global__ void grid_threshold_init_kernel(double *ws) {
    ws[0] = 0.0; ws[1] = 0.0;
    reinterpret_cast<unsigned int *>(ws + 2)[0] = 0u;
}

extern "C" int b2n_grid_threshold(const float *density_grid, int64_t n_cells, float density_threshold,
                                  double *workspace, float *stats_dev, void *stream) {
    B2N_CHECK_ARG(n_cells > 0 && workspace && stats_dev, "bad arguments (workspace = 3 doubles)");
    cudaStream_t st = (cudaStream_t)stream;
    grid_threshold_init_kernel<<<1, 1, 0, st>>>(workspace);
    grid_threshold_kernel<<<b2n_grid(b2n_blocks(n_cells, 256), 4), 256, 0, st>>>(
        density_grid, n_cells, density_threshold, workspace, reinterpret_cast<unsigned int *>(workspace + 2), stats_dev);
    B2N_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------ loss
__global__ void __launch_bounds__(256) nerf_loss_kernel(const float *__restrict__ rgb, const float *__restrict__ opacity,
                                                        const float *__restrict__ target, int64_t n, float bg,
                                                        float lambda_opa, float loss_scale, float *__restrict__ rgb_out,
                                                        float *loss, float *__restrict__ d_rgb,
                                                        float *__restrict__ d_opacity,
                                                        const float *__restrict__ loss_scale_dev) {
    if (loss_scale_dev != nullptr) loss_scale = __ldg(loss_scale_dev);
    float part = 0.f;
    const float inv3n = 1.0f / (3.0f * (float)n), invn = 1.0f / (float)n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float op = opacity[i];
        float dsum = 0.f;
        #pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = rgb[3 * i + c] + bg * (1.0f - op);      // rendering.py:163-164
            const float e = v - target[3 * i + c];
            if (rgb_out) rgb_out[3 * i + c] = v;
            part += e * e * inv3n;                                   // losses.py:34 + .mean()
            const float gr = 2.0f * e * inv3n * loss_scale;
            d_rgb[3 * i + c] = gr;
            dsum += gr;
        }
        const float o = op + 1e-10f;                                 // losses.py:36-38
        part += lambda_opa * (-o * logf(o)) * invn;
        d_opacity[i] = -bg * dsum + lambda_opa * (-(logf(o) + 1.0f)) * invn * loss_scale;
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
    __shared__ float sp[8];
    if ((threadIdx.x & 31) == 0) sp[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int k = 0; k < 8; ++k) s += sp[k];
        atomicAdd(loss, s);
    }
}

extern "C" int b2n_nerf_loss_fwbw(const float *rgb, const float *opacity, const float *target, int64_t n_rays,
                                  float bg, float lambda_opa, float loss_scale, float *rgb_out, float *loss_dev,
                                  float *dL_drgb, float *dL_dopacity, const float *loss_scale_dev, void *stream) {
    if (n_rays <= 0) return 0;
    nerf_loss_kernel<<<b2n_grid(b2n_blocks(n_rays, 256), 2), 256, 0, (cudaStream_t)stream>>>(
        rgb, opacity, target, n_rays, bg, lambda_opa, loss_scale, rgb_out, loss_dev, dL_drgb, dL_dopacity, loss_scale_dev);
    B2N_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------ depth prior
// Shift- and scale-invariant depth loss (ngp_pl/losses.py:5-23, the MiDaS loss the SCADE-style LeReS-prior path uses)
// of a ray batch, forward and backward in ONE single-CTA launch, no host synchronisation:
//   p_i = 1 / depth_i (predicted disparity), g_i = prior disparity, over the valid rays (prior > 0 and depth > 1e-6)
//   t = median (torch.median: the LOWER middle element), s = mean |x - t|, u = (p - t_p)/s_p, v = (g - t_g)/s_g
//   L = lambda / n * sum (u - v)^2
// The two medians are found by an 8-bit radix select over order-preserving integer keys (4 histogram passes in shared
// memory).  The gradient follows torch autograd through the median (it flows to the median element) and through the
// mean absolute deviation:  with r_i = 2 lambda (u_i - v_i) / n, A = sum r, B = sum r u / s_p, S = sum sign(p - t_p):
//   dL/dp_j = r_j / s_p - [j = m] A / s_p - B / n * (sign(p_j - t_p) - [j = m] S),   dL/ddepth_j = -dL/dp_j / depth_j^2.
__device__ __forceinline__ uint32_t float_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// block-wide sum of a double (1024 threads)
__device__ double block_sum(double v, double *scratch) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += scratch[k];
    return t;
}

__global__ void __launch_bounds__(1024) ssi_depth_loss_kernel(const float *__restrict__ depth,
                                                              const float *__restrict__ prior, int n, float lambda,
                                                              float loss_scale, const float *__restrict__ loss_scale_dev,
                                                              float *loss, float *__restrict__ dL_ddepth,
                                                              float *__restrict__ stats) {
    __shared__ unsigned int hist[2][256];
    __shared__ unsigned int sel_prefix[2], sel_k[2], med_idx;
    __shared__ double scratch[32];
    if (loss_scale_dev != nullptr) loss_scale = __ldg(loss_scale_dev);
    const int tid = threadIdx.x;
    // ---- valid rays
    int cnt = 0;
    for (int i = tid; i < n; i += blockDim.x) cnt += (prior[i] > 0.f && depth[i] > 1e-6f) ? 1 : 0;
    const int nv = (int)(block_sum((double)cnt, scratch) + 0.5);
    if (nv == 0) {
        for (int i = tid; i < n; i += blockDim.x) dL_ddepth[i] = 0.f;
        if (tid == 0 && stats != nullptr) { stats[0] = 0.f; stats[1] = stats[2] = stats[3] = stats[4] = 0.f; }
        return;
    }
    // ---- radix select of the lower median of both lists at once
    if (tid == 0) { sel_prefix[0] = sel_prefix[1] = 0u; sel_k[0] = sel_k[1] = (unsigned)((nv - 1) / 2); med_idx = 0xffffffffu; }
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int b = tid; b < 512; b += blockDim.x) hist[b >> 8][b & 255] = 0u;
        __syncthreads();
        const uint32_t hi_mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int i = tid; i < n; i += blockDim.x) {
            const float d = depth[i], g = prior[i];
            if (!(g > 0.f && d > 1e-6f)) continue;
            const uint32_t kp = float_key(1.0f / d), kg = float_key(g);
            if ((kp & hi_mask) == sel_prefix[0]) atomicAdd(&hist[0][(kp >> shift) & 255u], 1u);
            if ((kg & hi_mask) == sel_prefix[1]) atomicAdd(&hist[1][(kg >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 2) {
            unsigned k = sel_k[tid], acc = 0;
            int b = 0;
            for (; b < 256; ++b) {
                if (acc + hist[tid][b] > k) break;
                acc += hist[tid][b];
            }
            sel_k[tid] = k - acc;
            sel_prefix[tid] |= (uint32_t)b << shift;
        }
        __syncthreads();
    }
    const float tp = key_float(sel_prefix[0]), tg = key_float(sel_prefix[1]);
    // ---- mean absolute deviations, sign sum, the median element (smallest index among equals)
    double ap = 0.0, ag = 0.0, sg = 0.0;
    for (int i = tid; i < n; i += blockDim.x) {
        const float d = depth[i], g = prior[i];
        if (!(g > 0.f && d > 1e-6f)) continue;
        const float p = 1.0f / d;
        ap += fabsf(p - tp); ag += fabsf(g - tg);
        sg += (p > tp) ? 1.0 : ((p < tp) ? -1.0 : 0.0);
        if (p == tp) atomicMin(&med_idx, (unsigned)i);
    }
    const double sp = block_sum(ap, scratch) / nv, sgt = block_sum(ag, scratch) / nv, S = block_sum(sg, scratch);
    // ---- residual sums
    const float inv_sp = (float)(1.0 / sp), inv_sg = (float)(1.0 / sgt);
    double lsum = 0.0, A = 0.0, B = 0.0;
    for (int i = tid; i < n; i += blockDim.x) {
        const float d = depth[i], g = prior[i];
        if (!(g > 0.f && d > 1e-6f)) continue;
        const float u = (1.0f / d - tp) * inv_sp, v = (g - tg) * inv_sg, e = u - v;
        lsum += (double)e * e;
        A += e; B += (double)e * u;
    }
    const double Ls = block_sum(lsum, scratch), As = block_sum(A, scratch), Bs = block_sum(B, scratch);
    const double c = 2.0 * lambda / nv;                         // r_i = c * e_i
    const float Ar = (float)(c * As), Br = (float)(c * Bs / sp);
    const unsigned m = med_idx;
    for (int i = tid; i < n; i += blockDim.x) {
        const float d = depth[i], g = prior[i];
        float grad = 0.f;
        if (g > 0.f && d > 1e-6f) {
            const float p = 1.0f / d;
            const float u = (p - tp) * inv_sp, v = (g - tg) * inv_sg;
            const float r = (float)c * (u - v);
            const float sgn = (p > tp) ? 1.f : ((p < tp) ? -1.f : 0.f);
            const float is_m = ((unsigned)i == m) ? 1.f : 0.f;
            const float dp = r * inv_sp - is_m * Ar * inv_sp - Br / (float)nv * (sgn - is_m * (float)S);
            grad = -dp * p * p * loss_scale;
        }
        dL_ddepth[i] = grad;
    }
    if (tid == 0) {
        atomicAdd(loss, (float)(lambda * Ls / nv));
        if (stats != nullptr) { stats[0] = (float)nv; stats[1] = tp; stats[2] = (float)sp; stats[3] = tg; stats[4] = (float)sgt; }
    }
}

extern "C" int b2n_ssi_depth_loss_fwbw(const float *depth, const float *prior_disp, int64_t n_rays, float lambda,
                                       float loss_scale, const float *loss_scale_dev, float *loss_dev,
                                       float *dL_ddepth, float *stats, void *stream) {
    B2N_CHECK_ARG(depth && prior_disp && loss_dev && dL_ddepth, "null argument");
    B2N_CHECK_ARG(n_rays >= 0 && n_rays < (1ll << 30), "bad ray count");
    if (n_rays == 0) return 0;
    ssi_depth_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(depth, prior_disp, (int)n_rays, lambda, loss_scale,
                                                                 loss_scale_dev, loss_dev, dL_ddepth, stats);
    B2N_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------ membench
// Read-bandwidth probe used by bench.py to obtain the L2 roofline the hash-grid gather is reported against
// (MEASURED_PEAKS.json has no L2 figure): every thread streams 16-byte loads over a buffer `iters` times; with a
// buffer that fits the 126 MB L2 the steady state is served by L2, with a larger one by HBM.
__global__ void __launch_bounds__(256) membench_read_kernel(const uint4 *__restrict__ buf, int64_t n16, int iters,
                                                            uint32_t *__restrict__ sink) {
    uint32_t acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; ++it) {
        #pragma unroll 4
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
            uint4 v;
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + i));
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    if (acc == 0x9e3779b9u) sink[0] = acc;   // keeps the loads alive
}

// Random-gather probe: every thread issues independent 4-byte loads at pseudo-random word offsets of a buffer (an LCG
// per thread; addresses do not depend on loaded data, 8 loads in flight).  With a buffer that fits L2 but not L1 every
// load is one 32-byte sector request to L2: this measures the L2 REQUEST rate, which is what bounds the hash-grid
// gather (4 useful bytes per sector), rather than the byte bandwidth of streaming loads.
__global__ void __launch_bounds__(256) membench_gather_kernel(const uint32_t *__restrict__ buf, uint32_t word_mask, int iters,
                                                              uint32_t *__restrict__ sink) {
    uint32_t s = ((uint32_t)blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint32_t v[8];
        #pragma unroll
        for (int k = 0; k < 8; ++k) {
            s = s * 1664525u + 1013904223u;
            v[k] = __ldg(buf + ((s >> 7) & word_mask));
        }
        #pragma unroll
        for (int k = 0; k < 8; ++k) acc ^= v[k];
    }
    if (acc == 0x9e3779b9u) sink[0] = acc;   // keeps the loads alive
}

extern "C" int b2n_membench_gather(const void *buf, int64_t bytes, int iters, void *sink, int64_t *n_loads, void *stream) {
    B2N_CHECK_ARG(bytes >= 4096 && (bytes & (bytes - 1)) == 0 && iters >= 1, "bytes must be a power of two >= 4096");
    const unsigned grid = B2N_SMS * 8;
    membench_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint32_t *)buf, (uint32_t)(bytes / 4 - 1), iters,
                                                                   (uint32_t *)sink);
    B2N_LAUNCH_CHECK();
    if (n_loads != nullptr) *n_loads = (int64_t)grid * 256 * 8 * iters;
    return 0;
}

extern "C" int b2n_membench_read(const void *buf, int64_t bytes, int iters, void *sink, void *stream) {
    B2N_CHECK_ARG(bytes >= 16 && iters >= 1 && ((uintptr_t)buf & 15) == 0, "bad membench arguments");
    membench_read_kernel<<<B2N_SMS * 8, 256, 0, (cudaStream_t)stream>>>((const uint4 *)buf, bytes / 16, iters,
                                                                       (uint32_t *)sink);
    B2N_LAUNCH_CHECK();
    return 0;
}
