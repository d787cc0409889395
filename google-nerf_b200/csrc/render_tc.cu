// render_tc.cu -- test-time rendering of whole rays in ONE persistent kernel (HashGrid field, K1 = 32).
//
// Replaces the host loop of ngp_pl/models/rendering.py:42-114 (raymarching_test -> NGP.forward -> composite_test_fw,
// ~24 rounds per frame, five launches and a global live-ray count per round) for the case that matters at scale: a
// frame (or a rank's row tiles of it) whose per-round work is too small to fill 148 SMs.  At test time rays are
// independent -- the reference batches them only to fill its launches -- so here a CTA owns 128 ray slots and runs the
// reference's per-ray arithmetic start to finish without any grid-wide step:
//
//   round of a CTA (128 threads, thread = ray slot AND thread = sample row of the 128-row MMA tile):
//     A  rows of this round's tile are dealt out: a live ray of age a asks for 1 << min(a / 4, 5) samples (long rays
//        accelerate geometrically -- the reference grows its per-round count as rays die for the same reason), the rows
//        left over go to new rays taken from the global queue (one atomicAdd per CTA and round), one sample each;
//     B  every ray thread runs the reference's serial DDA loop (march.cuh, bit-identical positions) for its rows and
//        stages position / step / parameter / direction per row in shared memory;
//     C  every row thread gathers the 16 x 8 hash-grid corners of its sample (same summation order as
//        hashgrid_fw_kernel), writes the encoded row and SH-4(dir) straight into the canonical operand tiles,
//     D  the five layers run on tcgen05 with TMEM accumulators exactly as in field_mlp_fw_kernel,
//     E  every ray thread composites its rows front to back (T restarts from 1 - opacity each round like
//        composite_test_fw), retires on T <= T_threshold / volume exit and writes its pixel once.
//
// Nothing but the pixel leaves the SM: no xyzs / dirs / deltas / ts / enc / sigmas / rgbs arrays, no alive list.
// Per-ray arithmetic is the reference's; what differs from the round-synchronous loop is only WHERE a ray's rounds
// begin (T is re-derived from the accumulated opacity there) -- an fp32 rounding-level effect, within the 1e-5 the
// path is specified to -- and the per-call budget: the reference stops after >= max_samples scheduled samples per
// ray, this kernel counts rays that would cross max_samples (ctl[1]) so that the caller can fall back to the
// round-synchronous loop for such a frame (it does not happen in a box of scale 0.5: sqrt(3) / dt = 1024).
#include "field_tc.cuh"
#include "hashgrid.cuh"
#include "march.cuh"

#ifndef RENDER_AGE_SHIFT
#define RENDER_AGE_SHIFT 2     // a ray's rows per round double every 1 << RENDER_AGE_SHIFT rounds it survives
#endif
#ifndef RENDER_MAX_LOG2
#define RENDER_MAX_LOG2 5      // ... up to 32 rows
#endif
#ifndef RENDER_CTAS
#define RENDER_CTAS 4
#endif

struct RenderSmem {
    __half w[Img<32>::HALVES];                 // 20480 B  canonical weight images
    unsigned char a0[4 * ACT_LBO];             //  8256 B  encoded tile [128 x 32]
    unsigned char a1[TILE64_BYTES];            // 16512 B  hidden tile [128 x 64]; between rounds: per-row staging planes
    unsigned char a3[TILE32_BYTES];            //  8256 B  [SH16 | h16]
    uint64_t bar_w, bar_mma;
    uint32_t tmem_base;
    int32_t warp_tot[4];
    int32_t queue_base;
};

// staging planes in a1 (floats, 128 per plane): march -> row threads, then row threads -> compositor
enum { P_X = 0, P_Y, P_Z, P_DT, P_T, P_DX, P_DY, P_DZ, P_SIGMA, P_RDT, P_RT, P_R, P_G, P_B, N_PLANES };
static_assert(N_PLANES * 128 * 4 <= TILE64_BYTES, "staging planes must fit the hidden tile");

// all 16 levels of one sample -> its 64-byte encoded row (4 chunks of the canonical tile).  Per level the sum is
// (fma chain over the four x0 corners) + (fma chain over the four x0+1 corners): the order of hashgrid_fw_kernel's lane pair.
__device__ __forceinline__ void gather_row(float px, float py, float pz, const __half2 *__restrict__ table,
                                           const GridLevels &g, unsigned char *tile, int r) {
    #pragma unroll 1
    for (int c = 0; c < 4; ++c) {
        uint32_t packed[4];
        #pragma unroll
        for (int h = 0; h < 2; ++h) {
            Corner8 cn[2];
            __half2 v[2][8];
            #pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int l = 4 * c + 2 * h + q;
                level_corners(px, py, pz, g.scale[l], g.resolution[l], g.size[l], g.offset[l], g.mode[l], cn[q]);
                #pragma unroll
                for (int k = 0; k < 8; ++k) v[q][k] = __ldg(table + cn[q].idx[k]);
            }
            #pragma unroll
            for (int q = 0; q < 2; ++q) {
                float s0x = 0.f, s0y = 0.f, s1x = 0.f, s1y = 0.f;
                #pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 f0 = __half22float2(v[q][2 * k]), f1 = __half22float2(v[q][2 * k + 1]);
                    s0x = fmaf(cn[q].w[2 * k], f0.x, s0x); s0y = fmaf(cn[q].w[2 * k], f0.y, s0y);
                    s1x = fmaf(cn[q].w[2 * k + 1], f1.x, s1x); s1y = fmaf(cn[q].w[2 * k + 1], f1.y, s1y);
                }
                const __half2 o = __floats2half2_rn(__fadd_rn(s0x, s1x), __fadd_rn(s0y, s1y));
                packed[2 * h + q] = *reinterpret_cast<const uint32_t *>(&o);
            }
        }
        *reinterpret_cast<uint4 *>(tile + act_off(r, c)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
}

template <bool ESF_ZERO>
__global__ void __launch_bounds__(128, RENDER_CTAS) render_rays_kernel(
    const float *__restrict__ rays_o, const float *__restrict__ rays_d, const float *__restrict__ hits_t, int n_rays,
    const __grid_constant__ MarchParams p, const __grid_constant__ GridLevels g, const __half2 *__restrict__ table,
    const __half *__restrict__ image, float T_threshold, float *__restrict__ opacity, float *__restrict__ depth,
    float *__restrict__ rgb, int32_t *ctl, int32_t *__restrict__ ray_samples) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    RenderSmem &S = *reinterpret_cast<RenderSmem *>(smem_raw);
    using I = Img<32>;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&S.bar_w, 1);
        mbar_init(&S.bar_mma, 1);
        mbar_init_fence();
    }
    if (warp == 0) tmem_alloc<64>(&S.tmem_base);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = S.tmem_base;
    const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);
    if (tid == 0) {
        mbar_expect_tx(&S.bar_w, I::HALVES * 2);
        bulk_g2s(S.w, image, I::HALVES * 2, &S.bar_w);
    }
    mbar_wait(&S.bar_w, 0);
    const uint32_t w_addr = smem_u32(S.w), a0 = smem_u32(S.a0), a1 = smem_u32(S.a1), a3 = smem_u32(S.a3);
    float *plane = reinterpret_cast<float *>(S.a1);
    uint32_t phase = 0;

    // ---- ray slot state (registers)
    Ray q = {};
    float t = 0.f, t2 = 0.f, op = 0.f, dp = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
    int ray = -1, marched = 0, age = 0;
    bool have = false, queue_dry = false;
    int rounds = 0, truncated = 0;
    unsigned long long consumed = 0;

    while (true) {
        // ---- A: deal out the 128 rows of this round's tile
        int want = 0;
        if (have) want = min(1 << min(age >> RENDER_AGE_SHIFT, RENDER_MAX_LOG2), p.max_samples - marched);
        const int v = want | ((have ? 0 : 1) << 16);           // rows wanted | empty slot, scanned together
        int incl = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) S.warp_tot[warp] = incl;
        __syncthreads();                                       // (also: every compositor of the last round is done)
        int base = 0, total = 0;
        #pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int wt = S.warp_tot[w];
            if (w < warp) base += wt;
            total += wt;
        }
        const int excl = base + incl - v;
        const int want_start = excl & 0xffff, empty_rank = excl >> 16;
        const int total_want = total & 0xffff, total_empty = total >> 16;
        const int rows_used = min(total_want, 128);
        const int n_new = queue_dry ? 0 : min(128 - rows_used, total_empty);
        if (tid == 0) S.queue_base = n_new > 0 ? atomicAdd(ctl, n_new) : n_rays;
        plane[P_DT * 128 + tid] = 0.0f;                        // rows nobody fills stay invalid
        __syncthreads();
        const int qb = S.queue_base;
        const int avail = max(0, min(n_new, n_rays - qb));
        if (n_new > 0 && avail < n_new) queue_dry = true;
        if (total_want == 0 && avail == 0) break;              // CTA-uniform: no live ray and the queue is empty
        ++rounds;

        int ns = 0, row0 = 0;
        if (have) {
            row0 = want_start;
            ns = max(0, min(want, 128 - row0));
        } else if (empty_rank < avail) {
            ray = qb + empty_rank;
            q = load_ray(rays_o, rays_d, ray);
            t = __ldg(hits_t + 2 * ray); t2 = __ldg(hits_t + 2 * ray + 1);
            op = dp = cr = cg = cb = 0.f;
            marched = 0; age = 0; have = true;
            ns = 1; row0 = rows_used + empty_rank;
        }

        // ---- B: the reference's serial DDA loop for this ray's rows (rendering.py:79-83)
        int n_eff = 0;
        if (ns > 0) {
            while (t < t2 && n_eff < ns) {
                float dt, x, y, z, target;
                if (probe(q, t, p, dt, x, y, z, target)) {
                    const int row = row0 + n_eff;
                    plane[P_X * 128 + row] = x; plane[P_Y * 128 + row] = y; plane[P_Z * 128 + row] = z;
                    plane[P_DT * 128 + row] = dt; plane[P_T * 128 + row] = t;
                    plane[P_DX * 128 + row] = q.dx; plane[P_DY * 128 + row] = q.dy; plane[P_DZ * 128 + row] = q.dz;
                    t = __fadd_rn(t, dt);
                    ++n_eff;
                } else {
                    do {
                        t = __fadd_rn(t, ESF_ZERO ? p.dt0 : calc_dt(t, p));
                    } while (t < target);
                }
            }
            marched += n_eff;
            consumed += (unsigned long long)n_eff;
        }
        __syncthreads();

        // ---- C: thread = row: hash-grid gather + SH straight into the operand tiles
        const float s_dt = plane[P_DT * 128 + tid], s_t = plane[P_T * 128 + tid];
        const bool valid = s_dt > 0.0f;
        {
            const float x = plane[P_X * 128 + tid], y = plane[P_Y * 128 + tid], z = plane[P_Z * 128 + tid];
            float dx = plane[P_DX * 128 + tid], dy = plane[P_DY * 128 + tid], dz = plane[P_DZ * 128 + tid];
            if (valid) {
                gather_row((x - g.x_offset) * g.x_scale, (y - g.x_offset) * g.x_scale, (z - g.x_offset) * g.x_scale, table,
                           g, S.a0, tid);
            } else {
                #pragma unroll
                for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4 *>(S.a0 + act_off(tid, c)) = make_uint4(0, 0, 0, 0);
                dx = 0.f; dy = 0.f; dz = 1.f;
            }
            const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
            float sh[16];
            sh4_eval_dev(dx * inv, dy * inv, dz * inv, sh);
            if (!valid) {
                #pragma unroll
                for (int i = 0; i < 16; ++i) sh[i] = 0.f;
            }
            *reinterpret_cast<uint4 *>(S.a3 + act_off(tid, 0)) = pack8(sh);
            *reinterpret_cast<uint4 *>(S.a3 + act_off(tid, 1)) = pack8(sh + 8);
        }
        STEP_SYNC();
        // ---- D: the five layers (field_mlp_fw_kernel's chain)
        if (tid == 0) issue_layer(tmem, a0, w_addr + I::W1 * 2, 64, 32, &S.bar_mma);
        mbar_wait(&S.bar_mma, phase); phase ^= 1;
        fence_after_sync();
        {
            float acc[64];
            tmem_ld64(tmem_row, acc);
            relu_to_tile(acc, S.a1, tid);
        }
        STEP_SYNC();
        if (tid == 0) issue_layer(tmem, a1, w_addr + I::W2 * 2, 16, 64, &S.bar_mma);
        mbar_wait(&S.bar_mma, phase); phase ^= 1;
        fence_after_sync();
        float sigma;
        {
            float acc[16];
            tmem_ld16(tmem_row, acc);
            const uint4 p0 = pack8(acc), p1 = pack8(acc + 8);
            *reinterpret_cast<uint4 *>(S.a3 + act_off(tid, 2)) = p0;
            *reinterpret_cast<uint4 *>(S.a3 + act_off(tid, 3)) = p1;
            sigma = expf(__low2float(*reinterpret_cast<const __half2 *>(&p0)));      // TruncExp of the fp16-rounded h[0]
        }
        STEP_SYNC();
        if (tid == 0) issue_layer(tmem, a3, w_addr + I::W3 * 2, 64, 32, &S.bar_mma);
        mbar_wait(&S.bar_mma, phase); phase ^= 1;
        fence_after_sync();
        {
            float acc[64];
            tmem_ld64(tmem_row, acc);
            relu_to_tile(acc, S.a1, tid);
        }
        STEP_SYNC();
        if (tid == 0) issue_layer(tmem, a1, w_addr + I::W4 * 2, 64, 64, &S.bar_mma);
        mbar_wait(&S.bar_mma, phase); phase ^= 1;
        fence_after_sync();
        {
            float acc[64];
            tmem_ld64(tmem_row, acc);
            relu_to_tile(acc, S.a1, tid);
        }
        STEP_SYNC();
        if (tid == 0) issue_layer(tmem, a1, w_addr + I::W5 * 2, 16, 64, &S.bar_mma);
        mbar_wait(&S.bar_mma, phase); phase ^= 1;
        fence_after_sync();
        {
            float acc[16];
            tmem_ld16(tmem_row, acc);
            // the layer-5 MMAs have completed, nothing reads the hidden tile any more: it takes the per-row results
            plane[P_SIGMA * 128 + tid] = sigma; plane[P_RDT * 128 + tid] = s_dt; plane[P_RT * 128 + tid] = s_t;
            #pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float yv = 1.0f / (1.0f + __expf(-acc[c]));
                plane[(P_R + c) * 128 + tid] = __half2float(__float2half_rn(yv));    // rgb_net returns fp16
            }
        }
        fence_before_sync();
        __syncthreads();
        fence_after_sync();

        // ---- E: thread = ray: composite this round's rows (composite_test_fw: T restarts from 1 - opacity)
        if (ns > 0) {
            bool dead = n_eff < ns;                 // left the volume (the reference finds N_eff = 0 one round later)
            float T = 1.0f - op;
            for (int s = 0; s < n_eff; ++s) {
                const int k = row0 + s;
                const float a = 1.0f - expf(-plane[P_SIGMA * 128 + k] * plane[P_RDT * 128 + k]);
                const float wgt = a * T;
                cr = fmaf(wgt, plane[P_R * 128 + k], cr); cg = fmaf(wgt, plane[P_G * 128 + k], cg);
                cb = fmaf(wgt, plane[P_B * 128 + k], cb);
                dp = fmaf(wgt, plane[P_RT * 128 + k], dp);
                op += wgt;
                T *= 1.0f - a;
                if (T <= T_threshold) { dead = true; break; }
            }
            ++age;
            if (!dead && marched >= p.max_samples) { dead = true; ++truncated; }
            if (dead) {
                opacity[ray] = op; depth[ray] = dp;
                rgb[3 * ray] = cr; rgb[3 * ray + 1] = cg; rgb[3 * ray + 2] = cb;
                if (ray_samples != nullptr) ray_samples[ray] = marched;
                have = false;
            }
        }
    }
    // ---- statistics: rays cut at the sample budget, samples marched, the longest CTA's round count
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        truncated += __shfl_xor_sync(0xffffffffu, truncated, o);
        consumed += __shfl_xor_sync(0xffffffffu, consumed, o);
    }
    if (lane == 0) {
        if (truncated) atomicAdd(ctl + 1, truncated);
        if (consumed) atomicAdd(reinterpret_cast<unsigned long long *>(ctl + 2), consumed);
        if (warp == 0) atomicMax(ctl + 4, rounds);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem);
}

extern "C" int b2n_render_rays(const float *rays_o, const float *rays_d, const float *hits_t, int64_t n_rays,
                               const uint8_t *density_bitfield, int cascades, float scale, float exp_step_factor,
                               int grid_size, int max_samples, const b2n_grid_layout *layout, const b2n_half *table,
                               const b2n_half *image, float T_threshold, float *opacity, float *depth, float *rgb,
                               int32_t *ctl, int32_t *ray_samples, void *stream) {
    MarchParams p;
    if (fill_params(p, density_bitfield, cascades, scale, exp_step_factor, grid_size, max_samples)) return 1;
    GridLevels g;
    if (to_levels(layout, g)) return 1;
    B2N_CHECK_ARG(g.n_levels == 16, "the fused renderer is built for the 16-level HashGrid field (K1 = 32)");
    B2N_CHECK_ARG(n_rays >= 0 && n_rays < (1ll << 31) && ctl != nullptr && ((uintptr_t)ctl & 7) == 0, "bad arguments");
    B2N_CHECK_ARG(((uintptr_t)image & 15) == 0, "image must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(ctl, 0, 8 * sizeof(int32_t), st);
    if (n_rays == 0) return 0;
    const int smem = (int)sizeof(RenderSmem) + 256;
    const unsigned grid = b2n_grid((n_rays + 127) / 128, RENDER_CTAS);
    if (exp_step_factor == 0.0f) {
        cudaFuncSetAttribute(render_rays_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        render_rays_kernel<true><<<grid, 128, smem, st>>>(rays_o, rays_d, hits_t, (int)n_rays, p, g, (const __half2 *)table,
                                                          (const __half *)image, T_threshold, opacity, depth, rgb, ctl,
                                                          ray_samples);
    } else {
        cudaFuncSetAttribute(render_rays_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        render_rays_kernel<false><<<grid, 128, smem, st>>>(rays_o, rays_d, hits_t, (int)n_rays, p, g, (const __half2 *)table,
                                                           (const __half *)image, T_threshold, opacity, depth, rgb, ctl,
                                                           ray_samples);
    }
    B2N_LAUNCH_CHECK();
    return 0;
}
