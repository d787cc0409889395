// render_tc.cu -- test-time rendering of whole rays in ONE persistent kernel (HashGrid field, K1 = 32).
//
// Replaces the host loop of ngp_pl/models/rendering.py:42-114 (raymarching_test -> NGP.forward -> composite_test_fw,
// ~24 rounds per frame, five launches and a global live-ray count per round) for the case that matters at scale: a
// frame (or a rank's row tiles of it) whose per-round work is too small to fill 148 SMs.  At test time rays are
// independent -- the reference batches them only to fill its launches -- so here a CTA owns up to 128 ray slots and runs
// the reference's per-ray arithmetic start to finish without any grid-wide step:
//
//   pre-pass (render_first_hit_kernel, one thread per ray, every ray in flight): walk the empty space in front of the
//   ray; rays that meet nothing get their pixel, the others are queued at their first occupied rung.
//
//   round of a CTA (128 threads; a thread is a ray slot AND a sample row of the 128-row MMA tile):
//     A  guided self-scheduling: the CTA may hold its share of half the rays still queued (so the last rays of a frame
//        are spread over all CTAs) and takes new rays with one atomicAdd; the tile's 128 rows are dealt out: an even
//        share, more for rays that have survived long (1 << age / 4: long rays accelerate geometrically, like the
//        reference's growing per-round count);
//     B  march: every ROW thread computes and probes "its" rung of the owning ray's ladder in parallel, the ray thread
//        accepts the leading run of occupied rungs and continues the reference's serial DDA loop behind the first
//        empty one (march.cuh: bit-identical positions);
//     C  every row thread gathers the 16 x 8 hash-grid corners of its sample (summation order of hashgrid_fw_kernel)
//        and writes the encoded row and SH-4(dir) straight into the canonical operand tiles;
//     D  the five layers run on tcgen05 with TMEM accumulators exactly as in field_mlp_fw_kernel;
//     E  every ray thread composites its rows front to back (composite_test_fw's loop body), retires on
//        T <= T_threshold / volume exit and writes its pixel once (background blended, optionally into a rank's packed
//        block of a sharded frame).
//
// Nothing but the pixel leaves the SM: no xyzs / dirs / deltas / ts / enc / sigmas / rgbs arrays, no alive list.
// Per-ray arithmetic is the reference's with one difference: the transmittance is carried across a ray's rounds where
// the reference re-derives it as 1 - opacity at the start of each of ITS rounds (an fp32 rounding-level effect, within
// the 1e-5 the path is specified to).  A ray's pixel therefore does not depend on which rays share its CTA or launch:
// renders are reproducible bit for bit and a sharded frame equals the unsharded one.  The per-call sample budget of
// the reference (it stops after >= max_samples scheduled samples per ray, at a schedule-dependent point) cannot be
// reproduced without its rounds: the kernel counts rays that reach max_samples alive (ctl[1]) and the caller renders
// such a frame with the round loop (it cannot happen in a box of scale 0.5: sqrt(3) / dt = 1024).
#include "field_tc.cuh"
#include "hashgrid.cuh"
#include "march.cuh"

#ifndef RENDER_AGE_SHIFT
#define RENDER_AGE_SHIFT 2     // a ray's rows per round double every 1 << RENDER_AGE_SHIFT rounds it survives
#endif
#ifndef RENDER_MAX_LOG2
#define RENDER_MAX_LOG2 5      // ... up to 32 rows
#endif
#ifndef RENDER_MAX_SHARE
#define RENDER_MAX_SHARE 64    // rows per ray and round when a CTA holds few rays (the march of a ray's rows is serial)
#endif
#ifndef RENDER_CTAS
#define RENDER_CTAS 4
#endif
#ifndef RENDER_BALANCE
#define RENDER_BALANCE 2       // a CTA holds at most remaining / (RENDER_BALANCE x CTAs) rays ...
#endif
#ifndef RENDER_MIN_SLOTS
#define RENDER_MIN_SLOTS 2     // ... but at least this many
#endif

// where a finished ray's pixel goes: separate dense arrays (stride 3 / 1 / 1) or the columns of one packed (n, stride)
// block; bg >= 0 blends rgb + bg * (1 - opacity) (rendering.py:108-111) with torch's three separately rounded operations
struct PixelOut {
    float *opacity, *depth, *rgb;
    int stride_rgb, stride_1;
    float bg;
    float *tail;     // may be NULL: 4 floats written by the last CTA: samples marched as 16-bit digits (lo, mid, hi), rays cut
    __device__ __forceinline__ void put(int r, float op, float dp, float cr, float cg, float cb) const {
        if (bg >= 0.0f) {
            const float add = __fmul_rn(bg, __fsub_rn(1.0f, op));
            cr = __fadd_rn(cr, add); cg = __fadd_rn(cg, add); cb = __fadd_rn(cb, add);
        }
        opacity[(int64_t)r * stride_1] = op; depth[(int64_t)r * stride_1] = dp;
        float *c = rgb + (int64_t)r * stride_rgb;
        c[0] = cr; c[1] = cg; c[2] = cb;
    }
};

#ifdef RENDER_PROFILE
__device__ unsigned long long render_prof[8];       // cycles in phases A, B, C, D, E summed over CTAs, [5] rounds
#define PROF_MARK(i) do { if (tid == 0) { const long long c_ = clock64(); prof[i] += c_ - prof_t; prof_t = c_; } } while (0)
#else
#define PROF_MARK(i) do { } while (0)
#endif

struct RenderSmem {
    __half w[Img<32>::HALVES];                 // 20480 B  canonical weight images
    unsigned char a0[4 * ACT_LBO];             //  8256 B  encoded tile [128 x 32]
    unsigned char a1[TILE64_BYTES];            // 16512 B  hidden tile [128 x 64]; between rounds: per-row staging planes
    unsigned char a3[TILE32_BYTES];            //  8256 B  [SH16 | h16]
    uint64_t bar_w, bar_mma;
    uint32_t tmem_base;
    int32_t warp_tot[4], warp_scan[4];
    int32_t queue_base, queue_n;
};

// staging planes in a1 (floats, 128 per plane): march -> row threads, then row threads -> compositor
enum { P_X = 0, P_Y, P_Z, P_DT, P_T, P_DX, P_DY, P_DZ, P_SIGMA, P_RDT, P_RT, P_R, P_G, P_B,
       P_OWN, P_J,                                                          // row -> (ray slot, index of the row in the ray's run)
       R_OX, R_OY, R_OZ, R_DX, R_DY, R_DZ, R_IX, R_IY, R_IZ, R_T, R_T2,     // per ray slot: the ray and where it stands
       N_PLANES };
static_assert(N_PLANES * 128 * 4 <= TILE64_BYTES, "staging planes must fit the hidden tile");

// all 16 levels of one sample -> its 64-byte encoded row (4 chunks of the canonical tile), thread = row.  Per level the
// sum is (fma chain over the four x0 corners) + (fma chain over the four x0+1 corners): the order of hashgrid_fw_kernel's
// lane pair, so the features are bit-identical to the stand-alone kernel's.  RENDER_GATHER_LEVELS levels (8 gathers each)
// are in flight per thread.  (Measured alternatives: hashgrid_fw_kernel's two-lanes-per-sample mapping in two passes --
// half the L1 sector operations, twice the index arithmetic -- 3.50 vs 3.16 ms per 800x800 frame.)
#ifndef RENDER_GATHER_LEVELS
#define RENDER_GATHER_LEVELS 2
#endif
__device__ __forceinline__ void gather_row(float px, float py, float pz, const __half2 *__restrict__ table,
                                           const GridLevels &g, unsigned char *tile, int r) {
    constexpr int GL = RENDER_GATHER_LEVELS;
    #pragma unroll 1
    for (int c = 0; c < 4; ++c) {
        uint32_t packed[4];
        #pragma unroll
        for (int h = 0; h < 4 / GL; ++h) {
            Corner8 cn[GL];
            __half2 v[GL][8];
            #pragma unroll
            for (int q = 0; q < GL; ++q) {
                const int l = 4 * c + GL * h + q;
                level_corners(px, py, pz, g.scale[l], g.resolution[l], g.size[l], g.offset[l], g.mode[l], cn[q]);
                #pragma unroll
                for (int k = 0; k < 8; ++k) v[q][k] = __ldg(table + cn[q].idx[k]);
            }
            #pragma unroll
            for (int q = 0; q < GL; ++q) {
                float s0x = 0.f, s0y = 0.f, s1x = 0.f, s1y = 0.f;
                #pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 f0 = __half22float2(v[q][2 * k]), f1 = __half22float2(v[q][2 * k + 1]);
                    s0x = fmaf(cn[q].w[2 * k], f0.x, s0x); s0y = fmaf(cn[q].w[2 * k], f0.y, s0y);
                    s1x = fmaf(cn[q].w[2 * k + 1], f1.x, s1x); s1y = fmaf(cn[q].w[2 * k + 1], f1.y, s1y);
                }
                const __half2 o = __floats2half2_rn(__fadd_rn(s0x, s1x), __fadd_rn(s0y, s1y));
                packed[GL * h + q] = *reinterpret_cast<const uint32_t *>(&o);
            }
        }
        *reinterpret_cast<uint4 *>(tile + act_off(r, c)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
}

// Pre-pass, one thread per ray with every ray of the call in flight: walk the empty space in front of the ray (the
// serial loop's state is just t, so stopping at the first occupied probe and resuming there later is the same loop).
// Rays that never meet an occupied cell get their (zero) pixel here; the others are appended to `list` as
// (ray, bits of t at the first sample): inside the persistent kernel a new ray's first probe then hits at once instead
// of stalling its CTA's round behind a chain of dependent bitfield loads.
template <bool ESF_ZERO>
__global__ void __launch_bounds__(256) render_first_hit_kernel(
    const float *__restrict__ rays_o, const float *__restrict__ rays_d, const float *__restrict__ hits_t, int n_rays,
    const __grid_constant__ MarchParams p, const __grid_constant__ PixelOut out, int32_t *ctl, int2 *__restrict__ list,
    int32_t *__restrict__ ray_samples) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    float t = 0.f;
    if (r < n_rays) {
        const Ray q = load_ray(rays_o, rays_d, r);
        t = __ldg(hits_t + 2 * r);
        const float t2 = __ldg(hits_t + 2 * r + 1);
        while (t < t2) {
            float dt, x, y, z, target;
            if (probe(q, t, p, dt, x, y, z, target)) { hit = true; break; }
            do {
                t = __fadd_rn(t, ESF_ZERO ? p.dt0 : calc_dt(t, p));
            } while (t < target);
        }
        if (!hit) {
            out.put(r, 0.f, 0.f, 0.f, 0.f, 0.f);
            if (ray_samples != nullptr) ray_samples[r] = 0;
        }
    }
    const int lane = threadIdx.x & 31;
    const uint32_t m = __ballot_sync(0xffffffffu, hit);
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(ctl + 5, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (hit) list[base + __popc(m & ((1u << lane) - 1))] = make_int2(r, __float_as_int(t));
}

template <bool ESF_ZERO, bool LIST>
__global__ void __launch_bounds__(128, RENDER_CTAS) render_rays_kernel(
    const float *__restrict__ rays_o, const float *__restrict__ rays_d, const float *__restrict__ hits_t, int n_rays,
    const __grid_constant__ MarchParams p, const __grid_constant__ GridLevels g, const __half2 *__restrict__ table,
    const __half *__restrict__ image, float T_threshold, const __grid_constant__ PixelOut out, int32_t *ctl,
    int32_t *__restrict__ ray_samples, const int2 *__restrict__ list) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    if (LIST) n_rays = ctl[5];                                   // rays the pre-pass found a first sample for
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
    float t = 0.f, t2 = 0.f, op = 0.f, dp = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, T = 1.f;
    int ray = -1, marched = 0, age = 0;
    bool have = false, queue_dry = false;
    int rounds = 0, truncated = 0;
    unsigned long long consumed = 0;

#ifdef RENDER_PROFILE
    long long prof[5] = {0, 0, 0, 0, 0}, prof_t = clock64();
#endif
    while (true) {
        // ---- A1: which slots are free, how many rays may this CTA hold, take new rays from the queue
        const uint32_t have_m = __ballot_sync(0xffffffffu, have);
        if (lane == 0) S.warp_tot[warp] = __popc(have_m);
        __syncthreads();                                       // (also: every compositor of the last round is done)
        int n_have = 0, empty_rank = __popc(~have_m & ((1u << lane) - 1));
        #pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int c = S.warp_tot[w];
            n_have += c;
            if (w < warp) empty_rank += 32 - c;
        }
        if (tid == 0) {
            // guided self-scheduling: a CTA holds at most its share of HALF the rays still queued, so the last rays of a
            // frame are spread over all CTAs (each then marches several samples per ray and round) instead of 128 long
            // rays ending up in one CTA that needs 128 x their length / 128 rounds while the others have left
            int n_new = 0;
            if (!queue_dry) {
                const int remaining = max(0, n_rays - *reinterpret_cast<volatile int32_t *>(ctl));
                const int cap = min(128, max(RENDER_MIN_SLOTS, (remaining + RENDER_BALANCE * (int)gridDim.x - 1) /
                                                                     (RENDER_BALANCE * (int)gridDim.x)));
                n_new = max(0, min(128 - n_have, cap - n_have));
            }
            S.queue_n = n_new;
            S.queue_base = n_new > 0 ? atomicAdd(ctl, n_new) : n_rays;
        }
        plane[P_DT * 128 + tid] = 0.0f;                        // rows nobody fills stay invalid
        reinterpret_cast<int *>(plane)[P_OWN * 128 + tid] = -1;
        __syncthreads();
        const int n_new = S.queue_n;
        const int qb = S.queue_base;
        const int avail = max(0, min(n_new, n_rays - qb));
        if (n_new > 0 && qb + n_new >= n_rays) queue_dry = true;
        const int n_tot = n_have + avail;
        if (n_tot == 0) break;      // CTA-uniform: no live ray, and a CTA without rays always asks the queue (cap >= 1)
        if (++rounds > (1 << 20)) {                            // cannot happen (every round retires samples); never hang
            if (have) ++truncated;
            break;
        }
        if (!have && empty_rank < avail) {
            ray = qb + empty_rank;
            if (LIST) {
                const int2 e = __ldg(list + ray);
                ray = e.x;
                t = __int_as_float(e.y);
            } else {
                t = __ldg(hits_t + 2 * ray);
            }
            q = load_ray(rays_o, rays_d, ray);
            t2 = __ldg(hits_t + 2 * ray + 1);
            op = dp = cr = cg = cb = 0.f; T = 1.f;
            marched = 0; age = 0; have = true;
        }
        // ---- A2: deal out the 128 rows of this round's tile: an even share, more for rays that have survived long
        // (1 << age / 4, up to 32: long rays accelerate geometrically like the reference's growing per-round count)
        int want = 0;
        if (have) {
            const int share = min(RENDER_MAX_SHARE, 128 / n_tot);
            want = min(max(1 << min(age >> RENDER_AGE_SHIFT, RENDER_MAX_LOG2), share), p.max_samples - marched);
        }
        int incl = want;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) S.warp_scan[warp] = incl;
        __syncthreads();
        int row0 = incl - want;
        #pragma unroll
        for (int w = 0; w < 4; ++w)
            if (w < warp) row0 += S.warp_scan[w];
        const int ns = have ? max(0, min(want, 128 - row0)) : 0;

        PROF_MARK(0);
        // ---- B: the reference's serial DDA loop (rendering.py:79-83) for this ray's rows, in three steps so that a ray
        // with many rows does not march them one after the other while 127 threads wait:
        //   B1  the ray thread publishes its ray and parameter and labels its rows;
        //   B2  every ROW thread computes "its" rung -- the parameter the serial loop reaches after j emitted samples,
        //       t <- t + dt(t) j times, the very additions the loop performs -- probes it and stages the sample;
        //   B3  the ray thread accepts the leading run of occupied rungs (exactly what the serial loop would have
        //       emitted), discards what was staged behind the first empty rung and carries on from there serially
        //       (the skip through the empty cell and whatever follows).
        int *iplane = reinterpret_cast<int *>(plane);
        if (ns > 0) {
            plane[R_OX * 128 + tid] = q.ox; plane[R_OY * 128 + tid] = q.oy; plane[R_OZ * 128 + tid] = q.oz;
            plane[R_DX * 128 + tid] = q.dx; plane[R_DY * 128 + tid] = q.dy; plane[R_DZ * 128 + tid] = q.dz;
            plane[R_IX * 128 + tid] = q.ix; plane[R_IY * 128 + tid] = q.iy; plane[R_IZ * 128 + tid] = q.iz;
            plane[R_T * 128 + tid] = t; plane[R_T2 * 128 + tid] = t2;
            for (int s = 0; s < ns; ++s) {
                iplane[P_OWN * 128 + row0 + s] = tid;
                iplane[P_J * 128 + row0 + s] = s;
            }
        }
        __syncthreads();
        {
            const int own = iplane[P_OWN * 128 + tid];
            if (own >= 0) {
                const int j = iplane[P_J * 128 + tid];
                Ray r;
                r.ox = plane[R_OX * 128 + own]; r.oy = plane[R_OY * 128 + own]; r.oz = plane[R_OZ * 128 + own];
                r.dx = plane[R_DX * 128 + own]; r.dy = plane[R_DY * 128 + own]; r.dz = plane[R_DZ * 128 + own];
                r.ix = plane[R_IX * 128 + own]; r.iy = plane[R_IY * 128 + own]; r.iz = plane[R_IZ * 128 + own];
                float tj = plane[R_T * 128 + own];
                const float tj2 = plane[R_T2 * 128 + own];
                for (int i = 0; i < j; ++i) tj = __fadd_rn(tj, ESF_ZERO ? p.dt0 : calc_dt(tj, p));
                if (tj < tj2) {
                    float dt, x, y, z, target;
                    if (probe(r, tj, p, dt, x, y, z, target)) {
                        plane[P_X * 128 + tid] = x; plane[P_Y * 128 + tid] = y; plane[P_Z * 128 + tid] = z;
                        plane[P_DT * 128 + tid] = dt; plane[P_T * 128 + tid] = tj;
                        plane[P_DX * 128 + tid] = r.dx; plane[P_DY * 128 + tid] = r.dy; plane[P_DZ * 128 + tid] = r.dz;
                    }
                }
            }
        }
        __syncthreads();
        int n_eff = 0;
        if (ns > 0) {
            while (n_eff < ns && plane[P_DT * 128 + row0 + n_eff] > 0.0f) ++n_eff;
            if (n_eff > 0) t = __fadd_rn(plane[P_T * 128 + row0 + n_eff - 1], plane[P_DT * 128 + row0 + n_eff - 1]);
            if (n_eff < ns) {
                for (int s = n_eff + 1; s < ns; ++s) plane[P_DT * 128 + row0 + s] = 0.0f;
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
            }
            marched += n_eff;
            consumed += (unsigned long long)n_eff;
        }
        __syncthreads();
        PROF_MARK(1);

        // ---- C: thread = row: hash-grid gather + SH straight into the operand tiles
        const float s_dt = plane[P_DT * 128 + tid], s_t = plane[P_T * 128 + tid];
        const bool valid = s_dt > 0.0f;
        {
            float dx = plane[P_DX * 128 + tid], dy = plane[P_DY * 128 + tid], dz = plane[P_DZ * 128 + tid];
            if (valid) {
                const float x = plane[P_X * 128 + tid], y = plane[P_Y * 128 + tid], z = plane[P_Z * 128 + tid];
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
        PROF_MARK(2);
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

        PROF_MARK(3);
        // ---- E: thread = ray: composite this round's rows (composite_test_fw's loop body; the transmittance is carried in
        // a register across rounds -- the reference re-derives it as 1 - opacity at the start of each of its rounds --
        // so a ray's pixel does not depend on how its samples were dealt to rounds: results are reproducible bit for
        // bit although the queue order is not)
        if (ns > 0) {
            bool dead = !(t < t2);                  // left the volume (the reference finds N_eff = 0 one round later)
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
                out.put(ray, op, dp, cr, cg, cb);
                if (ray_samples != nullptr) ray_samples[ray] = marched;
                have = false;
            }
        }
        PROF_MARK(4);
    }
#ifdef RENDER_PROFILE
    if (tid == 0) {
        for (int i = 0; i < 5; ++i) atomicAdd(render_prof + i, (unsigned long long)prof[i]);
        atomicAdd(render_prof + 5, (unsigned long long)rounds);
    }
#endif
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
    if (tid == 0) {                      // the last CTA to get here publishes the call's totals next to the pixels
        __threadfence();
        if (atomicAdd(ctl + 6, 1) == (int)gridDim.x - 1 && out.tail != nullptr) {
            __threadfence();
            volatile int32_t *c = ctl;
            const uint32_t lo = (uint32_t)c[2], hi = (uint32_t)c[3];
            out.tail[0] = (float)(lo & 0xffffu); out.tail[1] = (float)(lo >> 16); out.tail[2] = (float)hi;
            out.tail[3] = (float)c[1];
        }
    }
}

template <bool ESF_ZERO>
static void launch_render(const float *rays_o, const float *rays_d, const float *hits_t, int n_rays, const MarchParams &p,
                          const GridLevels &g, const __half2 *table, const __half *image, float T_threshold,
                          const PixelOut &out, int32_t *ctl, int32_t *ray_samples, int2 *list, cudaStream_t st) {
    const int smem = (int)sizeof(RenderSmem) + 256;
    const unsigned grid = b2n_grid((n_rays + 127) / 128, RENDER_CTAS);
    if (list != nullptr) {
        render_first_hit_kernel<ESF_ZERO><<<b2n_blocks(n_rays, 256), 256, 0, st>>>(rays_o, rays_d, hits_t, n_rays, p, out, ctl,
                                                                                  list, ray_samples);
        cudaFuncSetAttribute(render_rays_kernel<ESF_ZERO, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        render_rays_kernel<ESF_ZERO, true><<<grid, 128, smem, st>>>(rays_o, rays_d, hits_t, n_rays, p, g, table, image,
                                                                     T_threshold, out, ctl, ray_samples, list);
    } else {
        cudaFuncSetAttribute(render_rays_kernel<ESF_ZERO, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        render_rays_kernel<ESF_ZERO, false><<<grid, 128, smem, st>>>(rays_o, rays_d, hits_t, n_rays, p, g, table, image,
                                                                      T_threshold, out, ctl, ray_samples, nullptr);
    }
}

extern "C" int b2n_render_rays(const float *rays_o, const float *rays_d, const float *hits_t, int64_t n_rays,
                               const uint8_t *density_bitfield, int cascades, float scale, float exp_step_factor,
                               int grid_size, int max_samples, const b2n_grid_layout *layout, const b2n_half *table,
                               const b2n_half *image, float T_threshold, float background, float *opacity, float *depth,
                               float *rgb, int out_stride, float *tail, int32_t *ctl, int32_t *ray_samples,
                               void *first_hit_list, void *stream) {
    MarchParams p;
    if (fill_params(p, density_bitfield, cascades, scale, exp_step_factor, grid_size, max_samples)) return 1;
    GridLevels g;
    if (to_levels(layout, g)) return 1;
    B2N_CHECK_ARG(g.n_levels == 16, "the fused renderer is built for the 16-level HashGrid field (K1 = 32)");
    B2N_CHECK_ARG(n_rays >= 0 && n_rays < (1ll << 31) && ctl != nullptr && ((uintptr_t)ctl & 7) == 0, "bad arguments");
    B2N_CHECK_ARG(((uintptr_t)image & 15) == 0 && ((uintptr_t)first_hit_list & 7) == 0, "image / list misaligned");
    B2N_CHECK_ARG(out_stride == 0 || out_stride >= 3, "out_stride: 0 (dense arrays) or the row stride of a packed block");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(ctl, 0, 8 * sizeof(int32_t), st);
    if (tail != nullptr) cudaMemsetAsync(tail, 0, 4 * sizeof(float), st);
    if (n_rays == 0) return 0;
    PixelOut out{opacity, depth, rgb, out_stride ? out_stride : 3, out_stride ? out_stride : 1, background, tail};
    if (exp_step_factor == 0.0f)
        launch_render<true>(rays_o, rays_d, hits_t, (int)n_rays, p, g, (const __half2 *)table, (const __half *)image,
                            T_threshold, out, ctl, ray_samples, (int2 *)first_hit_list, st);
    else
        launch_render<false>(rays_o, rays_d, hits_t, (int)n_rays, p, g, (const __half2 *)table, (const __half *)image,
                             T_threshold, out, ctl, ray_samples, (int2 *)first_hit_list, st);
    B2N_LAUNCH_CHECK();
    return 0;
}

#ifdef RENDER_PROFILE
extern "C" __attribute__((visibility("default"))) int b2n_render_profile(unsigned long long *out8, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out8, render_prof, sizeof(unsigned long long) * 8);
    if (reset) {
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(render_prof, z, sizeof(z));
    }
    return 0;
}
#endif
