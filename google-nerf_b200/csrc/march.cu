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
// loop would have visited: every rung knows its successor in the serial walk (a binary search over the
// chunk's rungs for the skip landing), and the visited set is the orbit of the entry rung, found with five
// rounds of pointer doubling (shuffle + warp OR-reduce); emitted samples are compacted with a popcount rank.  Results are bit-identical to the serial loop (oracle cross-check).
//
// The ladder itself is a serial fp32 recurrence, but where the step is constant (exp_step_factor == 0, or the
// clamped ends of the exponential schedule) and the 32 rungs stay inside one binade, repeated rounding adds
// a constant increment q = fl(t + dt) - t, so rung j = t + j*q exactly and every lane computes its rung in
// O(1); chunks that cross a binade, hit a rounding tie or sit in the geometric part fall back to the serial
// recurrence.  The count pass records each chunk's emission mask (64 words per ray) so that the write pass
// replays the ladder without touching the bitfield again.
#include "march.cuh"      // MarchParams, Ray, probe (the loop body, spelled in single IEEE operations), fill_params

#define FULL 0xffffffffu
#define WS_WORDS 64          // workspace words per ray: [0] = n_chunks | overflow << 31, [1..63] = chunk masks
#define MAX_CHUNKS (WS_WORDS - 1)
#define SERIAL_MIN_RAYS 16384  // test marcher: at or above this many live rays, thread-per-ray; below, warp-per-ray

// 32 rungs of the ladder starting at t_base: lane j gets f^j(t_base), t_next = f^32(t_base).
template <bool ESF_ZERO>
__device__ __forceinline__ void ladder(float t_base, const MarchParams &p, int lane, float &tj, float &t_next) {
    const float dt_c = ESF_ZERO ? p.dt0 : calc_dt(t_base, p);
    // constant-step regime over the whole chunk?  (calc_dt is monotone in t)
    const bool constant = ESF_ZERO || (calc_dt(t_base + 33.0f * dt_c, p) == dt_c);
    const float q = (t_base + dt_c) - t_base;             // the increment rounding actually applies
    const float t_end = t_base + 32.0f * q;
    const uint32_t b0 = __float_as_uint(t_base), b1 = __float_as_uint(t_end);
    // same binade for all 33 rungs; t_base > dt_c keeps ulp(t) >= ulp(dt) (and excludes zero / negative t)
    const bool same_binade = ((b0 ^ b1) >> 23) == 0 && (b0 >> 23) > 24 && t_base > dt_c;
    // ulp(t_base) = 2^(e-23); a residual of exactly half an ulp is a round-to-even tie whose outcome depends on t
    const float ulp = __uint_as_float((b0 & 0x7f800000u) - (23u << 23));
    const bool tie = fabsf(dt_c - q) * 2.0f == ulp;
    if (constant && same_binade && !tie) {
        tj = t_base + (float)lane * q;                    // both the product and the sum are exact
        t_next = t_end;
        return;
    }
    float t = t_base;
    tj = t_base;
    #pragma unroll 8
    for (int i = 0; i < 32; ++i) {
        t = t + (ESF_ZERO ? p.dt0 : calc_dt(t, p));
        if (i < lane) tj = t;
    }
    t_next = t;
}

// Sample sinks: count only / packed train output / (n_alive, n_samples) test output.
struct CountSink {
    __device__ __forceinline__ void put(int, float, float, float, float, float, const Ray &) const {}
};
struct PackedSink {
    float *xyzs, *dirs, *deltas, *ts;
    int64_t base;
    __device__ __forceinline__ void put(int k, float x, float y, float z, float t, float dt, const Ray &r) const {
        const int64_t s = base + k;
        xyzs[3 * s] = x; xyzs[3 * s + 1] = y; xyzs[3 * s + 2] = z;
        dirs[3 * s] = r.dx; dirs[3 * s + 1] = r.dy; dirs[3 * s + 2] = r.dz;
        ts[s] = t; deltas[s] = dt;
    }
};

// Marches one ray with the whole warp.  Emits at most `limit` samples.  Returns the number emitted;
// t_after_last = parameter after the last emitted sample (test marcher state), unchanged if none.
// ws (optional): per-ray workspace that receives the emission mask of every chunk.
template <bool ESF_ZERO, class Sink>
__device__ __forceinline__ int march_ray(const Ray &r, float t_start, float t2, int limit,
                                         const MarchParams &p, const Sink &sink, float &t_after_last,
                                         uint32_t *ws) {
    const int lane = threadIdx.x & 31;
    const float NEG_INF = __int_as_float(0xff800000);
    float t_base = t_start;
    float pending = NEG_INF;  // skip target carried over from the previous chunk
    int n = 0, chunk = 0;
    bool ws_overflow = false;
    while (true) {
        float tj, t_next;
        ladder<ESF_ZERO>(t_base, p, lane, tj, t_next);
        // ---- probe all 32 rungs in parallel
        const bool valid = tj < t2;
        float dt = 0.f, x = 0.f, y = 0.f, z = 0.f, target = NEG_INF;
        bool occ = false;
        if (valid) occ = probe(r, tj, p, dt, x, y, z, target);
        const uint32_t valid_m = __ballot_sync(FULL, valid);
        const uint32_t occ_m = __ballot_sync(FULL, occ) & valid_m;
        uint32_t emit_m = 0;
        bool done = false;
        if (occ_m == valid_m && pending == NEG_INF) {
            // every live rung is occupied and nothing is being skipped: the serial loop emits them all
            emit_m = valid_m;
            done = valid_m != FULL;
        } else {
            // next rung the serial loop probes after an empty rung j: first k > j with t_k >= target_j
            int lo = 0, hi = 32;
            #pragma unroll
            for (int s = 0; s < 5; ++s) {
                const int mid = (lo + hi) >> 1;
                const float tm = __shfl_sync(FULL, tj, mid);
                if (tm < target) lo = mid + 1; else hi = mid;
            }
            // rung the serial loop probes after rung j: j+1 after an emitted (occupied) one, the skip landing after an
            // empty one; 32 = leaves the chunk (or the ray: rungs at or beyond t2 are terminal)
            int hop = !valid ? 32 : (occ ? lane + 1 : max(lane + 1, lo));
            // which rungs does the serial loop visit?  The orbit of the entry rung under `hop`, found by pointer
            // doubling: after round k the set holds every rung within 2^(k+1)-1 hops (5 rounds cover 31 hops).
            const uint32_t reach_m = __ballot_sync(FULL, tj >= pending);
            uint32_t visit_m = reach_m ? (1u << (__ffs(reach_m) - 1)) : 0u;
            if (visit_m) {
                #pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const uint32_t contrib = (((visit_m >> lane) & 1) && hop < 32) ? (1u << hop) : 0u;
                    visit_m |= __reduce_or_sync(FULL, contrib);
                    const int hop2 = __shfl_sync(FULL, hop, hop & 31);
                    hop = hop < 32 ? hop2 : 32;
                }
                emit_m = visit_m & occ_m;
                done = (visit_m & ~valid_m) != 0;
                const int last = 31 - __clz(visit_m);                    // the rung the walk leaves the chunk from
                const float last_target = __shfl_sync(FULL, target, last);
                const bool last_occ = (occ_m >> last) & 1;
                // an occupied last rung is rung 31 (the next chunk starts at its successor); an empty one carries its
                // skip target into the next chunk
                pending = last_occ ? NEG_INF : last_target;
            }
            // visit_m == 0: no rung of this chunk reaches the pending skip target; it stays pending
        }
        // ---- compaction: rank among emitted rungs, honour the per-ray limit
        const int rank = __popc(emit_m & ((1u << lane) - 1));
        const int room = limit - n;
        const bool mine = ((emit_m >> lane) & 1) && rank < room;
        if (mine) sink.put(n + rank, x, y, z, tj, dt, r);
        const int took = min(__popc(emit_m), room);
        if (took > 0) {
            const uint32_t last_m = __ballot_sync(FULL, mine && rank == took - 1);
            t_after_last = __shfl_sync(FULL, tj + dt, __ffs(last_m) - 1);
        }
        if (ws != nullptr) {
            if (chunk < MAX_CHUNKS) {
                const uint32_t took_m = __ballot_sync(FULL, mine);
                if (lane == 0) ws[1 + chunk] = took_m;
            } else {
                ws_overflow = true;
            }
        }
        ++chunk;
        n += took;
        if (done || n >= limit) break;
        t_base = t_next;
        if (!(t_base < t2)) break;  // the next chunk's first rung already fails the loop test
    }
    if (ws != nullptr && lane == 0) ws[0] = (uint32_t)min(chunk, MAX_CHUNKS) | (ws_overflow ? 0x80000000u : 0u);
    return n;
}

// Write pass from recorded masks: recompute the ladder, no bitfield access.
template <bool ESF_ZERO>
__device__ __forceinline__ void replay_ray(const Ray &r, float t_start, int n_chunks, int limit,
                                           const MarchParams &p, const PackedSink &sink, const uint32_t *ws) {
    const int lane = threadIdx.x & 31;
    float t_base = t_start;
    int n = 0;
    for (int c = 0; c < n_chunks && n < limit; ++c) {
        float tj, t_next;
        ladder<ESF_ZERO>(t_base, p, lane, tj, t_next);
        const uint32_t m = __ldg(ws + 1 + c);
        if (m != 0) {
            const int rank = __popc(m & ((1u << lane) - 1));
            if (((m >> lane) & 1) && n + rank < limit) {
                const float x = r.ox + tj * r.dx, y = r.oy + tj * r.dy, z = r.oz + tj * r.dz;
                sink.put(n + rank, x, y, z, tj, calc_dt(tj, p), r);
            }
            n += __popc(m);
        }
        t_base = t_next;
    }
}

// ------------------------------------------------------------------------------------------------ train
// WRITE=false: count pass (rays_a[r] = [r, -, N]); WRITE=true: write pass using rays_a[r] = [r, start, N].
template <bool ESF_ZERO, bool WRITE>
__global__ void __launch_bounds__(256) march_train_kernel(const float *__restrict__ rays_o,
                                                          const float *__restrict__ rays_d,
                                                          const float *__restrict__ hits_t,
                                                          const float *__restrict__ noise, const __grid_constant__ MarchParams p,
                                                          int64_t n_rays, int64_t *rays_a, float *xyzs,
                                                          float *dirs, float *deltas, float *ts, uint32_t *workspace) {
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rays; r += warps) {
        const Ray q = load_ray(rays_o, rays_d, r);
        float t1 = __ldg(hits_t + 2 * r);
        const float t2 = __ldg(hits_t + 2 * r + 1);
        uint32_t *ws = workspace ? workspace + r * WS_WORDS : nullptr;
        int n = 0;
        bool marched = false;
        if (t1 >= 0.0f) {  // the reference loop needs 0 <= t; a miss (-1) emits nothing
            t1 += calc_dt(t1, p) * __ldg(noise + r);
            float unused;
            if (t1 < t2) {
                if (WRITE) {
                    const int limit = (int)rays_a[3 * r + 2];
                    PackedSink sink{xyzs, dirs, deltas, ts, rays_a[3 * r + 1]};
                    if (limit > 0) {
                        const uint32_t head = ws ? __ldg(ws) : 0x80000000u;
                        if (head & 0x80000000u) march_ray<ESF_ZERO>(q, t1, t2, limit, p, sink, unused, nullptr);
                        else replay_ray<ESF_ZERO>(q, t1, (int)head, limit, p, sink, ws);
                    }
                } else {
                    n = march_ray<ESF_ZERO>(q, t1, t2, p.max_samples, p, CountSink{}, unused, ws);
                    marched = true;
                }
            }
        }
        if (!WRITE && (threadIdx.x & 31) == 0) {
            rays_a[3 * r] = r;
            rays_a[3 * r + 2] = n;
            if (ws && !marched) ws[0] = 0;
        }
    }
}

// Count pass, one THREAD per ray: the reference's serial loop, recording the emission mask of every 32-rung chunk of
// the ladder (same workspace format as the warp form, so the write pass replays it unchanged).  It takes longer per
// ray than a warp, but ~10x fewer issue slots: the trainer runs it on the side stream for the NEXT batch, where its
// latency is hidden and what matters is how little it takes away from the kernels of the current step.
// The workspace must be zero on entry (chunks without a sample are not written).
template <bool ESF_ZERO>
__global__ void __launch_bounds__(128) march_train_count_serial_kernel(const float *__restrict__ rays_o,
                                                                       const float *__restrict__ rays_d,
                                                                       const float *__restrict__ hits_t,
                                                                       const float *__restrict__ noise,
                                                                       const __grid_constant__ MarchParams p,
                                                                       int64_t n_rays, int64_t *rays_a,
                                                                       uint32_t *workspace) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rays; r += (int64_t)gridDim.x * blockDim.x) {
        const Ray q = load_ray(rays_o, rays_d, r);
        float t = __ldg(hits_t + 2 * r);
        const float t2 = __ldg(hits_t + 2 * r + 1);
        uint32_t *ws = workspace + r * WS_WORDS;
        int n = 0;
        uint32_t head = 0;
        if (t >= 0.0f) {
            t += calc_dt(t, p) * __ldg(noise + r);
            if (t < t2) {
                uint32_t k = 0, word = 0, mask = 0;          // rung index on the ladder, its chunk, the chunk's mask
                bool overflow = false;
                while (t < t2 && n < p.max_samples) {
                    float dt, x, y, z, target;
                    if (probe(q, t, p, dt, x, y, z, target)) {
                        if ((k >> 5) != word) {
                            if (mask) { if (word < MAX_CHUNKS) ws[1 + word] = mask; else overflow = true; }
                            word = k >> 5; mask = 0;
                        }
                        mask |= 1u << (k & 31);
                        ++n;
                        t = t + dt; ++k;
                    } else {
                        do {
                            t = t + (ESF_ZERO ? p.dt0 : calc_dt(t, p)); ++k;
                        } while (t < target);
                    }
                }
                if (mask) { if (word < MAX_CHUNKS) ws[1 + word] = mask; else overflow = true; }
                const uint32_t chunks = n ? word + 1 : 0;
                head = (chunks < MAX_CHUNKS ? chunks : MAX_CHUNKS) | (overflow ? 0x80000000u : 0u);
            }
        }
        ws[0] = head;
        rays_a[3 * r] = r;
        rays_a[3 * r + 2] = n;
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

static inline unsigned march_grid(int64_t n_warps, int per_sm = 8) { return b2n_grid((n_warps + 7) / 8, per_sm); }

extern "C" int b2n_raymarching_train_count(const float *rays_o, const float *rays_d, const float *hits_t,
                                           const uint8_t *density_bitfield, int cascades, float scale,
                                           float exp_step_factor, const float *noise, int grid_size,
                                           int max_samples, int64_t n_rays, int64_t capacity,
                                           int64_t *rays_a, int32_t *counter, uint32_t *workspace, void *stream) {
    MarchParams p;
    if (fill_params(p, density_bitfield, cascades, scale, exp_step_factor, grid_size, max_samples)) return 1;
    B2N_CHECK_ARG(n_rays >= 0, "n_rays < 0");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rays > 0) {
        if (exp_step_factor == 0.0f)
            march_train_kernel<true, false><<<march_grid(n_rays), 256, 0, st>>>(
                rays_o, rays_d, hits_t, noise, p, n_rays, rays_a, nullptr, nullptr, nullptr, nullptr, workspace);
        else
            march_train_kernel<false, false><<<march_grid(n_rays), 256, 0, st>>>(
                rays_o, rays_d, hits_t, noise, p, n_rays, rays_a, nullptr, nullptr, nullptr, nullptr, workspace);
        B2N_LAUNCH_CHECK();
    }
    march_scan_kernel<<<1, 1024, 0, st>>>(rays_a, n_rays, capacity, counter);
    B2N_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2n_raymarching_train_count_serial(const float *rays_o, const float *rays_d, const float *hits_t,
                                                  const uint8_t *density_bitfield, int cascades, float scale,
                                                  float exp_step_factor, const float *noise, int grid_size,
                                                  int max_samples, int64_t n_rays, int64_t capacity,
                                                  int64_t *rays_a, int32_t *counter, uint32_t *workspace,
                                                  void *stream) {
    MarchParams p;
    if (fill_params(p, density_bitfield, cascades, scale, exp_step_factor, grid_size, max_samples)) return 1;
    B2N_CHECK_ARG(n_rays >= 0 && workspace != nullptr, "n_rays < 0 or no workspace");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rays > 0) {
        cudaMemsetAsync(workspace, 0, (size_t)n_rays * WS_WORDS * sizeof(uint32_t), st);
        const unsigned grid = b2n_grid(b2n_blocks(n_rays, 128), 16);
        if (exp_step_factor == 0.0f)
            march_train_count_serial_kernel<true><<<grid, 128, 0, st>>>(rays_o, rays_d, hits_t, noise, p, n_rays, rays_a,
                                                                        workspace);
        else
            march_train_count_serial_kernel<false><<<grid, 128, 0, st>>>(rays_o, rays_d, hits_t, noise, p, n_rays, rays_a,
                                                                         workspace);
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
                                           float *xyzs, float *dirs, float *deltas, float *ts,
                                           const uint32_t *workspace, void *stream) {
    MarchParams p;
    if (fill_params(p, density_bitfield, cascades, scale, exp_step_factor, grid_size, max_samples)) return 1;
    if (n_rays <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (exp_step_factor == 0.0f)
        march_train_kernel<true, true><<<march_grid(n_rays), 256, 0, st>>>(
            rays_o, rays_d, hits_t, noise, p, n_rays, (int64_t *)rays_a, xyzs, dirs, deltas, ts, (uint32_t *)workspace);
    else
        march_train_kernel<false, true><<<march_grid(n_rays), 256, 0, st>>>(
            rays_o, rays_d, hits_t, noise, p, n_rays, (int64_t *)rays_a, xyzs, dirs, deltas, ts, (uint32_t *)workspace);
    B2N_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------ test
template <bool ESF_ZERO>
__global__ void __launch_bounds__(256) march_test_kernel(const float *__restrict__ rays_o,
                                                         const float *__restrict__ rays_d, float *hits_t,
                                                         const int64_t *__restrict__ alive, const __grid_constant__ MarchParams p,
                                                         int n_samples, int64_t n_alive, float *xyzs,
                                                         float *dirs, float *deltas, float *ts,
                                                         int32_t *n_eff, const int32_t *__restrict__ ctl) {
    if (ctl != nullptr) {  // device-driven render loop: this round's ray count and samples per ray
        n_alive = ctl[0];
        n_samples = ctl[1];
    }
    if (n_alive >= SERIAL_MIN_RAYS) {
        // Many live rays (full frames): one THREAD per ray running the reference's serial loop.  With enough rays to
        // fill the machine this does ~10x less work than the warp form -- a thread touches one rung per voxel it
        // crosses and stops right after its few samples, where a warp probes 32 rungs at a time -- and neighbouring
        // pixels keep the lanes of a warp on similar paths.  Same ladder, same probe: bit-identical samples.
        for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < n_alive; n += (int64_t)gridDim.x * blockDim.x) {
            const int64_t r = alive[n];
            const Ray q = load_ray(rays_o, rays_d, r);
            float t = hits_t[2 * r];
            const float t2 = hits_t[2 * r + 1];
            const PackedSink sink{xyzs, dirs, deltas, ts, n * n_samples};
            int s = 0;
            while (t < t2 && s < n_samples) {
                float dt, x, y, z, target;
                if (probe(q, t, p, dt, x, y, z, target)) {
                    sink.put(s, x, y, z, t, dt, q);
                    t = t + dt;
                    hits_t[2 * r] = t;
                    ++s;
                } else {
                    do {
                        t = t + (ESF_ZERO ? p.dt0 : calc_dt(t, p));
                    } while (t < target);
                }
            }
            for (int k = s; k < n_samples; ++k) {            // unused slots stay zero (rendering.py:87)
                const int64_t o = n * n_samples + k;
                xyzs[3 * o] = 0.f; xyzs[3 * o + 1] = 0.f; xyzs[3 * o + 2] = 0.f;
                dirs[3 * o] = 0.f; dirs[3 * o + 1] = 0.f; dirs[3 * o + 2] = 0.f;
                ts[o] = 0.f; deltas[o] = 0.f;
            }
            n_eff[n] = s;
        }
        return;
    }
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < n_alive; n += warps) {
        const int64_t r = alive[n];
        const Ray q = load_ray(rays_o, rays_d, r);
        const float t1 = hits_t[2 * r], t2 = hits_t[2 * r + 1];
        PackedSink sink{xyzs, dirs, deltas, ts, n * n_samples};
        float t_after = t1;
        int s = 0;
        if (t1 < t2) s = march_ray<ESF_ZERO>(q, t1, t2, n_samples, p, sink, t_after, nullptr);
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

static int launch_march_test(const float *rays_o, const float *rays_d, float *hits_t, const int64_t *alive_indices,
                             const uint8_t *density_bitfield, int cascades, float scale, float exp_step_factor,
                             int grid_size, int max_samples, int n_samples, int64_t n_alive, float *xyzs, float *dirs,
                             float *deltas, float *ts, int32_t *n_eff, const int32_t *ctl, void *stream) {
    MarchParams p;
    if (fill_params(p, density_bitfield, cascades, scale, exp_step_factor, grid_size, max_samples)) return 1;
    B2N_CHECK_ARG(n_samples >= 1, "n_samples < 1");
    if (n_alive <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (exp_step_factor == 0.0f)
        march_test_kernel<true><<<march_grid(n_alive), 256, 0, st>>>(
            rays_o, rays_d, hits_t, alive_indices, p, n_samples, n_alive, xyzs, dirs, deltas, ts, n_eff, ctl);
    else
        march_test_kernel<false><<<march_grid(n_alive), 256, 0, st>>>(
            rays_o, rays_d, hits_t, alive_indices, p, n_samples, n_alive, xyzs, dirs, deltas, ts, n_eff, ctl);
    B2N_LAUNCH_CHECK();
    return 0;
}

extern "C" int b2n_raymarching_test(const float *rays_o, const float *rays_d, float *hits_t,
                                    const int64_t *alive_indices, const uint8_t *density_bitfield,
                                    int cascades, float scale, float exp_step_factor, int grid_size,
                                    int max_samples, int n_samples, int64_t n_alive, float *xyzs,
                                    float *dirs, float *deltas, float *ts, int32_t *n_eff, void *stream) {
    return launch_march_test(rays_o, rays_d, hits_t, alive_indices, density_bitfield, cascades, scale, exp_step_factor,
                             grid_size, max_samples, n_samples, n_alive, xyzs, dirs, deltas, ts, n_eff, nullptr, stream);
}

// Device-driven form: the number of live rays and the samples per ray of this round are read from ctl[0], ctl[1]
// (written by b2n_render_schedule); max_alive only sizes the grid.
extern "C" int b2n_raymarching_test_dev(const float *rays_o, const float *rays_d, float *hits_t,
                                        const int64_t *alive_indices, const uint8_t *density_bitfield, int cascades,
                                        float scale, float exp_step_factor, int grid_size, int max_samples,
                                        int64_t max_alive, const int32_t *ctl, float *xyzs, float *dirs,
                                        float *deltas, float *ts, int32_t *n_eff, void *stream) {
    B2N_CHECK_ARG(ctl != nullptr, "ctl is required");
    return launch_march_test(rays_o, rays_d, hits_t, alive_indices, density_bitfield, cascades, scale, exp_step_factor,
                             grid_size, max_samples, 1, max_alive, xyzs, dirs, deltas, ts, n_eff, ctl, stream);
}
