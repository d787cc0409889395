/*
 * b2n.h -- C ABI of libb2n.so, the B200 (sm_100a) implementation of the Instant-NGP hot path that
 * mikacuy/google-nerf's ngp_pl runs through its `vren` extension and tiny-cuda-nn.
 *
 * Boundary rules:
 *   - plain pointers and sizes only; every pointer is a CUDA device pointer unless marked HOST;
 *   - the caller allocates every output; the library owns no persistent memory;
 *   - every entry point takes the CUDA stream to launch on (`void*` = cudaStream_t) and is asynchronous;
 *   - return value 0 = ok, non-zero = error (message via b2n_last_error(), thread-local); never exits;
 *   - tensors are contiguous, row-major, shapes written as (rows, cols).
 *
 * Each declaration cites the reference interface it replaces (paths under /root/reference/ngp_pl).
 */
#ifndef B2N_H
#define B2N_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2N_VERSION 100
#if defined(__GNUC__)
#define B2N_API __attribute__((visibility("default")))
#else
#define B2N_API
#endif
typedef uint16_t b2n_half; /* IEEE binary16 bit pattern */

/* Device-resident optimiser control block (32 bytes).  The host writes lr / step before every step (a captured CUDA
 * graph replays with fresh values); the backward kernels set found_inf; b2n_adam_step[_peer] skip the update while
 * found_inf is set and count steps as step - skipped; b2n_scaler_update advances skipped / loss_scale / good_steps.
 * This is torch.cuda.amp.GradScaler (the reference trains with precision=16, train.py:265) kept on the device. */
typedef struct {
    float lr;
    int32_t step;            /* 1-based count of steps issued */
    int32_t found_inf;       /* != 0: a gradient left the fp16 range in this step */
    int32_t skipped;         /* steps skipped so far */
    float loss_scale;        /* current loss scale (<= 0: not managed here) */
    int32_t good_steps;      /* clean steps since the last change of the scale */
    int32_t growth_interval; /* double the scale after this many clean steps; 0 = never (tcnn's fixed scale) */
    int32_t reserved;
} b2n_hyper;

B2N_API int b2n_version(void);
B2N_API const char *b2n_last_error(void);

/* ---------------------------------------------------------------- vren: intersection ----------------- */
/* vren.ray_aabb_intersect  (models/custom_functions.py:29; used by models/rendering.py:27-28).
 * hits_cnt (n_rays) i32, hits_t (n_rays,max_hits,2) f32 (-1 = miss), hits_voxel_idx (n_rays,max_hits) i64. */
B2N_API int b2n_ray_aabb_intersect(const float *rays_o, const float *rays_d, const float *centers,
                           const float *half_sizes, int64_t n_rays, int64_t n_voxels, int max_hits,
                           int32_t *hits_cnt, float *hits_t, int64_t *hits_voxel_idx, void *stream);
/* vren.ray_sphere_intersect  (models/custom_functions.py:52).  radii (n_spheres). */
B2N_API int b2n_ray_sphere_intersect(const float *rays_o, const float *rays_d, const float *centers,
                             const float *radii, int64_t n_rays, int64_t n_spheres, int max_hits,
                             int32_t *hits_cnt, float *hits_t, int64_t *hits_sphere_idx, void *stream);
/* rendering.py:29 -- hits_t[(t1>=0)&(t1<near)] = near, in place on hits_t (n_rays,1,2). */
B2N_API int b2n_clamp_near(float *hits_t, int64_t n_rays, float near_distance, void *stream);

/* get_rays (datasets/ray_utils.py:152-175) of a batch drawn as indices (datasets/base.py:24-40; train.py:150-157):
 * directions (H*W,3) camera-frame, poses (N_img,3,4) c2w, img_idxs / pix_idxs (n_rays) i64 (no bounds checks) ->
 * rays_o, rays_d (n_rays,3). */
B2N_API int b2n_rays_from_indices(const float *directions, const float *poses, const int64_t *img_idxs,
                          const int64_t *pix_idxs, int64_t n_rays, float *rays_o, float *rays_d, void *stream);

/* ---------------------------------------------------------------- vren: Morton / bitfield ------------ */
/* vren.morton3D (models/networks.py:128,147): coords (n,3) i32 -> indices (n) i32. */
B2N_API int b2n_morton3D(const int32_t *coords, int64_t n, int32_t *indices, void *stream);
/* vren.morton3D_invert (models/networks.py:153): indices (n) i32 -> coords (n,3) i32. */
B2N_API int b2n_morton3D_invert(const int32_t *indices, int64_t n, int32_t *coords, void *stream);
/* vren.packbits (models/networks.py:251-252): bit i of byte n = grid[8n+i] > threshold.
 * threshold_dev, if non-NULL, is a device float that overrides `threshold` (min(mean, thr) computed on
 * the device by b2n_density_grid_update, removing the .item() sync at networks.py:249). */
B2N_API int b2n_packbits(const float *density_grid, int64_t n_bytes, float threshold, const float *threshold_dev,
                 uint8_t *density_bitfield, void *stream);

/* ---------------------------------------------------------------- vren: ray marching ----------------- */
/* vren.raymarching_train (models/custom_functions.py:86-90), split in two so that the caller can size the
 * sample buffers: _count fills rays_a (n_rays,3) i64 = [ray_idx, start_idx, N] with start_idx the exclusive
 * prefix sum of N in ray order, and counter (4) i32 = [total, n_rays, overflow, unclamped_total]; _write
 * re-marches and writes xyzs, dirs (total,3), deltas, ts (total).
 * capacity >= 0 clamps: rays whose samples would pass `capacity` rows are truncated (N reduced) and
 * counter[2] is set; capacity < 0 = unbounded.  hits_t (n_rays,2); noise (n_rays) in [0,1).
 * workspace (optional, may be NULL in both): 64 uint32 per ray; _count records every 32-rung chunk's emission
 * mask there and _write then replays them without reading the bitfield again. */
B2N_API int b2n_raymarching_train_count(const float *rays_o, const float *rays_d, const float *hits_t,
                                const uint8_t *density_bitfield, int cascades, float scale,
                                float exp_step_factor, const float *noise, int grid_size,
                                int max_samples, int64_t n_rays, int64_t capacity, int64_t *rays_a,
                                int32_t *counter, uint32_t *workspace, void *stream);
/* Same contract and results as b2n_raymarching_train_count with a workspace (required here), computed by one thread
 * per ray running the reference's serial loop: several times the latency, about a tenth of the issue slots -- for
 * marching the NEXT batch on a side stream underneath the current step's kernels. */
B2N_API int b2n_raymarching_train_count_serial(const float *rays_o, const float *rays_d, const float *hits_t,
                                       const uint8_t *density_bitfield, int cascades, float scale,
                                       float exp_step_factor, const float *noise, int grid_size,
                                       int max_samples, int64_t n_rays, int64_t capacity, int64_t *rays_a,
                                       int32_t *counter, uint32_t *workspace, void *stream);
B2N_API int b2n_raymarching_train_write(const float *rays_o, const float *rays_d, const float *hits_t,
                                const uint8_t *density_bitfield, int cascades, float scale,
                                float exp_step_factor, const float *noise, int grid_size,
                                int max_samples, int64_t n_rays, const int64_t *rays_a, float *xyzs,
                                float *dirs, float *deltas, float *ts, const uint32_t *workspace, void *stream);
/* RayMarcher.backward (models/custom_functions.py:103-113; reached with --optimize_ext): per-ray segment sums over
 * the packed samples of rays_a (n_rays,3) = [ray_idx, start, N]: dL_drays_o[ray] = sum dL_dxyzs,
 * dL_drays_d[ray] = sum (dL_dxyzs * ts + dL_ddirs).  dL_ddirs may be NULL.  Every ray row is written (rays_a holds
 * each ray once). */
B2N_API int b2n_raymarcher_bw(const float *dL_dxyzs, const float *dL_ddirs, const float *ts, const int64_t *rays_a,
                              int64_t n_rays, float *dL_drays_o, float *dL_drays_d, void *stream);
/* vren.raymarching_test (models/rendering.py:79-83).  hits_t (n_rays,2) is advanced IN PLACE;
 * outputs (n_alive,n_samples[,3]) are fully written (unused slots zero); n_eff (n_alive) i32. */
B2N_API int b2n_raymarching_test(const float *rays_o, const float *rays_d, float *hits_t,
                         const int64_t *alive_indices, const uint8_t *density_bitfield, int cascades,
                         float scale, float exp_step_factor, int grid_size, int max_samples,
                         int n_samples, int64_t n_alive, float *xyzs, float *dirs, float *deltas,
                         float *ts, int32_t *n_eff, void *stream);

/* Device-driven test-time loop (the body of models/rendering.py:64-102 without its per-iteration host syncs).
 * ctl: 8 x i32 on the device, 8-byte aligned: [0] live rays this round, [1] samples per ray, [2] slots = [0]*[1],
 * [3] samples marched per ray so far, [4] live rays left by the compositor (the next round's [0]; the caller seeds it
 * with n_rays), [5] rounds run, [6..7] u64 samples consumed.  One round = b2n_render_schedule ->
 * b2n_raymarching_test_dev -> field kernels with n_dev = ctl + 2 -> b2n_composite_test_fw_dev; the alive list
 * ping-pongs between two buffers.  Every launch has a fixed grid, so a round is CUDA-graph replayable and rounds
 * after the last ray died are no-ops.  b2n_render_schedule applies rendering.py:68-71:
 * N_samples = max(min(n_rays // N_alive, 64), min_samples), stop after max_samples. */
B2N_API int b2n_render_schedule(int32_t *ctl, int64_t n_rays, int min_samples, int max_samples, void *stream);
B2N_API int b2n_raymarching_test_dev(const float *rays_o, const float *rays_d, float *hits_t,
                             const int64_t *alive_indices, const uint8_t *density_bitfield, int cascades,
                             float scale, float exp_step_factor, int grid_size, int max_samples,
                             int64_t max_alive, const int32_t *ctl, float *xyzs, float *dirs, float *deltas,
                             float *ts, int32_t *n_eff, void *stream);
/* as b2n_composite_test_fw; rays that stay alive are appended to alive_next (count in ctl[4]). */
B2N_API int b2n_composite_test_fw_dev(const float *sigmas, const float *rgbs, const float *deltas, const float *ts,
                              int64_t *alive_indices, int64_t *alive_next, float T_threshold,
                              const int32_t *n_eff, int64_t max_alive, int32_t *ctl, float *opacity, float *depth,
                              float *rgb, void *stream);

/* ---------------------------------------------------------------- vren: compositing ------------------ */
/* vren.composite_train_fw (models/custom_functions.py:140-142).  n_total_dev (device i32, may be NULL)
 * is unused by the math; rays_a drives everything.  Outputs indexed by rays_a[:,0]. */
B2N_API int b2n_composite_train_fw(const float *sigmas, const float *rgbs, const float *deltas, const float *ts,
                           const int64_t *rays_a, float T_threshold, int64_t n_rays, float *opacity,
                           float *depth, float *depth_sq, float *rgb, void *stream);
/* vren.composite_train_bw (models/custom_functions.py:153-158).  Writes every sample owned by a ray
 * (zeros after an early stop), so dL_dsigmas/dL_drgbs need no pre-zeroing.  Optional (both or neither):
 * alive_idx (>= total samples) i32 receives the indices of the samples that carry gradient (those composited
 * before the ray's early stop), in no particular order, and alive_count (1) i32 their number. */
B2N_API int b2n_composite_train_bw(const float *dL_dopacity, const float *dL_ddepth, const float *dL_ddepth_sq,
                           const float *dL_drgb, const float *sigmas, const float *rgbs,
                           const float *deltas, const float *ts, const int64_t *rays_a,
                           const float *opacity, const float *depth, const float *depth_sq,
                           const float *rgb, float T_threshold, int64_t n_rays, float *dL_dsigmas,
                           float *dL_drgbs, int32_t *alive_idx, int32_t *alive_count, void *stream);
/* Training fast path: composite_train_fw + NeRFLoss (losses.py:20-40, incl. the random-background blend of
 * models/rendering.py:163-164) + composite_train_bw of each ray in one launch; the results equal the three separate
 * calls.  target (n_rays,3); writes opacity/depth (n_rays), rgb_out (n_rays,3, background-blended; may be NULL),
 * loss_dev (1) fp32 (unscaled loss), dL_dsigmas/dL_drgbs (scaled by loss_scale, or by *loss_scale_dev when that
 * device pointer is non-NULL: b2n_hyper.loss_scale) and the optional alive list. */
B2N_API int b2n_composite_loss_fwbw(const float *sigmas, const float *rgbs, const float *deltas, const float *ts,
                            const int64_t *rays_a, const float *target, float T_threshold, int64_t n_rays,
                            float bg, float lambda_opa, float loss_scale, float *opacity, float *depth,
                            float *rgb_out, float *loss_dev, float *dL_dsigmas, float *dL_drgbs,
                            int32_t *alive_idx, int32_t *alive_count, const float *loss_scale_dev, void *stream);
/* vren.composite_test_fw (models/rendering.py:97-100).  In place on alive_indices/opacity/depth/rgb. */
B2N_API int b2n_composite_test_fw(const float *sigmas, const float *rgbs, const float *deltas, const float *ts,
                          const float *hits_t, int64_t *alive_indices, float T_threshold,
                          const int32_t *n_eff, int n_samples, int64_t n_alive, float *opacity,
                          float *depth, float *rgb, void *stream);

/* ---------------------------------------------------------------- tiny-cuda-nn: encodings ------------ */
#define B2N_MAX_LEVELS 32
typedef struct {
    int32_t n_levels, n_features;        /* n_features must be 2 */
    float scale[B2N_MAX_LEVELS];         /* pos = fma(scale, x, 0.5) */
    uint32_t resolution[B2N_MAX_LEVELS];
    uint32_t size[B2N_MAX_LEVELS];       /* entries in the level (hashed iff res^3 does not fit) */
    uint32_t offset[B2N_MAX_LEVELS + 1]; /* entry offsets into the flat table */
    float x_offset, x_scale;             /* input affine x01 = (x - x_offset) * x_scale, (0,1) by default: folds
                                            NGP.density's box normalisation (networks.py:96) into the kernel */
} b2n_grid_layout;                       /* HOST struct, passed by pointer, copied at launch */

/* GridEncoding sizing of tcnn's "HashGrid" config (models/networks.py:39-47).  Fills *layout (HOST);
 * returns 0.  n_params = layout->offset[n_levels] * n_features. */
B2N_API int b2n_hashgrid_layout(int n_levels, int n_features, int log2_hashmap_size, int base_resolution,
                        double per_level_scale, b2n_grid_layout *layout);
/* HashGrid forward: x (n,3) f32 in [0,1], table fp16 (entries, 2) -> out (n, out_stride) fp16 columns
 * [0, 2*n_levels).  n_dev (device i32, may be NULL) overrides n with min(n, *n_dev). */
B2N_API int b2n_hashgrid_fw(const float *x, const b2n_half *table, const b2n_grid_layout *layout, int64_t n,
                    const int32_t *n_dev, b2n_half *out, int out_stride, void *stream);
/* HashGrid backward w.r.t. the table: dL_dout (n, dy_stride) fp16, accumulated (+=) into
 * grad_table fp32 (entries,2) as  grad += dL_dout * weight * grad_scale.  sample_idx (may be NULL): row i of dL_dout
 * belongs to position x[sample_idx[i]] (compacted backward). */
B2N_API int b2n_hashgrid_bw(const float *x, const b2n_half *dL_dout, int dy_stride, const b2n_grid_layout *layout,
                    int64_t n, const int32_t *n_dev, float grad_scale, float *grad_table, const int32_t *sample_idx,
                    void *stream);
/* Frequency encoding (models/networks.py:49-53): x (n,3) -> out (n, out_stride) fp16, 3*n_freq*2 columns
 * then ones up to the next multiple of 16.  The encoded value is (x - x_min) / x_extent: pass 0, 1 for the plain
 * tcnn module, or xyz_min, xyz_max - xyz_min to fold in the box normalisation of NGP.density (networks.py:96). */
B2N_API int b2n_frequency_fw(const float *x, int n_frequencies, int64_t n, const int32_t *n_dev, b2n_half *out,
                     int out_stride, float x_min, float x_extent, void *stream);
/* SphericalHarmonics degree 4 (models/networks.py:63-70): d01 (n,3) f32 in [0,1] -> out (n,out_stride)
 * fp16 columns [0,16).  If normalize != 0 the input is a raw direction d and the kernel computes
 * d/|d| first (the fused form of networks.py:113-114; (d+1)/2 then *2-1 is folded away exactly). */
B2N_API int b2n_sh4_fw(const float *d, int normalize, int64_t n, const int32_t *n_dev, b2n_half *out,
               int out_stride, void *stream);

/* ---------------------------------------------------------------- tiny-cuda-nn: FullyFusedMLP -------- */
/* FullyFusedMLP, 64 neurons, ReLU, no bias (models/networks.py:54-60,76-82).
 * weights: fp16, layers concatenated, each row-major (out,in): (64,in_width), (64,64)*(n_hidden-1),
 * (16,64).  in (n,in_stride) fp16 [in_width columns used, in_width in {16,32,48,64,80}];
 * hidden (n_hidden, n, 64) fp16 post-ReLU activations saved for backward (may be NULL);
 * out (n,16) fp16.  output_activation: 0 none, 1 sigmoid. */
B2N_API int b2n_mlp_fw(const b2n_half *in, int in_stride, int in_width, const b2n_half *weights, int n_hidden,
               int output_activation, int64_t n, const int32_t *n_dev, b2n_half *hidden, b2n_half *out,
               void *stream);
/* Backward.  dL_dout (n,16) fp16 (w.r.t. the activated output), out = forward output (for sigmoid').
 * dL_din (n,in_stride) fp16 or NULL; grad_weights fp32, same layout as weights, accumulated (+=) as
 * grad += g * grad_scale. */
B2N_API int b2n_mlp_bw(const b2n_half *dL_dout, const b2n_half *in, int in_stride, int in_width,
               const b2n_half *weights, int n_hidden, int output_activation, int64_t n,
               const int32_t *n_dev, const b2n_half *hidden, const b2n_half *out, float grad_scale,
               b2n_half *dL_din, float *grad_weights, void *stream);

/* ---------------------------------------------------------------- fused field MLPs (tcgen05) --------- */
/* The dense half of NGP.forward (models/networks.py:96-115) in one kernel:
 * enc(k1) -> 64 -> 16 = h, sigma = exp(h[0]) (TruncExp), [SH4(dir/|dir|) | h] -> 64 -> 64 -> 3, sigmoid.
 * k1 = 32: the HashGrid configuration (networks.py:39-47, enc = b2n_hashgrid_fw output);
 * k1 = 80: the Frequency-12 configuration this fork has active (networks.py:49-53, enc = b2n_frequency_fw output).
 * image: the two FullyFusedMLP weight sets (sigma: 64*k1 + 1024 halves, rgb: 7168 halves, flat row-major (out,in))
 * repacked by b2n_field_pack_weights into b2n_field_image_halves(k1) halves (scaler_update, may be NULL: a single-GPU
 * trainer's b2n_hyper; the launch then also does what b2n_scaler_update does and clears found_inf).
 * enc (n,k1) fp16, dirs (n,3) fp32 raw directions -> sigmas (n) fp32 (may be NULL), rgbs (n,3) fp32 (fp16-rounded),
 * h (n,16) fp16 (may be NULL; the one activation the backward pass wants saved).  rgbs == NULL selects the
 * density-only form (NGP.density, models/networks.py:87-100): the chain stops after h, dirs is not read. */
B2N_API int b2n_field_image_halves(int k1);
B2N_API int b2n_field_pack_weights(const b2n_half *sigma_weights, const b2n_half *rgb_weights, b2n_half *image, int k1,
                                   b2n_hyper *scaler_update, void *stream);
B2N_API int b2n_field_mlp_fw(const b2n_half *enc, int k1, const float *dirs, const b2n_half *image, int64_t n,
                             const int32_t *n_dev, float *sigmas, float *rgbs, b2n_half *h, void *stream);

/* Whole rays in ONE persistent kernel: the test-time loop of models/rendering.py:42-114 for the 16-level HashGrid field
 * (k1 = 32) without grid-wide rounds, launches or intermediate arrays.  A CTA owns up to 128 ray slots; per round it
 * marches (the reference's serial DDA loop), gathers the hash grid, runs the five layers on tcgen05 and composites, rays take
 * their next samples as soon as their own previous ones are composited and new rays come from a global queue.
 * rays_o / rays_d (n,3), hits_t (n,2) = [t_near, t_far] (read only), layout / table / image as for b2n_hashgrid_fw /
 * b2n_field_mlp_fw.  opacity, depth, rgb are written once per ray: dense arrays (n), (n), (n,3) when out_stride == 0,
 * otherwise columns of one packed block with out_stride floats per ray (e.g. rgb = base, depth = base + 3, opacity =
 * base + 4, out_stride = 5: a rank's part of a sharded frame, ready for the all-gather).  background >= 0: rgb +=
 * background * (1 - opacity) (rendering.py:108-111); < 0: no blend.  tail (4 floats, may be NULL): written by the last
 * CTA: samples marched as three 16-bit digits and the number of rays cut at the budget (ctl[1]) -- the totals travel
 * with the pixels, the host needs no separate read-back.
 * ctl: 8 x i32 on the device, 8-byte aligned, zeroed by the call: [1] rays that reached max_samples while alive (the
 * reference's per-call sample budget would have cut them at a schedule-dependent point: the caller re-renders such a
 * frame with the round loop), [2..3] u64 samples marched, [4] rounds of the longest CTA, [5] rays with at least one
 * sample.  ray_samples (n) i32, may be NULL: samples marched per ray.  first_hit_list: n x 8 bytes of workspace, may be
 * NULL: with it a pre-pass (one thread per ray) walks the empty space in front of every ray, writes the pixels of rays
 * that meet nothing and queues the others at their first sample, so that the persistent kernel never waits on a
 * new ray's chain of dependent occupancy loads.
 * Per-ray arithmetic (positions, encoding, MLP, compositing order) is that of the round loop, except that the
 * transmittance is carried across a ray's rounds (the round loop re-derives it as 1 - opacity at the start of each
 * round): pixels agree with the round loop to fp32 rounding (1e-5), do not depend on which rays share a launch, and
 * are reproducible bit for bit. */
B2N_API int b2n_render_rays(const float *rays_o, const float *rays_d, const float *hits_t, int64_t n_rays,
                            const uint8_t *density_bitfield, int cascades, float scale, float exp_step_factor,
                            int grid_size, int max_samples, const b2n_grid_layout *layout, const b2n_half *table,
                            const b2n_half *image, float T_threshold, float background, float *opacity, float *depth,
                            float *rgb, int out_stride, float *tail, int32_t *ctl, int32_t *ray_samples,
                            void *first_hit_list, void *stream);
/* Backward of the same chain.  dL_dsigmas (n), dL_drgbs (n,3) fp32 (already multiplied by the loss scale); the hidden
 * activations are recomputed from enc / dirs / h (the forward pass saves nothing else).  Writes dL_denc (n,32) fp16
 * for b2n_hashgrid_bw (k1 = 32; NULL to skip, must be NULL for k1 = 80) and accumulates (+=) grad_sigma_w
 * (64*k1 + 1024) / grad_rgb_w (7168) fp32 in the flat row-major (out,in) layout of the weights, times grad_scale.
 * sample_idx (may be NULL): compacted list of sample rows to process (b2n_composite_train_bw's alive_idx); then n /
 * n_dev count list entries and row i of dL_denc belongs to sample sample_idx[i].
 * serialize != 0: MMA issue of the CTA's tile groups goes through a shared-memory lock (validation of the default
 * lock-free accumulation).  found_inf (may be NULL): set to 1 when a gradient left the fp16 range (inf / NaN), the
 * GradScaler signal of the reference's precision=16 training (train.py:265). */
B2N_API int b2n_field_mlp_bw(const float *dL_dsigmas, const float *dL_drgbs, const b2n_half *enc, int k1,
                             const float *dirs, const b2n_half *image, int64_t n, const int32_t *n_dev,
                             const float *rgbs, const b2n_half *h, float grad_scale, b2n_half *dL_denc,
                             float *grad_sigma_w, float *grad_rgb_w, const int32_t *sample_idx, int serialize,
                             int32_t *found_inf, void *stream);

/* ---------------------------------------------------------------- multi-GPU: NVLink peer memory ------ */
/* Data-parallel training (ngp_pl/train.py:197-208: DDPPlugin -> one NCCL gradient all-reduce per step, every rank
 * runs the whole optimiser) re-designed for one NVSwitch node: gradient reduce-scatter + Adam + parameter
 * all-gather as ONE kernel over peer memory, bracketed by flag barriers (csrc/peer.cu).  One process per GPU.
 *
 * b2n_peer_alloc: cudaMalloc `bytes` (zero-filled) on the current device and return its CUDA-IPC handle (64 bytes)
 * for the other ranks; b2n_peer_open maps another rank's handle (peer access enabled lazily); _close / _free undo. */
B2N_API int b2n_peer_alloc(int64_t bytes, void **ptr, void *handle64);
B2N_API int b2n_peer_open(const void *handle64, void **ptr);
B2N_API int b2n_peer_close(void *ptr);
B2N_API int b2n_peer_free(void *ptr);
/* Barrier across the `world` ranks.  flag_ptrs: HOST array of `world` device pointers, entry r = rank r's flag
 * block (>= 16 u32, zero-initialised, peer-mapped).  state: 2 local u32 {completed epoch, sticky error (1 + the
 * rank that timed out)}.  Everything the ranks wrote before the barrier (also into peer memory) is visible to
 * kernels launched after it.  A peer that does not arrive within timeout_s sets the error word; no hang. */
B2N_API int b2n_peer_barrier(void *const *flag_ptrs, int rank, int world, uint32_t *state, double timeout_s,
                             void *stream);
/* The fused step for this rank's shard [shard_first, shard_first + shard_n) of the flat parameter vector:
 * grad = sum_r grad[r][i] (rank order), FusedAdam update (see b2n_adam_step) of param_shard / exp_avg / exp_avg_sq
 * (local, shard_n elements), fp16 result stored to half_ptrs[r][i] for every r.  grad_ptrs / grad16_ptrs / half_ptrs:
 * HOST arrays of `world` device pointers to each rank's full fp32 gradient / fp16 gradient copy / fp16 parameter
 * vector.  Elements with index in [half_lo, half_hi) are read from the fp16 copies made by b2n_grad_pack_half (half
 * the NVLink bytes; grad16_ptrs == NULL or an empty range: everything fp32).  fp32 gradients are left untouched:
 * their holder clears them after the closing barrier.
 * hyper_ptrs (may be NULL): HOST array of `world` device pointers to every rank's b2n_hyper (peer memory): an
 * overflow flagged by ANY rank skips the step on all of them.  barrier_state (may be NULL): the state word pair of
 * b2n_peer_barrier; when its sticky error is set the kernel does nothing (a peer did not arrive: gradients incomplete). */
B2N_API int b2n_adam_step_peer(float *param_shard, float *exp_avg, float *exp_avg_sq, void *const *grad_ptrs,
                               void *const *grad16_ptrs, int64_t half_lo, int64_t half_hi, void *const *half_ptrs,
                               int world, int64_t shard_first, int64_t shard_n, float lr, float beta1, float beta2,
                               float eps, float inv_scale, int step, const b2n_hyper *hyper_dev,
                               void *const *hyper_ptrs, const uint32_t *barrier_state, void *stream);
/* For i in [lo, hi): grad16[i] = saturate_fp16(grad[i]); grad[i] = 0 -- the wire copy of this rank's table gradient. */
B2N_API int b2n_grad_pack_half(float *grad, b2n_half *grad16, int64_t lo, int64_t hi, void *stream);

/* ---------------------------------------------------------------- optimiser / grid maintenance ------- */
/* apex FusedAdam step (train.py:112: lr, eps=1e-15, betas (0.9,0.999), bias-corrected, no weight decay)
 * over one flat fp32 parameter; grad is multiplied by inv_scale and, with zero_grad != 0, then ZEROED (the fused
 * trainer accumulates into it again next step; apex leaves it alone: zero_grad = 0); half_copy (may be NULL)
 * receives the fp16 copy of the updated parameter.  hyper_dev (may be NULL): device b2n_hyper that overrides
 * `lr` / `step` (effective step = step - skipped), divides inv_scale by its loss_scale, and turns the call into
 * "clear the gradient only" while found_inf is set. */
B2N_API int b2n_adam_step(float *param, float *grad, float *exp_avg, float *exp_avg_sq, b2n_half *half_copy,
                  int64_t n, float lr, float beta1, float beta2, float eps, float inv_scale, int step,
                  const b2n_hyper *hyper_dev, int zero_grad, void *stream);
/* GradScaler bookkeeping after the optimiser step: found_inf (OR-ed with the blocks in hyper_ptrs, a HOST array of
 * `world` device pointers, or NULL) halves loss_scale and counts a skipped step; otherwise growth_interval clean
 * steps in a row double it.  found_inf is left for the caller to clear. */
B2N_API int b2n_scaler_update(b2n_hyper *hyper_dev, void *const *hyper_ptrs, int world, void *stream);
/* fp32 -> fp16 parameter copy (what tcnn does before each forward). */
B2N_API int b2n_cast_half(const float *src, b2n_half *dst, int64_t n, void *stream);

/* NGP.update_density_grid pieces (models/networks.py:225-252).
 * cells -> world positions: xyz = (coords/(G-1)*2-1)*(s - s/G) + (noise*2-1)*s/G, normalised to [0,1]
 * with (x - xyz_min)/(xyz_max - xyz_min) when unit_cube != 0 (networks.py:96,227-231). */
B2N_API int b2n_grid_cell_positions(const int32_t *coords, const float *noise, int64_t n, int grid_size, float s,
                            float xyz_min, float xyz_max, int unit_cube, float *xyz, void *stream);
/* grid[idx[i]] update: tmp scatter + EMA-max: grid = grid<0 ? grid : max(grid*decay, tmp)  (:232-237)
 * done in two launches: scatter (tmp[indices[i]] = sigma[i]) and ema over the whole cascade set. */
B2N_API int b2n_grid_scatter(const int64_t *indices, const float *sigmas, int64_t n, float *tmp, int64_t n_cells,
                             void *stream);
B2N_API int b2n_grid_ema(float *density_grid, const float *tmp, int64_t n_cells, float decay, void *stream);
/* mean of grid[grid>0] -> stats_dev[0] = min(mean, density_threshold), stats_dev[1] = mean, stats_dev[2]
 * = count (networks.py:249-251), no host sync.  workspace: 3 doubles (24 B), initialised by the call. */
B2N_API int b2n_grid_threshold(const float *density_grid, int64_t n_cells, float density_threshold,
                       double *workspace, float *stats_dev, void *stream);

/* NGP.mark_invisible_cells (models/networks.py:159-214) in one launch: K (3,3), poses (n_img,3,4) camera-to-world,
 * density_grid (cascades, grid_size^3) in Morton order <- 0 for cells that at least one camera sees at depth >=
 * near_distance and none has in view closer than that, -1 otherwise (every cell is written). */
B2N_API int b2n_mark_invisible_cells(const float *K, const float *poses, int n_img, int img_w, int img_h,
                                     float near_distance, int grid_size, int cascades, float scale,
                                     float *density_grid, void *stream);

/* Random 4-byte gathers over a power-of-two buffer (L2 request-rate probe for bench.py: the roofline of the hash-grid
 * gather).  *n_loads (host, may be NULL) receives the number of loads the launch issues. */
B2N_API int b2n_membench_gather(const void *buf, int64_t bytes, int iters, void *sink, int64_t *n_loads, void *stream);
/* Read-bandwidth probe (bench.py): streams `bytes` of buf `iters` times with ld.global.cg; a buffer that fits the
 * L2 measures L2 bandwidth, a larger one HBM bandwidth.  sink: 4 bytes, never written in practice. */
B2N_API int b2n_membench_read(const void *buf, int64_t bytes, int iters, void *sink, void *stream);

/* ---------------------------------------------------------------- loss (losses.py:26-40) ------------- */
/* NeRFLoss + background blend fused: rgb_out = rgb + bg*(1-opacity) (rendering.py:159-164);
 * loss = mean((rgb_out-target)^2) + lambda_opa*mean(-o log o), o = opacity+1e-10 (train.py:160).
 * Writes loss_dev[0] += loss (caller zeroes) and the gradients dL_drgb (n,3), dL_dopacity (n) of
 * loss*loss_scale with respect to the composited rgb / opacity. */
B2N_API int b2n_nerf_loss_fwbw(const float *rgb, const float *opacity, const float *target, int64_t n_rays,
                       float bg, float lambda_opa, float loss_scale, float *rgb_out, float *loss_dev,
                       float *dL_drgb, float *dL_dopacity, const float *loss_scale_dev, void *stream);

/* shiftscale_inv_depthloss (losses.py:5-23) of a ray batch, forward + backward in one launch without host
 * synchronisation: p = 1/depth (rendered depth, n_rays), g = prior_disp (image-based disparity prior, e.g. LeReS);
 * rays with prior_disp <= 0 or depth <= 1e-6 are left out.  t = torch.median (lower middle), s = mean |x - t|,
 * loss += lambda * mean(((p - t_p)/s_p - (g - t_g)/s_g)^2) is ADDED to *loss_dev, dL_ddepth (n_rays) receives the
 * gradient (autograd-exact through the median and the deviation), times loss_scale (or *loss_scale_dev when
 * non-NULL).  stats (may be NULL): 5 floats {n_valid, t_p, s_p, t_g, s_g}. */
B2N_API int b2n_ssi_depth_loss_fwbw(const float *depth, const float *prior_disp, int64_t n_rays, float lambda,
                                    float loss_scale, const float *loss_scale_dev, float *loss_dev,
                                    float *dL_ddepth, float *stats, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B2N_H */
