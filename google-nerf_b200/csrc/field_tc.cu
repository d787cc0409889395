// field_tc.cu -- the NGP field's dense part on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Replaces the chain
//     xyz_encoder's FullyFusedMLP (K1 -> 64 -> 16)  ->  TruncExp  ->  SH-4(dir)  ->  cat  ->  rgb_net (32 -> 64 -> 64 -> 3)
// of ngp_pl/models/networks.py:96-115 (five tcnn/torch launches plus casts in the reference) with ONE kernel
// forward and ONE kernel backward.  K1 = 32 is the HashGrid configuration (networks.py:39-47), K1 = 80 the
// Frequency-12 configuration this fork has active (networks.py:49-53).
//
// Forward: layer activations never leave the SM: each layer is one tcgen05.mma group (M = 128 samples, N = layer
// width, K = 16 per instruction) whose fp32 accumulator lives in TMEM; the epilogue threads pull their row with
// tcgen05.ld, apply ReLU / exp / sigmoid, and write the fp16 row straight back into the shared-memory operand tile of
// the next layer.  Weights (20-27 KB fp16, laid out once in the UMMA canonical layout by field_pack_weights) arrive
// with one TMA bulk copy per CTA.  The only activation saved for the backward pass is h (16 fp16 per sample).
//
// Backward: the three hidden activations are RECOMPUTED on the tensor cores from enc / [SH | h] (20 kFLOP per sample
// on a pipe that idles) instead of being written by the forward pass and read back (416 B per sample each way).
#include "field_tc.cuh"

__global__ void __launch_bounds__(256) field_pack_weights_kernel(const __half *__restrict__ sigma_w,
                                                                 const __half *__restrict__ rgb_w,
                                                                 __half *__restrict__ image, int k1, b2n_hyper *hyper) {
    // single-GPU trainer: the loss-scaler bookkeeping that follows the optimiser step (b2n_scaler_update) rides along,
    // including the clearing of found_inf -- two launches less per step
    if (hyper != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        if (hyper->found_inf) {
            hyper->skipped += 1; hyper->good_steps = 0;
            hyper->loss_scale = fmaxf(hyper->loss_scale * 0.5f, 1.0f);
        } else if (hyper->growth_interval > 0 && ++hyper->good_steps >= hyper->growth_interval) {
            hyper->good_steps = 0;
            hyper->loss_scale = fminf(hyper->loss_scale * 2.0f, 65536.0f);
        }
        hyper->found_inf = 0;
    }
    // flat row-major (out,in) matrices -> canonical K-major no-swizzle images
    const int w2 = 64 * k1, w3 = w2 + 1024, w4 = w3 + 2048, w5 = w4 + 4096, total = w5 + 1024;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const __half *src; int base, N, K, idx;
        if (e < w2)      { src = sigma_w;        base = 0;  N = 64; K = k1; idx = e; }
        else if (e < w3) { src = sigma_w + w2;   base = w2; N = 16; K = 64; idx = e - w2; }
        else if (e < w4) { src = rgb_w;          base = w3; N = 64; K = 32; idx = e - w3; }
        else if (e < w5) { src = rgb_w + 2048;   base = w4; N = 64; K = 64; idx = e - w4; }
        else             { src = rgb_w + 6144;   base = w5; N = 16; K = 64; idx = e - w5; }
        const int n = idx / K, k = idx - n * K;
        image[base + w_off_halves(n, k, N)] = src[idx];
    }
}

extern "C" int b2n_field_image_halves(int k1) { return (k1 == 32 || k1 == 80) ? 64 * k1 + 8192 : -1; }

extern "C" int b2n_field_pack_weights(const b2n_half *sigma_w, const b2n_half *rgb_w, b2n_half *image, int k1,
                                      b2n_hyper *scaler_update, void *stream) {
    B2N_CHECK_ARG(k1 == 32 || k1 == 80, "first-layer width must be 32 (HashGrid) or 80 (Frequency-12)");
    field_pack_weights_kernel<<<14, 256, 0, (cudaStream_t)stream>>>((const __half *)sigma_w, (const __half *)rgb_w,
                                                                    (__half *)image, k1, scaler_update);
    B2N_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
template <int K1>
__global__ void __launch_bounds__(128, K1 == 32 ? 4 : 3) field_mlp_fw_kernel(
    const __half *__restrict__ enc, const float *__restrict__ dirs, const __half *__restrict__ image, int64_t n,
    const int32_t *__restrict__ n_dev, float *__restrict__ sigmas, float *__restrict__ rgbs,
    __half *__restrict__ h_out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    FieldFwSmem<K1> &S = *reinterpret_cast<FieldFwSmem<K1> *>(smem_raw);
    using I = Img<K1>;
    constexpr int NCH = K1 / 8;
    const int tid = threadIdx.x, warp = tid >> 5;
    n = b2n_eff_n(n, n_dev);
    const int64_t n_tiles = (n + 127) / 128;
    if ((int64_t)blockIdx.x >= n_tiles) return;     // uniform per CTA: nothing allocated yet

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
    const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32 TMEM lanes
    if (tid == 0) {
        mbar_expect_tx(&S.bar_w, I::HALVES * 2);
        bulk_g2s(S.w, image, I::HALVES * 2, &S.bar_w);
    }
    mbar_wait(&S.bar_w, 0);

    const uint32_t w_addr = smem_u32(S.w), a0 = smem_u32(S.a0), a1 = smem_u32(S.a1), a3 = smem_u32(S.a3);
    const bool density_only = (rgbs == nullptr);
    uint32_t phase = 0;

    // the encoded tile of the NEXT tile is fetched with cp.async into a0 as soon as layer 1 has consumed the
    // current one (four layers ahead of its use); the first tile is fetched here
    tile_load_async<NCH>(S.a0, enc + (int64_t)blockIdx.x * 128 * K1, n - (int64_t)blockIdx.x * 128, tid);
    cp_async_commit();
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * 128, row = row0 + tid;
        const bool live = row < n;
        const int64_t next = tile + gridDim.x;
        // ---- stage 0: SH(dir) into the colour-net operand tile; the encoded features are already in flight
        if (!density_only) sh_to_tile(dirs, row, live, S.a3, tid);
        cp_async_wait_all();
        STEP_SYNC();
        // ---- layer 1: enc(K1) -> 64, ReLU
        if (tid == 0) issue_layer(tmem, a0, w_addr + I::W1 * 2, 64, K1, &S.bar_mma);
        mbar_wait(&S.bar_mma, phase); phase ^= 1;
        fence_after_sync();
        if (next < n_tiles) {                         // a0 is free again: next tile's encoded features
            tile_load_async<NCH>(S.a0, enc + next * 128 * K1, n - next * 128, tid);
            cp_async_commit();
        }
        {
            float v[64];
            tmem_ld64(tmem_row, v);
            relu_to_tile(v, S.a1, tid);
        }
        STEP_SYNC();
        // ---- layer 2: 64 -> 16 (h); sigma = exp(h[0])   (TruncExp forward, custom_functions.py:165-167)
        if (tid == 0) issue_layer(tmem, a1, w_addr + I::W2 * 2, 16, 64, &S.bar_mma);
        mbar_wait(&S.bar_mma, phase); phase ^= 1;
        fence_after_sync();
        {
            float v[16];
            tmem_ld16(tmem_row, v);
            const uint4 p0 = pack8(v), p1 = pack8(v + 8);
            *reinterpret_cast<uint4 *>(S.a3 + act_off(tid, 2)) = p0;
            *reinterpret_cast<uint4 *>(S.a3 + act_off(tid, 3)) = p1;
            if (live) {
                const float h0 = __low2float(*reinterpret_cast<const __half2 *>(&p0));   // fp16-rounded like the reference
                if (sigmas != nullptr) sigmas[row] = expf(h0);
                if (h_out != nullptr) {
                    uint4 *dst = reinterpret_cast<uint4 *>(h_out + row * 16);
                    dst[0] = p0; dst[1] = p1;
                }
            }
        }
        if (density_only) {            // NGP.density (networks.py:87-100): only sigma (and h) are wanted
            fence_before_sync();
            __syncthreads();
            fence_after_sync();
            continue;
        }
        STEP_SYNC();
        // ---- layer 3: [SH16 | h16] -> 64, ReLU
        if (tid == 0) issue_layer(tmem, a3, w_addr + I::W3 * 2, 64, 32, &S.bar_mma);
        mbar_wait(&S.bar_mma, phase); phase ^= 1;
        fence_after_sync();
        {
            float v[64];
            tmem_ld64(tmem_row, v);
            relu_to_tile(v, S.a1, tid);
        }
        STEP_SYNC();
        // ---- layer 4: 64 -> 64, ReLU
        if (tid == 0) issue_layer(tmem, a1, w_addr + I::W4 * 2, 64, 64, &S.bar_mma);
        mbar_wait(&S.bar_mma, phase); phase ^= 1;
        fence_after_sync();
        {
            float v[64];
            tmem_ld64(tmem_row, v);
            relu_to_tile(v, S.a1, tid);       // own row only; the layer-4 MMAs that read a1 are complete
        }
        STEP_SYNC();
        // ---- layer 5: 64 -> 3 (padded 16), sigmoid
        if (tid == 0) issue_layer(tmem, a1, w_addr + I::W5 * 2, 16, 64, &S.bar_mma);
        mbar_wait(&S.bar_mma, phase); phase ^= 1;
        fence_after_sync();
        {
            float v[16];
            tmem_ld16(tmem_row, v);
            if (live) {
                #pragma unroll
                for (int c = 0; c < 3; ++c) {
                    // the reference's rgb_net returns fp16: round the sigmoid like it does
                    const float y = 1.0f / (1.0f + __expf(-v[c]));
                    rgbs[3 * row + c] = __half2float(__float2half_rn(y));
                }
            }
        }
        fence_before_sync();
        __syncthreads();       // all TMEM reads of this tile are done before the next tile's first MMA
        fence_after_sync();
    }
    if (warp == 0) tmem_dealloc<64>(tmem);
}

template <int K1>
static int launch_fw(const b2n_half *enc, const float *dirs, const b2n_half *image, int64_t n, const int32_t *n_dev,
                     float *sigmas, float *rgbs, b2n_half *h, void *stream) {
    const int smem = (int)sizeof(FieldFwSmem<K1>) + 256;
    cudaFuncSetAttribute(field_mlp_fw_kernel<K1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    field_mlp_fw_kernel<K1><<<b2n_grid((n + 127) / 128, K1 == 32 ? 4 : 3), 128, smem, (cudaStream_t)stream>>>(
        (const __half *)enc, dirs, (const __half *)image, n, n_dev, sigmas, rgbs, (__half *)h);
    return 0;
}

extern "C" int b2n_field_mlp_fw(const b2n_half *enc, int k1, const float *dirs, const b2n_half *image, int64_t n,
                                const int32_t *n_dev, float *sigmas, float *rgbs, b2n_half *h, void *stream) {
    B2N_CHECK_ARG(k1 == 32 || k1 == 80, "first-layer width must be 32 (HashGrid) or 80 (Frequency-12)");
    B2N_CHECK_ARG(((uintptr_t)enc & 15) == 0 && ((uintptr_t)image & 15) == 0 && ((uintptr_t)h & 15) == 0,
                  "enc / image / h must be 16-byte aligned");
    B2N_CHECK_ARG(rgbs == nullptr || dirs != nullptr, "dirs are required unless only the density is wanted");
    if (n <= 0) return 0;
    if (k1 == 32) launch_fw<32>(enc, dirs, image, n, n_dev, sigmas, rgbs, h, stream);
    else launch_fw<80>(enc, dirs, image, n, n_dev, sigmas, rgbs, h, stream);
    B2N_LAUNCH_CHECK();
    return 0;
}

// ================================================================================================
// Backward.  Per 128-sample tile the hidden activations are recomputed (F-steps: forward layers) and consumed by the
// gradient steps; each gradient step issues, back to back,
//   wgrad:  dW += g^T . act       (M = 64, K = 128 samples; both operands are the SAME shared-memory tiles the
//                                  forward/dgrad MMAs use, read MN-major; accumulators persist in TMEM across
//                                  all tiles of the CTA and are flushed once with red.global.add)
//   dgrad:  g_prev = g . W        (M = 128 samples; W read MN-major from the canonical weight image)
// and the epilogue threads apply ReLU' / TruncExp' and write the fp16 gradient tile of the next step in place.
//
//   tile buffers of a group: X, Y, G (each [128 x 64] fp16 canonical) (+ E = [128 x 80] for K1 = 80)
//   prologue   X[0:32) = [SH(dir) | h]      G = g5 = dL/drgb * sigmoid'(rgb)  (16 wide)
//   F3         Y = relu([SH|h] . W3^T)                                        (= hid_r1)
//   F4         X = relu(Y . W4^T)                                             (= hid_r2)
//   A          dW5 += X^T g5      G = (g5 . W5) * (X > 0)                      (g4)
//   B          dW4 += g4^T Y      G = (g4 . W4) * (Y > 0)                      (g3)
//              meanwhile X[0:32) = [SH|h] again, ENC = enc tile (X[32:64) for K1 = 32, E for K1 = 80)
//   C + F1     dW3 += g3^T [SH|h]     dh = g3 . W3[:, 16:32]     Y = relu(ENC . W1^T)   (= hid_s)
//              G = g2 = dh (+ TruncExp' * dL/dsigma on channel 0)             (16 wide)
//   D          dW2 += Y^T g2      G = (g2 . W2) * (Y > 0)                      (g1)
//   E          dW1 += g1^T ENC    dL/denc = g1 . W1  (HashGrid only)  -> global
//
// Latency hiding: the chain of a tile is 7 dependent MMA groups, and the weight-gradient TMEM columns cap co-resident
// CTAs, so ONE CTA per SM runs GROUPS independent 128-thread groups, each with its own tile, shared-memory tiles,
// mbarrier and 80 accumulator columns, all accumulating into the SAME weight-gradient columns.
// TMEM columns: [80 g, 80 g + 64) D0, [80 g + 64, 80 g + 80) Dc of group g | 80 GROUPS.. dW5^T(16) dW4(64) dW3(32)
// dW2^T(16) dW1(K1).
template <int K1, int GROUPS>
struct FieldBwSmem {
    __half w[Img<K1>::HALVES];
    unsigned char X[GROUPS][TILE64_BYTES];              // 16512 B each
    unsigned char Y[GROUPS][TILE64_BYTES];
    unsigned char G[GROUPS][TILE64_BYTES];
    unsigned char E[K1 == 32 ? 1 : GROUPS][K1 == 32 ? 16 : (K1 / 8) * ACT_LBO];   // K1 = 80: encoded tile [128 x 80]
    int32_t idx[GROUPS][128];                           // sample rows of the group's current tile
    uint64_t bar_w, bar_mma[GROUPS];
    uint32_t tmem_base;
    int lock;
};

// wgrad: D[64 x N] += A_tile^T[64 x 128] * B_tile[128 x N]; A/B tiles are [128 samples x features]
__device__ __forceinline__ void issue_wgrad(uint32_t tmem_d, uint32_t a_tile, uint32_t b_tile, int N) {
    const uint32_t idesc = make_idesc(64, N, 1, 1);
    #pragma unroll
    for (int k = 0; k < 8; ++k) {                    // 16 samples per instruction = two 8-sample groups
        const uint64_t da = make_desc(a_tile + (uint32_t)k * 2 * ACT_SBO, ACT_SBO, ACT_LBO);
        const uint64_t db = make_desc(b_tile + (uint32_t)k * 2 * ACT_SBO, ACT_SBO, ACT_LBO);
        mma_f16_ss(tmem_d, da, db, idesc, 1u);       // accumulators are zero-initialised at kernel start
    }
}
// dgrad: D[128 x N_in] = G_tile[128 x K_out] (K-major) * W[K_out x N_in] (W image has K_out rows: MN-major B);
// in0 = first input feature (multiple of 8) of the N_in-wide slice of W that is used
__device__ __forceinline__ void issue_dgrad(uint32_t tmem_d, uint32_t g_tile, uint32_t w_img, int N_in, int K_out,
                                            int in0 = 0) {
    const uint32_t idesc = make_idesc(128, N_in, 0, 1);
    const uint32_t w_lbo = (uint32_t)(K_out >> 3) * 128;          // byte stride between in-chunks of the image
    w_img += (uint32_t)(in0 >> 3) * w_lbo;
    for (int k = 0; k < K_out / 16; ++k) {
        const uint64_t da = make_desc(g_tile + (uint32_t)k * 2 * ACT_LBO, ACT_LBO, ACT_SBO);
        const uint64_t db = make_desc(w_img + (uint32_t)k * 2 * 128, 128, w_lbo);
        mma_f16_ss(tmem_d, da, db, idesc, k > 0 ? 1u : 0u);
    }
}

__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }

#define GROUP_STEP_SYNC()      \
    do {                       \
        fence_async_smem();    \
        fence_before_sync();   \
        group_sync(grp);       \
        fence_after_sync();    \
    } while (0)

// MMA issue from the group leaders is NOT serialised by default: tcgen05.mma executes in the tensor pipe one
// instruction at a time and D += A*B is a single read-modify-write of TMEM inside that pipe, so accumulations into the
// shared weight-gradient columns from different issuing threads commute (only their order, i.e. fp32 summation order,
// is unspecified).  lock != nullptr (b2n_field_mlp_bw's `serialize` flag) brackets every issue sequence with a
// shared-memory lock; tests/test_gpu_field_tc.py::test_field_tc_backward_stress compares both against fp64.
__device__ __forceinline__ void issue_lock(int *lock) {
    if (lock != nullptr)
        while (atomicCAS(lock, 0, 1) != 0) __nanosleep(32);
    fence_after_sync();
}
__device__ __forceinline__ void issue_unlock(int *lock) {
    fence_before_sync();
    if (lock != nullptr) {
        __threadfence_block();
        atomicExch(lock, 0);
    }
}

// g_next = (act > 0) ? dgrad : 0 for this thread's 64 columns (read from TMEM in two halves to bound registers).
// The mask is applied in the packed fp16 domain: cvt.rn.f16x2 of the gradient pair, HSETP2-style (act > 0) -> {1,0}
// and one HMUL2 -- three instructions per two values instead of convert / compare / select / pack per value.
// found: set when a gradient left the fp16 range (inf/NaN after the conversion).
__device__ __forceinline__ void relu_bw_epilogue(uint32_t tmem_work, const unsigned char *act_tile, unsigned char *g_tile, int r) {
    const __half2 zero2 = __float2half2_rn(0.f);
    #pragma unroll
    for (int half32 = 0; half32 < 2; ++half32) {
        float v[32];
        tmem_ld32(tmem_work + half32 * 32, v);
        #pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int c = half32 * 4 + cc;
            const uint4 a = *reinterpret_cast<const uint4 *>(act_tile + act_off(r, c));
            const __half2 *ah = reinterpret_cast<const __half2 *>(&a);
            uint4 o;
            __half2 *oh = reinterpret_cast<__half2 *>(&o);
            #pragma unroll
            for (int j = 0; j < 4; ++j)
                oh[j] = __hmul2(__floats2half2_rn(v[8 * cc + 2 * j], v[8 * cc + 2 * j + 1]), __hgt2(ah[j], zero2));
            *reinterpret_cast<uint4 *>(g_tile + act_off(r, c)) = o;
        }
    }
}

template <int K1, int GROUPS>
__global__ void __launch_bounds__(128 * GROUPS, 1) field_mlp_bw_kernel(
    const float *__restrict__ dL_dsigmas, const float *__restrict__ dL_drgbs, const __half *__restrict__ enc,
    const float *__restrict__ dirs, const __half *__restrict__ image, int64_t n, const int32_t *__restrict__ n_dev,
    const float *__restrict__ rgbs, const __half *__restrict__ h_in, float grad_scale, __half *__restrict__ dL_denc,
    float *__restrict__ grad_sigma_w, float *__restrict__ grad_rgb_w, const int32_t *__restrict__ sample_idx,
    int serialize, int32_t *__restrict__ found_inf) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    using SM = FieldBwSmem<K1, GROUPS>;
    using I = Img<K1>;
    SM &S = *reinterpret_cast<SM *>(smem_raw);
    constexpr int TM_WG = 80 * GROUPS, TM_DW5T = TM_WG, TM_DW4 = TM_WG + 16, TM_DW3 = TM_WG + 80, TM_DW2T = TM_WG + 112,
                  TM_DW1 = TM_WG + 128, TM_WG_COLS = 128 + K1;
    constexpr int TM_ALLOC = (TM_WG + TM_WG_COLS) <= 256 ? 256 : 512;
    static_assert(TM_WG + TM_WG_COLS <= 512, "TMEM budget");
    const int grp = threadIdx.x >> 7, tid = threadIdx.x & 127, warp = tid >> 5, lane = tid & 31;
    n = b2n_eff_n(n, n_dev);
    const int64_t n_tiles = (n + 127) / 128;
    if ((int64_t)blockIdx.x * GROUPS >= n_tiles) return;          // uniform per CTA

    if (threadIdx.x == 0) {
        mbar_init(&S.bar_w, 1);
        for (int gq = 0; gq < GROUPS; ++gq) mbar_init(&S.bar_mma[gq], 1);
        S.lock = 0;
        mbar_init_fence();
    }
    if (threadIdx.x < 32) tmem_alloc<TM_ALLOC>(&S.tmem_base);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = S.tmem_base;
    const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);      // warp % 4 selects the TMEM lane quadrant
    const uint32_t tmem_work = tmem_row + grp * 80;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&S.bar_w, I::HALVES * 2);
        bulk_g2s(S.w, image, I::HALVES * 2, &S.bar_w);
    }
    if (grp == 0) {                                                      // zero the shared weight-gradient columns
        for (int c16 = 0; c16 < TM_WG_COLS / 16; ++c16) tmem_st16_zero(tmem_row + TM_WG + c16 * 16);
        tmem_st_wait();
    }
    mbar_wait(&S.bar_w, 0);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();

    const uint32_t w_addr = smem_u32(S.w);
    unsigned char *X = S.X[grp], *Y = S.Y[grp], *G = S.G[grp];
    unsigned char *ENC = (K1 == 32) ? (X + 4 * ACT_LBO) : S.E[K1 == 32 ? 0 : grp];
    const uint32_t x_a = smem_u32(X), y_a = smem_u32(Y), g_a = smem_u32(G), enc_a = smem_u32(ENC);
    uint64_t *bar = &S.bar_mma[grp];
    const uint32_t tm_d0 = tmem + grp * 80, tm_dc = tm_d0 + 64;          // accumulators (lane 0 base) of this group
    int *lock = serialize ? &S.lock : nullptr;
    int32_t *IDX = S.idx[grp];
    uint32_t phase = 0;
    bool bad = false;                                                    // a gradient left the fp16 range
    // The groups start together and would march through their (identical) phases in lock step -- all issuing
    // MMAs, then all in the issue-bound epilogue.  A one-off skew per group interleaves them.
#ifndef BW_SKEW_NS
#define BW_SKEW_NS 1000
#endif
    if (grp) __nanosleep(grp * BW_SKEW_NS);

    const int64_t tile_stride = (int64_t)gridDim.x * GROUPS;
    // Per-tile inputs are fetched ONE TILE AHEAD (during step E of the tile before): [SH | h] -> X[0:32) (cp.async for
    // h, registers for SH), the packed g5 = dL/drgb * sigmoid'(rgb) and dL/dsigma in registers -- the L2 latency of
    // these loads is off the tile's dependent chain.
    int64_t nx_s = 0;
    uint32_t nx_g5a = 0u, nx_g5b = 0u;
    float nx_dsig = 0.f;
    bool nx_live = false;
    auto prefetch = [&](int64_t t) {
        const int64_t r = t * 128 + tid;
        nx_live = (t < n_tiles) && (r < n);
        // Row r of a tile is sample IDX[r]: the identity, or an entry of the compacted alive list (sample_idx).
        nx_s = nx_live ? (sample_idx ? (int64_t)__ldg(sample_idx + r) : r) : 0;
        cp_async16(X + act_off(tid, 2), h_in + nx_s * 16, nx_live);
        cp_async16(X + act_off(tid, 3), h_in + nx_s * 16 + 8, nx_live);
        cp_async_commit();
        sh_to_tile(dirs, nx_s, nx_live, X, tid);
        float g5[8];
        #pragma unroll
        for (int i = 0; i < 8; ++i) g5[i] = 0.f;
        nx_dsig = 0.f;
        if (nx_live) {
            #pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float y = __ldg(rgbs + 3 * nx_s + c);
                g5[c] = __ldg(dL_drgbs + 3 * nx_s + c) * y * (1.0f - y);
            }
            nx_dsig = __ldg(dL_dsigmas + nx_s);
        }
        const uint4 p0 = pack8(g5);
        const __half2 *ph = reinterpret_cast<const __half2 *>(&p0);
        bad |= __hisinf(ph[0].x) || __hisnan(ph[0].x) || __hisinf(ph[0].y) || __hisnan(ph[0].y) ||
               __hisinf(ph[1].x) || __hisnan(ph[1].x);
        nx_g5a = p0.x; nx_g5b = p0.y;
    };
    if ((int64_t)blockIdx.x * GROUPS + grp < n_tiles) prefetch((int64_t)blockIdx.x * GROUPS + grp);
    for (int64_t tile = (int64_t)blockIdx.x * GROUPS + grp; tile < n_tiles; tile += tile_stride) {
        const int64_t row0 = tile * 128, rows_valid = n - row0;
        const bool live = nx_live;
        const int64_t my_s = nx_s;
        const float dsig = nx_dsig;
        IDX[tid] = (int32_t)my_s;
        // ---- prologue: [SH | h] is already in X[0:32) (or in flight); g5 -> G[0:16)
        *reinterpret_cast<uint4 *>(G + act_off(tid, 0)) = make_uint4(nx_g5a, nx_g5b, 0u, 0u);
        *reinterpret_cast<uint4 *>(G + act_off(tid, 1)) = make_uint4(0, 0, 0, 0);
        if (K1 != 32) {       // the dedicated encoded tile is fetched here (after IDX is visible)
            group_sync(grp);
            tile_gather_async<K1 / 8>(ENC, enc, IDX, rows_valid, tid);
            cp_async_commit();
        }
        if (K1 == 32) cp_async_wait_all(); else asm volatile("cp.async.wait_group 1;" ::: "memory");   // h has landed
        GROUP_STEP_SYNC();
        // ---- F3: hid_r1 = relu([SH|h] . W3^T) -> Y
        if (tid == 0) {
            issue_lock(lock);
            issue_layer(tm_d0, x_a, w_addr + I::W3 * 2, 64, 32, bar);
            issue_unlock(lock);
        }
        mbar_wait(bar, phase); phase ^= 1;
        fence_after_sync();
        relu_tmem_to_tile(tmem_work, Y, tid);
        GROUP_STEP_SYNC();
        // ---- F4: hid_r2 = relu(hid_r1 . W4^T) -> X
        if (tid == 0) {
            issue_lock(lock);
            issue_layer(tm_d0, y_a, w_addr + I::W4 * 2, 64, 64, bar);
            issue_unlock(lock);
        }
        mbar_wait(bar, phase); phase ^= 1;
        fence_after_sync();
        relu_tmem_to_tile(tmem_work, X, tid);
        GROUP_STEP_SYNC();
        // ---- step A: layer 5 (64 -> 16)
        if (tid == 0) {
            issue_lock(lock);
            issue_wgrad(tmem + TM_DW5T, x_a, g_a, 16);                        // dW5^T[in][out] += hid_r2^T . g5
            issue_dgrad(tm_d0, g_a, w_addr + I::W5 * 2, 64, 16);             // g4 = g5 . W5
            mma_commit(bar);
            issue_unlock(lock);
        }
        mbar_wait(bar, phase); phase ^= 1;
        fence_after_sync();
        relu_bw_epilogue(tmem_work, X, G, tid);                              // g4 = . * (hid_r2 > 0)
        GROUP_STEP_SYNC();
        // ---- step B: layer 4 (64 -> 64)
        if (tid == 0) {
            issue_lock(lock);
            issue_wgrad(tmem + TM_DW4, g_a, y_a, 64);                         // dW4[out][in] += g4^T . hid_r1
            issue_dgrad(tm_d0, g_a, w_addr + I::W4 * 2, 64, 64);              // g3 = g4 . W4
            mma_commit(bar);
            issue_unlock(lock);
        }
        {   // X is free (step A done): [SH | h] again for step C, and (K1 = 32) the encoded tile for F1 / step E
            cp_async16(X + act_off(tid, 2), h_in + my_s * 16, live);
            cp_async16(X + act_off(tid, 3), h_in + my_s * 16 + 8, live);
            if (K1 == 32) tile_gather_async<4>(ENC, enc, IDX, rows_valid, tid);
            cp_async_commit();
            sh_to_tile(dirs, my_s, live, X, tid);
        }
        mbar_wait(bar, phase); phase ^= 1;
        fence_after_sync();
        relu_bw_epilogue(tmem_work, Y, G, tid);                              // g3 = . * (hid_r1 > 0)
        cp_async_wait_all();
        GROUP_STEP_SYNC();
        // ---- step C (layer 3, 32 -> 64) together with F1: hid_s = relu(enc . W1^T)
        if (tid == 0) {
            issue_lock(lock);
            issue_wgrad(tmem + TM_DW3, g_a, x_a, 32);                         // dW3[out][in] += g3^T . [SH|h]
            issue_dgrad(tm_dc, g_a, w_addr + I::W3 * 2, 16, 64, 16);          // dL/dh = g3 . W3[:, 16:32]
            issue_layer_nc(tm_d0, enc_a, w_addr + I::W1 * 2, 64, K1);         // F1
            mma_commit(bar);
            issue_unlock(lock);
        }
        mbar_wait(bar, phase); phase ^= 1;
        fence_after_sync();
        {
            float v[16];
            tmem_ld16(tmem_work + 64, v);                                    // dL/dh from the colour net
            // TruncExp backward (custom_functions.py:171-173) joins on channel 0
            const __half2 hh = *reinterpret_cast<const __half2 *>(X + act_off(tid, 2));
            const float h0 = __low2float(hh);
            if (live) v[0] += dsig * expf(fminf(fmaxf(h0, -15.f), 15.f));
            const uint4 p0 = pack8(v);
            const __half2 *ph = reinterpret_cast<const __half2 *>(&p0);
            bad |= __hisinf(ph[0].x) || __hisnan(ph[0].x);
            *reinterpret_cast<uint4 *>(G + act_off(tid, 0)) = p0;
            *reinterpret_cast<uint4 *>(G + act_off(tid, 1)) = pack8(v + 8);
        }
        relu_tmem_to_tile(tmem_work, Y, tid);                                // hid_s (step B's readers of Y are done)
        GROUP_STEP_SYNC();
        // ---- step D: layer 2 (64 -> 16)
        if (tid == 0) {
            issue_lock(lock);
            issue_wgrad(tmem + TM_DW2T, y_a, g_a, 16);                        // dW2^T[in][out] += hid_s^T . g2
            issue_dgrad(tm_d0, g_a, w_addr + I::W2 * 2, 64, 16);              // g1 = g2 . W2
            mma_commit(bar);
            issue_unlock(lock);
        }
        mbar_wait(bar, phase); phase ^= 1;
        fence_after_sync();
        relu_bw_epilogue(tmem_work, Y, G, tid);                              // g1 = . * (hid_s > 0)
        GROUP_STEP_SYNC();
        // ---- step E: layer 1 (K1 -> 64)
        if (tid == 0) {
            issue_lock(lock);
            issue_wgrad(tmem + TM_DW1, g_a, enc_a, K1);                       // dW1[out][in] += g1^T . enc
            if (dL_denc != nullptr) issue_dgrad(tm_d0, g_a, w_addr + I::W1 * 2, 32, 64);   // g_enc = g1 . W1
            mma_commit(bar);
            issue_unlock(lock);
        }
        prefetch(tile + tile_stride);             // X[0:32) is free since step C: the next tile's [SH | h], g5, dL/dsigma
        mbar_wait(bar, phase); phase ^= 1;
        fence_after_sync();
        if (dL_denc != nullptr) {
            // dL/denc: own row -> G tile (its readers, the step-E MMAs, are done), then one coalesced copy to global
            float v[32];
            tmem_ld32(tmem_work, v);
            #pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint4 p = pack8(v + 8 * c);
                const __half2 *ph = reinterpret_cast<const __half2 *>(&p);
                #pragma unroll
                for (int j = 0; j < 4; ++j) bad |= __hisinf(ph[j].x) || __hisnan(ph[j].x) || __hisinf(ph[j].y) || __hisnan(ph[j].y);
                *reinterpret_cast<uint4 *>(G + act_off(tid, c)) = p;
            }
            fence_before_sync();
            group_sync(grp);
            fence_after_sync();
            tile_store<4>(G, dL_denc + row0 * 32, rows_valid, tid);
        }
        fence_before_sync();
        group_sync(grp);                 // the next tile's prologue overwrites G / X / IDX rows other threads just used
        fence_after_sync();
    }
    // ---- flush the weight gradients: M = 64 accumulators sit in lanes 32*w + (0..15) <-> rows 16*w + lane
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (grp == 0) {
        const int m = warp * 16 + lane;      // valid for lane < 16
        #pragma unroll 1
        for (int c16 = 0; c16 < TM_WG_COLS / 16; ++c16) {
            float v[16];
            tmem_ld16(tmem_row + TM_WG + c16 * 16, v);
            if (lane < 16) {
                #pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int col = TM_WG + c16 * 16 + j;
                    const float val = v[j] * grad_scale;
                    if (val == 0.f) continue;
                    if (!(fabsf(val) <= 3.0e38f)) bad = true;                                               // inf / NaN
                    if (col < TM_DW4)       atomicAdd(grad_rgb_w + 6144 + (col - TM_DW5T) * 64 + m, val);   // dW5^T[in=m][out]
                    else if (col < TM_DW3)  atomicAdd(grad_rgb_w + 2048 + m * 64 + (col - TM_DW4), val);    // dW4[out=m][in]
                    else if (col < TM_DW2T) atomicAdd(grad_rgb_w + m * 32 + (col - TM_DW3), val);           // dW3[out=m][in]
                    else if (col < TM_DW1)  atomicAdd(grad_sigma_w + 64 * K1 + (col - TM_DW2T) * 64 + m, val); // dW2^T[in=m][out]
                    else                    atomicAdd(grad_sigma_w + m * K1 + (col - TM_DW1), val);         // dW1[out=m][in]
                }
            }
        }
    }
    if (found_inf != nullptr && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(found_inf, 1);
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<TM_ALLOC>(tmem);
}

template <int K1, int GROUPS>
static void launch_bw(const float *dL_dsigmas, const float *dL_drgbs, const b2n_half *enc, const float *dirs,
                      const b2n_half *image, int64_t n, const int32_t *n_dev, const float *rgbs, const b2n_half *h,
                      float grad_scale, b2n_half *dL_denc, float *grad_sigma_w, float *grad_rgb_w,
                      const int32_t *sample_idx, int serialize, int32_t *found_inf, void *stream) {
    const int smem = (int)sizeof(FieldBwSmem<K1, GROUPS>) + 256;
    cudaFuncSetAttribute(field_mlp_bw_kernel<K1, GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int64_t n_tiles = (n + 127) / 128;
    field_mlp_bw_kernel<K1, GROUPS><<<b2n_grid((n_tiles + GROUPS - 1) / GROUPS, 1), 128 * GROUPS, smem, (cudaStream_t)stream>>>(
        dL_dsigmas, dL_drgbs, (const __half *)enc, dirs, (const __half *)image, n, n_dev, rgbs, (const __half *)h,
        grad_scale, (__half *)dL_denc, grad_sigma_w, grad_rgb_w, sample_idx, serialize, found_inf);
}

extern "C" int b2n_field_mlp_bw(const float *dL_dsigmas, const float *dL_drgbs, const b2n_half *enc, int k1,
                                const float *dirs, const b2n_half *image, int64_t n, const int32_t *n_dev,
                                const float *rgbs, const b2n_half *h, float grad_scale, b2n_half *dL_denc,
                                float *grad_sigma_w, float *grad_rgb_w, const int32_t *sample_idx, int serialize,
                                int32_t *found_inf, void *stream) {
    B2N_CHECK_ARG(k1 == 32 || k1 == 80, "first-layer width must be 32 (HashGrid) or 80 (Frequency-12)");
    B2N_CHECK_ARG(dL_dsigmas && dL_drgbs && enc && dirs && image && rgbs && h && grad_sigma_w && grad_rgb_w,
                  "gradients in, enc, dirs, weights, rgbs, h and the weight-gradient outputs are required");
    B2N_CHECK_ARG(k1 == 32 || dL_denc == nullptr, "dL/denc exists for the HashGrid configuration only");
    B2N_CHECK_ARG(((uintptr_t)enc & 15) == 0 && ((uintptr_t)h & 15) == 0 && ((uintptr_t)dL_denc & 15) == 0,
                  "enc / h / dL_denc must be 16-byte aligned");
    if (n <= 0) return 0;
    if (k1 == 32)
        launch_bw<32, 4>(dL_dsigmas, dL_drgbs, enc, dirs, image, n, n_dev, rgbs, h, grad_scale, dL_denc, grad_sigma_w,
                         grad_rgb_w, sample_idx, serialize, found_inf, stream);
    else
        launch_bw<80, 2>(dL_dsigmas, dL_drgbs, enc, dirs, image, n, n_dev, rgbs, h, grad_scale, dL_denc, grad_sigma_w,
                         grad_rgb_w, sample_idx, serialize, found_inf, stream);
    B2N_LAUNCH_CHECK();
    return 0;
}
